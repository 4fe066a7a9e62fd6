#!/bin/bash
# Round 2, two B200s (gpurun --gpus 2): the multi-GPU test, then the strong-scaling bench line.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2_test_multi_gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps ${STEPS:-3} --warmup ${WARMUP:-2} --iters-per-step ${ITERS:-100} \
  > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 3000 gpurun_out/r2_bench_n2.err
cat gpurun_out/r2_bench_n2.json
