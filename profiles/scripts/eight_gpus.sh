set -x
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 8 --steps 2 --warmup 3 --iters-per-step 40 --no-cpu > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 runs rc=$?"
timeout 300 $T bench.py --gpus 8 --steps 2 --warmup 3 --iters-per-step 20 --no-cpu --no-e2e --workload netflix --shard ratings > gpurun_out/bench_n8_netflix_ratings.json 2> gpurun_out/bench_n8_netflix_ratings.err; echo "n8 ratings rc=$?"
timeout 300 $T bench.py --gpus 8 --steps 2 --warmup 3 --iters-per-step 20 --no-cpu --no-e2e --shard ratings > gpurun_out/bench_n8_ml20m_ratings.json 2> gpurun_out/bench_n8_ml20m_ratings.err; echo "n8 ratings ml20m rc=$?"
