#!/bin/bash
# Round 2 records on one B200: GPU test-suite, bench lines, ncu metric pass (all kernels of the
# iterations) and the --set full capture of the segment passes.  TAG names the kernel version.
TAG=${TAG:-r2_v14}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
python bench.py > gpurun_out/${TAG}_bench_ml20m.json 2> gpurun_out/${TAG}_bench_ml20m.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
for w in ml1m:500 netflix:100 ml100k:200; do
  python bench.py --workload ${w%%:*} --iters-per-step ${w##*:} --no-cpu > gpurun_out/${TAG}_bench_${w%%:*}.json 2> gpurun_out/${TAG}_bench_${w%%:*}.err
done
python bench.py --ids zipf --iters-per-step 100 --no-cpu --no-api-e2e --no-e2e > gpurun_out/${TAG}_bench_ml20m_zipf.json 2> gpurun_out/${TAG}_bench_zipf.err
if [ -z "$SKIP_NCU" ]; then
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
  for w in ml20m ml1m netflix ml100k; do
    CMD="python bench.py --workload $w --steps 1 --warmup 1 --iters-per-step 2 --no-e2e --no-cpu --no-api-e2e"
    $CMD > gpurun_out/plain_$w.log 2>&1 &&
    ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_metrics_$w.csv $CMD > gpurun_out/ncu_metrics_$w.log 2>&1
  done
  CMD="python bench.py --steps 1 --warmup 1 --iters-per-step 2 --no-e2e --no-cpu --no-api-e2e"
  ncu --set full --clock-control none --import-source on -k regex:segment_pass --launch-skip 3 -c 3 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/ncu_full.log 2>&1
fi
echo done
