python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python bench.py --workload ml1m --iters-per-step 500 --no-cpu > gpurun_out/bench_ml1m.json 2> gpurun_out/bench_ml1m.err
python bench.py --workload netflix --iters-per-step 100 --no-cpu > gpurun_out/bench_netflix.json 2> gpurun_out/bench_netflix.err
python bench.py --workload ml100k --iters-per-step 200 --no-cpu > gpurun_out/bench_ml100k.json 2> gpurun_out/bench_ml100k.err
python bench.py --ids zipf --iters-per-step 100 --no-cpu > gpurun_out/bench_zipf.json 2> gpurun_out/bench_zipf.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v13.csv python bench.py --steps 1 --warmup 1 --iters-per-step 2 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:segment_pass --launch-skip 3 -c 3 -f -o gpurun_out/prof_r1_v13 python bench.py --steps 1 --warmup 1 --iters-per-step 2 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
echo done
