#!/bin/bash
# Round 2, N B200s (gpurun --gpus N): multi-GPU test, the strong-scaling bench line of the ML-20M
# configuration (+ the Netflix sharded record), the ML-1M runs-sharded line and the cv_fit line.
N=${N:-2}
TAG=${TAG:-r2_v14}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_TESTS" ]; then
  python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_test_multi_gpu_n$N.txt
fi
$TR --master-port 29511 bench.py --gpus $N --steps ${STEPS:-5} --warmup ${WARMUP:-2} \
  > gpurun_out/${TAG}_bench_ml20m_n$N.json 2> gpurun_out/${TAG}_bench_ml20m_n$N.err; echo "ml20m rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_ml20m_n$N.err
$TR --master-port 29512 bench.py --gpus $N --workload ml1m --iters-per-step 500 --steps ${STEPS:-5} --warmup ${WARMUP:-2} --no-netflix \
  > gpurun_out/${TAG}_bench_ml1m_n$N.json 2> gpurun_out/${TAG}_bench_ml1m_n$N.err; echo "ml1m rc=$?"
$TR --master-port 29513 bench.py --gpus $N --workload ml1m_cv --steps 2 --warmup 1 \
  > gpurun_out/${TAG}_bench_ml1m_cv_n$N.json 2> gpurun_out/${TAG}_bench_ml1m_cv_n$N.err; echo "cv rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_ml1m_cv_n$N.err
if [ -n "$ZIPF" ]; then
  $TR --master-port 29514 bench.py --gpus $N --ids zipf --iters-per-step 100 --steps 3 --warmup 1 --no-netflix --no-e2e \
    > gpurun_out/${TAG}_bench_ml20m_zipf_n$N.json 2> gpurun_out/${TAG}_bench_ml20m_zipf_n$N.err; echo "zipf rc=$?"
fi
for f in gpurun_out/${TAG}_bench_*_n$N.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], "unreadable", e); raise SystemExit
print(sys.argv[1], "value %.3e" % d["value"], "ms/step %.2f" % d["ms_per_step"], d["config"].get("mode"), d["config"].get("parallelism", "")[:60])
for m, r in (d.get("modes") or {}).items():
    print("   ", m, "%.3e" % r["value"], "ms/it %.4f" % r["ms_per_iteration"], {k: r[k] for k in ("pr_wait_ms_per_iteration", "runs_per_gpu", "setup_s", "stage_ms_when_profiled") if k in r})
nf = d.get("netflix_ratings_sharded")
if nf: print("    netflix %.3e" % nf["value"], "ms/it %.4f" % nf["ms_per_iteration"], "wait", nf["pr_wait_ms_per_iteration"], nf["stage_ms_when_profiled"], "lik diff", nf.get("likelihood_rel_diff_vs_one_gpu"))
if d.get("e2e"): print("    e2e %.3e" % d["e2e"]["value"], "ms/step %.1f" % d["e2e"]["ms_per_step"])
if "cv_accuracies" in d: print("    cv", d["cv_accuracies"], d["wall_s_per_cv"], d["config"]["parallelism"])
PY
done
