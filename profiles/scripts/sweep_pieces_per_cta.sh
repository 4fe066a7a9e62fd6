for wl in netflix ml20m ml1m; do for spc in 0 8 16 32; do
  MMSBM_SPC=$spc python bench.py --workload $wl --steps 2 --warmup 3 --iters-per-step 30 --no-e2e --no-cpu > gpurun_out/sw.json 2> gpurun_out/sw.err || tail -3 gpurun_out/sw.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/sw.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms"]
print("$wl SPC=$spc iter %.4f by_user %.4f by_item %.4f" % (d["ms_per_iteration"], k["by_user"], k["by_item"]))
PY
done; done
python profiles/scripts/shard_profile.py netflix 8
MMSBM_SPC=16 python profiles/scripts/shard_profile.py netflix 8
MMSBM_SPC=8 python profiles/scripts/shard_profile.py netflix 8
MMSBM_SPC=8 python profiles/scripts/shard_profile.py ml20m 8
