for v in "3 3" "4 3" "4 2" "5 2" "6 2" "5 3" "3 4" "2 4" "2 3"; do set -- $v
  MMSBM_UN=$1 MMSBM_OCC=$2 python bench.py --steps 2 --warmup 3 --iters-per-step 30 --no-e2e --no-cpu > gpurun_out/sw.json 2> gpurun_out/sw.err || tail -3 gpurun_out/sw.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/sw.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms"]
print("UN=$1 OCC=$2 iter %.3f by_user %.3f by_item %.3f" % (d["ms_per_iteration"], k["by_user"], k["by_item"]))
PY
done
