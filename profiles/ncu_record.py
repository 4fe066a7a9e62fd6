"""Per-iteration record of an `ncu --csv --log-file X --metrics ...` pass over bench.py, written into
profiles/ncu_records.json (what bench.py's roofline block reads).

usage: python profiles/ncu_record.py <csv> <workload> <source-name>

The launches are cut into EM iterations at every prep_p_kernel; the record is the median complete
iteration: DRAM bytes (read + write) of ALL its kernels, of the segment_pass_kernel launches alone,
L1 global-load bytes (sectors x 32), and time-weighted fp64-pipe / L2-hit percentages of the
segment passes."""
import collections, csv, json, os, sys

M_T = "gpu__time_duration.sum"
M_R, M_W = "dram__bytes_read.sum", "dram__bytes_write.sum"
M_L1 = "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"
M_F64 = "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"
M_L2 = "lts__t_sector_hit_rate.pct"
M_LSU = "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
        "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}          # times in ms, sizes in bytes


def extract(path, source=None):
    """The per-iteration record of one metrics CSV (see the module docstring)."""
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0]})
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        d[r["Metric Name"]] = v
    its, cur = [], None
    for d in launches.values():
        if "prep_p_kernel" in d["name"]:
            cur = []
            its.append(cur)
        if cur is not None:
            cur.append(d)
    its = [it for it in its if any("pr_finalize" in d["name"] for d in it)]
    # an iteration ends with pr_finalize_kernel: drop what follows it (likelihood, next phase)
    cut = []
    for it in its:
        end = max(i for i, d in enumerate(it) if "pr_finalize" in d["name"])
        cut.append(it[:end + 1])
    cut.sort(key=lambda it: sum(d.get(M_T, 0.0) for d in it))
    it = cut[len(cut) // 2]
    seg = [d for d in it if "segment_pass_kernel" in d["name"]]
    dram = lambda ds: sum(d.get(M_R, 0.0) + d.get(M_W, 0.0) for d in ds)
    tw = lambda ds, m: (sum(d.get(m, 0.0) * d.get(M_T, 0.0) for d in ds) / max(sum(d.get(M_T, 0.0) for d in ds), 1e-30))
    per_kernel = collections.OrderedDict()
    for d in it:
        k = per_kernel.setdefault(d["name"].replace("void ", "").replace("mmsbm::", ""), {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
        k["launches"] += 1; k["ms"] += d.get(M_T, 0.0); k["dram_bytes"] += d.get(M_R, 0.0) + d.get(M_W, 0.0)
    rec = {"source": source or path, "iterations_seen": len(cut), "launches_per_iteration": len(it),
           "dram_bytes_per_iteration": dram(it), "segment_pass_dram_bytes": dram(seg),
           "l1_global_load_bytes_per_iteration": sum(d.get(M_L1, 0.0) for d in it) * 32.0,
           "ncu_ms_per_iteration": sum(d.get(M_T, 0.0) for d in it),
           "segment_pass_share_of_ncu_time": sum(d.get(M_T, 0.0) for d in seg) / sum(d.get(M_T, 0.0) for d in it),
           "fp64_pipe_pct": tw(seg, M_F64), "l2_hit_pct": tw(seg, M_L2), "lsu_pipe_pct": tw(seg, M_LSU) or None,
           "per_kernel": per_kernel}
    return rec


def main(path, workload, source):
    rec = extract(path, source)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_records.json")
    allrec = json.load(open(out)) if os.path.exists(out) else {}
    allrec[workload] = rec
    json.dump(allrec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
