"""Summarise `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.
usage: python profiles/hotlines.py <csv> [kernel-index] [top-n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
secs, cur = [], None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = {"file": r[1], "rows": []}; secs.append(cur)
    elif r[0] == "Function Name" and cur is not None: cur["func"] = r[1]
    elif r[0] == "Line No" and cur is not None: cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r[0] != "": cur["rows"].append(r)
funcs = []
for s in secs:
    if s.get("func") not in funcs: funcs.append(s.get("func"))
print("kernels:", funcs)
f = funcs[want]
tot = totS = 0
allr = []
for s in secs:
    if s.get("func") != f: continue
    h = s["hdr"]; iS = h.index("# Samples"); iI = h.index("Instructions Executed")
    for r in s["rows"]:
        try: n = int(r[iI] or 0); m = int(r[iS] or 0)
        except ValueError: continue
        tot += n; totS += m
        allr.append((s["file"].split("/")[-1], int(r[0]), n, m, r[1]))
print(f, "total warp-inst", tot, "samples", totS)
top = sorted(allr, key=lambda x: -x[2])[:topn]
for fl, ln, n, m, src in sorted(top, key=lambda x: (x[0], x[1])):
    print(f"{fl:>14}:{ln:<4} inst={n/tot*100:5.1f}% smp={m/max(totS,1)*100:5.1f}%  {src.strip()[:100]}")
