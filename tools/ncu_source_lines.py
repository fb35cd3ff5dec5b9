"""Per source line of an ncu report's source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv): stall
samples, warp instructions, threads active per instruction. usage: ncu_source_lines.py f.csv [file substring] [top] [--regions a-b:name,...]"""
import csv, sys, collections
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
regions = []
for a in sys.argv[4:]:
    if a.startswith("--regions="):
        for item in a[len("--regions="):].split(","):
            rng, name = item.split(":")
            lo, hi = rng.split("-")
            regions.append((int(lo), int(hi), name))
rows = csv.reader(open(path, newline=""))
cur = None; hdr = None; sec = -1; named = None
per = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Name": named = r[1]; continue
    if r[0] == "Line No":  # a new file section (the combined cuda,sass view does not name them: number them)
        hdr = {h: i for i, h in enumerate(r)}; hdr_list = r; sec += 1; cur = named or f"section{sec}"; named = None; continue
    if r[0] == "" or hdr is None or len(r) < 8: continue
    if want not in (cur or ""): continue
    try: line = int(r[0])
    except ValueError: continue
    i_s = hdr_list.index("# Samples"); i_w = hdr_list.index("Instructions Executed"); i_t = hdr_list.index("Thread Instructions Executed")
    e = per[(cur, line)]
    num = lambda x: int(x) if x not in ("", "-") else 0
    e[0] += num(r[i_s]); e[1] += num(r[i_w]); e[2] += num(r[i_t]); e[3] = r[1]
tot_s = sum(e[0] for e in per.values()); tot_w = sum(e[1] for e in per.values()); tot_t = sum(e[2] for e in per.values())
print(f"total samples {tot_s}, warp instructions {tot_w:.3e}, threads per instruction {tot_t / max(tot_w, 1):.2f}")
if regions:
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (f, line), e in per.items():
        name = next((n for lo, hi, n in regions if lo <= line <= hi), "other")
        a = agg[name]; a[0] += e[0]; a[1] += e[1]; a[2] += e[2]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * a[0] / tot_s:5.1f}% samples  {100 * a[1] / tot_w:5.1f}% warp-inst  {a[2] / max(a[1], 1):5.1f} threads/inst  {name}")
else:
    for (f, line), e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * e[0] / tot_s:5.1f}%  {100 * e[1] / tot_w:5.1f}%w  {e[2] / max(e[1], 1):5.1f}thr  {(f or '?').split('/')[-1]}:{line}  {e[3].strip()[:100]}")
