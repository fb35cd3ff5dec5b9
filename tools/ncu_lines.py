"""Join an ncu SASS source-page CSV with nvdisasm line info: instructions / stall samples per CUDA source line.
usage: ncu_lines.py <src.csv from `ncu --page source --csv`> <nvdisasm --print-line-info output> <kernel substring> [top]"""
import collections, csv, re, sys, os
csv_path, dis_path, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis_path).read().split('\n')
start = next(i for i, l in enumerate(lines) if '.section' in l and '.text.' in l and kern in l)
addr2line, cur, fresh = {}, None, True
for l in lines[start + 1:]:
    if l.startswith('\t.section') and addr2line:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:  # with --print-line-info-inline the first line of a group is the innermost frame
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            fresh = False
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
        fresh = True
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][ix["Address"]], 16)
agg, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
tot = ts = 0
for r in rows[2:]:
    a = int(r[ix["Address"]], 16) - base
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    key = addr2line.get(a)
    agg[key] += n; samp[key] += s; tot += n; ts += s
    thr[key] += int(r[ix["Thread Instructions Executed"]] or 0)
src = {}
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bcm3_b200", "csrc")
for f in os.listdir(root):
    if not os.path.isfile(os.path.join(root, f)):
        continue
    src[f] = open(os.path.join(root, f), errors="replace").read().split('\n')
print(f"total warp instructions {tot}, samples {ts}, sass lines {len(rows) - 2}")
for k, v in agg.most_common(top):
    if not k:
        print(f"{'?':18s} inst {v / tot * 100:5.2f}%")
        continue
    f, ln = k
    text = src[f][ln - 1].strip()[:90] if f in src and ln - 1 < len(src[f]) else ''
    print(f"{f[:13]:13s}:{ln:4d} inst {v / tot * 100:5.2f}% samp {samp[k] / ts * 100:5.2f}% thr {thr[k] / max(v, 1):4.1f} | {text}")
