"""Per source line: share of one stall reason's samples. usage: ncu_stalls.py <src.csv> <nvdisasm --print-line-info-inline> <kernel substring> <stall column> [top]"""
import collections, csv, re, sys, os
csv_path, dis_path, kern, col = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
lines = open(dis_path).read().split('\n')
start = next(i for i, l in enumerate(lines) if '.section' in l and '.text.' in l and kern in l)
addr2line, cur, fresh = {}, None, True
for l in lines[start + 1:]:
    if l.startswith('\t.section') and addr2line:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            fresh = False
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
        fresh = True
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][ix["Address"]], 16)
agg = collections.Counter(); ops = collections.defaultdict(collections.Counter)
tot = 0
for r in rows[2:]:
    a = int(r[ix["Address"]], 16) - base
    s = int(r[ix[col]] or 0)
    key, sass = addr2line.get(a, (None, ''))
    agg[key] += s; tot += s
    ops[key][sass.split()[0] if sass else '?'] += s
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bcm3_b200", "csrc")
src = {f: open(os.path.join(root, f), errors="replace").read().split('\n') for f in os.listdir(root) if os.path.isfile(os.path.join(root, f))}
print(f"{col}: {tot} samples")
for key, s in agg.most_common(top):
    text = ''
    if key and key[0] in src and key[1] - 1 < len(src[key[0]]):
        text = src[key[0]][key[1] - 1].strip()[:90]
    print(f"{100 * s / max(tot, 1):5.1f}%  {key}  {dict(ops[key].most_common(3))}  | {text}")
