"""Static code size of a kernel by source statement: for every SASS instruction take the inline chain printed by
`nvdisasm --print-line-info-inline` and credit the instruction to the frame that was inlined at <call_line> of <file>
(i.e. the statement of the callee body), or to the kernel-body line when there is none.
usage: sass_static.py <dis.txt> <kernel substring> <file basename> <call line> [min count]"""
import collections, re, sys
dis, kern, fname, call = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
minc = int(sys.argv[5]) if len(sys.argv) > 5 else 40
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if '.section' in l and '.text.' in l and kern in l)
cnt = collections.Counter(); body = collections.Counter()
block, inblock, n = [], False, 0
for l in lines[start + 1:]:
    if l.startswith('\t.section') and n:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if not inblock:
            block, inblock = [], True
        block.append((m.group(1).split('/')[-1], int(m.group(2)), (m.group(3) or '').split('/')[-1], int(m.group(4) or 0)))
        continue
    if re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l):
        inblock = False
        n += 1
        key = None
        for f, a, g, b in block:
            if g == fname and b == call and f == fname:
                key = a
        if key is not None:
            cnt[key] += 1
        else:
            last = block[-1] if block else ('?', 0, '', 0)
            body[(last[0], last[1])] += 1
import os
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bcm3_b200', 'csrc', fname)).read().split('\n')
print('instructions', n, 'in callee', sum(cnt.values()))
for ln, v in sorted(cnt.items()):
    if v >= minc:
        print(f'{ln:5d} {v:6d}  {src[ln - 1].strip()[:110]}')
print('--- kernel body / other')
for k, v in body.most_common(15):
    print(k, v)
