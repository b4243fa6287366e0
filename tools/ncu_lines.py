"""Map the stall samples of an `ncu --page source --csv` dump (SASS level) to source lines with nvdisasm's line info.

    ncu -i X.ncu-rep --page source --csv > src.csv ;  cuobjdump -xelf all build/K.o ; nvdisasm -g -c K.cubin > dis.txt
    python tools/ncu_lines.py src.csv dis.txt <kernel-name-substring> [file-substring]

Prints, per source line of the kernel's own file (inlined callee lines are attributed to the innermost nvdisasm line tag), the samples
and the two dominant stall reasons, sorted by samples."""
import csv, re, sys
from collections import defaultdict

src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
kname, dname = kname.split(':') if ':' in kname else (kname, kname)   # 'ncu (demangled) substring:nvdisasm (mangled) substring'
only = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(src_csv)))
# the dump may hold several kernels: blocks start with a "Kernel Name" row
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
blk = [b for b in blocks if kname in b['name']][0]
hdr, data = blk['rows'][0], blk['rows'][1:]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
# nvdisasm: instruction lines in order, tagged with the last "//## File ... line N" (+ "inlined at" chains collapse to the innermost tag)
lines, tag, in_fn = [], None, False
for l in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        tag = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*\.text\.', l) or '.section' in l:
        in_fn = dname in l if '.text.' in l else in_fn
    if in_fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', l):
        lines.append(tag)
n = min(len(lines), len(data))
if len(lines) != len(data):
    print(f'warning: {len(lines)} disassembled instructions vs {len(data)} profiled', file=sys.stderr)
agg = defaultdict(lambda: [0, defaultdict(int)])
for i in range(n):
    s = int(data[i][idx['# Samples']] or 0)
    if not s:
        continue
    a = agg[lines[i]]
    a[0] += s
    for h in stall_cols:
        a[1][h] += int(data[i][idx[h]] or 0)
tot = sum(a[0] for a in agg.values())
print('total samples', tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    if only and k and only not in k[0]:
        continue
    top = sorted(a[1].items(), key=lambda kv: -kv[1])[:2]
    print(f'{a[0]:6d} {100.0 * a[0] / tot:5.1f}%  {k}  {[(h[6:], v) for h, v in top]}')
