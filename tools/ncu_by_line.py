"""Join an `ncu --page source --csv` SASS listing with `nvdisasm --print-line-info` to attribute executed
instructions / stall samples to CUDA source lines.  usage: ncu_by_line.py src.csv kernel.sass [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('Source'); iI = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
cnt = []
for r in data:
    try: cnt.append((int(r[iI]), int(r[iN]), r[iS]))
    except Exception: pass
lines = []
cur = ('?', 0)
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
print('sass', len(lines), 'csv', len(cnt))
n = min(len(lines), len(cnt))
by = collections.Counter(); bys = collections.Counter()
tot = sum(c[0] for c in cnt); ts = sum(c[1] for c in cnt)
for i in range(n):
    by[lines[i]] += cnt[i][0]; bys[lines[i]] += cnt[i][1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
for k, v in by.most_common(top):
    print(f'{k[0]:22s} L{k[1]:4d} {v:11d} {100*v/tot:5.1f}%  samples {100*bys[k]/ts:5.1f}%')
