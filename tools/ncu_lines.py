"""Per-source-line instruction counts and stall samples: joins an `ncu --page source --csv` SASS dump with the line table
of `nvdisasm -g -c` output of the same cubin (instructions matched by order within the function)."""
import csv, re, sys, collections
sass_csv, dis, func_pat = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 5 and r[2].isdigit()]
half = len(data)
# several captured launches of the same kernel repeat the listing: keep the first
addr0 = data[0][0]
for i in range(1, len(data)):
    if data[i][0] == addr0:
        half = i
        break
data = data[:half]
iS, iE = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
lines = []
cur = None
infunc = False
for l in open(dis):
    if l.startswith('.text.'):
        infunc = func_pat in l
        continue
    if not infunc:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        lines.append(cur)
print('instructions: ncu', len(data), 'nvdisasm', len(lines))
agg = collections.defaultdict(lambda: [0, 0])
tot = 0
for r, ln in zip(data, lines):
    e = int(r[iE]); s = int(r[iS])
    agg[ln][0] += e; agg[ln][1] += s; tot += e
stot = sum(v[1] for v in agg.values())
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    f, n = k if k else ('?', 0)
    if f not in src:
        try:
            src[f] = open('/root/repo/audio-tokens_b200/csrc/' + f).read().split('\n')
        except Exception:
            src[f] = []
    text = src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ''
    print(f'{v[0] / tot * 100:5.1f}% inst {v[1] / max(stot, 1) * 100:5.1f}% stall  {f}:{n}: {text}')
