"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) of `bench.py`: kernels of the second
applyChanges call (the first one after warm-up), grouped by kernel, sorted by time. python tools/launch_summary.py file.csv"""
import csv, re, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
r = csv.reader(lines); hdr = next(r)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
rows = []
for row in r:
    v = float(row[iv].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(row[iu], 1)
    rows.append((row[ik], v))
def short(n):
    m = re.search(r'k_foreach(?:_warp)?<(?:amg::)?(\w+)', n)
    if m: return m.group(1)
    m = re.search(r'(k_\w+)', n)
    return m.group(1) if m else n[:40]
names = [short(n) for n, _ in rows]
starts = [i for i, n in enumerate(names) if n in ('OffsetsToRangesKernel', 'SplitPairsKernel') and (i == 0 or names[i - 1] not in ('OffsetsToRangesKernel',))]
starts = [i for k, i in enumerate(starts) if k == 0 or i - starts[k - 1] > 20]   # one per call
a = starts[1] if len(starts) > 1 else starts[0]
b = min(len(rows), a + (starts[1] - starts[0])) if len(starts) > 1 else len(rows)   # same number of launches as the call before it (what follows in bench.py is the decode re-run)
seg = list(zip(names[a:b], [v for _, v in rows[a:b]]))
tot = sum(v for _, v in seg)
print('%d launches in the list; call = launches [%d, %d): %d launches, %.3f ms of kernel time' % (len(rows), a, b, len(seg), tot / 1e6))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in seg:
    agg[n][0] += 1; agg[n][1] += v
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print('%-30s %4d %8.3f ms %5.1f%%' % (n, c, v / 1e6, 100 * v / tot))
