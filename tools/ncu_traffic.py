"""Reads an ncu report (`ncu --set full`) and writes the DRAM traffic of the decode kernel per batch into a small JSON file that
bench.py quotes as roofline.traffic:  python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/traffic_r02.json <changes in batch>
The report is captured over `tools/ab_decode.py` (amg_bench_decode re-runs the decode of the resident batch: one launch of
k_decode_tiles over the whole batch, then the list launch over the inflated changes). Traffic per batch = DRAM bytes of a
whole-batch launch + the list launch that follows it, averaged over the pairs in the capture."""
import csv, json, subprocess, sys
rep, out, changes = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def num(r, name):
    v = float(r[col[name]].replace(',', '')); u = units[col[name]]
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'second': 1}.get(u, 1)
launches = []
for r in rows[2:]:
    if 'k_decode_tiles' not in r[col['Kernel Name']]:
        continue
    launches.append({'grid': int(float(r[col['launch__grid_size']])), 'block': int(float(r[col['launch__block_size']])),
                     'read': num(r, 'dram__bytes_read.sum'), 'write': num(r, 'dram__bytes_write.sum'), 'seconds': num(r, 'gpu__time_duration.sum'),
                     'registers': int(float(r[col['launch__registers_per_thread']]))})
full = -(-changes // launches[0]['block'])
batches = []
for i, l in enumerate(launches):
    if l['grid'] != full:
        continue
    b = {'range_launch': l, 'bytes': l['read'] + l['write'], 'seconds': l['seconds']}
    if i + 1 < len(launches) and launches[i + 1]['grid'] < full // 4:
        b['list_launch'] = launches[i + 1]; b['bytes'] += launches[i + 1]['read'] + launches[i + 1]['write']; b['seconds'] += launches[i + 1]['seconds']
    batches.append(b)
if not batches:
    sys.exit('no whole-batch launch of k_decode_tiles (grid %d) in the capture' % full)
total = sum(b['bytes'] for b in batches) / len(batches)
json.dump({'source': rep, 'kernel': 'k_decode_tiles', 'captured_batches': batches, 'dram_bytes_per_change': total / changes, 'changes_in_batch': changes,
           'decode_total_bytes': int(total),
           'note': 'dram__bytes_read.sum + dram__bytes_write.sum of a whole-batch launch of k_decode_tiles plus its list launch (inflated changes), averaged over the captured batches'}, open(out, 'w'), indent=1)
print(open(out).read()[:1500])
