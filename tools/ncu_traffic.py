"""Reads an ncu report (`ncu --set full`) and writes the DRAM traffic of the decode kernel per batch into a small JSON file that
bench.py quotes as roofline.traffic:  python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/traffic_r02.json <changes in batch>
The capture holds launches of k_decode_tiles over PIECES of the batch (one launch per uploaded piece); traffic per batch =
bytes per change of the captured launches x changes of the batch."""
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
    grid = int(float(r[col['launch__grid_size']])); block = int(float(r[col['launch__block_size']]))
    launches.append({'grid': grid, 'changes': grid * block, 'read': num(r, 'dram__bytes_read.sum'), 'write': num(r, 'dram__bytes_write.sum'), 'seconds': num(r, 'gpu__time_duration.sum'),
                     'registers': int(float(r[col['launch__registers_per_thread']]))})
n = sum(l['changes'] for l in launches)
per_change = sum(l['read'] + l['write'] for l in launches) / n
json.dump({'source': rep, 'kernel': 'k_decode_tiles', 'captured_launches': launches, 'dram_bytes_per_change': per_change, 'changes_in_batch': changes,
           'decode_total_bytes': int(per_change * changes),
           'note': 'dram__bytes_read.sum + dram__bytes_write.sum of the captured launches (each covers one uploaded piece of the batch), scaled to the batch'}, open(out, 'w'), indent=1)
print(open(out).read())
