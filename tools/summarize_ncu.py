"""Turn an .ncu-rep into a small committed summary: python tools/summarize_ncu.py gpurun_out/x.ncu-rep profiles/x.md"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
units = rows[1]
with open(out, 'w') as fh:
    fh.write('# ncu summary of `{}`\n\n(`ncu --set full --clock-control none --import-source on`; raw page; one row per captured launch)\n\n'.format(rep))
    fh.write('| ' + ' | '.join('{} [{}]'.format(w, units[i]) if units[i] else w for w, i in idx) + ' |\n')
    fh.write('|' + '---|' * len(idx) + '\n')
    for r in rows[2:]:
        fh.write('| ' + ' | '.join(r[i][:70] for _, i in idx) + ' |\n')
print(open(out).read()[:3000])
