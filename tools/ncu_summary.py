"""Condenses `ncu -i X.ncu-rep --page raw --csv` output into the per-kernel summary kept under
profiles/:  python tools/ncu_summary.py out.csv raw1.csv [raw2.csv ...]"""
import csv
import sys

WANT = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum',
    'dram__bytes_write.sum', 'lts__t_sectors_srcunit_tex.sum',
    'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
    'launch__block_size', 'launch__cluster_dim_x', 'launch__occupancy_limit_shared_mem',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]
out_path, raws = sys.argv[1], sys.argv[2:]
rows_out = []
for path in raws:
  rows = list(csv.reader(open(path)))
  hdr, units = rows[0], rows[1]
  ki = hdr.index('Kernel Name')
  for r in rows[2:]:
    d = {'source': path.split('/')[-1], 'kernel': r[ki]}
    for w in WANT:
      if w in hdr:
        d[w] = f'{r[hdr.index(w)]} {units[hdr.index(w)]}'.strip()
    st = [(hdr[i], r[i]) for i in range(len(hdr))
          if 'smsp__average_warps_issue_stalled' in hdr[i]
          and hdr[i].endswith('_per_issue_active.ratio') and 'not_issued' not in hdr[i]]
    st = sorted(st, key=lambda x: -float(x[1] or 0))[:4]
    d['top_stalls_per_issue'] = '; '.join(
        f"{a.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}"
        f"={float(b):.2f}" for a, b in st)
    rows_out.append(d)
keys = ['source', 'kernel'] + WANT + ['top_stalls_per_issue']
with open(out_path, 'w', newline='') as f:
  w = csv.DictWriter(f, fieldnames=keys)
  w.writeheader()
  for d in rows_out:
    w.writerow(d)
print('wrote', out_path, len(rows_out), 'kernels')
