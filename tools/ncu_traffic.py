"""profiles/r02_ncu_traffic.json from the raw CSV of an `ncu --set full` capture of
`bench.py --steps 1 --warmup 1 --no-cpu --no-extras` (tools/gpu_round.sh prof):

    ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv "r02 session N, commit <hash>"

Per C-ABI entry point: DRAM bytes (read + write) per launch, averaged over the captured launches
of the kernels behind it, and the SHA-256 of the sources the kernel is compiled from -- bench.py
reports `roofline.traffic` from this file only while those sources are unchanged."""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = 'last_torch_b200/csrc/'
# entry point -> (kernel-name substrings, any of; sources)
MAP = {
    'lt_lattice_forward': (['lattice_forward_fast2<'], ['lattice_fast2.cu', 'fast2.cuh', 'fast_ptx.cuh']),
    # the resident step's K2 writes fp32 gradients (SPLIT = 0)
    'lt_lattice_backward': (['lattice_backward_fast2<1, 256, 0,'], ['lattice_fast2.cu', 'fast2.cuh', 'fast_ptx.cuh']),
    'lt_lattice_backward[split rows]': (['lattice_backward_fast2<1, 256, 1,'], ['lattice_fast2.cu', 'fast2.cuh', 'fast_ptx.cuh']),
    'lt_joint_forward': (['joint_forward_ts_kernel', 'joint_forward_tc_kernel'], ['joint_fwd_ts.cu', 'joint_tc.cu', 'umma.cuh']),
    'lt_joint_backward[dgrad]': (['joint_dgrad2_kernel'], ['joint_dgrad2.cu', 'umma.cuh']),
    'lt_joint_backward[wgrad]': (['joint_wgrad_tc_kernel'], ['joint_tc.cu', 'umma.cuh']),
}


def digest(files):
  h = hashlib.sha256()
  for f in files:
    with open(os.path.join(ROOT, CSRC, f), 'rb') as fh:
      h.update(fh.read())
  return h.hexdigest()


def main():
  raw, capture = sys.argv[1], sys.argv[2]
  rows = list(csv.reader(open(raw)))
  hdr = rows[0]
  ix = {h: i for i, h in enumerate(hdr)}
  units = rows[1]

  def gb(r, key):
    v, u = float(r[ix[key]]), units[ix[key]]
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]

  out = {'capture': capture, 'kernels': {}}
  for entry, (subs, files) in MAP.items():
    sel = [r for r in rows[2:] if any(s in r[ix['Kernel Name']] for s in subs)]
    if not sel:
      continue
    tot = [gb(r, 'dram__bytes_read.sum') + gb(r, 'dram__bytes_write.sum') for r in sel]
    ms = [float(r[ix['gpu__time_duration.sum']]) for r in sel]
    out['kernels'][entry] = {
        'kernel': sel[0][ix['Kernel Name']][:120], 'launches_captured': len(sel),
        'dram_bytes_per_launch': sum(tot) / len(tot), 'ms_under_ncu': sum(ms) / len(ms),
        'source': [CSRC + f for f in files], 'source_sha256': digest(files)}
  path = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
  with open(path, 'w') as f:
    json.dump(out, f, indent=1)
  print('wrote', path, list(out['kernels']))


if __name__ == '__main__':
  main()
