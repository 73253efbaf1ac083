"""Writes the GPU-vs-reference error table (run on the GPU box):

    python tools/parity_table.py [out.json]          # default gpurun_out/r02_parity_errors.json

For every fixture: max |reference_fp32 - truth| and max |gpu - truth| (absolute, and relative
over the entries above 1e-4 of the largest), truth = the oracle in float64 on the same fp32 inputs
(tests/parity_common.py).  The headline-size rows are produced twice: with the renormalised
recursion (default) and with the plain fp32 recursion (LT_NO_NORM=1, in a subprocess).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402


def headline_rows(tag):
  import parity_common as P
  rows = []
  for f in sorted(os.listdir(P.GOLDEN_DIR)):
    if not (f.startswith('headline_') and f.endswith('.npz')):
      continue
    g = np.load(os.path.join(P.GOLDEN_DIR, f))
    table, labels = P.headline_inputs(g)
    truth = P.headline_truth(g, table, labels)
    gpu = P.gpu_headline(g, table, labels)
    rows += P.headline_rows(f[:-4] + tag, g, truth, gpu)
    nf = g['num_frames']
    sums = np.concatenate([gpu['grad_frame_sums'][b, :nf[b]] for b in range(len(nf))])
    rows.append({'case': f[:-4] + tag, 'quantity': 'max |sum of the gradient of a real frame| (= 0)',
                 'gpu_abs': float(np.abs(sums).max())})
  for k in (-1, 2, 3):          # context_size 2 (thread-per-column / row kernels), T = 200
    rows += P.trigram_rows(k, tag=tag)[0]
  return rows


def main():
  out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out',
                                                          'r02_parity_errors.json')
  if os.environ.get('LT_PARITY_HEADLINE_ONLY'):
    print(json.dumps(headline_rows(' [plain fp32 recursion, LT_NO_NORM=1]')))
    return
  import __graft_entry__ as ge
  ge.build()
  import torch
  import parity_common as P
  rows = []
  for f in sorted(os.listdir(P.GOLDEN_DIR)):
    if f.startswith('lattice_') and f.endswith('.npz'):
      loss, grad = P.gpu_lattice_golden(f)
      rows += P.lattice_golden_rows(f, loss, grad)
  rows += headline_rows('')
  plain = subprocess.run([sys.executable, os.path.abspath(__file__)], capture_output=True,
                         text=True, env=dict(os.environ, LT_NO_NORM='1',
                                             LT_PARITY_HEADLINE_ONLY='1'))
  if plain.returncode == 0:
    rows += json.loads(plain.stdout.strip().splitlines()[-1])
  else:
    rows.append({'case': 'plain fp32 recursion', 'error': plain.stderr[-400:]})
  for f in sorted(os.listdir(P.GOLDEN_DIR)):
    if f.startswith('jointlattice_') and f.endswith('.npz'):
      g = np.load(os.path.join(P.GOLDEN_DIR, f))
      for split in (True, False):
        loss, grads = P.gpu_joint_lattice(g, split)
        rows += P.joint_lattice_rows(f, loss, grads,
                                     ' [split rows]' if split else ' [fp32 hand-over]')
  g = P.synthetic_joint_case(seed=5, vocab=256, hidden=512, emb=96, feat=80, batch=2, t_max=200,
                             u=40)
  loss64, grads64 = P.joint_lattice_truth_large(g)
  for split in (True, False):
    loss, grads = P.gpu_joint_lattice(g, split)
    tag = 'joint_v256_h512_t200' + (' [split rows]' if split else ' [fp32 hand-over]')
    rows.append(P.row(tag, 'loss', None, loss, loss64))
    rows += [P.row(tag, 'grad_' + p, None, grads[p], grads64[p]) for p in P.PARAMS]
  for r in rows:
    if 'reference_fp32_abs' in r:
      r['within_2x_bar'] = bool(P.within_bar(r))
  doc = {
      'what': ('max error against a float64 evaluation of the oracle on identical fp32 inputs; '
               'reference_fp32 = outputs of the unmodified reference (tests/golden), gpu = this '
               'repo through the public API; rel = over entries >= 1e-4 of the largest'),
      'device': torch.cuda.get_device_name(0),
      'rows': rows,
  }
  os.makedirs(os.path.dirname(out), exist_ok=True)
  with open(out, 'w') as f:
    json.dump(doc, f, indent=1)
  bad = [r for r in rows if r.get('within_2x_bar') is False]
  print(f'{len(rows)} rows -> {out}; outside the 2x bar: {len(bad)}')
  for r in bad:
    print('  ', r)


if __name__ == '__main__':
  main()
