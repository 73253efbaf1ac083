"""Expectations under the lattice's path distribution (the expectation semiring of
semirings.py:404-484 run through the recursion -- which the reference cannot do):
RecognitionLattice.expectation / .entropy against a brute-force enumeration of all paths, and the
fused kernel (lt_lattice_expectation: no [B,T,C,V] posterior tensor) against the composition of
lt_lattice_backward's posteriors with the values."""
import itertools

import numpy as np
import numpy.testing as npt
import pytest
import torch

from test_gpu_lattice import cuda, frames_for, make_lattice

pytestmark = pytest.mark.gpu


def _enumerate(blank, lex, vals_b, vals_l, t_len, vocab):
  """All alignment paths of a bigram FrameDependent lattice: per frame blank (0) or a label."""
  ws, vs = [], []
  for path in itertools.product(range(vocab + 1), repeat=t_len):
    state, w, v = 0, 0.0, 0.0
    for t, y in enumerate(path):
      if y == 0:
        w += blank[t, state]; v += vals_b[t, state]
      else:
        w += lex[t, state, y - 1]; v += vals_l[t, state, y - 1]
        state = y
    ws.append(w); vs.append(v)
  ws, vs = np.array(ws), np.array(vs)
  log_z = np.logaddexp.reduce(ws)
  p = np.exp(ws - log_z)
  return log_z, float((p * vs).sum()), float(-(p * (ws - log_z)).sum())


def test_expectation_and_entropy_against_path_enumeration():
  vocab, t, b = 2, 5, 2
  rng = np.random.RandomState(3)
  table = rng.randn(b, t, vocab + 1, vocab + 1)
  vals = rng.randn(b, t, vocab + 1, vocab + 1)
  nf = np.array([5, 3])
  lattice = make_lattice(vocab, 1, -1, cuda(table))
  log_z, e = lattice.expectation(frames_for(b, t), cuda(nf), cuda(vals[..., 0]), cuda(vals[..., 1:]),
                                 cache=None)
  h = lattice.entropy(frames_for(b, t), cuda(nf), cache=None)
  for i in range(b):
    z, ev, ent = _enumerate(table[i, :, :, 0], table[i, :, :, 1:], vals[i, :, :, 0],
                            vals[i, :, :, 1:], nf[i], vocab)
    npt.assert_allclose(float(log_z[i]), z, rtol=1e-6)
    npt.assert_allclose(float(e[i]), ev, rtol=1e-5, atol=1e-6)
    npt.assert_allclose(float(h[i]), ent, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('vocab,t', [(64, 11), (192, 7), (256, 300)])
def test_fused_expectation_kernel_equals_posteriors_times_values(vocab, t):
  """Bigram fast-path shapes (odd batch, ragged and empty utterances, pruned arcs): the fused
  kernel against posteriors x values through the generic kernels, for given values and for the
  arcs' own weights (entropy)."""
  from last_torch_b200 import _native as N
  b = 3
  rng = np.random.RandomState(vocab + t)
  table = rng.randn(b, t, vocab + 1, vocab + 1).astype(np.float32)
  drop = rng.rand(*table.shape) < 0.03
  drop[..., 0] = False
  table[drop] = -np.inf
  vals = rng.rand(b, t, vocab + 1, vocab + 1).astype(np.float32)
  nf = np.array([t, (2 * t) // 3, 0])
  assert N.lib().lt_lattice_expectation_supported(vocab, 1, -1, 0) == 1
  assert N.lib().lt_lattice_expectation_supported(vocab, 1, -1, 1) == 0
  assert N.lib().lt_lattice_expectation_supported(vocab, 2, -1, 0) == 0
  out = {}
  for flags in (0, 1):             # 1: LT_FLAG_FORCE_GENERIC -> composed from the posteriors
    lattice = make_lattice(vocab, 1, -1, cuda(table), flags)
    launches = N.lib().lt_launch_count()
    z, e = lattice.expectation(frames_for(b, t), cuda(nf), cuda(vals[..., 0]), cuda(vals[..., 1:]),
                               cache=None)
    h = lattice.entropy(frames_for(b, t), cuda(nf), cache=None)
    out[flags] = (z.cpu().numpy(), e.cpu().numpy(), h.cpu().numpy(), N.lib().lt_launch_count() - launches)
  npt.assert_allclose(out[0][0], out[1][0], rtol=1e-6)
  npt.assert_allclose(out[0][1], out[1][1], rtol=2e-5, atol=1e-6)
  npt.assert_allclose(out[0][2], out[1][2], rtol=2e-5, atol=2e-4)
  assert out[0][1][2] == 0.0 and np.all(np.isfinite(out[0][1])) and np.all(np.isfinite(out[0][2]))
  # values in [0, 1): the expectation is at most the number of arcs on a path
  assert np.all(out[0][1][:2] > 0) and np.all(out[0][1][:2] < nf[:2])
  assert np.all(out[0][2][:2] > 0)         # a distribution over many paths has positive entropy
