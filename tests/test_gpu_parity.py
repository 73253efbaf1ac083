"""Parity at the sizes the headline is quoted on, and the GPU-vs-reference error bar.

Every test compares the CUDA path (public API -> C ABI) with a float64 evaluation of the oracle on
identical fp32 inputs; where the fixture also holds the fp32 outputs of the unmodified reference,
the test asserts  max|gpu - truth| <= 2 x max|reference - truth|  (tests/parity_common.py).  The
same rows are written to profiles/r02_parity_errors.json by tools/parity_table.py.

Sizes: tests/golden/lattice_*.npz (test scale, incl. BASELINE configs[0]); the reference run at
configs[1] geometry, T = 1000 (tests/golden/headline_*.npz); lattice-only B = 4, T = 1000,
V = 256 against the double C oracle; RecognitionLattice.forward with JointWeightFn at V = 256,
H = 512, T = 200 with and without the split-row hand-over; configs[2] (trigram vocab 64,
FrameLabelDependent(2)) at T = 200.
"""
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

import parity_common as P
from conftest import GOLDEN_DIR, golden_files
from oracle import c_oracle
from oracle import lattice_oracle as O

pytestmark = pytest.mark.gpu


def _show(rows):
  for r in rows:
    print(f"{r['case']:34s} {r['quantity']:30s} scale {r['scale']:.3g}  ref abs "
          f"{r['reference_fp32_abs']} rel {r['reference_fp32_rel']}  gpu abs {r['gpu_abs']:.3g} "
          f"rel {r['gpu_rel']:.3g}")


@pytest.mark.parametrize('fname', golden_files('lattice_'))
def test_gpu_error_within_twice_the_reference_error(fname):
  """Loss and gradient of every reference fixture: the GPU is at most twice as far from the fp64
  truth as the fp32 reference is."""
  loss, grad = P.gpu_lattice_golden(fname)
  rows = P.lattice_golden_rows(fname, loss, grad)
  _show(rows)
  for r in rows:
    assert P.within_bar(r), r


@pytest.mark.parametrize('fname', golden_files('headline_'))
def test_headline_size_against_the_reference_run(fname):
  """BASELINE configs[1] geometry (bigram vocab 256, T = 1000, U = 120, one ragged utterance):
  the reference's own fp32 run vs the GPU vs the double C oracle.  Loss, logZ and gradients meet
  the 2x bar; with the renormalised recursion the gradients are also within 1e-5 relative of
  the truth (north_star), which the fp32 reference is not at this length."""
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  table, labels = P.headline_inputs(g)
  truth = P.headline_truth(g, table, labels)
  gpu = P.gpu_headline(g, table, labels)
  rows = P.headline_rows(fname[:-4], g, truth, gpu)
  _show(rows)
  for r in rows:
    assert P.within_bar(r), r
  # north_star: 1e-5 relative -- loss / logZ of their value, gradients of every significant entry
  npt.assert_allclose(gpu['loss'], truth['loss'], rtol=1e-6)
  npt.assert_allclose(gpu['log_z'], truth['log_z'], rtol=1e-6)
  # measured (profiles/r02_parity_errors.json): 7e-6 / 3e-6 relative, 2.6e-6 / 1e-6 absolute --
  # the floor of MUFU ex2 / lg2 (2^-22 relative each) accumulated over 1000 frames; the fp32
  # reference is at 8e-4 / 1.2e-3 relative on the same inputs
  by = {r['quantity']: r for r in rows}
  assert by['grad_blank']['gpu_rel'] < 5e-5, by['grad_blank']
  assert by['grad_lexical (sampled frames)']['gpu_rel'] < 5e-5
  assert by['grad_blank']['gpu_abs'] < 1e-5
  assert by['grad_blank']['gpu_abs'] < 0.2 * by['grad_blank']['reference_fp32_abs']
  # marginals of every real frame sum to one (denominator) minus one (numerator)
  nf = g['num_frames']
  for b in range(len(nf)):
    npt.assert_allclose(gpu['grad_frame_sums'][b, :nf[b]], 0.0, atol=2e-5)
    npt.assert_array_equal(gpu['grad_frame_sums'][b, nf[b]:], 0.0)


@pytest.mark.parametrize('fname', golden_files('headline_'))
def test_headline_size_viterbi_matches_the_reference_path(fname):
  """MaxTropical at T = 1000: distance within 1e-6 of the reference, Viterbi arcs bit-exact
  (the reference's own MaxTropical autograd gives the one-hot arcs)."""
  lt = P._lt()
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  table, _ = P.headline_inputs(g)
  v, b, t = int(g['vocab']), int(g['batch']), int(g['t_max'])
  leaf = table.cuda().requires_grad_()
  lattice = P.table_lattice(v, 1, -1, leaf)
  nf = P._cuda(g['num_frames'])
  dist, _ = lattice._forward(cache=None, frames=P.frames_for(b, t), num_frames=nf,
                             semiring=lt.semirings.MaxTropical)
  npt.assert_allclose(dist.detach().cpu(), g['maxtropical_dist'], rtol=1e-6)
  (gd,) = torch.autograd.grad(dist.sum(), leaf)
  flat = gd.reshape(b, t, -1)
  taken = flat.sum(-1).cpu().numpy()
  arc = flat.argmax(-1).cpu().numpy()
  npt.assert_array_equal(taken, (np.arange(t)[None] < g['num_frames'][:, None]).astype(np.float32))
  state = np.where(taken > 0, arc // (1 + v), -1)
  label = np.where(taken > 0, arc % (1 + v), -1)
  npt.assert_array_equal(state, g['viterbi_state'])
  npt.assert_array_equal(label, g['viterbi_label'])
  labels, num, weights = lattice.shortest_path(frames=P.frames_for(b, t), num_frames=nf, cache=None)
  npt.assert_array_equal(labels.cpu().numpy()[taken > 0], g['viterbi_label'][taken > 0])
  npt.assert_allclose(weights.cpu(), g['maxtropical_dist'], rtol=1e-6)


def test_lattice_only_b4_t1000_against_the_double_oracle():
  """Lattice-only Log loss + FULL gradients at B = 4, T = 1000, V = 256 (ragged) against the
  double build of the C oracle: loss 1e-6, every gradient entry within 5e-5 relative (entries
  above 1e-4 of the largest; measured 1.4e-5) and 1e-5 absolute (measured 4.4e-6)."""
  b, t, v, u = 4, 1000, 256, 120
  rng = np.random.RandomState(7)
  gen = torch.Generator().manual_seed(77)
  table = torch.randn([b, t, v + 1, 1 + v], generator=gen)
  nf = np.array([1000, 640, 1000, 873])
  labels = rng.randint(1, v + 1, size=(b, u))
  nl = np.array([120, 77, 0, 120])
  tab = table.numpy()
  loss64, gb64, gl64, _, _ = c_oracle.lattice_loss_and_grads(
      np.ascontiguousarray(tab[..., 0]), np.ascontiguousarray(tab[..., 1:]), nf, labels, nl, v, 1,
      real='f64')
  leaf = table.cuda().requires_grad_()
  lattice = P.table_lattice(v, 1, -1, leaf)
  loss = lattice(frames=P.frames_for(b, t), num_frames=P._cuda(nf), labels=P._cuda(labels),
                 num_labels=P._cuda(nl), cache=None)
  (gt,) = torch.autograd.grad(loss.sum(), leaf)
  npt.assert_allclose(loss.detach().cpu().numpy(), loss64, rtol=1e-6)
  gt = gt.cpu().numpy()
  rows = [P.row('b4_t1000_v256', 'grad_blank', None, gt[..., 0], gb64),
          P.row('b4_t1000_v256', 'grad_lexical', None, gt[..., 1:], gl64)]
  _show([dict(r, reference_fp32_abs=float('nan'), reference_fp32_rel=float('nan')) for r in rows])
  for r in rows:
    assert r['gpu_rel'] < 5e-5 and r['gpu_abs'] < 1e-5, r


@pytest.mark.parametrize('split', [True, False])
def test_joint_lattice_v256_h512_t200_against_the_double_oracle(split):
  """RecognitionLattice.forward with JointWeightFn at the headline widths (V = 256, H = 512,
  tcgen05 forward, TMA lattice kernels, tensor-core backward) and T = 200, ragged: loss and all
  seven parameter gradients against float64 (joint network in numpy, lattice in the double C
  oracle).  The bf16x3 operand split bounds the logits at ~1e-5 of their scale; gradients are
  asserted at 1e-4 of theirs (the measured values are in profiles/r02_parity_errors.json)."""
  g = P.synthetic_joint_case(seed=5, vocab=256, hidden=512, emb=96, feat=80, batch=2, t_max=200,
                             u=40)
  loss64, grads64 = P.joint_lattice_truth_large(g)
  loss, grads = P.gpu_joint_lattice(g, split)
  npt.assert_allclose(loss, loss64, rtol=2e-5)
  rows = [P.row('joint_v256_h512_t200', 'grad_' + p, None, grads[p], grads64[p]) for p in P.PARAMS]
  _show([dict(r, reference_fp32_abs=float('nan'), reference_fp32_rel=float('nan')) for r in rows])
  for r in rows:
    assert r['gpu_abs'] <= 1e-4 * r['scale'] + 1e-7, r


def test_config2_trigram_fld2_t200_against_the_oracle():
  """BASELINE configs[2]: FullNGram(vocab 64, context_size 2) = 4161 states,
  FrameLabelDependent(2), T = 200: MaxTropical distance (1e-6) and Viterbi labels (bit-exact)
  against the numpy oracle, Log loss + gradients against the double C oracle."""
  lt = P._lt()
  b, t, v, n, k, u = 2, 200, 64, 2, 2, 30
  rng = np.random.RandomState(11)
  c = 1 + v + v * v
  table_np = rng.randn(b, t, c, 1 + v).astype(np.float32)
  nf = np.array([200, 131])
  labels = rng.randint(1, v + 1, size=(b, u))
  nl = np.array([30, 17])
  blank = np.ascontiguousarray(table_np[..., 0])
  lex = np.ascontiguousarray(table_np[..., 1:])
  o_dist, _, _, o_labels = O.viterbi(blank, lex, nf, O.FullNGram(v, n), k, False)
  loss64, gb64, gl64, _, _ = c_oracle.lattice_loss_and_grads(blank, lex, nf, labels, nl, v, n, k,
                                                             real='f64')
  leaf = P._cuda(table_np).requires_grad_()
  lattice = P.table_lattice(v, n, k, leaf)
  frames = P.frames_for(b, t)
  path, num, weights = lattice.shortest_path(frames=frames, num_frames=P._cuda(nf), cache=None)
  npt.assert_allclose(weights.cpu(), o_dist, rtol=1e-6)
  npt.assert_array_equal(path.cpu(), o_labels)
  npt.assert_array_equal(num.cpu(), (k + 1) * nf)
  loss = lattice(frames=frames, num_frames=P._cuda(nf), labels=P._cuda(labels),
                 num_labels=P._cuda(nl), cache=None)
  (gt,) = torch.autograd.grad(loss.sum(), leaf)
  npt.assert_allclose(loss.detach().cpu().numpy(), loss64, rtol=1e-5)
  gt = gt.cpu().numpy()
  # renormalised state in the thread-per-column forward / 8-lanes-per-row backward (logZ ~ 1.2e3
  # here; the plain fp32 recursion, LT_NO_NORM=1, sits at 3e-5 absolute)
  for got, want in ((gt[..., 0], gb64), (gt[..., 1:], gl64)):
    assert np.abs(got - want).max() < 1e-5
    npt.assert_allclose(got, want, rtol=1e-4, atol=5e-6)


@pytest.mark.parametrize('k', [-1, 2, 3])
def test_trigram_v32_t200_against_the_double_oracle(k):
  """FullNGram(vocab 32, context_size 2) = 1057 states, T = 200, ragged, FrameDependent and
  FrameLabelDependent(2 / 3): Log loss and every gradient entry of the TMA kernels
  (lattice_cols.cu forward, lattice_rows.cu backward, renormalised state) against the double
  build of the C oracle."""
  rows, loss_rel = P.trigram_rows(k)
  _show([dict(r, reference_fp32_abs=float('nan'), reference_fp32_rel=float('nan')) for r in rows])
  assert loss_rel < 1e-6
  for r in rows:
    assert r['gpu_abs'] < 2e-5, r


@pytest.mark.parametrize('k', [-1, 2])
def test_trigram_v32_t200_wide_weights_against_the_double_oracle(k):
  """The same lattices with weights ~ N(0, 8^2): alpha grows by ~45 log2 units per frame, which is
  what the one-frame-late, predicted shift of the FrameDependent kernel has to follow without
  ringing (lattice_cols.cu: measured 1.0e-5 absolute; half-weight feedback of the late maximum
  gave 2.6e-5).  FrameLabelDependent(2) shifts once per frame, so its level vectors sit up to two
  such steps above the shifted alpha and fp32 resolves their sums to ~7e-5 (gradients up to 2)."""
  rows, loss_rel = P.trigram_rows(k, scale=8.0)
  _show([dict(r, reference_fp32_abs=float('nan'), reference_fp32_rel=float('nan')) for r in rows])
  assert loss_rel < 1e-6
  for r in rows:
    assert r['gpu_abs'] < (2e-5 if k < 0 else 1.5e-4), r
