"""GPU parity tests of the lattice kernels (through the C ABI / public API)
against the golden fixtures produced by the reference and against the numpy
oracle.  Tolerances: fp32 Log/Real values and gradients 1e-5 relative to the
magnitude of the quantity (rtol 1e-5 on O(1) values, with an atol that covers
fp32 cancellation in alpha + w + beta - logZ); MaxTropical distances 1e-6;
Viterbi one-hot gradients / labels bit-exact.
"""
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

from conftest import GOLDEN_DIR, golden_files
from oracle import lattice_oracle as O

pytestmark = pytest.mark.gpu

SR = ['Real', 'Log', 'MaxTropical']
CLUSTER_FLAGS = [0, 1 << 8, 2 << 8, 4 << 8]


def _lt():
  import last_torch_b200 as last_torch
  return last_torch


def make_lattice(vocab, ctx, k, table, flags=0):
  lt = _lt()
  alignment = (lt.alignments.FrameDependent() if k < 0 else
               lt.alignments.FrameLabelDependent(max_expansions=k))
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=vocab, context_size=ctx),
      alignment=alignment,
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  lattice.kernel_flags = flags
  return lattice


def frames_for(b, t):
  return torch.arange(t, device='cuda', dtype=torch.float32)[None, :, None].expand(b, t, 1)


def cuda(x, dtype=torch.float32):
  return torch.as_tensor(np.asarray(x), device='cuda').to(dtype)


@pytest.mark.parametrize('flags', CLUSTER_FLAGS)
@pytest.mark.parametrize('fname', golden_files('lattice_'))
def test_lattice_golden(fname, flags):
  lt = _lt()
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  vocab, ctx, k = int(g['vocab']), int(g['context_size']), int(g['k'])
  b, t = g['table'].shape[:2]
  frames = frames_for(b, t)
  nf, lab, nl = cuda(g['num_frames']), cuda(g['labels']), cuda(g['num_labels'])
  for name in SR:
    semiring = getattr(lt.semirings, name)
    tab = g['Real_table'] if name == 'Real' else g['table']
    table = cuda(tab).requires_grad_()
    lattice = make_lattice(vocab, ctx, k, table, flags)
    dist, alphas = lattice._forward(cache=None, frames=frames, num_frames=nf, semiring=semiring)
    npt.assert_allclose(dist.detach().cpu(), g[f'{name}_dist'], rtol=2e-5, atol=1e-6,
                        err_msg=f'{name} dist')
    npt.assert_allclose(alphas.detach().cpu(), g[f'{name}_alphas'], rtol=2e-5, atol=2e-5,
                        err_msg=f'{name} alphas')
    (gd,) = torch.autograd.grad(dist.sum(), table)
    if name == 'MaxTropical':
      npt.assert_array_equal(gd.cpu(), g[f'{name}_dist_grad'])
    else:
      npt.assert_allclose(gd.cpu(), g[f'{name}_dist_grad'], rtol=2e-4, atol=2e-6,
                          err_msg=f'{name} dist grad')
    table = cuda(tab).requires_grad_()
    lattice = make_lattice(vocab, ctx, k, table, flags)
    sd = lattice._string_forward(cache=None, frames=frames, num_frames=nf, labels=lab,
                                 num_labels=nl, semiring=semiring)
    npt.assert_allclose(sd.detach().cpu(), g[f'{name}_string'], rtol=2e-5, atol=1e-6,
                        err_msg=f'{name} string')
    reach = torch.isfinite(sd) if name != 'Real' else sd != 0
    if bool(reach.any()):
      (gs,) = torch.autograd.grad(torch.where(reach, sd, torch.zeros_like(sd)).sum(), table)
      if name == 'MaxTropical':
        npt.assert_array_equal(gs.cpu(), g[f'{name}_string_grad'])
      else:
        npt.assert_allclose(gs.cpu(), g[f'{name}_string_grad'], rtol=2e-4, atol=2e-6,
                            err_msg=f'{name} string grad')
  # loss value + gradient (= denominator marginals - numerator marginals)
  table = cuda(g['table']).requires_grad_()
  lattice = make_lattice(vocab, ctx, k, table, flags)
  loss = lattice(frames=frames, num_frames=nf, labels=lab, num_labels=nl, cache=None)
  npt.assert_allclose(loss.detach().cpu(), g['loss'], rtol=2e-5, atol=2e-5)
  fin = torch.isfinite(loss)
  (gt,) = torch.autograd.grad(torch.where(fin, loss, torch.zeros_like(loss)).sum(), table)
  expect = g['Log_dist_grad'] - g['Log_string_grad']
  fin_np = fin.cpu().numpy()
  # 'wide' uses weights ~N(0, 8^2): |alpha| ~ 1e2 and the reference's own two gradient
  # oracles (patched autograd vs alignment.backward) already differ by 1.8e-5 there.
  atol = 1e-4 if 'wide' in fname else 4e-6
  npt.assert_allclose(gt.cpu().numpy()[fin_np], expect[fin_np], rtol=2e-4, atol=atol)
  # Viterbi labels agree with the oracle's true labels; weights with the reference
  labels, num_labels, weights = lattice.shortest_path(frames=frames, num_frames=nf, cache=None)
  kk, fd = (0, True) if k < 0 else (k, False)
  blank, lex = np.ascontiguousarray(g['table'][..., 0]), np.ascontiguousarray(g['table'][..., 1:])
  o_dist, _, _, o_labels = O.viterbi(blank, lex, g['num_frames'], O.FullNGram(vocab, ctx), kk, fd)
  npt.assert_allclose(weights.cpu(), g['MaxTropical_dist'], rtol=1e-6)
  npt.assert_array_equal(labels.cpu(), o_labels)
  npt.assert_array_equal(num_labels.cpu(), (kk + 1) * g['num_frames'])


def _random_case(seed, b, t, vocab, ctx, k, u, scale=1.0, ragged=True):
  rng = np.random.RandomState(seed)
  c = sum(vocab**i for i in range(ctx + 1))
  table = (rng.randn(b, t, c, 1 + vocab) * scale).astype(np.float32)
  nf = rng.randint(t // 2, t + 1, size=b) if ragged else np.full(b, t)
  nf[0] = t
  labels = rng.randint(1, vocab + 1, size=(b, u))
  per_frame = 1 if k < 0 else k
  nl = np.minimum(rng.randint(0, u + 1, size=b), nf * per_frame)
  return table, nf, labels, nl


ORACLE_CASES = [
    # seed, B, T, V, n, k, U, scale
    (1, 4, 50, 16, 1, -1, 12, 1.0),      # BASELINE.json configs[0]
    (2, 3, 24, 8, 2, -1, 10, 1.0),
    (3, 2, 16, 64, 1, -1, 8, 2.0),
    (4, 3, 12, 6, 2, 2, 10, 1.0),        # FrameLabelDependent(2), trigram states
    (5, 2, 10, 5, 1, 3, 9, 1.0),
    (6, 2, 12, 256, 1, -1, 8, 1.0),      # configs[1] width (C=257, V=256), short T
    (7, 2, 8, 7, 0, -1, 5, 1.0),         # unigram: single context state
    (8, 2, 6, 64, 2, 2, 6, 1.0),         # configs[2] width (C=4161), FLD(2)
    (9, 5, 40, 33, 1, -1, 11, 4.0),      # odd vocab (no 16-byte alignment), wide range
    (10, 3, 20, 128, 1, -1, 10, 1.0),    # TMA fast path, cluster of 2
    (11, 2, 9, 192, 1, -1, 6, 3.0),      # TMA fast path, cluster of 3
    (12, 5, 37, 64, 1, -1, 9, 1.0),      # TMA fast path, single CTA, T > ring depth
    (13, 3, 9, 16, 2, -1, 5, 1.0),       # thread-per-column path (n=2), single CTA, FrameDependent
    (14, 2, 7, 32, 2, 3, 5, 1.0),        # thread-per-column path, cluster of 4, FLD(3)
    (15, 2, 5, 64, 2, -1, 5, 2.0),       # configs[2] width, cluster of 8, FrameDependent
    (16, 3, 8, 8, 3, 2, 6, 1.0),         # thread-per-column path, 4-gram states (n=3), cluster of 2
    (17, 3, 9, 32, 2, -1, 5, 1.5),       # cols forward + 8-lanes-per-row backward, C = 1057
    (18, 2, 40, 32, 2, 1, 8, 1.0),       # same kernels, FLD(1): one level per frame, late shift
    (19, 2, 60, 32, 2, -1, 9, 6.0),      # same kernels, wide weights over 60 frames
]


def test_fast_path_ragged_and_empty_utterances():
  """Fast (TMA) path with num_frames in {0, 1, < ring depth, T}: same results
  as the generic kernels and as the oracle."""
  b, t, vocab, ctx, u = 5, 23, 64, 1, 4
  rng = np.random.RandomState(5)
  table_np = rng.randn(b, t, 1 + vocab, 1 + vocab).astype(np.float32)
  nf = np.array([0, 1, 2, 7, 23])
  labels = rng.randint(1, vocab + 1, size=(b, u))
  nl = np.array([0, 1, 2, 4, 3])
  tab64 = table_np.astype(np.float64)
  o_loss, o_gb, o_gl = O.lattice_loss_and_grads(
      tab64[..., 0].copy(), tab64[..., 1:].copy(), nf, labels, nl, O.FullNGram(vocab, ctx))
  outs = []
  for flags in [1, 0]:   # 1: LT_FLAG_FORCE_GENERIC, 0: TMA fast path
    table = cuda(table_np).requires_grad_()
    lattice = make_lattice(vocab, ctx, -1, table, flags)
    loss = lattice(frames=frames_for(b, t), num_frames=cuda(nf), labels=cuda(labels),
                   num_labels=cuda(nl), cache=None)
    (gt,) = torch.autograd.grad(loss.sum(), table)
    npt.assert_allclose(loss.detach().cpu(), o_loss, rtol=1e-5, atol=1e-5)
    npt.assert_allclose(gt.cpu().numpy()[..., 0], o_gb, rtol=1e-4, atol=1e-5)
    npt.assert_allclose(gt.cpu().numpy()[..., 1:], o_gl, rtol=1e-4, atol=1e-5)
    outs.append((loss.detach().cpu().numpy(), gt.cpu().numpy()))
  for other in outs[1:]:
    npt.assert_allclose(outs[0][0], other[0], rtol=2e-6, atol=2e-6)
    npt.assert_allclose(outs[0][1], other[1], rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize('flags', [0, 1 << 8, 4 << 8, 8 << 8])
@pytest.mark.parametrize('case', ORACLE_CASES)
def test_loss_and_grads_vs_oracle(case, flags):
  seed, b, t, vocab, ctx, k, u, scale = case
  table_np, nf, labels, nl = _random_case(seed, b, t, vocab, ctx, k, u, scale)
  kk, fd = (0, True) if k < 0 else (k, False)
  octx = O.FullNGram(vocab, ctx)
  tab64 = table_np.astype(np.float64)
  blank, lex = np.ascontiguousarray(tab64[..., 0]), np.ascontiguousarray(tab64[..., 1:])
  o_loss, o_gb, o_gl = O.lattice_loss_and_grads(blank, lex, nf, labels, nl, octx, kk, fd)
  table = cuda(table_np).requires_grad_()
  lattice = make_lattice(vocab, ctx, k, table, flags)
  loss = lattice(frames=frames_for(b, t), num_frames=cuda(nf), labels=cuda(labels),
                 num_labels=cuda(nl), cache=None)
  fin = np.isfinite(o_loss)
  npt.assert_array_equal(torch.isfinite(loss).cpu().numpy(), fin)
  # loss: 1e-5 relative (fp32 kernel vs fp64 oracle)
  npt.assert_allclose(loss.detach().cpu().numpy()[fin], o_loss[fin], rtol=1e-5, atol=1e-5)
  (gt,) = torch.autograd.grad(
      torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum(), table)
  gt = gt.cpu().numpy()
  # gradients are posteriors in [0, 1]; fp32 rounding of alpha+w+beta-logZ at |alpha| ~ 1e2
  # bounds the achievable accuracy at ~1e-5 of the value; atol covers tiny posteriors.
  npt.assert_allclose(gt[fin][..., 0], o_gb[fin], rtol=1e-4, atol=1e-5)
  npt.assert_allclose(gt[fin][..., 1:], o_gl[fin], rtol=1e-4, atol=1e-5)
  if fd:
    # free invariant: denominator marginals of each frame sum to one
    dist, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                               semiring=_lt().semirings.Log)
    (gd,) = torch.autograd.grad(dist.sum(), table)
    npt.assert_allclose(gd.sum((1, 2, 3)).cpu(), nf, rtol=1e-4)


@pytest.mark.parametrize('case', ORACLE_CASES)
def test_forward_all_semirings_vs_oracle(case):
  lt = _lt()
  seed, b, t, vocab, ctx, k, u, scale = case
  table_np, nf, labels, nl = _random_case(seed + 100, b, t, vocab, ctx, k, u, scale)
  kk, fd = (0, True) if k < 0 else (k, False)
  octx = O.FullNGram(vocab, ctx)
  for name, sr in [('Log', O.LOG), ('MaxTropical', O.MAXTROPICAL), ('Real', O.REAL)]:
    tab = table_np
    if name == 'Real':
      tab = (np.exp(np.clip(table_np, -40, 5) * 0.25) / (1 + vocab)).astype(np.float32)
    tab64 = tab.astype(np.float64)
    blank, lex = np.ascontiguousarray(tab64[..., 0]), np.ascontiguousarray(tab64[..., 1:])
    o_dist, o_alphas = O.lattice_forward(blank, lex, nf, octx, sr, kk, fd)
    o_sd = O.lattice_string_forward(blank, lex, nf, labels, nl, octx, sr, kk, fd)
    table = cuda(tab)
    lattice = make_lattice(vocab, ctx, k, table)
    dist, alphas = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                                    semiring=getattr(lt.semirings, name))
    sd = lattice._string_forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                                 labels=cuda(labels), num_labels=cuda(nl),
                                 semiring=getattr(lt.semirings, name))
    tol = dict(rtol=1e-6, atol=1e-5) if name == 'MaxTropical' else dict(rtol=1e-5, atol=1e-5)
    npt.assert_allclose(dist.cpu(), o_dist, err_msg=name, **tol)
    npt.assert_allclose(sd.cpu(), o_sd, err_msg=name, **tol)
    a = alphas.cpu().numpy()
    if name == 'Real':
      npt.assert_allclose(a, o_alphas, rtol=2e-5, atol=1e-30, err_msg=name)
    else:
      fin = np.isfinite(o_alphas)
      npt.assert_array_equal(np.isfinite(a), fin)
      npt.assert_allclose(a[fin], o_alphas[fin], rtol=1e-5, atol=2e-4, err_msg=name)


@pytest.mark.parametrize('case', ORACLE_CASES)
def test_viterbi_vs_oracle(case):
  seed, b, t, vocab, ctx, k, u, scale = case
  table_np, nf, _, _ = _random_case(seed + 200, b, t, vocab, ctx, k, u, scale)
  kk, fd = (0, True) if k < 0 else (k, False)
  blank, lex = np.ascontiguousarray(table_np[..., 0]), np.ascontiguousarray(table_np[..., 1:])
  o_dist, o_gb, o_gl, o_labels = O.viterbi(blank, lex, nf, O.FullNGram(vocab, ctx), kk, fd)
  table = cuda(table_np).requires_grad_()
  lattice = make_lattice(vocab, ctx, k, table)
  dist, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                             semiring=_lt().semirings.MaxTropical)
  npt.assert_allclose(dist.detach().cpu(), o_dist, rtol=1e-6)
  (gd,) = torch.autograd.grad(dist.sum(), table)
  # random continuous weights: the maximum is unique, so the path is bit-exact
  npt.assert_array_equal(gd.cpu().numpy()[..., 0], o_gb)
  npt.assert_array_equal(gd.cpu().numpy()[..., 1:], o_gl)
  labels, _, weights = lattice.shortest_path(frames=frames_for(b, t), num_frames=cuda(nf),
                                             cache=None)
  npt.assert_array_equal(labels.cpu(), o_labels)
  npt.assert_allclose(weights.cpu(), o_dist, rtol=1e-6)


@pytest.mark.parametrize('flags', [0])
@pytest.mark.parametrize('vocab', [64, 128, 192, 256])
def test_fast_path_generations(vocab, flags):
  """The TMA / cluster fast path on every supported vocabulary, with an odd batch, ragged lengths and
  -inf arcs: Log / Real values, Log loss + gradients, MaxTropical distances and the
  bit-exact Viterbi one-hot gradient against the oracle."""
  lt = _lt()
  b, t, ctx, u = 3, 11, 1, 5
  rng = np.random.RandomState(vocab + flags)
  c = 1 + vocab
  table_np = (rng.randn(b, t, c, 1 + vocab) * 2.0).astype(np.float32)
  drop = rng.rand(b, t, c, 1 + vocab) < 0.05
  drop[..., 0] = False
  table_np[drop] = -np.inf
  nf = np.array([11, 4, 9])
  labels = rng.randint(1, vocab + 1, size=(b, u))
  nl = np.array([5, 2, 3])
  octx = O.FullNGram(vocab, ctx)
  tab64 = table_np.astype(np.float64)
  blank, lex = np.ascontiguousarray(tab64[..., 0]), np.ascontiguousarray(tab64[..., 1:])
  with np.errstate(all='ignore'):
    o_loss, o_gb, o_gl = O.lattice_loss_and_grads(blank, lex, nf, labels, nl, octx)
    o_vd, o_vgb, o_vgl, _ = O.viterbi(table_np[..., 0].copy(), table_np[..., 1:].copy(), nf, octx,
                                      0, True)
  table = cuda(table_np).requires_grad_()
  lattice = make_lattice(vocab, ctx, -1, table, flags)
  loss = lattice(frames=frames_for(b, t), num_frames=cuda(nf), labels=cuda(labels),
                 num_labels=cuda(nl), cache=None)
  fin = np.isfinite(o_loss)
  npt.assert_array_equal(torch.isfinite(loss).cpu().numpy(), fin)
  npt.assert_allclose(loss.detach().cpu().numpy()[fin], o_loss[fin], rtol=1e-5, atol=1e-5)
  (gt,) = torch.autograd.grad(
      torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum(), table)
  gt = gt.cpu().numpy()
  assert np.all(np.isfinite(gt))
  assert np.all(gt[drop] == 0)
  npt.assert_allclose(gt[fin][..., 0], o_gb[fin], rtol=1e-4, atol=1e-5)
  npt.assert_allclose(gt[fin][..., 1:], o_gl[fin], rtol=1e-4, atol=1e-5)
  vd, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                           semiring=lt.semirings.MaxTropical)
  npt.assert_allclose(vd.detach().cpu(), o_vd, rtol=1e-6)
  (gv,) = torch.autograd.grad(vd.sum(), table)
  npt.assert_array_equal(gv.cpu().numpy()[..., 0], o_vgb)
  npt.assert_array_equal(gv.cpu().numpy()[..., 1:], o_vgl)
  # Real semiring on probabilities
  prob = (np.exp(np.clip(table_np, -40, 5) * 0.25) / (1 + vocab)).astype(np.float32)
  p64 = prob.astype(np.float64)
  o_rd, o_ra = O.lattice_forward(np.ascontiguousarray(p64[..., 0]),
                                 np.ascontiguousarray(p64[..., 1:]), nf, octx, O.REAL, 0, True)
  ptab = cuda(prob).requires_grad_()
  rl = make_lattice(vocab, ctx, -1, ptab, flags)
  rd, ra = rl._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                       semiring=lt.semirings.Real)
  npt.assert_allclose(rd.detach().cpu(), o_rd, rtol=1e-5)
  npt.assert_allclose(ra.cpu(), o_ra, rtol=2e-5, atol=1e-30)
  # d(dist)/d(weights) in the Real semiring against a generic-kernel run
  (gr,) = torch.autograd.grad(rd.sum(), ptab)
  ptab2 = cuda(prob).requires_grad_()
  rd2, _ = make_lattice(vocab, ctx, -1, ptab2, 1)._forward(
      cache=None, frames=frames_for(b, t), num_frames=cuda(nf), semiring=lt.semirings.Real)
  (gr2,) = torch.autograd.grad(rd2.sum(), ptab2)
  npt.assert_allclose(gr.cpu(), gr2.cpu(), rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize('flags', [0])
def test_fast_path_ties(flags):
  """All-equal weights on a fast-path shape: blank beats lexical, lowest source
  row wins inside the reduction (semirings.py:363, :382)."""
  b, t, vocab, ctx = 2, 6, 64, 1
  c = 1 + vocab
  tab = np.zeros([b, t, c, 1 + vocab], np.float32)
  tab[1, :, :, 0] = -1.0        # utterance 1: lexical arcs beat blank, all tie with each other
  nf = np.array([6, 4])
  o_vd, o_gb, o_gl, _ = O.viterbi(tab[..., 0].copy(), tab[..., 1:].copy(), nf,
                                  O.FullNGram(vocab, ctx), 0, True)
  table = cuda(tab).requires_grad_()
  lattice = make_lattice(vocab, ctx, -1, table, flags)
  dist, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                             semiring=_lt().semirings.MaxTropical)
  npt.assert_array_equal(dist.detach().cpu(), o_vd)
  (gd,) = torch.autograd.grad(dist.sum(), table)
  npt.assert_array_equal(gd.cpu().numpy()[..., 0], o_gb)
  npt.assert_array_equal(gd.cpu().numpy()[..., 1:], o_gl)


@pytest.mark.parametrize('vocab', [32, 64])
def test_rows_backward_vs_generic(vocab):
  """context_size 2, FrameDependent: the TMA column-forward / row-backward kernels against the
  generic kernels -- Log and Real distances, alphas and gradients, with -inf arcs, ragged and
  empty utterances and an upstream gradient that differs per utterance."""
  lt = _lt()
  b, t, ctx = 4, 7, 2
  c = 1 + vocab + vocab * vocab
  rng = np.random.RandomState(vocab)
  table_np = rng.randn(b, t, c, 1 + vocab).astype(np.float32)
  drop = rng.rand(b, t, c, 1 + vocab) < 0.03
  drop[..., 0] = False
  table_np[drop] = -np.inf
  nf = cuda(np.array([7, 3, 0, 6]))
  cot = cuda(np.array([1.0, -0.5, 2.0, 0.25]))
  frames = frames_for(b, t)
  for name in ['Log', 'Real']:
    sr = getattr(lt.semirings, name)
    tab = table_np if name == 'Log' else (np.exp(np.clip(table_np, -40, 5) * 0.25) /
                                          (1 + vocab)).astype(np.float32)
    res = []
    for flags in [1, 0]:
      table = cuda(tab).requires_grad_()
      lattice = make_lattice(vocab, ctx, -1, table, flags)
      dist, alphas = lattice._forward(cache=None, frames=frames, num_frames=nf, semiring=sr)
      (gd,) = torch.autograd.grad((dist * cot).sum(), table)
      res.append((dist.detach().cpu().numpy(), alphas.cpu().numpy(), gd.cpu().numpy()))
    (d0, a0, g0), (d1, a1, g1) = res
    npt.assert_allclose(d1, d0, rtol=2e-6, atol=1e-6, err_msg=name)
    fin = np.isfinite(a0)
    npt.assert_array_equal(np.isfinite(a1), fin)
    npt.assert_allclose(a1[fin], a0[fin], rtol=2e-6, atol=2e-5, err_msg=name)
    assert np.all(np.isfinite(g1))
    if name == 'Log':
      assert np.all(g1[drop] == 0)
    npt.assert_allclose(g1, g0, rtol=1e-4, atol=2e-6, err_msg=name)
    assert np.all(g1[2] == 0)                               # empty utterance
    assert np.all(g1[1, 3:] == 0)                           # padding frames


@pytest.mark.parametrize('k', [-1, 2])
def test_cols_path_ties(k):
  """All-equal weights on a thread-per-column shape (context_size 2): blank beats lexical,
  fewer expansions win, lowest source row block wins (semirings.py:363, :382)."""
  b, t, vocab, ctx = 2, 5, 16, 2
  c = 1 + vocab + vocab * vocab
  tab = np.zeros([b, t, c, 1 + vocab], np.float32)
  tab[1, :, :, 0] = -1.0        # utterance 1: lexical arcs beat blank, all tie with each other
  nf = np.array([5, 3])
  kk, fd = (0, True) if k < 0 else (k, False)
  o_vd, o_gb, o_gl, o_labels = O.viterbi(tab[..., 0].copy(), tab[..., 1:].copy(), nf,
                                         O.FullNGram(vocab, ctx), kk, fd)
  for flags in [0, 1]:
    table = cuda(tab).requires_grad_()
    lattice = make_lattice(vocab, ctx, k, table, flags)
    dist, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                               semiring=_lt().semirings.MaxTropical)
    npt.assert_array_equal(dist.detach().cpu(), o_vd)
    (gd,) = torch.autograd.grad(dist.sum(), table)
    npt.assert_array_equal(gd.cpu().numpy()[..., 0], o_gb)
    npt.assert_array_equal(gd.cpu().numpy()[..., 1:], o_gl)
    labels, _, _ = lattice.shortest_path(frames=frames_for(b, t), num_frames=cuda(nf), cache=None)
    npt.assert_array_equal(labels.cpu(), o_labels)


def test_viterbi_ties_match_reference_rules():
  """All-equal weights: blank beats lexical, so the all-blank path wins
  (semirings.py:363; SURVEY section 7 'MaxTropical ties')."""
  b, t, vocab = 2, 5, 3
  for ctx, k in [(1, -1), (2, -1), (1, 2)]:
    c = sum(vocab**i for i in range(ctx + 1))
    table = torch.zeros([b, t, c, 1 + vocab], device='cuda', requires_grad=True)
    lattice = make_lattice(vocab, ctx, k, table)
    labels, _, weights = lattice.shortest_path(frames=frames_for(b, t),
                                               num_frames=cuda([5, 3]), cache=None)
    npt.assert_array_equal(labels.cpu(), 0)
    npt.assert_array_equal(weights.cpu(), 0)
    kk, fd = (0, True) if k < 0 else (k, False)
    tab = np.zeros([b, t, c, 1 + vocab], np.float32)
    _, o_gb, o_gl, _ = O.viterbi(tab[..., 0].copy(), tab[..., 1:].copy(), np.array([5, 3]),
                                 O.FullNGram(vocab, ctx), kk, fd)
    dist, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda([5, 3]),
                               semiring=_lt().semirings.MaxTropical)
    (gd,) = torch.autograd.grad(dist.sum(), table)
    npt.assert_array_equal(gd.cpu().numpy()[..., 0], o_gb)
    npt.assert_array_equal(gd.cpu().numpy()[..., 1:], o_gl)


def test_neg_inf_weights_and_unreachable():
  """-inf arcs: values stay finite where a path exists, gradients are 0 on
  the -inf arcs (semirings.py:222-241); an unreachable label string gives
  loss=+inf (tests/lattices_test.py:57)."""
  rng = np.random.RandomState(0)
  b, t, vocab, ctx = 3, 7, 4, 1
  c = 1 + vocab
  tab = rng.randn(b, t, c, 1 + vocab).astype(np.float32)
  drop = rng.rand(b, t, c, 1 + vocab) < 0.3
  drop[..., 0] = False
  tab[drop] = -np.inf
  nf = np.array([7, 5, 2])
  labels = rng.randint(1, vocab + 1, size=(b, 4))
  nl = np.array([2, 1, 4])      # utterance 2: 4 labels in 2 frames is unreachable
  tab64 = tab.astype(np.float64)
  with np.errstate(all='ignore'):
    o_loss, o_gb, o_gl = O.lattice_loss_and_grads(
        tab64[..., 0].copy(), tab64[..., 1:].copy(), nf, labels, nl, O.FullNGram(vocab, ctx))
  table = cuda(tab).requires_grad_()
  lattice = make_lattice(vocab, ctx, -1, table)
  loss = lattice(frames=frames_for(b, t), num_frames=cuda(nf), labels=cuda(labels),
                 num_labels=cuda(nl), cache=None)
  l = loss.detach().cpu().numpy()
  assert np.isposinf(l[2])
  fin = np.isfinite(o_loss)
  npt.assert_allclose(l[fin], o_loss[fin], rtol=1e-5, atol=1e-5)
  (gt,) = torch.autograd.grad(loss[:2].sum(), table)
  gt = gt.cpu().numpy()
  assert np.all(np.isfinite(gt))
  assert np.all(gt[drop] == 0)
  npt.assert_allclose(gt[:2][..., 0], o_gb[:2], rtol=1e-4, atol=1e-5)
  npt.assert_allclose(gt[:2][..., 1:], o_gl[:2], rtol=1e-4, atol=1e-5)


def test_empty_and_zero_length():
  lt = _lt()
  vocab, ctx = 3, 1
  c = 4
  table = torch.randn([2, 4, c, 1 + vocab], device='cuda')
  lattice = make_lattice(vocab, ctx, -1, table)
  # num_frames == 0: distance is the semiring one (tests/lattices_test.py:210-221)
  for name, one in [('Log', 0.0), ('MaxTropical', 0.0), ('Real', 1.0)]:
    dist, alphas = lattice._forward(cache=None, frames=frames_for(2, 4), num_frames=cuda([0, 0]),
                                    semiring=getattr(lt.semirings, name))
    npt.assert_array_equal(dist.cpu(), [one, one])
    assert alphas.shape == (2, 4, c)
  loss = lattice(frames=frames_for(2, 4), num_frames=cuda([0, 0]), labels=cuda([[1], [2]]),
                 num_labels=cuda([0, 0]), cache=None)
  npt.assert_array_equal(loss.cpu(), [0, 0])


def test_full_size_properties_config1_and_2():
  """Size-independent properties at BASELINE.json sizes (configs[1]: B=32 is cut
  to B=4 to bound memory; T=1000, V=256): marginals of every real frame sum to
  one, padding frames get zero gradient, the gradient is invariant to padding,
  and chunked evaluation of the batch gives identical results."""
  lt = _lt()
  b, t, vocab, ctx = 4, 1000, 256, 1
  c = 257
  g = torch.Generator(device='cuda').manual_seed(0)
  table = torch.randn([b, t, c, 1 + vocab], device='cuda', generator=g).requires_grad_()
  nf = torch.tensor([1000, 731, 512, 1000], device='cuda')
  lattice = make_lattice(vocab, ctx, -1, table)
  dist, alphas = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=nf,
                                  semiring=lt.semirings.Log)
  (gd,) = torch.autograd.grad(dist.sum(), table)
  per_frame = gd.sum((2, 3))
  expect = (torch.arange(t, device='cuda')[None, :] < nf[:, None]).float()
  # logZ ~ 5.5e3 here, where one fp32 ulp is 4.9e-4: every posterior
  # exp(alpha + w + beta - logZ) inherits ~1e-3 relative error from the fp32
  # REPRESENTATION of alpha/beta/logZ (the fp32 reference has the same limit).
  npt.assert_allclose(per_frame.cpu(), expect.cpu(), rtol=0, atol=3e-3)
  assert float(gd.min()) >= 0
  # chunked over the batch == whole batch (utterances are independent)
  lattice2 = make_lattice(vocab, ctx, -1, table[2:].detach())
  dist2, _ = lattice2._forward(cache=None, frames=frames_for(2, t), num_frames=nf[2:],
                               semiring=lt.semirings.Log)
  npt.assert_array_equal(dist2.cpu(), dist[2:].detach().cpu())
  # logZ bounds: max-path <= logZ <= max-path + log(#paths) with #paths <= (V+1)^T
  vd, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=nf,
                           semiring=lt.semirings.MaxTropical)
  assert bool((vd <= dist + 1e-3).all())
  assert bool((dist <= vd + nf.float() * np.log(1 + vocab) + 1e-2).all())


@pytest.mark.parametrize('k', [-1, 2])
@pytest.mark.parametrize('vocab', [64, 128, 192, 256])
@pytest.mark.parametrize('semiring', ['log', 'real'])
def test_backward_split_row_emission(vocab, semiring, k):
  """lt_lattice_backward with LT_FLAG_GRAD_SPLIT (FrameDependent and FrameLabelDependent(2) fast
  paths): every row of grad_lexical holds [V bf16 hi | V bf16 lo] in its V*4 bytes and hi + lo
  is the fp32 gradient to 2^-17 -- with an odd batch, ragged and empty utterances (zero rows on
  padding frames) and a non-unit upstream gradient; grad_blank is unchanged."""
  _lt()
  from last_torch_b200 import ops, _native as N
  sr = N.LOG if semiring == 'log' else N.REAL
  b, t, c = 3, 11, vocab + 1
  g = torch.Generator(device='cuda').manual_seed(vocab)
  blank = torch.randn([b, t, c], device='cuda', generator=g)
  lex = torch.randn([b, t, c, vocab], device='cuda', generator=g)
  if semiring == 'real':
    blank, lex = blank * 0.01 + 1.0 / c, lex * 0.01 + 1.0 / c
  nf = torch.tensor([t, 6, 0], dtype=torch.int32, device='cuda')
  assert N.lib().lt_lattice_backward_split_supported(sr, vocab, 1, -1, 0) == 1
  assert N.lib().lt_lattice_backward_split_supported(sr, vocab, 2, -1, 0) == 0
  assert N.lib().lt_lattice_backward_split_supported(sr, vocab, 1, 2, 0) == 1
  assert N.lib().lt_lattice_backward_split_supported(sr, vocab, 1, 4, 0) == 0
  out = ops._lattice_forward_raw(sr, vocab, 1, k, blank, lex, nf, 0, True, False)
  dist, alphas, levels = out[0], out[1], out[3]
  gd = torch.tensor([1.0, -0.5, 2.0], device='cuda')

  def bwd(flags):
    gb = torch.full_like(blank, 7.0)
    gl = torch.full_like(lex, 7.0)
    N.check(N.lib().lt_lattice_backward(
        sr, vocab, 1, k, N.ptr(blank), N.ptr(lex), N.ptr(nf), b, t, N.ptr(alphas), N.ptr(levels),
        N.ptr(dist), N.ptr(gd), N.ptr(gb), N.ptr(gl), None, flags, N.stream_ptr(blank.device)),
        'lt_lattice_backward')
    return gb, gl

  gb0, gl0 = bwd(0)
  gb1, gs = bwd(N.FLAG_GRAD_SPLIT)
  npt.assert_array_equal(gb1.cpu().numpy(), gb0.cpu().numpy())
  rows = gs.view(torch.bfloat16).reshape(b, t, c, 2, vocab).float()
  rec = rows[..., 0, :] + rows[..., 1, :]
  err = (rec - gl0).abs()
  assert bool((err <= gl0.abs() * 2.0 ** -16 + 1e-37).all()), float(err.max())
  assert float(rec[1, 6:].abs().max()) == 0.0 and float(rec[2].abs().max()) == 0.0


@pytest.mark.parametrize('k', [1, 2, 3])
@pytest.mark.parametrize('vocab', [64, 192, 256])
def test_fast_path_frame_label_dependent(vocab, k):
  """FrameLabelDependent(k) on the bigram fast path (lattice_fast2_fld.cu: the frame's k levels
  against the resident tile): Log loss + gradients and Real values against the oracle, MaxTropical
  distance + bit-exact Viterbi path, and everything against the generic kernels
  (LT_FLD_GENERIC); odd batch, ragged lengths, -inf arcs."""
  from last_torch_b200 import _native as N
  lt = _lt()
  b, t, ctx, u = 3, 9, 1, 5
  rng = np.random.RandomState(vocab * 10 + k)
  c = 1 + vocab
  table_np = (rng.randn(b, t, c, 1 + vocab) * 2.0).astype(np.float32)
  drop = rng.rand(b, t, c, 1 + vocab) < 0.05
  drop[..., 0] = False
  table_np[drop] = -np.inf
  nf = np.array([9, 4, 7])
  labels = rng.randint(1, vocab + 1, size=(b, u))
  nl = np.array([5, 2, 3])
  octx = O.FullNGram(vocab, ctx)
  tab64 = table_np.astype(np.float64)
  blank, lex = np.ascontiguousarray(tab64[..., 0]), np.ascontiguousarray(tab64[..., 1:])
  with np.errstate(all='ignore'):
    o_loss, o_gb, o_gl = O.lattice_loss_and_grads(blank, lex, nf, labels, nl, octx, k, False)
    o_vd, o_vgb, o_vgl, o_labels = O.viterbi(table_np[..., 0].copy(), table_np[..., 1:].copy(),
                                             nf, octx, k, False)
    o_dist, o_alphas = O.lattice_forward(blank, lex, nf, octx, O.LOG, k, False)

  def run(generic):
    with N.option('LT_FLD_GENERIC', generic):
      table = cuda(table_np).requires_grad_()
      lattice = make_lattice(vocab, ctx, k, table)
      loss = lattice(frames=frames_for(b, t), num_frames=cuda(nf), labels=cuda(labels),
                     num_labels=cuda(nl), cache=None)
      (gt,) = torch.autograd.grad(
          torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum(), table)
      dist, alphas = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                                      semiring=lt.semirings.Log)
      vd, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                               semiring=lt.semirings.MaxTropical)
      (gv,) = torch.autograd.grad(vd.sum(), table)
      path, num, weights = lattice.shortest_path(frames=frames_for(b, t), num_frames=cuda(nf),
                                                 cache=None)
      return [x.detach().cpu().numpy() for x in (loss, gt, dist, alphas, vd, gv, path, weights)]

  fast, gen = run(0), run(1)
  loss, gt, dist, alphas, vd, gv, path, weights = fast
  fin = np.isfinite(o_loss)
  npt.assert_array_equal(np.isfinite(loss), fin)
  npt.assert_allclose(loss[fin], o_loss[fin], rtol=1e-5, atol=1e-5)
  assert np.all(np.isfinite(gt)) and np.all(gt[drop] == 0)
  npt.assert_allclose(gt[fin][..., 0], o_gb[fin], rtol=1e-4, atol=1e-5)
  npt.assert_allclose(gt[fin][..., 1:], o_gl[fin], rtol=1e-4, atol=1e-5)
  npt.assert_allclose(dist, o_dist, rtol=1e-5, atol=1e-5)
  afin = np.isfinite(o_alphas)
  npt.assert_array_equal(np.isfinite(alphas), afin)
  npt.assert_allclose(alphas[afin], o_alphas[afin], rtol=1e-5, atol=2e-4)
  npt.assert_allclose(vd, o_vd, rtol=1e-6)
  npt.assert_array_equal(gv[..., 0], o_vgb)
  npt.assert_array_equal(gv[..., 1:], o_vgl)
  npt.assert_array_equal(path, o_labels)
  npt.assert_allclose(weights, o_vd, rtol=1e-6)
  # the generic kernels on the same inputs
  npt.assert_allclose(loss[fin], gen[0][fin], rtol=2e-6, atol=2e-6)
  npt.assert_allclose(gt, gen[1], rtol=1e-4, atol=2e-6)
  npt.assert_array_equal(gv, gen[5])
  npt.assert_array_equal(path, gen[6])
  # Real semiring on probabilities: values against the oracle, gradient against the generic kernels
  prob = (np.exp(np.clip(table_np, -40, 5) * 0.25) / (1 + vocab)).astype(np.float32)
  p64 = prob.astype(np.float64)
  o_rd, o_ra = O.lattice_forward(np.ascontiguousarray(p64[..., 0]),
                                 np.ascontiguousarray(p64[..., 1:]), nf, octx, O.REAL, k, False)
  grads = []
  for generic in (0, 1):
    with N.option('LT_FLD_GENERIC', generic):
      ptab = cuda(prob).requires_grad_()
      rd, ra = make_lattice(vocab, ctx, k, ptab)._forward(
          cache=None, frames=frames_for(b, t), num_frames=cuda(nf), semiring=lt.semirings.Real)
      npt.assert_allclose(rd.detach().cpu(), o_rd, rtol=1e-5)
      npt.assert_allclose(ra.cpu(), o_ra, rtol=2e-5, atol=1e-30)
      grads.append(torch.autograd.grad(rd.sum(), ptab)[0].cpu().numpy())
  npt.assert_allclose(grads[0], grads[1], rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize('k', [1, 2])
def test_fast_path_frame_label_dependent_ties(k):
  """All-equal weights: blank terms with fewer expansions win, the lowest source row wins
  inside a level's reduction (semirings.py:363, :382) -- bit-exact against the oracle."""
  b, t, vocab, ctx = 2, 5, 64, 1
  c = 1 + vocab
  tab = np.zeros([b, t, c, 1 + vocab], np.float32)
  tab[1, :, :, 0] = -1.0
  tab[1, :, :, 1:] = 0.25       # utterance 1: expanding pays, every label ties
  nf = np.array([5, 3])
  o_vd, o_gb, o_gl, o_labels = O.viterbi(tab[..., 0].copy(), tab[..., 1:].copy(), nf,
                                         O.FullNGram(vocab, ctx), k, False)
  table = cuda(tab).requires_grad_()
  lattice = make_lattice(vocab, ctx, k, table)
  dist, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf),
                             semiring=_lt().semirings.MaxTropical)
  npt.assert_array_equal(dist.detach().cpu(), o_vd)
  (gd,) = torch.autograd.grad(dist.sum(), table)
  npt.assert_array_equal(gd.cpu().numpy()[..., 0], o_gb)
  npt.assert_array_equal(gd.cpu().numpy()[..., 1:], o_gl)
  path, _, _ = lattice.shortest_path(frames=frames_for(b, t), num_frames=cuda(nf), cache=None)
  npt.assert_array_equal(path.cpu(), o_labels)


def test_fast_path_frame_label_dependent_t300_against_the_double_oracle():
  """FrameLabelDependent(2), bigram vocab 256, T = 300, ragged: Log loss and ALL gradient entries
  against the double build of the C oracle (renormalised recursion, double-precision numerator
  chain: 1e-6 on the loss, 5e-5 relative + 1e-6 absolute on the gradients)."""
  from oracle import c_oracle
  b, t, v, k, u = 3, 300, 256, 2, 40
  rng = np.random.RandomState(21)
  gen = torch.Generator().manual_seed(212)
  table = torch.randn([b, t, v + 1, 1 + v], generator=gen)
  nf = np.array([300, 177, 251])
  labels = rng.randint(1, v + 1, size=(b, u))
  nl = np.array([40, 13, 0])
  tab = table.numpy()
  loss64, gb64, gl64, _, _ = c_oracle.lattice_loss_and_grads(
      np.ascontiguousarray(tab[..., 0]), np.ascontiguousarray(tab[..., 1:]), nf, labels, nl, v, 1,
      k, real='f64')
  leaf = table.cuda().requires_grad_()
  lattice = make_lattice(v, 1, k, leaf)
  loss = lattice(frames=frames_for(b, t), num_frames=cuda(nf), labels=cuda(labels),
                 num_labels=cuda(nl), cache=None)
  (gt,) = torch.autograd.grad(loss.sum(), leaf)
  npt.assert_allclose(loss.detach().cpu().numpy(), loss64, rtol=1e-6)
  gt = gt.cpu().numpy()
  # entries on the label string are differences of two O(1) posteriors (denominator minus
  # numerator): their error is a few fp32 ulps of ONE, whatever is left after the cancellation
  for got, want in ((gt[..., 0], gb64), (gt[..., 1:], gl64)):
    assert np.abs(got - want).max() < 1e-5
    npt.assert_allclose(got, want, rtol=5e-5, atol=1e-6)
