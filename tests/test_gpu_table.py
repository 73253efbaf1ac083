"""GPU parity tests of the table-driven kernels (contexts.NextStateTable,
csrc/lattice_table.cu): against the reference's known answers, against the
FullNGram kernels on the same DFA, and against the numpy oracle on random DFAs.
Tolerances as in test_gpu_lattice.py.
"""
import numpy as np
import numpy.testing as npt
import pytest
import torch

from oracle import lattice_oracle as O
from test_gpu_lattice import cuda, frames_for

pytestmark = pytest.mark.gpu


def _lt():
  import last_torch_b200 as last_torch
  return last_torch


def make_lattice(context, k, table, flags=0):
  lt = _lt()
  alignment = (lt.alignments.FrameDependent() if k < 0 else
               lt.alignments.FrameLabelDependent(max_expansions=k))
  lattice = lt.RecognitionLattice(
      context=context, alignment=alignment,
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  lattice.kernel_flags = flags
  return lattice


def test_forward_reduce_known_answer_and_semirings():
  """tests/contexts_test.py:214-220 (Real golden) + every semiring against the closed-form
  FullNGram reduction on the same DFA, values and gradients."""
  lt = _lt()
  full = lt.contexts.FullNGram(vocab_size=3, context_size=2)
  ctx = lt.contexts.NextStateTable(full.next_state_table().to(torch.int32))
  w = torch.arange(39, device='cuda', dtype=torch.float32).reshape(1, 13, 3)
  npt.assert_array_equal(
      ctx.forward_reduce(w, lt.semirings.Real).cpu(),
      [[0, 0, 1, 2, 3 * 4 + 54, 4 * 4 + 54, 5 * 4 + 54, 6 * 4 + 54, 7 * 4 + 54, 8 * 4 + 54,
        9 * 4 + 54, 10 * 4 + 54, 11 * 4 + 54]])
  g = torch.Generator(device='cuda').manual_seed(0)
  for name in ['Real', 'Log', 'MaxTropical']:
    sr = getattr(lt.semirings, name)
    x = torch.randn([2, 5, 13, 3], device='cuda', generator=g)
    x[0, 0, 4:7] = float('-inf') if name != 'Real' else 0.0
    a = x.clone().requires_grad_()
    b = x.clone().requires_grad_()
    ra, rb = ctx.forward_reduce(a, sr), full.forward_reduce(b, sr)
    fin = torch.isfinite(rb)
    npt.assert_array_equal(torch.isfinite(ra).cpu(), fin.cpu())
    npt.assert_allclose(ra[fin].detach().cpu(), rb[fin].detach().cpu(), rtol=1e-6, atol=1e-6)
    cot = torch.randn(ra.shape, device='cuda', generator=g)
    (ga,) = torch.autograd.grad((torch.where(fin, ra, torch.zeros_like(ra)) * cot).sum(), a)
    (gb,) = torch.autograd.grad((torch.where(fin, rb, torch.zeros_like(rb)) * cot).sum(), b)
    npt.assert_allclose(ga.cpu(), gb.cpu(), rtol=1e-5, atol=1e-6)
  # a state without incoming arcs gets the semiring zero
  t = torch.tensor([[1, 1], [1, 1]], dtype=torch.int32)
  c2 = lt.contexts.NextStateTable(t)
  x = torch.randn([2, 2], device='cuda')
  assert float(c2.forward_reduce(x, lt.semirings.Log)[0]) == float('-inf')
  assert float(c2.forward_reduce(x, lt.semirings.Real)[0]) == 0.0
  npt.assert_allclose(float(c2.forward_reduce(x, lt.semirings.Log)[1]),
                      float(torch.logsumexp(x.reshape(-1), 0)), rtol=1e-6)


@pytest.mark.parametrize('vocab,ctx_size,k', [(5, 1, -1), (3, 2, -1), (4, 2, 2), (6, 1, 3),
                                              (64, 1, -1)])
def test_table_lattice_equals_full_ngram(vocab, ctx_size, k):
  """A NextStateTable holding FullNGram's transitions is the same lattice: distances in all
  semirings, alphas, Log loss + gradients, Viterbi gradients and labels agree with the
  FullNGram kernels (which are pinned to the reference goldens and the oracle)."""
  lt = _lt()
  full = lt.contexts.FullNGram(vocab_size=vocab, context_size=ctx_size)
  tab_ctx = lt.contexts.NextStateTable(full.next_state_table().to(torch.int32))
  c = full.num_states()
  b, t, u = 3, 9, 4
  rng = np.random.RandomState(vocab * 10 + ctx_size)
  table_np = rng.randn(b, t, c, 1 + vocab).astype(np.float32)
  drop = rng.rand(b, t, c, 1 + vocab) < 0.05
  drop[..., 0] = False
  table_np[drop] = -np.inf
  nf = cuda(np.array([9, 5, 0]))
  labels = cuda(rng.randint(1, vocab + 1, size=(b, u)))
  nl = cuda(np.array([4, 2, 0]))
  frames = frames_for(b, t)
  for name in ['Real', 'Log', 'MaxTropical']:
    sr = getattr(lt.semirings, name)
    tab = table_np if name != 'Real' else (np.exp(np.clip(table_np, -40, 5) * 0.25) /
                                           (1 + vocab)).astype(np.float32)
    res = []
    for context, flags in [(full, 1), (tab_ctx, 0)]:
      table = cuda(tab).requires_grad_()
      lattice = make_lattice(context, k, table, flags)
      dist, alphas = lattice._forward(cache=None, frames=frames, num_frames=nf, semiring=sr)
      (gd,) = torch.autograd.grad(dist.sum(), table)
      res.append((dist.detach().cpu().numpy(), alphas.cpu().numpy(), gd.cpu().numpy()))
    (d0, a0, g0), (d1, a1, g1) = res
    npt.assert_allclose(d1, d0, rtol=2e-6, atol=1e-6, err_msg=name)
    fin = np.isfinite(a0)
    npt.assert_array_equal(np.isfinite(a1), fin)
    npt.assert_allclose(a1[fin], a0[fin], rtol=2e-6, atol=1e-5, err_msg=name)
    if name == 'MaxTropical':
      npt.assert_array_equal(g1, g0)
    else:
      npt.assert_allclose(g1, g0, rtol=1e-4, atol=1e-6, err_msg=name)
  res = []
  for context, flags in [(full, 1), (tab_ctx, 0)]:
    table = cuda(table_np).requires_grad_()
    lattice = make_lattice(context, k, table, flags)
    loss = lattice(frames=frames, num_frames=nf, labels=labels, num_labels=nl, cache=None)
    fin = torch.isfinite(loss)
    (gt,) = torch.autograd.grad(torch.where(fin, loss, torch.zeros_like(loss)).sum(), table)
    path = lattice.shortest_path(frames=frames, num_frames=nf, cache=None)
    res.append((loss.detach().cpu().numpy(), gt.cpu().numpy(), [x.cpu().numpy() for x in path]))
  (l0, g0, p0), (l1, g1, p1) = res
  npt.assert_array_equal(np.isfinite(l1), np.isfinite(l0))
  fin = np.isfinite(l0)
  npt.assert_allclose(l1[fin], l0[fin], rtol=1e-5, atol=1e-5)
  npt.assert_allclose(g1, g0, rtol=1e-4, atol=2e-6)
  for x, y in zip(p0, p1):
    npt.assert_array_equal(y, x)


@pytest.mark.parametrize('seed,c,vocab,k', [(0, 7, 3, -1), (1, 20, 5, -1), (2, 9, 4, 2),
                                            (3, 300, 17, -1)])
def test_random_dfa_vs_oracle(seed, c, vocab, k):
  """A random (non n-gram) DFA: some states have many incoming arcs, some none."""
  lt = _lt()
  rng = np.random.RandomState(seed)
  nst = rng.randint(0, max(2, c - 2), size=(c, vocab)).astype(np.int32)   # last states unreachable
  nst[0, 0] = 1
  ctx = lt.contexts.NextStateTable(torch.from_numpy(nst))
  octx = O.NextStateTable(nst)
  b, t, u = 3, 8, 4
  table_np = rng.randn(b, t, c, 1 + vocab).astype(np.float32)
  nf = np.array([8, 5, 2])
  labels = rng.randint(1, vocab + 1, size=(b, u))
  nl = np.array([3, 4, 1])
  kk, fd = (0, True) if k < 0 else (k, False)
  tab64 = table_np.astype(np.float64)
  blank, lex = np.ascontiguousarray(tab64[..., 0]), np.ascontiguousarray(tab64[..., 1:])
  with np.errstate(all='ignore'):
    o_loss, o_gb, o_gl = O.lattice_loss_and_grads(blank, lex, nf, labels, nl, octx, kk, fd)
  table = cuda(table_np).requires_grad_()
  lattice = make_lattice(ctx, k, table)
  frames = frames_for(b, t)
  loss = lattice(frames=frames, num_frames=cuda(nf), labels=cuda(labels), num_labels=cuda(nl),
                 cache=None)
  fin = np.isfinite(o_loss)
  npt.assert_array_equal(torch.isfinite(loss).cpu().numpy(), fin)
  npt.assert_allclose(loss.detach().cpu().numpy()[fin], o_loss[fin], rtol=1e-5, atol=1e-5)
  (gt,) = torch.autograd.grad(
      torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum(), table)
  gt = gt.cpu().numpy()
  npt.assert_allclose(gt[fin][..., 0], o_gb[fin], rtol=1e-4, atol=1e-5)
  npt.assert_allclose(gt[fin][..., 1:], o_gl[fin], rtol=1e-4, atol=1e-5)
  for name, sr in [('Log', O.LOG), ('MaxTropical', O.MAXTROPICAL)]:
    o_dist, o_alphas = O.lattice_forward(blank, lex, nf, octx, sr, kk, fd)
    dist, alphas = lattice._forward(cache=None, frames=frames, num_frames=cuda(nf),
                                    semiring=getattr(lt.semirings, name))
    npt.assert_allclose(dist.detach().cpu(), o_dist, rtol=1e-5, atol=1e-5, err_msg=name)
    a = alphas.cpu().numpy()
    f2 = np.isfinite(o_alphas)
    npt.assert_array_equal(np.isfinite(a), f2)
    npt.assert_allclose(a[f2], o_alphas[f2], rtol=1e-5, atol=2e-4, err_msg=name)
  # Viterbi: the returned path is a real path of the DFA and its score is the distance
  path_labels, num_path, weights = lattice.shortest_path(frames=frames, num_frames=cuda(nf),
                                                         cache=None)
  npt.assert_allclose(weights.cpu(), o_dist, rtol=1e-6)
  lab = path_labels.cpu().numpy().reshape(b, t, -1)
  for bi in range(b):
    q, score = 0, 0.0
    for ti in range(nf[bi]):
      took = 0
      for y in lab[bi, ti]:
        if y == 0:
          break
        score += table_np[bi, ti, q, y]
        q = nst[q, y - 1]
        took += 1
      if not fd or took == 0:      # FrameDependent: ONE arc per frame, blank or lexical
        score += table_np[bi, ti, q, 0]
    npt.assert_allclose(score, o_dist[bi], rtol=1e-5)
    assert np.all(lab[bi, nf[bi]:] == 0)


@pytest.mark.parametrize('cluster', [0, 1, 2, 4, 8])
@pytest.mark.parametrize('seed,c,vocab,ties', [(10, 20, 8, False), (11, 257, 12, False),
                                               (12, 130, 64, False), (13, 40, 16, True),
                                               (14, 257, 256, False)])
def test_cluster_table_kernels(seed, c, vocab, ties, cluster):
  """The cluster-per-utterance kernels (csrc/lattice_table2.cu) for every cluster size against
  the fp64 oracle (loss + gradients, Log / MaxTropical / Real forward) and, bit for bit in the
  MaxTropical semiring (integer weights = ties everywhere when `ties`), against the one-CTA
  kernels.  cluster 0 = the size the library picks itself."""
  lt = _lt()
  from last_torch_b200 import _native as N
  rng = np.random.RandomState(seed)
  nst = rng.randint(0, max(2, c - 2), size=(c, vocab)).astype(np.int32)
  nst[0, 0] = 1
  nst[:, vocab - 1] = c // 2          # one state with a large in-degree
  ctx = lt.contexts.NextStateTable(torch.from_numpy(nst))
  octx = O.NextStateTable(nst)
  b, t, u = 4, 11, 4
  if ties:
    table_np = rng.randint(-2, 3, size=(b, t, c, 1 + vocab)).astype(np.float32)
  else:
    table_np = rng.randn(b, t, c, 1 + vocab).astype(np.float32)
    drop = rng.rand(b, t, c, 1 + vocab) < 0.03
    drop[..., 0] = False
    table_np[drop] = -np.inf
  nf = np.array([11, 6, 0, 1])
  labels = rng.randint(1, vocab + 1, size=(b, u))
  nl = np.array([4, 3, 0, 1])
  frames = frames_for(b, t)

  def run():
    table = cuda(table_np).requires_grad_()
    lattice = make_lattice(ctx, -1, table)
    loss = lattice(frames=frames, num_frames=cuda(nf), labels=cuda(labels), num_labels=cuda(nl),
                   cache=None)
    fin = torch.isfinite(loss)
    (gt,) = torch.autograd.grad(torch.where(fin, loss, torch.zeros_like(loss)).sum(), table)
    out = {'loss': loss.detach().cpu().numpy(), 'grad': gt.cpu().numpy()}
    for name in ['Log', 'MaxTropical', 'Real']:
      tab = table
      if name == 'Real':
        tab = cuda((np.exp(np.clip(table_np, -40, 5) * 0.25) / (1 + vocab)).astype(np.float32))
        tab.requires_grad_()
      lat = make_lattice(ctx, -1, tab)
      dist, alphas = lat._forward(cache=None, frames=frames, num_frames=cuda(nf),
                                  semiring=getattr(lt.semirings, name))
      (gd,) = torch.autograd.grad(dist.sum(), tab)
      out[name] = (dist.detach().cpu().numpy(), alphas.cpu().numpy(), gd.cpu().numpy())
    path = lattice.shortest_path(frames=frames, num_frames=cuda(nf), cache=None)
    out['path'] = [x.cpu().numpy() for x in path]
    return out

  with N.option('LT_TABLE_V1', 1):
    assert N.lib().lt_table_lattice_cluster(c, vocab, -1, 0) == 0
    v1 = run()
  with N.option('LT_TABLE_CLUSTER', cluster):
    _cluster_body(c, vocab, cluster, run, v1, octx, table_np, nf, labels, nl, ties)


def _cluster_body(c, vocab, cluster, run, v1, octx, table_np, nf, labels, nl, ties):
  lt = _lt()
  from last_torch_b200 import _native as N
  used = N.lib().lt_table_lattice_cluster(c, vocab, -1, 0)
  used_bwd = N.lib().lt_table_lattice_cluster(c, vocab, -1, 1)
  if cluster:
    assert used in (0, cluster) and used_bwd in (0, cluster)   # 0: shape outside its limits
    if (c, vocab, cluster) in [(257, 256, 8), (20, 8, 2), (130, 64, 4), (257, 12, 8)]:
      assert used == cluster and used_bwd == cluster
  else:
    assert used >= 1 and used_bwd >= 1
    if (c, vocab) == (257, 256):
      assert used == 8 and used_bwd == 8
  v2 = run()

  # against the one-CTA kernels
  npt.assert_array_equal(np.isfinite(v2['loss']), np.isfinite(v1['loss']))
  fin = np.isfinite(v1['loss'])
  npt.assert_allclose(v2['loss'][fin], v1['loss'][fin], rtol=1e-5, atol=1e-5)
  npt.assert_allclose(v2['grad'][fin], v1['grad'][fin], rtol=1e-4, atol=2e-6)
  for name in ['Log', 'MaxTropical', 'Real']:
    d1, a1, g1 = v1[name]
    d2, a2, g2 = v2[name]
    f = np.isfinite(a1)
    npt.assert_array_equal(np.isfinite(a2), f, err_msg=name)
    if name == 'MaxTropical':
      npt.assert_array_equal(d2, d1)
      npt.assert_array_equal(a2[f], a1[f])
      npt.assert_array_equal(g2, g1)          # same winning arcs, ties included
    else:
      npt.assert_allclose(d2, d1, rtol=1e-5, atol=1e-6, err_msg=name)
      npt.assert_allclose(a2[f], a1[f], rtol=1e-5, atol=1e-5, err_msg=name)
      npt.assert_allclose(g2, g1, rtol=1e-4, atol=2e-6, err_msg=name)
  for x, y in zip(v1['path'], v2['path']):
    npt.assert_array_equal(y, x)

  # against the fp64 oracle
  tab64 = table_np.astype(np.float64)
  blank, lex = np.ascontiguousarray(tab64[..., 0]), np.ascontiguousarray(tab64[..., 1:])
  with np.errstate(all='ignore'):
    o_loss, o_gb, o_gl = O.lattice_loss_and_grads(blank, lex, nf, labels, nl, octx, 0, True)
  ofin = np.isfinite(o_loss)
  npt.assert_array_equal(np.isfinite(v2['loss']), ofin)
  npt.assert_allclose(v2['loss'][ofin], o_loss[ofin], rtol=1e-5, atol=1e-5)
  npt.assert_allclose(v2['grad'][ofin][..., 0], o_gb[ofin], rtol=1e-4, atol=1e-5)
  npt.assert_allclose(v2['grad'][ofin][..., 1:], o_gl[ofin], rtol=1e-4, atol=1e-5)
  for name, sr in [('Log', O.LOG), ('MaxTropical', O.MAXTROPICAL)]:
    with np.errstate(all='ignore'):
      o_dist, o_alphas = O.lattice_forward(blank, lex, nf, octx, sr, 0, True)
    d2, a2, _ = v2[name]
    npt.assert_allclose(d2, o_dist, rtol=1e-5, atol=1e-5, err_msg=name)
    f2 = np.isfinite(o_alphas)
    npt.assert_array_equal(np.isfinite(a2), f2)
    npt.assert_allclose(a2[f2], o_alphas[f2], rtol=1e-5, atol=2e-4, err_msg=name)
