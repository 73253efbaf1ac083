"""GPU tests of the public API surface: semirings, contexts, alignments,
weight functions and RecognitionLattice, written to read like the reference's
own tests (file:line cited) plus the golden fixtures produced by the reference.
"""
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

from conftest import GOLDEN_DIR, golden_files

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _cuda_default_device():
  prev = torch.get_default_device()
  torch.set_default_device('cuda')
  yield
  torch.set_default_device(prev)


def _lt():
  import last_torch_b200 as last_torch
  return last_torch


def T(x):
  return torch.tensor(x, dtype=torch.float32)


# ---- semirings (tests/semirings_test.py) -----------------------------------

@pytest.mark.parametrize('name', ['Real', 'Log', 'MaxTropical'])
def test_zero_and_one(name):
  # tests/semirings_test.py:25-63
  semiring = getattr(_lt().semirings, name)
  one, zero = semiring.ones([3]), semiring.zeros([3])
  xs = T([1., 2., 3.])
  for args in [(one, xs), (xs, one)]:
    npt.assert_array_equal(semiring.times(*args).cpu(), xs.cpu())
    npt.assert_array_equal(semiring.prod(torch.stack(args), dim=0).cpu(), xs.cpu())
  for a, b, exp in [('ones', 'zeros', 'zeros'), ('zeros', 'ones', 'zeros'),
                    ('ones', 'ones', 'ones'), ('zeros', 'zeros', 'zeros')]:
    npt.assert_array_equal(
        semiring.times(getattr(semiring, a)((1, 2)), getattr(semiring, b)((3, 1))).cpu(),
        getattr(semiring, exp)((3, 2)).cpu())
  for a, b, exp in [('ones', 'zeros', 'ones'), ('zeros', 'ones', 'ones'),
                    ('zeros', 'zeros', 'zeros')]:
    npt.assert_array_equal(
        semiring.plus(getattr(semiring, a)((1, 2)), getattr(semiring, b)((3, 1))).cpu(),
        getattr(semiring, exp)((3, 2)).cpu())
  npt.assert_array_equal(semiring.sum(torch.zeros([3, 0]), dim=0).cpu(), np.zeros([0]))
  npt.assert_array_equal(semiring.sum(torch.zeros([3, 0]), dim=1).cpu(), zero.cpu())
  npt.assert_array_equal(semiring.prod(torch.zeros([3, 0]), dim=1).cpu(), one.cpu())


def test_log_basics_and_grads():
  # tests/semirings_test.py:194-204
  log = _lt().semirings.Log
  npt.assert_allclose(log.plus(T([2]), T([3])).cpu(), 3.31326169, rtol=1e-6)
  npt.assert_allclose(log.sum(T([2, 3]), dim=0).cpu(), 3.31326169, rtol=1e-6)
  # gradients: the intent documented at semirings.py:222-241
  a = T([[1., 2., -float('inf')], [0., 2., -float('inf')]]).requires_grad_()
  y = log.plus(a[0], a[1])
  (g,) = torch.autograd.grad(y.sum(), a)
  e = torch.softmax(torch.tensor([[1., 0.], [2., 2.]]), dim=-1).T
  npt.assert_allclose(g.cpu()[:, :2], e.cpu(), rtol=1e-6)
  npt.assert_array_equal(g.cpu()[:, 2], [0, 0])         # all -inf: zero gradient, no NaN
  a = torch.randn([2, 3, 4, 5]).requires_grad_()
  for dim in range(-4, 4):
    y = log.sum(a, dim=dim)
    npt.assert_allclose(y.detach().cpu(), torch.logsumexp(a.detach(), dim=dim).cpu(), rtol=1e-5,
                        atol=1e-6)
    (g,) = torch.autograd.grad(y.sum(), a)
    npt.assert_allclose(g.cpu(), torch.softmax(a.detach(), dim=dim).cpu(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('name', ['Log', 'MaxTropical'])
def test_sum_axis_errors_and_zero_sized(name):
  # tests/semirings_test.py:148-189
  semiring = getattr(_lt().semirings, name)
  xs = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape([2, 3, 4, 5])
  assert semiring.sum(xs, dim=1).shape == (2, 4, 5)
  assert semiring.sum(xs, dim=-1).shape == (2, 3, 4)
  with pytest.raises(ValueError, match='Invalid reduction axis'):
    semiring.sum(xs, dim=4)
  with pytest.raises(ValueError, match='Invalid reduction axis'):
    semiring.sum(xs, dim=-5)
  with pytest.raises(ValueError, match='Only int axis'):
    semiring.sum(xs, dim=None)
  z = torch.zeros([0, 2])
  npt.assert_array_equal(semiring.sum(z, dim=0).cpu(), semiring.zeros([2]).cpu())
  assert semiring.sum(z, dim=1).shape == (0,)


def test_maxtropical_tie_gradients():
  # tests/semirings_test.py:226-247
  mt = _lt().semirings.MaxTropical
  a = T([[1., 2., 3.], [0., 2., 4.]]).requires_grad_()
  (g,) = torch.autograd.grad(mt.plus(a[0], a[1]).sum(), a)
  npt.assert_array_equal(g.cpu(), [[1, 1, 0], [0, 0, 1]])
  (g,) = torch.autograd.grad(mt.sum(a, dim=0).sum(), a)
  npt.assert_array_equal(g.cpu(), [[1, 1, 0], [0, 0, 1]])
  at = T([[1., 2., 3.], [0., 2., 4.]]).requires_grad_()
  (g,) = torch.autograd.grad(mt.sum(at.T, dim=-1).sum(), at)
  npt.assert_array_equal(g.cpu(), [[1, 1, 0], [0, 0, 1]])


def test_cartesian_and_expectation():
  # tests/semirings_test.py:305-388 (op-level composition)
  s = _lt().semirings
  cart = s.Cartesian(s.Real, s.MaxTropical)
  a = (T([2.]), T([3.]))
  b = (T([4.]), T([5.]))
  x, y = cart.plus(a, b)
  npt.assert_array_equal(x.cpu(), [6]); npt.assert_array_equal(y.cpu(), [5])
  x, y = cart.times(a, b)
  npt.assert_array_equal(x.cpu(), [8]); npt.assert_array_equal(y.cpu(), [8])
  x, y = cart.sum((T([1., 2.]), T([1., 2.])), 0)
  npt.assert_array_equal(x.cpu(), 3); npt.assert_array_equal(y.cpu(), 2)
  # entropy via the expectation semiring (tests/semirings_test.py:305-324)
  probs = T([0.25, 0.25, 0.5])
  w = torch.log(probs)
  ent = s.LogLogExpectation.sum(s.LogLogExpectation.weighted(w, torch.log(-w)), 0)
  npt.assert_allclose(torch.exp(ent[1]).cpu(), float(-(probs * torch.log(probs)).sum()), rtol=1e-5)


# ---- contexts / alignments (golden per-frame ops from the reference) ----------

@pytest.mark.parametrize('fname', golden_files('frameops_'))
def test_frame_ops_golden(fname):
  lt = _lt()
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  k = int(g['k'])
  context = lt.contexts.FullNGram(int(g['vocab']), int(g['context_size']))
  alignment = (lt.alignments.FrameDependent() if k < 0 else
               lt.alignments.FrameLabelDependent(max_expansions=k))
  n = alignment.num_states()
  alpha, blank, lexical, beta, log_z = (T(g[x]) for x in ['alpha', 'blank', 'lexical', 'beta',
                                                          'log_z'])
  for name in ['Real', 'Log', 'MaxTropical']:
    semiring = getattr(lt.semirings, name)
    npt.assert_allclose(
        alignment.forward(alpha, [blank] * n, [lexical] * n, context, semiring).cpu(),
        g[f'{name}_forward'], rtol=1e-5, atol=1e-6)
    npt.assert_allclose(
        alignment.string_forward(T(g['salpha']), [T(g['sblank'])] * n, [T(g['slex'])] * n,
                                 semiring).cpu(),
        g[f'{name}_string_forward'], rtol=1e-5, atol=1e-6)
    npt.assert_allclose(context.forward_reduce(lexical, semiring).cpu(),
                        g[f'{name}_forward_reduce'], rtol=1e-5, atol=1e-6)
  nb, bm, lm = alignment.backward(alpha, [blank] * n, [lexical] * n, beta, log_z, context)
  npt.assert_allclose(nb.cpu(), g['backward_next_beta'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(torch.stack(bm).sum(0).cpu(), g['backward_blank_marginal'], rtol=1e-5,
                      atol=1e-7)
  npt.assert_allclose(torch.stack(lm).sum(0).cpu(), g['backward_lexical_marginal'], rtol=1e-5,
                      atol=1e-7)
  npt.assert_array_equal(context.backward_broadcast(beta).cpu(), g['backward_broadcast'])
  npt.assert_array_equal(context.next_state_table().cpu(), g['next_state_table'])


def test_alignment_arity_errors():
  # tests/alignments_test.py:77-91
  lt = _lt()
  context = lt.contexts.FullNGram(2, 1)
  al = lt.alignments.FrameDependent()
  a, b, l = torch.rand([3]), torch.rand([3]), torch.rand([3, 2])
  with pytest.raises(ValueError, match='blank should be'):
    al.forward(a, [b, b], [l], context, lt.semirings.Real)
  with pytest.raises(ValueError, match='lexical should be'):
    al.forward(a, [b], [l, l], context, lt.semirings.Real)


# ---- JointWeightFn (golden from the reference body with injected weights) ------

@pytest.mark.parametrize('fname', golden_files('joint_'))
def test_joint_golden(fname):
  lt = _lt()
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  v = int(g['vocab'])
  h, e = g['w_ctx'].shape
  d = g['w_frame'].shape[1]
  fn = lt.weight_fns.JointWeightFn(vocab_size=v, hidden_size=h, device='cuda', embedding_size=e,
                                   feature_size=d)
  with torch.no_grad():
    fn.context_projection.weight.copy_(T(g['w_ctx']))
    fn.blank_projection.weight.copy_(T(g['w_frame']))
    fn.joint_projection_to_blank.weight.copy_(T(g['w_blank'])[None])
    fn.joint_projection_to_blank.bias.copy_(T(g['b_blank']).reshape(1))
    fn.joint_projection_to_vocab.weight.copy_(T(g['w_vocab']))
    fn.joint_projection_to_vocab.bias.copy_(T(g['b_vocab']))
  cache, frame = T(g['cache']), T(g['frame'])
  blank, lexical = fn(cache, frame)
  npt.assert_allclose(blank.detach().cpu(), g['blank'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(lexical.detach().cpu(), g['lexical'], rtol=1e-5, atol=1e-6)
  sb, sl = fn(cache, frame, torch.tensor(g['state']))
  npt.assert_allclose(sb.detach().cpu(), g['state_blank'], rtol=1e-5, atol=1e-6)
  npt.assert_allclose(sl.detach().cpu(), g['state_lexical'], rtol=1e-5, atol=1e-6)
  # the whole-utterance kernel path gives the same numbers (and gradients)
  frames = frame[:, None, :].expand(-1, 3, -1).contiguous()
  kb, kl = fn.all_frames(cache, frames)
  npt.assert_allclose(kb[:, 0].detach().cpu(), g['blank'], rtol=1e-5, atol=2e-6)
  npt.assert_allclose(kl[:, 1].detach().cpu(), g['lexical'], rtol=1e-5, atol=2e-6)
  params = list(fn.parameters())
  cache_r = cache.clone().requires_grad_()
  frames_r = frames.clone().requires_grad_()
  wb = torch.randn_like(kb); wl = torch.randn_like(kl)
  kb, kl = fn.all_frames(cache_r, frames_r)
  got = torch.autograd.grad((kb * wb).sum() + (kl * wl).sum(), params + [cache_r, frames_r])
  rb, rl = fn(cache_r, frames_r.reshape(-1, d))
  rb, rl = rb.reshape(kb.shape), rl.reshape(kl.shape)
  ref = torch.autograd.grad((rb * wb).sum() + (rl * wl).sum(), params + [cache_r, frames_r])
  for a, b in zip(got, ref):
    npt.assert_allclose(a.cpu(), b.cpu(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('fname', golden_files('jointlattice_'))
@pytest.mark.parametrize('split', [True, False])
def test_joint_lattice_golden(fname, split):
  """The reference's whole GNAT loss with JointWeightFn inside the lattice -- vocab 128 / hidden
  128 (tcgen05 forward, TMA fast path, tensor-core backward with and without the split-row
  hand-over), vocab 64 (tcgen05 forward and dgrad, CUDA-core weight gradient) and a trigram
  FrameLabelDependent(2) lattice (CUDA-core joint, generic lattice kernels): loss of the
  reference as shipped, parameter gradients of its patched autograd."""
  import os
  lt = _lt()
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  v, h = int(g['vocab']), int(g['hidden'])
  e, d = g['cache'].shape[1], g['frames'].shape[2]
  k = int(g['k'])
  context = lt.contexts.FullNGram(vocab_size=v, context_size=int(g['context_size']))
  lattice = lt.RecognitionLattice(
      context=context,
      alignment=(lt.alignments.FrameDependent() if k < 0 else
                 lt.alignments.FrameLabelDependent(max_expansions=k)),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=e, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=v, hidden_size=h, device='cuda', embedding_size=e, feature_size=d))
  fn, cacher = lattice.weight_fn, lattice.weight_fn_cacher
  with torch.no_grad():
    cacher.embedding.weight.copy_(T(g['cache']))
    fn.context_projection.weight.copy_(T(g['w_ctx']))
    fn.blank_projection.weight.copy_(T(g['w_frame']))
    fn.joint_projection_to_blank.weight.copy_(T(g['w_blank']))
    fn.joint_projection_to_blank.bias.copy_(T(g['b_blank']))
    fn.joint_projection_to_vocab.weight.copy_(T(g['w_vocab']))
    fn.joint_projection_to_vocab.bias.copy_(T(g['b_vocab']))
  lattice.split_grad_handover = split
  loss = lattice(frames=T(g['frames']), num_frames=T(g['num_frames']), labels=T(g['labels']),
                 num_labels=T(g['num_labels']))
  loss.sum().backward()
  npt.assert_allclose(loss.detach().cpu(), g['loss'], rtol=1e-5)
  got = {'cache': cacher.embedding.weight.grad, 'w_ctx': fn.context_projection.weight.grad,
         'w_frame': fn.blank_projection.weight.grad,
         'w_blank': fn.joint_projection_to_blank.weight.grad,
         'b_blank': fn.joint_projection_to_blank.bias.grad,
         'w_vocab': fn.joint_projection_to_vocab.weight.grad,
         'b_vocab': fn.joint_projection_to_vocab.bias.grad}
  for name, a in got.items():
    ref = g['grad_' + name]
    scale = float(np.abs(ref).max())
    err = float(np.abs(a.cpu().numpy() - ref).max())
    # + 5e-6: grad_b_blank of the FrameLabelDependent fixture is 0 up to round-off (the
    # reference's own value there is 1.7e-6 of noise)
    assert err <= 1e-4 * scale + 5e-6, (name, err / scale)


def test_joint_lattice_full_size_properties():
  """BASELINE configs[1] + configs[3] end to end (B=32, T=1000, bigram vocab 256, U=120, joint
  hidden 512) through RecognitionLattice.forward: size-independent properties of the parameter
  gradients, with and without the split-row hand-over.
    * adding a constant to every arc weight of a FrameDependent lattice moves logZ and the
      numerator by the same T * const, so sum_v grad_b_vocab + grad_b_blank = 0;
    * both hand-over forms give the same loss and the same gradients."""
  import os
  lt = _lt()
  b, t, v, h, u = 32, 1000, 256, 512, 120
  torch.manual_seed(0)
  context = lt.contexts.FullNGram(vocab_size=v, context_size=1)
  lattice = lt.RecognitionLattice(
      context=context, alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=64, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=h, device='cuda'))
  x = torch.randn([b, t, 80], device='cuda')
  num_frames = torch.randint(t // 2, t + 1, [b], device='cuda')
  labels = torch.randint(1, v + 1, [b, u], device='cuda')
  num_labels = torch.full([b], u, device='cuda')

  def run(no_split):
    lattice.split_grad_handover = not no_split
    lattice.zero_grad()
    loss = lattice(frames=x, num_frames=num_frames, labels=labels, num_labels=num_labels)
    loss.sum().backward()
    return loss.detach().clone(), {n: p.grad.clone() for n, p in lattice.named_parameters()}

  loss_s, g_s = run(False)
  loss_f, g_f = run(True)
  assert bool(torch.isfinite(loss_s).all())
  npt.assert_array_equal(loss_s.cpu(), loss_f.cpu())
  mass = float(num_frames.sum())
  for grads in (g_s, g_f):
    bv = [g for n, g in grads.items() if n.endswith('joint_projection_to_vocab.bias')][0]
    bb = [g for n, g in grads.items() if n.endswith('joint_projection_to_blank.bias')][0]
    assert float(bv.abs().sum()) > 0.1 * mass          # the gradient is not trivially small
    assert abs(float(bv.double().sum() + bb.double().sum())) < 5e-3 * mass
  for n in g_s:
    scale = float(g_f[n].abs().max()) + 1e-12
    assert float((g_s[n] - g_f[n]).abs().max()) / scale < 1e-4, n


# ---- RecognitionLattice API (tests/lattices_test.py) ---------------------------

def _joint_lattice(vocab_size, context_size, alignment):
  lt = _lt()
  context = lt.contexts.FullNGram(vocab_size=vocab_size, context_size=context_size)
  return lt.RecognitionLattice(
      context=context, alignment=alignment,
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=24, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=16, device='cuda'))


def test_call_joint_weight_fn():
  # tests/lattices_test.py:39-89
  lt = _lt()
  lattice = _joint_lattice(2, 1, lt.alignments.FrameDependent())
  frames = torch.rand([4, 6, 8])
  num_frames = T([6, 3, 2, 1])
  labels = T([[1, 1, 1, 1], [2, 2, 2, 2], [1, 2, 1, 2], [2, 1, 2, 1]])
  num_labels = T([4, 3, 1, 2])
  loss = lattice(frames=frames, num_frames=num_frames, labels=labels, num_labels=num_labels)
  npt.assert_array_equal(torch.isfinite(loss).cpu(), [True, True, True, False])
  # padded inputs give the SAME loss (the reference can only check rtol=2 because its
  # weights are re-randomised on every call)
  padded = lattice(frames=torch.nn.functional.pad(frames, (0, 0, 0, 1, 0, 0)),
                   num_frames=num_frames,
                   labels=torch.nn.functional.pad(labels, (0, 2, 0, 0)), num_labels=num_labels)
  npt.assert_allclose(padded.detach().cpu()[:3], loss.detach().cpu()[:3], rtol=1e-6)
  with pytest.raises(ValueError, match='frames and num_frames have different batch_dims'):
    lattice(frames=frames[:1], num_frames=num_frames, labels=labels, num_labels=num_labels)
  with pytest.raises(ValueError, match='labels and num_frames have different batch_dims'):
    lattice(frames=frames, num_frames=num_frames, labels=labels[:1], num_labels=num_labels)
  with pytest.raises(ValueError, match='num_labels and num_frames have different batch_dims'):
    lattice(frames=frames, num_frames=num_frames, labels=labels, num_labels=num_labels[:1])
  # gradients reach every parameter of the weight function and the cacher
  loss[:3].sum().backward()
  grads = [p.grad for p in lattice.parameters()]
  assert len(grads) == 7 and all(g is not None and torch.isfinite(g).all() for g in grads)
  assert all(float(g.abs().sum()) > 0 for g in grads)


@pytest.mark.parametrize('vocab,hidden,batch,frames', [(128, 128, 3, 40), (256, 256, 2, 70)])
def test_joint_lattice_split_row_gradients(vocab, hidden, batch, frames):
  """RecognitionLattice.forward with JointWeightFn: the lattice backward kernel hands its arc
  posteriors to the tensor-core joint backward as split rows ([V bf16 hi | V bf16 lo], read by
  TMA; ops.JointLatticeLoss).  Same parameter gradients as the fp32 hand-over
  (split_grad_handover = False),
  with ragged utterances (zero rows on padding frames) and label strings with repeated bigrams
  (several numerator positions land on one arc of the split buffer)."""
  import os
  lt = _lt()
  from last_torch_b200 import ops
  torch.manual_seed(vocab + batch)
  context = lt.contexts.FullNGram(vocab_size=vocab, context_size=1)
  lattice = lt.RecognitionLattice(
      context=context, alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=24, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=hidden, device='cuda'))
  x = torch.randn([batch, frames, 16], device='cuda')
  num_frames = torch.tensor([frames, frames - 7, frames // 2][:batch], device='cuda')
  labels = torch.randint(1, vocab + 1, [batch, 12], device='cuda')
  labels[:, 4:8] = labels[:, 0:4]            # repeated bigrams
  labels[0, 8:] = labels[0, 8]               # ... and a run of one label
  num_labels = torch.tensor([12, 9, 5][:batch], device='cuda')
  weights = torch.rand([batch], device='cuda') + 0.5

  def run(no_split):
    lattice.split_grad_handover = not no_split
    lattice.zero_grad()
    loss = lattice(frames=x, num_frames=num_frames, labels=labels, num_labels=num_labels)
    (loss * weights).sum().backward()
    return loss.detach(), [p.grad.clone() for p in lattice.parameters()]

  loss_s, g_s = run(False)
  loss_f, g_f = run(True)
  npt.assert_array_equal(loss_s.cpu(), loss_f.cpu())
  assert len(g_s) == 7
  for a, b in zip(g_s, g_f):
    scale = float(b.abs().max()) + 1e-12
    assert float((a - b).abs().max()) / scale < 2e-5, (tuple(b.shape), float((a - b).abs().max()) / scale)
  # the split path really ran: the handshake object reports it for this shape
  from last_torch_b200 import _native as N
  c = vocab + 1
  assert N.lib().lt_joint_backward_split_supported(batch * frames, c, hidden, vocab) == 1
  assert N.lib().lt_lattice_backward_split_supported(N.LOG, vocab, 1, -1, 0) == 1


def test_shortest_path_api():
  # tests/lattices_test.py:91-127 and :151-176
  lt = _lt()
  frames = torch.rand([4, 6, 8])
  lattice = _joint_lattice(2, 1, lt.alignments.FrameDependent())
  nf = T([6, 3, 2, 0])
  labels, num, weights = lattice.shortest_path(frames, nf)
  npt.assert_array_equal(num.cpu(), [6, 3, 2, 0])
  is_padding = (torch.arange(6)[None, :] >= nf[:, None])
  assert bool((labels[is_padding] == 0).all())
  assert bool(((labels >= 0) & (labels <= 2)).all())
  npt.assert_array_equal(torch.isfinite(weights).cpu(), [True] * 4)
  npt.assert_array_equal((weights == 0).cpu(), [False, False, False, True])
  lattice = _joint_lattice(2, 1, lt.alignments.FrameLabelDependent(max_expansions=2))
  nf = T([6, 3, 2, 1])
  labels, num, weights = lattice.shortest_path(frames, nf)
  npt.assert_array_equal(num.cpu(), (3 * nf).cpu())
  npt.assert_array_equal(labels.reshape([4, 6, 3])[..., -1].cpu(), np.zeros([4, 6]))
  loss = lattice(frames=frames, num_frames=nf,
                 labels=T([[1, 1, 1, 1], [2, 2, 2, 2], [1, 2, 1, 2], [2, 1, 2, 1]]),
                 num_labels=T([4, 3, 4, 3]))
  npt.assert_array_equal(torch.isfinite(loss).cpu(), [True, True, True, False])


def test_frame_dependent_known_answer():
  # tests/lattices_test.py:181-288 (TableWeightFn golden)
  lt = _lt()
  b, t, v, c = 3, 2, 2, 3
  frames = torch.arange(t)[None, :, None].expand(b, t, 1).float()
  num_frames = T([2, 1, 0])
  table = 1 + torch.arange(b * t * c * (1 + v)).reshape([b, t, c, 1 + v]).float()
  table = table * T([[-1, 1], [1, -1], [1, 1]])[:, :, None, None]
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=v, context_size=1),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  lse = lambda xs: float(torch.logsumexp(T(xs), 0))
  den = lse([-1 + 10, -1 + 11, -1 + 12, -2 + 13, -2 + 14, -2 + 15, -3 + 16, -3 + 17, -3 + 18])
  for name, expected in [('MaxTropical', [15, 21, 0]), ('Real', [-33 - 84 - 153, 60, 1]),
                         ('Log', [den, lse([19, 20, 21]), 0.])]:
    got = lattice._forward(cache=None, frames=frames, num_frames=num_frames,
                           semiring=getattr(lt.semirings, name))[0]
    npt.assert_allclose(got.cpu(), expected, rtol=1e-6)
  labels_out, num_out, weights = lattice.shortest_path(frames=frames, num_frames=num_frames,
                                                       cache=None)
  npt.assert_array_equal(num_out.cpu(), num_frames.cpu())
  npt.assert_allclose(weights.cpu(), [15, 21, 0])
  # true labels; the reference prints [[1,1],[0,0],[0,0]] because of SURVEY D4/D5
  npt.assert_array_equal(labels_out.cpu(), [[2, 2], [2, 0], [0, 0]])
  labels = T([[1, 2, 0], [2, 1, 0], [1, 2, 0]])
  num_labels = T([1, 1, 0])
  for name, expected in [('MaxTropical', [11, 21, 0]), ('Real', [-11 - 26, 21, 1]),
                         ('Log', [lse([10, 11]), 21, 0])]:
    semiring = getattr(lt.semirings, name)
    got = lattice._string_forward(cache=None, frames=frames, num_frames=num_frames, labels=labels,
                                  num_labels=num_labels, semiring=semiring)
    npt.assert_allclose(got.cpu(), expected, rtol=1e-6)
    got = lattice._string_forward(cache=None, frames=frames, num_frames=num_frames, labels=labels,
                                  num_labels=T([3, 2, 1]), semiring=semiring)
    npt.assert_array_equal(got.cpu(), semiring.zeros([3]).cpu())
  loss = lattice(frames=frames, num_frames=num_frames, labels=labels, num_labels=num_labels,
                 cache=None)
  npt.assert_allclose(loss.cpu(), [den - lse([10, 11]), lse([19, 20, 21]) - 21., 0.], rtol=1e-6,
                      atol=1e-6)


def test_no_cpu_fallback():
  lt = _lt()
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    lt.semirings.Log.plus(torch.zeros([2], device='cpu'), torch.zeros([2], device='cpu'))


@pytest.mark.parametrize('v', [2, 31, 256])
def test_local_normalizers(v):
  """hat_normalize / log_softmax_normalize kernels (weight_fns.py:99-136): the reference's
  known answers (tests/weight_fns_test.py:25-41), probabilities summing to one, and values and
  gradients against the reference formula evaluated in float64."""
  import torch.nn.functional as F
  import last_torch_b200 as lt
  W = lt.weight_fns
  blank = torch.tensor([2., 7.], device='cuda')
  lexical = torch.tensor([[0., 1.], [3., 5.]], device='cuda')
  nb, nl = W.hat_normalize(blank, lexical)
  sp = np.log1p(np.exp(np.array([2., 7.])))                      # softplus(blank)
  lse = np.log(np.exp(np.array([[0., 1.], [3., 5.]])).sum(-1))
  npt.assert_allclose(nb.cpu(), np.array([2., 7.]) - sp, rtol=1e-5, atol=1e-7)
  npt.assert_allclose(nl.cpu(), np.array([[0., 1.], [3., 5.]]) - lse[:, None] - sp[:, None],
                      rtol=1e-5)
  npt.assert_allclose((nb.exp() + nl.exp().sum(-1)).cpu(), [1, 1], rtol=1e-6)
  nb, nl = W.log_softmax_normalize(blank, lexical)
  npt.assert_allclose(nb.cpu(), [-0.40760595, -0.14293164], rtol=1e-5)
  npt.assert_allclose(nl.cpu(), [[-2.407606, -1.4076059], [-4.1429315, -2.1429315]], rtol=1e-5)
  g = torch.Generator(device='cuda').manual_seed(v)
  b = (torch.randn([3, 5, 7], device='cuda', generator=g) * 3).requires_grad_()
  l = (torch.randn([3, 5, 7, v], device='cuda', generator=g) * 3).requires_grad_()
  cb = torch.randn(b.shape, device='cuda', generator=g)
  cl = torch.randn(l.shape, device='cuda', generator=g)
  for name in ['hat', 'log_softmax']:
    fn = W.hat_normalize if name == 'hat' else W.log_softmax_normalize
    ob, ol = fn(b, l)
    npt.assert_allclose((ob.exp() + ol.exp().sum(-1)).detach().cpu(), 1.0, rtol=2e-5)
    gb, gl = torch.autograd.grad((ob * cb).sum() + (ol * cl).sum(), [b, l])
    b64, l64 = b.detach().double().requires_grad_(), l.detach().double().requires_grad_()
    if name == 'hat':
      z = F.softplus(b64)
      rb, rl = b64 - z, F.log_softmax(l64, dim=-1) - z.unsqueeze(-1)
    else:
      a = F.log_softmax(torch.cat([b64.unsqueeze(-1), l64], dim=-1), dim=-1)
      rb, rl = a[..., 0], a[..., 1:]
    rgb, rgl = torch.autograd.grad((rb * cb.double()).sum() + (rl * cl.double()).sum(), [b64, l64])
    npt.assert_allclose(ob.detach().cpu(), rb.detach().cpu(), rtol=1e-5, atol=1e-5)
    npt.assert_allclose(ol.detach().cpu(), rl.detach().cpu(), rtol=1e-5, atol=1e-5)
    npt.assert_allclose(gb.cpu(), rgb.cpu(), rtol=1e-4, atol=1e-5)
    npt.assert_allclose(gl.cpu(), rgl.cpu(), rtol=1e-4, atol=1e-5)


def test_locally_normalized_lattice_loss():
  """A locally normalised weight function skips the denominator (lattices.py:178-179):
  loss = -numerator; probabilities of all label strings of a tiny lattice sum to one."""
  import itertools
  import last_torch_b200 as lt
  torch.manual_seed(0)
  vocab, t = 2, 3
  context = lt.contexts.FullNGram(vocab_size=vocab, context_size=1)
  table = torch.randn([1, t, context.num_states(), 1 + vocab], device='cuda')
  lattice = lt.RecognitionLattice(
      context=context, alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher(),
      weight_fn_factory=lambda _: lt.weight_fns.LocallyNormalizedWeightFn(
          lt.weight_fns.TableWeightFn(table), lt.weight_fns.log_softmax_normalize))
  frames = torch.arange(t, device='cuda', dtype=torch.float32)[None, :, None]
  total = 0.0
  for u in range(t + 1):
    for labels in itertools.product(range(1, vocab + 1), repeat=u):
      lab = torch.tensor([list(labels) + [1] * (t - u)], device='cuda')
      loss = lattice(frames=frames, num_frames=torch.tensor([t], device='cuda'), labels=lab,
                     num_labels=torch.tensor([u], device='cuda'), cache=None)
      total += float(torch.exp(-loss))
  npt.assert_allclose(total, 1.0, rtol=1e-5)


# ---- round-2 additions: label validation, lazily-shaped parameters, autograd hygiene ---------

def _bigram_table_lattice(vocab, table):
  lt = _lt()
  return lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=vocab, context_size=1),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())


@pytest.mark.parametrize('vocab', [5, 64])
def test_labels_out_of_range(vocab):
  """labels are range-checked.  Padding after num_labels may hold anything (-1,
  vocab_size + 1): same loss and gradients as zero padding.  A label outside [0, vocab_size]
  BEFORE num_labels raises ValueError (the reference fails in one_hot, lattices.py:322); with
  validate_labels = False the kernels stay inside their buffers."""
  b, t, u = 3, 9, 5
  c = vocab + 1
  torch.manual_seed(vocab)
  table = torch.randn([b, t, c, 1 + vocab], device='cuda', requires_grad=True)
  frames = torch.arange(t, device='cuda', dtype=torch.float32)[None, :, None].expand(b, t, 1)
  nf = T([9, 7, 4])
  good = torch.randint(1, vocab + 1, [b, u], device='cuda')
  nl = torch.tensor([5, 3, 0], device='cuda')
  pos = torch.arange(u, device='cuda')[None]
  zero_pad = torch.where(pos < nl[:, None], good, torch.zeros_like(good))
  junk_pad = torch.where(pos < nl[:, None], good,
                         torch.where(pos % 2 == 0, torch.full_like(good, -1),
                                     torch.full_like(good, vocab + 1)))
  lattice = _bigram_table_lattice(vocab, table)
  res = []
  for lab in (zero_pad, junk_pad):
    loss = lattice(frames=frames, num_frames=nf, labels=lab, num_labels=nl, cache=None)
    (g,) = torch.autograd.grad(loss.sum(), table)
    res.append((loss.detach().cpu().numpy(), g.cpu().numpy()))
  npt.assert_array_equal(res[0][0], res[1][0])
  npt.assert_array_equal(res[0][1], res[1][1])
  for bad_value in (-1, vocab + 1, 10 ** 6):
    bad = good.clone()
    bad[0, 1] = bad_value
    with pytest.raises(ValueError, match='labels must be in'):
      lattice(frames=frames, num_frames=nf, labels=bad, num_labels=nl, cache=None)
    with pytest.raises(ValueError, match='labels must be in'):
      lattice._string_forward(cache=None, frames=frames, num_frames=nf, labels=bad,
                              num_labels=nl, semiring=_lt().semirings.Log)
    lattice.validate_labels = False
    guard = torch.full_like(table, 3.0)                    # neighbouring allocation
    loss = lattice(frames=frames, num_frames=nf, labels=bad, num_labels=nl, cache=None)
    (g,) = torch.autograd.grad(loss.sum(), table)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(g).all()) and float(guard.min()) == 3.0
    lattice.validate_labels = True


def test_joint_weight_fn_parameters_exist_before_the_first_call():
  """an optimizer built BEFORE the first forward must train all four projections."""
  lt = _lt()
  torch.manual_seed(3)
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=128, context_size=1),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=24, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=128, device='cuda'))
  names = [n for n, _ in lattice.named_parameters()]
  assert len(names) == 7 and not lattice.weight_fn.is_materialized()
  opt = torch.optim.SGD(lattice.parameters(), lr=0.05)
  x = torch.randn([2, 6, 16], device='cuda')
  args = dict(frames=x, num_frames=T([6, 4]), labels=T([[3, 9, 100], [7, 7, 0]]),
              num_labels=T([3, 2]))
  first = None
  for _ in range(3):
    opt.zero_grad()
    loss = lattice(**args).sum()
    loss.backward()
    opt.step()
    first = float(loss) if first is None else first
  assert lattice.weight_fn.is_materialized()
  assert all(p.grad is not None and float(p.grad.abs().max()) > 0 for p in lattice.parameters())
  assert float(loss) < first


def test_split_row_handover_is_invisible_to_autograd():
  """The arc-weight gradients never become autograd tensors on the JointWeightFn path
  (ops.JointLatticeLoss): parameter hooks, gradient accumulation over two backward passes and
  a non-trivial upstream gradient behave as with the fp32 hand-over."""
  lt = _lt()
  torch.manual_seed(9)
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=128, context_size=1),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=24, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=128, device='cuda', embedding_size=24,
          feature_size=16))
  x = torch.randn([3, 20, 16], device='cuda', requires_grad=True)
  args = dict(num_frames=T([20, 11, 5]), labels=T([[3, 9, 100, 1], [7, 7, 0, 0], [64, 0, 0, 0]]),
              num_labels=T([4, 2, 1]))
  w = torch.tensor([1.0, -0.5, 2.0], device='cuda')
  seen = []
  hook = lattice.weight_fn.joint_projection_to_vocab.weight.register_hook(
      lambda g: seen.append(g.detach().clone()))
  out = {}
  for split in (True, False):
    lattice.split_grad_handover = split
    lattice.zero_grad()
    x.grad = None
    for _ in range(2):                                   # accumulate two passes
      (lattice(frames=x, **args) * w).sum().backward()
    out[split] = [p.grad.clone() for p in lattice.parameters()] + [x.grad.clone()]
  hook.remove()
  assert len(seen) == 4
  for a, b in zip(out[True], out[False]):
    scale = float(b.abs().max()) + 1e-12
    assert float((a - b).abs().max()) / scale < 2e-5
  npt.assert_allclose(seen[0].cpu() * 2, out[True][5].cpu(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('fname', golden_files('rnncacher_'))
def test_shared_rnn_cacher_golden_on_gpu(fname):
  """SharedRNNCacher on the GPU against the reference's own forward
  (weight_fns.py:265-294) run with an injected LSTMCell (tests/golden/make_golden.py), then
  through RecognitionLattice.forward with JointWeightFn: the loss back-propagates into the
  LSTM cell and the label embedding (the reference rebuilds a random cell per call, D7)."""
  lt = _lt()
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  v, n = int(g['vocab']), int(g['context_size'])
  rnn_size, emb = int(g['rnn_size']), int(g['emb_size'])
  cell = torch.nn.LSTMCell(emb, rnn_size, device='cuda')
  with torch.no_grad():
    cell.weight_ih.copy_(T(g['weight_ih'])); cell.weight_hh.copy_(T(g['weight_hh']))
    cell.bias_ih.copy_(T(g['bias_ih'])); cell.bias_hh.copy_(T(g['bias_hh']))
  context = lt.contexts.FullNGram(vocab_size=v, context_size=n)

  def make_cacher(_):
    cacher = lt.weight_fns.SharedRNNCacher(vocab_size=v, context_size=n, rnn_size=rnn_size,
                                           rnn_embedding_size=emb, rnn_cell=cell, device='cuda')
    with torch.no_grad():
      cacher.embedding.weight.copy_(T(g['embedding']))
    return cacher

  lattice = lt.RecognitionLattice(
      context=context, alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=make_cacher,
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=16, device='cuda'))
  cache = lattice.build_cache()
  assert cache.is_cuda and tuple(cache.shape) == (context.num_states(), rnn_size)
  npt.assert_allclose(cache.detach().cpu(), g['cache'], rtol=1e-5, atol=1e-6)
  torch.manual_seed(0)
  x = torch.randn([2, 7, 5], device='cuda')
  loss = lattice(frames=x, num_frames=T([7, 4]), labels=T([[1, 2, 1], [3, 0, 0]]),
                 num_labels=T([3, 1]))
  loss.sum().backward()
  assert bool(torch.isfinite(loss).all())
  for p in list(cell.parameters()) + [lattice.weight_fn_cacher.embedding.weight]:
    assert p.grad is not None and float(p.grad.abs().max()) > 0


@pytest.mark.parametrize('normalize', ['hat', 'log_softmax'])
@pytest.mark.parametrize('vocab,hidden,ctx', [(128, 128, 1), (6, 16, 2)])
def test_locally_normalized_loss_uses_the_state_gathered_weights(normalize, vocab, hidden, ctx):
  """lattices.py:178-179 + :300-313: with LocallyNormalizedWeightFn the loss is -numerator and the
  weight function only has to be evaluated on the U+1 context states of each label string
  (WeightFn.string_frames -> [B,T,U+1,V]); same loss and parameter gradients as gathering out
  of the dense [B,T,C,V] weights (gathered_numerator = False), tensor-core and CUDA-core joint
  kernels, bigram and trigram contexts."""
  lt = _lt()
  torch.manual_seed(vocab + ctx)
  norm = lt.weight_fns.hat_normalize if normalize == 'hat' else lt.weight_fns.log_softmax_normalize
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=vocab, context_size=ctx),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=24, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.LocallyNormalizedWeightFn(
          lt.weight_fns.JointWeightFn(vocab_size=c.shape()[1], hidden_size=hidden, device='cuda',
                                      embedding_size=24, feature_size=16), norm))
  b, t, u = 3, 21, 7
  x = torch.randn([b, t, 16], device='cuda', requires_grad=True)
  args = dict(num_frames=T([21, 13, 8]),
              labels=torch.randint(1, vocab + 1, [b, u], device='cuda'), num_labels=T([7, 4, 2]))
  out = {}
  for gathered in (True, False):
    lattice.gathered_numerator = gathered
    lattice.zero_grad()
    x.grad = None
    torch.cuda.reset_peak_memory_stats()
    loss = lattice(frames=x, **args)
    loss.sum().backward()
    out[gathered] = (loss.detach().clone(),
                     [p.grad.clone() for p in lattice.parameters()] + [x.grad.clone()])
  lattice.gathered_numerator = None      # automatic = gathered for a locally normalised fn
  auto = lattice(frames=x, **args)
  npt.assert_array_equal(auto.detach().cpu(), out[True][0].cpu())
  assert bool(torch.isfinite(out[True][0]).all()) and bool((out[True][0] > 0).all())
  npt.assert_allclose(out[True][0].cpu(), out[False][0].cpu(), rtol=2e-5)
  for a, r in zip(out[True][1], out[False][1]):
    scale = float(r.abs().max()) + 1e-12
    assert float((a - r).abs().max()) / scale < 5e-5, tuple(r.shape)


def test_table_weight_fn_string_frames_matches_dense_gather():
  """WeightFn.string_frames of TableWeightFn (and the per-position default of the base class)
  against gathering the string's arcs out of all_frames, and _string_forward through it for the
  three semirings."""
  lt = _lt()
  torch.manual_seed(2)
  b, t, vocab, u = 3, 9, 5, 4
  c = 1 + vocab + vocab * vocab
  table = torch.randn([b, t, c, 1 + vocab], device='cuda')
  fn = lt.weight_fns.TableWeightFn(table)
  frames = torch.arange(t, device='cuda', dtype=torch.float32)[None, :, None].expand(b, t, 1)
  states = torch.randint(0, c, [b, u + 1], device='cuda')
  bl, lx = fn.string_frames(None, frames, states)
  db, dl = fn.all_frames(None, frames)
  bi = torch.arange(b, device='cuda')[:, None, None]
  ti = torch.arange(t, device='cuda')[None, :, None]
  npt.assert_array_equal(bl.cpu(), db[bi, ti, states[:, None, :]].cpu())
  npt.assert_array_equal(lx.cpu(), dl[bi, ti, states[:, None, :]].cpu())
  b2, l2 = lt.weight_fns.WeightFn.string_frames(fn, None, frames, states)   # base-class default
  npt.assert_array_equal(b2.cpu(), bl.cpu())
  npt.assert_array_equal(l2.cpu(), lx.cpu())
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=vocab, context_size=2),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_factory=lambda _: fn, weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  labels = torch.randint(1, vocab + 1, [b, u], device='cuda')
  for name in ['Log', 'MaxTropical', 'Real']:
    res = []
    for gathered in (True, False):
      lattice.gathered_numerator = gathered
      res.append(lattice._string_forward(cache=None, frames=frames, num_frames=T([9, 6, 2]),
                                         labels=labels, num_labels=T([4, 2, 1]),
                                         semiring=getattr(lt.semirings, name)))
    npt.assert_allclose(res[0].cpu(), res[1].cpu(), rtol=1e-6, err_msg=name)


@pytest.mark.parametrize('ctx,vocab,k', [(1, 4, 2), (2, 3, 3)])
@pytest.mark.parametrize('name', ['Log', 'MaxTropical', 'Real'])
def test_frame_label_dependent_per_state_masks(name, ctx, vocab, k):
  """lattices.py:447-453: blank_mask / lexical_mask hold ONE mask per alignment state of
  FrameLabelDependent and are added to that state's weights.  (1) zero masks: same distance, and
  the mask gradients summed over the states equal the gradient w.r.t. the shared weights;
  (2) non-zero masks: distance and mask gradients against the per-frame host recursion
  (alignment.forward, alignments.py:362-376, driven frame by frame like lattices.py:856-892)."""
  lt = _lt()
  torch.manual_seed(7 + ctx)
  semiring = getattr(lt.semirings, name)
  context = lt.contexts.FullNGram(vocab_size=vocab, context_size=ctx)
  c = context.num_states()
  b, t, n_align = 3, 6, k + 1
  raw = torch.randn([b, t, c, 1 + vocab], device='cuda')
  table = (torch.exp(raw * 0.25) / (1 + vocab) if name == 'Real' else raw).requires_grad_()
  alignment = lt.alignments.FrameLabelDependent(max_expansions=k)
  lattice = lt.RecognitionLattice(
      context=context, alignment=alignment,
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  frames = torch.arange(t, device='cuda', dtype=torch.float32)[None, :, None].expand(b, t, 1)
  nf = T([6, 4, 0])
  w = torch.tensor([1.0, -0.5, 2.0], device='cuda')

  def masks(scale):
    g = torch.Generator(device='cuda').manual_seed(3)
    bm = [(scale * torch.randn([b, t, c], device='cuda', generator=g)).requires_grad_()
          for _ in range(n_align)]
    lm = [(scale * torch.randn([b, t, 1, vocab], device='cuda', generator=g)).requires_grad_()
          for _ in range(n_align)]                     # broadcast over context states
    return bm, lm

  # (1) zero masks
  bm, lm = masks(0.0)
  dist, _ = lattice._forward(cache=None, frames=frames, num_frames=nf, semiring=semiring,
                             blank_mask=bm, lexical_mask=lm)
  plain, _ = lattice._forward(cache=None, frames=frames, num_frames=nf, semiring=semiring)
  npt.assert_allclose(dist.detach().cpu(), plain.detach().cpu(), rtol=1e-6)
  grads = torch.autograd.grad((dist * w).sum(), bm + lm)
  (gt,) = torch.autograd.grad((plain * w).sum(), table)
  tol = dict(rtol=0, atol=0) if name == 'MaxTropical' else dict(rtol=1e-4, atol=1e-6)
  npt.assert_allclose(sum(grads[:n_align]).cpu(), gt[..., 0].cpu(), **tol)
  npt.assert_allclose(sum(grads[n_align:]).cpu(), gt[..., 1:].sum(2, keepdim=True).cpu(), **tol)
  assert float(grads[2 * n_align - 1].abs().max()) == 0.0    # lexical[k] is never used

  # (2) non-zero masks against the per-frame host recursion
  bm, lm = masks(0.0 if name == 'Real' else 0.7)
  if name == 'Real':
    bm = [(0.01 * (i + 1) * torch.ones([b, t, c], device='cuda')).requires_grad_()
          for i in range(n_align)]
  dist, _ = lattice._forward(cache=None, frames=frames, num_frames=nf, semiring=semiring,
                             blank_mask=bm, lexical_mask=lm)
  got = torch.autograd.grad((dist * w).sum(), bm + lm)
  start = torch.zeros([b, c], device='cuda')
  alpha = semiring.zeros([b, c]).to('cuda').clone()
  alpha[:, 0] = semiring.ones([b]).to('cuda')
  for i in range(t):
    blank_i, lex_i = table[:, i, :, 0], table[:, i, :, 1:]
    nxt = alignment.forward(alpha=alpha, blank=[blank_i + m[:, i] for m in bm],
                            lexical=[lex_i + m[:, i] for m in lm], context=context,
                            semiring=semiring)
    alpha = torch.where((i >= nf)[:, None], alpha, nxt)
  ref = semiring.sum(alpha, dim=-1)
  npt.assert_allclose(dist.detach().cpu(), ref.detach().cpu(), rtol=2e-5, atol=1e-6)
  want = torch.autograd.grad((ref * w).sum(), bm + lm, allow_unused=True)
  for a, r in zip(got, want):
    r = torch.zeros_like(a) if r is None else r
    if name == 'MaxTropical':
      npt.assert_array_equal(a.cpu(), r.cpu())
    else:
      npt.assert_allclose(a.cpu(), r.cpu(), rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize('vocab,hidden', [(64, 128), (32, 64), (5, 32)])
def test_fused_joint_lattice_inference_matches_the_unfused_path(vocab, hidden):
  """north_star (4): JointWeightFn fused into the forward recursion (no [B,T,C,V] logits in
  memory) against the unfused path (tensor-core / CUDA-core joint kernel -> HBM -> K1): Log
  distances and alphas to 2e-5 of their scale, MaxTropical distances to 1e-5 (the unfused
  logits carry the bf16x3 rounding, the fused ones are fp32 FMAs), Viterbi labels identical
  wherever the two best paths differ by more than that."""
  lt = _lt()
  torch.manual_seed(vocab)
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=vocab, context_size=1),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=24, device='cuda'),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=hidden, device='cuda', embedding_size=24,
          feature_size=16))
  with torch.no_grad():      # larger weights: well separated paths
    lattice.weight_fn.joint_projection_to_vocab.weight.mul_(6.0)
    lattice.weight_fn.joint_projection_to_blank.weight.mul_(6.0)
  b, t = 4, 37
  x = torch.randn([b, t, 16], device='cuda')
  nf = T([37, 20, 1, 0])
  res = {}
  for fused in (False, True):
    lattice.fused_inference = fused
    with torch.no_grad():
      log_z, alphas = lattice._forward(cache=lattice.build_cache(), frames=x, num_frames=nf,
                                       semiring=lt.semirings.Log)
      vd, _ = lattice._forward(cache=lattice.build_cache(), frames=x, num_frames=nf,
                               semiring=lt.semirings.MaxTropical)
    labels, num, weights = lattice.shortest_path(frames=x, num_frames=nf)
    res[fused] = (log_z.cpu().numpy(), alphas.cpu().numpy(), vd.cpu().numpy(),
                  labels.cpu().numpy(), num.cpu().numpy(), weights.cpu().numpy())
  u, f = res[False], res[True]
  scale = np.abs(u[0]).max() + 1.0
  npt.assert_allclose(f[0], u[0], rtol=0, atol=2e-5 * scale)
  fin = np.isfinite(u[1])
  npt.assert_array_equal(np.isfinite(f[1]), fin)
  npt.assert_allclose(f[1][fin], u[1][fin], rtol=0, atol=2e-5 * scale)
  npt.assert_allclose(f[2], u[2], rtol=0, atol=1e-5 * scale)
  npt.assert_allclose(f[5], f[2], rtol=1e-6)
  npt.assert_array_equal(f[4], u[4])
  npt.assert_array_equal(f[3], u[3])
  assert f[3].min() >= 0 and f[3].max() <= vocab and (f[3] > 0).any()
  # a training call still takes the differentiable path
  loss = lattice(frames=x[:2], num_frames=nf[:2], labels=T([[1, 2], [3, 0]]), num_labels=T([2, 1]))
  loss.sum().backward()
  assert lattice.weight_fn.joint_projection_to_vocab.weight.grad is not None


def test_example_training_loop_reduces_the_loss():
  """examples/train_step.py end to end (one GPU): the loss goes down, Viterbi and entropy run."""
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, os.path.join(root, 'examples', 'train_step.py')],
                     capture_output=True, text=True, timeout=600,
                     env={k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE')})
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
  losses = [float(line.split()[-1]) for line in r.stdout.splitlines() if line.startswith('step')]
  assert len(losses) == 4 and losses[-1] < losses[0], r.stdout
  assert 'entropy' in r.stdout


@pytest.mark.parametrize('m,k,n', [(32000, 80, 512), (257, 512, 512), (130, 33, 70), (1, 5, 3),
                                   (4097, 128, 129), (32000, 512, 512), (5001, 192, 256),
                                   (1500, 64, 128), (2049, 320, 384)])
def test_input_projection_kernels_match_torch(m, k, n):
  """lt_linear_forward / lt_linear_wgrad (the bias-free input projections of JointWeightFn,
  weight_fns.py:208-211) against torch in float64: values, both gradients, odd shapes.  Shapes
  with 64-multiples take the tcgen05 kernels (bf16x3 split: 1e-5 of the scale), the others the
  fp32 FMA kernels (2e-6)."""
  from last_torch_b200.joint import _Linear
  g = torch.Generator(device='cuda').manual_seed(m + k + n)
  x = torch.randn([m, k], device='cuda', generator=g, requires_grad=True)
  w = (torch.randn([n, k], device='cuda', generator=g) / k ** 0.5).requires_grad_()
  gy = torch.randn([m, n], device='cuda', generator=g)
  y = _Linear.apply(x, w)
  gx, gw = torch.autograd.grad(y, [x, w], gy)
  xd, wd = x.detach().double().requires_grad_(), w.detach().double().requires_grad_()
  yd = xd @ wd.T
  gxd, gwd = torch.autograd.grad(yd, [xd, wd], gy.double())
  tol = 1e-5 if (k % 64 == 0 or n % 64 == 0) and m >= 128 else 2e-6
  for got, want in ((y, yd), (gx, gxd), (gw, gwd)):
    scale = float(want.abs().max()) + 1e-30
    assert float((got.double() - want).abs().max()) <= tol * scale
  # the weight gradient is reduced in a fixed order: bit-identical from run to run
  (gw2,) = torch.autograd.grad(_Linear.apply(x, w), [w], gy)
  assert torch.equal(gw, gw2)


def test_own_input_projection_kernels_through_the_lattice(monkeypatch):
  """LT_OWN_LINEAR: RecognitionLattice.forward + backward with the input projections on
  lt_linear_forward / lt_linear_wgrad instead of the library GEMM -- same loss and gradients."""
  lt = _lt()
  from last_torch_b200 import joint

  def run(own):
    monkeypatch.setattr(joint, 'OWN_LINEAR', own)
    torch.manual_seed(21)
    lattice = lt.RecognitionLattice(
        context=lt.contexts.FullNGram(vocab_size=64, context_size=1),
        alignment=lt.alignments.FrameDependent(),
        weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
            num_context_states=c.shape()[0], embedding_size=40, device='cuda'),
        weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
            vocab_size=c.shape()[1], hidden_size=128, device='cuda', embedding_size=40,
            feature_size=24))
    g = torch.Generator(device='cuda').manual_seed(4)
    x = torch.randn([3, 17, 24], device='cuda', generator=g)
    loss = lattice(frames=x, num_frames=T([17, 9, 12]), labels=T([[3, 9, 60], [7, 7, 0], [1, 0, 0]]),
                   num_labels=T([3, 2, 1]))
    grads = torch.autograd.grad(loss.sum(), list(lattice.parameters()))
    return loss.detach(), grads

  l0, g0 = run(False)
  l1, g1 = run(True)
  npt.assert_allclose(l1.cpu(), l0.cpu(), rtol=2e-6)
  for a, b in zip(g1, g0):
    scale = float(b.abs().max()) + 1e-30
    assert float((a - b).abs().max()) <= 2e-5 * scale


def test_tensor_core_frame_projection_through_the_lattice(monkeypatch):
  """A frame projection large enough for the tcgen05 kernels (B T = 4400 frames, feature size 64,
  hidden size 128: lt_linear_tensor_core) against nn.Linear's sgemm through
  RecognitionLattice.forward + backward: same loss, same gradients to the bf16x3 split's 1e-5."""
  lt = _lt()
  from last_torch_b200 import joint
  from last_torch_b200 import _native as N
  assert N.lib().lt_linear_tensor_core(4400, 64, 128) == 1
  assert N.lib().lt_linear_tensor_core(300, 64, 128) == 0          # too small to pay
  assert N.lib().lt_linear_tensor_core(4400, 80, 128) == 0         # K % 64

  def run(no_tc):
    monkeypatch.setattr(joint, 'NO_TC_LINEAR', no_tc)
    torch.manual_seed(33)
    lattice = lt.RecognitionLattice(
        context=lt.contexts.FullNGram(vocab_size=64, context_size=1),
        alignment=lt.alignments.FrameDependent(),
        weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
            num_context_states=c.shape()[0], embedding_size=40, device='cuda'),
        weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
            vocab_size=c.shape()[1], hidden_size=128, device='cuda', embedding_size=40,
            feature_size=64))
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn([4, 1100, 64], device='cuda', generator=g)
    labels = torch.randint(1, 65, [4, 30], device='cuda', generator=g)
    before = N.lib().lt_launch_count()
    loss = lattice(frames=x, num_frames=T([1100, 800, 1100, 640]), labels=labels,
                   num_labels=T([30, 12, 0, 25]))
    grads = torch.autograd.grad(loss.sum(), list(lattice.parameters()))
    return loss.detach(), grads, N.lib().lt_launch_count() - before

  l0, g0, n0 = run(True)
  l1, g1, n1 = run(False)
  # frame projection: forward + weight gradient (partials, reduction) = 3 launches; context
  # projection through the same entry points: forward, input gradient, weight gradient (2) = 4
  assert n1 == n0 + 7
  npt.assert_allclose(l1.cpu(), l0.cpu(), rtol=1e-5)
  for a, b in zip(g1, g0):
    scale = float(b.abs().max()) + 1e-30
    assert float((a - b).abs().max()) <= 2e-5 * scale
