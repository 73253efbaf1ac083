"""Generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/last_torch as is.  Two runtime patches are applied
to the imported module (no reference file is edited), exactly as SURVEY.md
section 8c describes, because the shipped Log-semiring autograd functions
cannot run (D1) or return zeros (D2):

  * _LogAddExp.backward / _LogSumExp.backward read ctx.saved_tensors and apply
    the "safe gradient" rule documented at semirings.py:222-241.

Everything else (all forward values, MaxTropical and Real autograd) is the
reference as shipped.  Arc weights are fed through the reference's own
TableWeightFn (weight_fns.py:307-342) with frames[b, t, 0] = t, the same
device tests/lattices_test.py:188-206 uses.
"""

import os
import sys

import numpy as np
import torch

REFERENCE = os.environ.get('LAST_TORCH_REFERENCE', '/root/reference')
sys.path.insert(0, REFERENCE)
import last_torch  # noqa: E402  (the reference)
from last_torch import semirings as ref_semirings  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


# --- runtime patches ---------------------------------------

class _PatchedLogAddExp(torch.autograd.Function):
  @staticmethod
  def forward(ctx, a, b):
    c = torch.max(a, b)
    c = torch.where(torch.isfinite(c), c, 0)
    ea, eb = torch.exp(a - c), torch.exp(b - c)
    z = ea + eb
    ctx.save_for_backward(ea, eb, z)
    return c + torch.log(z)

  @staticmethod
  def backward(ctx, grad):
    ea, eb, z = ctx.saved_tensors
    z = torch.where(z != 0, z, 1)
    scale = grad / z
    return scale * ea, scale * eb


class _PatchedLogSumExp(torch.autograd.Function):
  @staticmethod
  def forward(ctx, a, dim):
    c = torch.max(a, dim=dim, keepdim=True).values
    c = torch.where(torch.isfinite(c), c, 0)
    e = torch.exp(a - c)
    z = torch.sum(e, dim=dim, keepdim=True)
    ctx.save_for_backward(e, z)
    ctx.dim = dim
    return torch.squeeze(c, dim=dim) + torch.log(torch.squeeze(z, dim=dim))

  @staticmethod
  def backward(ctx, grad):
    e, z = ctx.saved_tensors
    z = torch.where(z != 0, z, 1)
    return torch.unsqueeze(grad, dim=ctx.dim) / z * e, None


def apply_patches():
  ref_semirings._logaddexp = lambda a, b: (_PatchedLogAddExp.apply(a, b),)
  ref_semirings._logsumexp = lambda a, dim: (_PatchedLogSumExp.apply(a, dim),)


def remove_patches():
  ref_semirings._logaddexp = ref_semirings._LogAddExp.apply
  ref_semirings._logsumexp = ref_semirings._LogSumExp.apply


# --- lattice cases ----------------------------------------------------------

def make_lattice(vocab, ctx, k, table):
  alignment = (last_torch.alignments.FrameDependent() if k is None else
               last_torch.alignments.FrameLabelDependent(max_expansions=k))
  return last_torch.RecognitionLattice(
      context=last_torch.contexts.FullNGram(vocab_size=vocab, context_size=ctx),
      alignment=alignment,
      weight_fn_factory=lambda _: last_torch.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: last_torch.weight_fns.NullCacher())


def lattice_case(name, vocab, ctx, k, batch, t_max, num_frames, labels,
                 num_labels, seed, scale=1.0, neg_inf_frac=0.0):
  g = torch.Generator().manual_seed(seed)
  c = sum(vocab**i for i in range(ctx + 1))
  table = torch.randn([batch, t_max, c, 1 + vocab], generator=g) * scale
  if neg_inf_frac > 0:
    drop = torch.rand([batch, t_max, c, 1 + vocab], generator=g) < neg_inf_frac
    drop[..., 0] = False      # keep blank arcs so every lattice stays connected
    # TableWeightFn looks weights up with a one-hot einsum (0 * -inf = NaN), so
    # pruned arcs use a huge finite negative instead of -inf.
    table = torch.where(drop, torch.tensor(-1e30), table)
  frames = torch.broadcast_to(
      torch.arange(t_max)[None, :, None], [batch, t_max, 1]).float()
  nf = torch.tensor(num_frames).float()
  lab = torch.tensor(labels).float()
  nl = torch.tensor(num_labels).float()
  out = dict(vocab=vocab, context_size=ctx, k=-1 if k is None else k,
             table=table.numpy(), num_frames=np.asarray(num_frames),
             labels=np.asarray(labels), num_labels=np.asarray(num_labels))

  for sr_name in ['Real', 'Log', 'MaxTropical']:
    semiring = getattr(last_torch.semirings, sr_name)
    # Real semiring: positive weights scaled by the out-degree so that T=50
    # products stay inside fp32 range.
    real_table = (torch.exp(table.clamp(min=-40) * 0.25) / (1 + vocab)
                  if sr_name == 'Real' else table)
    # --- forward values: reference exactly as shipped ---
    remove_patches()
    lattice = make_lattice(vocab, ctx, k, real_table)
    with torch.no_grad():
      dist, alphas = lattice._forward(
          cache=None, frames=frames, num_frames=nf, semiring=semiring)
      sdist = lattice._string_forward(
          cache=None, frames=frames, num_frames=nf, labels=lab, num_labels=nl,
          semiring=semiring)
    out[f'{sr_name}_dist'] = dist.numpy()
    out[f'{sr_name}_alphas'] = alphas.numpy()
    out[f'{sr_name}_string'] = sdist.numpy()
    # --- gradients ---
    if sr_name == 'Log':
      apply_patches()
    leaf = real_table.clone().requires_grad_()
    lattice = make_lattice(vocab, ctx, k, leaf)
    dist, _ = lattice._forward(
        cache=None, frames=frames, num_frames=nf, semiring=semiring)
    (gd,) = torch.autograd.grad(dist.sum(), leaf)
    out[f'{sr_name}_dist_grad'] = gd.numpy()
    leaf = real_table.clone().requires_grad_()
    lattice = make_lattice(vocab, ctx, k, leaf)
    sdist = lattice._string_forward(
        cache=None, frames=frames, num_frames=nf, labels=lab, num_labels=nl,
        semiring=semiring)
    reach = torch.isfinite(sdist) if sr_name != 'Real' else sdist != 0
    if bool(reach.any()) and sdist.requires_grad:
      (gs,) = torch.autograd.grad(torch.where(reach, sdist, 0).sum(), leaf)
    else:
      gs = torch.zeros_like(leaf)
    out[f'{sr_name}_string_grad'] = gs.numpy()
    remove_patches()
    if sr_name == 'Real':
      out['Real_table'] = real_table.numpy()

  # loss value: reference as shipped (lattices.py:131-183).
  lattice = make_lattice(vocab, ctx, k, table)
  with torch.no_grad():
    loss = lattice(frames=frames, num_frames=nf, labels=lab, num_labels=nl,
                   cache=None)
  out['loss'] = loss.numpy()

  # Second, unpatched oracle for the denominator marginals: the reference's
  # own alignment.backward (alignments.py:300-318 / :378-418) driven by a
  # harness loop with the padding masks of lattices.py:775-779.
  with torch.no_grad():
    log_z, alphas = lattice._forward(
        cache=None, frames=frames, num_frames=nf,
        semiring=last_torch.semirings.Log)
    n_align = lattice.alignment.num_states()
    beta = torch.zeros([batch, c])
    marg = torch.zeros_like(table)
    for t in reversed(range(t_max)):
      bl, lx = table[:, t, :, 0], table[:, t, :, 1:]
      nb, bm, lm = lattice.alignment.backward(
          alpha=alphas[:, t], blank=[bl] * n_align, lexical=[lx] * n_align,
          beta=beta, log_z=log_z, context=lattice.context)
      pad = (t >= nf)[:, None]
      beta = torch.where(pad, beta, nb)
      marg[:, t, :, 0] = torch.where(pad, 0, torch.stack(bm).sum(0))
      marg[:, t, :, 1:] = torch.where(pad[..., None], 0, torch.stack(lm).sum(0))
  out['Log_marginals_fb'] = marg.numpy()

  np.savez_compressed(os.path.join(OUT, f'lattice_{name}.npz'), **out)
  err = np.abs(out['Log_marginals_fb'] - out['Log_dist_grad']).max()
  print(f'{name}: C={c} loss={out["loss"]} |fb-autograd|max={err:.2e}')


# --- per-frame alignment ops -------------------------------------------------

def frame_ops_case(name, vocab, ctx, k, batch, seed):
  g = torch.Generator().manual_seed(seed)
  context = last_torch.contexts.FullNGram(vocab_size=vocab, context_size=ctx)
  alignment = (last_torch.alignments.FrameDependent() if k is None else
               last_torch.alignments.FrameLabelDependent(max_expansions=k))
  n_align = alignment.num_states()
  c = context.num_states()
  alpha = torch.randn([batch, c], generator=g)
  blank = torch.randn([batch, c], generator=g)
  lexical = torch.randn([batch, c, vocab], generator=g)
  beta = torch.randn([batch, c], generator=g)
  log_z = torch.randn([batch], generator=g) + 3
  u1 = 5
  salpha = torch.randn([batch, u1], generator=g)
  sblank = torch.randn([batch, u1], generator=g)
  slex = torch.randn([batch, u1], generator=g)
  out = dict(vocab=vocab, context_size=ctx, k=-1 if k is None else k,
             alpha=alpha.numpy(), blank=blank.numpy(), lexical=lexical.numpy(),
             beta=beta.numpy(), log_z=log_z.numpy(), salpha=salpha.numpy(),
             sblank=sblank.numpy(), slex=slex.numpy())
  remove_patches()
  with torch.no_grad():
    for sr_name in ['Real', 'Log', 'MaxTropical']:
      semiring = getattr(last_torch.semirings, sr_name)
      out[f'{sr_name}_forward'] = alignment.forward(
          alpha=alpha, blank=[blank] * n_align, lexical=[lexical] * n_align,
          context=context, semiring=semiring).numpy()
      out[f'{sr_name}_string_forward'] = alignment.string_forward(
          alpha=salpha, blank=[sblank] * n_align, lexical=[slex] * n_align,
          semiring=semiring).numpy()
      out[f'{sr_name}_forward_reduce'] = context.forward_reduce(
          lexical, semiring).numpy()
    nb, bm, lm = alignment.backward(
        alpha=alpha, blank=[blank] * n_align, lexical=[lexical] * n_align,
        beta=beta, log_z=log_z, context=context)
    out['backward_next_beta'] = nb.numpy()
    out['backward_blank_marginal'] = torch.stack(bm).sum(0).numpy()
    out['backward_lexical_marginal'] = torch.stack(lm).sum(0).numpy()
    out['backward_broadcast'] = context.backward_broadcast(beta).numpy()
    out['next_state_table'] = context.next_state_table().numpy()
  np.savez_compressed(os.path.join(OUT, f'frameops_{name}.npz'), **out)
  print(f'{name}: frame ops ok')


# --- JointWeightFn ------------------------------------------------------------

def joint_case(name, vocab, ctx, hidden, emb, feat, batch, seed):
  """Reference JointWeightFn.forward body (weight_fns.py:194-227) run with
  `last_torch.weight_fns.nn` replaced by a shim whose Linear returns a cached,
  seeded module per (in, out, bias) key."""
  import last_torch.weight_fns as wf
  torch.manual_seed(seed)
  c = sum(vocab**i for i in range(ctx + 1))
  cache_mods = {}

  class ShimNN:
    def __getattr__(self, item):
      return getattr(torch.nn, item)

    @staticmethod
    def Linear(i, o, bias=True, device=None):
      key = (i, o, bias)
      if key not in cache_mods:
        cache_mods[key] = torch.nn.Linear(i, o, bias=bias)
      return cache_mods[key]

  saved = wf.nn
  wf.nn = ShimNN()
  try:
    fn = wf.JointWeightFn(vocab_size=vocab, hidden_size=hidden)
    cache = torch.randn([c, emb])
    frame = torch.randn([batch, feat])
    with torch.no_grad():
      blank, lexical = fn(cache, frame)
      state = torch.randint(0, c, [batch])
      sblank, slexical = fn(cache, frame, state)
  finally:
    wf.nn = saved
  assert emb != feat, 'shim keys must be distinguishable'
  out = dict(
      vocab=vocab, context_size=ctx, cache=cache.numpy(), frame=frame.numpy(),
      w_ctx=cache_mods[(emb, hidden, False)].weight.detach().numpy(),
      w_frame=cache_mods[(feat, hidden, False)].weight.detach().numpy(),
      w_blank=cache_mods[(hidden, 1, True)].weight.detach().numpy()[0],
      b_blank=cache_mods[(hidden, 1, True)].bias.detach().numpy()[0],
      w_vocab=cache_mods[(hidden, vocab, True)].weight.detach().numpy(),
      b_vocab=cache_mods[(hidden, vocab, True)].bias.detach().numpy(),
      blank=blank.numpy(), lexical=lexical.numpy(), state=state.numpy(),
      state_blank=sblank.numpy(), state_lexical=slexical.numpy())
  np.savez_compressed(os.path.join(OUT, f'joint_{name}.npz'), **out)
  print(f'{name}: joint ok blank{tuple(blank.shape)} lexical{tuple(lexical.shape)}')


def joint_lattice_case(name, vocab, hidden, emb, feat, batch, t_max, num_frames, labels,
                       num_labels, seed, ctx=1, k=None):
  """The whole GNAT loss of the reference with its own JointWeightFn inside the lattice
  (lattices.py:131-183 + weight_fns.py:194-227), bigram context, in the shape envelope of the
  tensor-core kernels.  The Linear shim of joint_case makes the weights deterministic;
  the cache tensor is passed explicitly (D7).  Loss value: the reference as shipped.  Parameter
  gradients: patched Log autograd (D1/D2) through `_forward - _string_forward`, because
  `forward()` detaches the denominator (D3)."""
  import last_torch.weight_fns as wf
  torch.manual_seed(seed)
  c = sum(vocab**i for i in range(ctx + 1))
  cache_mods = {}

  class ShimNN:
    def __getattr__(self, item):
      return getattr(torch.nn, item)

    @staticmethod
    def Linear(i, o, bias=True, device=None):
      key = (i, o, bias)
      if key not in cache_mods:
        cache_mods[key] = torch.nn.Linear(i, o, bias=bias)
      return cache_mods[key]

  assert emb != feat, 'shim keys must be distinguishable'
  saved = wf.nn
  wf.nn = ShimNN()
  try:
    lattice = last_torch.RecognitionLattice(
        context=last_torch.contexts.FullNGram(vocab_size=vocab, context_size=ctx),
        alignment=(last_torch.alignments.FrameDependent() if k is None else
                   last_torch.alignments.FrameLabelDependent(max_expansions=k)),
        weight_fn_factory=lambda _: wf.JointWeightFn(vocab_size=vocab, hidden_size=hidden),
        weight_fn_cacher_factory=lambda _: last_torch.weight_fns.NullCacher())
    cache = torch.randn([c, emb]).requires_grad_()
    frames = torch.randn([batch, t_max, feat])
    nf = torch.tensor(num_frames).float()
    lab = torch.tensor(labels).float()
    nl = torch.tensor(num_labels).float()
    remove_patches()
    with torch.no_grad():
      loss = lattice(frames=frames, num_frames=nf, labels=lab, num_labels=nl, cache=cache)
    apply_patches()
    log_z, _ = lattice._forward(cache=cache, frames=frames, num_frames=nf,
                                semiring=last_torch.semirings.Log)
    num = lattice._string_forward(cache=cache, frames=frames, num_frames=nf, labels=lab,
                                  num_labels=nl, semiring=last_torch.semirings.Log)
    mods = dict(w_ctx=cache_mods[(emb, hidden, False)], w_frame=cache_mods[(feat, hidden, False)],
                blank=cache_mods[(hidden, 1, True)], vocab=cache_mods[(hidden, vocab, True)])
    leaves = [cache, mods['w_ctx'].weight, mods['w_frame'].weight, mods['blank'].weight,
              mods['blank'].bias, mods['vocab'].weight, mods['vocab'].bias]
    grads = torch.autograd.grad((log_z - num).sum(), leaves)
    remove_patches()
  finally:
    wf.nn = saved
    remove_patches()
  names = ['cache', 'w_ctx', 'w_frame', 'w_blank', 'b_blank', 'w_vocab', 'b_vocab']
  out = dict(vocab=vocab, hidden=hidden, context_size=ctx, k=-1 if k is None else k,
             frames=frames.numpy(), num_frames=np.asarray(num_frames),
             labels=np.asarray(labels), num_labels=np.asarray(num_labels),
             loss=loss.numpy(), loss_patched=(log_z - num).detach().numpy())
  for n, leaf, g in zip(names, leaves, grads):
    out[n] = leaf.detach().numpy()
    out['grad_' + n] = g.numpy()
  np.savez_compressed(os.path.join(OUT, f'jointlattice_{name}.npz'), **out)
  print(f'{name}: joint lattice ok loss={out["loss"]} |grad w_vocab|max='
        f'{np.abs(out["grad_w_vocab"]).max():.3e}')


def rnn_cacher_case(name, vocab, ctx, rnn_size, emb_size, seed):
  """Reference SharedRNNCacher.forward (weight_fns.py:265-294) with an injected, seeded LSTMCell
  (as shipped it builds a fresh random cell per call when none is given)."""
  torch.manual_seed(seed)
  cell = torch.nn.LSTMCell(emb_size, rnn_size)
  cacher = last_torch.weight_fns.SharedRNNCacher(
      vocab_size=vocab, context_size=ctx, rnn_size=rnn_size, rnn_embedding_size=emb_size,
      rnn_cell=cell)
  with torch.no_grad():
    cache = cacher()
  out = dict(vocab=vocab, context_size=ctx, rnn_size=rnn_size, emb_size=emb_size,
             embedding=cacher.embedding.weight.detach().numpy(),
             weight_ih=cell.weight_ih.detach().numpy(), weight_hh=cell.weight_hh.detach().numpy(),
             bias_ih=cell.bias_ih.detach().numpy(), bias_hh=cell.bias_hh.detach().numpy(),
             cache=cache.numpy())
  np.savez_compressed(os.path.join(OUT, f'rnncacher_{name}.npz'), **out)
  print(f'{name}: rnn cacher ok cache{tuple(cache.shape)}')


def main():
  torch.set_num_threads(4)
  if os.environ.get('LT_GOLDEN_ONLY') == 'rnncacher':
    rnn_cacher_case('bigram_v5', vocab=5, ctx=1, rnn_size=12, emb_size=7, seed=50)
    rnn_cacher_case('trigram_v3', vocab=3, ctx=2, rnn_size=8, emb_size=6, seed=51)
    return
  lattice_case('fd_bigram_v3', vocab=3, ctx=1, k=None, batch=4, t_max=6,
               num_frames=[6, 4, 1, 0],
               labels=[[1, 3, 2], [2, 2, 0], [3, 0, 0], [1, 2, 3]],
               num_labels=[3, 2, 1, 0], seed=1)
  lattice_case('fd_bigram_unreachable', vocab=2, ctx=1, k=None, batch=4,
               t_max=6, num_frames=[6, 3, 2, 1],
               labels=[[1, 1, 1, 1], [2, 2, 2, 2], [1, 2, 1, 2], [2, 1, 2, 1]],
               num_labels=[4, 3, 1, 2], seed=2)
  lattice_case('fd_trigram_v3', vocab=3, ctx=2, k=None, batch=3, t_max=5,
               num_frames=[5, 3, 2], labels=[[1, 2, 3, 1], [3, 3, 0, 0],
                                             [2, 1, 0, 0]],
               num_labels=[4, 2, 2], seed=3)
  lattice_case('fd_unigram_v4', vocab=4, ctx=0, k=None, batch=2, t_max=5,
               num_frames=[5, 2], labels=[[4, 1, 2], [3, 0, 0]],
               num_labels=[3, 1], seed=4)
  lattice_case('fd_4gram_v2', vocab=2, ctx=3, k=None, batch=2, t_max=6,
               num_frames=[6, 5], labels=[[1, 2, 2, 1, 1], [2, 1, 2, 0, 0]],
               num_labels=[5, 3], seed=5)
  lattice_case('fld2_bigram_v3', vocab=3, ctx=1, k=2, batch=4, t_max=5,
               num_frames=[5, 3, 1, 0],
               labels=[[1, 3, 2, 2, 1, 3], [2, 2, 1, 0, 0, 0],
                       [3, 1, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]],
               num_labels=[6, 3, 2, 0], seed=6)
  lattice_case('fld2_trigram_v2', vocab=2, ctx=2, k=2, batch=3, t_max=4,
               num_frames=[4, 3, 2], labels=[[1, 2, 2, 1, 2], [2, 2, 1, 0, 0],
                                             [1, 0, 0, 0, 0]],
               num_labels=[5, 3, 1], seed=7)
  lattice_case('fld1_bigram_v4', vocab=4, ctx=1, k=1, batch=2, t_max=6,
               num_frames=[6, 4], labels=[[4, 1, 2, 3], [3, 3, 0, 0]],
               num_labels=[4, 2], seed=8)
  lattice_case('fd_bigram_pruned', vocab=4, ctx=1, k=None, batch=3, t_max=6,
               num_frames=[6, 5, 3], labels=[[1, 2, 3], [4, 4, 0], [2, 0, 0]],
               num_labels=[3, 2, 1], seed=9, neg_inf_frac=0.3)
  lattice_case('fd_bigram_v16_cfg1', vocab=16, ctx=1, k=None, batch=4,
               t_max=50, num_frames=[50, 37, 25, 50],
               labels=np.random.RandomState(0).randint(1, 17, [4, 12]).tolist(),
               num_labels=[12, 9, 5, 12], seed=10)
  lattice_case('fd_bigram_wide', vocab=5, ctx=1, k=None, batch=2, t_max=12,
               num_frames=[12, 9], labels=[[1, 5, 2, 3], [4, 4, 1, 0]],
               num_labels=[4, 3], seed=11, scale=8.0)
  frame_ops_case('fd_trigram_v3', vocab=3, ctx=2, k=None, batch=3, seed=20)
  frame_ops_case('fld2_trigram_v3', vocab=3, ctx=2, k=2, batch=3, seed=21)
  frame_ops_case('fld3_bigram_v4', vocab=4, ctx=1, k=3, batch=2, seed=22)
  frame_ops_case('fd_unigram_v3', vocab=3, ctx=0, k=None, batch=2, seed=23)
  joint_case('bigram_v5', vocab=5, ctx=1, hidden=16, emb=24, feat=8, batch=3,
             seed=30)
  joint_case('trigram_v3', vocab=3, ctx=2, hidden=32, emb=12, feat=20, batch=4,
             seed=31)
  joint_lattice_case('bigram_v128_h128', vocab=128, hidden=128, emb=24, feat=16, batch=2,
                     t_max=5, num_frames=[5, 3], labels=[[7, 100, 7], [128, 1, 0]],
                     num_labels=[3, 2], seed=40)
  joint_lattice_case('bigram_v64_h128', vocab=64, hidden=128, emb=24, feat=16, batch=3,
                     t_max=6, num_frames=[6, 4, 2], labels=[[5, 5, 64, 1], [9, 33, 0, 0], [2, 0, 0, 0]],
                     num_labels=[4, 2, 1], seed=41)
  joint_lattice_case('trigram_v3_fld2', vocab=3, hidden=32, emb=12, feat=20, batch=3, t_max=5,
                     num_frames=[5, 4, 2], labels=[[1, 3, 2, 2, 1], [2, 2, 0, 0, 0], [3, 1, 0, 0, 0]],
                     num_labels=[5, 2, 2], seed=42, ctx=2, k=2)
  rnn_cacher_case('bigram_v5', vocab=5, ctx=1, rnn_size=12, emb_size=7, seed=50)
  rnn_cacher_case('trigram_v3', vocab=3, ctx=2, rnn_size=8, emb_size=6, seed=51)


if __name__ == '__main__':
  main()
