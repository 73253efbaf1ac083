"""Error measurement shared by the parity tests and tools/parity_table.py (TEST INFRASTRUCTURE).

For every fixture the same three things are compared on identical fp32 inputs:

  truth      the oracle evaluated in float64 (oracle/lattice_oracle.py, or the double build of
             oracle/lattice_oracle.c at the headline size)
  reference  the fp32 outputs of the UNMODIFIED reference stored in tests/golden/*.npz
             (tests/golden/make_golden.py, make_golden_headline.py)
  gpu        the CUDA path through the public API

and reported as max |x - truth| (absolute) and max |x - truth| / |truth| over the entries with
|truth| >= REL_FLOOR * max |truth| (relative).  `north_star` asks for 1e-5 relative agreement with
the reference; where the reference itself is further than that from the truth (long utterances:
logZ ~ 5e3, one fp32 ulp = 4.9e-4), the bar is  gpu error <= 2 x reference error.
"""
import hashlib
import os

import numpy as np

from oracle import lattice_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
REL_FLOOR = 1e-4


def err_stats(x, truth):
  """(max abs error, max relative error over the significant entries, max |truth|)."""
  x = np.asarray(x, np.float64)
  truth = np.asarray(truth, np.float64)
  fin = np.isfinite(truth)
  if not fin.any():
    return 0.0, 0.0, 0.0
  d = np.abs(x[fin] - truth[fin])
  t = np.abs(truth[fin])
  scale = float(t.max())
  big = t >= REL_FLOOR * scale if scale > 1e-9 else np.zeros_like(t, bool)
  rel = float((d[big] / t[big]).max()) if big.any() else 0.0
  return float(d.max()), rel, scale


def row(name, quantity, ref, gpu, truth):
  ra, rr, scale = err_stats(ref, truth) if ref is not None else (None, None, None)
  ga, gr, scale2 = err_stats(gpu, truth)
  return {'case': name, 'quantity': quantity, 'scale': scale2,
          'reference_fp32_abs': ra, 'reference_fp32_rel': rr, 'gpu_abs': ga, 'gpu_rel': gr}


def within_bar(r, factor=2.0, abs_floor=2e-7, north_star=1e-5):
  """The parity bar, on the max absolute error against the fp64 truth:

      gpu error <= max( factor x reference error + 2e-7 x scale,  north_star / 2 x scale )

  i.e. the GPU is at most twice as far from the truth as the fp32 reference is (the floor covers
  fixtures where both sit at fp32 round-off), OR it is inside half of north_star's 1e-5 of the
  quantity's magnitude -- in which case |gpu - reference| <= 1e-5 x scale holds whenever the
  reference itself does.  The elementwise relative errors are reported next to it; they are
  dominated by near-cancelling entries (gradient = denominator - numerator posteriors) and are
  not part of the bar."""
  if r['reference_fp32_abs'] is None:
    return True
  scale = max(r['scale'], 1.0)
  return r['gpu_abs'] <= max(factor * r['reference_fp32_abs'] + abs_floor * scale,
                             0.5 * north_star * scale)


# --------------------------------------------------------------------------- lattice goldens --

def lattice_truth(g):
  """fp64 oracle on the fixture's fp32 table: loss [B], d sum(finite loss) / d table."""
  vocab, ctx, k = int(g['vocab']), int(g['context_size']), int(g['k'])
  kk, fd = (0, True) if k < 0 else (k, False)
  tab = g['table'].astype(np.float64)
  blank, lex = np.ascontiguousarray(tab[..., 0]), np.ascontiguousarray(tab[..., 1:])
  with np.errstate(all='ignore'):
    loss, gb, gl = O.lattice_loss_and_grads(blank, lex, g['num_frames'], g['labels'],
                                            g['num_labels'], O.FullNGram(vocab, ctx), kk, fd)
  grad = np.concatenate([gb[..., None], gl], axis=-1)
  return loss, grad


def lattice_golden_rows(fname, gpu_loss, gpu_grad):
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  loss64, grad64 = lattice_truth(g)
  fin = np.isfinite(loss64)
  ref_grad = g['Log_dist_grad'] - g['Log_string_grad']
  name = fname[:-4]
  return [row(name, 'loss', g['loss'][fin], gpu_loss[fin], loss64[fin]),
          row(name, 'grad', ref_grad[fin], gpu_grad[fin], grad64[fin])]


# ----------------------------------------------------------------------------- headline golden --

def headline_inputs(g):
  """Regenerates the arc weights of a headline fixture from its seed (torch CPU generator) and
  checks them against the stored checksum."""
  import torch
  seed, v, b, t, u = (int(g[k]) for k in ('seed', 'vocab', 'batch', 't_max', 'u'))
  gen = torch.Generator().manual_seed(seed)
  table = torch.randn([b, t, v + 1, 1 + v], generator=gen)
  digest = hashlib.sha256(table.numpy().tobytes()).hexdigest()
  if digest != str(g['sha256']):
    raise RuntimeError('regenerated headline weights differ from the fixture (torch build?)')
  labels = np.random.RandomState(seed).randint(1, v + 1, size=(b, u))
  return table, labels


def headline_truth(g, table, labels):
  """Double-precision C oracle at the headline size (a few seconds)."""
  from oracle import c_oracle
  tab = table.numpy()
  blank = np.ascontiguousarray(tab[..., 0])
  lex = np.ascontiguousarray(tab[..., 1:])
  loss, gb, gl, log_z, alphas = c_oracle.lattice_loss_and_grads(
      blank, lex, g['num_frames'], labels, g['num_labels'], int(g['vocab']), 1, real='f64')
  return dict(loss=loss, grad_blank=gb, grad_lexical=gl, log_z=log_z, alphas=alphas)


def headline_rows(name, g, truth, gpu):
  """gpu: dict(loss, grad_blank [B,T,C], grad_lexical_sample [B,S,C,V], log_z)."""
  sf = g['sample_frames']
  return [
      row(name, 'loss', g['loss'], gpu['loss'], truth['loss']),
      row(name, 'log_z', g['log_z'], gpu['log_z'], truth['log_z']),
      row(name, 'grad_blank', g['grad_blank'], gpu['grad_blank'], truth['grad_blank']),
      row(name, 'grad_lexical (sampled frames)', g['grad_lexical_sample'],
          gpu['grad_lexical_sample'], truth['grad_lexical'][:, sf]),
  ]


# ------------------------------------------------------------------------ joint-lattice goldens --

PARAMS = ['cache', 'w_ctx', 'w_frame', 'w_blank', 'b_blank', 'w_vocab', 'b_vocab']


def joint_lattice_truth(g):
  """fp64: loss [B] and d sum(loss) / d parameter of the whole JointWeightFn + lattice loss."""
  f = lambda n: g[n].astype(np.float64)
  vocab, ctx, k = int(g['vocab']), int(g['context_size']), int(g['k'])
  kk, fd = (0, True) if k < 0 else (k, False)
  cache, frames, w_ctx, w_frame = f('cache'), f('frames'), f('w_ctx'), f('w_frame')
  w_blank, w_vocab, b_vocab = f('w_blank').reshape(-1), f('w_vocab'), f('b_vocab')
  b_blank = float(np.asarray(g['b_blank']).reshape(-1)[0])
  blank, lex = O.joint_weights(cache, frames, w_ctx, w_frame, w_blank, b_blank, w_vocab, b_vocab)
  loss, gb, gl = O.lattice_loss_and_grads(blank, lex, g['num_frames'], g['labels'],
                                          g['num_labels'], O.FullNGram(vocab, ctx), kk, fd)
  joint = np.tanh((cache @ w_ctx.T)[None, None] + (frames @ w_frame.T)[:, :, None, :])
  dpre = (gl @ w_vocab + gb[..., None] * w_blank) * (1.0 - joint * joint)
  g_pc = dpre.sum(axis=(0, 1))
  g_pf = dpre.sum(axis=2)
  grads = {
      'w_vocab': np.einsum('btcv,btch->vh', gl, joint), 'b_vocab': gl.sum(axis=(0, 1, 2)),
      'w_blank': np.einsum('btc,btch->h', gb, joint).reshape(g['w_blank'].shape),
      'b_blank': np.asarray(gb.sum()).reshape(g['b_blank'].shape),
      'w_ctx': g_pc.T @ cache, 'cache': g_pc @ w_ctx,
      'w_frame': np.einsum('bth,btd->hd', g_pf, frames),
  }
  return loss, grads


def joint_lattice_rows(fname, gpu_loss, gpu_grads, tag=''):
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  loss64, grads64 = joint_lattice_truth(g)
  name = fname[:-4] + tag
  rows = [row(name, 'loss', g['loss'], gpu_loss, loss64)]
  for p in PARAMS:
    rows.append(row(name, 'grad_' + p, g['grad_' + p], gpu_grads[p], grads64[p]))
  return rows


# ------------------------------------------------------------------------------ GPU runners ----
# (import torch / the product package lazily: the functions above are also used on CPU)

def _lt():
  import last_torch_b200 as last_torch
  return last_torch


def _cuda(x, dtype=None):
  import torch
  t = torch.as_tensor(np.asarray(x), device='cuda')
  return t.to(dtype or torch.float32)


def table_lattice(vocab, ctx, k, table, flags=0):
  lt = _lt()
  alignment = (lt.alignments.FrameDependent() if k < 0 else
               lt.alignments.FrameLabelDependent(max_expansions=k))
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=vocab, context_size=ctx), alignment=alignment,
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  lattice.kernel_flags = flags
  return lattice


def frames_for(b, t):
  import torch
  return torch.arange(t, device='cuda', dtype=torch.float32)[None, :, None].expand(b, t, 1)


def gpu_lattice_golden(fname, flags=0):
  """(loss [B], d sum(finite loss) / d table) of a lattice_*.npz fixture on the GPU."""
  import torch
  g = np.load(os.path.join(GOLDEN_DIR, fname))
  vocab, ctx, k = int(g['vocab']), int(g['context_size']), int(g['k'])
  b, t = g['table'].shape[:2]
  table = _cuda(g['table']).requires_grad_()
  lattice = table_lattice(vocab, ctx, k, table, flags)
  loss = lattice(frames=frames_for(b, t), num_frames=_cuda(g['num_frames']),
                 labels=_cuda(g['labels']), num_labels=_cuda(g['num_labels']), cache=None)
  fin = torch.isfinite(loss)
  (gt,) = torch.autograd.grad(torch.where(fin, loss, torch.zeros_like(loss)).sum(), table)
  return loss.detach().cpu().numpy(), gt.cpu().numpy()


def gpu_headline(g, table, labels):
  """Loss, logZ and gradients of a headline fixture on the GPU (weights as a TableWeightFn
  leaf, like the reference run that produced the fixture)."""
  import torch
  lt = _lt()
  v, b, t = int(g['vocab']), int(g['batch']), int(g['t_max'])
  leaf = table.cuda().requires_grad_()
  lattice = table_lattice(v, 1, -1, leaf)
  nf = _cuda(g['num_frames'])
  loss = lattice(frames=frames_for(b, t), num_frames=nf, labels=_cuda(labels),
                 num_labels=_cuda(g['num_labels']), cache=None)
  (gt,) = torch.autograd.grad(loss.sum(), leaf)
  with torch.no_grad():
    log_z, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=nf,
                                semiring=lt.semirings.Log)
  sf = torch.as_tensor(g['sample_frames'], device='cuda')
  out = dict(loss=loss.detach().cpu().numpy(), log_z=log_z.cpu().numpy(),
             grad_blank=gt[..., 0].cpu().numpy(),
             grad_lexical_sample=gt[:, sf][..., 1:].cpu().numpy(),
             grad_frame_sums=gt.double().sum((2, 3)).cpu().numpy())
  del gt, leaf
  return out


def joint_lattice_for(g, device='cuda'):
  """RecognitionLattice with the fixture's JointWeightFn parameters injected."""
  import torch
  lt = _lt()
  v, h = int(g['vocab']), int(g['hidden'])
  e, d = g['cache'].shape[1], g['frames'].shape[2]
  k = int(g['k'])
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=v, context_size=int(g['context_size'])),
      alignment=(lt.alignments.FrameDependent() if k < 0 else
                 lt.alignments.FrameLabelDependent(max_expansions=k)),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=e, device=device),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=v, hidden_size=h, device=device, embedding_size=e, feature_size=d))
  fn, cacher = lattice.weight_fn, lattice.weight_fn_cacher
  with torch.no_grad():
    cacher.embedding.weight.copy_(_cuda(g['cache']))
    fn.context_projection.weight.copy_(_cuda(g['w_ctx']))
    fn.blank_projection.weight.copy_(_cuda(g['w_frame']))
    fn.joint_projection_to_blank.weight.copy_(_cuda(g['w_blank']).reshape(1, -1))
    fn.joint_projection_to_blank.bias.copy_(_cuda(g['b_blank']).reshape(1))
    fn.joint_projection_to_vocab.weight.copy_(_cuda(g['w_vocab']))
    fn.joint_projection_to_vocab.bias.copy_(_cuda(g['b_vocab']))
  return lattice


def joint_param_grads(lattice):
  fn, cacher = lattice.weight_fn, lattice.weight_fn_cacher
  got = {'cache': cacher.embedding.weight.grad, 'w_ctx': fn.context_projection.weight.grad,
         'w_frame': fn.blank_projection.weight.grad,
         'w_blank': fn.joint_projection_to_blank.weight.grad,
         'b_blank': fn.joint_projection_to_blank.bias.grad,
         'w_vocab': fn.joint_projection_to_vocab.weight.grad,
         'b_vocab': fn.joint_projection_to_vocab.bias.grad}
  return {k: v.detach().cpu().numpy() for k, v in got.items()}


def gpu_joint_lattice(g, split):
  lattice = joint_lattice_for(g)
  lattice.split_grad_handover = bool(split)
  loss = lattice(frames=_cuda(g['frames']), num_frames=_cuda(g['num_frames']),
                 labels=_cuda(g['labels']), num_labels=_cuda(g['num_labels']))
  loss.sum().backward()
  grads = joint_param_grads(lattice)
  return loss.detach().cpu().numpy(), {k: v.reshape(np.shape(g[k])) for k, v in grads.items()}


def synthetic_joint_case(seed, vocab, hidden, emb, feat, batch, t_max, u, ragged=True):
  """A JointWeightFn + bigram lattice problem of arbitrary size in the fixture format (no
  reference outputs: the truth is the fp64 oracle)."""
  rng = np.random.RandomState(seed)
  c = 1 + vocab
  nf = rng.randint(t_max // 2, t_max + 1, size=batch) if ragged else np.full(batch, t_max)
  nf[0] = t_max
  return dict(
      vocab=vocab, hidden=hidden, context_size=1, k=-1,
      cache=rng.randn(c, emb).astype(np.float32),
      frames=rng.randn(batch, t_max, feat).astype(np.float32),
      w_ctx=(rng.randn(hidden, emb) / np.sqrt(emb)).astype(np.float32),
      w_frame=(rng.randn(hidden, feat) / np.sqrt(feat)).astype(np.float32),
      w_blank=(rng.randn(1, hidden) / np.sqrt(hidden)).astype(np.float32),
      b_blank=np.asarray([0.1], np.float32),
      w_vocab=(rng.randn(vocab, hidden) / np.sqrt(hidden)).astype(np.float32),
      b_vocab=(rng.randn(vocab) * 0.1).astype(np.float32),
      num_frames=nf, labels=rng.randint(1, vocab + 1, size=(batch, u)),
      num_labels=np.minimum(rng.randint(u // 2, u + 1, size=batch), nf))


def joint_lattice_truth_large(g):
  """joint_lattice_truth for sizes where the numpy lattice oracle is too slow: the joint network
  in float64 numpy (BLAS), the lattice in the double build of the C oracle."""
  from oracle import c_oracle
  f = lambda n: np.asarray(g[n], np.float64)
  vocab, ctx, k = int(g['vocab']), int(g['context_size']), int(g['k'])
  cache, frames, w_ctx, w_frame = f('cache'), f('frames'), f('w_ctx'), f('w_frame')
  w_blank, w_vocab, b_vocab = f('w_blank').reshape(-1), f('w_vocab'), f('b_vocab')
  b_blank = float(np.asarray(g['b_blank']).reshape(-1)[0])
  b, t, _ = frames.shape
  h = w_ctx.shape[0]
  joint = np.tanh((cache @ w_ctx.T)[None, None] + (frames @ w_frame.T)[:, :, None, :])
  j2 = joint.reshape(-1, h)
  lex = (j2 @ w_vocab.T + b_vocab).reshape(b, t, -1, vocab)
  blank = (j2 @ w_blank + b_blank).reshape(b, t, -1)
  loss, gb, gl, _, _ = c_oracle.lattice_loss_and_grads(
      blank, lex, g['num_frames'], g['labels'], g['num_labels'], vocab, ctx, k, real='f64')
  gl2 = gl.reshape(-1, vocab)
  dpre = ((gl2 @ w_vocab + gb.reshape(-1, 1) * w_blank) * (1.0 - j2 * j2)).reshape(joint.shape)
  g_pc = dpre.sum(axis=(0, 1))
  g_pf = dpre.sum(axis=2)
  grads = {
      'w_vocab': gl2.T @ j2, 'b_vocab': gl2.sum(axis=0),
      'w_blank': (gb.reshape(-1) @ j2).reshape(np.shape(g['w_blank'])),
      'b_blank': np.asarray(gb.sum()).reshape(np.shape(g['b_blank'])),
      'w_ctx': g_pc.T @ cache, 'cache': g_pc @ w_ctx,
      'w_frame': g_pf.reshape(-1, h).T @ frames.reshape(b * t, -1),
  }
  return loss, grads


# ------------------------------------------------------------- context_size 2, long utterances --

def trigram_rows(k, vocab=32, t=200, b=2, u=30, tag='', scale=1.0):
  """Rows (grad_blank, grad_lexical) of a FullNGram(vocab, 2) lattice at T = t against the double
  build of the C oracle, and the relative error of the loss; k = -1: FrameDependent."""
  import torch
  from oracle import c_oracle
  n = 2
  c = 1 + vocab + vocab * vocab
  rng = np.random.RandomState(5)
  gen = torch.Generator().manual_seed(vocab)
  table = torch.randn([b, t, c, 1 + vocab], generator=gen) * scale
  nf = np.array([t] + [int(x) for x in rng.randint(t // 2, t + 1, size=b - 1)])
  labels = rng.randint(1, vocab + 1, size=(b, u))
  nl = rng.randint(0, u + 1, size=b)
  tab = table.numpy()
  loss64, gb64, gl64, _, _ = c_oracle.lattice_loss_and_grads(
      np.ascontiguousarray(tab[..., 0]), np.ascontiguousarray(tab[..., 1:]), nf, labels, nl, vocab,
      n, k, real='f64')
  leaf = table.cuda().requires_grad_()
  lattice = table_lattice(vocab, n, k, leaf)
  loss = lattice(frames=frames_for(b, t), num_frames=_cuda(nf), labels=_cuda(labels),
                 num_labels=_cuda(nl), cache=None)
  (gt,) = torch.autograd.grad(loss.sum(), leaf)
  gt = gt.cpu().numpy()
  name = f'trigram_v{vocab}_t{t}_' + ('fd' if k < 0 else f'fld{k}') + tag
  loss_rel = float(np.abs(loss.detach().cpu().numpy() - loss64).max() / np.abs(loss64).max())
  return [row(name, 'grad_blank', None, gt[..., 0], gb64),
          row(name, 'grad_lexical', None, gt[..., 1:], gl64)], loss_rel
