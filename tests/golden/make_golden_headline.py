"""Generates tests/golden/headline_*.npz: the UNMODIFIED reference run at the size the
headline number is quoted on (BASELINE.json configs[1] geometry: bigram, vocab 256, 257 context
states, T = 1000 frames, U = 120 labels; B = 2 with one ragged utterance).

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden_headline.py

The dense arc weights (2 x 1000 x 257 x 257 floats = 528 MB) are NOT stored: they are
regenerated on both sides from a seeded torch CPU generator (`headline_weights`, same torch
build here and on the GPU box) and pinned by a checksum.  Stored are the reference's fp32
outputs: loss / logZ / numerator / MaxTropical distance per utterance, the full blank-arc
gradient, the lexical-arc gradient and alphas on a few sampled frames, per-frame gradient sums,
and the Viterbi arcs (from the reference's own MaxTropical autograd, which works as shipped).

The reference's loop (lattices.py:865-886) is fed through a harness WeightFn that returns
per-frame leaf tensors (a WeightFn is any nn.Module with forward(cache, frame, state),
weight_fns.py:42-83) -- the reference's TableWeightFn would build a [T, C, V+1] one-hot product
per frame.  Log gradients use the same two runtime patches as make_golden.py; no
reference file is edited.
"""

import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (imports the reference, provides the D1/D2 patches)

last_torch = MG.last_torch

SAMPLE_FRAMES = [0, 1, 499, 776, 998, 999]


def headline_weights(seed, b, t, c, v):
  """The arc-weight table of a headline case: N(0,1) from a seeded CPU generator."""
  g = torch.Generator().manual_seed(seed)
  return torch.randn([b, t, c, 1 + v], generator=g)


def headline_labels(seed, b, u, v):
  return np.random.RandomState(seed).randint(1, v + 1, size=(b, u))


def checksum(table):
  return hashlib.sha256(table.numpy().tobytes()).hexdigest()


class PrecompWeightFn(last_torch.weight_fns.WeightFn):
  """frame[..., 0] holds the frame index t.  state=None (the denominator loop, one call per
  frame outside vmap): returns the per-frame leaf tensors.  With a state (the numerator, called
  under torch.vmap over frames): gathers rows of the stacked table."""

  def __init__(self, leaves=None, table=None):
    super().__init__()
    self.leaves = leaves        # list over t of [B, C, 1+V]
    self.table = table          # [B, T, C, 1+V]

  def forward(self, cache, frame, state=None):
    del cache
    if state is None:
      w = self.leaves[int(frame[0, 0])]
      return w[..., 0], w[..., 1:]
    idx = frame[..., 0].long()
    b = torch.arange(self.table.shape[0])
    w = self.table[b, idx, state.long()]          # [B, 1+V]
    return w[..., 0], w[..., 1:]


def lattice_for(vocab, weight_fn):
  return last_torch.RecognitionLattice(
      context=last_torch.contexts.FullNGram(vocab_size=vocab, context_size=1),
      alignment=last_torch.alignments.FrameDependent(),
      weight_fn_factory=lambda _: weight_fn,
      weight_fn_cacher_factory=lambda _: last_torch.weight_fns.NullCacher())


def headline_case(name, seed, vocab, batch, t_max, num_frames, u, num_labels):
  c = 1 + vocab
  table = headline_weights(seed, batch, t_max, c, vocab)
  labels = headline_labels(seed, batch, u, vocab)
  frames = torch.broadcast_to(torch.arange(t_max)[None, :, None], [batch, t_max, 1]).float()
  nf = torch.tensor(num_frames).float()
  lab = torch.tensor(labels).float()
  nl = torch.tensor(num_labels).float()
  out = dict(seed=seed, vocab=vocab, batch=batch, t_max=t_max, num_frames=np.asarray(num_frames),
             u=u, num_labels=np.asarray(num_labels), sha256=checksum(table),
             sample_frames=np.asarray(SAMPLE_FRAMES))

  # ---- Log: loss value by the reference as shipped --------------------------------
  t0 = time.time()
  MG.remove_patches()
  leaves = [table[:, t].clone() for t in range(t_max)]
  with torch.no_grad():
    lattice = lattice_for(vocab, PrecompWeightFn(leaves, table))
    out['loss'] = lattice(frames=frames, num_frames=nf, labels=lab, num_labels=nl,
                          cache=None).numpy()
  print(f'{name}: reference forward() {time.time() - t0:.1f} s, loss {out["loss"]}')

  # ---- Log: gradients by patched autograd (D1/D2) ------------------------------------
  t0 = time.time()
  MG.apply_patches()
  leaves = [table[:, t].clone().requires_grad_() for t in range(t_max)]
  big = table.clone().requires_grad_()
  lattice = lattice_for(vocab, PrecompWeightFn(leaves, big))
  log_z, alphas = lattice._forward(cache=None, frames=frames, num_frames=nf,
                                   semiring=last_torch.semirings.Log)
  g_den = torch.stack(torch.autograd.grad(log_z.sum(), leaves), dim=1)      # [B,T,C,1+V]
  num = lattice._string_forward(cache=None, frames=frames, num_frames=nf, labels=lab,
                                num_labels=nl, semiring=last_torch.semirings.Log)
  (g_num,) = torch.autograd.grad(num.sum(), big)
  MG.remove_patches()
  grad = (g_den - g_num).numpy()
  print(f'{name}: reference gradients {time.time() - t0:.1f} s')
  out['log_z'] = log_z.detach().numpy()
  out['numerator'] = num.detach().numpy()
  out['alphas_sample'] = alphas.detach().numpy()[:, SAMPLE_FRAMES]
  out['grad_blank'] = grad[..., 0]
  out['grad_lexical_sample'] = grad[:, SAMPLE_FRAMES][..., 1:]
  out['grad_den_frame_sums'] = g_den.numpy().astype(np.float64).sum((2, 3))
  out['grad_num_frame_sums'] = g_num.numpy().astype(np.float64).sum((2, 3))
  del g_den, g_num, grad, alphas

  # ---- MaxTropical: distance + Viterbi arcs (autograd works as shipped) -----------------
  t0 = time.time()
  leaves = [table[:, t].clone().requires_grad_() for t in range(t_max)]
  lattice = lattice_for(vocab, PrecompWeightFn(leaves, table))
  vd, _ = lattice._forward(cache=None, frames=frames, num_frames=nf,
                           semiring=last_torch.semirings.MaxTropical)
  g = torch.stack(torch.autograd.grad(vd.sum(), leaves), dim=1)             # one-hot arcs
  out['maxtropical_dist'] = vd.detach().numpy()
  flat = g.reshape(batch, t_max, -1)
  arc = flat.argmax(-1)                                                     # state * (1+V) + col
  taken = flat.sum(-1)                                                      # 1 on real frames
  assert bool(((taken == 1) | (taken == 0)).all())
  out['viterbi_state'] = torch.where(taken > 0, arc // (1 + vocab), -1).numpy().astype(np.int32)
  out['viterbi_label'] = torch.where(taken > 0, arc % (1 + vocab), -1).numpy().astype(np.int32)
  print(f'{name}: reference MaxTropical {time.time() - t0:.1f} s, dist {out["maxtropical_dist"]}')
  np.savez_compressed(os.path.join(HERE, f'headline_{name}.npz'), **out)


def main():
  torch.set_num_threads(8)
  headline_case('bigram_v256_t1000', seed=2024, vocab=256, batch=2, t_max=1000,
                num_frames=[1000, 777], u=120, num_labels=[120, 57])


if __name__ == '__main__':
  main()
