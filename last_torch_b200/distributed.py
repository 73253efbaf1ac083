"""Multi-GPU data parallelism for the lattice path: utterances are independent
units (every operation in lattices.py is batched over the leading dims), so the
batch is split contiguously across ranks -- one process per GPU -- and the ONLY
exchange is one all-reduce of the summed loss and of the weight-function
parameter gradients (NCCL over NVLink on the GPU box; gloo in the CPU tests).
Per-arc gradients never leave the rank that owns the utterance.
"""

from __future__ import annotations

from typing import Iterable, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(batch_size: int, rank: int, world_size: int) -> tuple[int, int]:
  """Contiguous [lo, hi) slice of the batch owned by `rank`; sizes differ by at
  most one and every utterance is owned by exactly one rank."""
  if world_size <= 0 or not 0 <= rank < world_size:
    raise ValueError(f'invalid rank {rank} for world_size {world_size}')
  base, extra = divmod(batch_size, world_size)
  lo = rank * base + min(rank, extra)
  return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world_size: int):
  """Slices every tensor along dim 0 with shard_range."""
  lo, hi = shard_range(tensors[0].shape[0], rank, world_size)
  return [t[lo:hi] for t in tensors]


_VERIFIED_LAYOUTS = set()


def _check_same_layout(numel: int, device, group) -> None:
  """All ranks must flatten the same number of elements into the all-reduce (a rank whose
  parameters were never materialised would otherwise corrupt or hang the collective).  Checked
  once per (group, layout) with one tiny MAX all-reduce of [numel, -numel]."""
  key = (id(group), numel)
  if key in _VERIFIED_LAYOUTS:
    return
  probe = torch.tensor([numel, -numel], dtype=torch.int64, device=device)
  dist.all_reduce(probe, op=dist.ReduceOp.MAX, group=group)
  hi, lo = int(probe[0]), -int(probe[1])
  if hi != lo:
    raise RuntimeError(f'ranks disagree on the size of the loss / gradient bucket ({lo} .. {hi} '
                       'elements): every rank must hold the same (materialised) parameters')
  _VERIFIED_LAYOUTS.add(key)


def all_reduce_loss_and_grads(loss_sum: torch.Tensor, grads: Iterable[Optional[torch.Tensor]],
                              group=None):
  """One flat all-reduce(sum) of [loss_sum, *grads]; returns (loss_sum, grads).

  `grads` are parameter gradients (same shapes on every rank); None entries are
  passed through.  A single bucket keeps it to one collective launch: the payload
  is O(#parameters), tiny next to the per-rank lattice work."""
  grads = list(grads)
  if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
    return loss_sum, grads
  live = [g for g in grads if g is not None]
  flat = torch.cat([loss_sum.reshape(-1).to(torch.float32)] +
                   [g.reshape(-1).to(torch.float32) for g in live])
  _check_same_layout(flat.numel(), flat.device, group)
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  out_loss = flat[:loss_sum.numel()].reshape(loss_sum.shape).to(loss_sum.dtype)
  offset = loss_sum.numel()
  out = []
  for g in grads:
    if g is None:
      out.append(None)
      continue
    out.append(flat[offset:offset + g.numel()].reshape(g.shape).to(g.dtype))
    offset += g.numel()
  return out_loss, out


def _materialize_parameters(lattice, frames, cache) -> None:
  """Gives lazily-shaped parameters (weight_fns.JointWeightFn without embedding_size /
  feature_size) their shapes on a rank that evaluates nothing."""
  fn = getattr(lattice, 'weight_fn', None)
  if fn is None or not hasattr(fn, 'materialize') or fn.is_materialized():
    return
  if cache is None:
    cache = lattice.build_cache()
  fn.materialize(cache.shape[-1], frames.shape[-1])


def local_loss_and_grads(lattice, frames, num_frames, labels, num_labels, cache=None):
  """Loss and parameter gradients of THIS rank's utterances, no communication.
  Returns (loss sum, list of gradients in lattice.parameters() order, losses [b]).
  The parameter list is read AFTER the forward pass, when lazily-shaped parameters exist; a
  rank without utterances materialises them explicitly and contributes zeros."""
  if frames.shape[0] > 0:
    loss = lattice(frames=frames, num_frames=num_frames, labels=labels, num_labels=num_labels,
                   cache=cache)
    params = [p for p in lattice.parameters() if p.requires_grad]
    total = loss.sum()
    grads = list(torch.autograd.grad(total, params, allow_unused=True)) if params else []
  else:
    _materialize_parameters(lattice, frames, cache)
    params = [p for p in lattice.parameters() if p.requires_grad]
    loss = frames.new_zeros([0])
    total = frames.new_zeros([])
    grads = [None] * len(params)
  grads = [torch.zeros_like(p) if g is None else g for g, p in zip(grads, params)]
  return total.detach(), grads, loss


def sharded_loss_and_grads(lattice, frames, num_frames, labels, num_labels, cache=None,
                           group=None):
  """Data-parallel GNAT loss: each rank evaluates `lattice(...)` on its slice of
  the global batch; the summed loss and the parameter gradients are all-reduced.
  Returns (global loss sum, list of reduced parameter gradients, local losses)."""
  rank = dist.get_rank(group) if dist.is_initialized() else 0
  world = dist.get_world_size(group) if dist.is_initialized() else 1
  f, nf, lab, nl = shard_batch([frames, num_frames, labels, num_labels], rank, world)
  total, grads, loss = local_loss_and_grads(lattice, f, nf, lab, nl, cache)
  total, grads = all_reduce_loss_and_grads(total, grads, group)
  return total, grads, loss
