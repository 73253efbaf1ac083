"""N > 1 path on CPU: two gloo processes shard the batch by utterance, each
computes loss + gradients for its slice (with the oracle standing in for the
GPU kernels -- the data path has no collective), and a single all-reduce of
[loss sum, parameter gradients] reproduces the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _case():
  rng = np.random.RandomState(3)
  b, t, v, n, u = 5, 7, 4, 1, 3
  c = v + 1
  table = rng.randn(b, t, c, 1 + v)
  nf = np.array([7, 5, 6, 2, 7])
  labels = rng.randint(1, v + 1, (b, u))
  nl = np.array([3, 2, 3, 1, 0])
  return table, nf, labels, nl, v, n


def _local(table, nf, labels, nl, v, n, scale):
  from oracle import lattice_oracle as O
  if table.shape[0] == 0:
    return 0.0, np.zeros([3])
  loss, gb, gl = O.lattice_loss_and_grads(
      np.ascontiguousarray(table[..., 0]), np.ascontiguousarray(table[..., 1:]), nf, labels, nl,
      O.FullNGram(v, n))
  # a toy "parameter": arc weights = scale * table, so d loss / d scale = sum(grad * table)
  gparam = np.array([(gb * table[..., 0]).sum() + (gl * table[..., 1:]).sum(), gb.sum(), gl.sum()])
  return float(loss.sum()), gparam


def _worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from last_torch_b200 import distributed as D
  table, nf, labels, nl, v, n = _case()
  lo, hi = D.shard_range(table.shape[0], rank, world)
  loss, gparam = _local(table[lo:hi], nf[lo:hi], labels[lo:hi], nl[lo:hi], v, n, 1.0)
  total, grads = D.all_reduce_loss_and_grads(
      torch.tensor(loss, dtype=torch.float32),
      [torch.tensor(gparam, dtype=torch.float32), None])
  assert grads[1] is None
  out[rank] = (float(total), grads[0].numpy().copy(), (lo, hi))
  dist.barrier()
  dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
  sys.path.insert(0, ROOT)
  from last_torch_b200 import distributed as D
  for b in [0, 1, 5, 32, 33]:
    for w in [1, 2, 3, 8]:
      spans = [D.shard_range(b, r, w) for r in range(w)]
      assert spans[0][0] == 0 and spans[-1][1] == b
      assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
      sizes = [hi - lo for lo, hi in spans]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    D.shard_range(4, 2, 2)


@pytest.mark.timeout(120)
def test_two_rank_gloo_matches_single_process():
  world = 2
  port = _free_port()
  manager = mp.Manager()
  out = manager.dict()
  mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
  table, nf, labels, nl, v, n = _case()
  full_loss, full_g = _local(table, nf, labels, nl, v, n, 1.0)
  assert [out[r][2] for r in range(world)] == [(0, 3), (3, 5)]
  for r in range(world):
    np.testing.assert_allclose(out[r][0], full_loss, rtol=1e-5)
    np.testing.assert_allclose(out[r][1], full_g, rtol=1e-4, atol=1e-5)


# ---- sharded_loss_and_grads: lazily-shaped parameters, a rank with an empty shard ------------

class _FakeLattice(torch.nn.Module):
  """Stands in for RecognitionLattice on CPU: same call signature, parameters that only get
  their shapes on the first call (weight_fns.JointWeightFn without embedding / feature sizes)."""

  def __init__(self):
    super().__init__()
    sys.path.insert(0, ROOT)
    import last_torch_b200 as lt
    torch.manual_seed(0)
    self.weight_fn_cacher = lt.weight_fns.SharedEmbCacher(num_context_states=4, embedding_size=6)
    self.weight_fn = lt.weight_fns.JointWeightFn(vocab_size=3, hidden_size=5)

  def build_cache(self):
    return self.weight_fn_cacher()

  def forward(self, frames, num_frames, labels, num_labels, cache=None):
    cache = self.build_cache() if cache is None else cache
    blank, lexical = self.weight_fn(cache, frames[:, 0])
    return torch.tanh(blank).sum(-1) + (lexical ** 2).sum((-1, -2)) * num_frames


def _fake_inputs():
  g = torch.Generator().manual_seed(5)
  frames = torch.randn([1, 2, 7], generator=g)            # ONE utterance: rank 1 gets nothing
  return frames, torch.tensor([2.0]), torch.zeros([1, 1]), torch.zeros([1])


def _sharded_worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from last_torch_b200 import distributed as D
  lattice = _FakeLattice()
  assert not lattice.weight_fn.is_materialized()
  total, grads, loss = D.sharded_loss_and_grads(lattice, *_fake_inputs())
  names = [n for n, _ in lattice.named_parameters()]
  out[rank] = (float(total), {n: g.numpy().copy() for n, g in zip(names, grads)}, int(loss.numel()))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_loss_with_lazy_parameters_and_an_empty_shard():
  """parameters created on the first call must not be dropped from the first step,
  and a rank with no utterances must still join the all-reduce with the same bucket layout."""
  world = 2
  port = _free_port()
  out = mp.Manager().dict()
  mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
  lattice = _FakeLattice()
  loss = lattice(*_fake_inputs())
  params = dict(lattice.named_parameters())
  ref = torch.autograd.grad(loss.sum(), list(params.values()))
  assert out[0][2] == 1 and out[1][2] == 0
  for r in range(world):
    np.testing.assert_allclose(out[r][0], float(loss.sum()), rtol=1e-6)
    assert set(out[r][1]) == set(params)
    for (name, _), g in zip(params.items(), ref):
      np.testing.assert_allclose(out[r][1][name], g.numpy(), rtol=1e-5, atol=1e-7, err_msg=name)
    assert np.abs(out[r][1]['weight_fn.context_projection.weight']).max() > 0
