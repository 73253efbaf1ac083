"""Multi-rank correctness ON GPUS: two ranks shard a batch by utterance, each runs the real
kernels through RecognitionLattice.forward (JointWeightFn, tensor-core path) inside
distributed.sharded_loss_and_grads, and the all-reduced loss + parameter gradients must equal the
single-GPU result on the concatenated batch.  With two or more GPUs the ranks use NCCL (one GPU
each); on a one-GPU box both ranks share cuda:0 and the all-reduce goes through gloo -- the
sharding, the empty-/uneven-shard logic and the kernels are the same."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _problem(device):
  sys.path.insert(0, ROOT)
  import last_torch_b200 as lt
  torch.manual_seed(11)
  v, h, b, t, u = 128, 128, 3, 24, 6                 # uneven shards: 2 + 1 utterances
  lattice = lt.RecognitionLattice(
      context=lt.contexts.FullNGram(vocab_size=v, context_size=1),
      alignment=lt.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=24, device=device),
      weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=h, device=device))     # lazily-shaped inputs
  g = torch.Generator().manual_seed(3)
  frames = torch.randn([b, t, 16], generator=g).to(device)
  num_frames = torch.tensor([24, 17, 9], device=device)
  labels = torch.randint(1, v + 1, [b, u], generator=g).to(device)
  num_labels = torch.tensor([6, 4, 2], device=device)
  return lattice, (frames, num_frames, labels, num_labels)


def _worker(rank, world, port, ngpu, out):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  device = f'cuda:{rank % ngpu}'
  torch.cuda.set_device(device)
  backend = 'nccl' if ngpu >= world else 'gloo'
  dist.init_process_group(backend, rank=rank, world_size=world)
  from last_torch_b200 import distributed as D
  lattice, inputs = _problem(device)
  total, grads, loss = D.sharded_loss_and_grads(lattice, *inputs)
  torch.cuda.synchronize()
  names = [n for n, _ in lattice.named_parameters()]
  out[rank] = (backend, float(total), {n: g.cpu().numpy() for n, g in zip(names, grads)},
               loss.detach().cpu().numpy())
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_gpu_on_the_concatenated_batch():
  world = 2
  ngpu = torch.cuda.device_count()
  out = mp.Manager().dict()
  mp.spawn(_worker, args=(world, _free_port(), ngpu, out), nprocs=world, join=True)
  lattice, inputs = _problem('cuda:0')
  loss = lattice(frames=inputs[0], num_frames=inputs[1], labels=inputs[2], num_labels=inputs[3])
  params = dict(lattice.named_parameters())
  ref = torch.autograd.grad(loss.sum(), list(params.values()))
  full = loss.detach().cpu().numpy()
  np.testing.assert_array_equal(out[0][3], full[:2])
  np.testing.assert_array_equal(out[1][3], full[2:])
  print('backend:', out[0][0])
  for r in range(world):
    np.testing.assert_allclose(out[r][1], float(loss.sum()), rtol=1e-6)
    for (name, _), g in zip(params.items(), ref):
      want = g.cpu().numpy()
      scale = np.abs(want).max() + 1e-12
      # per-rank partial sums are added in a different order than the one-GPU atomics
      assert np.abs(out[r][2][name] - want).max() <= 2e-5 * scale, name
