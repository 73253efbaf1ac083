"""A complete training loop on the GNAT loss: synthetic utterances, Adam on the JointWeightFn /
SharedEmbCacher parameters, optional data parallelism.

    python examples/train_step.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        examples/train_step.py                          # one process per GPU, NCCL

Data parallelism follows DESIGN.md section 7: every rank takes a contiguous slice of the global
batch (utterances are independent: no collective inside the lattice kernels), and ONE flat
all-reduce carries [sum of losses, parameter gradients]."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import last_torch  # noqa: E402  (alias of last_torch_b200)
from last_torch_b200 import distributed as D  # noqa: E402


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = f'cuda:{local}'
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(dev))
  rank = dist.get_rank() if world > 1 else 0

  vocab, hidden, feat, batch, frames_t, labels_u = 128, 256, 80, 8 * world, 120, 20
  torch.manual_seed(0)                                  # same initial parameters on every rank
  lattice = last_torch.RecognitionLattice(
      context=last_torch.contexts.FullNGram(vocab_size=vocab, context_size=1),
      alignment=last_torch.alignments.FrameDependent(),
      weight_fn_cacher_factory=lambda c: last_torch.weight_fns.SharedEmbCacher(
          num_context_states=c.shape()[0], embedding_size=hidden, device=dev),
      weight_fn_factory=lambda c: last_torch.weight_fns.JointWeightFn(
          vocab_size=c.shape()[1], hidden_size=hidden, device=dev, embedding_size=hidden,
          feature_size=feat))
  params = list(lattice.parameters())
  opt = torch.optim.Adam(params, lr=1e-3)

  g = torch.Generator(device=dev).manual_seed(1)        # the same GLOBAL batch on every rank
  frames = torch.randn([batch, frames_t, feat], device=dev, generator=g)
  num_frames = torch.randint(frames_t // 2, frames_t + 1, [batch], device=dev, generator=g)
  labels = torch.randint(1, vocab + 1, [batch, labels_u], device=dev, generator=g)
  num_labels = torch.randint(1, labels_u + 1, [batch], device=dev, generator=g)

  for step in range(20):
    # every rank evaluates its shard; the loss sum and the parameter gradients come back reduced
    total, grads, _ = D.sharded_loss_and_grads(lattice, frames, num_frames, labels, num_labels)
    for p, gr in zip(params, grads):
      p.grad = gr / batch
    opt.step()
    if rank == 0 and step % 5 == 0:
      print(f'step {step:2d}  loss per utterance {float(total) / batch:9.3f}', flush=True)
  if rank == 0:
    with torch.no_grad():
      path, n, score = lattice.shortest_path(frames[:1], num_frames[:1])
      print('Viterbi score of utterance 0:', float(score[0]),
            ' entropy of its alignment distribution (nats):',
            float(lattice.entropy(frames[:1], num_frames[:1])[0]))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
