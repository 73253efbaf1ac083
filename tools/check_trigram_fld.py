"""context_size-2 lattices at T = 200 against the double build of the C oracle: loss and every
gradient entry, TMA kernels (flags 0) and generic kernels (flags 1), FrameDependent and
FrameLabelDependent(k).  Prints the maximum absolute gradient error of each.
    python tools/check_trigram_fld.py [V] [T] [B]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import last_torch_b200 as lt  # noqa: E402
from oracle import c_oracle  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = 2
C = 1 + V + V * V
U = 30
rng = np.random.RandomState(5)
gen = torch.Generator().manual_seed(V)
table = torch.randn([B, T, C, 1 + V], generator=gen)
nf = np.array([T] + [int(x) for x in rng.randint(T // 2, T + 1, size=B - 1)])
labels = rng.randint(1, V + 1, size=(B, U))
nl = rng.randint(0, U + 1, size=B)
tab = table.numpy()
frames = torch.arange(T, device='cuda', dtype=torch.float32)[None, :, None].expand(B, T, 1)


def cuda(x):
  return torch.as_tensor(np.asarray(x), device='cuda').float()


for k in [-1, 2, 3]:
  loss64, gb64, gl64, _, _ = c_oracle.lattice_loss_and_grads(
      np.ascontiguousarray(tab[..., 0]), np.ascontiguousarray(tab[..., 1:]), nf, labels, nl, V, n,
      k, real='f64')
  for flags in [0, 1]:
    leaf = table.cuda().requires_grad_()
    alignment = (lt.alignments.FrameDependent() if k < 0 else
                 lt.alignments.FrameLabelDependent(max_expansions=k))
    lattice = lt.RecognitionLattice(
        context=lt.contexts.FullNGram(vocab_size=V, context_size=n), alignment=alignment,
        weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(leaf),
        weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
    lattice.kernel_flags = flags
    loss = lattice(frames=frames, num_frames=cuda(nf), labels=cuda(labels), num_labels=cuda(nl),
                   cache=None)
    (gt,) = torch.autograd.grad(loss.sum(), leaf)
    gt = gt.cpu().numpy()
    el = np.abs(loss.detach().cpu().numpy() - loss64).max() / np.abs(loss64).max()
    eb = np.abs(gt[..., 0] - gb64).max()
    eg = np.abs(gt[..., 1:] - gl64).max()
    print(f'V={V} T={T} k={k:2d} flags={flags}: loss rel {el:.2e}  grad_blank abs {eb:.2e}  '
          f'grad_lexical abs {eg:.2e}', flush=True)
