"""contexts.NextStateTable holding FullNGram(V, 1)'s transitions, T = 1000: Log loss and gradients
against the double build of the C oracle (the same lattice as FullNGram).
    python tools/check_table_long.py [V] [T]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: E402
from oracle import c_oracle  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
B, U = 2, 60
C = V + 1
rng = np.random.RandomState(3)
gen = torch.Generator().manual_seed(V + T)
table = torch.randn([B, T, C, 1 + V], generator=gen)
nf = np.array([T, int(0.7 * T)])
labels = rng.randint(1, V + 1, size=(B, U))
nl = np.array([U, U // 2])
tab = table.numpy()
loss64, gb64, gl64, _, _ = c_oracle.lattice_loss_and_grads(
    np.ascontiguousarray(tab[..., 0]), np.ascontiguousarray(tab[..., 1:]), nf, labels, nl, V, 1, -1,
    real='f64')
frames = torch.arange(T, device='cuda', dtype=torch.float32)[None, :, None].expand(B, T, 1)
full = lt.contexts.FullNGram(vocab_size=V, context_size=1)
for name, ctx in [('FullNGram', full),
                  ('NextStateTable', lt.contexts.NextStateTable(full.next_state_table().to(torch.int32)))]:
  leaf = table.cuda().requires_grad_()
  lattice = lt.RecognitionLattice(
      context=ctx, alignment=lt.alignments.FrameDependent(),
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(leaf),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  c = lambda x: torch.as_tensor(np.asarray(x), device='cuda').float()
  loss = lattice(frames=frames, num_frames=c(nf), labels=c(labels), num_labels=c(nl), cache=None)
  (gt,) = torch.autograd.grad(loss.sum(), leaf)
  gt = gt.cpu().numpy()
  print(f'{name:16s} V={V} T={T}: loss rel {np.abs(loss.detach().cpu().numpy() - loss64).max() / np.abs(loss64).max():.2e}'
        f'  grad_blank abs {np.abs(gt[..., 0] - gb64).max():.2e}  grad_lexical abs {np.abs(gt[..., 1:] - gl64).max():.2e}',
        flush=True)
