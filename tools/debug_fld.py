import sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import c_oracle
from last_torch_b200 import _native as N
import last_torch_b200 as lt
from test_gpu_lattice import make_lattice, frames_for, cuda

b, t, v, k, u = 3, int(sys.argv[1]) if len(sys.argv) > 1 else 300, 256, 2, 40
rng = np.random.RandomState(21)
gen = torch.Generator().manual_seed(212)
table = torch.randn([b, t, v + 1, 1 + v], generator=gen)
nf = np.array([t, t * 177 // 300, t * 251 // 300])
labels = rng.randint(1, v + 1, size=(b, u))
nl = np.array([40, 13, 0])
tab = table.numpy()
loss64, gb64, gl64, logz64, _ = c_oracle.lattice_loss_and_grads(
    np.ascontiguousarray(tab[..., 0]), np.ascontiguousarray(tab[..., 1:]), nf, labels, nl, v, 1, k, real='f64')
want = np.concatenate([gb64[..., None], gl64], -1)
for generic in (0, 1):
  with N.option('LT_FLD_GENERIC', generic):
    leaf = table.cuda().requires_grad_()
    lattice = make_lattice(v, 1, k, leaf)
    loss = lattice(frames=frames_for(b, t), num_frames=cuda(nf), labels=cuda(labels), num_labels=cuda(nl), cache=None)
    (gt,) = torch.autograd.grad(loss.sum(), leaf)
    gt = gt.cpu().numpy()
    err = np.abs(gt - want)
    big = np.abs(want) > 1e-4 * np.abs(want).max()
    rel = np.where(big, err / np.maximum(np.abs(want), 1e-30), 0)
    idx = np.unravel_index(np.argsort(rel.ravel())[-8:], rel.shape)
    print('generic' if generic else 'fast', 'loss err', np.abs(loss.detach().cpu().numpy() - loss64) / np.abs(loss64), 'max abs', err.max(), 'max rel', rel.max())
    for i in range(8):
      ix = tuple(a[i] for a in idx)
      print('  ', ix, 'want', want[ix], 'got', gt[ix], 'rel', rel[ix])
    # denominator only
    dist, _ = lattice._forward(cache=None, frames=frames_for(b, t), num_frames=cuda(nf), semiring=lt.semirings.Log)
    print('  logZ rel err', np.abs(dist.detach().cpu().numpy() - logz64) / np.abs(logz64))
