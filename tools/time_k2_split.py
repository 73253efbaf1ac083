"""Times lt_lattice_backward (Log, bigram vocab 256, B=32, T=1000) writing fp32 gradients and
split rows (LT_FLAG_GRAD_SPLIT), and checks that hi + lo of the split rows equals the fp32 rows."""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N

B, T, V = 32, 1000, 256
C = V + 1
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
out = ops._lattice_forward_raw(N.LOG, V, 1, -1, blank, lex, nf, 0, False, False)
dist, alphas = out[0], out[1]
gd = torch.ones([B], device='cuda')
gb = torch.empty_like(blank)
gl = torch.empty_like(lex)
gs = torch.empty_like(lex)


def bwd(flags, dst):
  N.check(N.lib().lt_lattice_backward(
      N.LOG, V, 1, -1, N.ptr(blank), N.ptr(lex), N.ptr(nf), B, T, N.ptr(alphas), None,
      N.ptr(dist), N.ptr(gd), N.ptr(gb), N.ptr(dst), None, flags, N.stream_ptr(blank.device)), 'bwd')


def timeit(fn, n=10):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(n):
    fn()
  e.record()
  torch.cuda.synchronize()
  return s.elapsed_time(e) / n


for rep in range(2):
  print('fp32  %.3f ms' % timeit(lambda: bwd(0, gl)), ' split %.3f ms' % timeit(lambda: bwd(N.FLAG_GRAD_SPLIT, gs)))
rows = gs.view(torch.bfloat16).reshape(B, T, C, 2, V).float()
rec = rows[..., 0, :] + rows[..., 1, :]
print('max |hi + lo - fp32| / max|fp32| = %.2e' % ((rec - gl).abs().max() / gl.abs().max()).item())
