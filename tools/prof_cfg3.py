"""One launch of the configs[2] forward (FrameLabelDependent(2), MaxTropical) and one of the
FrameDependent variant, for ncu."""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N
B, T, V, n = 32, 200, 64, 2
C = 1 + V + V * V
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
for k in (2, -1):
  ops._lattice_forward_raw(N.MAXTROPICAL, V, n, k, blank, lex, nf, 0, False, True)
torch.cuda.synchronize()
