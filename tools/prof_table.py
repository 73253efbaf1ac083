"""One forward + one backward of the NextStateTable cluster kernels at the configs[1] geometry
(for ncu):  ncu --set full -k regex:table_ python tools/prof_table.py [B] [T]"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt
from last_torch_b200 import ops, _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
V, C = 256, 257
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g).requires_grad_()
lex = torch.randn([B, T, C, V], device='cuda', generator=g).requires_grad_()
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
full = lt.contexts.FullNGram(vocab_size=V, context_size=1)
table = lt.contexts.NextStateTable(full.next_state_table().to(torch.int32))
for _ in range(2):
  dist, _ = ops.TableLatticeForward.apply(blank, lex, nf, table, N.LOG, -1)
  torch.autograd.grad(dist.sum(), [blank, lex])
torch.cuda.synchronize()
print('ok', float(dist.sum()))
