"""One forward + backward of the trigram (FullNGram(64, 2)) FrameLabelDependent(2) Log loss for
ncu: lattice_forward_cols<Log, FLD> and lattice_backward_rows_fld<Log>.
    python tools/prof_trigram_fld.py [B] [T]"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
V, n, k = 64, 2, 2
C = 1 + V + V * V
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
gd = torch.ones([B], device='cuda')
gb = torch.empty_like(blank)
gl = torch.empty_like(lex)
for _ in range(2):
  dist, alphas, _, levels, _, _, an = ops._lattice_forward_raw(N.LOG, V, n, k, blank, lex, nf, 0,
                                                               True, False, norm=True)
  N.check(N.lib().lt_lattice_backward_norm(
      N.LOG, V, n, k, N.ptr(blank), N.ptr(lex), N.ptr(nf), B, T, N.ptr(alphas), N.ptr(levels),
      N.ptr(dist), N.ptr(gd), N.ptr(gb), N.ptr(gl), None, N.ptr(an), 0,
      N.stream_ptr(blank.device)), 'bwd')
torch.cuda.synchronize()
