"""Log loss + gradient on a context_size-2 lattice (FullNGram(64, 2), 4161 states):
forward (thread-per-column path) + backward kernel, FrameDependent and FrameLabelDependent(2).
    python tools/time_trigram_lossgrad.py [B] [T]"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 250
V, n = 64, 2
C = 1 + V + V * V
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
gd = torch.ones([B], device='cuda')
gb = torch.empty_like(blank)
gl = torch.empty_like(lex)
W = B * T * C * (V + 1) * 4 / 1e9


def timeit(fn, reps=5):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True)
  e = torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(reps):
    fn()
  e.record()
  torch.cuda.synchronize()
  return s.elapsed_time(e) / reps


for k, name in [(-1, 'FrameDependent'), (2, 'FrameLabelDependent(2)')]:
  ref = None
  for flags in [0, 1]:
    # renormalised state where the kernels support it (LT_NO_NORM=1: the plain recursion)
    out = ops._lattice_forward_raw(N.LOG, V, n, k, blank, lex, nf, flags, k >= 1, False, norm=True)
    dist, alphas, _, levels, _, _, an = out[:7]
    fms = timeit(lambda: ops._lattice_forward_raw(N.LOG, V, n, k, blank, lex, nf, flags, k >= 1,
                                                  False, norm=True))

    def bwd():
      N.check(N.lib().lt_lattice_backward_norm(
          N.LOG, V, n, k, N.ptr(blank), N.ptr(lex), N.ptr(nf), B, T, N.ptr(alphas), N.ptr(levels),
          N.ptr(dist), N.ptr(gd), N.ptr(gb), N.ptr(gl), None, N.ptr(an), flags,
          N.stream_ptr(blank.device)), 'bwd')
    bwd()
    bms = timeit(bwd)
    msg = (f'{name:24s} flags={flags} norm={int(an is not None)}: forward {fms:8.3f} ms ({W / fms * 1e3:6.0f} GB/s)  '
           f'backward {bms:8.3f} ms ({2 * W / bms * 1e3:6.0f} GB/s)  frame-sum '
           f'{float(gl[0, T // 2].sum() + gb[0, T // 2].sum()):.5f}')
    if ref is not None:
      msg += f'  vs flags=0: glex {float((gl - ref[1]).abs().max()):.2e} gblank {float((gb - ref[0]).abs().max()):.2e}'
    else:
      ref = (gb.clone(), gl.clone())
    print(msg, flush=True)
