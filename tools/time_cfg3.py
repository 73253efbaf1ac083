"""BASELINE.json configs[2]: FullNGram(context_size=2, vocab 64) = 4161 states,
FrameLabelDependent(2), MaxTropical shortest distance + Viterbi alignment.

    python tools/time_cfg3.py [B] [T] [flags...]
Times lt_lattice_forward (MaxTropical, with back-pointers), lt_viterbi_backtrace and
the Log forward on the same lattice; prints HBM GB/s for the 1*W algorithmic bytes.
"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
FLAGS = [int(x) for x in sys.argv[3:]] or [0, 1]
# geometry overrides (e.g. LT_V=256 LT_N=1 LT_K=2: FrameLabelDependent on the configs[1] bigram)
import os
V, n, k = int(os.environ.get('LT_V', 64)), int(os.environ.get('LT_N', 2)), int(os.environ.get('LT_K', 2))
C = sum(V ** i for i in range(n + 1))
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
W = B * T * C * (V + 1) * 4 / 1e9


def timeit(fn, n=5):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True)
  e = torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(n):
    fn()
  e.record()
  torch.cuda.synchronize()
  return s.elapsed_time(e) / n


ref = None
for kk, name in [(k, f'FrameLabelDependent({k})'), (-1, 'FrameDependent')]:
  for flags in FLAGS:
    out = ops._lattice_forward_raw(N.MAXTROPICAL, V, n, kk, blank, lex, nf, flags, False, True)
    ms = timeit(lambda: ops._lattice_forward_raw(N.MAXTROPICAL, V, n, kk, blank, lex, nf, flags,
                                                 False, True))
    msg = f'{name:24s} MaxTropical forward flags={flags}: {ms:8.3f} ms {W / ms * 1e3:7.0f} GB/s'
    if (kk, 'mt') in (ref or {}):
      r = ref[(kk, 'mt')]
      msg += (f'  vs first: dist {(out[0] - r[0]).abs().max().item():.1e} '
              f'backptr mismatches {(out[4] != r[4]).sum().item()}'
              + (f' termptr mismatches {(out[5] != r[5]).sum().item()}' if kk >= 1 else ''))
    else:
      ref = ref or {}
      ref[(kk, 'mt')] = out
    print(msg, flush=True)
    dist, _, alpha_final, _, backptr, termptr = out[:6]
    labels = torch.empty([B, T, max(kk, 0) + 1], dtype=torch.int32, device='cuda')
    states = torch.empty([B, T + 1], dtype=torch.int32, device='cuda')

    def trace():
      N.check(N.lib().lt_viterbi_backtrace(
          V, n, kk, N.ptr(backptr), N.ptr(termptr), N.ptr(alpha_final), N.ptr(nf), B, T,
          N.ptr(labels), N.ptr(states), None, None, None, N.stream_ptr(blank.device)), 'viterbi')
    ms = timeit(trace)
    print(f'{name:24s} Viterbi back-trace  flags={flags}: {ms:8.3f} ms', flush=True)
    out = ops._lattice_forward_raw(N.LOG, V, n, kk, blank, lex, nf, flags, kk >= 1, False)
    ms = timeit(lambda: ops._lattice_forward_raw(N.LOG, V, n, kk, blank, lex, nf, flags, kk >= 1,
                                                 False))
    msg = f'{name:24s} Log forward         flags={flags}: {ms:8.3f} ms {W / ms * 1e3:7.0f} GB/s'
    if (kk, 'log') in ref:
      r = ref[(kk, 'log')]
      msg += f'  vs first: dist rel {((out[0] - r[0]).abs() / r[0].abs()).max().item():.1e}'
    else:
      ref[(kk, 'log')] = out
    print(msg, flush=True)
