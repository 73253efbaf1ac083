"""Latency-bound shapes of the bigram fast path (configs[1] geometry): forward and backward at
B = 8 / 32 / 48 and a ragged B = 32 batch, renormalised Log recursion.
    python tools/time_chain.py [B ...]      # e.g. 16 24 40r (r = ragged)"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N

T, V = 1000, 256
C = V + 1


def timeit(fn, n=10):
  for _ in range(3):
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


g = torch.Generator(device='cuda').manual_seed(0)
CASES = [(8, False), (32, False), (32, True), (48, False)]
if len(sys.argv) > 1:      # e.g. 16 24 40r: batch sizes, r = ragged
  CASES = [(int(a.rstrip('r')), a.endswith('r')) for a in sys.argv[1:]]
for B, ragged in CASES:
  blank = torch.randn([B, T, C], device='cuda', generator=g)
  lex = torch.randn([B, T, C, V], device='cuda', generator=g)
  nf = (torch.randint(T // 2, T + 1, [B], device='cuda', generator=g, dtype=torch.int32) if ragged
        else torch.full([B], T, dtype=torch.int32, device='cuda'))
  gd = torch.ones([B], device='cuda')
  gb = torch.empty_like(blank)
  gl = torch.empty_like(lex)
  out = ops._lattice_forward_raw(N.LOG, V, 1, -1, blank, lex, nf, 0, False, False, norm=True)
  dist, alphas, an = out[0], out[1], out[6]
  f = timeit(lambda: ops._lattice_forward_raw(N.LOG, V, 1, -1, blank, lex, nf, 0, False, False,
                                              norm=True))

  def bwd():
    N.check(N.lib().lt_lattice_backward_norm(
        N.LOG, V, 1, -1, N.ptr(blank), N.ptr(lex), N.ptr(nf), B, T, N.ptr(alphas), None,
        N.ptr(dist), N.ptr(gd), N.ptr(gb), N.ptr(gl), None, N.ptr(an), 0,
        N.stream_ptr(blank.device)), 'bwd')
  k = timeit(bwd)
  # alternating K1, K2 (as in a training step, nothing else running): per-kernel CUDA events
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * 10)]
  for i in range(10):
    ev[3 * i].record()
    ops._lattice_forward_raw(N.LOG, V, 1, -1, blank, lex, nf, 0, False, False, norm=True)
    ev[3 * i + 1].record()
    bwd()
    ev[3 * i + 2].record()
  torch.cuda.synchronize()
  fa = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(2, 10)) / 8
  ka = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(2, 10)) / 8
  real = float(nf.sum()) * C
  print(f'B={B:3d} ragged={int(ragged)}: forward {f:6.3f} ms  backward {k:6.3f} ms  '
        f'{real / (f + k) / 1e6:7.1f} G real frames*states/s (K1+K2)  alternating: {fa:6.3f} / {ka:6.3f} ms  '
        f'dist[0]={float(dist[0]):.4f}',
        flush=True)
  del blank, lex, gb, gl
