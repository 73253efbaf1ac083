"""Times lt_lattice_forward / backward per semiring on the configs[1] shape."""
import sys, torch
sys.path.insert(0, '.')
import last_torch_b200 as lt
from last_torch_b200 import ops, _native as N
B, T, V = 32, 1000, 256
C = V + 1
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
def timeit(fn, n=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(n): fn()
  e.record(); torch.cuda.synchronize()
  return s.elapsed_time(e) / n
W = B * T * C * (V + 1) * 4 / 1e9
for name, sr in [('Log', N.LOG), ('MaxTropical', N.MAXTROPICAL), ('Real', N.REAL)]:
  for flags, tag in [(0, 'fast'), (1, 'generic')]:
    if tag == 'generic' and name != 'Log': continue
    scale = 0.01 if name == 'Real' else 1.0
    ms = timeit(lambda: ops._lattice_forward_raw(sr, V, 1, -1, blank, lex * scale if name == 'Real' else lex, nf, flags, False, False))
    print(f'forward {name:12s} {tag:8s} {ms:7.3f} ms  {W / ms * 1e3:7.0f} GB/s')
