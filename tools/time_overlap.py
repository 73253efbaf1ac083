"""Which numerator kernel costs K1 its 0.2 ms inside the step?  K1 on the current stream with, on
a side stream 40 us behind it: nothing / the gather / the label-lattice forward / both.
    python tools/time_overlap.py"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N

B, T, V, U = 32, 1000, 256, 120
C, U1 = V + 1, U + 1
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
labels = torch.randint(1, V + 1, [B, U], device='cuda', generator=g, dtype=torch.int32)
nl = torch.full([B], U, dtype=torch.int32, device='cuda')
states, next_labels, _ = ops.walk_states(labels, nl, V, 1)
dev = blank.device
side = torch.cuda.Stream()
L = N.lib()
bw = torch.empty([B, T, U1], device='cuda')
lw = torch.empty([B, T, U1], device='cuda')
dist = torch.empty([B], device='cuda')
alphas = torch.empty([B, T, U1], device='cuda')
aexp = torch.empty([B, T, U1], dtype=torch.int32, device='cuda')
dnorm = torch.empty([B, 2], dtype=torch.int32, device='cuda')


def numerator(gather, forward):
  with torch.cuda.stream(side):
    N.check(L.lt_stream_delay(40000, N.stream_ptr(dev)), 'delay')
    if gather:
      N.check(L.lt_string_gather(V, C, N.ptr(blank), N.ptr(lex), N.ptr(states), N.ptr(next_labels),
                                 B, T, U1, N.ptr(bw), N.ptr(lw), N.stream_ptr(dev)), 'gather')
    if forward:
      N.check(L.lt_string_forward_norm(N.LOG, -1, N.ptr(bw), N.ptr(lw), N.ptr(nf), N.ptr(nl), B, T,
                                       U1, N.ptr(dist), N.ptr(alphas), None, N.ptr(aexp),
                                       N.ptr(dnorm), N.stream_ptr(dev)), 'string forward')


numerator(True, True)
torch.cuda.synchronize()
for name, ga, fo in [('K1 alone', False, False), ('K1 + gather', True, False),
                     ('K1 + label-lattice forward', False, True), ('K1 + both', True, True)]:
  ts = []
  for i in range(8):
    cur = torch.cuda.current_stream()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    ready = torch.cuda.Event()
    ready.record(cur)
    s.record()
    ops._lattice_forward_raw(N.LOG, V, 1, -1, blank, lex, nf, 0, False, False, norm=True)
    e.record()
    if ga or fo:
      side.wait_event(ready)
      numerator(ga, fo)
    torch.cuda.synchronize()
    if i >= 2:
      ts.append(s.elapsed_time(e))
  print(f'{name:28s} K1 {sum(ts) / len(ts):6.3f} ms', flush=True)
