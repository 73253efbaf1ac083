import sys, torch
sys.path.insert(0, '.')
import last_torch_b200
from last_torch_b200.joint import linear_forward_raw, _Linear
from last_torch_b200 import _native as N
def t(fn, n=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(n): fn()
  e.record(); torch.cuda.synchronize()
  return s.elapsed_time(e) / n
for m, k, n in [(32000, 80, 512), (257, 512, 512), (32000, 512, 512)]:
  x = torch.randn([m, k], device='cuda'); w = torch.randn([n, k], device='cuda'); gy = torch.randn([m, n], device='cuda')
  gw = torch.empty_like(w)
  ws = torch.empty([int(N.lib().lt_linear_wgrad_workspace_bytes(m, k, n))], dtype=torch.uint8, device='cuda')
  def wg():
    N.check(N.lib().lt_linear_wgrad(N.ptr(gy), N.ptr(x), N.ptr(gw), m, k, n, N.ptr(ws), N.stream_ptr(x.device)), 'w')
  print((m, k, n), 'fwd mine %.3f ms torch %.3f ms | wgrad mine %.3f ms torch %.3f ms' % (
      t(lambda: linear_forward_raw(x, w)), t(lambda: torch.nn.functional.linear(x, w)), t(wg), t(lambda: gy.t() @ x)))
