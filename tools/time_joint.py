"""Times the JointWeightFn kernels (lt_joint_forward / lt_joint_backward) at the configs[1]
shape: N = B*T = 32000 frames, C = 257, H = 512, V = 256."""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import _native as N
from last_torch_b200.joint import _JointProjection

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32000
c, h, v = 257, 512, 256
g = torch.Generator(device='cuda').manual_seed(0)
pc = torch.randn([c, h], device='cuda', generator=g).requires_grad_()
pf = torch.randn([n, h], device='cuda', generator=g).requires_grad_()
wb = (torch.randn([1, h], device='cuda', generator=g) * 0.3).requires_grad_()
bb = torch.zeros([], device='cuda').requires_grad_()
wv = (torch.randn([v, h], device='cuda', generator=g) * 0.3).requires_grad_()
bv = torch.zeros([v], device='cuda').requires_grad_()
gb = torch.rand([n, c], device='cuda', generator=g) / c
gl = torch.rand([n, c, v], device='cuda', generator=g) / (c * v)


def timed(name):
  timer = []
  N.KERNEL_TIMER = timer
  blank, lexical = _JointProjection.apply(pc, pf, wb, bb, wv, bv)
  torch.autograd.grad([blank, lexical], [pc, pf, wb, bb, wv, bv], [gb, gl])
  torch.cuda.synchronize()
  N.KERNEL_TIMER = None
  return {k: s.elapsed_time(e) for k, s, e in timer}

for i in range(3):
  r = timed('x')
print({k: round(x, 3) for k, x in r.items()})
