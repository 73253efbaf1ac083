"""A/B of lt_joint_forward at the configs[1] shape (N = 32000 frames, C = 257, H = 512, V = 256):
tanh operand in tensor memory (default) vs in shared memory (LT_JOINT_FWD_SS=1); also checks the
two against each other and a float64 evaluation of sampled rows."""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import _native as N
from last_torch_b200.joint import joint_forward_raw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32000
c, h, v = 257, 512, 256
g = torch.Generator(device='cuda').manual_seed(0)
pc = torch.randn([c, h], device='cuda', generator=g)
pf = torch.randn([n, h], device='cuda', generator=g)
wb = torch.randn([1, h], device='cuda', generator=g) * 0.3
bb = torch.full([], 0.25, device='cuda')
wv = torch.randn([v, h], device='cuda', generator=g) * 0.3
bv = torch.randn([v], device='cuda', generator=g)


def run(ss, reps=5, cluster=0):
  with N.option('LT_JOINT_FWD_SS', ss), N.option('LT_JOINT_FWD_CLUSTER', cluster):
    out = joint_forward_raw(pc, pf, wb, bb, wv, bv)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      joint_forward_raw(pc, pf, wb, bb, wv, bv)
      e.record()
      torch.cuda.synchronize()
      ts.append(s.elapsed_time(e))
  return out, sorted(ts)[len(ts) // 2]


(b_ts, l_ts), t_ts = run(0)
(b_ss, l_ss), t_ss = run(1)
(b_1, l_1), t_1 = run(0, cluster=1)
print({'ts_cluster1_ms': round(t_1, 3), 'equal': bool(torch.equal(l_1, l_ts) and torch.equal(b_1, b_ts))})
rows = torch.randint(0, n, [64], device='cuda', generator=g)
j = torch.tanh(pc.double()[None] + pf.double()[rows][:, None])
rl = j @ wv.double().T + bv.double()
rb = j @ wb.double()[0] + bb.double()
scale = float(rl.abs().max())
print({'ts_ms': round(t_ts, 3), 'ss_ms': round(t_ss, 3),
       'ts_vs_ss_lexical': float((l_ts - l_ss).abs().max()) / scale,
       'ts_vs_ss_blank': float((b_ts - b_ss).abs().max()),
       'ts_err': float((l_ts[rows].double() - rl).abs().max()) / scale,
       'ss_err': float((l_ss[rows].double() - rl).abs().max()) / scale,
       'ts_err_blank': float((b_ts[rows].double() - rb).abs().max()) / float(rb.abs().max()),
       'ss_err_blank': float((b_ss[rows].double() - rb).abs().max()) / float(rb.abs().max())})
