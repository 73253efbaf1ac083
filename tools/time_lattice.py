"""Times lt_lattice_forward / lt_lattice_backward per semiring and kernel generation
on a bigram FrameDependent shape and cross-checks the generations against each other.

    python tools/time_lattice.py [B] [T] [V]
flags: 0 = second-generation fast path, two 256-thread CTAs per SM (lattice_fast2.cu),
4 = the same with both utterances in one 512-thread CTA, 2 = first-generation fast path
(lattice_fast.cu), 1 = generic kernels.
"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt  # noqa: F401
from last_torch_b200 import ops, _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
V = int(sys.argv[3]) if len(sys.argv) > 3 else 256
C = V + 1
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
lex_real = lex * 0.01 + 1.0 / C
blank_real = blank * 0.01 + 1.0 / C
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
W = B * T * C * (V + 1) * 4 / 1e9


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


def backward(sr, bl, lx, dist, alphas, flags, gb, gl):
  gd = torch.ones([B], device='cuda')
  N.check(N.lib().lt_lattice_backward(
      sr, V, 1, -1, N.ptr(bl), N.ptr(lx), N.ptr(nf), B, T, N.ptr(alphas), None, N.ptr(dist),
      N.ptr(gd), N.ptr(gb), N.ptr(gl), None, flags, N.stream_ptr(bl.device)), 'bwd')


ref = {}
gb = torch.empty_like(blank)
gl = torch.empty_like(lex)
for name, sr in [('Log', N.LOG), ('MaxTropical', N.MAXTROPICAL), ('Real', N.REAL)]:
  bl, lx = (blank_real, lex_real) if name == 'Real' else (blank, lex)
  for flags, tag in [(0, 'v2'), (4, 'v2pair'), (2, 'v1'), (1, 'generic')]:
    if tag == 'generic' and name != 'Log':
      continue
    want_bp = name == 'MaxTropical'
    out = ops._lattice_forward_raw(sr, V, 1, -1, bl, lx, nf, flags, False, want_bp)
    ms = timeit(lambda: ops._lattice_forward_raw(sr, V, 1, -1, bl, lx, nf, flags, False, want_bp))
    msg = f'forward  {name:12s} {tag:8s} {ms:7.3f} ms  {W / ms * 1e3:7.0f} GB/s'
    key = name
    if key in ref:
      d = (out[0] - ref[key][0]).abs().max().item() / max(1.0, ref[key][0].abs().max().item())
      a = (out[1] - ref[key][1]).abs().max().item()
      msg += f'   vs v2: dist rel {d:.2e} alphas abs {a:.2e}'
      if want_bp:
        msg += f' backptr mismatches {(out[4] != ref[key][4]).sum().item()}'
    else:
      ref[key] = out
    print(msg, flush=True)
    if name != 'MaxTropical' and tag != 'generic':
      dist, alphas = out[0], out[1]
      backward(sr, bl, lx, dist, alphas, flags, gb, gl)
      ms = timeit(lambda: backward(sr, bl, lx, dist, alphas, flags, gb, gl))
      msg = f'backward {name:12s} {tag:8s} {ms:7.3f} ms  {2 * W / ms * 1e3:7.0f} GB/s'
      if (key, 'g') in ref:
        rb, rl = ref[(key, 'g')]
        msg += (f'   vs v2: gblank {(gb - rb).abs().max().item():.2e} '
                f'glex {(gl - rl).abs().max().item():.2e} framesum {gl[0, T // 2].sum().item() + gb[0, T // 2].sum().item():.5f}')
      else:
        ref[(key, 'g')] = (gb.clone(), gl.clone())
      print(msg, flush=True)
