"""SURVEY section 8(f) "next" rows, timed at the BASELINE configs[1] geometry.

    python tools/time_next_rows.py [B] [T]
N2: contexts.NextStateTable lattice kernels (csrc/lattice_table.cu) on FullNGram(256, 1)'s own
    table, Log forward and Log backward, beside the FullNGram kernels on the same weights
    (also checks that the two agree);
N4: hat_normalize / log_softmax_normalize (csrc/normalize.cu), forward and backward, as HBM GB/s
    (forward = read W + write W, backward = read 2 W + write W).
"""
import sys
import torch
sys.path.insert(0, '.')
import last_torch_b200 as lt
from last_torch_b200 import ops, _native as N

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
V, n = 256, 1
C = V + 1
g = torch.Generator(device='cuda').manual_seed(0)
blank = torch.randn([B, T, C], device='cuda', generator=g)
lex = torch.randn([B, T, C, V], device='cuda', generator=g)
nf = torch.full([B], T, dtype=torch.int32, device='cuda')
W = B * T * C * (V + 1) * 4 / 1e9


def timeit(fn, n=3):
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


full = lt.contexts.FullNGram(vocab_size=V, context_size=n)
table = lt.contexts.NextStateTable(full.next_state_table().to(torch.int32))

# ---- N2 ------------------------------------------------------------------
ref = ops._lattice_forward_raw(N.LOG, V, n, -1, blank, lex, nf, 0, False, False)
ms = timeit(lambda: ops._lattice_forward_raw(N.LOG, V, n, -1, blank, lex, nf, 0, False, False))
print(f'FullNGram      Log forward : {ms:8.3f} ms {W / ms * 1e3:7.0f} GB/s (1 W = {W:.2f} GB)')
out = ops._table_forward_raw(N.LOG, -1, table, blank, lex, nf, False, False)
ms = timeit(lambda: ops._table_forward_raw(N.LOG, -1, table, blank, lex, nf, False, False))
rel = ((out[0] - ref[0]).abs() / ref[0].abs()).max().item()
print(f'NextStateTable Log forward : {ms:8.3f} ms {W / ms * 1e3:7.0f} GB/s   dist rel diff {rel:.1e}')

for name, ctx_obj in [('FullNGram', full), ('NextStateTable', table)]:
  b_ = blank.clone().requires_grad_()
  l_ = lex.clone().requires_grad_()
  if name == 'FullNGram':
    dist, _ = ops.LatticeForward.apply(b_, l_, nf, N.LOG, V, n, -1, 0)
  else:
    dist, _ = ops.TableLatticeForward.apply(b_, l_, nf, ctx_obj, N.LOG, -1)
  loss = dist.sum()
  ms = timeit(lambda: torch.autograd.grad(loss, [b_, l_], retain_graph=True))
  gb, gl = torch.autograd.grad(loss, [b_, l_], retain_graph=True)
  if name == 'FullNGram':
    ref_g = (gb, gl)
    print(f'{name:14s} Log backward: {ms:8.3f} ms {2 * W / ms * 1e3:7.0f} GB/s (2 W)')
  else:
    d = max((gb - ref_g[0]).abs().max().item(), (gl - ref_g[1]).abs().max().item())
    print(f'{name:14s} Log backward: {ms:8.3f} ms {2 * W / ms * 1e3:7.0f} GB/s (2 W)   '
          f'max |grad diff| {d:.1e}')
  del b_, l_, dist, loss, gb, gl

# ---- N4 ------------------------------------------------------------------
for name, fn in [('hat_normalize', lt.weight_fns.hat_normalize),
                 ('log_softmax_normalize', lt.weight_fns.log_softmax_normalize)]:
  ms = timeit(lambda: fn(blank, lex))
  print(f'{name:22s} forward : {ms:8.3f} ms {2 * W / ms * 1e3:7.0f} GB/s (read W + write W)')
  b_ = blank.clone().requires_grad_()
  l_ = lex.clone().requires_grad_()
  ob, ol = fn(b_, l_)
  gob, gol = torch.randn_like(ob), torch.randn_like(ol)
  ms = timeit(lambda: torch.autograd.grad([ob, ol], [b_, l_], [gob, gol], retain_graph=True))
  print(f'{name:22s} backward: {ms:8.3f} ms {3 * W / ms * 1e3:7.0f} GB/s (read 2 W + write W)')
  del b_, l_, ob, ol, gob, gol
