"""One test-scale launch of every kernel family that exchanges state through shared memory /
DSMEM (st.async + mbarrier complete_tx, TMA rings, TMEM) -- the target of tools/sanitize.sh:

    compute-sanitizer --tool {memcheck,racecheck,synccheck} python tools/sanitize_targets.py

compute-sanitizer is CLOSED on this GPU pool (profiles/r02_sanitizer_unavailable.log), so the
script carries its own checks: guard bands around every output buffer of the lattice kernels
(out-of-bounds writes) and bit-identical results over repeated runs (races), next to the
comparison of every fast kernel family with the generic kernels.

Families: lattice_fast2 (bigram TMA fast path, plain and renormalised, fp32 and split-row
gradients, expectation variant), lattice_fast2_fld (FrameLabelDependent on the same path),
joint_fwd_ts (tensor-memory operand, multicast pairs), lattice_cols / lattice_rows (context_size 2), the generic cluster kernels,
lattice_table2 (NextStateTable clusters), string_lattice (numerator, plain and (e, f) chain),
viterbi back-trace, joint_tc forward / wgrad and joint_dgrad2 (tcgen05).  Results are checked
against the generic kernels so that a sanitizer-induced slowdown cannot hide a wrong answer."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import last_torch_b200 as lt  # noqa: E402
from last_torch_b200 import _native as N  # noqa: E402


def table_lattice(context, k, table, flags=0):
  alignment = lt.alignments.FrameDependent() if k < 0 else lt.alignments.FrameLabelDependent(k)
  lattice = lt.RecognitionLattice(
      context=context, alignment=alignment,
      weight_fn_factory=lambda _: lt.weight_fns.TableWeightFn(table),
      weight_fn_cacher_factory=lambda _: lt.weight_fns.NullCacher())
  lattice.kernel_flags = flags
  return lattice


def loss_and_grad(context, k, table_np, nf, labels, nl, flags=0):
  table = torch.tensor(table_np, device='cuda', requires_grad=True)
  b, t = table.shape[:2]
  frames = torch.arange(t, device='cuda', dtype=torch.float32)[None, :, None].expand(b, t, 1)
  lattice = table_lattice(context, k, table, flags)
  args = dict(frames=frames, num_frames=torch.tensor(nf, device='cuda'), cache=None)
  loss = lattice(labels=torch.tensor(labels, device='cuda'),
                 num_labels=torch.tensor(nl, device='cuda'), **args)
  (g,) = torch.autograd.grad(loss.sum(), table)
  path = lattice.shortest_path(**args)
  torch.cuda.synchronize()
  return loss.detach().cpu().numpy(), g.cpu().numpy(), [p.cpu().numpy() for p in path]


def lattice_family(name, vocab, n, k, b, t, u):
  rng = np.random.RandomState(vocab + n)
  context = lt.contexts.FullNGram(vocab_size=vocab, context_size=n)
  c = context.num_states()
  table = rng.randn(b, t, c, 1 + vocab).astype(np.float32)
  nf = [t] + [max(1, t - 3 * i) for i in range(1, b)]
  labels = rng.randint(1, vocab + 1, size=(b, u))
  nl = [u] + [max(0, u - i) for i in range(1, b)]
  fast = loss_and_grad(context, k, table, nf, labels, nl, 0)
  ref = loss_and_grad(context, k, table, nf, labels, nl, 1)          # LT_FLAG_FORCE_GENERIC
  np.testing.assert_allclose(fast[0], ref[0], rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(fast[1], ref[1], rtol=1e-4, atol=1e-5)
  np.testing.assert_array_equal(fast[2][0], ref[2][0])
  print('ok', name, flush=True)


def table_family():
  rng = np.random.RandomState(3)
  c, vocab, b, t, u = 130, 64, 3, 9, 4
  nst = rng.randint(0, c, size=(c, vocab)).astype(np.int32)
  context = lt.contexts.NextStateTable(torch.from_numpy(nst))
  table = rng.randn(b, t, c, 1 + vocab).astype(np.float32)
  nf, nl = [9, 5, 0], [4, 2, 0]
  labels = rng.randint(1, vocab + 1, size=(b, u))
  with N.option('LT_TABLE_CLUSTER', 4):
    fast = loss_and_grad(context, -1, table, nf, labels, nl)
  with N.option('LT_TABLE_V1', 1):
    ref = loss_and_grad(context, -1, table, nf, labels, nl)
  np.testing.assert_allclose(fast[0], ref[0], rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(fast[1], ref[1], rtol=1e-4, atol=1e-5)
  print('ok lattice_table2 (cluster of 4)', flush=True)


def joint_family(vocab, hidden, b, t):
  torch.manual_seed(vocab)
  out = {}
  for split in (True, False):
    torch.manual_seed(vocab)
    lattice = lt.RecognitionLattice(
        context=lt.contexts.FullNGram(vocab_size=vocab, context_size=1),
        alignment=lt.alignments.FrameDependent(),
        weight_fn_cacher_factory=lambda c: lt.weight_fns.SharedEmbCacher(
            num_context_states=c.shape()[0], embedding_size=24, device='cuda'),
        weight_fn_factory=lambda c: lt.weight_fns.JointWeightFn(
            vocab_size=c.shape()[1], hidden_size=hidden, device='cuda', embedding_size=24,
            feature_size=16))
    lattice.split_grad_handover = split
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn([b, t, 16], device='cuda', generator=g)
    loss = lattice(frames=x, num_frames=torch.tensor([t, t - 5][:b], device='cuda'),
                   labels=torch.randint(1, vocab + 1, [b, 5], device='cuda', generator=g),
                   num_labels=torch.tensor([5, 3][:b], device='cuda'))
    loss.sum().backward()
    torch.cuda.synchronize()
    out[split] = [p.grad.clone() for p in lattice.parameters()]
  for a, r in zip(out[True], out[False]):
    assert float((a - r).abs().max()) <= 2e-5 * (float(r.abs().max()) + 1e-12)
  print(f'ok joint_tc / joint_dgrad2 / wgrad (vocab {vocab}, hidden {hidden})', flush=True)


def guarded(shape, dtype=torch.float32, pad=4096):
  """A tensor of `shape` inside a larger allocation whose margins hold a sentinel: a kernel
  that writes outside its output buffer (what memcheck would flag) destroys the sentinel."""
  n = 1
  for d in shape:
    n *= d
  sentinel = 12345 if dtype in (torch.int32, torch.int16, torch.uint8) else 12345.0
  if dtype == torch.uint8:
    sentinel = 123
  buf = torch.full([n + 2 * pad], sentinel, dtype=dtype, device='cuda')
  view = buf[pad:pad + n].view(shape)

  def check():
    torch.cuda.synchronize()
    assert bool((buf[:pad] == sentinel).all()) and bool((buf[pad + n:] == sentinel).all()), \
        f'write outside a {tuple(shape)} {dtype} output buffer'
  return view, check


def raw_lattice_pair(name, vocab, n, k, b, t, flags=0, repeats=6):
  """lt_lattice_forward_norm / lt_lattice_backward_norm called directly on guarded output
  buffers, `repeats` times on the same inputs: (1) no write lands outside a buffer; (2) every
  run gives BIT-IDENTICAL dist / alphas / gradients -- the kernels reduce in a fixed order, so a
  missed mbarrier wait or a DSMEM store that races with its reader shows up as run-to-run
  differences (the check compute-sanitizer's racecheck would make; it is closed on this pool)."""
  rng = np.random.RandomState(7 * vocab + n)
  c = sum(vocab**i for i in range(n + 1))
  blank = torch.tensor(rng.randn(b, t, c).astype(np.float32), device='cuda')
  lex = torch.tensor(rng.randn(b, t, c, vocab).astype(np.float32), device='cuda')
  nf = torch.tensor([t] + [max(0, t - 2 * i - 1) for i in range(1, b)], dtype=torch.int32,
                    device='cuda')
  gd = torch.tensor(rng.rand(b).astype(np.float32) + 0.5, device='cuda')
  L = N.lib()
  norm = bool(L.lt_lattice_norm_supported(N.LOG, vocab, n, k, flags))
  outs = []
  for _ in range(repeats):
    dist, c1 = guarded([b])
    alphas, c2 = guarded([b, t, c])
    afin, c3 = guarded([b, c])
    levels, c4 = guarded([b, t, max(k, 1), c])
    an, c5 = guarded([b, t + 3], torch.int32)
    gb, c6 = guarded([b, t, c])
    gl, c7 = guarded([b, t, c, vocab])
    stream = N.stream_ptr(blank.device)
    N.check(L.lt_lattice_forward_norm(
        N.LOG, vocab, n, k, N.ptr(blank), N.ptr(lex), N.ptr(nf), b, t, None, N.ptr(dist),
        N.ptr(alphas), N.ptr(afin), N.ptr(levels) if k >= 1 else None, None, None,
        N.ptr(an) if norm else None, flags, stream), 'forward')
    N.check(L.lt_lattice_backward_norm(
        N.LOG, vocab, n, k, N.ptr(blank), N.ptr(lex), N.ptr(nf), b, t, N.ptr(alphas),
        N.ptr(levels) if k >= 1 else None, N.ptr(dist), N.ptr(gd), N.ptr(gb), N.ptr(gl), None,
        N.ptr(an) if norm else None, flags, stream), 'backward')
    for chk in (c1, c2, c3, c4, c5, c6, c7):
      chk()
    outs.append([x.clone() for x in (dist, alphas, gb, gl)])
  for o in outs[1:]:
    for a, r in zip(o, outs[0]):
      assert torch.equal(a, r), f'{name}: run-to-run difference'
  print(f'ok {name}: {repeats} bit-identical runs, guard bands intact (renormalised: {norm})',
        flush=True)


def raw_joint_forward(name, c, vocab, hidden, n, repeats=6):
  """lt_joint_forward (tanh operand in tensor memory, W_vocab multicast in CTA pairs) on guarded
  outputs: bit-identical runs (TMEM stages, two TMA rings, multicast commits -- a missed wait
  changes a tile), equal to the shared-memory-operand kernel."""
  from last_torch_b200.joint import joint_forward_raw
  g = torch.Generator(device='cuda').manual_seed(c + n)
  pc = torch.randn([c, hidden], device='cuda', generator=g)
  pf = torch.randn([n, hidden], device='cuda', generator=g)
  wb = torch.randn([1, hidden], device='cuda', generator=g) * 0.3
  bb = torch.full([], 0.5, device='cuda')
  wv = torch.randn([vocab, hidden], device='cuda', generator=g) * 0.3
  bv = torch.randn([vocab], device='cuda', generator=g)
  L = N.lib()
  outs = []
  for _ in range(repeats):
    blank, c1 = guarded([n, c])
    lex, c2 = guarded([n, c, vocab])
    ws = torch.empty([int(L.lt_joint_workspace_bytes(n, c, hidden, vocab))], dtype=torch.uint8,
                     device='cuda')
    N.check(L.lt_joint_forward(N.ptr(pc), N.ptr(pf), N.ptr(wb.reshape(-1)), N.ptr(bb.reshape(-1)),
                               N.ptr(wv), N.ptr(bv), n, c, hidden, vocab, N.ptr(blank), N.ptr(lex),
                               N.ptr(ws), N.stream_ptr(pc.device)), 'lt_joint_forward')
    c1(); c2()
    outs.append((blank.clone(), lex.clone()))
  for o in outs[1:]:
    assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]), f'{name}: run-to-run'
  with N.option('LT_JOINT_FWD_SS', 1):
    sb, sl = joint_forward_raw(pc, pf, wb, bb, wv, bv)
  assert torch.equal(sl, outs[0][1]), f'{name}: differs from the shared-memory-operand kernel'
  assert float((sb - outs[0][0]).abs().max()) < 1e-5
  print(f'ok {name}: {repeats} bit-identical runs, guard bands intact', flush=True)


def raw_linear(name, m, k, n, repeats=6):
  """lt_linear_forward / lt_linear_wgrad on the tcgen05 path (operand tiles staged and split in
  shared memory, one mbarrier per CTA, split-K partials reduced in a fixed order): guarded outputs,
  bit-identical runs, equal to the CUDA-core kernels to the operand split's 1e-5."""
  g = torch.Generator(device='cuda').manual_seed(m + k + n)
  x = torch.randn([m, k], device='cuda', generator=g)
  w = torch.randn([n, k], device='cuda', generator=g) / k ** 0.5
  gy = torch.randn([m, n], device='cuda', generator=g)
  L = N.lib()
  ws = torch.empty([int(L.lt_linear_wgrad_workspace_bytes(m, k, n))], dtype=torch.uint8,
                   device='cuda')

  def once():
    y, c1 = guarded([m, n])
    gw, c2 = guarded([n, k])
    N.check(L.lt_linear_forward(N.ptr(x), N.ptr(w), N.ptr(y), m, k, n, N.stream_ptr(x.device)),
            'lt_linear_forward')
    N.check(L.lt_linear_wgrad(N.ptr(gy), N.ptr(x), N.ptr(gw), m, k, n, N.ptr(ws),
                              N.stream_ptr(x.device)), 'lt_linear_wgrad')
    c1(); c2()
    return y.clone(), gw.clone()

  outs = [once() for _ in range(repeats)]
  for o in outs[1:]:
    assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]), f'{name}: run-to-run'
  with N.option('LT_LINEAR_SIMT', 1):
    y0, gw0 = once()
  for a, b in zip(outs[0], (y0, gw0)):
    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), name
  print(f'ok {name}: {repeats} bit-identical runs, guard bands intact', flush=True)


def expectation_family(vocab, b, t, repeats=4):
  """lt_lattice_expectation (K2 with posterior x value summed on the fly): bit-identical runs and
  equal to posteriors x values through the generic kernels."""
  from last_torch_b200 import ops
  rng = np.random.RandomState(vocab)
  blank = torch.tensor(rng.randn(b, t, vocab + 1).astype(np.float32), device='cuda')
  lex = torch.tensor(rng.randn(b, t, vocab + 1, vocab).astype(np.float32), device='cuda')
  nf = torch.tensor([t] + [max(0, t - 3 * i) for i in range(1, b)], dtype=torch.int32,
                    device='cuda')
  outs = [ops.lattice_expectation(blank, lex, nf, vocab, 1, -1) for _ in range(repeats)]
  for z, e in outs[1:]:
    assert torch.equal(z, outs[0][0]) and torch.equal(e, outs[0][1])
  z2, e2 = ops.lattice_expectation(blank, lex, nf, vocab, 1, -1, flags=1)
  assert float((outs[0][1] - e2).abs().max()) <= 2e-5 * float(e2.abs().max())
  print(f'ok lattice expectation vocab {vocab}: {repeats} bit-identical runs', flush=True)


def main():
  raw_linear('lt_linear tcgen05 5001 x 192 -> 256 (ragged last row tile, 26 splits)', 5001, 192, 256)
  raw_linear('lt_linear tcgen05 32000 x 512 -> 512', 32000, 512, 512, repeats=3)
  raw_lattice_pair('lattice_fast2 vocab 256, cluster of 8, 33 utterances', 256, 1, -1, 33, 40)
  raw_lattice_pair('lattice_fast2_fld FrameLabelDependent(2) vocab 256, 33 utterances', 256, 1, 2,
                   33, 24)
  raw_lattice_pair('lattice_fast2_fld FrameLabelDependent(3) vocab 192', 192, 1, 3, 5, 15)
  raw_lattice_pair('lattice_fast2_fld FrameLabelDependent(1) vocab 64, single CTA', 64, 1, 1, 4, 19)
  raw_lattice_pair('lattice_fast2 vocab 64, single CTA', 64, 1, -1, 5, 23)
  raw_lattice_pair('lattice_fast2 vocab 192, cluster of 6', 192, 1, -1, 7, 17)
  raw_lattice_pair('lattice_cols + lattice_rows vocab 64 context 2 (cluster of 8)', 64, 2, -1, 3, 9)
  raw_lattice_pair('lattice_cols FrameLabelDependent(2) vocab 32 context 2', 32, 2, 2, 3, 7)
  raw_lattice_pair('generic kernels, cluster of 4, vocab 40', 40, 1, -1, 5, 13, flags=4 << 8)
  raw_lattice_pair('generic kernels FrameLabelDependent(3), cluster of 2', 12, 1, 3, 4, 11,
                   flags=2 << 8)
  lattice_family('lattice_fast2 vocab 64 (single CTA)', 64, 1, -1, 3, 11, 4)
  lattice_family('lattice_fast2 vocab 256 (cluster of 8)', 256, 1, -1, 2, 9, 4)
  lattice_family('lattice_cols + lattice_rows vocab 32 context 2', 32, 2, -1, 2, 7, 4)
  lattice_family('lattice_cols FrameLabelDependent(2) vocab 16 context 2', 16, 2, 2, 2, 6, 5)
  lattice_family('generic cluster kernels vocab 33', 33, 1, -1, 3, 9, 4)
  table_family()
  lattice_family('lattice_fast2_fld FrameLabelDependent(2) vocab 128', 128, 1, 2, 3, 9, 4)
  expectation_family(256, 5, 21)
  expectation_family(64, 3, 9)
  raw_joint_forward('joint_fwd_ts vocab 256 hidden 512, 257 states, 1300 frames (multicast pairs)',
                    257, 256, 512, 1300)
  raw_joint_forward('joint_fwd_ts vocab 64 hidden 128, 70 states (left-over tiles)', 70, 64, 128, 333)
  joint_family(128, 128, 2, 12)
  joint_family(256, 256, 2, 9)
  print('all sanitizer targets ran', flush=True)


if __name__ == '__main__':
  main()
