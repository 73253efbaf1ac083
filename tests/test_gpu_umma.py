"""tcgen05 building block: D = A * B^T through hand-written UMMA descriptors,
bf16 hi/lo split operands and an fp32 TMEM accumulator, against fp32 matmul."""
import ctypes

import numpy as np
import pytest
import torch

from last_torch_b200 import _native as N

pytestmark = pytest.mark.gpu


def _probe(a, b, terms):
  from last_torch_b200 import _native as N
  N.lib()
  handle = ctypes.CDLL(N.LIB_PATH)
  fn = handle.ltx_umma_probe
  fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
  fn.restype = ctypes.c_int
  n, k = b.shape
  d = torch.empty([128, n], device='cuda')
  rc = fn(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, terms,
          torch.cuda.current_stream().cuda_stream)
  assert rc == 0, handle.lt_last_error()
  torch.cuda.synchronize()
  return d


@pytest.mark.parametrize('n,k', [(256, 64), (256, 512), (64, 128), (32, 64), (128, 256)])
def test_umma_probe_matches_fp32_matmul(n, k):
  g = torch.Generator(device='cuda').manual_seed(n * 1000 + k)
  a = torch.randn([128, k], device='cuda', generator=g)
  b = torch.randn([n, k], device='cuda', generator=g)
  ref = (a.double() @ b.double().T)
  scale = float(ref.abs().max())
  d1 = _probe(a, b, 1)
  d3 = _probe(a, b, 3)
  err1 = float((d1.double() - ref).abs().max()) / scale
  err3 = float((d3.double() - ref).abs().max()) / scale
  # one bf16 term is ~2^-9 accurate, the three-term split ~2^-17
  assert err1 < 2e-2, err1
  assert err3 < 2e-5, (err1, err3)
  assert err3 < err1 / 50


@pytest.mark.parametrize('n,k', [(256, 64), (256, 512), (128, 128), (64, 192), (32, 64)])
def test_umma_probe_operand_in_tensor_memory(n, k):
  """The same product with the A operand written to TMEM by tcgen05.st (row = lane, two K
  elements per 32-bit column) and only B in shared memory."""
  N.lib()
  handle = ctypes.CDLL(N.LIB_PATH)
  fn = handle.ltx_umma_probe_ts
  fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 + [ctypes.c_void_p]
  fn.restype = ctypes.c_int
  g = torch.Generator(device='cuda').manual_seed(n * 1000 + k + 1)
  a = torch.randn([128, k], device='cuda', generator=g)
  b = torch.randn([n, k], device='cuda', generator=g)
  d = torch.empty([128, n], device='cuda')
  rc = fn(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, torch.cuda.current_stream().cuda_stream)
  assert rc == 0, handle.lt_last_error()
  torch.cuda.synchronize()
  ref = a.double() @ b.double().T
  err = float((d.double() - ref).abs().max()) / float(ref.abs().max())
  assert err < 2e-5, err


@pytest.mark.parametrize('c,v,h,n', [(257, 256, 512, 300), (65, 64, 128, 9), (130, 128, 192, 31),
                                     (7, 192, 1024, 77), (257, 256, 512, 2501), (64, 64, 64, 1),
                                     (33, 256, 128, 130)])
def test_joint_forward_operand_in_tensor_memory_equals_shared_memory_form(c, v, h, n):
  """lt_joint_forward with the tanh operand in tensor memory (default for V % 64 == 0,
  H <= 1024) against the shared-memory-operand kernel: same bf16x3 products, so the two agree to
  fp32 accumulation order."""
  from last_torch_b200.joint import joint_forward_raw
  g = torch.Generator(device='cuda').manual_seed(c * 31 + n)
  pc = torch.randn([c, h], device='cuda', generator=g)
  pf = torch.randn([n, h], device='cuda', generator=g)
  wb = torch.randn([1, h], device='cuda', generator=g) * 0.3
  bb = torch.full([], -0.5, device='cuda')
  wv = torch.randn([v, h], device='cuda', generator=g) * 0.3
  bv = torch.randn([v], device='cuda', generator=g)
  b_ts, l_ts = joint_forward_raw(pc, pf, wb, bb, wv, bv)
  with N.option('LT_JOINT_FWD_CLUSTER', 1):                # no multicast: one CTA per cluster
    b_1, l_1 = joint_forward_raw(pc, pf, wb, bb, wv, bv)
  assert torch.equal(l_1, l_ts) and torch.equal(b_1, b_ts)
  with N.option('LT_JOINT_FWD_SS', 1):
    b_ss, l_ss = joint_forward_raw(pc, pf, wb, bb, wv, bv)
  j = torch.tanh(pc.double()[None] + pf.double()[:, None])
  rl = j @ wv.double().T + bv.double()
  rb = j @ wb.double()[0] + bb.double()
  scale = float(rl.abs().max())
  assert float((l_ts.double() - rl).abs().max()) / scale < 1e-5
  assert float((l_ss.double() - rl).abs().max()) / scale < 1e-5
  assert float((l_ts - l_ss).abs().max()) / scale < 2e-6
  assert float((b_ts.double() - rb).abs().max()) / float(rb.abs().max()) < 1e-5
  assert float((b_ts - b_ss).abs().max()) / float(rb.abs().max()) < 2e-6


@pytest.mark.parametrize('c,v,h,n', [(257, 256, 512, 70), (65, 64, 128, 9), (33, 32, 64, 5),
                                     (130, 128, 192, 31)])
def test_joint_forward_tcgen05_matches_fp32(c, v, h, n):
  """lt_joint_forward on the tcgen05 path (bf16x3 split, fp32 TMEM accumulate)
  against the fp32 reference formula (weight_fns.py:208-227); 1e-5 of the
  logit scale, and identical (to 1e-5) to the CUDA-core fp32 kernels."""
  import last_torch_b200 as lt
  torch.manual_seed(c + v)
  fn = lt.weight_fns.JointWeightFn(vocab_size=v, hidden_size=h, device='cuda', embedding_size=48,
                                   feature_size=40)
  cache = torch.randn([c, 48], device='cuda')
  frames = torch.randn([n, 1, 40], device='cuda')
  with torch.no_grad():
    kb, kl = fn.all_frames(cache, frames)                  # kernel path: [n,1,c], [n,1,c,v]
    pcx = fn.context_projection(cache).double()
    pfx = fn.blank_projection(frames[:, 0]).double()
    joint = torch.tanh(pcx[None] + pfx[:, None])
    rl = joint @ fn.joint_projection_to_vocab.weight.double().T + fn.joint_projection_to_vocab.bias.double()
    rb = joint @ fn.joint_projection_to_blank.weight.double()[0] + fn.joint_projection_to_blank.bias.double()
    with N.option('LT_JOINT_SIMT', 1):
      sb, sl = fn.all_frames(cache, frames)
  scale = float(rl.abs().max())
  err_tc = float((kl[:, 0].double() - rl).abs().max()) / scale
  err_simt = float((sl[:, 0].double() - rl).abs().max()) / scale
  err_b = float((kb[:, 0].double() - rb).abs().max()) / max(float(rb.abs().max()), 1e-6)
  assert err_simt < 2e-6, err_simt
  assert err_tc < 1e-5, (err_tc, err_simt)
  assert err_b < 1e-5, err_b


@pytest.mark.parametrize('c,v,h,n', [(257, 256, 512, 40), (65, 64, 128, 9)])
def test_joint_forward_large_preactivations(c, v, h, n):
  """Pre-activations of +-20 .. 40 that partly cancel (context projection against frame
  projection): tanh(pc + pf) must come out of the exponential tables exactly, not from factors
  clamped so that their product stays normal (the tables clamp at 2^+-126 per factor and let the
  product overflow / flush inside 1 / (1 + E_c E_f); exact up to |x| = 43.6)."""
  import last_torch_b200 as lt
  torch.manual_seed(3 * c + v)
  fn = lt.weight_fns.JointWeightFn(vocab_size=v, hidden_size=h, device='cuda', embedding_size=48,
                                   feature_size=40)
  cache = torch.randn([c, 48], device='cuda')
  frames = torch.randn([n, 1, 40], device='cuda')
  with torch.no_grad():
    # scale both projections so that their entries reach +-38: a few per cent of the (state, frame,
    # hidden unit) triples then pair a factor beyond 21.8 with a sum inside the unsaturated range
    pcx = fn.context_projection(cache)
    pfx = fn.blank_projection(frames[:, 0])
    fn.context_projection.weight.mul_(38.0 / float(pcx.abs().max()))
    fn.blank_projection.weight.mul_(38.0 / float(pfx.abs().max()))
    kb, kl = fn.all_frames(cache, frames)
    pcx = fn.context_projection(cache).double()
    pfx = fn.blank_projection(frames[:, 0]).double()
    assert float(pcx.abs().max()) > 30 and float(pfx.abs().max()) > 30
    s = pcx[None] + pfx[:, None]
    # the case the clamp of round 1 got wrong must be present: one factor beyond 21.8, sum small
    assert int(((pcx[None].abs() > 22) & (s.abs() < 9)).sum()) > 0
    joint = torch.tanh(s)
    rl = joint @ fn.joint_projection_to_vocab.weight.double().T + fn.joint_projection_to_vocab.bias.double()
    rb = joint @ fn.joint_projection_to_blank.weight.double()[0] + fn.joint_projection_to_blank.bias.double()
  scale = float(rl.abs().max())
  assert float((kl[:, 0].double() - rl).abs().max()) / scale < 1e-5
  assert float((kb[:, 0].double() - rb).abs().max()) / max(float(rb.abs().max()), 1e-6) < 1e-5


@pytest.mark.parametrize('c,v,h,n', [(257, 256, 512, 40), (65, 64, 128, 9), (130, 128, 256, 17),
                                     (257, 256, 512, 300), (33, 64, 128, 261), (17, 192, 384, 130)])
def test_joint_backward_tcgen05_matches_autograd(c, v, h, n):
  """Gradients of the whole-utterance JointWeightFn kernel path (tcgen05 dgrad +
  streaming reduction, weight gradients) against torch autograd through the
  per-frame reference formula, and against the CUDA-core kernels."""
  import last_torch_b200 as lt
  torch.manual_seed(c * 7 + v)
  fn = lt.weight_fns.JointWeightFn(vocab_size=v, hidden_size=h, device='cuda', embedding_size=48,
                                   feature_size=40)
  params = list(fn.parameters())
  cache = torch.randn([c, 48], device='cuda', requires_grad=True)
  frames = torch.randn([n, 1, 40], device='cuda', requires_grad=True)
  wb = torch.randn([n, 1, c], device='cuda')
  wl = torch.randn([n, 1, c, v], device='cuda')

  def grads(path):
    with N.option('LT_JOINT_SIMT', 1 if path == 'simt' else 0):
      if path == 'ref':
        b, l = fn(cache, frames.reshape(n, 40))
        b, l = b.reshape(wb.shape), l.reshape(wl.shape)
      else:
        b, l = fn.all_frames(cache, frames)
      return torch.autograd.grad((b * wb).sum() + (l * wl).sum(), params + [cache, frames])

  ref, tc, simt = grads('ref'), grads('tc'), grads('simt')
  for r, a, s in zip(ref, tc, simt):
    scale = float(r.abs().max()) + 1e-12
    assert float((s - r).abs().max()) / scale < 3e-5
    assert float((a - r).abs().max()) / scale < 3e-5, (tuple(r.shape), float((a - r).abs().max()) / scale)


@pytest.mark.parametrize('c,v,h,n', [(65, 64, 128, 148 * 128 + 77), (130, 128, 256, 37 * 128 * 2 + 5)])
def test_joint_backward_frame_blocks_straddle_ctas(c, v, h, n):
  """More 128-frame blocks than CTAs per hidden block: the fused dgrad cuts the (frame block, c)
  tile sequence into equal ranges, so blocks straddle two CTAs and grad_proj_frame is flushed by
  both; the last block is partial (rows past N read the zero line).  Tensor-core path against the
  CUDA-core kernels (themselves checked against autograd above)."""
  import last_torch_b200 as lt
  torch.manual_seed(n)
  fn = lt.weight_fns.JointWeightFn(vocab_size=v, hidden_size=h, device='cuda', embedding_size=24,
                                   feature_size=16)
  params = list(fn.parameters())
  cache = torch.randn([c, 24], device='cuda', requires_grad=True)
  frames = torch.randn([n, 1, 16], device='cuda', requires_grad=True)
  wb = torch.randn([n, 1, c], device='cuda')
  wl = torch.randn([n, 1, c, v], device='cuda')

  def run(simt):
    with N.option('LT_JOINT_SIMT', 1 if simt else 0):
      b, l = fn.all_frames(cache, frames)
      g = torch.autograd.grad((b * wb).sum() + (l * wl).sum(), params + [cache, frames])
      return [b.detach(), l.detach()] + list(g)

  tc, simt = run(False), run(True)
  for a, s in zip(tc, simt):
    scale = float(s.abs().max()) + 1e-12
    assert float((a - s).abs().max()) / scale < 3e-5, (tuple(s.shape), float((a - s).abs().max()) / scale)


@pytest.mark.parametrize('variant', ['', 'LT_JOINT_DGRAD_PAIR', 'LT_JOINT_DGRAD_MULTICAST'])
def test_joint_backward_split_row_kernels(variant):
  """The split-row forms of the tensor-core backward kernels (fused dgrad fed by TMA, weight
  gradient copying its A operand) on a split copy of the fp32 gradient (lt_joint_split_rows),
  plus the opt-in CTA-pair and TMA-multicast variants of the dgrad: same gradients as the fp32
  kernels."""
  from last_torch_b200 import joint
  c, v, h, n = 257, 256, 512, 300
  torch.manual_seed(7)
  pc = torch.randn([c, h], device='cuda')
  pf = torch.randn([n, h], device='cuda')
  w_blank = torch.randn([1, h], device='cuda') / h ** 0.5
  w_vocab = torch.randn([v, h], device='cuda') / h ** 0.5
  gb = torch.rand([n, c], device='cuda') / c
  gl = torch.rand([n, c, v], device='cuda') / (c * v)
  assert N.lib().lt_joint_backward_split_supported(n, c, h, v) == 1
  ref = joint.joint_backward_raw(pc, pf, w_blank, w_vocab, gb, gl, fmt=0)
  split = torch.empty_like(gl)
  N.check(N.lib().lt_joint_split_rows(N.ptr(gl), N.ptr(split), n * c, v,
                                      N.stream_ptr(gl.device)), 'lt_joint_split_rows')
  rows = split.view(torch.bfloat16).reshape(n, c, 2, v).float()
  assert float((rows[:, :, 0] + rows[:, :, 1] - gl).abs().max()) <= float(gl.max()) * 2.0 ** -16
  if variant:
    with N.option(variant, 1):
      got = joint.joint_backward_raw(pc, pf, w_blank, w_vocab, gb, split, fmt=1)
  else:
    got = joint.joint_backward_raw(pc, pf, w_blank, w_vocab, gb, split, fmt=1)
  for a, r in zip(got, ref):
    scale = float(r.abs().max()) + 1e-12
    assert float((a - r).abs().max()) / scale < 2e-5, (tuple(r.shape), float((a - r).abs().max()) / scale)


def _probe_mn(at, bt, swap):
  from last_torch_b200 import _native as N
  N.lib()
  handle = ctypes.CDLL(N.LIB_PATH)
  fn = handle.ltx_umma_probe_mn
  fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
  fn.restype = ctypes.c_int
  k, n = bt.shape
  d = torch.empty([128, n], device='cuda')
  rc = fn(at.data_ptr(), bt.data_ptr(), d.data_ptr(), n, k, swap,
          torch.cuda.current_stream().cuda_stream)
  assert rc == 0
  torch.cuda.synchronize()
  return d


@pytest.mark.parametrize('n,k', [(256, 32), (256, 128), (64, 64), (128, 96)])
def test_umma_probe_mn_major(n, k):
  """MN-major operands (both contiguous along M / N): D = At^T Bt."""
  g = torch.Generator(device='cuda').manual_seed(n + k)
  at = torch.randn([k, 128], device='cuda', generator=g)
  bt = torch.randn([k, n], device='cuda', generator=g)
  ref = at.double().T @ bt.double()
  scale = float(ref.abs().max())
  err = float((_probe_mn(at, bt, 0).double() - ref).abs().max()) / scale
  assert err < 2e-5, err
