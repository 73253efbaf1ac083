"""JointWeightFn over all frames (weight_fns.py:194-227), whole-utterance form.

The two small input projections ([C,E]x[E,H] and [N,D]x[D,H]) are library GEMMs by default
(lt_linear_forward / lt_linear_wgrad with LT_OWN_LINEAR=1); the hot part -- tanh of the [N,C,H] joint and its projection to
[N,C,1+V] -- is lt_joint_forward / lt_joint_backward.
"""

from __future__ import annotations

import os

import torch

from . import _native as N


def joint_forward_raw(proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab):
  """(blank [N,C], lexical [N,C,V]) from proj_ctx [C,H], proj_frame [N,H]; b_blank is a device
  scalar read by the kernel (no host synchronisation)."""
  proj_ctx = N.require_cuda(proj_ctx, 'proj_ctx')
  proj_frame = N.require_cuda(proj_frame, 'proj_frame')
  w_blank = N.require_cuda(w_blank.reshape(-1), 'w_blank')
  b_blank = N.require_cuda(b_blank.reshape(-1), 'b_blank')
  w_vocab = N.require_cuda(w_vocab, 'w_vocab')
  b_vocab = N.require_cuda(b_vocab, 'b_vocab')
  n, h = proj_frame.shape
  c = proj_ctx.shape[0]
  v = w_vocab.shape[0]
  dev = proj_frame.device
  blank = torch.empty([n, c], dtype=torch.float32, device=dev)
  lexical = torch.empty([n, c, v], dtype=torch.float32, device=dev)
  workspace = torch.empty([int(N.lib().lt_joint_workspace_bytes(n, c, h, v))], dtype=torch.uint8,
                          device=dev)
  with torch.cuda.device(dev):
    N.check(N.lib().lt_joint_forward(
        N.ptr(proj_ctx), N.ptr(proj_frame), N.ptr(w_blank), N.ptr(b_blank), N.ptr(w_vocab),
        N.ptr(b_vocab), n, c, h, v, N.ptr(blank), N.ptr(lexical), N.ptr(workspace),
        N.stream_ptr(dev)), 'lt_joint_forward')
  return blank, lexical


def joint_backward_raw(proj_ctx, proj_frame, w_blank, w_vocab, g_blank, g_lexical, fmt=0):
  """Gradients w.r.t. (proj_ctx, proj_frame, w_blank [1,H], b_blank [], w_vocab, b_vocab).
  fmt=1: g_lexical holds split rows (include/last_lattice.h, lt_joint_backward)."""
  n, h = proj_frame.shape
  c = proj_ctx.shape[0]
  v = w_vocab.shape[0]
  dev = proj_frame.device
  w_blank = w_blank.reshape(-1)
  g_blank = N.require_cuda(g_blank, 'grad_blank')
  g_lexical = N.require_cuda(g_lexical, 'grad_lexical')
  g_pc = torch.zeros_like(proj_ctx)
  g_pf = torch.zeros_like(proj_frame)
  g_wb = torch.zeros_like(w_blank)
  g_bb = torch.zeros([1], dtype=torch.float32, device=dev)
  g_wv = torch.zeros_like(w_vocab)
  g_bv = torch.zeros([v], dtype=torch.float32, device=dev)
  workspace = torch.empty([int(N.lib().lt_joint_backward_workspace_bytes(n, c, h, v))],
                          dtype=torch.uint8, device=dev)
  with torch.cuda.device(dev):
    N.check(N.lib().lt_joint_backward(
        N.ptr(proj_ctx), N.ptr(proj_frame), N.ptr(w_blank), N.ptr(w_vocab), N.ptr(g_blank),
        N.ptr(g_lexical), n, c, h, v, N.ptr(g_pc), N.ptr(g_pf), N.ptr(g_wb), N.ptr(g_bb),
        N.ptr(g_wv), N.ptr(g_bv), N.ptr(workspace), fmt, N.stream_ptr(dev)), 'lt_joint_backward')
  return g_pc, g_pf, g_wb.reshape(1, -1), g_bb.reshape(()), g_wv, g_bv


class _JointProjection(torch.autograd.Function):
  """(blank [N,C], lexical [N,C,V]) from proj_ctx [C,H], proj_frame [N,H]."""

  @staticmethod
  def forward(ctx, proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab):
    blank, lexical = joint_forward_raw(proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab)
    ctx.save_for_backward(proj_ctx.contiguous(), proj_frame.contiguous(), w_blank, w_vocab)
    return blank, lexical

  @staticmethod
  def backward(ctx, g_blank, g_lexical):
    proj_ctx, proj_frame, w_blank, w_vocab = ctx.saved_tensors
    return joint_backward_raw(proj_ctx, proj_frame, w_blank, w_vocab.contiguous(),
                              g_blank.contiguous(), g_lexical.contiguous())


def linear_forward_raw(x, w):
  """x [M,K] . w [N,K]^T -> [M,N] (lt_linear_forward; fp32, contiguous, CUDA)."""
  m, k = x.shape
  n = w.shape[0]
  y = torch.empty([m, n], dtype=torch.float32, device=x.device)
  with torch.cuda.device(x.device):
    N.check(N.lib().lt_linear_forward(N.ptr(x), N.ptr(w), N.ptr(y), m, k, n,
                                      N.stream_ptr(x.device)), 'lt_linear_forward')
  return y


class _Linear(torch.autograd.Function):
  """nn.Linear without bias (weight_fns.py:208-211) on the library's own kernels."""

  @staticmethod
  def forward(ctx, x, w):
    x = N.require_cuda(x, 'input').contiguous().float()
    w = N.require_cuda(w, 'weight').contiguous().float()
    ctx.save_for_backward(x, w)
    return linear_forward_raw(x, w)

  @staticmethod
  def backward(ctx, gy):
    x, w = ctx.saved_tensors
    gy = gy.contiguous()
    gx = gw = None
    if ctx.needs_input_grad[0]:
      gx = linear_forward_raw(gy, w.t().contiguous())            # gy [M,N] . (w^T [K,N])^T
    if ctx.needs_input_grad[1]:
      m, k = x.shape
      n = w.shape[0]
      gw = torch.empty_like(w)
      ws = torch.empty([max(int(N.lib().lt_linear_wgrad_workspace_bytes(m, k, n)), 4)],
                       dtype=torch.uint8, device=x.device)
      with torch.cuda.device(x.device):
        N.check(N.lib().lt_linear_wgrad(N.ptr(gy), N.ptr(x), N.ptr(gw), m, k, n, N.ptr(ws),
                                        N.stream_ptr(x.device)), 'lt_linear_wgrad')
    return gx, gw


# The input projections are plain skinny fp32 GEMMs (2.6 GFLOP at the headline shape, 3 % of the
# step).  Measured on a B200: lt_linear_forward 0.113 ms / lt_linear_wgrad 0.170 ms against
# 0.074 / 0.136 ms for the library sgemm behind nn.Linear -- so the library call stays the default
# and the library's own kernels are the opt-in path (LT_OWN_LINEAR=1) for a build without cuBLAS.
OWN_LINEAR = os.environ.get('LT_OWN_LINEAR', '0') == '1'
# LT_NO_TC_LINEAR=1: nn.Linear also for projections the tcgen05 kernels would take
# (lt_linear_tensor_core: M >= 4096, K % 64 == 0, N % 128 == 0; 0.11 + 0.15 ms against the library
# sgemm's 0.33 + 0.38 ms at 32000 x 512 x 512)
NO_TC_LINEAR = os.environ.get('LT_NO_TC_LINEAR', '0') == '1'


_TC_LINEAR_SHAPES = {}


def _tensor_core_linear(m, k, n):
  """lt_linear_tensor_core(m, k, n), asked once per shape (a host-side query, no kernel)."""
  key = (int(m), int(k), int(n))
  if key not in _TC_LINEAR_SHAPES:
    _TC_LINEAR_SHAPES[key] = bool(N.lib().lt_linear_tensor_core(*key))
  return _TC_LINEAR_SHAPES[key]


def joint_projections(fn, cache, frames):
  """The two bias-free input projections in front of the joint kernel: proj_ctx [C,H],
  proj_frame [N,H] (weight_fns.py:208-211)."""
  flat = frames.reshape(-1, frames.shape[-1])
  if OWN_LINEAR:
    return (_Linear.apply(cache, fn.context_projection.weight),
            _Linear.apply(flat, fn.blank_projection.weight))

  # A production-sized frame projection ([B T, D] x [H, D], 16.8 GFLOP at the headline shape) runs
  # on the library's tcgen05 kernels, and the small context projection then goes through the
  # library's kernels as well (0.03 ms against the sgemm's 0.015 ms: no library GEMM is left on the
  # path).  Where the frame projection does not qualify, both stay with nn.Linear.
  wf = fn.blank_projection.weight
  if (not NO_TC_LINEAR and flat.is_cuda and flat.dtype == torch.float32 and
      wf.dtype == torch.float32 and cache.dtype == torch.float32 and
      _tensor_core_linear(flat.shape[0], flat.shape[1], wf.shape[0]) and
      flat.data_ptr() % 16 == 0):
    return (_Linear.apply(cache, fn.context_projection.weight), _Linear.apply(flat, wf))
  return fn.context_projection(cache), fn.blank_projection(flat)


def joint_all_frames(fn, cache, frames):
  """fn: weight_fns.JointWeightFn; cache [C,E]; frames [batch..., T, D]."""
  batch_shape = frames.shape[:-1]
  proj_ctx, proj_frame = joint_projections(fn, cache, frames)
  blank, lexical = _JointProjection.apply(
      proj_ctx, proj_frame, fn.joint_projection_to_blank.weight,
      fn.joint_projection_to_blank.bias.reshape(()), fn.joint_projection_to_vocab.weight,
      fn.joint_projection_to_vocab.bias)
  c, v = proj_ctx.shape[0], fn.vocab_size
  return blank.reshape(*batch_shape, c), lexical.reshape(*batch_shape, c, v)


def joint_string_frames(fn, cache, frames, states):
  """fn: weight_fns.JointWeightFn; cache [C,E]; frames [B,T,D]; states [B,U1] ->
  (blank [B,T,U1], lexical [B,T,U1,V]): the vocabulary projection on the context states of each
  utterance's label string only (lattices.py:300-313)."""
  b, t, _ = frames.shape
  proj_ctx, proj_frame = joint_projections(fn, cache, frames)              # [C,H], [B*T,H]
  proj_frame = proj_frame.reshape(b, t, -1)
  idx = states.to(torch.int64)
  blanks, lexicals = [], []
  for i in range(b):
    bl, lx = _JointProjection.apply(
        proj_ctx.index_select(0, idx[i]), proj_frame[i], fn.joint_projection_to_blank.weight,
        fn.joint_projection_to_blank.bias.reshape(()), fn.joint_projection_to_vocab.weight,
        fn.joint_projection_to_vocab.bias)
    blanks.append(bl)
    lexicals.append(lx)
  return torch.stack(blanks), torch.stack(lexicals)
