"""torch.autograd.Function wrappers over the C ABI (include/last_lattice.h).

torch is used for device memory, streams and the autograd graph only; every
arithmetic step below is a hand-written sm_100a kernel.
"""

from __future__ import annotations

import math
import os

import torch

from . import _native as N

_SR_ID = {'Real': N.REAL, 'Log': N.LOG, 'MaxTropical': N.MAXTROPICAL}

# Renormalised recursion state for the Log semiring (include/last_lattice.h: lt_*_norm): on by
# default wherever a kernel supports it; LT_NO_NORM=1 (read once, at import) selects the plain
# fp32 recursion everywhere -- that is the "plain" column of profiles/r02_parity_errors.json.
USE_NORM = not os.environ.get('LT_NO_NORM')


def _as_i32(x: torch.Tensor, device) -> torch.Tensor:
  """Index tensors may arrive as float (tests/lattices_test.py:49-51)."""
  return x.to(device=device, dtype=torch.int32).contiguous()


# ---------------------------------------------------------------------------
# semiring (+) on arbitrary tensors
# ---------------------------------------------------------------------------

class SemiringPlus(torch.autograd.Function):
  """Log / MaxTropical `plus` (semirings.py:202-204, :330-332) with the
  gradients of semirings.py:264-269 (safe) and :360-369 (a >= b)."""

  @staticmethod
  def forward(ctx, a, b, sr):
    a, b = torch.broadcast_tensors(a, b)
    a = N.require_cuda(a, 'a')
    b = N.require_cuda(b, 'b')
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
      N.check(N.lib().lt_semiring_plus_forward(sr, N.ptr(a), N.ptr(b), N.ptr(out), a.numel(),
                                               N.stream_ptr(a.device)), 'semiring plus')
    ctx.save_for_backward(a, b)
    ctx.sr = sr
    return out

  @staticmethod
  def backward(ctx, g):
    a, b = ctx.saved_tensors
    g = N.require_cuda(g, 'grad')
    ga, gb = torch.empty_like(a), torch.empty_like(b)
    with torch.cuda.device(a.device):
      N.check(N.lib().lt_semiring_plus_backward(ctx.sr, N.ptr(a), N.ptr(b), N.ptr(g), N.ptr(ga),
                                                N.ptr(gb), a.numel(), N.stream_ptr(a.device)),
              'semiring plus backward')
    return ga, gb, None


class SemiringSum(torch.autograd.Function):
  """Log / MaxTropical `sum` along one axis (semirings.py:211-220, :339-348)
  with the gradients of :296-300 (safe) and :389-398 (first arg-max)."""

  @staticmethod
  def forward(ctx, a, dim, sr):
    a = N.require_cuda(a, 'a')
    dim = dim % a.ndim
    outer = 1
    for s in a.shape[:dim]:
      outer *= s
    inner = 1
    for s in a.shape[dim + 1:]:
      inner *= s
    red = a.shape[dim]
    out_shape = a.shape[:dim] + a.shape[dim + 1:]
    out = torch.empty(out_shape, dtype=a.dtype, device=a.device)
    argmax = (torch.empty(out_shape, dtype=torch.int32, device=a.device)
              if sr == N.MAXTROPICAL else None)
    with torch.cuda.device(a.device):
      N.check(N.lib().lt_semiring_sum_forward(sr, N.ptr(a), outer, red, inner, N.ptr(out),
                                              N.ptr(argmax), N.stream_ptr(a.device)),
              'semiring sum')
    ctx.save_for_backward(a, out, argmax)
    ctx.geom = (sr, outer, red, inner)
    return out

  @staticmethod
  def backward(ctx, g):
    a, out, argmax = ctx.saved_tensors
    sr, outer, red, inner = ctx.geom
    g = N.require_cuda(g, 'grad')
    ga = torch.empty_like(a)
    with torch.cuda.device(a.device):
      N.check(N.lib().lt_semiring_sum_backward(sr, N.ptr(a), N.ptr(out), N.ptr(argmax), N.ptr(g),
                                               outer, red, inner, N.ptr(ga),
                                               N.stream_ptr(a.device)), 'semiring sum backward')
    return ga, None, None


# ---------------------------------------------------------------------------
# K1 / K2 / K5: recognition-lattice forward with gradients
# ---------------------------------------------------------------------------

def _lattice_forward_raw(sr, V, n, k, blank, lexical, num_frames, flags, want_levels,
                         want_backptr, alpha_init=None, stream=None, norm=False):
  """Buffers are allocated on the current stream; the kernel is enqueued on
  `stream` when one is given (the caller orders it against the current stream).

  norm=True asks for the renormalised recursion state where a kernel supports it
  (lt_lattice_forward_norm): the 7th result is then the alpha_norm buffer [B, T+3] and `alphas`
  holds alpha~ (only lt_lattice_backward_norm / lt_alphas_denormalize understand the pair);
  otherwise the 7th result is None."""
  B, T, C = blank.shape
  dev = blank.device
  alpha_norm = None
  if (norm and USE_NORM and T > 0 and B > 0 and
      N.lib().lt_lattice_norm_supported(sr, V, n, k, flags)):
    alpha_norm = torch.empty([B, T + 3], dtype=torch.int32, device=dev)
  dist = torch.empty([B], dtype=torch.float32, device=dev)
  alphas = torch.empty([B, T, C], dtype=torch.float32, device=dev)
  alpha_final = torch.empty([B, C], dtype=torch.float32, device=dev)
  fld = k >= 1
  levels = (torch.empty([B, T, k, C], dtype=torch.float32, device=dev)
            if (fld and want_levels) else None)
  backptr = termptr = None
  if want_backptr and sr == N.MAXTROPICAL:
    backptr = torch.empty([B, T, max(k, 1), C], dtype=torch.int16, device=dev)
    if fld:
      termptr = torch.empty([B, T, C], dtype=torch.uint8, device=dev)
  with torch.cuda.device(dev), torch.cuda.stream(stream):       # stream(None) is a no-op
    N.check(N.lib().lt_lattice_forward_norm(
        sr, V, n, k, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T, N.ptr(alpha_init),
        N.ptr(dist), N.ptr(alphas), N.ptr(alpha_final), N.ptr(levels), N.ptr(backptr),
        N.ptr(termptr), N.ptr(alpha_norm), flags, N.stream_ptr(dev)), 'lt_lattice_forward')
  return dist, alphas, alpha_final, levels, backptr, termptr, alpha_norm


def _denormalized(alphas, alpha_norm):
  """True alphas (lattices.py:496) from the renormalised pair; a copy, the pair stays intact."""
  if alpha_norm is None:
    return alphas
  out = alphas.clone()
  B, T, C = out.shape
  with torch.cuda.device(out.device):
    N.check(N.lib().lt_alphas_denormalize(N.ptr(out), N.ptr(alpha_norm), B, T, C,
                                          N.stream_ptr(out.device)), 'lt_alphas_denormalize')
  return out


def lattice_expectation(blank, lexical, num_frames, V, n, k, value_blank=None, value_lexical=None,
                        flags=0):
  """(log_z [B], expect [B]), both float64: expect[b] = sum over arcs of posterior_b(arc) * value(arc)
  under the Log-semiring path distribution of the lattice -- the first-order expectation semiring
  (semirings.py:404-484) evaluated by forward-backward.  Without values the arc's own weight is
  used: entropy = log_z - expect.  No gradients flow (inputs are detached).

  On the bigram TMA fast path this is K1 + ONE more pass over the weights with no [B,T,C,V]
  posterior tensor (lt_lattice_expectation); other lattices compose lt_lattice_backward's
  posteriors with the values."""
  blank, lexical = blank.detach(), lexical.detach()
  B, T, C = blank.shape
  dev = blank.device
  if (value_blank is None) != (value_lexical is None):
    raise ValueError('value_blank and value_lexical go together')
  if value_blank is not None:
    value_blank = N.require_cuda(value_blank.detach().float().contiguous(), 'value_blank')
    value_lexical = N.require_cuda(value_lexical.detach().float().contiguous(), 'value_lexical')
    if value_blank.shape != blank.shape or value_lexical.shape != lexical.shape:
      raise ValueError(f'values must have the shapes of the weights {tuple(blank.shape)} / '
                       f'{tuple(lexical.shape)}; got {tuple(value_blank.shape)} / '
                       f'{tuple(value_lexical.shape)}')
  dist, alphas, _, levels, _, _, alpha_norm = _lattice_forward_raw(
      N.LOG, V, n, k, blank, lexical, num_frames, flags, True, False, norm=True)
  if B == 0 or T == 0:
    return dist.double(), torch.zeros([B], dtype=torch.float64, device=dev)
  log_z = dist.double()
  if alpha_norm is not None:
    # logZ = offset + residual: exact in float64 (dist alone is rounded to fp32 at |logZ| ~ 1e3)
    off = alpha_norm[:, T].double()
    res = alpha_norm[:, T + 1].contiguous().view(torch.float32).double()
    unit = torch.where(alpha_norm[:, T + 2] == 0, math.log(2.0), 1.0)
    log_z = torch.where(torch.isfinite(dist), (off + res) * unit, log_z)
  if N.lib().lt_lattice_expectation_supported(V, n, k, flags) and lexical.data_ptr() % 16 == 0:
    part = torch.empty([B, max(V // 32, 1)], dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
      N.check(N.lib().lt_lattice_expectation(
          V, n, k, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T, N.ptr(alphas),
          N.ptr(dist), N.ptr(alpha_norm), N.ptr(value_blank), N.ptr(value_lexical), N.ptr(part),
          flags, N.stream_ptr(dev)), 'lt_lattice_expectation')
    return log_z, part.sum(-1)
  gb = torch.empty_like(blank)
  gl = torch.empty_like(lexical)
  with torch.cuda.device(dev):
    N.check(N.lib().lt_lattice_backward_norm(
        N.LOG, V, n, k, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T, N.ptr(alphas),
        N.ptr(levels), N.ptr(dist), None, N.ptr(gb), N.ptr(gl), None, N.ptr(alpha_norm), flags,
        N.stream_ptr(dev)), 'lt_lattice_backward')
  vb = blank if value_blank is None else value_blank
  vl = lexical if value_lexical is None else value_lexical
  zero = torch.zeros([], device=dev)
  expect = (torch.where(gb > 0, gb * vb, zero).double().sum((1, 2)) +
            torch.where(gl > 0, gl * vl, zero).double().sum((1, 2, 3)))
  return log_z, expect


def _check_weights(blank, lexical, V, C):
  blank = N.require_cuda(blank, 'blank')
  lexical = N.require_cuda(lexical, 'lexical')
  if blank.ndim != 3 or lexical.ndim != 4:
    raise ValueError(f'blank must be [B,T,C] and lexical [B,T,C,V]; got {tuple(blank.shape)} '
                     f'and {tuple(lexical.shape)}')
  if blank.shape[-1] != C or tuple(lexical.shape) != (*blank.shape, V):
    raise ValueError(f'weights do not match the context shape ({C}, {V}): blank '
                     f'{tuple(blank.shape)}, lexical {tuple(lexical.shape)}')
  return blank, lexical


class LatticeForward(torch.autograd.Function):
  """(dist, alphas) = RecognitionLattice._forward on dense weights
  (lattices.py:379-496).  Gradients flow through `dist` only: Log / Real use
  the beta-recursion kernel, MaxTropical the Viterbi back-trace."""

  @staticmethod
  def forward(ctx, blank, lexical, num_frames, sr, V, n, k, flags):
    C = blank.shape[-1]
    blank, lexical = _check_weights(blank, lexical, V, C)
    need_grad = any(ctx.needs_input_grad[:2])
    dist, alphas, alpha_final, levels, backptr, termptr, alpha_norm = _lattice_forward_raw(
        sr, V, n, k, blank, lexical, num_frames, flags, want_levels=need_grad,
        want_backptr=need_grad, norm=True)
    ctx.geom = (sr, V, n, k, flags)
    ctx.save_for_backward(blank, lexical, num_frames, dist, alphas, alpha_final, levels, backptr,
                          termptr, alpha_norm)
    out_alphas = _denormalized(alphas, alpha_norm)
    ctx.mark_non_differentiable(out_alphas)
    return dist, out_alphas

  @staticmethod
  def backward(ctx, g_dist, _g_alphas):
    sr, V, n, k, flags = ctx.geom
    (blank, lexical, num_frames, dist, alphas, alpha_final, levels, backptr, termptr,
     alpha_norm) = ctx.saved_tensors
    B, T, C = blank.shape
    dev = blank.device
    g_dist = N.require_cuda(g_dist, 'grad_dist')
    with torch.cuda.device(dev):
      if sr == N.MAXTROPICAL:
        gb = torch.zeros_like(blank)
        gl = torch.zeros_like(lexical)
        labels = torch.empty([B, T, max(k, 0) + 1], dtype=torch.int32, device=dev)
        N.check(N.lib().lt_viterbi_backtrace(
            V, n, k, N.ptr(backptr), N.ptr(termptr), N.ptr(alpha_final), N.ptr(num_frames), B, T,
            N.ptr(labels), None, N.ptr(g_dist), N.ptr(gb), N.ptr(gl), N.stream_ptr(dev)),
            'lt_viterbi_backtrace')
      else:
        gb = torch.empty_like(blank)
        gl = torch.empty_like(lexical)
        N.check(N.lib().lt_lattice_backward_norm(
            sr, V, n, k, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T, N.ptr(alphas),
            N.ptr(levels), N.ptr(dist), N.ptr(g_dist), N.ptr(gb), N.ptr(gl), None,
            N.ptr(alpha_norm), flags, N.stream_ptr(dev)), 'lt_lattice_backward')
    return gb, gl, None, None, None, None, None, None


class LatticeForwardLevels(torch.autograd.Function):
  """LatticeForward for FrameLabelDependent(k) with ONE SET OF WEIGHTS PER ALIGNMENT STATE
  (lattices.py:447-453: per-state masks added to the weights): blank [B,T,k+1,C], lexical
  [B,T,k+1,C,V] (LT_FLAG_LEVEL_WEIGHTS).  Gradients have the same layout: Log / Real from the
  beta recursion, MaxTropical from the back-trace (one arc per taken expansion plus the closing
  blank arc of every frame)."""

  @staticmethod
  def forward(ctx, blank, lexical, num_frames, sr, V, n, k, flags):
    blank = N.require_cuda(blank, 'blank')
    lexical = N.require_cuda(lexical, 'lexical')
    B, T, L, C = blank.shape
    if L != k + 1 or tuple(lexical.shape) != (B, T, L, C, V):
      raise ValueError(f'per-state weights must be [B,T,{k + 1},C] / [B,T,{k + 1},C,V]; got '
                       f'{tuple(blank.shape)} and {tuple(lexical.shape)}')
    dev = blank.device
    flags = flags | N.FLAG_LEVEL_WEIGHTS
    dist = torch.empty([B], dtype=torch.float32, device=dev)
    alphas = torch.empty([B, T, C], dtype=torch.float32, device=dev)
    alpha_final = torch.empty([B, C], dtype=torch.float32, device=dev)
    need_grad = any(ctx.needs_input_grad[:2])
    levels = torch.empty([B, T, k, C], dtype=torch.float32, device=dev) if need_grad else None
    backptr = termptr = None
    if sr == N.MAXTROPICAL:
      backptr = torch.empty([B, T, k, C], dtype=torch.int16, device=dev)
      termptr = torch.empty([B, T, C], dtype=torch.uint8, device=dev)
    alpha_norm = None
    if (USE_NORM and T > 0 and B > 0 and
        N.lib().lt_lattice_norm_supported(sr, V, n, k, flags | N.FLAG_FORCE_GENERIC)):
      alpha_norm = torch.empty([B, T + 3], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
      N.check(N.lib().lt_lattice_forward_norm(
          sr, V, n, k, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T, None,
          N.ptr(dist), N.ptr(alphas), N.ptr(alpha_final), N.ptr(levels), N.ptr(backptr),
          N.ptr(termptr), N.ptr(alpha_norm), flags, N.stream_ptr(dev)), 'lt_lattice_forward')
    ctx.geom = (sr, V, n, k, flags)
    ctx.save_for_backward(blank, lexical, num_frames, dist, alphas, alpha_final, levels, backptr,
                          termptr, alpha_norm)
    out_alphas = _denormalized(alphas, alpha_norm)
    ctx.mark_non_differentiable(out_alphas)
    return dist, out_alphas

  @staticmethod
  def backward(ctx, g_dist, _g_alphas):
    sr, V, n, k, flags = ctx.geom
    (blank, lexical, num_frames, dist, alphas, alpha_final, levels, backptr, termptr,
     alpha_norm) = ctx.saved_tensors
    B, T, L, C = blank.shape
    dev = blank.device
    g_dist = N.require_cuda(g_dist, 'grad_dist').contiguous()
    with torch.cuda.device(dev):
      if sr != N.MAXTROPICAL:
        gb = torch.empty_like(blank)
        gl = torch.empty_like(lexical)
        N.check(N.lib().lt_lattice_backward_norm(
            sr, V, n, k, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T, N.ptr(alphas),
            N.ptr(levels), N.ptr(dist), N.ptr(g_dist), N.ptr(gb), N.ptr(gl), None,
            N.ptr(alpha_norm), flags, N.stream_ptr(dev)), 'lt_lattice_backward')
        return gb, gl, None, None, None, None, None, None
      labels = torch.empty([B, T, k + 1], dtype=torch.int32, device=dev)
      states = torch.empty([B, T + 1], dtype=torch.int32, device=dev)
      N.check(N.lib().lt_viterbi_backtrace(
          V, n, k, N.ptr(backptr), N.ptr(termptr), N.ptr(alpha_final), N.ptr(num_frames), B, T,
          N.ptr(labels), N.ptr(states), None, None, None, N.stream_ptr(dev)),
          'lt_viterbi_backtrace')
    # one-hot gradients per alignment state from the path (index arithmetic on [B,T,k+1] ints):
    # expansion i of frame t leaves context state s_i with label y_i > 0; the frame closes with
    # the blank arc of state e = number of expansions taken.
    from . import contexts as _contexts
    context = _contexts.FullNGram(vocab_size=V, context_size=n)
    gb = torch.zeros_like(blank)
    gl = torch.zeros_like(lexical)
    real = torch.arange(T, device=dev)[None, :] < num_frames[:, None]
    bi = torch.arange(B, device=dev)[:, None].expand(B, T)
    ti = torch.arange(T, device=dev)[None, :].expand(B, T)
    g = g_dist[:, None].expand(B, T)
    s = states[:, :T].long()
    alive = real
    for i in range(k + 1):
      y = labels[:, :, i].long() if i < k else torch.zeros_like(s)
      take = alive & (y > 0)
      close = alive & ~take                     # the blank arc of alignment state i
      gb.index_put_((bi[close], ti[close], torch.full_like(s[close], i), s[close]), g[close],
                    accumulate=True)
      if i < k:
        gl.index_put_((bi[take], ti[take], torch.full_like(s[take], i), s[take], y[take] - 1),
                      g[take], accumulate=True)
        s = torch.where(take, context.next_state(s, y), s)
      alive = take
    return gb, gl, None, None, None, None, None, None


def joint_lattice_forward_fused(sr, V, proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab,
                                num_frames, want_alphas=True, want_path=False):
  """JointWeightFn fused into the forward recursion (lt_joint_lattice_forward_fused): no
  [B,T,C,V] tensor exists.  proj_ctx [C,H], proj_frame [B,T,H].  Returns (dist [B], alphas
  [B,T,C] or None, labels [B,T,1] / path_states [B,T+1] when want_path (MaxTropical))."""
  proj_ctx = N.require_cuda(proj_ctx, 'proj_ctx')
  proj_frame = N.require_cuda(proj_frame, 'proj_frame')
  w_blank = N.require_cuda(w_blank.reshape(-1), 'w_blank')
  b_blank = N.require_cuda(b_blank.reshape(-1), 'b_blank')
  w_vocab = N.require_cuda(w_vocab, 'w_vocab')
  b_vocab = N.require_cuda(b_vocab, 'b_vocab')
  B, T, H = proj_frame.shape
  C = proj_ctx.shape[0]
  dev = proj_frame.device
  dist = torch.empty([B], dtype=torch.float32, device=dev)
  alphas = torch.empty([B, T, C], dtype=torch.float32, device=dev) if want_alphas else None
  backptr = alpha_final = None
  if want_path:
    backptr = torch.empty([B, T, 1, C], dtype=torch.int16, device=dev)
    alpha_final = torch.empty([B, C], dtype=torch.float32, device=dev)
  with torch.cuda.device(dev):
    N.check(N.lib().lt_joint_lattice_forward_fused(
        sr, V, N.ptr(proj_ctx), N.ptr(proj_frame), N.ptr(w_blank), N.ptr(b_blank), N.ptr(w_vocab),
        N.ptr(b_vocab), N.ptr(num_frames), B, T, H, N.ptr(dist), N.ptr(alphas),
        N.ptr(alpha_final), N.ptr(backptr), N.stream_ptr(dev)), 'lt_joint_lattice_forward_fused')
    labels = states = None
    if want_path:
      labels = torch.empty([B, T, 1], dtype=torch.int32, device=dev)
      states = torch.empty([B, T + 1], dtype=torch.int32, device=dev)
      N.check(N.lib().lt_viterbi_backtrace(
          V, 1, N.FRAME_DEPENDENT, N.ptr(backptr), None, N.ptr(alpha_final), N.ptr(num_frames),
          B, T, N.ptr(labels), N.ptr(states), None, None, None, N.stream_ptr(dev)),
          'lt_viterbi_backtrace')
  return dist, alphas, labels, states


def viterbi_path(blank, lexical, num_frames, V, n, k, flags=0):
  """MaxTropical forward + back-trace; returns (labels [B,T,k+1] int32 with true
  1-based labels, path_states [B,T+1], path_weights [B])."""
  C = blank.shape[-1]
  blank, lexical = _check_weights(blank.detach(), lexical.detach(), V, C)
  B, T, _ = blank.shape
  dev = blank.device
  dist, _, alpha_final, _, backptr, termptr, _ = _lattice_forward_raw(
      N.MAXTROPICAL, V, n, k, blank, lexical, num_frames, flags, want_levels=False,
      want_backptr=True)
  labels = torch.empty([B, T, max(k, 0) + 1], dtype=torch.int32, device=dev)
  states = torch.empty([B, T + 1], dtype=torch.int32, device=dev)
  with torch.cuda.device(dev):
    N.check(N.lib().lt_viterbi_backtrace(
        V, n, k, N.ptr(backptr), N.ptr(termptr), N.ptr(alpha_final), N.ptr(num_frames), B, T,
        N.ptr(labels), N.ptr(states), None, None, None, N.stream_ptr(dev)),
        'lt_viterbi_backtrace')
  return labels, states, dist


# ---------------------------------------------------------------------------
# generic context DFA (contexts.NextStateTable): table-driven kernels
# ---------------------------------------------------------------------------

class TableReduce(torch.autograd.Function):
  """NextStateTable.forward_reduce: out[..., q] = (+)_{p -y-> q} w[..., p, y]
  (contexts.py:74-90) with the semiring's gradient (safe Log gradient,
  first-arg-max MaxTropical gradient)."""

  @staticmethod
  def forward(ctx, weights, context, sr):
    w = N.require_cuda(weights, 'weights')
    table, offsets, arcs = context.kernel_tables(w.device)
    c, v = table.shape
    outer = w.numel() // (c * v)
    out = torch.empty(w.shape[:-1], dtype=torch.float32, device=w.device)
    argarc = (torch.empty(w.shape[:-1], dtype=torch.int32, device=w.device)
              if sr == N.MAXTROPICAL else None)
    with torch.cuda.device(w.device):
      N.check(N.lib().lt_table_reduce_forward(
          sr, N.ptr(w), N.ptr(offsets), N.ptr(arcs), outer, c, v, N.ptr(out), N.ptr(argarc),
          N.stream_ptr(w.device)), 'lt_table_reduce_forward')
    ctx.save_for_backward(w, out, argarc, table)
    ctx.geom = (sr, outer, c, v)
    return out

  @staticmethod
  def backward(ctx, g):
    w, out, argarc, table = ctx.saved_tensors
    sr, outer, c, v = ctx.geom
    g = N.require_cuda(g, 'grad')
    gw = torch.empty_like(w)
    with torch.cuda.device(w.device):
      N.check(N.lib().lt_table_reduce_backward(
          sr, N.ptr(w), N.ptr(out), N.ptr(argarc), N.ptr(g), N.ptr(table), outer, c, v,
          N.ptr(gw), N.stream_ptr(w.device)), 'lt_table_reduce_backward')
    return gw, None, None


def _table_forward_raw(sr, k, context, blank, lexical, num_frames, want_levels, want_backarc):
  B, T, C = blank.shape
  dev = blank.device
  table, offsets, arcs = context.kernel_tables(dev)
  V = table.shape[1]
  dist = torch.empty([B], dtype=torch.float32, device=dev)
  alphas = torch.empty([B, T, C], dtype=torch.float32, device=dev)
  alpha_final = torch.empty([B, C], dtype=torch.float32, device=dev)
  fld = k >= 1
  levels = (torch.empty([B, T, k, C], dtype=torch.float32, device=dev)
            if (fld and want_levels) else None)
  backarc = termptr = None
  if want_backarc and sr == N.MAXTROPICAL:
    backarc = torch.empty([B, T, max(k, 1), C], dtype=torch.int32, device=dev)
    if fld:
      termptr = torch.empty([B, T, C], dtype=torch.uint8, device=dev)
  with torch.cuda.device(dev):
    N.check(N.lib().lt_table_lattice_forward(
        sr, k, N.ptr(table), N.ptr(offsets), N.ptr(arcs), C, V, N.ptr(blank), N.ptr(lexical),
        N.ptr(num_frames), B, T, None, N.ptr(dist), N.ptr(alphas), N.ptr(alpha_final),
        N.ptr(levels), N.ptr(backarc), N.ptr(termptr), N.stream_ptr(dev)),
        'lt_table_lattice_forward')
  return dist, alphas, alpha_final, levels, backarc, termptr


class TableLatticeForward(torch.autograd.Function):
  """(dist, alphas) = RecognitionLattice._forward for a NextStateTable context."""

  @staticmethod
  def forward(ctx, blank, lexical, num_frames, context, sr, k):
    C, V = context.shape()
    blank, lexical = _check_weights(blank, lexical, V, C)
    need_grad = any(ctx.needs_input_grad[:2])
    dist, alphas, alpha_final, levels, backarc, termptr = _table_forward_raw(
        sr, k, context, blank, lexical, num_frames, need_grad, need_grad)
    ctx.geom = (sr, k, context)
    ctx.save_for_backward(blank, lexical, num_frames, dist, alphas, alpha_final, levels, backarc,
                          termptr)
    ctx.mark_non_differentiable(alphas)
    return dist, alphas

  @staticmethod
  def backward(ctx, g_dist, _g_alphas):
    sr, k, context = ctx.geom
    blank, lexical, num_frames, dist, alphas, alpha_final, levels, backarc, termptr = \
        ctx.saved_tensors
    B, T, C = blank.shape
    V = lexical.shape[-1]
    dev = blank.device
    g_dist = N.require_cuda(g_dist, 'grad_dist')
    table, _, _ = context.kernel_tables(dev)
    with torch.cuda.device(dev):
      if sr == N.MAXTROPICAL:
        gb = torch.zeros_like(blank)
        gl = torch.zeros_like(lexical)
        labels = torch.empty([B, T, max(k, 0) + 1], dtype=torch.int32, device=dev)
        N.check(N.lib().lt_table_viterbi_backtrace(
            k, C, V, N.ptr(backarc), N.ptr(termptr), N.ptr(alpha_final), N.ptr(num_frames), B, T,
            N.ptr(labels), None, N.ptr(g_dist), N.ptr(gb), N.ptr(gl), N.stream_ptr(dev)),
            'lt_table_viterbi_backtrace')
      else:
        gb = torch.empty_like(blank)
        gl = torch.empty_like(lexical)
        N.check(N.lib().lt_table_lattice_backward(
            sr, k, N.ptr(table), C, V, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T,
            N.ptr(alphas), N.ptr(levels), N.ptr(dist), N.ptr(g_dist), N.ptr(gb), N.ptr(gl),
            N.stream_ptr(dev)), 'lt_table_lattice_backward')
    return gb, gl, None, None, None, None


def table_viterbi_path(blank, lexical, num_frames, context, k):
  """MaxTropical forward + back-trace for a NextStateTable context; returns
  (labels [B,T,k+1] int32 true 1-based labels, path_states [B,T+1], path_weights [B])."""
  C, V = context.shape()
  blank, lexical = _check_weights(blank.detach(), lexical.detach(), V, C)
  B, T, _ = blank.shape
  dev = blank.device
  dist, _, alpha_final, _, backarc, termptr = _table_forward_raw(
      N.MAXTROPICAL, k, context, blank, lexical, num_frames, False, True)
  labels = torch.empty([B, T, max(k, 0) + 1], dtype=torch.int32, device=dev)
  states = torch.empty([B, T + 1], dtype=torch.int32, device=dev)
  with torch.cuda.device(dev):
    N.check(N.lib().lt_table_viterbi_backtrace(
        k, C, V, N.ptr(backarc), N.ptr(termptr), N.ptr(alpha_final), N.ptr(num_frames), B, T,
        N.ptr(labels), N.ptr(states), None, None, None, N.stream_ptr(dev)),
        'lt_table_viterbi_backtrace')
  return labels, states, dist


# ---------------------------------------------------------------------------
# K3: numerator on the label lattice
# ---------------------------------------------------------------------------

def walk_states(labels: torch.Tensor, num_labels, V: int, n: int):
  """labels [B,U] int32 (CUDA), num_labels [B] int32 or None ->
  (states [B,U+1], next_labels [B,U+1], bad [1]) int32:
  FullNGram.walk_states (contexts.py:109-146) and labels ++ [1] with label 0
  read as label 1 (lattices.py:314-315, :336-338), in one kernel.  Positions
  u >= num_labels[b] are read as epsilon; `bad` counts labels outside [0, V] before that."""
  labels = N.require_cuda(labels, 'labels', torch.int32)
  B, U = labels.shape
  states = torch.empty([B, U + 1], dtype=torch.int32, device=labels.device)
  next_labels = torch.empty([B, U + 1], dtype=torch.int32, device=labels.device)
  bad = torch.zeros([1], dtype=torch.int32, device=labels.device)
  with torch.cuda.device(labels.device):
    N.check(N.lib().lt_walk_states_checked(
        V, n, N.ptr(labels), N.ptr(num_labels), B, U, N.ptr(states), N.ptr(next_labels),
        N.ptr(bad), N.stream_ptr(labels.device)), 'lt_walk_states')
  return states, next_labels, bad


_SIDE_STREAMS = {}


def _side_stream(device):
  """One extra stream per device for the (tiny, latency-bound) numerator
  kernels, so that they overlap the HBM-bound denominator kernels."""
  key = (device.type, device.index)
  if key not in _SIDE_STREAMS:
    _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
  return _SIDE_STREAMS[key]


# Head start (ns) K1 gets over the numerator kernels on the side stream (lt_stream_delay)
NUMERATOR_STAGGER_NS = int(os.environ.get('LT_NUMERATOR_STAGGER_NS', '40000'))


def _string_forward_raw(sr, k, V, C, blank, lexical, num_frames, states, next_labels, num_labels,
                        need_grad, side=None, ready=None, stagger_ns=0):
  """gather + string forward.  With `side` (a torch.cuda.Stream) the two kernels
  are enqueued there, after the event `ready` (or, without one, after everything
  already on the current stream); the caller joins with
  current_stream().wait_stream(side).  Buffers are always allocated on the
  current stream, so the caching allocator stays correct."""
  B, T, _ = blank.shape
  U1 = states.shape[-1]
  dev = blank.device
  bw = torch.empty([B, T, U1], dtype=torch.float32, device=dev)
  lw = torch.empty([B, T, U1], dtype=torch.float32, device=dev)
  dist = torch.empty([B], dtype=torch.float32, device=dev)
  # (integer part, fraction) chain where the kernels have it (lt_string_forward_norm); it is
  # used with and without gradients so that the value does not depend on requires_grad
  use_ext = bool(USE_NORM and B > 0 and T > 0 and N.lib().lt_string_norm_supported(sr, k, U1))
  alphas = (torch.empty([B, T, U1], dtype=torch.float32, device=dev)
            if (need_grad and sr != N.MAXTROPICAL) or use_ext else None)
  backptr = (torch.empty([B, T, U1], dtype=torch.uint8, device=dev)
             if need_grad and sr == N.MAXTROPICAL else None)
  alpha_exp = dist_norm = None
  if use_ext:
    alpha_exp = torch.empty([B, T, U1], dtype=torch.int32, device=dev)
    dist_norm = torch.empty([B, 2], dtype=torch.int32, device=dev)
  with torch.cuda.device(dev):
    if side is not None:
      if ready is not None:
        side.wait_event(ready)
      else:
        side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):           # stream(None) is a no-op
      L = N.lib()
      if side is not None and stagger_ns:
        N.check(L.lt_stream_delay(stagger_ns, N.stream_ptr(dev)), 'lt_stream_delay')
      N.check(L.lt_string_gather(V, C, N.ptr(blank), N.ptr(lexical), N.ptr(states),
                                 N.ptr(next_labels), B, T, U1, N.ptr(bw), N.ptr(lw),
                                 N.stream_ptr(dev)), 'lt_string_gather')
      N.check(L.lt_string_forward_norm(sr, k, N.ptr(bw), N.ptr(lw), N.ptr(num_frames),
                                       N.ptr(num_labels), B, T, U1, N.ptr(dist), N.ptr(alphas),
                                       N.ptr(backptr), N.ptr(alpha_exp), N.ptr(dist_norm),
                                       N.stream_ptr(dev)), 'lt_string_forward')
  return dist, bw, lw, alphas, backptr, (alpha_exp, dist_norm)


def _string_backward(sr, k, bw, lw, num_frames, num_labels, alphas, backptr, dist, g_dist,
                     side=None, ready=None, out=None, ext=(None, None), no_wait=False):
  """Numerator posteriors on the label lattice: (grad_blank_w, grad_lexical_w).
  `ext` = (alpha_exp, dist_norm) of a lt_string_forward_norm run, or (None, None)."""
  alpha_exp, dist_norm = ext
  B, T, U1 = bw.shape
  dev = bw.device
  gbw, glw = out if out is not None else (torch.empty_like(bw), torch.empty_like(lw))
  with torch.cuda.device(dev):
    if side is not None and not no_wait:      # no_wait: `side` already follows the producers
      if ready is not None:
        side.wait_event(ready)
      else:
        side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      N.check(N.lib().lt_string_backward_norm(
          sr, k, N.ptr(bw), N.ptr(lw), N.ptr(num_frames), N.ptr(num_labels), B, T, U1,
          N.ptr(alphas), N.ptr(backptr), N.ptr(dist), N.ptr(g_dist), N.ptr(gbw), N.ptr(glw),
          N.ptr(alpha_exp), N.ptr(dist_norm), N.stream_ptr(dev)), 'lt_string_backward')
  return gbw, glw


def _string_scatter(V, C, gbw, glw, states, next_labels, scale, gb, gl, utt_scale=None,
                    split=False):
  """grad_dense[b, t, states[u], next_labels[u] - 1] += scale * utt_scale[b] * grad_w[b, t, u]."""
  B, T, U1 = gbw.shape
  dev = gbw.device
  fn = N.lib().lt_string_scatter_add_split if split else N.lib().lt_string_scatter_add
  with torch.cuda.device(dev):
    N.check(fn(
        V, C, N.ptr(gbw), N.ptr(glw), N.ptr(states), N.ptr(next_labels), B, T, U1, float(scale),
        N.ptr(utt_scale), N.ptr(gb), N.ptr(gl), N.stream_ptr(dev)), 'lt_string_scatter_add')


class StringChainForward(torch.autograd.Function):
  """Shortest distance on the T x (U+1) label lattice from the ALREADY GATHERED per-position
  weights blank_w / lexical_w [B,T,U1] (shortest_distance_step_scan, lattices.py:347-377, with
  alignment.string_forward): the numerator when the weight function was only evaluated on the
  U+1 states of the label string (lattices.py:300-313) instead of on all context states."""

  @staticmethod
  def forward(ctx, bw, lw, num_frames, num_labels, sr, k):
    bw = N.require_cuda(bw, 'blank_w')
    lw = N.require_cuda(lw, 'lexical_w')
    B, T, U1 = bw.shape
    dev = bw.device
    need_grad = any(ctx.needs_input_grad[:2])
    dist = torch.empty([B], dtype=torch.float32, device=dev)
    use_ext = bool(USE_NORM and B > 0 and T > 0 and N.lib().lt_string_norm_supported(sr, k, U1))
    alphas = (torch.empty([B, T, U1], dtype=torch.float32, device=dev)
              if (need_grad and sr != N.MAXTROPICAL) or use_ext else None)
    backptr = (torch.empty([B, T, U1], dtype=torch.uint8, device=dev)
               if need_grad and sr == N.MAXTROPICAL else None)
    alpha_exp = dist_norm = None
    if use_ext:
      alpha_exp = torch.empty([B, T, U1], dtype=torch.int32, device=dev)
      dist_norm = torch.empty([B, 2], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
      N.check(N.lib().lt_string_forward_norm(
          sr, k, N.ptr(bw), N.ptr(lw), N.ptr(num_frames), N.ptr(num_labels), B, T, U1,
          N.ptr(dist), N.ptr(alphas), N.ptr(backptr), N.ptr(alpha_exp), N.ptr(dist_norm),
          N.stream_ptr(dev)), 'lt_string_forward')
    ctx.geom = (sr, k)
    ctx.save_for_backward(bw, lw, num_frames, num_labels, alphas, backptr, dist, alpha_exp,
                          dist_norm)
    return dist

  @staticmethod
  def backward(ctx, g_dist):
    sr, k = ctx.geom
    bw, lw, num_frames, num_labels, alphas, backptr, dist, alpha_exp, dist_norm = \
        ctx.saved_tensors
    g_dist = N.require_cuda(g_dist, 'grad_dist')
    gbw, glw = _string_backward(sr, k, bw, lw, num_frames, num_labels, alphas, backptr, dist,
                                g_dist, ext=(alpha_exp, dist_norm))
    return gbw, glw, None, None, None, None


class StringForward(torch.autograd.Function):
  """RecognitionLattice._string_forward on dense weights (lattices.py:250-377)."""

  @staticmethod
  def forward(ctx, blank, lexical, num_frames, states, next_labels, num_labels, sr, V, k):
    C = blank.shape[-1]
    blank, lexical = _check_weights(blank, lexical, V, C)
    need_grad = any(ctx.needs_input_grad[:2])
    dist, bw, lw, alphas, backptr, ext = _string_forward_raw(
        sr, k, V, C, blank, lexical, num_frames, states, next_labels, num_labels, need_grad)
    ctx.geom = (sr, V, C, k, tuple(blank.shape))
    ctx.save_for_backward(bw, lw, num_frames, num_labels, alphas, backptr, dist, states,
                          next_labels, *ext)
    return dist

  @staticmethod
  def backward(ctx, g_dist):
    sr, V, C, k, shape = ctx.geom
    (bw, lw, num_frames, num_labels, alphas, backptr, dist, states, next_labels, alpha_exp,
     dist_norm) = ctx.saved_tensors
    g_dist = N.require_cuda(g_dist, 'grad_dist')
    gb = torch.zeros(shape, dtype=torch.float32, device=bw.device)
    gl = torch.zeros((*shape, V), dtype=torch.float32, device=bw.device)
    gbw, glw = _string_backward(sr, k, bw, lw, num_frames, num_labels, alphas, backptr, dist,
                                g_dist, ext=(alpha_exp, dist_norm))
    _string_scatter(V, C, gbw, glw, states, next_labels, 1.0, gb, gl)
    return gb, gl, None, None, None, None, None, None, None


def _loss_forward(blank, lexical, num_frames, states, next_labels, num_labels, V, n, k, flags,
                  need_grad):
  """K1 (Log) over the dense lattice on a high-priority stream; beside it, on a side stream,
  the whole numerator: gather, K3 forward and (when gradients are needed) K3 backward, i.e. the
  UNSCALED numerator posteriors -- the label lattice is tiny, so its three kernels fit under the
  HBM-bound K1.  Returns (log_z, num, saved) with `saved` the tensors _loss_backward needs."""
  C = blank.shape[-1]
  dev = blank.device
  cur = torch.cuda.current_stream(dev)
  side = _side_stream(dev)
  ready = torch.cuda.Event()
  ready.record(cur)
  # K1 stays on the CURRENT stream, directly behind the kernel that produced the weights: it
  # becomes launchable the moment that kernel retires, while the numerator kernels first have to
  # see the cross-stream event -- so K1's CTAs (which fill an SM's register file exactly, two
  # per SM on 128 of the 148 SMs) are placed first and the small numerator CTAs take the SMs
  # that are left.  The other way round, every SM a numerator CTA sits on is lost to K1 and a
  # whole cluster (utterance) waits for it to retire.
  # NOTE: every buffer handed to a kernel on `side` must stay referenced until the join below
  # -- a tensor dropped earlier returns to the current stream's pool and the next allocation
  # may alias it while the other stream still writes to it.
  fwd = _lattice_forward_raw(
      N.LOG, V, n, k, blank, lexical, num_frames, flags, want_levels=need_grad,
      want_backptr=False, norm=True)
  log_z, alphas, _, levels, _, _, alpha_norm = fwd
  strf = _string_forward_raw(
      N.LOG, k, V, C, blank, lexical, num_frames, states, next_labels, num_labels, need_grad,
      side=side, ready=ready, stagger_ns=NUMERATOR_STAGGER_NS)
  num, bw, lw, s_alphas, _, ext = strf
  gbw = glw = None
  if need_grad:
    # The loss only needs the numerator's VALUE: the current stream joins after K3 forward, and
    # K3 backward (the unscaled numerator posteriors, unit upstream gradient) runs on the side
    # stream behind it, under whatever the current stream does next -- normally K2 -- and is
    # joined by _loss_backward before the scatter.  For FrameDependent it fits under K1 either
    # way; the FrameLabelDependent chain (double state, k + 1 terms per step) does not.
    # record_stream: these buffers are in use on `side` beyond this function, the allocator must
    # not hand them back to the current stream's pool before the side stream is done with them.
    num_ready = torch.cuda.Event()
    num_ready.record(side)
    cur.wait_event(num_ready)
    gbw, glw = _string_backward(N.LOG, k, bw, lw, num_frames, num_labels, s_alphas, None, num,
                                None, side=side, ext=ext, no_wait=True)
    for x in (bw, lw, s_alphas, num, gbw, glw, num_frames, num_labels) + tuple(ext):
      if x is not None:
        x.record_stream(side)
  else:
    cur.wait_stream(side)
  del fwd, strf
  return log_z, num, (alphas, levels, gbw, glw, alpha_norm)


def _loss_backward(blank, lexical, num_frames, log_z, saved, states, next_labels, V, n, k, flags,
                   g_den, g_numr, split):
  """K2 writes g_den * (arc posteriors) straight into the weight-gradient buffers (as split rows
  when `split`); the stored numerator posteriors are then scattered in, weighted by g_numr."""
  alphas, levels, gbw, glw, alpha_norm = saved
  B, T, C = blank.shape
  dev = blank.device
  g_den = N.require_cuda(g_den, 'grad').contiguous()
  g_numr = N.require_cuda(g_numr, 'grad').contiguous()
  gb = torch.empty_like(blank)
  gl = torch.empty_like(lexical)
  if split:
    flags |= N.FLAG_GRAD_SPLIT
  with torch.cuda.device(dev):
    N.check(N.lib().lt_lattice_backward_norm(
        N.LOG, V, n, k, N.ptr(blank), N.ptr(lexical), N.ptr(num_frames), B, T, N.ptr(alphas),
        N.ptr(levels), N.ptr(log_z), N.ptr(g_den), N.ptr(gb), N.ptr(gl), None,
        N.ptr(alpha_norm), flags, N.stream_ptr(dev)), 'lt_lattice_backward')
  torch.cuda.current_stream(dev).wait_stream(_side_stream(dev))   # K3 backward (_loss_forward)
  _string_scatter(V, C, gbw, glw, states, next_labels, 1.0, gb, gl, utt_scale=g_numr,
                  split=split)
  return gb, gl


class LatticeLoss(torch.autograd.Function):
  """loss = logZ - numerator (lattices.py:131-183) on dense arc weights as ONE autograd node
  (see _loss_forward / _loss_backward).  Returns (loss, log_z, numerator)."""

  @staticmethod
  def forward(ctx, blank, lexical, num_frames, states, next_labels, num_labels, V, n, k, flags):
    C = blank.shape[-1]
    blank, lexical = _check_weights(blank, lexical, V, C)
    need_grad = any(ctx.needs_input_grad[:2])
    log_z, num, saved = _loss_forward(blank, lexical, num_frames, states, next_labels,
                                      num_labels, V, n, k, flags, need_grad)
    ctx.geom = (V, n, k, flags)
    ctx.save_for_backward(blank, lexical, num_frames, log_z, states, next_labels, *saved)
    return log_z - num, log_z, num

  @staticmethod
  def backward(ctx, g_loss, g_logz, g_num):
    V, n, k, flags = ctx.geom
    blank, lexical, num_frames, log_z, states, next_labels, *saved = ctx.saved_tensors
    g_den = g_loss if g_logz is None else g_loss + g_logz
    g_numr = -g_loss if g_num is None else g_num - g_loss
    gb, gl = _loss_backward(blank, lexical, num_frames, log_z, saved, states, next_labels, V, n,
                            k, flags, g_den, g_numr, split=False)
    return gb, gl, None, None, None, None, None, None, None, None


# Hand the arc posteriors from the lattice backward to the joint network's backward as "split
# rows" (include/last_lattice.h: LT_FLAG_GRAD_SPLIT) where both kernels support it.
# LT_NO_SPLIT_GRAD=1 (read once, at import) or RecognitionLattice.split_grad_handover = False
# keeps them in float32.
SPLIT_GRAD_DEFAULT = not os.environ.get('LT_NO_SPLIT_GRAD')


class JointLatticeLoss(torch.autograd.Function):
  """JointWeightFn over all frames (weight_fns.py:194-227) + GNAT loss (lattices.py:131-183) as
  ONE autograd node: proj_ctx [C,H], proj_frame [N,H] and the two output projections in, loss
  [B] out.  The dense arc weights and their gradients are internals of the node -- no autograd
  tensor, hook or accumulation ever sees them -- which is what makes the split-row hand-over
  (K2 emits [V bf16 hi | V bf16 lo] rows, the tensor-core dgrad / wgrad load them as operands)
  safe: this is the streaming-gradient intent of lattices.py:644-799 (_backward +
  BackwardStepCallback), with the callback replaced by the joint network's own backward."""

  @staticmethod
  def forward(ctx, proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab, num_frames, states,
              next_labels, num_labels, B, T, V, n, k, flags, allow_split):
    from . import joint
    blank, lexical = joint.joint_forward_raw(proj_ctx, proj_frame, w_blank, b_blank, w_vocab,
                                             b_vocab)
    C = proj_ctx.shape[0]
    blank = blank.reshape(B, T, C)
    lexical = lexical.reshape(B, T, C, V)
    need_grad = any(ctx.needs_input_grad[:6])
    log_z, num, saved = _loss_forward(blank, lexical, num_frames, states, next_labels,
                                      num_labels, V, n, k, flags, need_grad)
    ctx.geom = (B, T, V, n, k, flags, bool(allow_split))
    ctx.save_for_backward(proj_ctx, proj_frame, w_blank, w_vocab, blank, lexical, num_frames,
                          log_z, states, next_labels, *saved)
    return log_z - num, log_z, num

  @staticmethod
  def backward(ctx, g_loss, g_logz, g_num):
    from . import joint
    B, T, V, n, k, flags, allow_split = ctx.geom
    (proj_ctx, proj_frame, w_blank, w_vocab, blank, lexical, num_frames, log_z, states,
     next_labels, *saved) = ctx.saved_tensors
    g_den = g_loss if g_logz is None else g_loss + g_logz
    g_numr = -g_loss if g_num is None else g_num - g_loss
    nfr, h = proj_frame.shape
    c = proj_ctx.shape[0]
    U1 = states.shape[-1]
    split = bool(allow_split and U1 <= 4096 and
                 N.lib().lt_joint_backward_split_supported(nfr, c, h, V) and
                 N.lib().lt_lattice_backward_split_supported(N.LOG, V, n, k, flags))
    gb, gl = _loss_backward(blank, lexical, num_frames, log_z, saved, states, next_labels, V, n,
                            k, flags, g_den, g_numr, split=split)
    grads = joint.joint_backward_raw(proj_ctx, proj_frame, w_blank, w_vocab, gb.reshape(nfr, c),
                                     gl.reshape(nfr, c, V), fmt=1 if split else 0)
    return (*grads, None, None, None, None, None, None, None, None, None, None, None)
