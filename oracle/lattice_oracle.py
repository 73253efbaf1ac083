"""CPU oracle for the last_torch lattice hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference algorithm
(theadamsabra/last_torch, /root/reference at survey time).  It exists only to
check the CUDA kernels: nothing in the product package `last_torch_b200/`
imports it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import this module.

Parity pinning: every function here is checked in `tests/test_oracle_golden.py`
against (1) the hand-expanded known-answer vectors of the reference's own test
suite (restated there with file:line citations) and (2) fixtures under
`tests/golden/*.npz` that were produced by importing the unmodified reference
in the build container (`tests/golden/make_golden.py`, committed).

Every function cites the reference file:line it follows.  Array layout:
  blank   [B, T, C]       weight of the blank arc leaving context state c at frame t
  lexical [B, T, C, V]    weight of the arc with label y+1 leaving context state c
All arithmetic runs in the dtype of the inputs (float32 or float64).
"""

from __future__ import annotations

import numpy as np

REAL, LOG, MAXTROPICAL = 0, 1, 2
SEMIRING_NAMES = {'Real': REAL, 'Log': LOG, 'MaxTropical': MAXTROPICAL}


# ---------------------------------------------------------------------------
# Semirings (reference: last_torch/semirings.py)
# ---------------------------------------------------------------------------

def sr_zero(sr, dtype=np.float32):
  """semirings.py:147-150 (Real), :188-191 (Log), :316-319 (MaxTropical)."""
  return dtype(0.0) if sr == REAL else dtype(-np.inf)


def sr_one(sr, dtype=np.float32):
  """semirings.py:153-154, :194-195, :322-323."""
  return dtype(1.0) if sr == REAL else dtype(0.0)


def sr_times(sr, a, b):
  """semirings.py:157-158 (a*b), :198-199 and :326-327 (a+b)."""
  return a * b if sr == REAL else a + b


def _logaddexp(a, b):
  """semirings.py:247-255: c=max(a,b); non-finite c replaced by 0."""
  a, b = np.broadcast_arrays(a, b)
  with np.errstate(all='ignore'):
    c = np.maximum(a, b)
    c = np.where(np.isfinite(c), c, 0).astype(a.dtype)
    z = np.exp(a - c) + np.exp(b - c)
    return c + np.log(z)


def _logsumexp(a, axis):
  """semirings.py:279-286."""
  with np.errstate(all='ignore'):
    c = np.max(a, axis=axis, keepdims=True)
    c = np.where(np.isfinite(c), c, 0).astype(a.dtype)
    z = np.sum(np.exp(a - c), axis=axis, keepdims=True)
    return np.squeeze(c, axis=axis) + np.log(np.squeeze(z, axis=axis))


def sr_plus(sr, a, b):
  """semirings.py:161-162, :202-204, :330-332."""
  if sr == REAL:
    return a + b
  if sr == LOG:
    return _logaddexp(a, b)
  return np.maximum(a, b)


def sr_sum(sr, a, axis):
  """semirings.py:169-170, :211-220, :339-348 (empty axis -> semiring zero)."""
  a = np.asarray(a)
  if sr == REAL:
    return np.sum(a, axis=axis)
  if a.size == 0:
    ax = axis + a.ndim if axis < 0 else axis
    shape = a.shape[:ax] + a.shape[ax + 1:]
    return np.full(shape, -np.inf, dtype=a.dtype)
  if sr == LOG:
    return _logsumexp(a, axis)
  return np.max(a, axis=axis)


def logaddexp_grad(a, b, g):
  """Intended safe gradient of Log.plus, semirings.py:222-241 and :264-269."""
  a, b = np.broadcast_arrays(a, b)
  with np.errstate(all='ignore'):
    c = np.maximum(a, b)
    c = np.where(np.isfinite(c), c, 0)
    ea, eb = np.exp(a - c), np.exp(b - c)
    z = ea + eb
    z = np.where(z != 0, z, 1)
    return g / z * ea, g / z * eb


def logsumexp_grad(a, axis, g):
  """Intended safe gradient of Log.sum, semirings.py:222-241 and :296-300."""
  with np.errstate(all='ignore'):
    c = np.max(a, axis=axis, keepdims=True)
    c = np.where(np.isfinite(c), c, 0)
    e = np.exp(a - c)
    z = np.sum(e, axis=axis, keepdims=True)
    z = np.where(z != 0, z, 1)
    return np.expand_dims(g, axis) / z * e


def maximum_grad(a, b, g):
  """semirings.py:360-369: ties go to `a` (a >= b)."""
  a, b = np.broadcast_arrays(a, b)
  choose_a = (a >= b).astype(a.dtype)
  return g * choose_a, g * (1 - choose_a)


def max_grad(a, axis, g):
  """semirings.py:380-398: one-hot at the FIRST arg-max along `axis`."""
  idx = np.argmax(a, axis=axis)
  mask = np.zeros_like(a)
  np.put_along_axis(mask, np.expand_dims(idx, axis), 1, axis=axis)
  return np.expand_dims(g, axis) * mask


# ---------------------------------------------------------------------------
# FullNGram context dependency (reference: last_torch/contexts.py:149-263)
# ---------------------------------------------------------------------------

class FullNGram:
  """contexts.py:149-263."""

  def __init__(self, vocab_size, context_size):
    self.vocab_size = int(vocab_size)
    self.context_size = int(context_size)

  def num_states(self):
    """contexts.py:181-182."""
    return sum(self.vocab_size**i for i in range(self.context_size + 1))

  def shape(self):
    return self.num_states(), self.vocab_size

  def start(self):
    return 0

  def next_state(self, state, label):
    """contexts.py:190-205 (epsilon label 0 stays in place)."""
    state = np.asarray(state).astype(np.int64)
    label = np.asarray(label).astype(np.int64)
    v, n = self.vocab_size, self.context_size
    num_asc = sum(v**i for i in range(n))
    ascend = state * v + label
    if n == 0:
      full = np.zeros_like(ascend)
    else:
      full = (state - num_asc) % (v**(n - 1)) * v + num_asc + label - 1
    nxt = np.where(state < num_asc, ascend, full)
    return np.where(label == 0, state, nxt)

  def next_state_table(self):
    """contexts.py:258-263."""
    c, v = self.shape()
    return self.next_state(np.arange(c)[:, None], np.arange(v)[None, :] + 1)

  def forward_reduce(self, weights, sr):
    """contexts.py:207-230: out[q] = (+)_{p -y-> q} weights[p, y]."""
    v, n = self.vocab_size, self.context_size
    batch = weights.shape[:-2]
    parts = []
    if n > 0:
      parts.append(np.full(batch + (1,), sr_zero(sr, weights.dtype.type),
                           dtype=weights.dtype))
    low = sum(v**i for i in range(0, n - 1))
    parts.append(weights[..., :low, :].reshape(batch + (-1,)))
    parts.append(sr_sum(sr, weights[..., low:, :].reshape(batch + (-1, v**n)),
                        axis=-2))
    return np.concatenate(parts, axis=-1)

  def backward_broadcast(self, weights):
    """contexts.py:232-256: out[p, y] = weights[next_state(p, y)]."""
    v, n = self.vocab_size, self.context_size
    batch = weights.shape[:-1]
    if n == 0:
      return np.broadcast_to(weights[..., None], weights.shape + (v,))
    num_asc = sum(v**i for i in range(n))
    part_a = weights[..., 1:num_asc].reshape(batch + (-1, v))
    part_b = np.broadcast_to(weights[..., None, num_asc:],
                             batch + (1 + v, v**n)).reshape(batch + (-1, v))
    return np.concatenate([part_a, part_b], axis=-2)

  def walk_states(self, labels):
    """contexts.py:109-146 (int64 here; the reference returns float32)."""
    labels = np.asarray(labels).astype(np.int64)
    state = np.zeros(labels.shape[:-1], dtype=np.int64)
    out = [state]
    for i in range(labels.shape[-1]):
      state = self.next_state(state, labels[..., i])
      out.append(state)
    return np.stack(out, axis=-1)


class NextStateTable:
  """contexts.py:266-324: a context DFA given as a [num_states, vocab_size]
  table, next_state_table[p, y - 1] = state reached from p with label y.

  forward_reduce follows the interface contract out[q] = (+)_{p -y-> q} w[p, y]
  (contexts.py:74-90) in EVERY semiring; the reference body (contexts.py:306-317)
  only realises it for the Real semiring, where its golden test
  (tests/contexts_test.py:214-220) agrees with this restatement.
  """

  def __init__(self, next_state_table):
    self.table = np.asarray(next_state_table).astype(np.int64)
    assert self.table.ndim == 2 and self.table.size > 0

  def shape(self):
    return self.table.shape

  def start(self):
    return 0

  def next_state(self, state, label):
    """contexts.py:297-304 (epsilon label 0 stays in place)."""
    state = np.asarray(state).astype(np.int64)
    label = np.asarray(label).astype(np.int64)
    nxt = self.table[state, np.where(label == 0, 0, label - 1)]
    return np.where(label == 0, state, nxt)

  def next_state_table(self):
    return self.table

  def forward_reduce(self, weights, sr):
    c, v = self.table.shape
    batch = weights.shape[:-2]
    flat = weights.reshape(batch + (c * v,))
    dest = self.table.reshape(-1)
    out = np.full(batch + (c,), sr_zero(sr, weights.dtype.type), dtype=weights.dtype)
    for q in range(c):
      arcs = np.nonzero(dest == q)[0]
      if arcs.size:
        out[..., q] = sr_sum(sr, flat[..., arcs], axis=-1)
    return out

  def backward_broadcast(self, weights):
    """contexts.py:319-324."""
    return weights[..., self.table]

  def walk_states(self, labels):
    """contexts.py:109-146."""
    labels = np.asarray(labels).astype(np.int64)
    state = np.zeros(labels.shape[:-1], dtype=np.int64)
    out = [state]
    for i in range(labels.shape[-1]):
      state = self.next_state(state, labels[..., i])
      out.append(state)
    return np.stack(out, axis=-1)


# ---------------------------------------------------------------------------
# Alignment lattices (reference: last_torch/alignments.py)
#   max_expansions == 0  <=>  FrameDependent (alignments.py:266-329)
#   max_expansions == k  <=>  FrameLabelDependent(k) (alignments.py:331-432)
# Blank / lexical weights are alignment-state invariant (lattices.py:447-449),
# so a single (blank, lexical) pair per frame is passed.
# ---------------------------------------------------------------------------

def shift_down(x, sr):
  """alignments.py:233-248."""
  z = np.full(x.shape[:-1] + (1,), sr_zero(sr, x.dtype.type), dtype=x.dtype)
  return np.concatenate([z, x[..., :-1]], axis=-1)


def frame_forward(alpha, blank, lexical, context, sr, max_expansions=0,
                  frame_dependent=True):
  """One frame of the forward recursion.

  FrameDependent: alignments.py:294-297.
  FrameLabelDependent: alignments.py:370-376.
  """
  if frame_dependent:
    return sr_plus(
        sr, sr_times(sr, alpha, blank),
        context.forward_reduce(sr_times(sr, alpha[..., None], lexical), sr))
  terminated = [sr_times(sr, alpha, blank)]
  last = alpha
  for _ in range(max_expansions):
    last = context.forward_reduce(sr_times(sr, last[..., None], lexical), sr)
    terminated.append(sr_times(sr, last, blank))
  return sr_sum(sr, np.stack(terminated), axis=0)


def frame_backward(alpha, blank, lexical, beta, log_z, context,
                   max_expansions=0, frame_dependent=True):
  """One frame of the (Log-semiring) backward recursion with arc marginals.

  FrameDependent: alignments.py:300-318.
  FrameLabelDependent: alignments.py:378-418.
  Returns (next_beta, blank_marginal summed over alignment states,
  lexical_marginal summed over alignment states) as lattices.py:772-773 does.
  """
  with np.errstate(all='ignore'):
    if frame_dependent:
      blank_beta = blank + beta
      lexical_beta = lexical + context.backward_broadcast(beta)
      log_scale = alpha - log_z[..., None]
      blank_marginal = np.exp(blank_beta + log_scale)
      lexical_marginal = np.exp(lexical_beta + log_scale[..., None])
      next_beta = _logaddexp(blank_beta, _logsumexp(lexical_beta, axis=-1))
      return next_beta, blank_marginal, lexical_marginal
    k = max_expansions
    lexical_alphas = [alpha]
    last = alpha
    for _ in range(k):
      last = context.forward_reduce(last[..., None] + lexical, LOG)
      lexical_alphas.append(last)
    blank_log_scale = beta - log_z[..., None]
    blank_marginal = 0
    for i in range(k + 1):
      blank_marginal = blank_marginal + np.exp(
          lexical_alphas[i] + blank + blank_log_scale)
    next_beta = blank + beta
    lexical_marginal = 0
    for i in range(k):
      j = k - 1 - i
      lexical_beta = lexical + context.backward_broadcast(next_beta)
      log_scale = lexical_alphas[j] - log_z[..., None]
      lexical_marginal = lexical_marginal + np.exp(
          lexical_beta + log_scale[..., None])
      next_beta = _logaddexp(blank + beta, _logsumexp(lexical_beta, axis=-1))
    if k == 0:
      lexical_marginal = np.zeros_like(lexical)
    return next_beta, blank_marginal, lexical_marginal


def frame_string_forward(alpha, blank, lexical, sr, max_expansions=0,
                         frame_dependent=True):
  """alignments.py:327-329 (FrameDependent), :427-432 (FrameLabelDependent)."""
  if frame_dependent:
    return sr_plus(sr, sr_times(sr, alpha, blank),
                   shift_down(sr_times(sr, alpha, lexical), sr))
  terminated = [sr_times(sr, alpha, blank)]
  last = alpha
  for _ in range(max_expansions):
    last = shift_down(sr_times(sr, last, lexical), sr)
    terminated.append(sr_times(sr, last, blank))
  return sr_sum(sr, np.stack(terminated), axis=0)


# ---------------------------------------------------------------------------
# Recognition lattice (reference: last_torch/lattices.py)
# ---------------------------------------------------------------------------

def _init_alpha(batch, num_states, start, sr, dtype):
  """lattices.py:801-807."""
  a = np.full((batch, num_states), sr_zero(sr, dtype.type), dtype=dtype)
  a[:, start] = sr_one(sr, dtype.type)
  return a


def lattice_forward(blank, lexical, num_frames, context, sr, max_expansions=0,
                    frame_dependent=True):
  """RecognitionLattice._forward, lattices.py:379-496 + scan :856-892.

  Returns (shortest_distance [B], alpha_0_to_T_minus_1 [B, T, C]).
  """
  b, t_max, c = blank.shape
  num_frames = np.asarray(num_frames)
  alpha = _init_alpha(b, c, context.start(), sr, blank.dtype)
  alphas = np.empty((b, t_max, c), dtype=blank.dtype)
  for t in range(t_max):
    alphas[:, t] = alpha
    nxt = frame_forward(alpha, blank[:, t], lexical[:, t], context, sr,
                        max_expansions, frame_dependent)
    is_padding = (t >= num_frames)[:, None]          # lattices.py:460-461
    alpha = np.where(is_padding, alpha, nxt)
  return sr_sum(sr, alpha, axis=-1), alphas


def lattice_marginals(blank, lexical, num_frames, context, max_expansions=0,
                      frame_dependent=True):
  """Intent of RecognitionLattice._backward, lattices.py:686-799.

  Forward alphas, then the backward recursion with alignment.backward and the
  padding masks of lattices.py:775-779.  Under the Log semiring the summed
  marginals equal d logZ / d weights (lattices.py:539-557).
  Returns (log_z [B], blank_marginal [B,T,C], lexical_marginal [B,T,C,V]).
  """
  b, t_max, c = blank.shape
  num_frames = np.asarray(num_frames)
  log_z, alphas = lattice_forward(blank, lexical, num_frames, context, LOG,
                                  max_expansions, frame_dependent)
  beta = np.zeros((b, c), dtype=blank.dtype)          # lattices.py:789-790
  gb = np.zeros_like(blank)
  gl = np.zeros_like(lexical)
  for t in range(t_max - 1, -1, -1):
    nb, bm, lm = frame_backward(alphas[:, t], blank[:, t], lexical[:, t], beta,
                                log_z, context, max_expansions,
                                frame_dependent)
    is_padding = (t >= num_frames)[:, None]
    beta = np.where(is_padding, beta, nb)
    gb[:, t] = np.where(is_padding, 0, bm)
    gl[:, t] = np.where(is_padding[..., None], 0, lm)
  return log_z, gb, gl


def real_lattice_grads(blank, lexical, num_frames, context, max_expansions=0,
                       frame_dependent=True):
  """d (sum_c alpha_T[c]) / d weights under the Real semiring.

  The reference obtains this with plain torch autograd through
  lattices.py:436-462; this is the same derivative written as a backward
  recursion (beta_T = 1).
  """
  b, t_max, c = blank.shape
  num_frames = np.asarray(num_frames)
  dist, alphas = lattice_forward(blank, lexical, num_frames, context, REAL,
                                 max_expansions, frame_dependent)
  beta = np.ones((b, c), dtype=blank.dtype)
  gb = np.zeros_like(blank)
  gl = np.zeros_like(lexical)
  k = 0 if frame_dependent else max_expansions
  for t in range(t_max - 1, -1, -1):
    bl, lx = blank[:, t], lexical[:, t]
    lasts = [alphas[:, t]]
    for _ in range(k):
      lasts.append(context.forward_reduce(lasts[-1][..., None] * lx, REAL))
    if frame_dependent:
      bb = context.backward_broadcast(beta)
      g_b = alphas[:, t] * beta
      g_l = alphas[:, t][..., None] * bb
      nb = bl * beta + np.sum(lx * bb, axis=-1)
    else:
      g_b = sum(lasts) * beta
      nb = bl * beta
      g_l = np.zeros_like(lx)
      for j in range(k - 1, -1, -1):
        bb = context.backward_broadcast(nb)
        g_l = g_l + lasts[j][..., None] * bb
        nb = bl * beta + np.sum(lx * bb, axis=-1)
    is_padding = (t >= num_frames)[:, None]
    beta = np.where(is_padding, beta, nb)
    gb[:, t] = np.where(is_padding, 0, g_b)
    gl[:, t] = np.where(is_padding[..., None], 0, g_l)
  return dist, gb, gl


def viterbi(blank, lexical, num_frames, context, max_expansions=0,
            frame_dependent=True):
  """MaxTropical shortest distance + the one-hot arc gradient.

  Same tie-breaking as autograd through the reference:
    * blank beats lexical (`a >= b`, semirings.py:363) for FrameDependent;
    * fewer expansions win for FrameLabelDependent (first arg-max over the
      stacked `terminated` list, alignments.py:376 + semirings.py:382);
    * inside forward_reduce the first row block wins (contexts.py:226-229);
    * the final state is the first arg-max of alpha_T (lattices.py:496).
  Returns (dist [B], grad_blank [B,T,C], grad_lexical [B,T,C,V],
           alignment_labels [B, T*(k+1)] with TRUE 1-based labels).
  """
  b, t_max, c = blank.shape
  v = lexical.shape[-1]
  n = context.context_size
  num_frames = np.asarray(num_frames).astype(np.int64)
  k = 0 if frame_dependent else max_expansions
  nlev = k + 1
  low = sum(v**i for i in range(0, n - 1))
  vn = v**n
  offset = 1 if n > 0 else 0
  gb = np.zeros_like(blank)
  gl = np.zeros_like(lexical)
  labels = np.zeros((b, t_max, nlev), dtype=np.int64)
  dist = np.zeros((b,), dtype=blank.dtype)

  def reduce_with_arg(w):
    # w: [C, V]; returns (values [C], source flat arc index per dest or -1).
    out = np.full((c,), -np.inf, dtype=w.dtype)
    arg = np.full((c,), -1, dtype=np.int64)
    flat = w.reshape(-1)
    nlow = low * v
    out[offset:offset + nlow] = flat[:nlow]
    arg[offset:offset + nlow] = np.arange(nlow)
    tail = flat[nlow:].reshape(-1, vn)
    kk = np.argmax(tail, axis=0)
    out[offset + nlow:] = tail[kk, np.arange(vn)]
    arg[offset + nlow:] = nlow + kk * vn + np.arange(vn)
    return out, arg

  for bi in range(b):
    alpha = np.full((c,), -np.inf, dtype=blank.dtype)
    alpha[context.start()] = 0
    bps = []
    nf = int(min(max(num_frames[bi], 0), t_max))
    for t in range(nf):
      bl, lx = blank[bi, t], lexical[bi, t]
      if frame_dependent:
        red, arg = reduce_with_arg(alpha[:, None] + lx)
        a = alpha + bl
        take_blank = a >= red
        nxt = np.where(take_blank, a, red)
        bps.append((take_blank, [arg]))
      else:
        terms = [alpha + bl]
        args = []
        last = alpha
        for _ in range(k):
          last, arg = reduce_with_arg(last[:, None] + lx)
          args.append(arg)
          terms.append(last + bl)
        st = np.stack(terms)
        which = np.argmax(st, axis=0)
        nxt = st[which, np.arange(c)]
        bps.append((which, args))
      alpha = nxt
    q = int(np.argmax(alpha))
    dist[bi] = alpha[q]
    for t in range(nf - 1, -1, -1):
      first, args = bps[t]
      if frame_dependent:
        if first[q]:
          gb[bi, t, q] += 1
        else:
          flat = int(args[0][q])
          p, y = flat // v, flat % v
          gl[bi, t, p, y] += 1
          labels[bi, t, 0] = y + 1
          q = p
      else:
        nexp = int(first[q])
        gb[bi, t, q] += 1
        for i in range(nexp - 1, -1, -1):
          flat = int(args[i][q])
          p, y = flat // v, flat % v
          gl[bi, t, p, y] += 1
          labels[bi, t, i] = y + 1
          q = p
  return dist, gb, gl, labels.reshape(b, t_max * nlev)


def gather_string_weights(blank, lexical, labels, context):
  """weight_step_scan, lattices.py:300-342 + :830-845.

  labels [B, U] -> context_states [B, U+1] (walk_states), next labels
  labels ++ [1]; label 0 is read as label 1 (make_safe_classes, :314-315).
  Returns (blank_w [B,T,U+1], lexical_w [B,T,U+1], states, safe_labels).
  """
  labels = np.asarray(labels).astype(np.int64)
  b = labels.shape[0]
  states = context.walk_states(labels)                           # :336
  nxt = np.concatenate([labels, np.ones((b, 1), np.int64)], -1)  # :337-338
  safe = np.where(nxt - 1 < 0, 1, nxt)                           # :314-315
  bi = np.arange(b)[:, None]
  blank_w = blank[bi, :, states].transpose(0, 2, 1)              # [B,T,U+1]
  lexical_w = lexical[bi, :, states, safe - 1].transpose(0, 2, 1)
  return blank_w, lexical_w, states, safe


def string_forward(blank_w, lexical_w, num_frames, num_labels, sr,
                   max_expansions=0, frame_dependent=True):
  """shortest_distance_step_scan + final selection, lattices.py:347-377."""
  b, t_max, u1 = blank_w.shape
  num_frames = np.asarray(num_frames)
  num_labels = np.asarray(num_labels)
  alpha = _init_alpha(b, u1, 0, sr, blank_w.dtype)
  for t in range(t_max):
    nxt = frame_string_forward(alpha, blank_w[:, t], lexical_w[:, t], sr,
                               max_expansions, frame_dependent)
    is_padding = (t >= num_frames)[:, None]                       # :357-358
    alpha = np.where(is_padding, alpha, nxt)
  is_final = num_labels[:, None] == np.arange(u1)[None, :]        # :375
  masked = np.where(is_final, alpha, sr_zero(sr, alpha.dtype.type))
  return sr_sum(sr, masked.astype(alpha.dtype), axis=-1)


def lattice_string_forward(blank, lexical, num_frames, labels, num_labels,
                           context, sr, max_expansions=0, frame_dependent=True):
  """RecognitionLattice._string_forward, lattices.py:250-377."""
  bw, lw, _, _ = gather_string_weights(blank, lexical, labels, context)
  return string_forward(bw, lw, num_frames, num_labels, sr, max_expansions,
                        frame_dependent)


def string_marginals(blank_w, lexical_w, num_frames, num_labels,
                     max_expansions=0, frame_dependent=True):
  """d numerator / d (blank_w, lexical_w) under the Log semiring.

  The reference has no backward for the numerator; it relies on autograd
  through alignments.py:327-329 / :427-432 (broken as shipped).
  This is the equivalent chain forward-backward; it is cross-checked against
  patched reference autograd by tests/golden/make_golden.py.
  Returns (numerator [B], grad_blank_w, grad_lexical_w); utterances whose
  numerator is -inf get zero gradients (safe-gradient rule,
  semirings.py:222-241).
  """
  b, t_max, u1 = blank_w.shape
  dt = blank_w.dtype
  num_frames = np.clip(np.asarray(num_frames).astype(np.int64), 0, t_max)
  num_labels = np.asarray(num_labels).astype(np.int64)
  k = 0 if frame_dependent else max_expansions
  gb = np.zeros_like(blank_w)
  gl = np.zeros_like(lexical_w)
  num = np.full((b,), -np.inf, dtype=dt)
  ninf = dt.type(-np.inf)

  def shift_up(x):      # out[u] = x[u+1], out[U] = -inf
    return np.concatenate([x[1:], np.full((1,), ninf, dt)])

  def shift_dn(x):
    return np.concatenate([np.full((1,), ninf, dt), x[:-1]])

  with np.errstate(all='ignore'):
    for bi in range(b):
      nf = int(num_frames[bi])
      alphas = np.full((nf + 1, u1), ninf, dtype=dt)
      alphas[0, 0] = 0
      for t in range(nf):
        alphas[t + 1] = frame_string_forward(
            alphas[t], blank_w[bi, t], lexical_w[bi, t], LOG, max_expansions,
            frame_dependent)
      nl = int(num_labels[bi])
      if nl < 0 or nl >= u1:
        continue
      z = alphas[nf, nl]
      num[bi] = z
      if not np.isfinite(z):
        continue
      beta = np.full((u1,), ninf, dtype=dt)
      beta[nl] = 0
      for t in range(nf - 1, -1, -1):
        bl, lx, a = blank_w[bi, t], lexical_w[bi, t], alphas[t]
        if frame_dependent:
          lex_beta = lx + shift_up(beta)
          gb[bi, t] = np.exp(a + bl + beta - z)
          gl[bi, t] = np.exp(a + lex_beta - z)
          beta = _logaddexp(bl + beta, lex_beta)
        else:
          lasts = [a]
          for _ in range(k):
            lasts.append(shift_dn(lasts[-1] + lx))
          gb[bi, t] = sum(np.exp(l + bl + beta - z) for l in lasts)
          nb = bl + beta
          acc = np.zeros((u1,), dtype=dt)
          for j in range(k - 1, -1, -1):
            lex_beta = lx + shift_up(nb)
            acc = acc + np.exp(lasts[j] + lex_beta - z)
            nb = _logaddexp(bl + beta, lex_beta)
          gl[bi, t] = acc
          beta = nb
  return num, gb, gl


def scatter_string_grads(gbw, glw, states, safe_labels, shape_blank,
                         shape_lexical):
  """Transpose of gather_string_weights (adds repeated (state,label) hits)."""
  gb = np.zeros(shape_blank, dtype=gbw.dtype)
  gl = np.zeros(shape_lexical, dtype=glw.dtype)
  b, t_max, u1 = gbw.shape
  for bi in range(b):
    for u in range(u1):
      s, y = int(states[bi, u]), int(safe_labels[bi, u]) - 1
      gb[bi, :, s] += gbw[bi, :, u]
      gl[bi, :, s, y] += glw[bi, :, u]
  return gb, gl


def lattice_loss_and_grads(blank, lexical, num_frames, labels, num_labels,
                           context, max_expansions=0, frame_dependent=True):
  """RecognitionLattice.forward, lattices.py:131-183: loss = logZ - numerator,
  with d sum(loss) / d (blank, lexical)."""
  log_z, gb, gl = lattice_marginals(blank, lexical, num_frames, context,
                                    max_expansions, frame_dependent)
  bw, lw, states, safe = gather_string_weights(blank, lexical, labels, context)
  num, gbw, glw = string_marginals(bw, lw, num_frames, num_labels,
                                   max_expansions, frame_dependent)
  sb, sl = scatter_string_grads(gbw, glw, states, safe, blank.shape,
                                lexical.shape)
  with np.errstate(all='ignore'):
    loss = log_z - num
  return loss, gb - sb, gl - sl


# ---------------------------------------------------------------------------
# JointWeightFn (reference: last_torch/weight_fns.py:194-227)
# ---------------------------------------------------------------------------

def joint_weights(cache, frames, w_ctx, w_frame, w_blank, b_blank, w_vocab,
                  b_vocab):
  """weight_fns.py:208-227 with explicit (deterministic) parameters.

  cache [C,E]; frames [..., D]; w_ctx [H,E]; w_frame [H,D]; w_blank [H];
  b_blank scalar; w_vocab [V,H]; b_vocab [V].
  Returns blank [..., C], lexical [..., C, V].
  """
  pc = cache @ w_ctx.T                                  # [C,H]   :215
  pf = frames @ w_frame.T                               # [...,H] :216
  joint = np.tanh(pc + pf[..., None, :])                # [...,C,H] :218
  blank = joint @ w_blank + b_blank                     # :223-225
  lexical = joint @ w_vocab.T + b_vocab                 # :226
  return blank.astype(frames.dtype), lexical.astype(frames.dtype)
