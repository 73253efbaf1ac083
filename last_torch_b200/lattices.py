"""Recognition lattice (drop-in for last_torch.RecognitionLattice).

Same constructor and methods as the reference
(/root/reference/last_torch/lattices.py:35-799): forward (the GNAT loss),
shortest_path, build_cache and the unit-tested privates _forward,
_string_forward, _forward_backward.  What changed underneath:

  * the weight function is evaluated ONCE over all frames
    (WeightFn.all_frames) instead of once per Python loop iteration;
  * the T-frame semiring recursion (lattices.py:436-462, :856-892) is a single
    persistent CUDA kernel (ops.LatticeForward / K1);
  * gradients come from a hand-written backward (beta) recursion that writes
    arc posteriors as the weight gradient (K2) -- the reference's
    _forward_backward has no working backward;
  * shortest_path follows back-pointers (K5) instead of differentiating through
    the unrolled loop, and reports the true 1-based labels.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from typing import Generic, Optional, TypeVar

import torch
from torch import nn

from . import _native as N
from . import alignments
from . import contexts
from . import ops
from . import semirings
from . import weight_fns

T = TypeVar('T')


class RecognitionLattice(nn.Module, Generic[T]):
  """Recognition lattice in the GNAT formulation: context dependency x
  alignment lattice x weight function (lattices.py:35-116)."""

  def __init__(self, context: contexts.ContextDependency,
               alignment: alignments.TimeSyncAlignmentLattice,
               weight_fn_cacher_factory: Callable[[contexts.ContextDependency],
                                                  weight_fns.WeightFnCacher[T]],
               weight_fn_factory: Callable[[contexts.ContextDependency],
                                           weight_fns.WeightFn[T]]):
    super().__init__()
    self.context = context
    self.alignment = alignment
    self.weight_fn_cacher_factory = weight_fn_cacher_factory
    self.weight_fn_factory = weight_fn_factory
    self.weight_fn_cacher = self.weight_fn_cacher_factory(self.context)
    self.weight_fn = self.weight_fn_factory(self.context)
    # kernel dispatch options (see include/last_lattice.h); 0 = automatic
    self.kernel_flags = 0
    # JointWeightFn + FullNGram: hand the arc posteriors to the joint network's backward as
    # split rows (ops.JointLatticeLoss); False keeps them in float32
    self.split_grad_handover = ops.SPLIT_GRAD_DEFAULT
    # Numerator from the weight function evaluated on the U+1 states of the label string only
    # (lattices.py:300-313) instead of gathered out of the dense [B,T,C,V] weights.  None =
    # automatic: when the loss needs no denominator (LocallyNormalizedWeightFn).
    self.gathered_numerator = None
    # Inference without materialising the arc weights: JointWeightFn fused into the forward
    # recursion (ops.joint_lattice_forward_fused; bigram, vocab <= 64, hidden in {32, 64, 128}).
    # Off by default: it saves the O(B*T*C*V) logits but is slower than the tensor-core joint
    # kernel + K1.  Used by shortest_path and by _forward when no
    # gradient is required.
    self.fused_inference = False
    # shortest_path: reproduce the label encoding of the reference as shipped (see there)
    self.reference_compat = False
    # Range-check the reference labels (0 <= label <= vocab_size for the first num_labels
    # positions) and raise ValueError like the reference's one_hot (lattices.py:317-324) -- one
    # 4-byte device-to-host read per call.  False skips the read; the kernels then treat an
    # out-of-range label as epsilon (memory-safe, but the loss of that utterance is not defined).
    self.validate_labels = True

  def build_cache(self) -> T:
    """Builds the weight function cache (lattices.py:118-129)."""
    return self.weight_fn_cacher()

  # -- helpers ---------------------------------------------------------------

  def _geometry(self):
    """(V, n, k) for the kernels; raises for anything they do not implement."""
    if not isinstance(self.alignment, alignments.TimeSyncAlignmentLattice):
      raise NotImplementedError(f'unsupported alignment {type(self.alignment).__name__}')
    k = self.alignment.kernel_max_expansions()
    if isinstance(self.context, contexts.NextStateTable):
      # table-driven kernels (csrc/lattice_table.cu); n is meaningless for a generic DFA
      return self.context.shape()[1], None, k
    if not isinstance(self.context, contexts.FullNGram):
      raise NotImplementedError(
          'the lattice kernels implement contexts.FullNGram and contexts.NextStateTable; got '
          f'{type(self.context).__name__} (no fallback path exists)')
    return self.context.vocab_size, self.context.context_size, k

  def _is_table(self) -> bool:
    return isinstance(self.context, contexts.NextStateTable)

  @staticmethod
  def _check_frames(frames, num_frames):
    batch_dims = tuple(num_frames.shape)
    if tuple(frames.shape[:-2]) != batch_dims:
      raise ValueError('frames and num_frames have different batch_dims: '
                       f'{tuple(frames.shape[:-2])} vs {batch_dims}')
    return batch_dims

  @staticmethod
  def _check_labels(labels, num_labels, batch_dims):
    if tuple(labels.shape[:-1]) != batch_dims:
      raise ValueError('labels and num_frames have different batch_dims: '
                       f'{tuple(labels.shape[:-1])} vs {batch_dims}')
    if tuple(num_labels.shape) != batch_dims:
      raise ValueError('num_labels and num_frames have different batch_dims: '
                       f'{tuple(num_labels.shape)} vs {batch_dims}')

  def _arc_weights(self, cache, frames, batch_dims):
    """Dense arc weights of all frames, flattened to one batch axis:
    blank [B,T,C], lexical [B,T,C,V] (fp32, contiguous, CUDA)."""
    blank, lexical = self.weight_fn.all_frames(cache, frames)
    c, v = self.context.shape()
    t = frames.shape[-2]
    blank = blank.reshape(-1, t, c)
    lexical = lexical.reshape(-1, t, c, v)
    if not blank.is_cuda:
      raise RuntimeError(
          f'arc weights live on {blank.device}; last_torch_b200 runs on CUDA (sm_100a) only '
          'and has no CPU fallback -- move frames / weight function to a CUDA device')
    if blank.dtype != torch.float32 or lexical.dtype != torch.float32:
      raise TypeError('the lattice kernels compute in float32; weight function returned '
                      f'{blank.dtype} / {lexical.dtype}')
    return blank.contiguous(), lexical.contiguous()

  def _string_indices(self, labels, num_labels, device):
    """context states along the label string and the label leaving each of
    them (lattices.py:336-338; label 0 is read as label 1, :314-315).

    Positions u >= num_labels[b] are padding: whatever they hold (-1, vocab_size + 1, ...) is
    read as epsilon -- they cannot influence alpha[num_labels] (alignments.py:327-329 only moves
    forward along the string).  A label outside [0, vocab_size] BEFORE num_labels is an error:
    the reference fails in one_hot (lattices.py:322); here the kernel counts such labels and,
    with validate_labels, a ValueError is raised."""
    labels = labels.reshape(-1, labels.shape[-1]).to(device=device, dtype=torch.int32)
    nl = ops._as_i32(num_labels.reshape(-1), device)
    v = self.context.shape()[1]
    if self._is_table():
      # generic DFA: integer gathers through the table (contexts.py:109-146)
      pos = torch.arange(labels.shape[1], device=device, dtype=torch.int32)[None, :]
      real = pos < nl[:, None]
      bad = (real & ((labels < 0) | (labels > v))).sum().to(torch.int32).reshape(1)
      labels = torch.where(real & (labels >= 0) & (labels <= v), labels,
                           torch.zeros_like(labels))
      states = self.context.walk_states(labels.long()).to(torch.int32).contiguous()
      ones = torch.ones_like(labels[:, :1])
      next_labels = torch.cat([torch.where(labels == 0, torch.ones_like(labels), labels), ones],
                              dim=1).contiguous()
    else:
      states, next_labels, bad = ops.walk_states(labels.contiguous(), nl,
                                                 self.context.vocab_size,
                                                 self.context.context_size)
    if self.validate_labels and int(bad) != 0:
      raise ValueError(f'labels must be in [0, vocab_size={v}] for the first num_labels '
                       f'positions of every utterance; found {int(bad)} label(s) outside')
    return states, next_labels

  # -- public API --------------------------------------------------------------

  def forward(self, frames: torch.Tensor, num_frames: torch.Tensor, labels: torch.Tensor,
              num_labels: torch.Tensor, cache: Optional[T] = None) -> torch.Tensor:
    """Negative sequence log-probability, [batch_dims...] (lattices.py:131-183)."""
    batch_dims = self._check_frames(frames, num_frames)
    self._check_labels(labels, num_labels, batch_dims)
    if cache is None:
      cache = self.weight_fn_cacher()
    if isinstance(self.weight_fn, weight_fns.LocallyNormalizedWeightFn):
      return -self._string_forward(cache=cache, frames=frames, num_frames=num_frames,
                                   labels=labels, num_labels=num_labels,
                                   semiring=semirings.Log)
    v, n, k = self._geometry()
    if type(self.weight_fn) is weight_fns.JointWeightFn and not self._is_table():
      return self._joint_loss(cache, frames, num_frames, labels, num_labels, batch_dims, v, n, k)
    blank, lexical = self._arc_weights(cache, frames, batch_dims)
    dev = blank.device
    states, next_labels = self._string_indices(labels, num_labels, dev)
    if self._is_table():
      nf = ops._as_i32(num_frames.reshape(-1), dev)
      log_z, _ = ops.TableLatticeForward.apply(blank, lexical, nf, self.context, N.LOG, k)
      num = ops.StringForward.apply(blank, lexical, nf, states, next_labels,
                                    ops._as_i32(num_labels.reshape(-1), dev), N.LOG, v, k)
      return (log_z - num).reshape(batch_dims)
    loss, _, _ = ops.LatticeLoss.apply(
        blank, lexical, ops._as_i32(num_frames.reshape(-1), dev), states, next_labels,
        ops._as_i32(num_labels.reshape(-1), dev), v, n, k, self.kernel_flags)
    return loss.reshape(batch_dims)

  def _joint_loss(self, cache, frames, num_frames, labels, num_labels, batch_dims, v, n, k):
    """JointWeightFn inside a FullNGram lattice: the two input projections are library GEMMs,
    everything after them -- tanh joint, vocabulary projection, denominator, numerator and the
    whole backward -- is ONE autograd node (ops.JointLatticeLoss); the dense arc weights and
    their gradients never become autograd tensors."""
    from . import joint
    if not frames.is_cuda:
      raise RuntimeError(
          f'frames live on {frames.device}; last_torch_b200 runs on CUDA (sm_100a) only '
          'and has no CPU fallback')
    fn = self.weight_fn
    fn._check_lazy(cache, frames)
    dev = frames.device
    t = frames.shape[-2]
    b = 1
    for d in batch_dims:
      b *= d
    states, next_labels = self._string_indices(labels, num_labels, dev)
    proj_ctx, proj_frame = joint.joint_projections(fn, cache, frames)
    loss, _, _ = ops.JointLatticeLoss.apply(
        proj_ctx, proj_frame, fn.joint_projection_to_blank.weight,
        fn.joint_projection_to_blank.bias.reshape(()), fn.joint_projection_to_vocab.weight,
        fn.joint_projection_to_vocab.bias, ops._as_i32(num_frames.reshape(-1), dev), states,
        next_labels, ops._as_i32(num_labels.reshape(-1), dev), b, t, v, n, k, self.kernel_flags,
        self.split_grad_handover)
    return loss.reshape(batch_dims)

  def shortest_path(self, frames: torch.Tensor, num_frames: torch.Tensor,
                    cache: Optional[T] = None, reference_compat: Optional[bool] = None):
    """Highest scoring path (lattices.py:185-247).

    Returns (alignment_labels [batch_dims..., T * num_alignment_states],
    num_alignment_labels [batch_dims...], path_weights [batch_dims...]).
    Labels are the TRUE labels: 0 = blank, 1..vocab_size lexical.

    reference_compat=True (default: self.reference_compat) reproduces the labels the reference
    prints as shipped, for code and tests that depend on them:
      * lattices.py:242-244 takes argmax over the V-wide mask without the `1 +`:
        lexical label y is reported as y - 1, so label 1 is indistinguishable from blank;
      * FrameDependent with one batch dimension: lattices.py:875-879 hands batch element 0's
        mask to every utterance, so row 0 holds, per frame, the most frequent label
        (first on ties) over the paths of ALL utterances and every other row is zero
        (tests/lattices_test.py:238-242 pins exactly this).
    path_weights and num_alignment_labels are unaffected.
    """
    batch_dims = self._check_frames(frames, num_frames)
    if cache is None:
      cache = self.weight_fn_cacher()
    v, n, k = self._geometry()
    if self._fused_applicable(N.MAXTROPICAL, frames, cache):
      path_weights, _, labels, _ = self._fused_forward(
          N.MAXTROPICAL, cache, frames, num_frames, batch_dims, False, True)
      dev = frames.device
      if reference_compat is None:
        reference_compat = self.reference_compat
      if reference_compat:
        labels = self._reference_labels(labels, num_frames.to(dev).reshape(-1), v, len(batch_dims))
      return (labels.to(torch.int64).reshape(*batch_dims, -1),
              self.alignment.num_states() * num_frames.to(dev), path_weights.reshape(batch_dims))
    with torch.no_grad():
      blank, lexical = self._arc_weights(cache, frames, batch_dims)
      dev = blank.device
      if self._is_table():
        labels, _, path_weights = ops.table_viterbi_path(
            blank, lexical, ops._as_i32(num_frames.reshape(-1), dev), self.context, k)
      else:
        labels, _, path_weights = ops.viterbi_path(
            blank, lexical, ops._as_i32(num_frames.reshape(-1), dev), v, n, k,
            self.kernel_flags)
    num_alignment_states = self.alignment.num_states()
    if reference_compat is None:
      reference_compat = self.reference_compat
    if reference_compat:
      labels = self._reference_labels(labels, num_frames.to(dev).reshape(-1), v, len(batch_dims))
    alignment_labels = labels.to(torch.int64).reshape(*batch_dims, -1)
    num_alignment_labels = num_alignment_states * num_frames.to(dev)
    return alignment_labels, num_alignment_labels, path_weights.reshape(batch_dims)

  def _fused_applicable(self, sr, frames, cache) -> bool:
    if not self.fused_inference or self._is_table() or not frames.is_cuda:
      return False
    if type(self.weight_fn) is not weight_fns.JointWeightFn:
      return False
    if not isinstance(self.context, contexts.FullNGram):
      return False
    k = self.alignment.kernel_max_expansions()
    return bool(N.lib().lt_joint_lattice_fused_supported(
        sr, self.context.vocab_size, self.context.context_size, k, self.weight_fn.hidden_size))

  def _fused_forward(self, sr, cache, frames, num_frames, batch_dims, want_alphas, want_path):
    from . import joint
    fn = self.weight_fn
    fn._check_lazy(cache, frames)
    t = frames.shape[-2]
    with torch.no_grad():
      proj_ctx, proj_frame = joint.joint_projections(fn, cache, frames)
      return ops.joint_lattice_forward_fused(
          sr, self.context.vocab_size, proj_ctx, proj_frame.reshape(-1, t, proj_frame.shape[-1]),
          fn.joint_projection_to_blank.weight, fn.joint_projection_to_blank.bias,
          fn.joint_projection_to_vocab.weight, fn.joint_projection_to_vocab.bias,
          ops._as_i32(num_frames.reshape(-1), frames.device), want_alphas, want_path)

  def _reference_labels(self, labels, num_frames, v, num_batch_dims):
    """labels [B, T, N] true labels -> what the reference as shipped prints (D4, D5)."""
    if isinstance(self.alignment, alignments.FrameDependent) and num_batch_dims == 1:
      b, t, _ = labels.shape
      real = torch.arange(t, device=labels.device)[None, :] < num_frames[:, None]
      lab = labels[..., 0].long()
      taken = real & (lab > 0)
      counts = torch.zeros([t, v], dtype=torch.int64, device=labels.device)
      tt = torch.arange(t, device=labels.device)[None, :].expand(b, t)
      counts.index_put_((tt[taken], lab[taken] - 1), torch.ones_like(lab[taken]), accumulate=True)
      row0 = torch.where(counts.sum(-1) > 0, counts.argmax(-1), torch.zeros_like(counts[:, 0]))
      out = torch.zeros_like(labels)
      out[0, :, 0] = row0.to(labels.dtype)
      return out
    return torch.clamp(labels - 1, min=0)

  # -- "private" methods that the reference's tests call directly -------------------

  def _string_forward(self, cache: T, frames: torch.Tensor, num_frames: torch.Tensor,
                      labels: torch.Tensor, num_labels: torch.Tensor,
                      semiring: semirings.Semiring[torch.Tensor]) -> torch.Tensor:
    """Shortest distance on the lattice intersected with the label string
    (lattices.py:250-377)."""
    batch_dims = self._check_frames(frames, num_frames)
    self._check_labels(labels, num_labels, batch_dims)
    sr = semirings.kernel_id(semiring)
    v, n, k = self._geometry()
    gathered = self.gathered_numerator
    if gathered is None:
      gathered = isinstance(self.weight_fn, weight_fns.LocallyNormalizedWeightFn)
    if gathered and len(batch_dims) == 1 and frames.is_cuda:
      return self._string_forward_gathered(cache, frames, num_frames, labels, num_labels, sr, k)
    blank, lexical = self._arc_weights(cache, frames, batch_dims)
    dev = blank.device
    states, next_labels = self._string_indices(labels, num_labels, dev)
    dist = ops.StringForward.apply(
        blank, lexical, ops._as_i32(num_frames.reshape(-1), dev), states, next_labels,
        ops._as_i32(num_labels.reshape(-1), dev), sr, v, k)
    return dist.reshape(batch_dims)

  def _string_forward_gathered(self, cache, frames, num_frames, labels, num_labels, sr, k):
    """_string_forward with the reference's own evaluation order (lattices.py:300-342): the
    weight function sees only the context states along the label string, [B,T,U+1,V] instead of
    [B,T,C,V] -- (U+1)/C of the joint network's work and memory (121 / 257 at configs[3], 121 /
    4161 for a trigram context).  With LocallyNormalizedWeightFn the loss is this alone
    (lattices.py:178-179), so the dense weights are never materialised."""
    dev = frames.device
    states, next_labels = self._string_indices(labels, num_labels, dev)
    bw, lw_all = self.weight_fn.string_frames(cache, frames, states)        # [B,T,U1], [B,T,U1,V]
    b, t, u1 = bw.shape
    pick = (next_labels.to(torch.int64) - 1)[:, None, :, None].expand(b, t, u1, 1)
    lw = torch.gather(lw_all, 3, pick).squeeze(3)
    return ops.StringChainForward.apply(
        bw.contiguous().float(), lw.contiguous().float(), ops._as_i32(num_frames.reshape(-1), dev),
        ops._as_i32(num_labels.reshape(-1), dev), sr, k)

  def _forward(self, cache: T, frames: torch.Tensor, num_frames: torch.Tensor,
               semiring: semirings.Semiring[torch.Tensor],
               blank_mask: Optional[Sequence[torch.Tensor]] = None,
               lexical_mask: Optional[Sequence[torch.Tensor]] = None):
    """Shortest distance on the recognition lattice (lattices.py:379-496).

    Returns (shortest_distance [batch_dims...],
             alpha_0_to_T_minus_1 [batch_dims..., T, num_context_states]).
    Zero-valued masks may be passed to differentiate w.r.t. arc weights
    (lattices.py:390-396); they are added to the dense weights.
    """
    batch_dims = self._check_frames(frames, num_frames)
    num_align = self.alignment.num_states()
    if blank_mask is not None and len(blank_mask) != num_align:
      raise ValueError('The length of blank_mask should be equal to '
                       f'{num_align} (the number of alignment states), '
                       f'but is {len(blank_mask)}')
    if lexical_mask is not None and len(lexical_mask) != num_align:
      raise ValueError('The length of lexical_mask should be equal to '
                       f'{num_align} (the number of alignment states), '
                       f'but is {len(lexical_mask)}')
    sr = semirings.kernel_id(semiring)
    v, n, k = self._geometry()
    if (blank_mask is None and lexical_mask is None and sr != N.REAL and
        not (torch.is_grad_enabled() and (frames.requires_grad or any(
            p.requires_grad for p in self.parameters()))) and
        self._fused_applicable(sr, frames, cache)):
      dist, alphas, _, _ = self._fused_forward(sr, cache, frames, num_frames, batch_dims, True,
                                               False)
      return dist.reshape(batch_dims), alphas.reshape(*batch_dims, frames.shape[-2], -1)
    blank, lexical = self._arc_weights(cache, frames, batch_dims)
    c = blank.shape[-1]
    t = blank.shape[1]
    if (blank_mask is not None or lexical_mask is not None) and num_align != 1:
      # lattices.py:447-453: one set of weights per alignment state, blank[i] + blank_mask[i] /
      # lexical[i] + lexical_mask[i]; the generic kernels take them as [B,T,k+1,C(,V)]
      if self._is_table():
        raise NotImplementedError('per-alignment-state masks need a contexts.FullNGram context')
      full_b, full_l = (*batch_dims, t, c), (*batch_dims, t, c, v)
      bl = [blank if blank_mask is None else
            blank + torch.broadcast_to(blank_mask[i], full_b).reshape(-1, t, c)
            for i in range(num_align)]
      lx = [lexical if lexical_mask is None else
            lexical + torch.broadcast_to(lexical_mask[i], full_l).reshape(-1, t, c, v)
            for i in range(num_align)]
      dist, alphas = ops.LatticeForwardLevels.apply(
          torch.stack(bl, dim=2).contiguous(), torch.stack(lx, dim=2).contiguous(),
          ops._as_i32(num_frames.reshape(-1), blank.device), sr, v, n, k, self.kernel_flags)
      return dist.reshape(batch_dims), alphas.reshape(*batch_dims, t, c)
    if blank_mask is not None or lexical_mask is not None:
      if blank_mask is not None:
        blank = blank + torch.broadcast_to(blank_mask[0], (*batch_dims, t, c)).reshape(-1, t, c)
      if lexical_mask is not None:
        lexical = lexical + torch.broadcast_to(
            lexical_mask[0], (*batch_dims, t, c, v)).reshape(-1, t, c, v)
    dev = blank.device
    if self._is_table():
      dist, alphas = ops.TableLatticeForward.apply(
          blank, lexical, ops._as_i32(num_frames.reshape(-1), dev), self.context, sr, k)
    else:
      dist, alphas = ops.LatticeForward.apply(
          blank, lexical, ops._as_i32(num_frames.reshape(-1), dev), sr, v, n, k,
          self.kernel_flags)
    return dist.reshape(batch_dims), alphas.reshape(*batch_dims, t, c)

  def expectation(self, frames: torch.Tensor, num_frames: torch.Tensor,
                  value_blank: Optional[torch.Tensor] = None,
                  value_lexical: Optional[torch.Tensor] = None, cache: Optional[T] = None):
    """(log_z, E) with E[b] = the expected total value of a lattice path under the Log-semiring
    path distribution, for an additive arc value function value_blank [batch_dims..., T, C] /
    value_lexical [batch_dims..., T, C, V] -- what running the recursion in the expectation
    semiring (semirings.Expectation / LogLogExpectation, semirings.py:404-484) yields; the
    reference cannot do that (`torch.where` on tuples in lattices.py:804-806).
    Without values the arc weights themselves are used (E = expected path weight).
    No gradients.  E is float64."""
    batch_dims = self._check_frames(frames, num_frames)
    if self._is_table():
      raise NotImplementedError('expectation needs a contexts.FullNGram context')
    if cache is None:
      cache = self.build_cache()
    with torch.no_grad():
      blank, lexical = self._arc_weights(cache, frames, batch_dims)
      v, n, k = self._geometry()
      t, c = blank.shape[1], blank.shape[2]
      if value_blank is not None:
        value_blank = torch.broadcast_to(value_blank, (*batch_dims, t, c)).reshape(-1, t, c)
      if value_lexical is not None:
        value_lexical = torch.broadcast_to(value_lexical, (*batch_dims, t, c, v)).reshape(-1, t, c, v)
      log_z, expect = ops.lattice_expectation(
          blank, lexical, ops._as_i32(num_frames.reshape(-1), blank.device), v, n, k, value_blank,
          value_lexical, self.kernel_flags)
    return log_z.reshape(batch_dims), expect.reshape(batch_dims)

  def entropy(self, frames: torch.Tensor, num_frames: torch.Tensor, cache: Optional[T] = None):
    """Entropy (nats) of the lattice's distribution over alignment paths,
    H = log Z - E[path weight]: the textbook use of the expectation semiring
    (semirings_test.py:305-324), in two passes over the arc weights."""
    log_z, expect = self.expectation(frames, num_frames, cache=cache)
    return (log_z.double() - expect).float()

  def _forward_backward(self, cache: T, frames: torch.Tensor, num_frames: torch.Tensor):
    """Log-semiring shortest distance whose gradient is computed by the
    backward algorithm (lattices.py:498-642).  Returns (log_z, alphas)."""
    return self._forward(cache=cache, frames=frames, num_frames=num_frames,
                         semiring=semirings.Log)
