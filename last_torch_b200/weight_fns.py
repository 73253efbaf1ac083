"""Weight functions (drop-in for last_torch.weight_fns).

Same class names and constructor arguments as the reference
(/root/reference/last_torch/weight_fns.py).  Differences that are deliberate
fixes of reference defects (documented in DESIGN.md):

  * JointWeightFn owns its four projections as parameters registered in the
    constructor (the two input projections as torch lazy modules when their
    widths are not given) instead of building fresh random nn.Linear layers on
    every call.
  * SharedEmbCacher returns the [num_context_states, embedding_size] table.

Every WeightFn additionally exposes `all_frames(cache, frames)`, which
evaluates the arc weights of ALL T frames in one call; RecognitionLattice uses
it to materialise `blank [B,T,C]` / `lexical [B,T,C,V]` once and hands them to
the persistent lattice kernels (a WeightFn is frame-independent by contract,
weight_fns.py:57-82).
"""

from __future__ import annotations

import abc
from typing import Callable, Generic, Optional, TypeVar

import torch
from torch import nn
from torch.nn import functional as F

T = TypeVar('T')


class WeightFn(nn.Module, Generic[T], abc.ABC):
  """Interface (weight_fns.py:42-83)."""

  @abc.abstractmethod
  def forward(self, cache: T, frame: torch.Tensor,
              state: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """frame [batch_dims..., feature_size] -> (blank [..., C], lexical [..., C, V]),
    or ([...], [..., V]) for the given `state`."""
    raise NotImplementedError

  def all_frames(self, cache: T, frames: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """frames [batch_dims..., T, feature_size] ->
    (blank [batch_dims..., T, C], lexical [batch_dims..., T, C, V]).

    Default: one `forward` call per frame (always valid); subclasses override
    it with a single batched evaluation.
    """
    outs = [self(cache, frames[..., t, :]) for t in range(frames.shape[-2])]
    nb = frames.ndim - 2
    return (torch.stack([o[0] for o in outs], dim=nb),
            torch.stack([o[1] for o in outs], dim=nb))


  def string_frames(self, cache: T, frames: torch.Tensor,
                    states: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Arc weights of all frames on the context states of a label string ONLY
    (lattices.py:300-313, weight_step_scan :830-845): frames [B, T, feature_size], states
    [B, U1] -> (blank [B, T, U1], lexical [B, T, U1, V]).  Default: what the reference does --
    for every string position one `forward(cache, frame, state)` vectorised over the frames with
    torch.vmap (falling back to a loop over frames for weight functions vmap cannot trace);
    subclasses override it with batched evaluations."""
    t = frames.shape[1]
    blanks, lexicals = [], []
    for u in range(states.shape[1]):
      state = states[:, u]
      try:
        bl, lx = torch.vmap(lambda fr: self(cache, fr, state), in_dims=1, out_dims=1)(frames)
      except Exception:      # data-dependent control flow, custom autograd functions, ...
        outs = [self(cache, frames[:, i], state) for i in range(t)]
        bl = torch.stack([o[0] for o in outs], dim=1)
        lx = torch.stack([o[1] for o in outs], dim=1)
      blanks.append(bl)
      lexicals.append(lx)
    return torch.stack(blanks, dim=2), torch.stack(lexicals, dim=2)


class WeightFnCacher(nn.Module, Generic[T], abc.ABC):
  """Interface (weight_fns.py:86-96)."""

  @abc.abstractmethod
  def forward(self) -> T:
    """Builds the cached data."""


class _LocalNormalize(torch.autograd.Function):
  """Fused row-wise normaliser (csrc/normalize.cu) for CUDA fp32 weights."""

  @staticmethod
  def forward(ctx, blank, lexical, mode):
    from . import _native as N
    blank = N.require_cuda(blank, 'blank')
    lexical = N.require_cuda(lexical, 'lexical')
    v = lexical.shape[-1]
    m = blank.numel()
    ob, ol = torch.empty_like(blank), torch.empty_like(lexical)
    with torch.cuda.device(blank.device):
      N.check(N.lib().lt_local_normalize_forward(
          mode, N.ptr(blank), N.ptr(lexical), m, v, N.ptr(ob), N.ptr(ol),
          N.stream_ptr(blank.device)), 'lt_local_normalize_forward')
    ctx.save_for_backward(blank, lexical)
    ctx.mode = mode
    return ob, ol

  @staticmethod
  def backward(ctx, gb, gl):
    from . import _native as N
    blank, lexical = ctx.saved_tensors
    gb = N.require_cuda(gb, 'grad_blank')
    gl = N.require_cuda(gl, 'grad_lexical')
    db, dl = torch.empty_like(blank), torch.empty_like(lexical)
    with torch.cuda.device(blank.device):
      N.check(N.lib().lt_local_normalize_backward(
          ctx.mode, N.ptr(blank), N.ptr(lexical), N.ptr(gb), N.ptr(gl), blank.numel(),
          lexical.shape[-1], N.ptr(db), N.ptr(dl), N.stream_ptr(blank.device)),
          'lt_local_normalize_backward')
    return db, dl, None


def _on_kernel_path(blank, lexical) -> bool:
  return (blank.is_cuda and lexical.is_cuda and blank.dtype == torch.float32 and
          lexical.dtype == torch.float32 and lexical.shape[:-1] == blank.shape)


def hat_normalize(blank: torch.Tensor, lexical: torch.Tensor):
  """HAT local normalisation (weight_fns.py:99-117).  CUDA fp32 weights go through
  the fused kernel; host tensors (the reference's unit-test use) keep the formula."""
  if _on_kernel_path(blank, lexical):
    return _LocalNormalize.apply(blank, lexical, 0)
  z = F.softplus(blank)
  return blank - z, F.log_softmax(lexical, dim=-1) - z.unsqueeze(-1)


def log_softmax_normalize(blank: torch.Tensor, lexical: torch.Tensor):
  """Joint log-softmax over blank ++ lexical (weight_fns.py:120-136)."""
  if _on_kernel_path(blank, lexical):
    return _LocalNormalize.apply(blank, lexical, 1)
  all_weights = F.log_softmax(torch.cat([blank.unsqueeze(-1), lexical], dim=-1), dim=-1)
  return all_weights[..., 0], all_weights[..., 1:]


class LocallyNormalizedWeightFn(WeightFn[T]):
  """Wrapper that makes any weight function locally normalised
  (weight_fns.py:139-171); RecognitionLattice.forward then skips the
  denominator (lattices.py:178-179)."""

  def __init__(self, weight_fn: WeightFn[T],
               normalize: Callable[[torch.Tensor, torch.Tensor],
                                   tuple[torch.Tensor, torch.Tensor]] = hat_normalize,
               *args, **kwargs) -> None:
    super().__init__(*args, **kwargs)
    self.weight_fn = weight_fn
    self.normalize = normalize

  def forward(self, cache, frame, state=None):
    return self.normalize(*self.weight_fn(cache, frame, state))

  def all_frames(self, cache, frames):
    return self.normalize(*self.weight_fn.all_frames(cache, frames))

  def string_frames(self, cache, frames, states):
    return self.normalize(*self.weight_fn.string_frames(cache, frames, states))


class JointWeightFn(WeightFn[torch.Tensor]):
  r"""tanh(W_c emb[c] + W_f frame) -> Linear(H, 1), Linear(H, V)
  (weight_fns.py:174-227).

  Works with any cacher that yields a [num_context_states, embedding_size]
  table.  Parameters (names follow the local variables of the reference body):
    context_projection   Linear(E, H, bias=False)   weight_fns.py:208-209
    blank_projection     Linear(D, H, bias=False)   weight_fns.py:210-211 (projects the frame)
    joint_projection_to_blank  Linear(H, 1)         weight_fns.py:220
    joint_projection_to_vocab  Linear(H, V)         weight_fns.py:221
  """

  def __init__(self, vocab_size: int, hidden_size: int, device: Optional[str] = 'cpu',
               *args, embedding_size: Optional[int] = None, feature_size: Optional[int] = None,
               **kwargs) -> None:
    super().__init__(*args, **kwargs)
    self.vocab_size = vocab_size
    self.hidden_size = hidden_size
    self.device = device
    dev = torch.device(device or 'cpu')
    h, v = hidden_size, vocab_size
    # All four projections are REGISTERED in the constructor, so an optimizer, a DDP wrapper or
    # load_state_dict built before the first call sees them.  The reference signature does not
    # carry the two input widths (weight_fns.py:187-192): without them the input projections are
    # torch lazy modules -- their (uninitialised) parameters exist from the start and are
    # materialised in place by the first call, by materialize(), or by load_state_dict.
    self.context_projection = (nn.Linear(embedding_size, h, bias=False, device=dev)
                               if embedding_size is not None else
                               nn.LazyLinear(h, bias=False, device=dev))
    self.blank_projection = (nn.Linear(feature_size, h, bias=False, device=dev)
                             if feature_size is not None else
                             nn.LazyLinear(h, bias=False, device=dev))
    self.joint_projection_to_blank = nn.Linear(h, 1, device=dev)
    self.joint_projection_to_vocab = nn.Linear(h, v, device=dev)

  def materialize(self, embedding_size: int, feature_size: int) -> None:
    """Gives the two input projections their shapes without evaluating anything (a rank with an
    empty shard, a parameter count before the first batch).  No-op once they are known."""
    for layer, width in ((self.context_projection, embedding_size),
                         (self.blank_projection, feature_size)):
      if torch.nn.parameter.is_lazy(layer.weight):
        dev = self.joint_projection_to_blank.weight.device
        with torch.no_grad():
          layer(torch.zeros([1, width], device=dev))

  def is_materialized(self) -> bool:
    return not any(torch.nn.parameter.is_lazy(l.weight)
                   for l in (self.context_projection, self.blank_projection))

  def _check_lazy(self, cache, frame):
    """torch's lazy modules crash (segfault in the in-place initialiser) when they take their
    shapes inside a torch.func transform: give the shapes first, outside of it."""
    if self.is_materialized():
      return
    from torch._C import _functorch
    if _functorch.maybe_current_level() is not None:
      raise RuntimeError(
          'JointWeightFn was built without embedding_size / feature_size and is used for the '
          'first time inside a torch.func transform; call materialize(embedding_size, '
          'feature_size) (or evaluate it once) before')
    self.materialize(cache.shape[-1], frame.shape[-1])

  def forward(self, cache, frame, state=None):
    self._check_lazy(cache, frame)
    context_embeddings = cache
    if state is None:
      frame = frame.unsqueeze(-2)                      # [..., 1, D]
    else:
      context_embeddings = torch.index_select(context_embeddings, 0, state.reshape(-1).long())
      context_embeddings = context_embeddings.reshape(*state.shape, -1)
    joint = torch.tanh(self.context_projection(context_embeddings) +
                       self.blank_projection(frame))
    blank = self.joint_projection_to_blank(joint).squeeze(-1)
    lexical = self.joint_projection_to_vocab(joint)
    return blank, lexical

  def all_frames(self, cache, frames):
    self._check_lazy(cache, frames)
    from . import joint as joint_ops   # CUDA (tcgen05) vocabulary projection
    return joint_ops.joint_all_frames(self, cache, frames)

  def string_frames(self, cache, frames, states):
    """The joint network on the U+1 states of every utterance's label string: per utterance one
    kernel call with proj_ctx gathered at its states (C := U+1, N := T), i.e. (U+1) / C of the
    work and memory of all_frames."""
    self._check_lazy(cache, frames)
    from . import joint as joint_ops
    return joint_ops.joint_string_frames(self, cache, frames, states)


class SharedEmbCacher(WeightFnCacher[torch.Tensor]):
  """A trainable, independent context embedding table (weight_fns.py:230-242);
  returns the [num_context_states, embedding_size] tensor JointWeightFn expects."""

  def __init__(self, num_context_states: int, embedding_size: int,
               device: Optional[str] = None, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self.num_context_states = num_context_states
    self.embedding_size = embedding_size
    self.device = device        # None: torch's current default device
    self.embedding = nn.Embedding(num_context_states, embedding_size, device=device)

  def forward(self) -> torch.Tensor:
    return self.embedding.weight


class SharedRNNCacher(WeightFnCacher[torch.Tensor]):
  """Builds the context embedding table by running the n-gram context labels
  through an RNN (weight_fns.py:245-294).

  Used with contexts.FullNGram: row s of the result is the RNN state after
  reading <start> followed by the labels of n-gram s, in FullNGram's state
  order (start, unigrams, bigrams, ...).  Unlike the reference, which builds a
  fresh randomly initialised LSTMCell on every call when `rnn_cell` is None
  (SURVEY D7), the default cell is created ONCE and registered, so the cacher
  has trainable state.  This runs once per optimiser step (C rows), outside
  the per-frame hot path, on whatever device the module lives on.
  """

  def __init__(self, vocab_size: int, context_size: int, rnn_size: int, rnn_embedding_size: int,
               rnn_cell: Optional[nn.RNNCellBase] = None, device: Optional[str] = None,
               *args, **kwargs):
    super().__init__(*args, **kwargs)
    self.vocab_size = vocab_size
    self.context_size = context_size
    self.rnn_size = rnn_size
    self.rnn_embedding_size = rnn_embedding_size
    self.rnn_cell = (rnn_cell if rnn_cell is not None else
                     nn.LSTMCell(rnn_embedding_size, rnn_size, device=device))
    self.embedding = nn.Embedding(vocab_size + 1, rnn_embedding_size, device=device)

  def _step(self, inputs, carry):
    """One cell step; returns (hidden, cell) like the reference's unpacking."""
    out = self.rnn_cell(inputs) if carry is None else self.rnn_cell(inputs, carry)
    if isinstance(out, tuple):
      return out
    return out, out

  def forward(self) -> torch.Tensor:
    dev = self.embedding.weight.device
    feed_cell_state = isinstance(self.rnn_cell, nn.LSTMCell)
    v = self.vocab_size
    hidden, cell = self._step(self.embedding(torch.zeros([1], dtype=torch.long, device=dev)), None)
    parts = [cell]
    inputs = None
    for i in range(self.context_size):
      if i == 0:
        inputs = self.embedding(torch.arange(1, v + 1, device=dev))
      else:
        inputs = inputs.repeat(v, *([1] * (inputs.ndim - 1)))          # 'n ... -> (v n) ...'
      tiled_hidden = torch.repeat_interleave(hidden, v, dim=0)          # 'n ... -> (n v) ...'
      if feed_cell_state:
        carry = (tiled_hidden, torch.repeat_interleave(cell, v, dim=0))
      else:
        carry = tiled_hidden
      hidden, cell = self._step(inputs, carry)
      parts.append(cell)
    return torch.concatenate(parts, dim=0)


class NullCacher(WeightFnCacher[type(None)]):
  """Returns None; used with TableWeightFn (weight_fns.py:297-304)."""

  def forward(self) -> None:
    return None


class TableWeightFn(WeightFn[type(None)]):
  """Looks arc weights up in a fixed table, for tests (weight_fns.py:307-342).

  table: [batch_dims..., input_vocab_size, num_context_states, 1 + vocab_size];
  frame[..., 0] is cast to an integer "input label"; table[..., 0] holds the
  blank weights, table[..., 1:] the lexical weights.
  """

  def __init__(self, table: torch.Tensor, *args, **kwargs) -> None:
    super().__init__(*args, **kwargs)
    self.table = table

  def forward(self, cache, frame, state=None):
    del cache
    *batch_dims, input_vocab_size, num_context_states, _ = self.table.shape
    if tuple(frame.shape[:-1]) != tuple(batch_dims):
      raise ValueError(f'frame should have batch_dims={tuple(batch_dims)} but '
                       f'got ({frame.shape[:-1]})')
    index = frame[..., 0].to(torch.int64)
    table = self.table.float()
    idx = index.reshape(*index.shape, 1, 1, 1).expand(*index.shape, 1, *table.shape[-2:])
    weights = torch.gather(table, len(batch_dims), idx).squeeze(len(batch_dims))
    if state is not None:
      state = torch.broadcast_to(state, tuple(batch_dims)).to(torch.int64)
      sidx = state.reshape(*state.shape, 1, 1).expand(*state.shape, 1, weights.shape[-1])
      weights = torch.gather(weights, len(batch_dims), sidx).squeeze(len(batch_dims))
    return weights[..., 0], weights[..., 1:]

  def string_frames(self, cache, frames, states):
    del cache
    *batch_dims, input_vocab_size, num_context_states, width = self.table.shape
    if tuple(frames.shape[:-2]) != tuple(batch_dims) or len(batch_dims) != 1:
      raise ValueError(f'frame should have batch_dims={tuple(batch_dims)} but '
                       f'got ({tuple(frames.shape[:-2])})')
    b, t = frames.shape[:2]
    u1 = states.shape[1]
    index = frames[..., 0].to(torch.int64)                                  # [B, T]
    table = self.table.float()
    bi = torch.arange(b, device=table.device)[:, None, None]
    w = table[bi, index[:, :, None], states.to(torch.int64)[:, None, :]]    # [B, T, U1, 1+V]
    return w[..., 0].contiguous(), w[..., 1:].contiguous()

  def all_frames(self, cache, frames):
    del cache
    *batch_dims, input_vocab_size, num_context_states, width = self.table.shape
    if tuple(frames.shape[:-2]) != tuple(batch_dims):
      raise ValueError(f'frame should have batch_dims={tuple(batch_dims)} but '
                       f'got ({tuple(frames.shape[:-2])})')
    index = frames[..., 0].to(torch.int64)                       # [batch..., T]
    table = self.table.float()
    idx = index.reshape(*index.shape, 1, 1).expand(*index.shape, num_context_states, width)
    weights = torch.gather(table, len(batch_dims), idx)          # [batch..., T, C, 1+V]
    return weights[..., 0].contiguous(), weights[..., 1:].contiguous()
