"""Context dependencies (drop-in for last_torch.contexts).

FullNGram keeps the reference's constructor, state numbering and methods
(/root/reference/last_torch/contexts.py:149-263).  The index math is integer
host/torch plumbing; the semiring reductions inside forward_reduce go through
the CUDA semiring kernels.  Inside RecognitionLattice the whole T-frame
recursion is fused into the lattice kernels and these per-frame methods are
not on the path.
"""

from __future__ import annotations

import abc
import dataclasses

import torch

from . import semirings


class ContextDependency(abc.ABC):
  """Interface (contexts.py:25-146)."""

  @abc.abstractmethod
  def shape(self) -> tuple[int, int]:
    """(num_states, vocab_size)."""

  @abc.abstractmethod
  def start(self) -> int:
    """The start state id."""

  @abc.abstractmethod
  def next_state(self, state: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """Takes a transition; label 0 (epsilon) stays in place."""

  @abc.abstractmethod
  def forward_reduce(self, weights: torch.Tensor,
                     semiring: semirings.Semiring[torch.Tensor]) -> torch.Tensor:
    """result[..., q] = sum_{p -y-> q} weights[..., p, y]."""

  @abc.abstractmethod
  def backward_broadcast(self, weights: torch.Tensor) -> torch.Tensor:
    """result[..., p, y] = weights[..., q] for p -y-> q."""

  def walk_states(self, labels: torch.Tensor) -> torch.Tensor:
    """States visited along label sequences (contexts.py:109-146).

    labels [batch_dims..., num_labels] -> [batch_dims..., num_labels + 1];
    int64 on the labels' device (the reference returns float32).
    """
    labels = labels.to(torch.int64)
    state = torch.full(labels.shape[:-1], self.start(), dtype=torch.int64, device=labels.device)
    states = [state]
    for i in range(labels.shape[-1]):
      state = self.next_state(state, labels[..., i])
      states.append(state)
    return torch.stack(states, dim=-1)


@dataclasses.dataclass(frozen=True)
class FullNGram(ContextDependency):
  """Full n-gram context dependency (contexts.py:149-263).

  States are all n-grams of length 0..context_size in lexicographic order:
  state 0 is the empty n-gram, 1..vocab_size the unigrams, and so on.
  """
  vocab_size: int
  context_size: int

  def __post_init__(self):
    if self.vocab_size <= 0:
      raise ValueError('vocab_size should be > 0, but got '
                       f'vocab_size={self.vocab_size}')
    if self.context_size < 0:
      raise ValueError('context_size should be >= 0, but got '
                       f'context_size={self.context_size}')

  def num_states(self) -> int:
    return sum(int(self.vocab_size**i) for i in range(self.context_size + 1))

  def shape(self) -> tuple[int, int]:
    return self.num_states(), self.vocab_size

  def start(self) -> int:
    return 0

  def _num_ascending(self) -> int:
    return sum(self.vocab_size**i for i in range(self.context_size))

  def next_state(self, state: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """contexts.py:190-205; accepts int or float index tensors."""
    v, n = self.vocab_size, self.context_size
    num_asc = self._num_ascending()
    ascend = state * v + label
    if n == 0:
      full = torch.zeros_like(ascend)
    else:
      full = (state - num_asc) % (v**(n - 1)) * v + num_asc + label - 1
    nxt = torch.where(state < num_asc, ascend, full)
    return torch.where(label == 0, state, nxt)

  def forward_reduce(self, weights: torch.Tensor,
                     semiring: semirings.Semiring[torch.Tensor]) -> torch.Tensor:
    """contexts.py:207-230."""
    batch_dims = tuple(weights.shape[:-2])
    if tuple(weights.shape[-2:]) != self.shape():
      raise ValueError(f'weights.shape[-2:] should be {self.shape()} but got'
                       f' {tuple(weights.shape[-2:])}')
    v, n = self.vocab_size, self.context_size
    parts = []
    if n > 0:
      parts.append(semiring.zeros(batch_dims + (1,), weights.dtype).to(weights.device))
    low = sum(v**i for i in range(0, n - 1))
    parts.append(weights[..., :low, :].reshape(batch_dims + (-1,)))
    parts.append(
        semiring.sum(weights[..., low:, :].reshape(batch_dims + (-1, v**n)), dim=-2))
    return torch.concatenate(parts, dim=-1)

  def backward_broadcast(self, weights: torch.Tensor) -> torch.Tensor:
    """contexts.py:232-256."""
    batch_dims = tuple(weights.shape[:-1])
    num_states = weights.shape[-1]
    if num_states != self.num_states():
      raise ValueError(f'weights.shape[-1] should be {self.num_states()} but '
                       f'got {num_states}')
    v, n = self.vocab_size, self.context_size
    if n == 0:
      return torch.broadcast_to(weights.unsqueeze(-1), tuple(weights.shape) + (v,))
    num_asc = self._num_ascending()
    part_a = weights[..., 1:num_asc].reshape(batch_dims + (-1, v))
    part_b = torch.broadcast_to(weights[..., None, num_asc:],
                                batch_dims + (1 + v, v**n)).reshape(batch_dims + (-1, v))
    return torch.concatenate([part_a, part_b], dim=-2)

  def next_state_table(self) -> torch.Tensor:
    """[num_states, vocab_size] table of next states (contexts.py:258-263)."""
    num_states, vocab_size = self.shape()
    return self.next_state(
        torch.arange(num_states).unsqueeze(-1), torch.arange(vocab_size).unsqueeze(0) + 1)


@dataclasses.dataclass(frozen=True)
class NextStateTable(ContextDependency):
  """Context dependency described as a transition lookup table
  (contexts.py:266-324).

  next_state_table: [num_states, vocab_size] int32; next_state_table[p, y - 1]
  is the state reached from p with label y.

  forward_reduce implements the documented contract out[q] = (+)_{p -y-> q}
  w[p, y] (contexts.py:74-90) for EVERY semiring through a CSR of incoming arcs
  (the reference only works for the Real semiring); MaxTropical ties
  keep the first arc in (p, y) order.  Inside RecognitionLattice the table-driven
  lattice kernels (csrc/lattice_table.cu) are used.
  """
  next_state_table: torch.Tensor

  def __post_init__(self):
    if self.next_state_table.ndim != 2:
      raise ValueError(
          'next_state_table should have shape [num_states, vocab_size], but'
          f'got shape {self.next_state_table.shape}')
    if 0 in self.next_state_table.size():
      raise ValueError('next_state_table should have a non-zero size, but '
                       f'got shape {self.next_state_table.shape}')
    if self.next_state_table.dtype != torch.int32:
      raise ValueError('next_state_table should be an int32 ndarray, but '
                       f'got dtype {self.next_state_table.dtype}')
    object.__setattr__(self, '_device_cache', {})

  def shape(self) -> tuple[int, int]:
    return self.next_state_table.shape

  def start(self) -> int:
    return 0

  def next_state(self, state: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """contexts.py:297-304; accepts int or float index tensors."""
    table = self.next_state_table.to(state.device)
    is_epsilon = label == 0
    zero_based_label = torch.where(is_epsilon, torch.zeros_like(label), label - 1)
    nextstate = table[state.long(), zero_based_label.long()].to(state.dtype)
    return torch.where(is_epsilon, state, nextstate)

  def kernel_tables(self, device):
    """(table [C,V], in_offsets [C+1], in_arcs [C*V]) int32 on `device`: the CSR of
    incoming arcs the kernels pull from (arcs of a destination in ascending p*V+y order)."""
    key = str(device)
    if key not in self._device_cache:
      table = self.next_state_table.to(device=device, dtype=torch.int64).contiguous()
      c = table.shape[0]
      if int(table.min()) < 0 or int(table.max()) >= c:
        raise ValueError(f'next_state_table entries should be in [0, {c}), but got '
                         f'[{int(table.min())}, {int(table.max())}]')
      flat = table.reshape(-1)
      order = torch.sort(flat, stable=True).indices
      counts = torch.bincount(flat, minlength=c)
      offsets = torch.zeros([c + 1], dtype=torch.int64, device=device)
      offsets[1:] = torch.cumsum(counts, 0)
      self._device_cache[key] = (table.to(torch.int32).contiguous(),
                                 offsets.to(torch.int32).contiguous(),
                                 order.to(torch.int32).contiguous())
    return self._device_cache[key]

  def forward_reduce(self, weights: torch.Tensor,
                     semiring: semirings.Semiring[torch.Tensor]) -> torch.Tensor:
    if weights.shape[-2:] != self.shape():
      raise ValueError(f'weights.shape[-2:] should be {self.shape()} but got'
                       f' {weights.shape[-2:]}')
    from . import ops
    if not weights.is_floating_point():
      weights = weights.to(torch.float32)     # the reference's torch ops take integer tensors
    return ops.TableReduce.apply(weights, self, semirings.kernel_id(semiring))

  def backward_broadcast(self, weights: torch.Tensor) -> torch.Tensor:
    num_states = weights.shape[-1]
    if num_states != self.shape()[0]:
      raise ValueError(f'weights.shape[-1] should be {self.shape()[0]} but '
                       f'got {num_states}')
    return weights[..., self.next_state_table.to(device=weights.device, dtype=torch.int64)]
