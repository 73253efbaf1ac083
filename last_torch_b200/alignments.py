"""Time-synchronous alignment lattices (drop-in for last_torch.alignments).

FrameDependent and FrameLabelDependent keep the reference's topology accessors
and per-frame `forward` / `backward` / `string_forward` methods
(/root/reference/last_torch/alignments.py).  The per-frame methods compose the
CUDA-backed semiring ops; RecognitionLattice does NOT call them frame by frame
-- it asks the alignment for its `max_expansions` and runs the whole T-frame
recursion inside the persistent lattice kernels (ops.LatticeForward).
"""

from __future__ import annotations

import abc
from collections.abc import Sequence
from typing import Optional

import torch

from . import _native as N
from . import contexts
from . import semirings


class TimeSyncAlignmentLattice(abc.ABC):
  """Interface (alignments.py:26-230): a frame-local acyclic automaton over
  {blank, lexical}, repeated once per frame."""

  @abc.abstractmethod
  def num_states(self) -> int:
    """Number of non-final frame-local alignment states."""

  @abc.abstractmethod
  def start(self) -> int:
    """Start state of the frame-local lattice."""

  @abc.abstractmethod
  def blank_next(self, state: int) -> Optional[int]:
    """Next state after the blank arc (the start state if it ends the frame)."""

  @abc.abstractmethod
  def lexical_next(self, state: int) -> Optional[int]:
    """Next state after a lexical arc; None if there is none."""

  @abc.abstractmethod
  def topological_visit(self) -> list[int]:
    """Non-final states in topological order."""

  @abc.abstractmethod
  def kernel_max_expansions(self) -> int:
    """Alignment id understood by the lattice kernels (LT_FRAME_DEPENDENT or k)."""

  @abc.abstractmethod
  def forward(self, alpha, blank, lexical, context, semiring):
    """One frame of the forward algorithm on the recognition lattice."""

  @abc.abstractmethod
  def backward(self, alpha, blank, lexical, beta, log_z, context):
    """One frame of the backward algorithm (Log semiring): next beta + marginals."""

  @abc.abstractmethod
  def string_forward(self, alpha, blank, lexical, semiring):
    """One frame of the forward algorithm after intersection with a string."""


def shift_down(x: torch.Tensor, semiring: semirings.Semiring[torch.Tensor]) -> torch.Tensor:
  """output[..., i + 1] = x[..., i], output[..., 0] = semiring zero
  (alignments.py:233-248)."""
  pad = semiring.zeros((*x.shape[:-1], 1), x.dtype).to(x.device)
  return torch.cat([pad, x[..., :-1]], dim=-1)


def check_num_weights(alignment: TimeSyncAlignmentLattice, blank: Sequence[torch.Tensor],
                      lexical: Sequence[torch.Tensor]) -> None:
  """alignments.py:251-263."""
  num_states = alignment.num_states()
  if len(blank) != num_states:
    raise ValueError(f'blank should be a length {num_states} sequence of ndarrays, '
                     f'but got length {len(blank)}')
  if len(lexical) != num_states:
    raise ValueError(f'lexical should be a length {num_states} sequence of ndarrays, '
                     f'but got length {len(lexical)}')


class FrameDependent(TimeSyncAlignmentLattice):
  """Each frame emits exactly one label: blank or lexical (alignments.py:266-329)."""

  def num_states(self) -> int:
    return 1

  def start(self) -> int:
    return 0

  def blank_next(self, state: int) -> Optional[int]:
    return 0

  def lexical_next(self, state: int) -> Optional[int]:
    return 0

  def topological_visit(self) -> list[int]:
    return [0]

  def kernel_max_expansions(self) -> int:
    return N.FRAME_DEPENDENT

  def forward(self, alpha, blank, lexical, context, semiring):
    check_num_weights(self, blank, lexical)
    stay = semiring.times(alpha, blank[0])
    move = context.forward_reduce(semiring.times(alpha.unsqueeze(-1), lexical[0]), semiring)
    return semiring.plus(stay, move)

  def backward(self, alpha, blank, lexical, beta, log_z, context):
    check_num_weights(self, blank, lexical)
    log = semirings.Log
    blank_beta = blank[0] + beta
    lexical_beta = lexical[0] + context.backward_broadcast(beta)
    log_scale = alpha - log_z.unsqueeze(-1)
    blank_marginal = torch.exp(blank_beta + log_scale)
    lexical_marginal = torch.exp(lexical_beta + log_scale.unsqueeze(-1))
    next_beta = log.plus(blank_beta, log.sum(lexical_beta, dim=-1))
    return next_beta, [blank_marginal], [lexical_marginal]

  def string_forward(self, alpha, blank, lexical, semiring):
    check_num_weights(self, blank, lexical)
    stay = semiring.times(alpha, blank[0])
    move = shift_down(semiring.times(alpha, lexical[0]), semiring)
    return semiring.plus(stay, move)


class FrameLabelDependent(TimeSyncAlignmentLattice):
  """Each frame emits up to `max_expansions` lexical labels and then a blank
  (alignments.py:331-432)."""

  def __init__(self, max_expansions: int) -> None:
    super().__init__()
    self.max_expansions = max_expansions

  def num_states(self) -> int:
    return self.max_expansions + 1

  def start(self) -> int:
    return 0

  def blank_next(self, state: int) -> Optional[int]:
    return 0

  def lexical_next(self, state: int) -> Optional[int]:
    return state + 1 if state + 1 <= self.max_expansions else None

  def topological_visit(self) -> list[int]:
    return list(range(self.max_expansions + 1))

  def kernel_max_expansions(self) -> int:
    if self.max_expansions < 1:
      raise NotImplementedError('the lattice kernels need max_expansions >= 1')
    return self.max_expansions

  def forward(self, alpha, blank, lexical, context, semiring):
    check_num_weights(self, blank, lexical)
    terminated = [semiring.times(alpha, blank[0])]
    last = alpha
    for i in range(self.max_expansions):
      last = context.forward_reduce(semiring.times(last.unsqueeze(-1), lexical[i]), semiring)
      terminated.append(semiring.times(last, blank[i + 1]))
    return semiring.sum(torch.stack(terminated), dim=0)

  def backward(self, alpha, blank, lexical, beta, log_z, context):
    check_num_weights(self, blank, lexical)
    log = semirings.Log
    k = self.max_expansions
    # forward weights at every expansion level
    level_alpha = [alpha]
    for i in range(k):
      level_alpha.append(
          context.forward_reduce(level_alpha[-1].unsqueeze(-1) + lexical[i], log))
    blank_log_scale = beta - log_z.unsqueeze(-1)
    blank_marginals = [torch.exp(level_alpha[i] + blank[i] + blank_log_scale)
                       for i in range(k + 1)]
    next_beta = blank[k] + beta
    lexical_marginals = [None] * k
    for j in range(k - 1, -1, -1):
      lexical_beta = lexical[j] + context.backward_broadcast(next_beta)
      log_scale = level_alpha[j] - log_z.unsqueeze(-1)
      lexical_marginals[j] = torch.exp(lexical_beta + log_scale.unsqueeze(-1))
      next_beta = log.plus(blank[j] + beta, log.sum(lexical_beta, dim=-1))
    lexical_marginals.append(torch.zeros_like(lexical[k]))
    return next_beta, blank_marginals, lexical_marginals

  def string_forward(self, alpha, blank, lexical, semiring):
    check_num_weights(self, blank, lexical)
    terminated = [semiring.times(alpha, blank[0])]
    last = alpha
    for i in range(self.max_expansions):
      last = shift_down(semiring.times(last, lexical[i]), semiring)
      terminated.append(semiring.times(last, blank[i + 1]))
    return semiring.sum(torch.stack(terminated), dim=0)
