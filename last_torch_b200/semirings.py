"""Semirings (drop-in for last_torch.semirings).

Same names and call signatures as the reference
(/root/reference/last_torch/semirings.py): Real, Log, MaxTropical, Expectation,
LogLogExpectation, Cartesian, value_shape, value_dtype.  The (+) operations of
Log and MaxTropical run as CUDA kernels through the C ABI with hand-written
gradients (the reference's Log autograd functions are broken as shipped; the
documented "safe gradient" semantics of semirings.py:222-241 are implemented).
"""

from __future__ import annotations

from collections.abc import Sequence
import dataclasses
from typing import Any, Callable, Generic, Optional, TypeVar

import torch
import torch.utils._pytree as pytree

from . import _native as N
from . import ops

DType = Any
PyTree = Any
T = TypeVar('T')
S = TypeVar('S')


def value_shape(x: PyTree) -> tuple[int, ...]:
  """Common shape of the leaves of a semiring value (semirings.py:30-62)."""
  shapes = []
  for leaf in pytree.tree_leaves(x):
    if leaf is None:
      raise ValueError(f'No common shape can be derived for an empty PyTree: {x!r}')
    shapes.append(tuple(leaf.shape))
  if not shapes:
    raise ValueError(f'No common shape can be derived for an empty PyTree: {x!r}')
  result = shapes[0]
  for s in shapes[1:]:
    if s != result:
      raise ValueError('A semiring value must consist of ndarrays of a common shape. '
                       f'Got inconsistent shapes {result} vs {s} for PyTree: {x!r}')
  return result


def value_dtype(x: PyTree) -> DType:
  """dtypes in the same structure as x (semirings.py:64-78)."""
  return pytree.tree_map(lambda leaf: leaf.dtype, x)


class Semiring(Generic[T]):
  """Interface (semirings.py:80-141)."""

  def zeros(self, shape: Sequence[int], dtype: Optional[DType] = None) -> T:
    raise NotImplementedError

  def ones(self, shape: Sequence[int], dtype: Optional[DType] = None) -> T:
    raise NotImplementedError

  def times(self, a: T, b: T) -> T:
    raise NotImplementedError

  def plus(self, a: T, b: T) -> T:
    raise NotImplementedError

  def prod(self, a: T, dim: int) -> T:
    raise NotImplementedError

  def sum(self, a: T, dim: int) -> T:
    raise NotImplementedError


def _check_axis(a: torch.Tensor, axis: int) -> None:
  """semirings.py:176-181."""
  if not isinstance(axis, int) or isinstance(axis, bool):
    raise ValueError(f'Only int axis is supported, got axis={axis!r}')
  if not -a.ndim <= axis < a.ndim:
    raise ValueError(f'Invalid reduction axis={axis!r} for input shape {a.shape}')


class _Real(Semiring[torch.Tensor]):
  """Real semiring (semirings.py:143-173).  (+) and (x) are the native tensor
  + and *, exactly as in the reference; the lattice recursions over this
  semiring run in the CUDA kernels (kernel id LT_REAL)."""
  name = 'Real'
  kernel_id = N.REAL

  @staticmethod
  def zeros(shape, dtype=None):
    return torch.zeros(tuple(shape), dtype=dtype)

  @staticmethod
  def ones(shape, dtype=None):
    return torch.ones(tuple(shape), dtype=dtype)

  @staticmethod
  def times(a, b):
    return a * b

  @staticmethod
  def plus(a, b):
    return a + b

  @staticmethod
  def prod(a, dim):
    return torch.prod(a, dim)

  @staticmethod
  def sum(a, dim):
    return torch.sum(a, dim)


Real = _Real()


class _KernelSemiring(Semiring[torch.Tensor]):
  """Shared body of Log and MaxTropical: zero = -inf, one = 0, (x) = +."""
  name = ''
  kernel_id = -1

  @staticmethod
  def zeros(shape, dtype=None):
    return torch.full(tuple(shape), -torch.inf, dtype=dtype)

  @staticmethod
  def ones(shape, dtype=None):
    return torch.zeros(tuple(shape), dtype=dtype)

  @staticmethod
  def times(a, b):
    return a + b

  @staticmethod
  def prod(a, dim):
    return torch.sum(a, dim)

  @classmethod
  def plus(cls, a, b):
    a, b = torch.broadcast_tensors(a, b)
    return ops.SemiringPlus.apply(a, b, cls.kernel_id)

  @classmethod
  def sum(cls, a, dim):
    _check_axis(a, dim)
    if torch.numel(a) > 0:
      return ops.SemiringSum.apply(a, dim, cls.kernel_id)
    # summing an empty axis gives semiring zeros (semirings.py:216-220)
    if dim < 0:
      dim += a.ndim
    shape = a.shape[:dim] + a.shape[dim + 1:]
    return torch.full(shape, -torch.inf, dtype=a.dtype, device=a.device)


class _Log(_KernelSemiring):
  """Log semiring (semirings.py:184-220): (+) is a max-shifted log-add-exp."""
  name = 'Log'
  kernel_id = N.LOG


class _MaxTropical(_KernelSemiring):
  """Max tropical semiring (semirings.py:308-348): (+) is max with exactly one
  non-zero gradient entry even on ties."""
  name = 'MaxTropical'
  kernel_id = N.MAXTROPICAL


Log = _Log()
MaxTropical = _MaxTropical()


@dataclasses.dataclass(frozen=True)
class Expectation(Generic[T, S], Semiring[tuple[T, S]]):
  """Eisner's expectation semiring (semirings.py:404-479); a thin composition
  over the component semirings."""
  w: Semiring[T]
  x: Semiring[S]
  w_to_x: Callable[[T], S]

  def weighted(self, w, v):
    w_is_zero = w == self.w.zeros([], w.dtype).to(w.device)
    safe_v = torch.where(w_is_zero, 0, v)
    return w, self.x.times(self.w_to_x(w), safe_v)

  def zeros(self, shape, dtype=None):
    dw, dx = (None, None) if dtype is None else dtype
    return self.w.zeros(shape, dw), self.x.zeros(shape, dx)

  def ones(self, shape, dtype=None):
    dw, dx = (None, None) if dtype is None else dtype
    return self.w.ones(shape, dw), self.x.zeros(shape, dx)

  def times(self, a, b):
    w_a, x_a = a
    w_b, x_b = b
    w = self.w.times(w_a, w_b)
    x = self.x.plus(self.x.times(self.w_to_x(w_a), x_b), self.x.times(self.w_to_x(w_b), x_a))
    return w, x

  def plus(self, a, b):
    return self.w.plus(a[0], b[0]), self.x.plus(a[1], b[1])

  def sum(self, a, axis):
    return self.w.sum(a[0], axis), self.x.sum(a[1], axis)


LogLogExpectation = Expectation(w=Log, x=Log, w_to_x=lambda x: x)


@dataclasses.dataclass(frozen=True)
class Cartesian(Generic[T, S], Semiring[tuple[T, S]]):
  """Cartesian product of two semirings (semirings.py:487-533)."""
  x: Semiring[T]
  y: Semiring[S]

  def zeros(self, shape, dtype=None):
    dx, dy = (None, None) if dtype is None else dtype
    return self.x.zeros(shape, dx), self.y.zeros(shape, dy)

  def ones(self, shape, dtype=None):
    dx, dy = (None, None) if dtype is None else dtype
    return self.x.ones(shape, dx), self.y.ones(shape, dy)

  def times(self, a, b):
    return self.x.times(a[0], b[0]), self.y.times(a[1], b[1])

  def plus(self, a, b):
    return self.x.plus(a[0], b[0]), self.y.plus(a[1], b[1])

  def sum(self, a, axis):
    return self.x.sum(a[0], axis), self.y.sum(a[1], axis)

  def prod(self, a, axis):
    return self.x.prod(a[0], axis), self.y.prod(a[1], axis)


def kernel_id(semiring) -> int:
  """Kernel-side id of a semiring object; raises for anything the lattice
  kernels do not implement (no fallback)."""
  kid = getattr(semiring, 'kernel_id', None)
  if kid in (N.REAL, N.LOG, N.MAXTROPICAL):
    return kid
  raise NotImplementedError(
      f'{semiring!r} is not supported by the lattice kernels; supported semirings are Real, '
      'Log and MaxTropical')
