"""ctypes binding of the C ABI in include/last_lattice.h.

The shared library is built in-tree by `__graft_entry__.build()` (nvcc,
sm_100a) as last_torch_b200/_C/liblast_lattice.so.  There is NO fallback: if
the library is missing, or an op is handed a tensor that is not a contiguous
fp32/int32 CUDA tensor, the call raises.
"""

from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LT_LIBRARY selects another build of the same sources (A/B experiments with compile-time switches)
LIB_PATH = os.environ.get('LT_LIBRARY') or os.path.join(_HERE, '_C', 'liblast_lattice.so')

REAL, LOG, MAXTROPICAL = 0, 1, 2
FRAME_DEPENDENT = -1
FLAG_FORCE_GENERIC = 1
FLAG_CLUSTER_SHIFT = 8
FLAG_GRAD_SPLIT = 16
FLAG_LEVEL_WEIGHTS = 32

_c_int = ctypes.c_int
_c_i64 = ctypes.c_int64
_ptr = ctypes.c_void_p
_c_float = ctypes.c_float
_c_uint = ctypes.c_uint

# name -> argtypes; every symbol declared in include/last_lattice.h.
SIGNATURES = {
    'lt_version': [],
    'lt_last_error': [],
    'lt_launch_count': [],
    'lt_device_info': [_ptr, _ptr, _ptr],
    'lt_lattice_forward': [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _c_int,
                           _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_uint, _ptr],
    'lt_lattice_backward': [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _c_int,
                            _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_uint, _ptr],
    'lt_lattice_norm_supported': [_c_int, _c_int, _c_int, _c_int, _c_uint],
    'lt_lattice_forward_norm': [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _c_int,
                                _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_uint, _ptr],
    'lt_lattice_backward_norm': [_c_int, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _c_int,
                                 _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_uint, _ptr],
    'lt_alphas_denormalize': [_ptr, _ptr, _c_int, _c_int, _c_int, _ptr],
    'lt_linear_forward': [_ptr, _ptr, _ptr, _c_i64, _c_int, _c_int, _ptr],
    'lt_linear_tensor_core': [_c_i64, _c_int, _c_int],
    'lt_linear_wgrad_workspace_bytes': [_c_i64, _c_int, _c_int],
    'lt_linear_wgrad': [_ptr, _ptr, _ptr, _c_i64, _c_int, _c_int, _ptr, _ptr],
    'lt_lattice_expectation_supported': [_c_int, _c_int, _c_int, _c_uint],
    'lt_lattice_expectation': [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _c_int, _ptr, _ptr,
                               _ptr, _ptr, _ptr, _ptr, _c_uint, _ptr],
    'lt_string_norm_supported': [_c_int, _c_int, _c_int],
    'lt_string_forward_norm': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int,
                               _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    'lt_string_backward_norm': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int,
                                _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    'lt_viterbi_backtrace': [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int,
                             _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    'lt_walk_states': [_c_int, _c_int, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr],
    'lt_walk_states_checked': [_c_int, _c_int, _ptr, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr,
                               _ptr],
    'lt_stream_delay': [_c_uint, _ptr],
    'lt_string_gather': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int,
                         _ptr, _ptr, _ptr],
    'lt_string_scatter_add': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int,
                              _c_float, _ptr, _ptr, _ptr, _ptr],
    'lt_string_scatter_add_split': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int,
                                    _c_float, _ptr, _ptr, _ptr, _ptr],
    'lt_string_forward': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int,
                          _ptr, _ptr, _ptr, _ptr],
    'lt_string_backward': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int,
                           _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    'lt_semiring_plus_forward': [_c_int, _ptr, _ptr, _ptr, _c_i64, _ptr],
    'lt_semiring_plus_backward': [_c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64, _ptr],
    'lt_semiring_sum_forward': [_c_int, _ptr, _c_i64, _c_i64, _c_i64, _ptr, _ptr, _ptr],
    'lt_semiring_sum_backward': [_c_int, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_i64, _c_i64, _ptr,
                                 _ptr],
    'lt_joint_workspace_bytes': [_c_i64, _c_int, _c_int, _c_int],
    'lt_joint_forward': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _c_int, _c_int,
                         _ptr, _ptr, _ptr, _ptr],
    'lt_joint_lattice_fused_supported': [_c_int, _c_int, _c_int, _c_int, _c_int],
    'lt_joint_lattice_forward_fused': [_c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                       _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr],
    'lt_joint_split_rows': [_ptr, _ptr, _c_i64, _c_int, _ptr],
    'lt_set_option': [ctypes.c_char_p, _c_int],
    'lt_get_option': [ctypes.c_char_p],
    'lt_joint_backward': [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _c_int, _c_int,
                          _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr],
    'lt_joint_backward_split_supported': [_c_i64, _c_int, _c_int, _c_int],
    'lt_lattice_backward_split_supported': [_c_int, _c_int, _c_int, _c_int, ctypes.c_uint],
    'lt_joint_backward_workspace_bytes': [_c_i64, _c_int, _c_int, _c_int],
    'lt_table_lattice_forward': [_c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _c_int, _ptr, _ptr,
                                 _ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                 _ptr],
    'lt_table_lattice_backward': [_c_int, _c_int, _ptr, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int,
                                  _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    'lt_table_viterbi_backtrace': [_c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int,
                                   _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    'lt_table_lattice_cluster': [_c_int, _c_int, _c_int, _c_int],
    'lt_table_reduce_forward': [_c_int, _ptr, _ptr, _ptr, _c_i64, _c_int, _c_int, _ptr, _ptr,
                                _ptr],
    'lt_local_normalize_forward': [_c_int, _ptr, _ptr, _c_i64, _c_int, _ptr, _ptr, _ptr],
    'lt_local_normalize_backward': [_c_int, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _ptr, _ptr,
                                    _ptr],
    'lt_table_reduce_backward': [_c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _c_int,
                                 _ptr, _ptr],
}

_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
  pass


def lib():
  """Loads liblast_lattice.so (once).  Raises loudly if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  with _lock:
    if _lib is not None:
      return _lib
    if not os.path.exists(LIB_PATH):
      raise NativeLibraryError(
          f'{LIB_PATH} not found: build the CUDA extension first '
          '(python -c "import __graft_entry__ as g; g.build()"). '
          'last_torch_b200 has no CPU or eager fallback.')
    handle = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
      fn = getattr(handle, name)     # AttributeError if a symbol is missing
      fn.argtypes = argtypes
      fn.restype = (ctypes.c_char_p if name == 'lt_last_error' else
                    ctypes.c_ulonglong if name == 'lt_launch_count' else
                    ctypes.c_int64 if name in ('lt_joint_workspace_bytes',
                                                'lt_joint_backward_workspace_bytes',
                                                'lt_linear_wgrad_workspace_bytes') else _c_int)
    _lib = _TimedLib(handle)
  return _lib


# Optional per-call CUDA-event timing (bench.py sets KERNEL_TIMER to a list; every
# native call then appends (name, start_event, end_event) recorded on the
# current stream of the current device -- the stream the kernel is launched on).
KERNEL_TIMER = None
_UNTIMED = ('lt_last_error', 'lt_version', 'lt_device_info', 'lt_launch_count', 'lt_set_option',
            'lt_get_option',
            'lt_joint_backward_split_supported', 'lt_lattice_backward_split_supported',
            'lt_lattice_norm_supported', 'lt_string_norm_supported',
            'lt_lattice_expectation_supported',
            'lt_joint_lattice_fused_supported',
            'lt_joint_workspace_bytes', 'lt_joint_backward_workspace_bytes',
            'lt_linear_wgrad_workspace_bytes')


# NVTX range per C-ABI call (nsys / ncu --nvtx show the entry point around its kernels); the
# push/pop cost ~0.1 us when no tool is attached.  LT_NO_NVTX=1 (read once) switches them off.
_NVTX = not os.environ.get('LT_NO_NVTX')


class _TimedLib:
  def __init__(self, handle):
    self._handle = handle
    self._wrapped = {}

  def __getattr__(self, name):
    fn = getattr(self._handle, name)
    if name in _UNTIMED:
      return fn
    cached = self._wrapped.get(name)
    if cached is not None:
      return cached
    nvtx = torch.cuda.nvtx if (_NVTX and torch.cuda.is_available()) else None

    def call(*args):
      timer = KERNEL_TIMER
      if nvtx is not None:
        nvtx.range_push(name)
      try:
        if timer is None:
          return fn(*args)
        start = torch.cuda.Event(enable_timing=True)
        end = torch.cuda.Event(enable_timing=True)
        start.record()
        rc = fn(*args)
        end.record()
        timer.append((name, start, end))
        return rc
      finally:
        if nvtx is not None:
          nvtx.range_pop()
    self._wrapped[name] = call
    return call


class option:
  """Context manager for a debug / test switch of the library (lt_set_option):
      with N.option('LT_JOINT_SIMT', 1): ...
  """

  def __init__(self, name: str, value: int = 1):
    self.name = name.encode()
    self.value = int(value)
    self.saved = 0

  def __enter__(self):
    self.saved = lib().lt_get_option(self.name)
    if self.saved < 0:
      raise KeyError(self.name.decode())
    check(lib().lt_set_option(self.name, self.value), 'lt_set_option')
    return self

  def __exit__(self, *exc):
    lib().lt_set_option(self.name, self.saved)
    return False


def check(rc: int, what: str) -> None:
  if rc == 0:
    return
  msg = lib().lt_last_error()
  msg = msg.decode('utf-8', 'replace') if msg else ''
  if rc == 1:
    raise ValueError(f'{what}: {msg}')
  raise RuntimeError(f'{what} failed (status {rc}): {msg}')


def ptr(t):
  """Device pointer of a tensor (None -> NULL)."""
  return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
  """Validates that `t` is something the kernels can read: CUDA, `dtype`,
  contiguous.  Never copies to or computes on the CPU."""
  if not isinstance(t, torch.Tensor):
    raise TypeError(f'{name} must be a torch.Tensor, got {type(t)}')
  if not t.is_cuda:
    raise RuntimeError(
        f'{name} lives on {t.device}; last_torch_b200 runs on CUDA (sm_100a) only and has '
        'no CPU fallback')
  if t.dtype != dtype:
    raise TypeError(f'{name} must have dtype {dtype}, got {t.dtype}')
  return t.contiguous()
