"""ctypes wrapper of oracle/lattice_oracle.c (TEST INFRASTRUCTURE ONLY).

Built by `__graft_entry__.build()` into oracle/_build/liblattice_oracle.so.
Used by tests (cross-check against the numpy oracle) and by bench.py's
cpu_baseline / --impl reference legs as the multi-threaded CPU port.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'liblattice_oracle.so')
LIB_PATH_F64 = os.path.join(_HERE, '_build', 'liblattice_oracle_f64.so')   # -DLT_ORACLE_REAL=double
_libs = {}


def available(real='f32'):
  return os.path.exists(LIB_PATH if real == 'f32' else LIB_PATH_F64)


def lib(real='f32'):
  if real not in _libs:
    handle = ctypes.CDLL(LIB_PATH if real == 'f32' else LIB_PATH_F64)
    handle.oracle_num_threads.restype = ctypes.c_int
    handle.oracle_real_size.restype = ctypes.c_int
    assert handle.oracle_real_size() == (4 if real == 'f32' else 8)
    _libs[real] = handle
  return _libs[real]


def num_threads():
  return int(lib().oracle_num_threads())


def set_threads(n, real='f32'):
  lib(real).oracle_set_threads(ctypes.c_int(int(n)))


def _p(a):
  return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def lattice_loss_and_grads(blank, lexical, num_frames, labels, num_labels, vocab_size,
                           context_size, max_expansions=-1, with_grads=True, real='f32'):
  """loss = logZ - numerator and d sum(loss) / d (blank, lexical), computed in float32
  (real='f32', the arithmetic of the reference) or float64 (real='f64', the truth the
  parity tables are measured against)."""
  dt = np.float32 if real == 'f32' else np.float64
  blank = np.ascontiguousarray(blank, dt)
  lexical = np.ascontiguousarray(lexical, dt)
  b, t, c = blank.shape
  nf = np.ascontiguousarray(num_frames, np.int32)
  lab = np.ascontiguousarray(labels, np.int32)
  nl = np.ascontiguousarray(num_labels, np.int32)
  u = lab.shape[1]
  log_z = np.empty([b], dt)
  num = np.empty([b], dt)
  alphas = np.empty([b, t, c], dt)
  gb = np.empty_like(blank) if with_grads else None
  gl = np.empty_like(lexical) if with_grads else None
  L = lib(real)
  neg_one = ctypes.c_float(-1.0) if real == 'f32' else ctypes.c_double(-1.0)
  L.oracle_lattice_log(ctypes.c_int(vocab_size), ctypes.c_int(context_size),
                       ctypes.c_int(max_expansions), _p(blank), _p(lexical), _p(nf),
                       ctypes.c_int(b), ctypes.c_int(t), None, _p(log_z), _p(alphas), _p(gb),
                       _p(gl))
  L.oracle_string_log(ctypes.c_int(vocab_size), ctypes.c_int(context_size),
                      ctypes.c_int(max_expansions), _p(blank), _p(lexical), _p(nf), _p(lab),
                      _p(nl), ctypes.c_int(b), ctypes.c_int(t), ctypes.c_int(u), None,
                      neg_one, _p(num), _p(gb), _p(gl))
  with np.errstate(all='ignore'):
    loss = log_z - num
  return loss, gb, gl, log_z, alphas
