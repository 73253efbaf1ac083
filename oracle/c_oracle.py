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
_lib = None


def available():
  return os.path.exists(LIB_PATH)


def lib():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(LIB_PATH)
    _lib.oracle_num_threads.restype = ctypes.c_int
    _lib.oracle_real_size.restype = ctypes.c_int
    assert _lib.oracle_real_size() == 4
  return _lib


def num_threads():
  return int(lib().oracle_num_threads())


def set_threads(n):
  lib().oracle_set_threads(ctypes.c_int(int(n)))


def _p(a):
  return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def lattice_loss_and_grads(blank, lexical, num_frames, labels, num_labels, vocab_size,
                           context_size, max_expansions=-1, with_grads=True):
  """loss = logZ - numerator and d sum(loss) / d (blank, lexical), float32."""
  blank = np.ascontiguousarray(blank, np.float32)
  lexical = np.ascontiguousarray(lexical, np.float32)
  b, t, c = blank.shape
  nf = np.ascontiguousarray(num_frames, np.int32)
  lab = np.ascontiguousarray(labels, np.int32)
  nl = np.ascontiguousarray(num_labels, np.int32)
  u = lab.shape[1]
  log_z = np.empty([b], np.float32)
  num = np.empty([b], np.float32)
  alphas = np.empty([b, t, c], np.float32)
  gb = np.empty_like(blank) if with_grads else None
  gl = np.empty_like(lexical) if with_grads else None
  L = lib()
  L.oracle_lattice_log(ctypes.c_int(vocab_size), ctypes.c_int(context_size),
                       ctypes.c_int(max_expansions), _p(blank), _p(lexical), _p(nf),
                       ctypes.c_int(b), ctypes.c_int(t), None, _p(log_z), _p(alphas), _p(gb),
                       _p(gl))
  L.oracle_string_log(ctypes.c_int(vocab_size), ctypes.c_int(context_size),
                      ctypes.c_int(max_expansions), _p(blank), _p(lexical), _p(nf), _p(lab),
                      _p(nl), ctypes.c_int(b), ctypes.c_int(t), ctypes.c_int(u), None,
                      ctypes.c_float(-1.0), _p(num), _p(gb), _p(gl))
  with np.errstate(all='ignore'):
    loss = log_z - num
  return loss, gb, gl, log_z, alphas
