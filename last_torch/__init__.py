"""`last_torch` import alias of last_torch_b200 (the drop-in name, reference __init__.py:18-22):

    import last_torch
    last_torch.{alignments, contexts, semirings, weight_fns, RecognitionLattice}

Code written against the reference imports this package unchanged; everything resolves to the
B200 implementation in last_torch_b200/ (this file holds no logic).
"""
import sys as _sys

import last_torch_b200 as _impl
from last_torch_b200 import alignments, contexts, lattices, semirings, weight_fns
from last_torch_b200.lattices import RecognitionLattice

for _name in ('alignments', 'contexts', 'lattices', 'semirings', 'weight_fns'):
  _sys.modules[f'{__name__}.{_name}'] = getattr(_impl, _name)

__all__ = ['alignments', 'contexts', 'semirings', 'weight_fns', 'RecognitionLattice']
