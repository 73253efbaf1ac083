"""last_torch_b200: B200 (sm_100a) kernels behind the last_torch API.

Same import surface as the reference package
(/root/reference/last_torch/__init__.py:18-22):

    import last_torch_b200 as last_torch
    last_torch.{alignments, contexts, semirings, weight_fns, RecognitionLattice}
"""

from . import alignments
from . import contexts
from . import semirings
from . import weight_fns
from .lattices import RecognitionLattice

__all__ = ['alignments', 'contexts', 'semirings', 'weight_fns', 'RecognitionLattice']
