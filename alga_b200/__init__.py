"""alga_b200 -- B200-native overlap-graph construction behind ALGA's GraphCreator interface.

Only the hot path of the reference (``GraphCreatorPrefSuf`` and the ``AlignmentControllers``
verification) lives here.  The compute path is hand-written CUDA for sm_100a in
``alga_b200/csrc`` exported through the C ABI of ``include/alga_gpu.h``; there is no CPU fallback.
"""
from .readset import ReadSet, from_code_list, from_code_matrix  # noqa: F401

__version__ = "0.1.0"
