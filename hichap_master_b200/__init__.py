"""hichap_master_b200 -- B200-native (sm_100a) implementation of the hot path of HiCHap's
``matrix`` stage: valid-pair binning, ICE balancing and the two-step allelic correction.

``hichap_master_b200.matrixBuilding`` mirrors ``HiCHap/matrixBuilding.py``; the compute lives
in ``libhichap_b200.so`` (C ABI in ``include/hichap_b200.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _abi  # noqa: F401


def library_path() -> str:
    return _abi.LIB_PATH
