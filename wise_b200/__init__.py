"""wise_b200 - a B200-native (sm_100a) search-index backend for WISE.

Replaces the faiss CPU indices behind /root/reference/src/index/feature_search_index.py with
hand-written CUDA kernels reached through a C-ABI (include/wise_b200.h).  Use
`wise_b200.faiss_compat` wherever the reference says `import faiss`.
"""
from . import faiss_compat  # noqa: F401

__all__ = ["faiss_compat"]
__version__ = "0.1"
