"""phi_b200 — B200-native front end (ILP_index stage) of PHI, behind a C ABI.

The product is phi_b200/libphi_gpu_index.so (hand-written sm_100a CUDA + C++ host code; see
include/phi_gpu_index.h).  This package is only the thin ctypes loader that tests, bench.py and
Python callers use; it never computes anything itself and has NO CPU fallback: if the library or a
GPU is missing, calls raise.
"""
from ._abi import Graph, Reads, IndexResultPy, PHI_OK  # noqa: F401
from .api import PhiGpuIndex, PhiGpuError, load_library, library_path, load_gfa, load_reads  # noqa: F401

__all__ = ["Graph", "Reads", "IndexResultPy", "PhiGpuIndex", "PhiGpuError", "load_library", "library_path", "load_gfa", "load_reads"]
