"""Test-only access to the verification build of the library (tests/cuda/libgcpb200_verify.so, compiled with
-DGCPB200_VERIFY by `make -C video_gcp_b200/csrc` / `__graft_entry__.build()`): the same sources plus the SIMT
cross-check kernels (gemm_ref_kernel, dec_tail_ref_kernel, dec_tail_nll_kernel) and the GCPB200_* environment
switches.  The shipped libgcpb200.so contains neither; the product package never loads this file."""
import os

from video_gcp_b200 import _C
from video_gcp_b200.engine import Engine

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda", "libgcpb200_verify.so")
_LIB = None


def verify_lib():
    global _LIB
    if _LIB is None:
        _LIB = _C.bind(PATH)
    return _LIB


def verify_engine(device, simt=True, **kw):
    """Engine on the verification build; simt=True runs every GEMM / decoder tail on the SIMT cross-check kernels."""
    return Engine(device, lib=verify_lib(), reserved0=int(simt), **kw)
