"""mbgc_b200 -- B200-native (sm_100a) sparse k-mer exact-match finder for MBGC archives.

The product is the shared library libmbgc_b200.so (hand-written CUDA kernels behind the C ABI declared in
include/mbgc_b200.h).  This package is only its Python face: `Matcher` mirrors the reference class
SlidingWindowSparseEMMatcher (/root/reference/matching/SlidingWindowSparseEMMatcher.h:86-126) method for
method through ctypes.  There is no CPU fallback: importing works anywhere, but creating a Matcher needs the
built library and a CUDA device and fails loudly otherwise.
"""
from .matcher import Matcher, MatcherError, NO_LOCK, PackParams, SeqConfig, lib_path, load_library  # noqa: F401

__all__ = ["Matcher", "MatcherError", "NO_LOCK", "PackParams", "SeqConfig", "lib_path", "load_library"]
