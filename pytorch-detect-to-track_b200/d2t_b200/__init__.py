"""d2t_b200 -- Blackwell-native (sm_100a) kernels for the Detect-to-Track per-frame-pair hot path.

``d2t_b200.ops`` is the tensor-level front end of the C-ABI library ``libd2t_b200.so``
(include/d2t_b200.h); the sibling ``model`` package mirrors the reference's ``lib/model``
operator API on top of it.  Importing this package does not need a GPU; calling an operator does.
"""
from ._lib import D2TError, D2TLibraryMissing, SO_PATH, build, lib  # noqa: F401

__all__ = ["D2TError", "D2TLibraryMissing", "SO_PATH", "build", "lib"]
