"""cqs_b200 — B200-native exact retrieval backend for cqs (host-side mirror).

The product is ``libcqs_b200.so`` (hand-written sm_100a kernels behind the C ABI
in ``include/cqs_b200.h``).  This package is the Python mirror of the
reference's Rust interface for the path (``VectorIndex``, ``SpladeIndex``,
``search_hybrid``, router alpha table) used by the tests and the benchmark; it
contains no arithmetic of its own and has no CPU fallback.
"""
from .capi import lib, B200Error, load_library  # noqa: F401
from .index import B200Index, IndexResult  # noqa: F401
from .router import (CATEGORIES, DEFAULT_ALPHA, resolve_splade_alpha,  # noqa: F401
                     CentroidClassifier, apply_centroid_floor)
from .hybrid import SpladeIndex, search_hybrid, candidate_count_for, cap_k_to_backend  # noqa: F401
