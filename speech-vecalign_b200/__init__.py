"""speech-vecalign_b200 — B200 (sm_100a) implementation of Speech-Vecalign's segment-alignment hot
path behind the reference's own Python entry points:

  vecalign.align(...)            <- svecalign/vecalign/vecalign.py:198-223   (called by seg_align/align.py:208)
  dp_utils.vecalign(...)         <- svecalign/vecalign/dp_utils.py:381-390
  dp_core.{make_dense_costs, dense_dp, score_path, make_sparse_costs, sparse_dp, make_x_y_offsets}
                                 <- svecalign/vecalign/dp_core.pyx

The directory name carries a hyphen (repo layout); import it as ``speech_vecalign_b200`` (alias
package at the repo root) or via ``importlib.import_module("speech-vecalign_b200")``.
All numerics live in ``libsvx.so`` (csrc/, C ABI in include/svx.h); without it, or without a CUDA
device, calls raise — there is no CPU fallback.
"""
from . import capi, dp_core, dp_utils, vecalign  # noqa: F401  (same module names as svecalign.vecalign.*)
from .dp_utils import vecalign_batch  # noqa: F401
from .vecalign import align, make_alignment_types, make_many_to_one_alignment_types, print_alignments  # noqa: F401

__all__ = ["capi", "dp_core", "dp_utils", "vecalign", "vecalign_batch", "align", "make_alignment_types",
           "make_many_to_one_alignment_types", "print_alignments"]
