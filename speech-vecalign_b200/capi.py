"""ctypes binding of ``libsvx.so`` (C ABI: ``include/svx.h``).

The library is the product: there is no CPU fallback.  Importing this module on a machine
without the built extension raises; calling a launcher without a CUDA device raises.

Job descriptors are numpy structured arrays whose layout mirrors the C structs (checked against
``svx_sizeof_job`` at load time); the launchers take the device copy and the host copy.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvx.so")
CSRC = os.path.join(_HERE, "csrc")

SVX_COST_EXACT, SVX_COST_FAST, SVX_COST_TC = 0, 1, 2
SVX_ST_LEFT_BAND, SVX_ST_NO_BACKPTR, SVX_ST_OVERFLOW = 1, 2, 4
SVX_BP_NONE = 255
SVX_MAX_TYPES = 126

_P = np.uint64  # device pointers travel as 64-bit integers

ROWS = np.dtype([("ptr", _P), ("nrows", np.int64)], align=True)
DOWN = np.dtype([("in", _P), ("out", _P), ("mean", _P), ("k", np.int32), ("n", np.int32)], align=True)
NORM = np.dtype([("vecs", _P), ("other", _P), ("idx", _P), ("mbar", _P), ("norms", _P),
                 ("k", np.int32), ("n", np.int32), ("ko", np.int32), ("no", np.int32), ("per", np.int32)],
                align=True)
ROW_SOURCE = np.dtype([("rows", _P), ("table", _P), ("nan_rows", _P), ("nrows", np.int32), ("is_fp16", np.int32)], align=True)
LEVEL = np.dtype([("vecs", _P), ("mean", _P), ("next", _P), ("other", _P), ("other_mean", _P), ("idx", _P),
                  ("mbar", _P), ("norms", _P),
                  ("k", np.int32), ("n", np.int32), ("ko", np.int32), ("no", np.int32), ("per", np.int32),
                  ("keep", np.int32), ("src", ROW_SOURCE), ("osrc", ROW_SOURCE)], align=True)
GATHER = np.dtype([("rows", _P), ("table", _P), ("out", _P), ("nan_rows", _P),
                   ("k", np.int32), ("n", np.int32), ("nrows", np.int32), ("is_fp16", np.int32)], align=True)
SCORE = np.dtype([("e", _P), ("f", _P), ("norm_e", _P), ("norm_f", _P), ("xi", _P), ("yi", _P),
                  ("scores", _P), ("del_penalty", _P), ("perm", _P), ("dots", _P),
                  ("ne", np.int32), ("nf", np.int32), ("nsamp", np.int32)], align=True)
DENSE = np.dtype([("v0", _P), ("v1", _P), ("n0", _P), ("n1", _P), ("costs", _P), ("dots", _P), ("tmap0", _P), ("tmap1", _P), ("lo0", _P), ("lo1", _P), ("del_penalty", _P),
                  ("bp", _P), ("csum", _P), ("ypath", _P), ("status_d", _P),
                  ("s0", np.int32), ("s1", np.int32), ("t0", np.int32), ("t1", np.int32),
                  ("upsample", np.int32), ("path_len", np.int32)], align=True)
REC = np.dtype([("x_end", np.int32), ("y_end", np.int32), ("nx", np.int32), ("ny", np.int32),
                ("score", np.float64)], align=True)
BAND = np.dtype([("v0", _P), ("v1", _P), ("n0", _P), ("n1", _P), ("ypath", _P), ("costs", _P),
                 ("del_penalty", _P), ("bp", _P), ("csum", _P), ("recs", _P), ("nrecs", _P),
                 ("next_ypath", _P), ("status_d", _P),
                 ("s0", np.int32), ("s1", np.int32), ("k0", np.int32), ("k1", np.int32),
                 ("a_len", np.int32), ("band", np.int32), ("width_over2", np.int32), ("ntypes", np.int32),
                 ("rec_cap", np.int32), ("t0", np.int32), ("t1", np.int32), ("next_len", np.int32),
                 ("xo", np.int8, (SVX_MAX_TYPES,)), ("yo", np.int8, (SVX_MAX_TYPES,)),
                 ("amax", np.int16)], align=True)

PARAMS = np.dtype([("k0", np.int32), ("k1", np.int32), ("dim", np.int32), ("ntypes", np.int32),
                   ("xo", np.int8, (SVX_MAX_TYPES,)), ("yo", np.int8, (SVX_MAX_TYPES,)),
                   ("del_percentile_frac", np.float64),
                   ("width_over2", np.int32), ("max_size_full_dp", np.int32), ("costs_sample_size", np.int32),
                   ("num_samps_for_norm", np.int32), ("cost_mode", np.int32), ("keep_all", np.int32),
                   ("unfused_prologue", np.int32), ("skip_norms0", np.int32), ("skip_norms1", np.int32)], align=True)
PLAN_INFO = np.dtype([("npairs", np.int32), ("nrecords", np.int32), ("max_depth", np.int32), ("band", np.int32),
                      ("width_over2", np.int32), ("per0", np.int32), ("per1", np.int32), ("fused_prologue", np.int32),
                      ("nlaunchers", np.int32), ("ndraw_calls", np.int64), ("arena_bytes", np.int64), ("host_bytes", np.int64),
                      ("result_offset", np.int64), ("result_bytes", np.int64), ("counts_offset", np.int64),
                      ("fallback_del_penalty", np.float64)], align=True)

_STRUCTS = [ROWS, DOWN, NORM, SCORE, DENSE, BAND, REC, LEVEL, GATHER, PARAMS, PLAN_INFO, ROW_SOURCE]

_lib = None


class SvxError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libsvx.so (csrc/Makefile: nvcc -gencode
    arch=compute_100a,code=sm_100a -lineinfo).  Cross-compiles without a GPU."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise SvxError("libsvx.so build failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


def lib():
    """The loaded C ABI.  Raises if the extension has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvxError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the CUDA extension is the only implementation; there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    sigs = {
        "svx_normalize_rows": [vp, vp, ci, ci, vp],
        "svx_downsample": [vp, vp, ci, ci, vp],
        "svx_sample_norms": [vp, vp, ci, ci, vp],
        "svx_level_prologue": [vp, vp, ci, ci, vp],
        "svx_gather_doc_embedding": [vp, vp, ci, ci, vp],
        "svx_score_pairs": [vp, vp, ci, ci, ci, vp],
        "svx_del_knob": [vp, vp, ci, cd, vp],
        "svx_host_del_knob": [vp, ci, cd, vp],
        "svx_dense_costs": [vp, vp, ci, ci, ci, vp],
        "svx_dense_dp": [vp, vp, ci, vp],
        "svx_dense_tmaps_encode": [vp, ci, ci, vp],
        "svx_path_len": [ci, ci, ci, ci, ci],
        "svx_banded_costs": [vp, vp, ci, ci, ci, vp],
        "svx_banded_dp": [vp, vp, ci, vp],
        "svx_host_banded_dp": [vp],
        "svx_upload_pinned": [vp, vp, ctypes.c_longlong, vp],
        "svx_host_memcpy": [vp, vp, ctypes.c_longlong, ci],
        "svx_host_randint_stream": [vp, vp, ci, vp, vp, vp],
        "svx_host_randint_seeded": [ci, vp, vp, vp, vp, vp, ci],
        "svx_host_dense_dp": [vp],
        "svx_version": [],
        "svx_sizeof_job": [ci],
        "svx_plan_create": [vp, ci, vp, vp, vp],
        "svx_plan_info": [vp, vp],
        "svx_plan_array": [vp, ci, vp, vp],
        "svx_plan_bind": [vp, vp, vp, vp, vp],
        "svx_plan_set_sources": [vp, vp, vp],
        "svx_plan_draw_seeded": [vp, vp, ci],
        "svx_plan_draw_stream": [vp, vp, vp],
        "svx_plan_upload": [vp, ci, vp],
        "svx_plan_restore": [vp, vp, vp],
        "svx_plan_launcher_name": [vp, ci, vp, ci],
        "svx_plan_enqueue": [vp, ci, ci, ci, vp],
        "svx_plan_fetch": [vp, vp, vp, vp, vp, vp, vp],
        "svx_workspace_bytes": [vp, ci, vp, vp, vp, vp],
        "svx_host_overlap_tables": [ci, vp, vp, vp, vp, ci, vp, vp, vp, ci],
        "svx_margin_workspace_bytes": [ci, ci, ci, ci, vp],
        "svx_margin_scores": [vp, vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, vp, vp, ctypes.c_int64, vp],
        "svx_align_batch": [vp, ci, vp, vp, vp, vp, vp, vp, ctypes.c_int64, vp, ctypes.c_int64, ci, vp, vp, vp, vp, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = ci
    L.svx_plan_destroy.argtypes = [vp]
    L.svx_plan_destroy.restype = None
    L.svx_launch_count.argtypes = [ci]
    L.svx_launch_count.restype = ctypes.c_longlong
    L.svx_last_error_string.argtypes = []
    L.svx_last_error_string.restype = ctypes.c_char_p
    for i, dt in enumerate(_STRUCTS):
        c_size = L.svx_sizeof_job(i)
        if c_size != dt.itemsize:
            raise SvxError(f"struct #{i} layout mismatch: C {c_size} B vs numpy {dt.itemsize} B")
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "svx_normalize_rows", "svx_downsample", "svx_sample_norms", "svx_level_prologue", "svx_gather_doc_embedding", "svx_score_pairs", "svx_del_knob",
    "svx_host_del_knob", "svx_dense_costs", "svx_dense_dp", "svx_dense_tmaps_encode", "svx_path_len", "svx_banded_costs",
    "svx_banded_dp", "svx_host_banded_dp", "svx_host_dense_dp", "svx_host_randint_stream",
    "svx_host_randint_seeded", "svx_upload_pinned", "svx_host_memcpy", "svx_version",
    "svx_last_error_string", "svx_sizeof_job", "svx_launch_count",
    "svx_plan_create", "svx_plan_destroy", "svx_plan_info", "svx_plan_array", "svx_plan_bind", "svx_plan_set_sources", "svx_plan_draw_seeded",
    "svx_plan_draw_stream", "svx_plan_upload", "svx_plan_restore", "svx_plan_launcher_name", "svx_plan_enqueue", "svx_plan_fetch",
    "svx_workspace_bytes", "svx_align_batch", "svx_margin_workspace_bytes", "svx_margin_scores",
    "svx_host_overlap_tables",
]


def check(rc: int, what: str):
    if rc != 0:
        raise SvxError(f"{what} failed (code {rc}): {lib().svx_last_error_string().decode()}")


def hptr(a: np.ndarray) -> int:
    return a.ctypes.data
