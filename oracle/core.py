"""oracle/core.py — TEST INFRASTRUCTURE (checker), not product code.

ctypes binding of ``liboracle.so`` (``dp_core_oracle.c``), exposing the same six callables, with
the same argument order, dtypes and return shapes, as the reference's native module
``svecalign/vecalign/dp_core.pyx`` (make_x_y_offsets :24-34, make_dense_costs :36-77,
dense_dp :79-141, score_path :143-161, make_sparse_costs :165-267, sparse_dp :269-404), so that
``oracle.vecalign_oracle`` can run on either this port or the reference's own compiled core
(``oracle/_ref``) interchangeably.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (oracle/Makefile).  Returns its path."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "dp_core_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a, ty):
    return a.ctypes.data_as(ty)


def _f32(a, ndim):
    a = np.ascontiguousarray(a)
    if a.dtype != np.float32 or a.ndim != ndim:
        raise ValueError("Buffer dtype mismatch, expected 'float' with ndim=%d" % ndim)
    return a


def make_x_y_offsets(alignment_types):
    for x, y in alignment_types:
        assert x > 0
        assert y > 0
    xo = np.array([x for x, _ in alignment_types], dtype=np.int32)
    yo = np.array([y for _, y in alignment_types], dtype=np.int32)
    return xo, yo


def make_dense_costs(vecs0, vecs1, norm0, norm1, offset0=0, offset1=0):
    vecs0, vecs1 = _f32(vecs0, 3), _f32(vecs1, 3)
    norm0, norm1 = _f32(norm0, 2), _f32(norm1, 2)
    assert vecs0.shape[0] > offset0 and vecs1.shape[0] > offset1
    assert norm0.shape[0] > offset0 and norm1.shape[0] > offset1
    s0, s1, d = vecs0.shape[1], vecs1.shape[1], vecs0.shape[2]
    assert norm0.shape[1] == s0 and norm1.shape[1] == s1 and vecs1.shape[2] == d
    out = np.empty((s0, s1), dtype=np.float32)
    _lib().svo_make_dense_costs(_p(vecs0, _f32p), _p(vecs1, _f32p), _p(norm0, _f32p), _p(norm1, _f32p),
                                s0, s1, d, int(offset0), int(offset1), _p(out, _f32p))
    return out


def dense_dp(alignment_cost, pen):
    cost = _f32(alignment_cost, 2)
    s0, s1 = cost.shape
    csum = np.empty((s0 + 1, s1 + 1), dtype=np.float64)
    bp = np.empty((s0 + 1, s1 + 1), dtype=np.int32)
    _lib().svo_dense_dp(_p(cost, _f32p), s0, s1, ctypes.c_float(float(pen)), _p(csum, _f64p), _p(bp, _i32p))
    return csum, bp


def score_path(xx, yy, norm1, norm2, vecs1, vecs2, out):
    for a in (xx, yy):
        if a.dtype != np.int32:
            raise ValueError("Buffer dtype mismatch, expected 'int'")
    xx, yy = np.ascontiguousarray(xx), np.ascontiguousarray(yy)
    norm1, norm2 = _f32(norm1, 1), _f32(norm2, 1)
    vecs1, vecs2 = _f32(vecs1, 2), _f32(vecs2, 2)
    assert out.dtype == np.float32 and out.flags.c_contiguous
    _lib().svo_score_path(_p(xx, _i32p), _p(yy, _i32p), int(xx.shape[0]), _p(norm1, _f32p), _p(norm2, _f32p),
                          _p(vecs1, _f32p), _p(vecs2, _f32p), int(vecs1.shape[1]), _p(out, _f32p))


def make_sparse_costs(vecs0, vecs1, norms0, norms1, x_y_path, alignment_types, width_over2):
    vecs0, vecs1 = _f32(vecs0, 3), _f32(vecs1, 3)
    norms0, norms1 = _f32(norms0, 2), _f32(norms1, 2)
    path = np.ascontiguousarray(np.array(x_y_path).astype(np.int32))
    assert vecs0.shape[0] == norms0.shape[0] and vecs1.shape[0] == norms1.shape[0]
    assert vecs0.shape[1] == norms0.shape[1] and vecs1.shape[1] == norms1.shape[1]
    mx = max([0] + [x for x, _ in alignment_types])
    my = max([0] + [y for _, y in alignment_types])
    if mx > vecs0.shape[0]:
        raise Exception('%d x overlaps requrested (via alignment_types), but vecs0 only has %d'
                        % (mx, vecs0.shape[0]))
    if my > vecs1.shape[0]:
        raise Exception('%d y overlaps requrested (via alignment_types), but vecs1 only has %d'
                        % (my, vecs1.shape[0]))
    assert vecs0.shape[2] == vecs1.shape[2]
    xo, yo = make_x_y_offsets(alignment_types)
    a_len, b_len = path.shape[0], 2 * int(width_over2)
    feats = np.empty((len(alignment_types), a_len, b_len), dtype=np.float32)
    boff = np.empty(a_len, dtype=np.int32)
    _lib().svo_make_sparse_costs(_p(vecs0, _f32p), _p(vecs1, _f32p), _p(norms0, _f32p), _p(norms1, _f32p),
                                 int(vecs0.shape[1]), int(vecs1.shape[1]), int(vecs0.shape[2]),
                                 _p(path, _i32p), a_len, _p(xo, _i32p), _p(yo, _i32p), len(alignment_types),
                                 int(width_over2), _p(feats, _f32p), _p(boff, _i32p))
    return feats, boff


def sparse_dp(a_b_costs, b_offset_in, alignment_types, del_penalty, x_in_size, y_in_size):
    costs = _f32(a_b_costs, 3)
    boff = np.ascontiguousarray(b_offset_in)
    assert boff.dtype == np.int32
    xo, yo = make_x_y_offsets(alignment_types)
    t, a_in, b_in = costs.shape
    csum = np.empty((a_in + 2, b_in), dtype=np.float64)
    xp = np.empty((a_in + 2, b_in), dtype=np.int32)
    yp = np.empty((a_in + 2, b_in), dtype=np.int32)
    boff_out = np.empty(a_in + 2, dtype=np.int32)
    _lib().svo_sparse_dp(_p(costs, _f32p), _p(boff, _i32p), _p(xo, _i32p), _p(yo, _i32p), int(t),
                         int(a_in), int(b_in), ctypes.c_double(float(del_penalty)),
                         int(x_in_size), int(y_in_size),
                         _p(csum, _f64p), _p(xp, _i32p), _p(yp, _i32p), _p(boff_out, _i32p))
    return csum, xp, yp, boff_out
