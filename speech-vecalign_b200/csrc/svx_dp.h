// svx_dp.h — DP cell updates, tracebacks and the search-path builder, written once and compiled
// both into the sm_100a kernels and into their host twins (svx_host_*), so that the integer /
// fp64 logic can be checked against the oracle on a machine without a GPU.
#pragma once
#include "../../include/svx.h"
#include "svx_math.h"

// ---------------------------------------------------------------------------------------------
// Search path of the next finer level (dp_utils.py:199-275: upsample_alignment,
// extend_alignments, alignment_to_search_path) built while walking a traceback BACKWARDS.
//
// The reference upsamples every alignment by 2, appends one final block that reaches
// (t0+1, t1+1), then lays unit steps: each run of consecutive deletions becomes one slanted
// segment, each n-m alignment another.  Because point i of a segment sits at path index
// x+y, segments can be emitted in any order; `Emit` expands one segment (lane-parallel on the
// device, serial on the host) and stores ypath[x+y] = y.
// ---------------------------------------------------------------------------------------------
template <class Emit>
struct SvxPathBuilder {
    Emit emit;
    int f;                 // 2 when upsampling, 1 when the target level is the same level
    long long pend_x, pend_y;   // pending deletion run, in target units

    SVX_HD SvxPathBuilder(Emit e) : emit(e), f(1), pend_x(0), pend_y(0) {}

    // c0,c1: coarse sizes (the traceback starts at node (c0,c1)); t0,t1: target sizes.
    SVX_HD void begin(int c0, int c1, int t0, int t1, int upsample)
    {
        pend_x = pend_y = 0;
        f = upsample ? 2 : 1;
        if (!upsample) return;
        // extend_alignments: xmax/ymax start at 0 and only grow over the upsampled ids
        const int xmax = c0 > 0 ? 2 * c0 - 1 : 0;
        const int ymax = c1 > 0 ? 2 * c1 - 1 : 0;
        const int lenx = t0 - xmax > 0 ? t0 - xmax : 0;   // len(range(xmax+1, t0+1))
        const int leny = t1 - ymax > 0 ? t1 - ymax : 0;
        if (lenx > 0 && leny > 0) emit(2LL * c0, 2LL * c1, (long long)lenx, (long long)leny);
        else if (lenx == 0) pend_y = leny;     // appended as single deletions -> joins the run
        else pend_x = lenx;
    }

    // one alignment ending at coarse node (x_end, y_end) consuming (nx, ny)
    SVX_HD void step(int x_end, int y_end, int nx, int ny)
    {
        if (nx > 0 && ny > 0) {
            if (pend_x | pend_y) emit((long long)f * x_end, (long long)f * y_end, pend_x, pend_y);
            pend_x = pend_y = 0;
            emit((long long)f * (x_end - nx), (long long)f * (y_end - ny), (long long)f * nx, (long long)f * ny);
        } else if (nx > 0) {
            pend_x += (long long)f * nx;
        } else {
            pend_y += (long long)f * ny;
        }
    }

    SVX_HD void finish()
    {
        if (pend_x | pend_y) emit(0LL, 0LL, pend_x, pend_y);
        pend_x = pend_y = 0;
    }
};

SVX_HD int svx_path_len_impl(int c0, int c1, int t0, int t1, int upsample)
{
    if (!upsample) return 1 + c0 + c1;
    const int xmax = c0 > 0 ? 2 * c0 - 1 : 0;
    const int ymax = c1 > 0 ? 2 * c1 - 1 : 0;
    const int lenx = t0 - xmax > 0 ? t0 - xmax : 0;
    const int leny = t1 - ymax > 0 ? t1 - ymax : 0;
    return 1 + 2 * c0 + lenx + 2 * c1 + leny;
}

// Serial segment expansion (host twin; also used by single-lane device code).
struct SvxSerialEmit {
    int32_t *ypath;
    int path_len;
    SVX_HD void operator()(long long xs, long long ys, long long xw, long long yw) const
    {
        const long long nn = xw + yw;
        for (long long i = 1; i <= nn; ++i) {
            int x, y;
            svx_slant_point(xs, ys, xw, yw, i, &x, &y);
            if (x + y < path_len) ypath[x + y] = y;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Dense 3-way DP (dp_core.pyx:79-141).  `penf` is the penalty narrowed to fp32 (dense_dp takes a
// C float); boundary values are int*float products in fp32 widened to double.
// ---------------------------------------------------------------------------------------------
SVX_HD double svx_dense_boundary(int k, float penf) { return (double)SVX_FMUL((float)k, penf); }

// interior node: diag = csum[r-1,c-1], left = csum[r,c-1], up = csum[r-1,c]
SVX_HD double svx_dense_cell(double diag, double left, double up, float cost, float penf, int *bp)
{
    const double c0 = SVX_DADD(diag, (double)cost);
    const double c1 = SVX_DADD(left, (double)penf);
    const double c2 = SVX_DADD(up, (double)penf);
    double best = c0; int b = 0;
    if (c1 < best) { best = c1; b = 1; }
    if (c2 < best) { best = c2; b = 2; }
    *bp = b;
    return best;
}

// dense traceback (dp_utils.py:146-174) feeding a path builder.  bp is (s0+1, s1+1) row-major.
// Returns SVX_ST_* status.
template <class Builder>
SVX_HD int svx_dense_walk(const uint8_t *bp, int s0, int s1, Builder &pb)
{
    int x = s0, y = s1;
    const int ld = s1 + 1;
    long long guard = (long long)s0 + s1 + 2;
    while (!(x == 0 && y == 0)) {
        if (--guard < 0) return SVX_ST_NO_BACKPTR;
        const int code = bp[(size_t)x * ld + y];
        if (code == 0) { pb.step(x, y, 1, 1); --x; --y; }
        else if (code == 1) { pb.step(x, y, 0, 1); --y; }
        else if (code == 2) { pb.step(x, y, 1, 0); --x; }
        else return SVX_ST_NO_BACKPTR;           // reference: 'got unknown value'
        if (x < 0 || y < 0) return SVX_ST_LEFT_BAND;
    }
    pb.finish();
    return SVX_ST_OK;
}

// ---------------------------------------------------------------------------------------------
// Banded DP node (dp_core.pyx:269-404 sparse_dp).
//
// Node (aa, bb) of the (A+2, B) lattice: yy = bb + boff(aa), xx = aa - yy, where
// boff(a) = b_offset_out[a] = ypath-derived: boff(0) = boff(1) = ypath[0]-w, boff(a) = ypath[a-2]-w+1.
// The cost cell of every type ending at this node is (xx-1, yy-1): anti-diagonal aa-2, band slot
// bb (because boff_out[a] = boff_in[a-2] + 1), so it is in the band whenever it is in the document.
// A candidate (xo,yo) needs that cost cell to exist (even for deletions — reference quirk,
// dp_core.pyx:382,390) and its predecessor (xx-xo, yy-yo) to be inside the node band.
//   Boff(a)            -> b_offset_out[a]
//   CsumAt(a, b)       -> csum of an earlier node
//   CostAt(t)          -> costs[aa-2][t][bb]
// Types: t < ntypes -> (xo[t], yo[t]); t == ntypes -> (0,1); t == ntypes+1 -> (1,0).
// ---------------------------------------------------------------------------------------------
template <class Boff, class CsumAt, class CostAt>
SVX_HD double svx_band_node(int aa, int bb, int s0, int s1, int a_len, int band, int ntypes,
                            const int8_t *xo, const int8_t *yo, double pen,
                            Boff boff, CsumAt csum_at, CostAt cost_at, int *bp_out)
{
    const int yy = bb + boff(aa);
    const int xx = aa - yy;
    if (xx == 0 && 0 <= yy && yy < s1 + 1) { *bp_out = ntypes; return SVX_DMUL(pen, (double)yy); }
    if (yy == 0 && 0 <= xx && xx < s0 + 1) { *bp_out = ntypes + 1; return SVX_DMUL(pen, (double)xx); }
    double best = INFINITY;
    int bcode = SVX_BP_NONE;
    const int xc = xx - 1, yc = yy - 1;
    if (0 <= xc && xc < s0 && 0 <= yc && yc < s1 && aa - 2 < a_len) {
        for (int t = 0; t < ntypes + 2; ++t) {
            const int dx = t < ntypes ? xo[t] : (t == ntypes ? 0 : 1);
            const int dy = t < ntypes ? yo[t] : (t == ntypes ? 1 : 0);
            const int xq = xx - dx, yq = yy - dy;
            if (xq < 0 || yq < 0) continue;           // upper bounds hold since xx <= s0, yy <= s1
            const int aq = xq + yq;
            const int bq = yq - boff(aq);
            if (bq < 0 || bq >= band) continue;
            const double stepc = (t < ntypes) ? (double)cost_at(t) : pen;
            const double total = SVX_DADD(csum_at(aq, bq), stepc);
            if (total < best) { best = total; bcode = t; }
        }
    }
    *bp_out = bcode;
    return best;
}

// b_offset_out from the search path (dp_core.pyx:241-243, 327-328)
SVX_HD int svx_boff_out(const int32_t *ypath, int a, int w)
{
    return a < 2 ? ypath[0] - w : ypath[a - 2] - w + 1;
}

// Banded traceback (dp_utils.py:105-143) + process_scores (:89-102).  Walks from (s0,s1) to
// (0,0); calls on_align(x_end, y_end, nx, ny, score) for every alignment, last one first.
template <class OnAlign>
SVX_HD int svx_band_walk(const uint8_t *bp, const double *csum, const int32_t *ypath, int w,
                         int s0, int s1, int a_len, int band, int ntypes,
                         const int8_t *xo, const int8_t *yo, OnAlign &on_align)
{
    int x = s0, y = s1;
    long long guard = (long long)s0 + s1 + 2;
    int a = x + y;
    if (a >= a_len + 2) return SVX_ST_LEFT_BAND;
    int b = y - svx_boff_out(ypath, a, w);
    if (b < 0 || b >= band) return SVX_ST_LEFT_BAND;
    double here = csum[(size_t)a * band + b];
    while (!(x == 0 && y == 0)) {
        if (--guard < 0) return SVX_ST_NO_BACKPTR;
        const int code = bp[(size_t)a * band + b];
        if (code == SVX_BP_NONE || code > ntypes + 1) return SVX_ST_NO_BACKPTR;
        const int dx = code < ntypes ? xo[code] : (code == ntypes ? 0 : 1);
        const int dy = code < ntypes ? yo[code] : (code == ntypes ? 1 : 0);
        const int px = x - dx, py = y - dy;
        if (px < 0 || py < 0) return SVX_ST_LEFT_BAND;          // reference: 'traceback bug'
        const int pa = px + py;
        const int pb = py - svx_boff_out(ypath, pa, w);
        if (pb < 0 || pb >= band) return SVX_ST_LEFT_BAND;
        const double prev = csum[(size_t)pa * band + pb];
        double sc = SVX_DSUB(here, prev);            // np.diff of the cumulative costs
        if (sc < 0.0) sc = 0.0;                      // np.clip(a_min=0); NaN stays NaN
        if (dx == 0 || dy == 0) sc = 0.0;
        else sc = SVX_DDIV(SVX_DDIV(sc, (double)dx), (double)dy);
        on_align(x, y, dx, dy, sc);
        x = px; y = py; a = pa; b = pb; here = prev;
    }
    return SVX_ST_OK;
}
