/*
 * oracle/dp_core_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the arithmetic of the reference's only native module,
 * svecalign/vecalign/dp_core.pyx (Cython), written from its documented behaviour so that
 * the CUDA path can be checked against it on machines where /root/reference does not exist.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.  The product package never does.
 *
 * Parity pin: tests/test_oracle_vs_reference.py compares every function below bit-for-bit with
 * the reference's own compiled dp_core (oracle/_ref/, built by oracle/Makefile from the
 * .pyx where it lies) whenever /root/reference is present, and tests/golden/ holds outputs
 * of the real reference for the GPU box.
 *
 * Compile WITHOUT -ffast-math and WITHOUT FMA contraction (-ffp-contract=off): the reference
 * is built by pyximport with plain -O2 on x86-64, so every fp32 product is rounded before it
 * is added.
 *
 * All arrays are C-contiguous.  vecs: (K, n, D) float32.  norms: (K, n) float32.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define SVO_API __attribute__((visibility("default")))

/* Sequential fp32 multiply-then-add dot product: dp_core.pyx:69-71, 157-159, 256-258. */
static inline float seq_dot(const float *a, const float *b, int d)
{
    float s = 0.0f;
    for (int j = 0; j < d; ++j)
        s += a[j] * b[j];
    return s;
}

/*
 * dp_core.pyx:36-77 make_dense_costs.
 * cost[x,y] = 2(1 - v0[o0,x].v1[o1,y]) / (1e-6 + n0[o0,x] + n1[o1,y]); quotient evaluated in
 * double and narrowed to fp32 (:73), then multiplied in fp32 by (o0+1) and (o1+1) (:75).
 */
SVO_API void svo_make_dense_costs(const float *vecs0, const float *vecs1,
                                  const float *norm0, const float *norm1,
                                  int size0, int size1, int dim, int off0, int off1,
                                  float *costs /* (size0,size1) */)
{
    const float *v0 = vecs0 + (size_t)off0 * size0 * dim;
    const float *v1 = vecs1 + (size_t)off1 * size1 * dim;
    const float *n0 = norm0 + (size_t)off0 * size0;
    const float *n1 = norm1 + (size_t)off1 * size1;
    for (int x = 0; x < size0; ++x) {
        for (int y = 0; y < size1; ++y) {
            float sumx = seq_dot(v0 + (size_t)x * dim, v1 + (size_t)y * dim, dim);
            double num = 2.0 * (1.0 - (double)sumx);
            double den = (1e-6 + (double)n0[x]) + (double)n1[y];
            float c = (float)(num / den);
            c = (c * (float)(off0 + 1)) * (float)(off1 + 1);
            costs[(size_t)x * size1 + y] = c;
        }
    }
}

/*
 * dp_core.pyx:79-141 dense_dp.  `pen` arrives as a C float (:79).  Nodes are
 * (size0+1) x (size1+1); bp: 0 = diagonal, 1 = came from (r, c-1), 2 = from (r-1, c), 4 at origin.
 * Row/column initialisation is an int*float product in fp32 widened to double (:110-117).
 * Candidate order and strict '<' per :126-139.
 */
SVO_API void svo_dense_dp(const float *cost, int size0, int size1, float pen,
                          double *csum, int32_t *bp)
{
    const int rmax = size0 + 1, cmax = size1 + 1;
#define CS(r, c) csum[(size_t)(r) * cmax + (c)]
#define BP(r, c) bp[(size_t)(r) * cmax + (c)]
    for (int c = 0; c < cmax; ++c) { CS(0, c) = (double)((float)c * pen); BP(0, c) = 1; }
    for (int r = 0; r < rmax; ++r) { CS(r, 0) = (double)((float)r * pen); BP(r, 0) = 2; }
    CS(0, 0) = 0.0;
    BP(0, 0) = 4;
    for (int c = 1; c < cmax; ++c) {
        for (int r = 1; r < rmax; ++r) {
            double c0 = CS(r - 1, c - 1) + (double)cost[(size_t)(r - 1) * size1 + (c - 1)];
            double c1 = CS(r, c - 1) + (double)pen;
            double c2 = CS(r - 1, c) + (double)pen;
            double best = c0; int32_t b = 0;
            if (c1 < best) { best = c1; b = 1; }
            if (c2 < best) { best = c2; b = 2; }
            CS(r, c) = best; BP(r, c) = b;
        }
    }
#undef CS
#undef BP
}

/*
 * dp_core.pyx:143-161 score_path.  The denominator n1[x]+n2[y] is an fp32 addition (both
 * operands are C floats) with no 1e-6 term; the quotient is taken in double and narrowed.
 */
SVO_API void svo_score_path(const int32_t *xx, const int32_t *yy, int n,
                            const float *norm1, const float *norm2,
                            const float *vecs1, const float *vecs2, int dim, float *out)
{
    for (int i = 0; i < n; ++i) {
        int xi = xx[i], yi = yy[i];
        float dot = seq_dot(vecs1 + (size_t)xi * dim, vecs2 + (size_t)yi * dim, dim);
        float den = norm1[xi] + norm2[yi];
        out[i] = (float)((2.0 * (1.0 - (double)dot)) / (double)den);
    }
}

/*
 * dp_core.pyx:165-267 make_sparse_costs.  path: (A,2) int32 (x,y); every step advances
 * x+y by one so point ii lands on anti-diagonal aa = x+y (:236-243).  feats: (T, A, B) with
 * B = 2*width_over2; cells outside [0,size0)x[0,size1) get +inf (:262-263).
 */
SVO_API void svo_make_sparse_costs(const float *vecs0, const float *vecs1,
                                   const float *norms0, const float *norms1,
                                   int size0, int size1, int dim,
                                   const int32_t *path, int a_len,
                                   const int32_t *xoffs, const int32_t *yoffs, int ntypes,
                                   int width_over2,
                                   float *feats, int32_t *b_offset)
{
    const int b_len = 2 * width_over2;
    for (int ii = 0; ii < a_len; ++ii) {
        const int x = path[2 * ii], y = path[2 * ii + 1];
        const int aa = x + y;
        b_offset[aa] = y - width_over2;
        for (int b = 0; b < b_len; ++b) {
            const int yy = y - width_over2 + b;
            const int xx = aa - yy;
            const int inside = (0 <= xx && xx < size0 && 0 <= yy && yy < size1);
            for (int t = 0; t < ntypes; ++t) {
                float feat;
                if (inside) {
                    const int xo = xoffs[t], yo = yoffs[t];
                    const float *a = vecs0 + ((size_t)(xo - 1) * size0 + xx) * dim;
                    const float *c = vecs1 + ((size_t)(yo - 1) * size1 + yy) * dim;
                    float sumx = seq_dot(a, c, dim);
                    double num = ((2.0 * (double)xo) * (double)yo) * (1.0 - (double)sumx);
                    double den = (1e-6 + (double)norms0[(size_t)(xo - 1) * size0 + xx])
                                 + (double)norms1[(size_t)(yo - 1) * size1 + yy];
                    feat = (float)(num / den);
                } else {
                    feat = INFINITY;
                }
                feats[((size_t)t * a_len + aa) * b_len + b] = feat;
            }
        }
    }
}

/*
 * dp_core.pyx:269-404 sparse_dp.  Nodes (A+2, B); b_offset_out = [b_in[0], b_in[0]] ++ (b_in+1)
 * (:327-328).  Types are the T cost types followed by (0,1) then (1,0) (:306-307), tried in
 * that order with strict '<' (:400).  A deletion is only allowed where the diagonal cost cell
 * (x-1,y-1) is inside the document and inside the cost band (:382,390) — reference quirk kept.
 */
SVO_API void svo_sparse_dp(const float *costs, const int32_t *b_offset_in,
                           const int32_t *xoffs_in, const int32_t *yoffs_in, int ntypes,
                           int a_in, int b_in, double pen, int x_in_size, int y_in_size,
                           double *csum, int32_t *xp, int32_t *yp, int32_t *b_offset_out)
{
    const int a_out = a_in + 2, b_out = b_in;
    const int x_out_size = x_in_size + 1, y_out_size = y_in_size + 1;
    const int nall = ntypes + 2;
    if (a_in > 0) {
        b_offset_out[0] = b_offset_in[0];
        b_offset_out[1] = b_offset_in[0];
    }
    for (int a = 0; a < a_in; ++a) b_offset_out[a + 2] = b_offset_in[a] + 1;

    for (int aa = 0; aa < a_out; ++aa) {
        for (int bb = 0; bb < b_out; ++bb) {
            const size_t o = (size_t)aa * b_out + bb;
            const int yy = bb + b_offset_out[aa];
            const int xx = aa - yy;
            if (xx == 0 && 0 <= yy && yy < y_out_size) {
                csum[o] = pen * (double)yy; xp[o] = 0; yp[o] = 1;
            } else if (yy == 0 && 0 <= xx && xx < x_out_size) {
                csum[o] = pen * (double)xx; xp[o] = 1; yp[o] = 0;
            } else {
                double best = INFINITY; int32_t bx = -42, by = -42;
                const int xc = xx - 1, yc = yy - 1;
                for (int t = 0; t < nall; ++t) {
                    const int xo = t < ntypes ? xoffs_in[t] : (t == ntypes ? 0 : 1);
                    const int yo = t < ntypes ? yoffs_in[t] : (t == ntypes ? 1 : 0);
                    const int xq = xx - xo, yq = yy - yo;
                    if (!(0 <= xc && xc < x_in_size && 0 <= yc && yc < y_in_size &&
                          0 <= xq && xq < x_out_size && 0 <= yq && yq < y_out_size))
                        continue;
                    const int ac = xc + yc;
                    const int bc = yc - b_offset_in[ac];
                    const int aq = xq + yq;
                    const int bq = yq - b_offset_out[aq];
                    if (!(0 <= ac && ac < a_in && 0 <= bc && bc < b_in &&
                          0 <= aq && aq < a_out && 0 <= bq && bq < b_out))
                        continue;
                    double step = (xo == 0 || yo == 0)
                                      ? pen
                                      : (double)costs[((size_t)t * a_in + ac) * b_in + bc];
                    double total = csum[(size_t)aq * b_out + bq] + step;
                    if (total < best) { best = total; bx = xo; by = yo; }
                }
                csum[o] = best; xp[o] = bx; yp[o] = by;
            }
        }
    }
}
