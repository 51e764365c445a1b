// svx_math.h — arithmetic shared by the sm_100a kernels and their host twins.
//
// Everything here is order- and rounding-exact with respect to the reference
// (svecalign/vecalign/dp_core.pyx and the numpy calls in dp_utils.py); the host build uses plain
// C operators (compiled with -ffp-contract=off), the device build uses the _rn intrinsics so that
// nvcc can never contract a multiply and an add into an FMA.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define SVX_HD __host__ __device__ __forceinline__
#define SVX_FMUL(a, b) __fmul_rn((a), (b))
#define SVX_FADD(a, b) __fadd_rn((a), (b))
#define SVX_FSUB(a, b) __fsub_rn((a), (b))
#define SVX_FDIV(a, b) __fdiv_rn((a), (b))
#define SVX_DMUL(a, b) __dmul_rn((a), (b))
#define SVX_DADD(a, b) __dadd_rn((a), (b))
#define SVX_DSUB(a, b) __dsub_rn((a), (b))
#define SVX_DDIV(a, b) __ddiv_rn((a), (b))
#else
#if defined(__CUDACC__)
#define SVX_HD __host__ __device__ inline
#else
#define SVX_HD inline
#endif
#define SVX_FMUL(a, b) ((float)(a) * (float)(b))
#define SVX_FADD(a, b) ((float)(a) + (float)(b))
#define SVX_FSUB(a, b) ((float)(a) - (float)(b))
#define SVX_FDIV(a, b) ((float)(a) / (float)(b))
#define SVX_DMUL(a, b) ((double)(a) * (double)(b))
#define SVX_DADD(a, b) ((double)(a) + (double)(b))
#define SVX_DSUB(a, b) ((double)(a) - (double)(b))
#define SVX_DDIV(a, b) ((double)(a) / (double)(b))
#endif

// ---------------------------------------------------------------------------------------------
// Cost formulas.  The numerator/denominator are evaluated in double and the quotient is narrowed
// to fp32, exactly as the C generated from dp_core.pyx does.
// ---------------------------------------------------------------------------------------------

// dp_core.pyx:259-260 (make_sparse_costs): ((2.0*xo)*yo)*(1.0-dot) / ((1e-6+n0)+n1)
SVX_HD float svx_band_cost(float dot, int xo, int yo, float n0, float n1)
{
    double num = SVX_DMUL(SVX_DMUL(SVX_DMUL(2.0, (double)xo), (double)yo), SVX_DSUB(1.0, (double)dot));
    double den = SVX_DADD(SVX_DADD(1e-6, (double)n0), (double)n1);
    return (float)SVX_DDIV(num, den);
}

// dp_core.pyx:73-75 (make_dense_costs, offsets 0): 2.0*(1.0-dot) / ((1e-6+n0)+n1); the trailing
// fp32 multiplications by (offset+1) == 1 are identities.
SVX_HD float svx_dense_cost(float dot, float n0, float n1)
{
    double num = SVX_DMUL(2.0, SVX_DSUB(1.0, (double)dot));
    double den = SVX_DADD(SVX_DADD(1e-6, (double)n0), (double)n1);
    return (float)SVX_DDIV(num, den);
}

// dp_core.pyx:161 (score_path): 2.0*(1.0-dot) / (float)(n1+n2)  — fp32 denominator, no 1e-6.
SVX_HD float svx_pair_score(float dot, float n1, float n2)
{
    double num = SVX_DMUL(2.0, SVX_DSUB(1.0, (double)dot));
    float den = SVX_FADD(n1, n2);
    return (float)SVX_DDIV(num, (double)den);
}

// ---------------------------------------------------------------------------------------------
// Search-path slant (dp_utils.py:177-196 append_slant).  Point i (1-based) of a segment that
// starts at (xs,ys) and spans (xw,yw): x = xs + round(xw*i/NN), y = ys + round(yw*i/NN) with
// Python's round (half to even), then forced to advance x+y by exactly one relative to the
// previous point — whose coordinate sum is xs+ys+i-1 by construction, so every point can be
// computed independently.  Integer arithmetic reproduces round(int/int) exactly.
// ---------------------------------------------------------------------------------------------
SVX_HD long long svx_round_half_even_div(long long p, long long q)  // p >= 0, q > 0
{
    long long quo = p / q, rem = p % q;
    if (2 * rem > q) return quo + 1;
    if (2 * rem < q) return quo;
    return quo + (quo & 1);
}

SVX_HD void svx_slant_point(long long xs, long long ys, long long xw, long long yw, long long i,
                            int *x_out, int *y_out)
{
    const long long nn = xw + yw;
    long long rx = svx_round_half_even_div(xw * i, nn);
    long long ry = svx_round_half_even_div(yw * i, nn);
    long long jump = rx + ry - (i - 1);
    if (jump == 2) rx -= 1;
    else if (jump == 0) rx += 1;
    *x_out = (int)(xs + rx);
    *y_out = (int)(ys + ry);
}

// ---------------------------------------------------------------------------------------------
// Deletion knob (dp_utils.py:43-79 DeletionKnob fed by make_del_knob :278-323), numpy >= 2.0
// arithmetic: np.histogram(samp fp32, bins=1000, range=[0, max fp32], density=True) -> fp32 bin
// edges from an fp32 linspace, fp32 index estimate with edge correction, fp64 density;
// cdf = cumsum(hist) * fp32 width; 27 interior knob points at k*(1/28) via searchsorted;
// np.interp at `frac`.
// ---------------------------------------------------------------------------------------------
#define SVX_KNOB_BINS 1000

SVX_HD float svx_knob_edge(int i, float step, float maxv)
{
    return i == SVX_KNOB_BINS ? maxv : SVX_FMUL((float)i, step);
}

// bin index of sample a, or -1 if numpy's `keep` mask drops it.
SVX_HD int svx_knob_bin(float a, float maxv, float step)
{
    if (!(a >= 0.0f && a <= maxv)) return -1;
    float f = SVX_FMUL(SVX_FDIV(a, maxv), (float)SVX_KNOB_BINS);
    int idx = (int)f;
    if (idx == SVX_KNOB_BINS) idx -= 1;
    if (a < svx_knob_edge(idx, step, maxv)) idx -= 1;
    if (a >= svx_knob_edge(idx + 1, step, maxv) && idx != SVX_KNOB_BINS - 1) idx += 1;
    return idx;
}

// density of bin i: hist[i] = counts[i] / width_i / total  (np.histogram(density=True), fp64)
SVX_HD double svx_knob_density(unsigned int count, int i, float step, float maxv, long long total)
{
    const double db = (double)SVX_FSUB(svx_knob_edge(i + 1, step, maxv), svx_knob_edge(i, step, maxv));
    return SVX_DDIV(SVX_DDIV((double)count, db), (double)total);
}

// density[1000] (svx_knob_density of every bin; may be computed in parallel) -> del_penalty.
// maxv is max(samples) (fp32).  Degenerate maxv <= 0 follows the reference's
// "res_max = res_min + 1e-4" branch, whose knob is 0 up to 27/28 and ends at 1e-4.
SVX_HD double svx_knob_finish(const double *density, float maxv, double frac)
{
    double ys[29], xs[29];
    const double qstep = SVX_DDIV(1.0, 28.0);
    xs[0] = 0.0; ys[0] = 0.0; xs[28] = 1.0;
    for (int k = 1; k < 28; ++k) xs[k] = SVX_DMUL((double)k, qstep);
    if (!(maxv > 0.0f)) {
        for (int k = 1; k < 28; ++k) ys[k] = 0.0;
        ys[28] = 1e-4;
    } else {
        const float step = SVX_FDIV(maxv, (float)SVX_KNOB_BINS);
        const double dx = (double)SVX_FSUB(svx_knob_edge(1, step, maxv), svx_knob_edge(0, step, maxv));
        double cum = 0.0;
        int k = 1;
        for (int i = 0; i < SVX_KNOB_BINS && k < 28; ++i) {
            cum = (i == 0) ? density[i] : SVX_DADD(cum, density[i]);      // np.cumsum: sequential fp64
            double cdf = SVX_DMUL(cum, dx);
            while (k < 28 && !(cdf < xs[k])) {   // searchsorted(left); NaN sorts last
                ys[k] = SVX_DMUL(SVX_DDIV((double)i, 1000.0), (double)maxv);
                ++k;
            }
        }
        for (; k < 28; ++k) ys[k] = SVX_DMUL(SVX_DDIV(1000.0, 1000.0), (double)maxv);
        ys[28] = (double)maxv;
    }
    // np.interp (numpy/_core/src/multiarray/compiled_base.c arr_interp), single query point
    if (frac != frac) return frac;
    if (frac < xs[0]) return ys[0];
    if (frac > xs[28]) return ys[28];
    int j = 0;
    while (j < 28 && xs[j + 1] <= frac) ++j;       // xs[j] <= frac < xs[j+1]
    if (j == 28) return ys[28];
    if (xs[j] == frac) return ys[j];
    const double slope = SVX_DDIV(SVX_DSUB(ys[j + 1], ys[j]), SVX_DSUB(xs[j + 1], xs[j]));
    double r = SVX_DADD(SVX_DMUL(slope, SVX_DSUB(frac, xs[j])), ys[j]);
    if (r != r) {
        r = SVX_DADD(SVX_DMUL(slope, SVX_DSUB(frac, xs[j + 1])), ys[j + 1]);
        if (r != r && ys[j] == ys[j + 1]) r = ys[j];
    }
    return r;
}
