// prologue.cu — HBM-streaming kernels in front of the DP (sm_100a):
//   svx_normalize_rows  (dp_utils.py:32-40   make_norm1)
//   svx_downsample      (dp_utils.py:362-378 downsample_vectors)
//   svx_sample_norms    (dp_utils.py:326-359 compute_norms, GEMV form)
//   svx_score_pairs     (dp_core.pyx:143-161 score_path)
//   svx_del_knob        (dp_utils.py:43-79   DeletionKnob, numpy >= 2 arithmetic)
//
// Row kernels use one warp per 4 KB embedding row: eight coalesced 16-byte loads per lane
// (one 128-float numpy "pairwise block" per load step), the numpy pairwise-sum tree is then
// replayed exactly through a padded, bank-conflict-free shared-memory transpose.
#include <cuda_fp16.h>
#include "svx_common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kRowPad = 136;  // 128-float block + 8 pad floats: (lane + 8 i) % 32 banks

// ---------------------------------------------------------------------------------------------
// numpy pairwise sum of one row's squares, bit-exact (numpy/_core/src/umath/loops_utils.h.src
// pairwise_sum: 8 sequential accumulators per <=128-element block, balanced tree above).
// sq: this warp's padded scratch, already holding the DIM squares.  Returns the sum in all lanes.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ float np_pairwise_from_smem(const float *sq, int lane)
{
    constexpr int NB = DIM / 128;            // pairwise leaf blocks
    constexpr int NACC = NB * 8;             // leaf accumulators
    constexpr int NPASS = (NACC + 31) / 32;
    float pass_sum[NPASS];
#pragma unroll
    for (int m = 0; m < NPASS; ++m) {
        const int q = lane + 32 * m;
        float acc = 0.0f;
        if (q < NACC) {
            const float *p = sq + (q >> 3) * kRowPad + (q & 7);
            acc = p[0];
#pragma unroll
            for (int i = 1; i < 16; ++i) acc = __fadd_rn(acc, p[8 * i]);
        }
        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) inside a block, then the block tree
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
        if (NB >= 2) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));
        if (NB >= 4) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
        pass_sum[m] = acc;
    }
    float total;
    if constexpr (NPASS == 1) total = pass_sum[0];
    else if constexpr (NPASS == 2) total = __fadd_rn(pass_sum[0], pass_sum[1]);
    else total = __fadd_rn(__fadd_rn(pass_sum[0], pass_sum[1]), __fadd_rn(pass_sum[2], pass_sum[3]));
    return __shfl_sync(0xffffffffu, total, 0);
}

// v[s] holds elements s*128 + 4*lane .. +3 of the row.  Normalises in registers.
template <int DIM>
__device__ __forceinline__ void warp_unit_row(float4 (&v)[DIM / 128], float *sq, int lane)
{
    constexpr int NB = DIM / 128;
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        float4 q;
        q.x = __fmul_rn(v[s].x, v[s].x); q.y = __fmul_rn(v[s].y, v[s].y);
        q.z = __fmul_rn(v[s].z, v[s].z); q.w = __fmul_rn(v[s].w, v[s].w);
        *reinterpret_cast<float4 *>(sq + s * kRowPad + 4 * lane) = q;
    }
    __syncwarp();
    const float total = np_pairwise_from_smem<DIM>(sq, lane);
    __syncwarp();
    const float den = __fadd_rn(__fsqrt_rn(total), 1e-5f);
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        v[s].x = __fdiv_rn(v[s].x, den); v[s].y = __fdiv_rn(v[s].y, den);
        v[s].z = __fdiv_rn(v[s].z, den); v[s].w = __fdiv_rn(v[s].w, den);
    }
}

// Row r = o * n + i of a level whose raw rows come from a source (SvxRowSource): the lane's elements
// b * 128 + 4 * lane .. + 3 of every 128-float block, exactly what svx_gather_doc_embedding would have written
// (widened fp16 / fp32, zeros without a source row, zeros when the source row holds a NaN).  Warp-uniform return
// value: the row was zeroed because of a NaN.
template <int DIM>
__device__ __forceinline__ bool warp_source_row(const SvxRowSource &src, int64_t r, int lane, float4 (&v)[DIM / 128])
{
    constexpr int NB = DIM / 128;
    const int s = src.table ? __ldg(src.table + r) : (int)r;
    const bool have = s >= 0 && s < src.nrows;
    bool bad = false;
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) {
        if (src.is_fp16) {
            const __half *p = reinterpret_cast<const __half *>(src.rows) + (size_t)s * DIM + 4 * lane;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p + b * 128));
                const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
                v[b] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
        } else {
            const float *p = reinterpret_cast<const float *>(src.rows) + (size_t)s * DIM + 4 * lane;
#pragma unroll
            for (int b = 0; b < NB; ++b) v[b] = ldg_f4(p + b * 128);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) bad |= (v[b].x != v[b].x) | (v[b].y != v[b].y) | (v[b].z != v[b].z) | (v[b].w != v[b].w);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (bad) {
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return bad;
}

template <int DIM>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_normalize(const SvxRows *jobs)
{
    constexpr int NB = DIM / 128;
    __shared__ __align__(16) float scratch[kWarpsPerCta][NB * kRowPad];
    const SvxRows job = jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + warp; row < job.nrows;
         row += (int64_t)gridDim.x * kWarpsPerCta) {
        float *p = job.ptr + row * DIM;
        float4 v[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) v[s] = *reinterpret_cast<const float4 *>(p + s * 128 + 4 * lane);
        warp_unit_row<DIM>(v, scratch[warp], lane);
#pragma unroll
        for (int s = 0; s < NB; ++s) *reinterpret_cast<float4 *>(p + s * 128 + 4 * lane) = v[s];
    }
}

// ---------------------------------------------------------------------------------------------
// downsample, pass A: out[o,j,:] = in[o,2j,:] + in[o,2j+1,:] and the per-overlap mean row.
// np.mean(axis=0) accumulates the rows sequentially in fp32, so one thread owns one column.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pairsum_colmean(const SvxDownJob *jobs, int dim)
{
    const SvxDownJob job = jobs[blockIdx.y];
    const int cblocks = dim >> 7;
    const int o = blockIdx.x / cblocks;
    if (o >= job.k) return;
    const int c = (blockIdx.x % cblocks) * 128 + threadIdx.x;
    const int m = job.n >> 1;
    if (m == 0) return;
    const float *src = job.in + (size_t)o * job.n * dim + c;
    float *dst = job.out + (size_t)o * m * dim + c;
    float acc = 0.0f;
    int j = 0;
    for (; j + 4 <= m; j += 4) {
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = __ldg(src + (size_t)(2 * j + u) * dim);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float h = __fadd_rn(a[2 * u], a[2 * u + 1]);
            dst[(size_t)(j + u) * dim] = h;
            acc = (j + u == 0) ? h : __fadd_rn(acc, h);
        }
    }
    for (; j < m; ++j) {
        const float h = __fadd_rn(__ldg(src + (size_t)(2 * j) * dim), __ldg(src + (size_t)(2 * j + 1) * dim));
        dst[(size_t)j * dim] = h;
        acc = (j == 0) ? h : __fadd_rn(acc, h);
    }
    job.mean[(size_t)o * dim + c] = __fdiv_rn(acc, (float)m);
}

// downsample, pass B: row -= mean row; unit-normalise.  In place on `out`.
template <int DIM>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_center_normalize(const SvxDownJob *jobs)
{
    constexpr int NB = DIM / 128;
    __shared__ __align__(16) float scratch[kWarpsPerCta][NB * kRowPad];
    const SvxDownJob job = jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = job.n >> 1;
    const int64_t nrows = (int64_t)job.k * m;
    for (int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + warp; row < nrows;
         row += (int64_t)gridDim.x * kWarpsPerCta) {
        const int o = (int)(row / m);
        float *p = job.out + row * DIM;
        const float *mu = job.mean + (size_t)o * DIM;
        float4 v[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            const float4 h = *reinterpret_cast<const float4 *>(p + s * 128 + 4 * lane);
            const float4 g = *reinterpret_cast<const float4 *>(mu + s * 128 + 4 * lane);
            v[s].x = __fsub_rn(h.x, g.x); v[s].y = __fsub_rn(h.y, g.y);
            v[s].z = __fsub_rn(h.z, g.z); v[s].w = __fsub_rn(h.w, g.w);
        }
        warp_unit_row<DIM>(v, scratch[warp], lane);
#pragma unroll
        for (int s = 0; s < NB; ++s) *reinterpret_cast<float4 *>(p + s * 128 + 4 * lane) = v[s];
    }
}

// ---------------------------------------------------------------------------------------------
// sample norms.  S1: mean sample vector in fp64; S2: norms = 1 - row . mbar (fp64 accumulate).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sample_mean(const SvxNormJob *jobs, int dim)
{
    const SvxNormJob job = jobs[blockIdx.y];
    const int c = blockIdx.x * 128 + threadIdx.x;
    const int nsamp = job.ko * job.per;
    if (nsamp <= 0 || job.no <= 0) return;
    double acc = 0.0;
    for (int o = 0; o < job.ko; ++o) {
        const float *base = job.other + (size_t)o * job.no * dim + c;
        const int32_t *ix = job.idx + (size_t)o * job.per;
        for (int s = 0; s < job.per; ++s) acc += (double)__ldg(base + (size_t)__ldg(ix + s) * dim);
    }
    job.mbar[c] = acc / (double)nsamp;
}

template <int DIM>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_norms_gemv(const SvxNormJob *jobs)
{
    constexpr int NB = DIM / 128;
    __shared__ double mb[DIM];
    const SvxNormJob job = jobs[blockIdx.y];
    const int64_t nrows = (int64_t)job.k * job.n;
    if ((int64_t)blockIdx.x * kWarpsPerCta >= nrows) return;
    for (int i = threadIdx.x; i < DIM; i += blockDim.x) mb[i] = job.mbar[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + warp; row < nrows;
         row += (int64_t)gridDim.x * kWarpsPerCta) {
        const float *p = job.vecs + row * DIM;
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            const float4 v = ldg_f4(p + s * 128 + 4 * lane);
            const double *q = mb + s * 128 + 4 * lane;
            acc = fma((double)v.x, q[0], acc); acc = fma((double)v.y, q[1], acc);
            acc = fma((double)v.z, q[2], acc); acc = fma((double)v.w, q[3], acc);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) job.norms[row] = __fsub_rn(1.0f, (float)acc);
    }
}

// ---------------------------------------------------------------------------------------------
// sampled pair scores: one thread per sample, the dot product in the reference's order.
//
// The samples are random (x, y) row pairs (np.random.choice draws), so a warp would touch 64
// different 4 KB rows per step.  k_sort_samples orders each job's samples by x with a counting
// sort in shared memory (one CTA per job; results are written back by original sample index, so
// the order inside a bucket is irrelevant): consecutive threads then share their x row
// (broadcast / L1 hits) and only the y row is a per-thread stream, which halves the L2->SM
// traffic that bounds this kernel.  When the job carries the dense dot matrix of its level
// (`dots`, coarsest level) the score is a gather.
// ---------------------------------------------------------------------------------------------
constexpr int kSortMaxRows = 16384;      // counters of one job in shared memory (64 KB)

__global__ void __launch_bounds__(512) k_sort_samples(const SvxScoreJob *jobs)
{
    extern __shared__ int cnt[];         // ne counters, then the running offsets
    __shared__ int wsum[16];
    const SvxScoreJob job = jobs[blockIdx.x];
    if (!job.perm || !job.xi || job.dots || job.ne > kSortMaxRows || job.nsamp <= 0) return;
    const int ne = job.ne, n = job.nsamp, tid = threadIdx.x;
    for (int i = tid; i < ne; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) atomicAdd(&cnt[job.xi[i]], 1);
    __syncthreads();
    // exclusive scan of cnt[0..ne): each thread owns a contiguous slice
    const int per = (ne + blockDim.x - 1) / blockDim.x;
    const int lo = min(tid * per, ne), hi = min(lo + per, ne);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += cnt[i];
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31) >= off) incl += v;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int wq = 0; wq < (tid >> 5); ++wq) base += wsum[wq];
    for (int i = lo; i < hi; ++i) { const int c = cnt[i]; cnt[i] = base; base += c; }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) job.perm[atomicAdd(&cnt[job.xi[i]], 1)] = i;
}

// A warp scores 32 samples.  Each lane must consume its own two rows in increasing d (the
// reference's summation order), but 32 lanes streaming 32 different rows would cost one L1 tag
// look-up per lane and 16 bytes: instead the warp fetches every 128-byte row piece with 8 lanes
// (coalesced 16-byte cp.async, 4 rows per instruction) into a 32 x 32-float block of shared memory
// (row stride 36 floats: conflict-free for the copy and for the per-lane LDS.128 read-back) and each
// lane then reads its own row back.
constexpr int kScWarps = 4;
constexpr int kScStride = 36;

template <bool EXACT>
__global__ void __launch_bounds__(kScWarps * 32) k_score_pairs(const SvxScoreJob *jobs, int dim)
{
    __shared__ __align__(16) float sy[kScWarps][2][32 * kScStride];
    const SvxScoreJob job = jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base = (blockIdx.x * kScWarps + warp) * 32;
    if (base >= job.nsamp) return;                       // warp-uniform
    int i = base + lane;
    const bool valid = i < job.nsamp;
    int xi = 0, yi = 0;
    if (valid) {
        if (job.xi) {
            if (job.perm && !job.dots && job.ne <= kSortMaxRows) i = job.perm[i];
            xi = job.xi[i]; yi = job.yi[i];
        } else { xi = i / job.nf; yi = i % job.nf; }
    }
    float dot = 0.0f;
    if (job.dots) {
        if (valid) dot = job.dots[(size_t)xi * job.nf + yi];
    } else {
        // y rows: moved global -> shared by cp.async (no register staging, no STS wavefronts - the LSU data
        // pipe is what bounds this kernel), this lane copying in step g the 16-byte piece lane%8 of sample
        // 4g + lane/8; double-buffered, so block d0 + 32 is in flight while block d0 is consumed.
        // x rows: the samples are sorted by x, so the warp's 32 samples share a handful of x rows; each
        // lane reads its own x row directly (lanes with the same row hit the same 16 bytes: one L1
        // tag per distinct row), no staging.
        const int sub = lane >> 3, piece = 4 * (lane & 7);
        const float *ysrc[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) ysrc[g] = job.f + (size_t)__shfl_sync(0xffffffffu, yi, 4 * g + sub) * dim + piece;
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(&sy[warp][0][0]) + (unsigned)((sub * kScStride + piece) * sizeof(float));
        auto issue = [&](int d0, int buf) {
            const unsigned dst = sbase + (unsigned)(buf * 32 * kScStride * sizeof(float));
#pragma unroll
            for (int g = 0; g < 8; ++g)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + (unsigned)(4 * g * kScStride * sizeof(float))),
                             "l"(ysrc[g] + d0));
            asm volatile("cp.async.commit_group;\n" ::);
        };
        const float *xrow = job.e + (size_t)xi * dim;
        issue(0, 0);
        for (int d0 = 0, buf = 0; d0 < dim; d0 += 32, buf ^= 1) {
            float4 vx[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) vx[q] = ldg_f4(xrow + d0 + 4 * q);
            if (d0 + 32 < dim) {
                issue(d0 + 32, buf ^ 1);
                asm volatile("cp.async.wait_group 1;\n" ::);
            } else {
                asm volatile("cp.async.wait_group 0;\n" ::);
            }
            __syncwarp();
            const float *my = &sy[warp][buf][lane * kScStride];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a = vx[q];
                const float4 b = *reinterpret_cast<const float4 *>(my + 4 * q);
                if (EXACT) {
                    dot = __fadd_rn(dot, __fmul_rn(a.x, b.x)); dot = __fadd_rn(dot, __fmul_rn(a.y, b.y));
                    dot = __fadd_rn(dot, __fmul_rn(a.z, b.z)); dot = __fadd_rn(dot, __fmul_rn(a.w, b.w));
                } else {
                    dot = fmaf(a.x, b.x, dot); dot = fmaf(a.y, b.y, dot);
                    dot = fmaf(a.z, b.z, dot); dot = fmaf(a.w, b.w, dot);
                }
            }
            __syncwarp();           // every lane is done with `buf` before the next iteration refills it
        }
    }
    if (valid) job.scores[i] = svx_pair_score(dot, job.norm_e[xi], job.norm_f[yi]);
}

// ---------------------------------------------------------------------------------------------
// deletion knob: one CTA per job.  max -> shared-memory histogram (integer counts are order
// independent) -> thread 0 replays numpy's fp64 cdf / searchsorted / interp.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_del_knob(const SvxScoreJob *jobs, double frac)
{
    __shared__ unsigned int hist[SVX_KNOB_BINS];
    __shared__ float wmax[8];
    const SvxScoreJob job = jobs[blockIdx.x];
    const int n = job.nsamp;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, job.scores[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mx;
    for (int i = threadIdx.x; i < SVX_KNOB_BINS; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    mx = wmax[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, wmax[w]);
    if (mx > 0.0f) {
        const float step = __fdiv_rn(mx, (float)SVX_KNOB_BINS);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int b = svx_knob_bin(job.scores[i], mx, step);
            if (b >= 0) atomicAdd(&hist[b], 1u);
        }
    }
    __syncthreads();
    // per-bin densities in parallel (exact per bin); only the fp64 running sum is sequential
    __shared__ double dens[SVX_KNOB_BINS];
    __shared__ long long tot_s;
    if (threadIdx.x < 32) {
        long long t = 0;
        for (int i = threadIdx.x; i < SVX_KNOB_BINS; i += 32) t += hist[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (threadIdx.x == 0) tot_s = t;
    }
    __syncthreads();
    if (mx > 0.0f) {
        const float step = __fdiv_rn(mx, (float)SVX_KNOB_BINS);
        for (int i = threadIdx.x; i < SVX_KNOB_BINS; i += blockDim.x) dens[i] = svx_knob_density(hist[i], i, step, mx, tot_s);
    }
    __syncthreads();
    if (threadIdx.x == 0) *job.del_penalty = svx_knob_finish(dens, mx, frac);
}

// ---------------------------------------------------------------------------------------------
// Fused level prologue (svx_level_prologue): the same per-row arithmetic as k_normalize /
// k_pairsum_colmean / k_center_normalize / k_sample_mean / k_norms_gemv, in fewer passes over HBM.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxLevelSamples = 2048;     // sampled rows per job whose denominators fit in smem

// step 1: mean row per overlap of the un-centred pair sums (sequential fp32 accumulation over rows
// = np.mean(axis=0)), one thread per column, 8 loads in flight.
__global__ void __launch_bounds__(128) k_level_colmean(const SvxLevelJob *jobs, int dim)
{
    const SvxLevelJob job = jobs[blockIdx.y];
    if (!job.mean || job.n <= 0) return;
    const int cblocks = dim >> 7;
    const int o = blockIdx.x / cblocks;
    if (o >= job.k) return;
    const int c = (blockIdx.x % cblocks) * 128 + threadIdx.x;
    const float *src = job.vecs + (size_t)o * job.n * dim + c;
    float acc = 0.0f;
    int j = 0;
    for (; j + 8 <= job.n; j += 8) {
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = __ldg(src + (size_t)(j + u) * dim);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = (j + u == 0) ? a[u] : __fadd_rn(acc, a[u]);
    }
    for (; j < job.n; ++j) {
        const float a = __ldg(src + (size_t)j * dim);
        acc = (j == 0) ? a : __fadd_rn(acc, a);
    }
    job.mean[(size_t)o * dim + c] = __fdiv_rn(acc, (float)job.n);
}

// step 2: mbar = mean over the sampled rows of the other side, each centred and unit-normalised on
// the fly exactly as step 3 will finish it.  (a) a warp per sampled row computes its denominator
// into the scratch behind mbar; (b) a thread per column accumulates in fp64 in sample order.
template <int DIM>
__global__ void __launch_bounds__(256) k_level_sample_den(const SvxLevelJob *jobs)
{
    constexpr int NB = DIM / 128;
    __shared__ __align__(16) float scratch[8][NB * kRowPad];
    const SvxLevelJob job = jobs[blockIdx.y];
    const int nsamp = job.ko * job.per;
    if (!job.idx || !job.norms || nsamp <= 0 || job.no <= 0 || job.n <= 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * 8 + warp;
    if (s >= nsamp) return;
    float *den = reinterpret_cast<float *>(job.mbar + DIM);
    const int o = s / job.per;
    const float *p = job.other + ((size_t)o * job.no + job.idx[s]) * DIM;
    float4 v[NB];
    // a sampled row that the source zeroes because of a NaN is flagged to step (b) by a negative denominator
    const bool zeroed = job.osrc.rows ? warp_source_row<DIM>(job.osrc, (int64_t)o * job.no + job.idx[s], lane, v) : false;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (!job.osrc.rows) v[b] = ldg_f4(p + b * 128 + 4 * lane);
        if (job.other_mean) {
            const float4 g = ldg_f4(job.other_mean + (size_t)o * DIM + b * 128 + 4 * lane);
            v[b].x = __fsub_rn(v[b].x, g.x); v[b].y = __fsub_rn(v[b].y, g.y);
            v[b].z = __fsub_rn(v[b].z, g.z); v[b].w = __fsub_rn(v[b].w, g.w);
        }
        float4 q;
        q.x = __fmul_rn(v[b].x, v[b].x); q.y = __fmul_rn(v[b].y, v[b].y);
        q.z = __fmul_rn(v[b].z, v[b].z); q.w = __fmul_rn(v[b].w, v[b].w);
        *reinterpret_cast<float4 *>(scratch[warp] + b * kRowPad + 4 * lane) = q;
    }
    __syncwarp();
    const float total = np_pairwise_from_smem<DIM>(scratch[warp], lane);
    if (lane == 0) {
        const float d = __fadd_rn(__fsqrt_rn(total), 1e-5f);
        den[s] = zeroed ? -d : d;
    }
}

__global__ void __launch_bounds__(128) k_level_sample_acc(const SvxLevelJob *jobs, int dim)
{
    const SvxLevelJob job = jobs[blockIdx.y];
    const int nsamp = job.ko * job.per;
    if (!job.idx || !job.norms || nsamp <= 0 || job.no <= 0 || job.n <= 0) return;
    const int c = blockIdx.x * 128 + threadIdx.x;
    const float *den = reinterpret_cast<const float *>(job.mbar + dim);
    double acc = 0.0;
    const SvxRowSource &os = job.osrc;
    for (int o = 0; o < job.ko; ++o) {
        const float *base = job.other + (size_t)o * job.no * dim + c;
        const float mu = job.other_mean ? __ldg(job.other_mean + (size_t)o * dim + c) : 0.0f;
        const int32_t *ix = job.idx + (size_t)o * job.per;
        const float *dn = den + (size_t)o * job.per;
        // element c of sampled row `row` of overlap o, raw: from `other`, or through the source (zeros without a
        // source row and - flagged by step (a) with a negative denominator - for a source row holding a NaN)
        auto element = [&](int row, float d) -> float {
            if (!os.rows) return __ldg(base + (size_t)row * dim);
            const int64_t r = (int64_t)o * job.no + row;
            const int sr = os.table ? __ldg(os.table + r) : (int)r;
            if (d < 0.0f || sr < 0 || sr >= os.nrows) return 0.0f;
            return os.is_fp16 ? __half2float(__ldg(reinterpret_cast<const __half *>(os.rows) + (size_t)sr * dim + c))
                              : __ldg(reinterpret_cast<const float *>(os.rows) + (size_t)sr * dim + c);
        };
        int s = 0;
        for (; s + 4 <= job.per; s += 4) {
            float x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = element(__ldg(ix + s + u), dn[s + u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float xc = job.other_mean ? __fsub_rn(x[u], mu) : x[u];
                acc += (double)__fdiv_rn(xc, fabsf(dn[s + u]));
            }
        }
        for (; s < job.per; ++s) {
            float x = element(__ldg(ix + s), dn[s]);
            if (job.other_mean) x = __fsub_rn(x, mu);
            acc += (double)__fdiv_rn(x, fabsf(dn[s]));
        }
    }
    job.mbar[c] = acc / (double)nsamp;
}

// step 3: a warp per row pair (2j, 2j+1) of one overlap.
#ifndef SVX_FIN_WARPS
#define SVX_FIN_WARPS 8
#define SVX_FIN_CTAS 2
#endif
constexpr int kFinWarps = SVX_FIN_WARPS, kFinCtas = SVX_FIN_CTAS;
constexpr int kFinPairs = 8;        // row pairs per warp when the launch is large enough (amortises the CTA's mbar load)
template <int DIM>
__global__ void __launch_bounds__(kFinWarps * 32, kFinCtas) k_level_finish(const SvxLevelJob *jobs)
{
    constexpr int NB = DIM / 128;
    __shared__ __align__(16) float scratch[kFinWarps][NB * kRowPad];
    __shared__ double mb[DIM];
    const SvxLevelJob job = jobs[blockIdx.y];
    const int npair = (job.n + 1) >> 1;
    const int64_t total = (int64_t)job.k * npair;
    if ((int64_t)blockIdx.x * kFinWarps >= total) return;
    const bool want_norms = job.norms && job.idx && job.ko * job.per > 0 && job.no > 0;
    if (want_norms) {
        for (int i = threadIdx.x; i < DIM; i += blockDim.x) mb[i] = job.mbar[i];
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = job.n >> 1;
    for (int64_t p = (int64_t)blockIdx.x * kFinWarps + warp; p < total; p += (int64_t)gridDim.x * kFinWarps) {
        const int o = (int)(p / npair), j = (int)(p % npair);
        float4 v[2][NB];
        const int nrow = (2 * j + 1 < job.n) ? 2 : 1;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (r >= nrow) break;
            if (job.src.rows) {
                const bool zeroed = warp_source_row<DIM>(job.src, (int64_t)o * job.n + 2 * j + r, lane, v[r]);
                if (zeroed && lane == 0 && job.src.nan_rows) atomicAdd(job.src.nan_rows, 1);
                continue;
            }
            float *q = job.vecs + ((size_t)o * job.n + 2 * j + r) * DIM;
#pragma unroll
            for (int b = 0; b < NB; ++b) v[r][b] = *reinterpret_cast<const float4 *>(q + b * 128 + 4 * lane);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (r >= nrow) break;
            float *q = job.vecs + ((size_t)o * job.n + 2 * j + r) * DIM;
            if (job.mean) {                 // the (k, dim) mean rows stay in L1
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float4 g = ldg_f4(job.mean + (size_t)o * DIM + b * 128 + 4 * lane);
                    v[r][b].x = __fsub_rn(v[r][b].x, g.x); v[r][b].y = __fsub_rn(v[r][b].y, g.y);
                    v[r][b].z = __fsub_rn(v[r][b].z, g.z); v[r][b].w = __fsub_rn(v[r][b].w, g.w);
                }
            }
            warp_unit_row<DIM>(v[r], scratch[warp], lane);
            if (o >= job.keep) continue;       // only feeds the pair sum below
#pragma unroll
            for (int b = 0; b < NB; ++b) *reinterpret_cast<float4 *>(q + b * 128 + 4 * lane) = v[r][b];
        }
        if (want_norms && o < job.keep) {
            // norms = 1 - row . mbar, fp64: both rows of the pair against ONE read of mbar from shared memory (8 KB per
            // read: with a read per row this kernel's shared-memory traffic, not HBM, set its time once the level-0 rows
            // came in as fp16); per-row accumulation order unchanged
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const double *m = mb + b * 128 + 4 * lane;
                const double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
                acc0 = fma((double)v[0][b].x, m0, acc0); acc0 = fma((double)v[0][b].y, m1, acc0);
                acc0 = fma((double)v[0][b].z, m2, acc0); acc0 = fma((double)v[0][b].w, m3, acc0);
                if (nrow == 2) {
                    acc1 = fma((double)v[1][b].x, m0, acc1); acc1 = fma((double)v[1][b].y, m1, acc1);
                    acc1 = fma((double)v[1][b].z, m2, acc1); acc1 = fma((double)v[1][b].w, m3, acc1);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
            }
            if (lane == 0) {
                job.norms[(size_t)o * job.n + 2 * j] = __fsub_rn(1.0f, (float)acc0);
                if (nrow == 2) job.norms[(size_t)o * job.n + 2 * j + 1] = __fsub_rn(1.0f, (float)acc1);
            }
        }
        if (job.next && j < half) {
            float *h = job.next + ((size_t)o * half + j) * DIM;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                float4 s4;
                s4.x = __fadd_rn(v[0][b].x, v[1][b].x); s4.y = __fadd_rn(v[0][b].y, v[1][b].y);
                s4.z = __fadd_rn(v[0][b].z, v[1][b].z); s4.w = __fadd_rn(v[0][b].w, v[1][b].w);
                *reinterpret_cast<float4 *>(h + b * 128 + 4 * lane) = s4;
            }
        }
    }
}

inline int rows_grid(int64_t max_rows)
{
    int64_t g = (max_rows + kWarpsPerCta - 1) / kWarpsPerCta;
    if (g < 1) g = 1;
    if (g > 148 * 32) g = 148 * 32;   // 32 resident CTAs of 8 warps would oversubscribe; rows loop
    return (int)g;
}

}  // namespace

#define SVX_DISPATCH_DIM(dim, CALL)            \
    switch (dim) {                             \
        case 128: { CALL(128); break; }        \
        case 256: { CALL(256); break; }        \
        case 512: { CALL(512); break; }        \
        case 1024: { CALL(1024); break; }      \
        default: break;                        \
    }

extern "C" int svx_normalize_rows(const SvxRows *jobs_d, const SvxRows *jobs_h, int njobs, int dim, void *stream)
{
    SVX_REQUIRE(svx_dim_supported(dim), SVX_ERR_UNSUPPORTED, "svx_normalize_rows: dim %d not in {128,256,512,1024}", dim);
    if (njobs <= 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int64_t mr = 0;
        for (int j = 0; j < nj; ++j) mr = jobs_h[j0 + j].nrows > mr ? jobs_h[j0 + j].nrows : mr;
        if (mr == 0) continue;
        dim3 grid(rows_grid(mr), nj);
#define CALL(D) k_normalize<D><<<grid, kWarpsPerCta * 32, 0, st>>>(jobs_d + j0)
        SVX_DISPATCH_DIM(dim, CALL)
#undef CALL
        SVX_LAUNCH_CHECK();
    }
    return SVX_OK;
}

extern "C" int svx_downsample(const SvxDownJob *jobs_d, const SvxDownJob *jobs_h, int njobs, int dim, void *stream)
{
    SVX_REQUIRE(svx_dim_supported(dim), SVX_ERR_UNSUPPORTED, "svx_downsample: dim %d unsupported", dim);
    if (njobs <= 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int kmax = 0; int64_t mr = 0;
        for (int j = 0; j < nj; ++j) {
            const SvxDownJob &jb = jobs_h[j0 + j];
            if (jb.k > kmax) kmax = jb.k;
            const int64_t r = (int64_t)jb.k * (jb.n / 2);
            if (r > mr) mr = r;
        }
        if (mr == 0) continue;
        dim3 ga(kmax * (dim / 128), nj);
        k_pairsum_colmean<<<ga, 128, 0, st>>>(jobs_d + j0, dim);
        SVX_LAUNCH_CHECK();
        dim3 gb(rows_grid(mr), nj);
#define CALL(D) k_center_normalize<D><<<gb, kWarpsPerCta * 32, 0, st>>>(jobs_d + j0)
        SVX_DISPATCH_DIM(dim, CALL)
#undef CALL
        SVX_LAUNCH_CHECK();
    }
    return SVX_OK;
}

extern "C" int svx_sample_norms(const SvxNormJob *jobs_d, const SvxNormJob *jobs_h, int njobs, int dim, void *stream)
{
    SVX_REQUIRE(svx_dim_supported(dim), SVX_ERR_UNSUPPORTED, "svx_sample_norms: dim %d unsupported", dim);
    if (njobs <= 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int j = 0; j < njobs; ++j)
        SVX_REQUIRE(jobs_h[j].no > 0 && jobs_h[j].ko * jobs_h[j].per > 0, SVX_ERR_ARG,
                    "svx_sample_norms: job %d has no samples (the caller fills norms with 1.0, dp_utils.py:356-357)", j);
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int64_t mr = 0;
        for (int j = 0; j < nj; ++j) {
            const int64_t r = (int64_t)jobs_h[j0 + j].k * jobs_h[j0 + j].n;
            if (r > mr) mr = r;
        }
        dim3 g1(dim / 128, nj);
        k_sample_mean<<<g1, 128, 0, st>>>(jobs_d + j0, dim);
        SVX_LAUNCH_CHECK();
        if (mr == 0) continue;
        dim3 g2(rows_grid(mr), nj);
#define CALL(D) k_norms_gemv<D><<<g2, kWarpsPerCta * 32, 0, st>>>(jobs_d + j0)
        SVX_DISPATCH_DIM(dim, CALL)
#undef CALL
        SVX_LAUNCH_CHECK();
    }
    return SVX_OK;
}

extern "C" int svx_score_pairs(const SvxScoreJob *jobs_d, const SvxScoreJob *jobs_h, int njobs, int dim, int mode,
                               void *stream)
{
    SVX_REQUIRE(dim > 0 && dim % 32 == 0, SVX_ERR_UNSUPPORTED, "svx_score_pairs: dim %d must be a multiple of 32", dim);
    if (njobs <= 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int sort_rows = 0;
    for (int j = 0; j < njobs; ++j) {
        const SvxScoreJob &jb = jobs_h[j];
        if (jb.perm && jb.xi && !jb.dots && jb.ne <= kSortMaxRows && jb.ne > sort_rows) sort_rows = jb.ne;
    }
    if (sort_rows > 0) {
        const size_t smem = (size_t)sort_rows * sizeof(int);
        if (smem > 48 * 1024)
            SVX_CUDA_OK(cudaFuncSetAttribute(k_sort_samples, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sort_samples<<<njobs, 512, smem, st>>>(jobs_d);
        SVX_LAUNCH_CHECK();
    }
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int ms = 0;
        for (int j = 0; j < nj; ++j) ms = jobs_h[j0 + j].nsamp > ms ? jobs_h[j0 + j].nsamp : ms;
        if (ms == 0) continue;
        dim3 grid((ms + 127) / 128, nj);
        if (mode == SVX_COST_EXACT) k_score_pairs<true><<<grid, kScWarps * 32, 0, st>>>(jobs_d + j0, dim);
        else k_score_pairs<false><<<grid, kScWarps * 32, 0, st>>>(jobs_d + j0, dim);
        SVX_LAUNCH_CHECK();
    }
    return SVX_OK;
}

extern "C" int svx_del_knob(const SvxScoreJob *jobs_d, const SvxScoreJob *jobs_h, int njobs, double frac, void *stream)
{
    (void)jobs_h;
    if (njobs <= 0) return SVX_OK;
    k_del_knob<<<njobs, 256, 0, (cudaStream_t)stream>>>(jobs_d, frac);
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

extern "C" int svx_host_del_knob(const float *scores, int n, double frac, double *del_penalty)
{
    SVX_REQUIRE(scores && del_penalty && n > 0, SVX_ERR_ARG, "svx_host_del_knob: bad arguments");
    float mx = scores[0];
    for (int i = 1; i < n; ++i) if (scores[i] > mx) mx = scores[i];   // Python max(): first maximal element
    unsigned int hist[SVX_KNOB_BINS];
    for (int i = 0; i < SVX_KNOB_BINS; ++i) hist[i] = 0u;
    if (mx > 0.0f) {
        const float step = mx / (float)SVX_KNOB_BINS;
        for (int i = 0; i < n; ++i) {
            const int b = svx_knob_bin(scores[i], mx, step);
            if (b >= 0) hist[b]++;
        }
    }
    static thread_local double dens[SVX_KNOB_BINS];
    long long total = 0;
    for (int i = 0; i < SVX_KNOB_BINS; ++i) total += hist[i];
    if (mx > 0.0f) {
        const float step = mx / (float)SVX_KNOB_BINS;
        for (int i = 0; i < SVX_KNOB_BINS; ++i) dens[i] = svx_knob_density(hist[i], i, step, mx, total);
    }
    *del_penalty = svx_knob_finish(dens, mx, frac);
    return SVX_OK;
}

extern "C" int svx_level_prologue(const SvxLevelJob *jobs_d, const SvxLevelJob *jobs_h, int njobs, int dim, void *stream)
{
    SVX_REQUIRE(svx_dim_supported(dim), SVX_ERR_UNSUPPORTED, "svx_level_prologue: dim %d not in {128,256,512,1024}", dim);
    if (njobs <= 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int j = 0; j < njobs; ++j) {
        SVX_REQUIRE(!jobs_h[j].idx || jobs_h[j].ko * jobs_h[j].per <= kMaxLevelSamples, SVX_ERR_UNSUPPORTED,
                    "svx_level_prologue: job %d draws %d samples (max %d)", j, jobs_h[j].ko * jobs_h[j].per, kMaxLevelSamples);
        SVX_REQUIRE(!(jobs_h[j].src.rows || jobs_h[j].osrc.rows) || (!jobs_h[j].mean && !jobs_h[j].other_mean), SVX_ERR_ARG,
                    "svx_level_prologue: job %d reads its rows from a source, which only the first level can (mean rows must be NULL)", j);
    }
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int kmax = 0, max_samples = 0; int64_t mp = 0; bool any_mean = false, any_samples = false;
        for (int j = 0; j < nj; ++j) {
            const SvxLevelJob &jb = jobs_h[j0 + j];
            if (jb.k > kmax) kmax = jb.k;
            const int64_t pairs = (int64_t)jb.k * ((jb.n + 1) / 2);
            if (pairs > mp) mp = pairs;
            any_mean |= jb.mean != nullptr && jb.n > 0;
            const bool samp = jb.idx != nullptr && jb.norms != nullptr && jb.ko * jb.per > 0 && jb.no > 0 && jb.n > 0;
            any_samples |= samp;
            if (samp && jb.ko * jb.per > max_samples) max_samples = jb.ko * jb.per;
        }
        if (any_mean && kmax > 0) {
            dim3 g(kmax * (dim / 128), nj);
            k_level_colmean<<<g, 128, 0, st>>>(jobs_d + j0, dim);
            SVX_LAUNCH_CHECK();
        }
        if (any_samples) {
            dim3 ga((max_samples + 7) / 8, nj);
#define CALL(D) k_level_sample_den<D><<<ga, 256, 0, st>>>(jobs_d + j0)
            SVX_DISPATCH_DIM(dim, CALL)
#undef CALL
            SVX_LAUNCH_CHECK();
            dim3 gb(dim / 128, nj);
            k_level_sample_acc<<<gb, 128, 0, st>>>(jobs_d + j0, dim);
            SVX_LAUNCH_CHECK();
        }
        if (mp > 0) {
            // a CTA starts by copying mbar (8 KB) into shared memory: with one row pair per warp that start-up
            // was ~10 % of the kernel (measured: 17.1 -> 15.1 ms on 256 pairs of 2000 x 2000 with 8 pairs per
            // warp).  Small launches keep one pair per warp so that they still fill the SMs.
            int64_t per_warp = (int64_t)nj * mp / ((int64_t)kFinWarps * 148 * 8);
            per_warp = per_warp < 1 ? 1 : (per_warp > kFinPairs ? kFinPairs : per_warp);
            int64_t gx = (mp + kFinWarps * per_warp - 1) / (kFinWarps * per_warp);
            if (gx > 148 * 32) gx = 148 * 32;
            dim3 g((unsigned)(gx < 1 ? 1 : gx), nj);
#define CALL(D) k_level_finish<D><<<g, kFinWarps * 32, 0, st>>>(jobs_d + j0)
            SVX_DISPATCH_DIM(dim, CALL)
#undef CALL
            SVX_LAUNCH_CHECK();
        }
    }
    return SVX_OK;
}

// ---------------------------------------------------------------------------------------------
// Descriptor upload without the DMA engines.  A batch's job descriptors and sample indices are a few
// MB in pinned (UVA-mapped) host memory; issued as a cudaMemcpyAsync they queue behind the bulk
// embedding copies of the following batches and delay this batch's first kernel by the whole
// transfer.  A kernel that reads the pinned buffer over PCIe is ordered only by its own stream.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_upload(uint4 *dst, const uint4 *src, long long n16)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}
}  // namespace

extern "C" int svx_upload_pinned(void *dst_d, const void *src_pinned_h, long long nbytes, void *stream)
{
    if (nbytes <= 0) return SVX_OK;
    SVX_REQUIRE(((uintptr_t)dst_d & 15) == 0 && ((uintptr_t)src_pinned_h & 15) == 0, SVX_ERR_ARG, "svx_upload_pinned: 16-byte alignment required");
    const long long n16 = (nbytes + 15) / 16;      // both buffers are padded to 256-byte multiples by the caller
    int grid = (int)((n16 + 255) / 256);
    if (grid > 148 * 4) grid = 148 * 4;
    k_upload<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4 *>(dst_d), reinterpret_cast<const uint4 *>(src_pinned_h), n16);
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

// ---------------------------------------------------------------------------------------------
// svx_gather_doc_embedding: a warp per output row; the source row (2 KB fp16 or 4 KB fp32) is read
// with 16-byte loads, checked for NaNs with a warp vote, widened and written as fp32.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_gather_rows(const SvxGatherJob *jobs, int dim)
{
    const SvxGatherJob job = jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total = (int64_t)job.k * job.n;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < total; r += (int64_t)gridDim.x * 8) {
        const int src = job.table ? job.table[r] : (int)r;      // no table: identity (a plain widening copy)
        float *dst = job.out + r * dim;
        const bool have = src >= 0 && src < job.nrows;
        bool bad = false;
        for (int d0 = 0; d0 < dim; d0 += 256) {              // 32 lanes x 8 elements
            const int d = d0 + 8 * lane;
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = 0.0f;
            if (have && d < dim) {
                if (job.is_fp16) {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(job.rows) + (size_t)src * dim + d));
                    const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const float2 f = __half22float2(h[u]); v[2 * u] = f.x; v[2 * u + 1] = f.y; }
                } else {
                    const float *p = reinterpret_cast<const float *>(job.rows) + (size_t)src * dim + d;
                    const float4 a = ldg_f4(p), b = ldg_f4(p + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) bad |= (v[u] != v[u]);
            }
            if (d < dim) {
                *reinterpret_cast<float4 *>(dst + d) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(dst + d + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        if (__any_sync(0xffffffffu, bad)) {                  // embedding_utils.py:196-200: reset to zero
            for (int d = 4 * lane; d < dim; d += 128) *reinterpret_cast<float4 *>(dst + d) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane == 0 && job.nan_rows) atomicAdd(job.nan_rows, 1);
        }
    }
}
}  // namespace

extern "C" int svx_gather_doc_embedding(const SvxGatherJob *jobs_d, const SvxGatherJob *jobs_h, int njobs, int dim, void *stream)
{
    SVX_REQUIRE(dim > 0 && dim % 8 == 0, SVX_ERR_UNSUPPORTED, "svx_gather_doc_embedding: dim %d must be a multiple of 8", dim);
    if (njobs <= 0) return SVX_OK;
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int64_t mr = 0;
        for (int j = 0; j < nj; ++j) {
            const int64_t r = (int64_t)jobs_h[j0 + j].k * jobs_h[j0 + j].n;
            if (r > mr) mr = r;
        }
        if (mr == 0) continue;
        dim3 grid(rows_grid(mr), nj);
        k_gather_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(jobs_d + j0, dim);
        SVX_LAUNCH_CHECK();
    }
    return SVX_OK;
}
