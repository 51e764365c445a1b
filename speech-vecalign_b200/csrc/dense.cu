// dense.cu — coarsest level (sm_100a): the full 1-1 cost matrix on CUDA cores in the reference's
// summation order, the 3-way anti-diagonal wavefront DP, its traceback and the search path of the
// next finer level.
//   svx_dense_costs  (dp_core.pyx:36-77  make_dense_costs)   [tcgen05 3xTF32 variant: dense_tc.cu]
//   svx_dense_dp     (dp_core.pyx:79-141 dense_dp, dp_utils.py:146-174 dense_traceback,
//                     dp_utils.py:177-275 path glue)
#include "svx_common.cuh"
#include "svx_dp.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Dense costs: 64x64 output tile per CTA, 256 threads, 4x4 outputs per thread (rows t, t+16, t+32,
// t+48 so that the 8 lanes of one LDS.128 phase hit 8 consecutive rows = 8 distinct bank groups with
// the 36-float row stride): 8 LDS.128 feed 64 multiply-adds, which keeps the FP32 pipe - not shared
// memory - the limit.  32-float slices arrive by cp.async into a double buffer.  Each accumulator
// adds its products in increasing d: the reference order.
// ---------------------------------------------------------------------------------------------
constexpr int kDT = 64;       // tile edge
constexpr int kDC = 32;       // floats of the embedding dimension staged per step
constexpr int kDS = kDC + 4;  // padded row stride

template <bool EXACT>
__device__ __forceinline__ void mac4(float &acc, const float4 &a, const float4 &b)
{
    if (EXACT) {
        acc = __fadd_rn(acc, __fmul_rn(a.x, b.x)); acc = __fadd_rn(acc, __fmul_rn(a.y, b.y));
        acc = __fadd_rn(acc, __fmul_rn(a.z, b.z)); acc = __fadd_rn(acc, __fmul_rn(a.w, b.w));
    } else {
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(256) k_dense_costs(const SvxDenseJob *jobs, int dim)
{
    __shared__ __align__(16) float xs[2][kDT * kDS];
    __shared__ __align__(16) float ys[2][kDT * kDS];
    const SvxDenseJob job = jobs[blockIdx.z];
    const int x0 = blockIdx.y * kDT, y0 = blockIdx.x * kDT;
    if (x0 >= job.s0 || y0 >= job.s1) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;       // y rows tx + 16 j; x rows ty + 16 i
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // this thread stages 2 pieces (row r, 16-byte piece c4) of each operand per slice
    auto issue = [&](int sl) {
        const int d0 = sl * kDC;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int f = tid + it * 256;
            const int r = f >> 3, c4 = f & 7;
            const bool okx = x0 + r < job.s0, oky = y0 + r < job.s1;
            const float *gx = okx ? job.v0 + (size_t)(x0 + r) * dim + d0 + 4 * c4 : job.v0;
            const float *gy = oky ? job.v1 + (size_t)(y0 + r) * dim + d0 + 4 * c4 : job.v1;
            const unsigned dx = (unsigned)__cvta_generic_to_shared(&xs[sl & 1][r * kDS + 4 * c4]);
            const unsigned dy = (unsigned)__cvta_generic_to_shared(&ys[sl & 1][r * kDS + 4 * c4]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dx), "l"(gx), "r"(okx ? 16 : 0));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dy), "l"(gy), "r"(oky ? 16 : 0));
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    const int slices = dim / kDC;
    issue(0);
    for (int sl = 0; sl < slices; ++sl) {
        if (sl + 1 < slices) {
            issue(sl + 1);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const float *bx = xs[sl & 1], *by = ys[sl & 1];
#pragma unroll 2
        for (int d = 0; d < kDC; d += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = *reinterpret_cast<const float4 *>(bx + (ty + 16 * i) * kDS + d);
                b[i] = *reinterpret_cast<const float4 *>(by + (tx + 16 * i) * kDS + d);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mac4<EXACT>(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + ty + 16 * i;
        if (x >= job.s0) continue;
        const float nx = job.n0[x];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = y0 + tx + 16 * j;
            if (y < job.s1) {
                job.costs[(size_t)x * job.s1 + y] = svx_dense_cost(acc[i][j], nx, job.n1[y]);
                if (job.dots) job.dots[(size_t)x * job.s1 + y] = acc[i][j];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dense DP: one CTA per job; nodes of anti-diagonal k = r + c are independent.  Three rotating
// diagonal buffers in shared memory are indexed by the coordinate of the SHORTER side (<= 300
// with the default max_size_full_dp, whatever the aspect ratio).  Warp 0 then walks the
// backpointers and lays the next level's search path, lanes expanding each slanted segment.
// ---------------------------------------------------------------------------------------------
struct WarpEmit {
    int32_t *ypath;
    int path_len;
    int lane;
    __host__ __device__ void operator()(long long xs, long long ys, long long xw, long long yw) const
    {
        const long long nn = xw + yw;
        for (long long i = 1 + lane; i <= nn; i += 32) {
            int x, y;
            svx_slant_point(xs, ys, xw, yw, i, &x, &y);
            if (x + y < path_len) ypath[x + y] = y;
        }
    }
};

__global__ void __launch_bounds__(256) k_dense_dp(const SvxDenseJob *jobs)
{
    extern __shared__ double diag[];   // 3 * (min(s0,s1)+1)
    const SvxDenseJob job = jobs[blockIdx.x];
    const int s0 = job.s0, s1 = job.s1;
    const bool swap = s0 > s1;               // p runs along the shorter side
    const int sp = swap ? s1 : s0, sq = swap ? s0 : s1;
    const int L = sp + 1;
    const float penf = (float)(*job.del_penalty);
    const int ld = s1 + 1;
    // The cost of a thread's cell on diagonal k + 1 is requested BEFORE the barrier that ends diagonal k (the
    // loads of one anti-diagonal are s1-1 floats apart: one L2 sector each, ~500 cycles that every warp then
    // waited for at the next barrier).  volatile asm: ptxas otherwise sinks the load to its use.
    const float *costs = job.costs;
    auto cost_prefetch = [&](int k, int p) -> float {
        const int q = k - p;
        const int r = swap ? q : p, c = swap ? p : q;
        float v = 0.0f;
        if (p <= sp && q >= 0 && q <= sq && r > 0 && c > 0)
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(costs + (size_t)(r - 1) * s1 + (c - 1)));
        return v;
    };
    const bool single = L <= (int)blockDim.x;          // one cell per thread and diagonal
    float cnext = single ? cost_prefetch(0, (int)threadIdx.x) : 0.0f;
    for (int k = 0; k <= s0 + s1; ++k) {
        double *cur = diag + (k % 3) * L;
        const double *d1 = diag + ((k + 2) % 3) * L;   // diagonal k-1
        const double *d2 = diag + ((k + 1) % 3) * L;   // diagonal k-2
        const int plo = k - sq > 0 ? k - sq : 0;
        const int phi = k < sp ? k : sp;
        const float cthis = cnext;
        if (single) cnext = cost_prefetch(k + 1, (k + 1 - sq > 0 ? k + 1 - sq : 0) + (int)threadIdx.x);   // p of this thread's cell on k + 1
        for (int p = plo + (int)threadIdx.x; p <= phi; p += blockDim.x) {
            const int q = k - p;
            const int r = swap ? q : p, c = swap ? p : q;
            double val; int bp;
            if (r == 0 && c == 0) { val = 0.0; bp = 4; }
            else if (r == 0) { val = svx_dense_boundary(c, penf); bp = 1; }
            else if (c == 0) { val = svx_dense_boundary(r, penf); bp = 2; }
            else {
                // (p,q-1) and (p-1,q) live on diagonal k-1 at slots p and p-1
                const double dg = d2[p - 1];
                const double left = swap ? d1[p - 1] : d1[p];   // csum[r, c-1]
                const double up = swap ? d1[p] : d1[p - 1];     // csum[r-1, c]
                val = svx_dense_cell(dg, left, up, single ? cthis : costs[(size_t)(r - 1) * s1 + (c - 1)], penf, &bp);
            }
            cur[p] = val;
            job.bp[(size_t)r * ld + c] = (uint8_t)bp;
            if (job.csum) job.csum[(size_t)r * ld + c] = val;
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        WarpEmit emit{job.ypath, job.path_len, lane};
        SvxPathBuilder<WarpEmit> pb(emit);
        pb.begin(s0, s1, job.t0, job.t1, job.upsample);
        const int st = svx_dense_walk(job.bp, s0, s1, pb);
        if (lane == 0) {
            if (job.path_len > 0) job.ypath[0] = 0;
            *job.status_d = st;
        }
    }
}

}  // namespace

int svx_dense_costs_tc_launch(const SvxDenseJob *jobs_d, const SvxDenseJob *jobs_h, int njobs, int dim, cudaStream_t st);

extern "C" int svx_dense_costs(const SvxDenseJob *jobs_d, const SvxDenseJob *jobs_h, int njobs, int dim, int mode,
                               void *stream)
{
    if (mode == SVX_COST_TC) return njobs > 0 ? svx_dense_costs_tc_launch(jobs_d, jobs_h, njobs, dim, (cudaStream_t)stream) : SVX_OK;
    SVX_REQUIRE(dim > 0 && dim % kDC == 0, SVX_ERR_UNSUPPORTED, "svx_dense_costs: dim %d must be a multiple of %d", dim, kDC);
    if (njobs <= 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int m0 = 0, m1 = 0;
        for (int j = 0; j < nj; ++j) {
            if (jobs_h[j0 + j].s0 > m0) m0 = jobs_h[j0 + j].s0;
            if (jobs_h[j0 + j].s1 > m1) m1 = jobs_h[j0 + j].s1;
        }
        if (m0 == 0 || m1 == 0) continue;
        const int gy = (m0 + kDT - 1) / kDT, gx = (m1 + kDT - 1) / kDT;
        SVX_REQUIRE(gy <= 65535, SVX_ERR_UNSUPPORTED, "svx_dense_costs: s0 %d too large", m0);
        dim3 grid(gx, gy, nj);
        if (mode == SVX_COST_EXACT) k_dense_costs<true><<<grid, 256, 0, st>>>(jobs_d + j0, dim);
        else k_dense_costs<false><<<grid, 256, 0, st>>>(jobs_d + j0, dim);
        SVX_LAUNCH_CHECK();
    }
    return SVX_OK;
}

extern "C" int svx_dense_dp(const SvxDenseJob *jobs_d, const SvxDenseJob *jobs_h, int njobs, void *stream)
{
    if (njobs <= 0) return SVX_OK;
    int lmax = 0;
    for (int j = 0; j < njobs; ++j) {
        const int m = jobs_h[j].s0 < jobs_h[j].s1 ? jobs_h[j].s0 : jobs_h[j].s1;
        if (m + 1 > lmax) lmax = m + 1;
    }
    const size_t smem = (size_t)3 * lmax * sizeof(double);
    SVX_REQUIRE(smem <= 200 * 1024, SVX_ERR_UNSUPPORTED,
                "svx_dense_dp: min(s0,s1)=%d needs %zu B of shared memory (max_size_full_dp too large)", lmax - 1, smem);
    if (smem > 48 * 1024)
        SVX_CUDA_OK(cudaFuncSetAttribute(k_dense_dp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dense_dp<<<njobs, 256, smem, (cudaStream_t)stream>>>(jobs_d);
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

extern "C" int svx_path_len(int c0, int c1, int t0, int t1, int upsample)
{
    return svx_path_len_impl(c0, c1, t0, t1, upsample);
}

// Host twin: same cell function, same walk, serial loops.  All pointers are host pointers.
extern "C" int svx_host_dense_dp(const SvxDenseJob *job)
{
    SVX_REQUIRE(job && job->bp && job->ypath && job->status_d && job->del_penalty, SVX_ERR_ARG, "svx_host_dense_dp: null pointer");
    const int s0 = job->s0, s1 = job->s1, ld = s1 + 1;
    const float penf = (float)(*job->del_penalty);
    double *cs = job->csum;
    double *tmp = nullptr;
    if (!cs) { tmp = new double[(size_t)(s0 + 1) * ld]; cs = tmp; }
    for (int r = 0; r <= s0; ++r) {
        for (int c = 0; c <= s1; ++c) {
            double val; int bp;
            if (r == 0 && c == 0) { val = 0.0; bp = 4; }
            else if (r == 0) { val = svx_dense_boundary(c, penf); bp = 1; }
            else if (c == 0) { val = svx_dense_boundary(r, penf); bp = 2; }
            else val = svx_dense_cell(cs[(size_t)(r - 1) * ld + c - 1], cs[(size_t)r * ld + c - 1],
                                      cs[(size_t)(r - 1) * ld + c], job->costs[(size_t)(r - 1) * s1 + c - 1], penf, &bp);
            cs[(size_t)r * ld + c] = val;
            job->bp[(size_t)r * ld + c] = (uint8_t)bp;
        }
    }
    SvxSerialEmit emit{job->ypath, job->path_len};
    SvxPathBuilder<SvxSerialEmit> pb(emit);
    pb.begin(s0, s1, job->t0, job->t1, job->upsample);
    *job->status_d = svx_dense_walk(job->bp, s0, s1, pb);
    if (job->path_len > 0) job->ypath[0] = 0;
    delete[] tmp;
    return SVX_OK;
}
