// banded_p2.cu — level-0 banded costs (dp_core.pyx:165-267 make_sparse_costs) for the standard type
// sets (make_alignment_types, vecalign.py:154-162), register-blocked AND packed: sm_100a's FFMA2.
//
// A thread owns a 2 x 2 block of POSITIONS (xx in {2X, 2X+1}, yy in {2Y, 2Y+1}); its cells lie on the
// anti-diagonals 2e, 2e+1, 2e+1, 2e+2 (e = X + Y is the block-diagonal) and share their operand rows.
// The two cells that share a y position are one packed fp32x2 accumulator {(2X, y), (2X+1, y)}:
//
//     FAST   acc = fma.rn.f32x2(x2, {y, y}, acc)                                   1 FFMA2 per 2 MACs
//     EXACT  p   = fma.rn.f32x2(x2, {y, y}, {-0.0, -0.0})   == round(x * y)        (the reference's order:
//            acc = fma.rn.f32x2(p, {1.0, 1.0}, acc)         == round(p + acc)       multiply, THEN add)
//
// with x2 = {x[2X][d], x[2X+1][d]}.  -0.0 and 1.0 are kernel ARGUMENTS: ptxas cannot fold them, so the two
// roundings stay separate (a literal -0.0 / 1.0 is simplified and contracted into one FFMA2, which is not
// the reference's arithmetic).  x * y + (-0.0) rounds once to round(x * y) with the product's own sign of
// zero, p * 1.0 + acc is round(p + acc): bit-identical to __fmul_rn + __fadd_rn (tests/test_gpu_functions.py).
// FFMA2 takes the {y, y} operand as a scalar broadcast (SASS `R.F32`), so no duplication moves are issued,
// and it occupies the FP32 pipe for two cycles per issue: the exact stream needs 1 issue slot per MAC
// instead of 2, which is what bounded the scalar kernel (issue-active 82 %, FMA pipe 67 %).
//
// Tile = ne (about 32 CW / (B/2 + 1), see the consumer mapping) consecutive block-diagonals of one job (2 ne + 1 anti-diagonals, the first and last shared
// with the neighbouring tiles cell by cell - every band cell belongs to exactly one block, so to exactly one tile).
// B/2 + 1 band slots (Y - Ymin(e)) cover every band cell of a block-diagonal (proved by enumeration in
// tests/test_host_logic.py); the consumer threads own the (block-diagonal, slot) pairs of the tile.
// Shared-memory layout of one embedding slice (BC floats): x rows as PAIR ROWS, the two positions of a block
// interleaved float by float ({x[2X][d], x[2X+1][d]} is then one aligned 8-byte word = one FFMA2 operand),
// staged with 4-byte cp.async; y rows as [even positions | odd positions], 16-byte cp.async; both through a
// ring of `nstages` slices with one __syncthreads per slice.
#include <limits.h>
#include "svx_common.cuh"
#include "svx_banded_p2.h"

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <bool EXACT>
__device__ __forceinline__ u64 mac2(u64 acc, u64 x2, float y, u64 one2, u64 nz2)
{
    const u64 y2 = pack2(y, y);
    if (EXACT) return fma2(fma2(x2, y2, nz2), one2, acc);
    return fma2(x2, y2, acc);
}


__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// arrives once every cp.async this thread has issued so far has landed (count pre-charged at init: .noinc)
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// K: overlaps per side; BC: floats of the embedding dimension per staged slice; DV: floats per operand load.
// Warps [0, 8) are consumers, the remaining 1-4 warps are producers: they move
// slice after slice of the tile's rows into a ring of `nstages` buffers with cp.async and signal each buffer's
// `full` mbarrier (cp.async.mbarrier.arrive); a consumer warp waits for `full`, multiplies, and arrives on the
// buffer's `empty` mbarrier.  No CTA-wide barrier in the slice loop: the consumer warps drift apart by up to the
// ring depth, and the copy address arithmetic stays off the consumers' FP32 pipe (IMAD shares it with FFMA2).
// CW consumer warps (8 = two per scheduler, 168 registers; 12 = three per scheduler, 128 registers: K <= 4 only - the third
// warp covers the FFMA2 dependency stalls the other two leave open, ncu: "wait" + "math throttle" on every other sample)
template <int K, int BC, int DV, bool EXACT, int CW>
__global__ void __launch_bounds__(32 * (CW + 4), 1)
k_banded_costs_p2(const SvxBandJob *jobs, int dim, int nstages, int nb, int ne, int grouped, float one, float nz)
{
    constexpr int T = K * (K + 1) / 2;
    constexpr int XS = 2 * BC + 4;      // pair-row stride in floats: 8 consecutive pair rows tile the 32 banks
    constexpr int YS = BC + 4;
    static_assert(DV == 2 || DV == 4, "operand loads are 8 or 16 bytes per row");
    extern __shared__ __align__(16) float tile[];
    const SvxBandJob &job = jobs[blockIdx.y];
    const int A = job.a_len;
    const int e_first = ne * (int)blockIdx.x - 1;            // e = -1 holds the (odd, odd) cells of diagonal 0
    const int d_first = max(2 * e_first, 0), d_last = min(2 * (e_first + ne - 1) + 2, A - 1);
    if (d_first > d_last) return;
    const int B = job.band, w = job.width_over2;
    const int s0 = job.s0, s1 = job.s1;
    const int32_t *ypath = job.ypath;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int kConsWarps = CW;
    const int nprod = (nthreads >> 5) - kConsWarps;

    // positions touched by the tile, rounded out to whole blocks.  The search path is monotone (x or y advances
    // by one per anti-diagonal), so the first / last diagonal bound every band offset of the tile.
    const int bf = ypath[d_first] - w, bl = ypath[d_last] - w;
    const int ylo = (bf >> 1) * 2, yhi = ((bl + B - 1) >> 1) * 2 + 1;
    const int xlo = ((d_first - (bf + B - 1)) >> 1) * 2, xhi = ((d_last - bl) >> 1) * 2 + 1;
    const int NX = xhi - xlo + 1, NY = yhi - ylo + 1, HX = NX >> 1, HY = NY >> 1;
    const int rows_cap = svx_p2_rows_cap(B, ne);
    if (NX <= 0 || NY <= 0 || NX + NY > rows_cap) return;    // not a search path (an earlier level failed: status_d says so)
    const int stage_floats = K * rows_cap * YS;
    // after the ring: the copy lists - one entry {source float offset from v0 / v1, byte offset in a stage} per staged row
    // that HAS a source row; rows outside the documents / overlaps are zeroed once in every stage and never copied -
    // then the mbarriers
    int2 *xlist = reinterpret_cast<int2 *>(tile + (size_t)nstages * stage_floats);
    int2 *ylist = xlist + K * rows_cap;
    int *counts = reinterpret_cast<int *>(ylist + K * rows_cap);                          // [0] x rows, [1] y rows
    const unsigned bars = (unsigned)__cvta_generic_to_shared(counts + 2);                 // full[nstages], empty[nstages]
    const int xstride = HX * XS;       // floats between overlaps in the x region
    const int ystride = NY * YS;
    const int ybase = K * HX * XS;     // the y region follows the x region
    if (tid < 2) counts[tid] = 0;
    if (tid < 2 * nstages) mbar_init(bars + 8 * tid, tid < nstages ? 32 * nprod : kConsWarps);
    __syncthreads();

    // x: slot k * NX + p is position xlo + p of overlap k, written into pair row p / 2 at float 2 d + (p & 1);
    // y: slot K * NX + k * NY + q is position ylo + 2 q (q < HY) or ylo + 2 (q - HY) + 1 of overlap k
    const int xrows = K * NX, yrows = K * NY;
    for (int r = tid; r < xrows + yrows; r += nthreads) {
        int off = -1, dst;
        const bool isx = r < xrows;
        if (isx) {
            const int k = r / NX, p = r - k * NX, seg = xlo + p;
            if (k < job.k0 && seg >= 0 && seg < s0) off = (int)(((size_t)k * s0 + seg) * dim);
            dst = k * xstride + (p >> 1) * XS + (p & 1);
        } else {
            const int r2 = r - xrows;
            const int k = r2 / NY, q = r2 - k * NY;
            const int seg = ylo + (q < HY ? 2 * q : 2 * (q - HY) + 1);
            if (k < job.k1 && seg >= 0 && seg < s1) off = (int)(((size_t)k * s1 + seg) * dim);
            dst = ybase + r2 * YS;
        }
        if (off >= 0) {
            const int slot = atomicAdd(&counts[isx ? 0 : 1], 1);
            (isx ? xlist : ylist)[slot] = make_int2(off, dst * (int)sizeof(float));
        } else {
            for (int st = 0; st < nstages; ++st)
                for (int d = 0; d < BC; ++d) tile[(size_t)st * stage_floats + dst + (isx ? 2 * d : d)] = 0.0f;
        }
    }
    __syncthreads();
    const unsigned tile_u32 = (unsigned)__cvta_generic_to_shared(tile);
    const int slices = dim / BC;

    if (warp >= kConsWarps) {
        // ---- producers: lane = position inside a row slice, sub-rows of a warp instruction = consecutive list entries --------
        const int pw = warp - kConsWarps;
        const char *gx = reinterpret_cast<const char *>(job.v0), *gy = reinterpret_cast<const char *>(job.v1);
        constexpr int XR = 32 / BC;            // x rows per warp instruction (4-byte copies, lane = d)
        constexpr int YR = 32 / (BC / 4);      // y rows per warp instruction (16-byte copies)
        constexpr int UB = 8;                  // list entries fetched before their copies are issued
        const int xd = lane % BC, xsub = lane / BC;
        const int yc = lane % (BC / 4), ysub = lane / (BC / 4);
        const int nx = counts[0], ny = counts[1];
        const int xstep = nprod * XR, ystep = nprod * YR;
        for (int sl = 0; sl < slices; ++sl) {
            const int s = sl % nstages;
            if (sl >= nstages) mbar_wait(bars + 8 * (nstages + s), (unsigned)((sl / nstages - 1) & 1));
            const unsigned buf = tile_u32 + (unsigned)(s * stage_floats * (int)sizeof(float));
            {
                const unsigned dst0 = buf + (unsigned)(8 * xd);
                const char *src0 = gx + (size_t)(sl * BC + xd) * 4;
                int i = pw * XR + xsub;
                for (; i + (UB - 1) * xstep < nx; i += UB * xstep) {
                    int2 ent[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) ent[u] = xlist[i + u * xstep];
#pragma unroll
                    for (int u = 0; u < UB; ++u)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst0 + (unsigned)ent[u].y), "l"(src0 + (long long)ent[u].x * 4));
                }
                for (; i < nx; i += xstep) {
                    const int2 e = xlist[i];
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst0 + (unsigned)e.y), "l"(src0 + (long long)e.x * 4));
                }
            }
            {
                const unsigned dst0 = buf + (unsigned)(16 * yc);
                const char *src0 = gy + (size_t)(sl * BC + 4 * yc) * 4;
                int i = pw * YR + ysub;
                for (; i + (UB - 1) * ystep < ny; i += UB * ystep) {
                    int2 ent[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) ent[u] = ylist[i + u * ystep];
#pragma unroll
                    for (int u = 0; u < UB; ++u)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst0 + (unsigned)ent[u].y), "l"(src0 + (long long)ent[u].x * 4));
                }
                for (; i < ny; i += ystep) {
                    const int2 e = ylist[i];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst0 + (unsigned)e.y), "l"(src0 + (long long)e.x * 4));
                }
            }
            mbar_arrive_cp_async(bars + 8 * s);
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        return;
    }

    // ---- consumers: this thread's block ----------------------------------------------------------------------
    // Consumer thread -> (block-diagonal of the tile, band slot).  Both orders give every scheduler CW / 4 consumer warps
    // whatever the band width (a warp per slot gave 9 or 10 warps at bands 16 / 18: 3 + 2 + 2 + 2 per scheduler).
    //   grouped: 4 CW groups of eight lanes, group g owns slot g % nb of the eight consecutive block-diagonals
    //     8 (g / nb) ..: ne = 8 * (4 CW / nb).  The eight lanes of an LDS.128 phase - one group - read rows that are equal
    //     or consecutive (X and Y advance by 0 or 1 along e): no bank conflicts, but 4 CW mod nb groups idle.
    //   slot-fastest: thread t owns slot t % nb of block-diagonal t / nb: ne = 32 CW / nb, at most nb - 1 idle lanes, and
    //     two-way conflicts in the phases that straddle a block-diagonal when nb is not a multiple of 8.
    // The launcher takes the grouped order unless it idles more than 7 % of the lanes the other would use (measured:
    // band 18, nb = 10: grouped 2 % faster with 24 against 25 block-diagonals; band 16, nb = 9: slot-fastest 10 % faster
    // with 28 against 24; the two coincide at nb = 8).
    int ei, yi;
    if (grouped) {
        const int grp = tid >> 3, eb = grp / nb;
        yi = grp - eb * nb;
        ei = 8 * eb + (tid & 7);
    } else {
        ei = tid / nb;
        yi = tid - ei * nb;
    }
    const int e = e_first + ei;
    const bool in_tile = ei < ne;
    const int dA = 2 * e, dB = 2 * e + 1, dC = 2 * e + 2;
    const bool vA = in_tile && dA >= 0 && dA < A, vB = in_tile && dB >= 0 && dB < A, vC = in_tile && dC >= 0 && dC < A;
    const int bA = vA ? ypath[dA] - w : 0, bB = vB ? ypath[dB] - w : 0, bC = vC ? ypath[dC] - w : 0;
    // smallest Y with a band cell on one of the three diagonals: even yy on dA, both parities on dB, odd yy on dC
    int ymin = INT_MAX;
    if (vA) ymin = min(ymin, (bA + 1) >> 1);
    if (vB) ymin = min(ymin, bB >> 1);
    if (vC) ymin = min(ymin, bC >> 1);
    const int Y = ymin + yi, X = e - Y;
    const int yy0 = 2 * Y, xx0 = 2 * X;
    unsigned inband = 0;               // bit c: cell c is a band cell; c = 0 (xx0, yy0), 1 (xx0+1, yy0), 2 (xx0, yy0+1), 3 (xx0+1, yy0+1)
    int bslot[4] = {0, 0, 0, 0};
    if (vA) { const int b = yy0 - bA; if (b >= 0 && b < B) { inband |= 1u; bslot[0] = b; } }
    if (vB) {
        const int b = yy0 - bB;
        if (b >= 0 && b < B) { inband |= 2u; bslot[1] = b; }
        if (b + 1 >= 0 && b + 1 < B) { inband |= 4u; bslot[2] = b + 1; }
    }
    if (vC) { const int b = yy0 + 1 - bC; if (b >= 0 && b < B) { inband |= 8u; bslot[3] = b; } }
    const int xs = (xx0 - xlo) >> 1, ys = (yy0 - ylo) >> 1;       // pair row / even-half row of the block
    const bool active = inband != 0 && xs >= 0 && xs < HX && ys >= 0 && ys < HY;
    const bool warp_active = __any_sync(0xffffffffu, active);

    u64 acc_e[T], acc_o[T];            // {(xx0, y), (xx0 + 1, y)} for y = yy0 / yy0 + 1, per type
#pragma unroll
    for (int t = 0; t < T; ++t) acc_e[t] = acc_o[t] = 0ull;
    const u64 one2 = pack2(one, one), nz2 = pack2(nz, nz);
    // inactive lanes of an active warp read the block of lane-clamped rows (always staged) and drop the result
    const int xs_c = min(max(xs, 0), HX - 1), ys_c = min(max(ys, 0), HY - 1);

    for (int sl = 0; sl < slices; ++sl) {
        const int s = sl % nstages;
        mbar_wait(bars + 8 * s, (unsigned)((sl / nstages) & 1));
        if (warp_active) {
            const float *buf = tile + (size_t)s * stage_floats;
            const float *px = buf + (size_t)xs_c * XS;
            const float *pye = buf + ybase + (size_t)ys_c * YS;
            const float *pyo = pye + (size_t)HY * YS;
#pragma unroll 2
            for (int d = 0; d < BC; d += DV) {
                float ye[K][DV], yo[K][DV];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (DV == 4) {
                        const float4 a = *reinterpret_cast<const float4 *>(pye + j * ystride + d);
                        const float4 b = *reinterpret_cast<const float4 *>(pyo + j * ystride + d);
                        ye[j][0] = a.x; ye[j][1] = a.y; ye[j][DV - 2] = a.z; ye[j][DV - 1] = a.w;
                        yo[j][0] = b.x; yo[j][1] = b.y; yo[j][DV - 2] = b.z; yo[j][DV - 1] = b.w;
                    } else {
                        const float2 a = *reinterpret_cast<const float2 *>(pye + j * ystride + d);
                        const float2 b = *reinterpret_cast<const float2 *>(pyo + j * ystride + d);
                        ye[j][0] = a.x; ye[j][1] = a.y;
                        yo[j][0] = b.x; yo[j][1] = b.y;
                    }
                }
                int t = 0;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    u64 xp[DV];           // {x[xx0][d + q], x[xx0 + 1][d + q]}
                    {
                        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(px + i * xstride + 2 * d);
                        xp[0] = a.x; xp[1] = a.y;
                        if (DV == 4) {
                            const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(px + i * xstride + 2 * d + 4);
                            xp[DV - 2] = b.x; xp[DV - 1] = b.y;
                        }
                    }
#pragma unroll
                    for (int j = 0; i + j <= K - 1; ++j, ++t) {
#pragma unroll
                        for (int q = 0; q < DV; ++q) {
                            acc_e[t] = mac2<EXACT>(acc_e[t], xp[q], ye[j][q], one2, nz2);
                            acc_o[t] = mac2<EXACT>(acc_o[t], xp[q], yo[j][q], one2, nz2);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (nstages + s));
    }
    if (!active) return;

    // ---- cost formula + store (dp_core.pyx:259-260), anti-diagonal-major (A, T, B) ----------------------------
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (!((inband >> c) & 1)) continue;
        const int i = c & 1, j = c >> 1;
        const int xx = xx0 + i, yy = yy0 + j, d = 2 * e + i + j;
        const bool inside = xx >= 0 && xx < s0 && yy >= 0 && yy < s1;
        float *out = job.costs + (size_t)d * T * B + bslot[c];
        int t = 0;
#pragma unroll
        for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int ky = 0; kx + ky <= K - 1; ++ky, ++t) {
                float lo, hi;
                unpack2(j ? acc_o[t] : acc_e[t], lo, hi);
                float cst = INFINITY;
                if (inside)
                    cst = svx_band_cost(i ? hi : lo, kx + 1, ky + 1, job.n0[(size_t)kx * s0 + xx], job.n1[(size_t)ky * s1 + yy]);
                out[(size_t)t * B] = cst;
            }
    }
}

template <int K, int BC, int DV, int CW>
int launch_p2_t(const SvxBandJob *jobs_d, int nj, int max_alen, int band, int dim, int mode, int nstages, int nprod, cudaStream_t st)
{
    constexpr int kConsWarps = CW;
    const int nb = band / 2 + 1;                                   // band slots (2 x 2 blocks) per block-diagonal
    if (nb > 32 * kConsWarps) return -1;
    const int ne_grouped = 8 * (4 * kConsWarps / nb);              // block-diagonals per tile, see the consumer mapping
    int ne = 32 * kConsWarps / nb;
    if (ne > 5 * CW) ne = 5 * CW;
    const int grouped = ne_grouped * 100 >= ne * 93;
    if (grouped) ne = ne_grouped;
    if (nprod < 1) nprod = 1;
    if (nprod > 4) nprod = 4;
    const int threads = 32 * (kConsWarps + nprod);
    const int rows_cap = svx_p2_rows_cap(band, ne);
    const size_t stage_bytes = (size_t)K * rows_cap * (BC + 4) * sizeof(float);
    const size_t table = (size_t)2 * K * rows_cap * sizeof(int2) + 16 + 2 * 8 * 8;
    while (nstages > 2 && nstages * stage_bytes + table > 225 * 1024) --nstages;
    const size_t smem = nstages * stage_bytes + table;
    if (smem > 225 * 1024 || dim % BC) return -1;
    const int emax = (max_alen - 1) >> 1;                          // block-diagonals -1 .. emax
    dim3 grid((emax + 2 + ne - 1) / ne, nj);
    if (mode == SVX_COST_EXACT) {
        auto kern = k_banded_costs_p2<K, BC, DV, true, CW>;
        SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(jobs_d, dim, nstages, nb, ne, grouped, 1.0f, -0.0f);
    } else {
        auto kern = k_banded_costs_p2<K, BC, DV, false, CW>;
        SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(jobs_d, dim, nstages, nb, ne, grouped, 1.0f, -0.0f);
    }
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

}  // namespace

int svx_launch_costs_p2(int K, const SvxBandJob *jobs_d, int nj, int max_alen, int band, int dim, int mode, int bc,
                        int nstages, int nprod, int cons_warps, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (band & 1) return -1;
    if (nstages < 2) nstages = 2;
    if (nstages > 8) nstages = 8;
#define P2(KK, BCC, DVV, CWW) return launch_p2_t<KK, BCC, DVV, CWW>(jobs_d, nj, max_alen, band, dim, mode, nstages, nprod, st)
#define P2K(KK)                                                    \
    if (cons_warps == 12) { if (bc == 32) P2(KK, 32, 4, 12); else P2(KK, 16, 4, 12); } \
    else { if (bc == 32) P2(KK, 32, 4, 8); else P2(KK, 16, 4, 8); }
    switch (K) {
        case 1: P2K(1)
        case 2: P2K(2)
        case 3: P2K(3)
        case 4: P2K(4)
        case 5: if (bc == 32) P2(5, 32, 4, 8); else P2(5, 16, 4, 8);
        case 6: P2(6, 16, 2, 8);
        case 7: P2(7, 16, 2, 8);
        default: return -1;
    }
#undef P2K
#undef P2
}
