// banded.cu — fine levels (sm_100a): multi-type cosine costs inside the search band, the banded
// anti-diagonal DP, its traceback (alignment records + scores) and the next level's search path.
//   svx_banded_costs (dp_core.pyx:165-267 make_sparse_costs)
//   svx_banded_dp    (dp_core.pyx:269-404 sparse_dp; dp_utils.py:89-143 sparse_traceback +
//                     process_scores; dp_utils.py:177-275 path glue)
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <type_traits>
#include <utility>
#include "svx_common.cuh"
#include "svx_dp.h"
#include "svx_banded_p2.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Banded costs.
//
// One CTA owns TA consecutive anti-diagonals of one job; thread = one band cell (a, b), which
// needs, for every alignment type (xo,yo), the dot product of overlap row xo-1 ending at segment
// xx with overlap row yo-1 ending at yy.  Because the search path advances x or y by one per
// anti-diagonal, the TA*B cells of a tile touch only NX + NY = TA + 2B - 1 distinct segment
// positions: those rows (x CX/CY overlaps) are staged through shared memory in DC-float slices
// with coalesced 16-byte loads, and every thread keeps all its type accumulators in registers,
// adding products in increasing d (the reference order; EXACT = separate multiply and add).
//
//   TRI  = true : the standard type set of make_alignment_types(a) (vecalign.py:154-162), K = a-1
//                 overlaps per side, accumulators (i,j) with i + j <= K-1, one pass.
//   TRI  = false: any type list; passes over CX x CY blocks of (x overlap, y overlap) with a
//                 lookup table (kx,ky) -> type index.
// ---------------------------------------------------------------------------------------------
constexpr int kBC = 32;         // floats of the embedding dimension per staged slice (128 B per row)
constexpr int kBS = kBC + 4;    // padded row stride (36 floats = 9 x 16 B: conflict-free LDS.128)

template <int CX, int CY, bool TRI, bool EXACT>
__global__ void __launch_bounds__(384, (CX <= 5 ? 2 : 1))
k_banded_costs(const SvxBandJob *jobs, int dim, int ta, int lb)
{
    extern __shared__ __align__(16) float tile[];  // two slice buffers of nrows_max * kBS floats
    __shared__ int16_t tmap[TRI ? 1 : 64 * 64];   // (kx,ky) -> type index, generic path only
    const SvxBandJob &job = jobs[blockIdx.y];
    const int a0 = blockIdx.x * ta;
    if (a0 >= job.a_len) return;
    const int na = min(ta, job.a_len - a0);
    const int B = job.band, w = job.width_over2, T = job.ntypes;
    const int s0 = job.s0, s1 = job.s1;
    const int boff_first = job.ypath[a0] - w;
    const int boff_last = job.ypath[a0 + na - 1] - w;
    const int ylo = boff_first, yhi = boff_last + B - 1;
    const int xlo = a0 - boff_first - B + 1, xhi = (a0 + na - 1) - boff_last;
    const int NX = xhi - xlo + 1, NY = yhi - ylo + 1;
    const int buf_floats = (ta + 2 * B - 1) * (CX > CY ? CX : CY) * kBS;

    const int tid = threadIdx.x;
    // lb = band width rounded up to a multiple of 8 lanes: a quarter-warp (one LDS.128 phase) then
    // never straddles two anti-diagonals, so its 8 lanes read 8 consecutive rows = 8 bank groups
    const bool has_cell = tid / lb < na && tid % lb < B;
    const int la = has_cell ? tid / lb : 0, b = has_cell ? tid % lb : 0;
    const int a = a0 + la;
    const int yy = job.ypath[a] - w + b;
    const int xx = a - yy;
    const bool inside = has_cell && xx >= 0 && xx < s0 && yy >= 0 && yy < s1;
    const int xr = xx - xlo, yr = yy - ylo;

    int kxmax = 0, kymax = 0;     // overlaps actually referenced by the type list
    if (!TRI) {
        for (int i = tid; i < 64 * 64; i += blockDim.x) tmap[i] = -1;
        __syncthreads();
        for (int t = tid; t < T; t += blockDim.x) tmap[(job.xo[t] - 1) * 64 + (job.yo[t] - 1)] = (int16_t)t;
        for (int t = 0; t < T; ++t) { kxmax = max(kxmax, (int)job.xo[t]); kymax = max(kymax, (int)job.yo[t]); }
        __syncthreads();
    } else {
        kxmax = CX; kymax = CY;
    }

    const int slices = dim / kBC;
    for (int kx0 = 0; kx0 < kxmax; kx0 += CX) {
        for (int ky0 = 0; ky0 < kymax; ky0 += CY) {
            if (TRI && (kx0 | ky0)) break;
            float acc[CX][CY];
#pragma unroll
            for (int i = 0; i < CX; ++i)
#pragma unroll
                for (int j = 0; j < CY; ++j) acc[i][j] = 0.0f;

            const int xrows = CX * NX, nrows = xrows + CY * NY;
            // Every slice moves the same (row, 16-byte piece) items per thread: resolve their source rows
            // once (the divisions by NX / NY are not cheap) and keep the pointers in registers.
            constexpr int kMaxItems = 8;          // nrows * 8 pieces / blockDim <= 8 (checked by the launcher)
            unsigned srco[kMaxItems];             // float offset from job.v0 / job.v1
            unsigned from_y = 0, valid = 0;       // bit it: item reads side 1 / item has a source row
            const int nitems = nrows * (kBC / 4);
#pragma unroll
            for (int it = 0; it < kMaxItems; ++it) {
                const int f = tid + it * (int)blockDim.x;
                srco[it] = 0;
                if (f < nitems) {
                    const int row = f / (kBC / 4), c4 = f % (kBC / 4);
                    if (row < xrows) {
                        const int k = kx0 + row / NX, seg = xlo + row % NX;
                        if (k < job.k0 && seg >= 0 && seg < s0) { srco[it] = (unsigned)(((size_t)k * s0 + seg) * dim + 4 * c4); valid |= 1u << it; }
                    } else {
                        const int r2 = row - xrows;
                        const int k = ky0 + r2 / NY, seg = ylo + r2 % NY;
                        if (k < job.k1 && seg >= 0 && seg < s1) { srco[it] = (unsigned)(((size_t)k * s1 + seg) * dim + 4 * c4); valid |= 1u << it; }
                        from_y |= 1u << it;
                    }
                }
            }
            const unsigned tile_u32 = (unsigned)__cvta_generic_to_shared(tile);
            const float *gv0 = job.v0, *gv1 = job.v1;
            // slice `sl` -> buffer sl & 1, 16 bytes per cp.async, zero-filled (src-size 0) for rows outside
            // the documents / overlaps
            const bool many_items = nitems > kMaxItems * (int)blockDim.x;     // very wide bands: sources resolved per copy
            auto issue = [&](int sl) {
                const unsigned buf = tile_u32 + (unsigned)((sl & 1) * buf_floats * (int)sizeof(float));
                if (many_items) {
                    for (int f = tid; f < nitems; f += (int)blockDim.x) {
                        const int row = f / (kBC / 4), c4 = f % (kBC / 4);
                        const float *gp = gv0;
                        int nbytes = 0;
                        if (row < xrows) {
                            const int k = kx0 + row / NX, seg = xlo + row % NX;
                            if (k < job.k0 && seg >= 0 && seg < s0) { gp = gv0 + ((size_t)k * s0 + seg) * dim + 4 * c4 + sl * kBC; nbytes = 16; }
                        } else {
                            const int r2 = row - xrows;
                            const int k = ky0 + r2 / NY, seg = ylo + r2 % NY;
                            if (k < job.k1 && seg >= 0 && seg < s1) { gp = gv1 + ((size_t)k * s1 + seg) * dim + 4 * c4 + sl * kBC; nbytes = 16; }
                        }
                        const unsigned dst = buf + (unsigned)(((f >> 3) * kBS + 4 * (f & 7)) * (int)sizeof(float));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gp), "r"(nbytes));
                    }
                    asm volatile("cp.async.commit_group;\n" ::);
                    return;
                }
#pragma unroll
                for (int it = 0; it < kMaxItems; ++it) {
                    const int f = tid + it * (int)blockDim.x;
                    if (f < nitems) {
                        const float *gp = ((from_y >> it) & 1 ? gv1 : gv0) + srco[it] + sl * kBC;
                        const int nbytes = (valid >> it) & 1 ? 16 : 0;
                        const unsigned dst = buf + (unsigned)(((f >> 3) * kBS + 4 * (f & 7)) * (int)sizeof(float));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gp), "r"(nbytes));
                    }
                }
                asm volatile("cp.async.commit_group;\n" ::);
            };
            __syncthreads();              // previous (kx0, ky0) pass is done with both buffers
            issue(0);
            for (int sl = 0; sl < slices; ++sl) {
                if (sl + 1 < slices) {
                    issue(sl + 1);        // lands while slice sl is being consumed
                    asm volatile("cp.async.wait_group 1;\n" ::);
                } else {
                    asm volatile("cp.async.wait_group 0;\n" ::);
                }
                __syncthreads();
                const float *buf = tile + (sl & 1) * buf_floats;
                if (inside) {
#pragma unroll 2
                    for (int d = 0; d < kBC; d += 4) {
                        float4 xv[CX], yv[CY];
#pragma unroll
                        for (int i = 0; i < CX; ++i)
                            xv[i] = *reinterpret_cast<const float4 *>(buf + (size_t)(i * NX + xr) * kBS + d);
#pragma unroll
                        for (int j = 0; j < CY; ++j)
                            yv[j] = *reinterpret_cast<const float4 *>(buf + (size_t)(xrows + j * NY + yr) * kBS + d);
#pragma unroll
                        for (int i = 0; i < CX; ++i)
#pragma unroll
                            for (int j = 0; j < CY; ++j) {
                                if (TRI && i + j > CX - 1) continue;
                                if (EXACT) {
                                    acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(xv[i].x, yv[j].x));
                                    acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(xv[i].y, yv[j].y));
                                    acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(xv[i].z, yv[j].z));
                                    acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(xv[i].w, yv[j].w));
                                } else {
                                    acc[i][j] = fmaf(xv[i].x, yv[j].x, acc[i][j]);
                                    acc[i][j] = fmaf(xv[i].y, yv[j].y, acc[i][j]);
                                    acc[i][j] = fmaf(xv[i].z, yv[j].z, acc[i][j]);
                                    acc[i][j] = fmaf(xv[i].w, yv[j].w, acc[i][j]);
                                }
                            }
                    }
                }
                __syncthreads();          // buffer sl & 1 is free for slice sl + 2
            }
            if (has_cell) {
                float *out = job.costs + (size_t)a * T * B + b;
                int tri_t = 0;
#pragma unroll
                for (int i = 0; i < CX; ++i)
#pragma unroll
                    for (int j = 0; j < CY; ++j) {
                        int t;
                        if (TRI) {
                            if (i + j > CX - 1) continue;
                            t = tri_t++;            // x outer, y inner: vecalign.py:154-162 order
                        } else {
                            const int kx = kx0 + i, ky = ky0 + j;
                            t = (kx < 64 && ky < 64) ? tmap[kx * 64 + ky] : -1;
                            if (t < 0) continue;
                        }
                        float c = INFINITY;
                        if (inside) {
                            const int kx = kx0 + i, ky = ky0 + j;
                            c = svx_band_cost(acc[i][j], kx + 1, ky + 1, job.n0[(size_t)kx * s0 + xx],
                                              job.n1[(size_t)ky * s1 + yy]);
                        }
                        out[(size_t)t * B] = c;
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Banded costs, register-blocked (standard type set, K <= 4).
//
// In the thread-per-cell kernel above every 4 multiply-adds of a type need 8 LDS.128 per K(K+1)/2
// types: the shared-memory pipe, not the FP32 pipe, is the limit.  Here a thread owns a 2 x 2 block of
// POSITIONS (xx in {2X, 2X+1}, yy in {2Y, 2Y+1}): the four cells lie on anti-diagonals 2e, 2e+1, 2e+1,
// 2e+2 (e = X + Y), all inside or next to the band, and share their operand rows, so 4K LDS.128 feed
// 4 x T accumulators - twice the arithmetic per shared-memory byte.  Blocks are independent of the
// path's direction, so there is no divergence; a block-diagonal e needs at most B/2 + 2 blocks to
// cover the band rows of its three diagonals, and every band cell belongs to exactly one block.
// Rows are kept in shared memory as [even positions | odd positions], so consecutive threads
// (Y+1: two positions further) read consecutive rows: conflict-free LDS.128.
// Tiles span 30 anti-diagonals (the 2B-1 halo rows are amortised over twice as many diagonals).
// ---------------------------------------------------------------------------------------------
constexpr int kBlkTA = 30;      // 16 block-diagonals per tile (kBlkTA / 2 + 1): a quarter-warp = 8 consecutive block-diagonals

template <int K, bool EXACT>
__global__ void __launch_bounds__(160, 3)
k_banded_costs_blk(const SvxBandJob *jobs, int dim)
{
    constexpr int T = K * (K + 1) / 2;
    extern __shared__ __align__(16) float tile[];      // two slice buffers, then the row source table
    const SvxBandJob &job = jobs[blockIdx.y];
    const int A = job.a_len;
    const int a0 = blockIdx.x * kBlkTA;
    if (a0 >= A) return;
    const int d_first = a0, d_last = min(a0 + kBlkTA, A) - 1;
    const int B = job.band, w = job.width_over2;
    const int s0 = job.s0, s1 = job.s1;
    const int32_t *ypath = job.ypath;
    const int tid = threadIdx.x;
    const int nb = B / 2 + 2, ne = kBlkTA / 2 + 1;

    // positions touched by the tile, rounded out to even / odd ends
    const int bf = ypath[d_first] - w, bl = ypath[d_last] - w;
    const int ylo = ((bf) >> 1) * 2, yhi = ((bl + B - 1) >> 1) * 2 + 1;
    const int xlo = ((d_first - (bf + B - 1)) >> 1) * 2, xhi = ((d_last - bl) >> 1) * 2 + 1;
    const int NX = xhi - xlo + 1, NY = yhi - ylo + 1, HX = NX >> 1, HY = NY >> 1;
    const int nrows = K * (NX + NY);
    const int rows_cap = K * (kBlkTA + 2 * B + 4);
    const int buf_floats = rows_cap * kBS;
    int *srcoff = reinterpret_cast<int *>(tile + 2 * buf_floats);        // float offset from v0 / v1, -1 = zero row

    // row slot -> source: x rows first (overlap-major, [even | odd] positions), then y rows
    for (int r = tid; r < nrows; r += blockDim.x) {
        int off = -1;
        if (r < K * NX) {
            const int k = r / NX, p = r % NX;
            const int seg = xlo + (p < HX ? 2 * p : 2 * (p - HX) + 1);
            if (k < job.k0 && seg >= 0 && seg < s0) off = (int)(((size_t)k * s0 + seg) * dim);
        } else {
            const int r2 = r - K * NX;
            const int k = r2 / NY, p = r2 % NY;
            const int seg = ylo + (p < HY ? 2 * p : 2 * (p - HY) + 1);
            if (k < job.k1 && seg >= 0 && seg < s1) off = (int)(((size_t)k * s1 + seg) * dim);
        }
        srcoff[r] = off;
    }

    // this thread's block.  Consecutive lanes = consecutive block-diagonals e at the same band slot yi:
    // along e both X and Y advance by 0 or 1, so the 8 lanes of an LDS.128 phase read rows that are
    // equal (broadcast) or consecutive - no bank conflicts.
    static_assert(kBlkTA / 2 + 1 == 16, "thread mapping assumes 16 block-diagonals per tile");
    const int ei = tid & 15, yi = tid >> 4;
    const int e = a0 / 2 - 1 + ei;
    const int dlo = max(2 * e, d_first), dhi = min(2 * e + 2, d_last);
    bool active = yi < nb && dlo <= dhi;
    int xx0 = 0, yy0 = 0;
    unsigned inband = 0;               // bit (2j + i): cell (xx0 + i, yy0 + j) is a band cell of this tile
    int bslot[4] = {0, 0, 0, 0};
    if (active) {
        const int Y = ((ypath[dlo] - w) >> 1) + yi;
        yy0 = 2 * Y; xx0 = 2 * (e - Y);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int d = 2 * e + i + j;
                if (d >= d_first && d <= d_last) {
                    const int b = yy0 + j - (ypath[d] - w);
                    if (b >= 0 && b < B) { inband |= 1u << (2 * j + i); bslot[2 * j + i] = b; }
                }
            }
        active = inband != 0;
    }
    const int xs_even = (xx0 - xlo) >> 1, ys_even = (yy0 - ylo) >> 1;      // slot of the even position in its half

    float acc[4][T];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int t = 0; t < T; ++t) acc[c][t] = 0.0f;

    const unsigned tile_u32 = (unsigned)__cvta_generic_to_shared(tile);
    const float *gv0 = job.v0, *gv1 = job.v1;
    const int xrows = K * NX;
    auto issue = [&](int sl) {
        const unsigned buf = tile_u32 + (unsigned)((sl & 1) * buf_floats * (int)sizeof(float));
        for (int f = tid; f < nrows * (kBC / 4); f += blockDim.x) {
            const int row = f >> 3, c4 = f & 7;
            const int off = srcoff[row];
            const float *gp = (row < xrows ? gv0 : gv1) + (off < 0 ? 0 : off) + sl * kBC + 4 * c4;
            const int nbytes = off < 0 ? 0 : 16;
            const unsigned dst = buf + (unsigned)((row * kBS + 4 * c4) * (int)sizeof(float));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gp), "r"(nbytes));
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    const int slices = dim / kBC;
    __syncthreads();                      // srcoff is complete
    issue(0);
    for (int sl = 0; sl < slices; ++sl) {
        if (sl + 1 < slices) {
            issue(sl + 1);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        if (active) {
            const float *buf = tile + (sl & 1) * buf_floats;
            const float *px = buf + (size_t)xs_even * kBS;                       // even x position, overlap 0
            const float *py = buf + (size_t)(xrows + ys_even) * kBS;
#pragma unroll 2
            for (int d = 0; d < kBC; d += 4) {
                // y operands of every overlap stay live; x operands are fetched overlap by overlap
                float4 ye[K], yo[K];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    ye[j] = *reinterpret_cast<const float4 *>(py + (size_t)(j * NY) * kBS + d);
                    yo[j] = *reinterpret_cast<const float4 *>(py + (size_t)(j * NY + HY) * kBS + d);
                }
                int t = 0;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const float4 xe = *reinterpret_cast<const float4 *>(px + (size_t)(i * NX) * kBS + d);
                    const float4 xo = *reinterpret_cast<const float4 *>(px + (size_t)(i * NX + HX) * kBS + d);
#pragma unroll
                    for (int j = 0; i + j <= K - 1; ++j, ++t) {
#define SVX_MAC4(ACC, XV, YV)                                                                       \
    if (EXACT) {                                                                                    \
        ACC = __fadd_rn(ACC, __fmul_rn(XV.x, YV.x)); ACC = __fadd_rn(ACC, __fmul_rn(XV.y, YV.y));   \
        ACC = __fadd_rn(ACC, __fmul_rn(XV.z, YV.z)); ACC = __fadd_rn(ACC, __fmul_rn(XV.w, YV.w));   \
    } else {                                                                                        \
        ACC = fmaf(XV.x, YV.x, ACC); ACC = fmaf(XV.y, YV.y, ACC);                                   \
        ACC = fmaf(XV.z, YV.z, ACC); ACC = fmaf(XV.w, YV.w, ACC);                                   \
    }
                        SVX_MAC4(acc[0][t], xe, ye[j])       // (xx0,     yy0)
                        SVX_MAC4(acc[1][t], xo, ye[j])       // (xx0 + 1, yy0)
                        SVX_MAC4(acc[2][t], xe, yo[j])       // (xx0,     yy0 + 1)
                        SVX_MAC4(acc[3][t], xo, yo[j])       // (xx0 + 1, yy0 + 1)
#undef SVX_MAC4
                    }
                }
            }
        }
        __syncthreads();
    }
    if (!active) return;
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
                float cst = INFINITY;
                if (inside)
                    cst = svx_band_cost(acc[c][t], kx + 1, ky + 1, job.n0[(size_t)kx * s0 + xx], job.n1[(size_t)ky * s1 + yy]);
                out[(size_t)t * B] = cst;
            }
    }
}

template <int K>
int launch_costs_blk(const SvxBandJob *jobs_d, int nj, int max_alen, int band, int dim, int mode, cudaStream_t st)
{
    const int nb = band / 2 + 2, ne = kBlkTA / 2 + 1;
    const int threads = ((nb * ne + 31) / 32) * 32;
    if (threads > 160) return -1;
    const int rows_cap = K * (kBlkTA + 2 * band + 4);
    const size_t smem = (size_t)2 * rows_cap * kBS * sizeof(float) + (size_t)rows_cap * sizeof(int);
    if (smem > 200 * 1024) return -1;
    dim3 grid((max_alen + kBlkTA - 1) / kBlkTA, nj);
    if (mode == SVX_COST_EXACT) {
        auto kern = k_banded_costs_blk<K, true>;
        if (smem > 48 * 1024) SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(jobs_d, dim);
    } else {
        auto kern = k_banded_costs_blk<K, false>;
        if (smem > 48 * 1024) SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(jobs_d, dim);
    }
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

// ---------------------------------------------------------------------------------------------
// Banded DP: one CTA of 4 warps per job.  Warp 0 runs the anti-diagonal recurrence, lane = band
// slot, with the last R diagonals of fp64 cumulative costs in a shared-memory ring; all warps
// stage the next chunk of cost diagonals (contiguous T*B floats each) and band offsets into a
// double buffer while warp 0 computes.  Backpointers are uint8 type indices in HBM; the walk and
// the next search path follow in the same kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxChunk = 32;   // anti-diagonals staged per buffer (fewer when T*B is large)

struct BandSmem {
    double *ring;      // R * B
    float *cost[2];    // kChunk * T * B each
    int *boff[2];      // kChunk + R each: b_offset_out of diagonals [chunk_start - R, chunk_end)
};

struct RecWriter {
    SvxAlignRec *recs;
    int cap;
    int count;
    int lane;
    int overflow;
};

template <class Builder>
struct OnAlignDev {
    RecWriter *rw;
    Builder *pb;     // may be null
    __host__ __device__ void operator()(int x_end, int y_end, int nx, int ny, double score)
    {
        if (rw->recs) {
            if (rw->count < rw->cap) {
                if (rw->lane == 0) {
                    SvxAlignRec r; r.x_end = x_end; r.y_end = y_end; r.nx = nx; r.ny = ny; r.score = score;
                    rw->recs[rw->cap - 1 - rw->count] = r;
                }
            } else rw->overflow = 1;
        }
        rw->count++;
        if (pb) pb->step(x_end, y_end, nx, ny);
    }
};

struct WarpEmitB {
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

__global__ void __launch_bounds__(128) k_banded_dp(const SvxBandJob *jobs, int R, int kChunk)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int8_t sxo[SVX_MAX_TYPES + 2], syo[SVX_MAX_TYPES + 2];
    const SvxBandJob &job = jobs[blockIdx.x];
    const int B = job.band, T = job.ntypes, A = job.a_len, w = job.width_over2;
    const int s0 = job.s0, s1 = job.s1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nodes_a = A + 2;
    const int tb = T * B;

    BandSmem sm;
    sm.ring = reinterpret_cast<double *>(smem_raw);
    sm.cost[0] = reinterpret_cast<float *>(sm.ring + (size_t)R * B);
    sm.cost[1] = sm.cost[0] + (size_t)kChunk * tb;
    sm.boff[0] = reinterpret_cast<int *>(sm.cost[1] + (size_t)kChunk * tb);
    sm.boff[1] = sm.boff[0] + (kChunk + R);

    for (int t = tid; t < T; t += blockDim.x) { sxo[t] = job.xo[t]; syo[t] = job.yo[t]; }
    const double pen = *job.del_penalty;

    // stage chunk c (node diagonals [c*kChunk, (c+1)*kChunk)) into buffer c & 1
    auto stage = [&](int c, int first_thread, int nthreads) {
        const int start = c * kChunk;
        float *cb = sm.cost[c & 1];
        int *bb = sm.boff[c & 1];
        // cost diagonal of node diagonal aa is aa - 2
        const int lo = start - 2, hi = min(start + kChunk, nodes_a) - 2;    // [lo, hi)
        const int clo = max(lo, 0), chi = min(hi, A);
        if (chi > clo) {
            const float *src = job.costs + (size_t)clo * tb;
            float *dst = cb + (size_t)(clo - lo) * tb;
            const int n = (chi - clo) * tb;
            for (int i = first_thread; i < n; i += nthreads) dst[i] = __ldg(src + i);
        }
        for (int i = first_thread; i < kChunk + R; i += nthreads) {
            const int aa = start - R + i;
            bb[i] = (aa >= 0 && aa < nodes_a) ? svx_boff_out(job.ypath, aa, w) : 0;
        }
    };

    const int nchunks = (nodes_a + kChunk - 1) / kChunk;
#ifdef SVX_DP_TIMING
    long long tm0 = clock64(), tm1 = 0, tm2 = 0;
#endif
    stage(0, tid, blockDim.x);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        if (warp != 0) {
            if (c + 1 < nchunks) stage(c + 1, tid - 32, blockDim.x - 32);
        } else {
            const float *cb = sm.cost[c & 1];
            const int *bo = sm.boff[c & 1];
            const int start = c * kChunk;
            const int end = min(start + kChunk, nodes_a);
            for (int aa = start; aa < end; ++aa) {
                // bands wider than a warp (large search_buffer_size): a lane takes slots lane, lane+32, ...;
                // the nodes of one diagonal only read older diagonals, so their order does not matter
                for (int bb = lane; bb < B; bb += 32) {
                    int bp;
                    auto boff = [&](int q) { return bo[q - start + R]; };            // q in (aa-R, aa]
                    auto csum_at = [&](int aq, int bq) { return sm.ring[(size_t)(aq & (R - 1)) * B + bq]; };
                    auto cost_at = [&](int t) { return cb[(size_t)(aa - start) * tb + (size_t)t * B + bb]; };
                    const double v = svx_band_node(aa, bb, s0, s1, A, B, T, sxo, syo, pen, boff, csum_at, cost_at, &bp);
                    // ring slot aa & (R-1) held diagonal aa - R, which no type can reach any more
                    sm.ring[(size_t)(aa & (R - 1)) * B + bb] = v;
                    job.bp[(size_t)aa * B + bb] = (uint8_t)bp;
                    job.csum[(size_t)aa * B + bb] = v;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    // traceback + next search path (warp 0; every lane walks, lane 0 writes the records)
    if (warp == 0) {
        RecWriter rw{job.recs, job.rec_cap, 0, lane, 0};
        int st;
        if (job.next_ypath) {
            WarpEmitB emit{job.next_ypath, job.next_len, lane};
            SvxPathBuilder<WarpEmitB> pb(emit);
            pb.begin(s0, s1, job.t0, job.t1, 1);
            OnAlignDev<SvxPathBuilder<WarpEmitB>> on{&rw, &pb};
            st = svx_band_walk(job.bp, job.csum, job.ypath, w, s0, s1, A, B, T, sxo, syo, on);
            if (st == SVX_ST_OK) pb.finish();
            if (lane == 0 && job.next_len > 0) job.next_ypath[0] = 0;
        } else {
            OnAlignDev<SvxPathBuilder<WarpEmitB>> on{&rw, nullptr};
            st = svx_band_walk(job.bp, job.csum, job.ypath, w, s0, s1, A, B, T, sxo, syo, on);
        }
        if (rw.overflow) st |= SVX_ST_OVERFLOW;
        if (lane == 0) {
            *job.status_d = st;
            if (job.nrecs) *job.nrecs = rw.count;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Banded DP, standard type set (make_alignment_types(K+1), vecalign.py:154-162; K = 1 for the
// coarser levels).  One CTA of 4 warps per job.
//
//   phase 1  warp 0 runs the recurrence, lane = band slot.  The fp64 cumulative costs of the last
//            K+1 diagonals live in a shared-memory ring whose rows are compile-time constants (the
//            diagonal loop is unrolled K+1 times); the predecessor of a candidate (dx,dy) on diagonal
//            aa-(dx+dy) sits at lane  b + (boff(aa) - boff(aa-dx-dy)) - dy,  a shift that is uniform over
//            the warp and precomputed per diagonal, so each candidate is one add, two LDS.64, one DADD
//            and one compare-select, in the reference's priority order (types, then (0,1), then (1,0);
//            strict '<' keeps the first minimum).  The two deletion candidates - the serial chain -
//            stay in registers (64-bit shuffles).  Warps 1-3 prepare everything that does not depend
//            on the chain, one chunk ahead: cost diagonals widened to fp64, band-offset differences,
//            forced values of boundary / outside nodes; and they flush the finished chunk's uint8
//            backpointers and fp64 csum to HBM.  (Measured: 705 -> 530 cycles per diagonal at K = 4;
//            the bare DADD -> SHFL -> compare chain is 68.)
//   phase 2  thread 0 walks the backpointers from (s0,s1) through a shared-memory window of
//            backpointers + band offsets (reloaded, coalesced, when the walk leaves it) and
//            writes the integer fields of the alignment records.
//   phase 3  all threads: scores = clipped csum differences (process_scores), and the search
//            path of the next finer level, one alignment (+ the deletion run that follows it)
//            per thread; long slanted segments are expanded by the whole CTA.
// ---------------------------------------------------------------------------------------------
template <class F, int... Us>
__device__ __forceinline__ void svx_static_for(F &f, std::integer_sequence<int, Us...>)
{
    (f(std::integral_constant<int, Us>{}), ...);
}

constexpr int kBigSeg = 96;      // segments longer than this are expanded cooperatively
constexpr int kSegQueue = 48;

struct SegQueue {
    long long xs[kSegQueue], ys[kSegQueue], xw[kSegQueue], yw[kSegQueue];
    int n;
};

__device__ __forceinline__ void emit_points(int32_t *ypath, int path_len, long long xs, long long ys, long long xw,
                                            long long yw, long long first, long long stride)
{
    const long long nn = xw + yw;
    for (long long i = first; i <= nn; i += stride) {
        int x, y;
        svx_slant_point(xs, ys, xw, yw, i, &x, &y);
        if (x + y < path_len) ypath[x + y] = y;
    }
}

__device__ __forceinline__ void emit_segment(SegQueue *q, int32_t *ypath, int path_len, long long xs, long long ys,
                                             long long xw, long long yw)
{
    if (xw + yw <= 0) return;
    if (xw + yw > kBigSeg) {
        const int slot = atomicAdd(&q->n, 1);
        if (slot < kSegQueue) {
            q->xs[slot] = xs; q->ys[slot] = ys; q->xw[slot] = xw; q->yw[slot] = yw;
            return;
        }
    }
    emit_points(ypath, path_len, xs, ys, xw, yw, 1, 1);
}

// backpointer code of the standard type set -> dx | dy << 4
__host__ __device__ constexpr int svx_dec_code(int K, int code)
{
    const int T = K * (K + 1) / 2;
    if (code == T) return 0 | (1 << 4);
    if (code == T + 1) return 1 | (0 << 4);
    int xq = 1, rem = code;
    while (rem >= K + 1 - xq) { rem -= K + 1 - xq; ++xq; }
    return xq | ((rem + 1) << 4);
}
__host__ __device__ constexpr unsigned long long svx_dec_packed(int K, int first)
{
    const int T = K * (K + 1) / 2;
    unsigned long long v = 0;
    for (int code = first; code < T + 2 && code < first + 8; ++code)
        v |= (unsigned long long)svx_dec_code(K, code) << (8 * (code - first));
    return v;
}

constexpr int kDpThreads = 256;      // warp 0 = recurrence, 7 staging warps (their work is global-load latency)

template <int K>
__global__ void __launch_bounds__(kDpThreads) k_banded_dp_tri(const SvxBandJob *jobs, int kChunk, int win_diags)
{
    constexpr int T = K * (K + 1) / 2;
    constexpr int NH = K + 1;                  // deepest diagonal a candidate reaches back to
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wst[5];                     // walk state: x, y, count, status, done
    __shared__ SegQueue segq;
    __shared__ uint8_t tdec[SVX_MAX_TYPES + 2];
    const SvxBandJob &job = jobs[blockIdx.x];
    const int B = job.band, A = job.a_len, w = job.width_over2;
    const int s0 = job.s0, s1 = job.s1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nodes_a = A + 2;
    const int tb = T * B;
    const int cstride = kChunk * tb + 32;      // +32: lanes >= B read past their diagonal's block

    // per chunk (double-buffered): the cost diagonals widened to fp64, and what the recurrence needs that does
    // NOT depend on the recurrence - all produced by the staging warps, off the serial chain:
    //   oc / ov  per (diagonal, lane): forced backpointer code + value (boundary nodes, nodes outside the
    //            lattice, lanes >= B), kNoOvr where the recurrence decides
    //   dl       per diagonal: [0] = boff(aa) - boff(aa-1); [s] = 8 * (boff(aa) - boff(aa-s)), s = 2..NH:
    //            the lane shift of a predecessor on diagonal aa - s, as a byte offset into the csum ring
    constexpr int kNoOvr = 0xFE;
    constexpr int DLS = NH + 1;
    constexpr int RS = 32 + K + NH + ((32 + K + NH) & 1);      // ring row: K pad | 32 lanes | NH pad, in doubles
    double *cbuf0 = reinterpret_cast<double *>(smem_raw);
    double *cbuf1 = cbuf0 + cstride;
    double *ov0 = cbuf1 + cstride;
    double *ov1 = ov0 + kChunk * 32;
    double *ring = ov1 + kChunk * 32;                          // NH rows of RS doubles
    int *dl0 = reinterpret_cast<int *>(ring + NH * RS);
    int *dl1 = dl0 + kChunk * DLS;
    uint8_t *oc0 = reinterpret_cast<uint8_t *>(dl1 + kChunk * DLS);
    uint8_t *oc1 = oc0 + kChunk * 32;
    // results of a chunk (csum, backpointer code per diagonal and lane): written by the recurrence with plain
    // shared-memory stores, flushed to HBM (coalesced) by the staging warps while the next chunk runs
    uint8_t *bpo0 = oc1 + kChunk * 32;
    uint8_t *bpo1 = bpo0 + kChunk * 32;
    double *cso0 = reinterpret_cast<double *>(bpo1 + kChunk * 32);      // 8-byte aligned: every block above is a multiple of 8 bytes
    double *cso1 = cso0 + kChunk * 32;
    int *bos = reinterpret_cast<int *>(cso1 + kChunk * 32);       // band offsets of diagonals start-NH .. start+kChunk-1
    const double pen = *job.del_penalty;
    const float *g_costs = job.costs;
    const int32_t *g_ypath = job.ypath;
    uint8_t *g_bp = job.bp;
    double *g_csum = job.csum;

    // chunk c = node diagonals [c*kChunk, (c+1)*kChunk); its cost diagonals (aa - 2) are one contiguous,
    // 16-byte aligned run in HBM (kChunk is even and so is the band width, so kChunk*tb and 2*tb floats are
    // multiples of 4): read as float4, stored as fp64 (the recurrence adds them to fp64 cumulative costs).
    // `sync` orders the two passes among the calling threads: band offsets of the chunk (+ NH diagonals of
    // history) go to shared memory first, so that every later value is computed from shared memory instead of
    // dependent global loads (the staging warps used to be as busy as the recurrence warp)
    auto stage = [&](int c, int first_thread, int nthreads, auto sync) {
        const int start = c * kChunk;
        double *cb = (c & 1) ? cbuf1 : cbuf0;
        double *ov = (c & 1) ? ov1 : ov0;
        int *dl = (c & 1) ? dl1 : dl0;
        uint8_t *oc = (c & 1) ? oc1 : oc0;
        for (int i = first_thread; i < kChunk + NH; i += nthreads) {
            const int aa = start - NH + i;                      // diagonals before the first one: offset 0
            bos[i] = (aa >= 0 && aa < nodes_a) ? svx_boff_out(g_ypath, aa, w) : 0;
        }
        const int lo = start - 2, hi = min(start + kChunk, nodes_a) - 2;   // cost diagonals [lo, hi)
        const int clo = max(lo, 0), chi = min(hi, A);
        if (chi > clo) {
            const float *src = g_costs + (size_t)clo * tb;
            double *dst = cb + (size_t)(clo - lo) * tb;
            const int n = (chi - clo) * tb;
            const int n4 = n >> 2;
            for (int i = first_thread; i < n4; i += nthreads) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
                reinterpret_cast<double2 *>(dst)[2 * i] = make_double2((double)v.x, (double)v.y);
                reinterpret_cast<double2 *>(dst)[2 * i + 1] = make_double2((double)v.z, (double)v.w);
            }
            for (int i = 4 * n4 + first_thread; i < n; i += nthreads) dst[i] = (double)__ldg(src + i);
        }
        sync();
        for (int idx = first_thread; idx < kChunk * 32; idx += nthreads) {
            const int i = idx >> 5, l = idx & 31, aa = start + i;
            int code = SVX_BP_NONE;
            double val = INFINITY;
            if (aa < nodes_a && l < B) {
                const int yy = l + bos[NH + i], xx = aa - yy;
                // every type ending here reads cost cell (xx-1, yy-1): it must exist (also for the deletions -
                // reference quirk, dp_core.pyx:382,390; its anti-diagonal aa - 2 < A always holds for aa <= A + 1)
                if ((unsigned)(xx - 1) < (unsigned)s0 && (unsigned)(yy - 1) < (unsigned)s1) code = kNoOvr;
                else if (xx == 0 && yy >= 0 && yy <= s1) { val = __dmul_rn(pen, (double)yy); code = T; }        // bp (0,1)
                else if (yy == 0 && xx >= 0 && xx <= s0) { val = __dmul_rn(pen, (double)xx); code = T + 1; }    // bp (1,0)
            }
            oc[idx] = (uint8_t)code;
            ov[idx] = val;
        }
        for (int idx = first_thread; idx < kChunk * DLS; idx += nthreads) {
            const int i = idx / DLS, sft = idx % DLS, aa = start + i;
            const int back = sft == 0 ? 1 : sft;
            int d = 0;
            if (aa < nodes_a && sft != 1) {
                d = bos[NH + i] - bos[NH + i - back];
                if (sft) d = min(max(d, -K), NH) * 8;       // inside the ring row's padding whatever the path does
            }
            dl[idx] = d;
        }
    };

    auto flush = [&](int c, int first_thread, int nthreads) {
        const int start = c * kChunk;
        const int nrows = min(kChunk, nodes_a - start);
        const double *cso = (c & 1) ? cso1 : cso0;
        const uint8_t *bpo = (c & 1) ? bpo1 : bpo0;
        double *gc = g_csum + (size_t)start * B;
        uint8_t *gb = g_bp + (size_t)start * B;
        const int l = first_thread & 31;
        if (l < B)
            for (int i = first_thread >> 5; i < nrows; i += nthreads >> 5) {      // a warp per row, a lane per slot
                gc[i * B + l] = cso[i * 32 + l];
                gb[i * B + l] = bpo[i * 32 + l];
            }
    };

    // ---- phase 1 ------------------------------------------------------------------------------
    // The fp64 cumulative costs of the last NH diagonals live in a shared-memory ring with COMPILE-TIME
    // rows: chunks hold whole groups of NH diagonals (kChunk % NH == 0), diagonal i of a chunk uses row
    // i % NH, and the group loop is fully unrolled.  A type candidate (x,y) is then one add (lane byte offset
    // + dl[x+y]) and one LDS.64 of the predecessor, one LDS.64 of the cost, DADD, compare-select; the rows
    // are padded with +inf on both sides and lanes >= B / lattice nodes outside the documents are forced to
    // +inf, so predecessors outside the node band never win the strict '<' - no range tests.  Candidates of
    // diagonal aa + 1 (they reach back >= 2 diagonals) are evaluated BEFORE the chain of diagonal aa:
    // csum(aa-1) -> + pen -> two 64-bit shuffles -> two compare-selects -> forced value -> csum(aa).
    // Two independent strict-'<' chains (first / second half of the type list) halve the dependent
    // compare-select latency and keep the first-minimum tie-break.
    for (int i = tid; i < NH * RS; i += blockDim.x) ring[i] = INFINITY;
    constexpr int SPLIT = T >= 6 ? (T + 1) / 2 : T;
    const char *ring_lane = reinterpret_cast<const char *>(ring + K + lane);
    struct Cand { double best; int code; };
    struct Shifts { int v[NH + 1]; };                            // dl row of one diagonal, in registers
    auto load_shifts = [&](const int *dl, int i) {
        Shifts sh;
#pragma unroll
        for (int q = 0; q <= NH; ++q) sh.v[q] = (q == 1) ? 0 : dl[i * DLS + q];
        return sh;
    };
    auto types = [&](auto slot_tag, const double *cb, const Shifts &sh, int i) -> Cand {
        constexpr int U = decltype(slot_tag)::value;          // ring row of diagonal i
        const double *crow = cb + (size_t)i * tb + lane;
        double b0 = INFINITY, b1 = INFINITY;
        int c0 = SVX_BP_NONE, c1 = SVX_BP_NONE;
        int t = 0;
#pragma unroll
        for (int x = 1; x <= K; ++x) {
#pragma unroll
            for (int y = 1; x + y <= K + 1; ++y, ++t) {
                const int sft = x + y;                                       // >= 2
                const int row = (U - sft + 2 * NH) % NH;
                const double pv = *reinterpret_cast<const double *>(ring_lane + sh.v[sft] + (row * RS - y) * 8);
                const double tot = __dadd_rn(pv, crow[t * B]);
                if (t < SPLIT) { if (tot < b0) { b0 = tot; c0 = t; } }
                else { if (tot < b1) { b1 = tot; c1 = t; } }
            }
        }
        if (SPLIT < T && b1 < b0) { b0 = b1; c0 = c1; }
        return Cand{b0, c0};
    };

    const int nchunks = (nodes_a + kChunk - 1) / kChunk;
#ifdef SVX_DP_TIMING
    long long tm0 = clock64(), tm1 = 0, tm2 = 0;
#endif
    stage(0, tid, blockDim.x, [] { __syncthreads(); });
    __syncthreads();
    auto stagers_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kDpThreads - 32) : "memory"); };
    double prev = INFINITY, prev2 = INFINITY;     // csum of the previous two diagonals, this lane
    for (int c = 0; c < nchunks; ++c) {
        if (warp != 0) {
            if (c >= 1) flush(c - 1, tid - 32, kDpThreads - 32);
            if (c + 1 < nchunks) stage(c + 1, tid - 32, kDpThreads - 32, stagers_sync);
        } else {
            const double *cb = (c & 1) ? cbuf1 : cbuf0;
            const double *ov = ((c & 1) ? ov1 : ov0) + lane;
            const uint8_t *oc = ((c & 1) ? oc1 : oc0) + lane;
            const int *dl = (c & 1) ? dl1 : dl0;
            const int start = c * kChunk;
            const int ndiag = min(kChunk, nodes_a - start);
            uint8_t *bp_out = ((c & 1) ? bpo1 : bpo0) + lane;
            double *cs_out = ((c & 1) ? cso1 : cso0) + lane;
            // software pipeline: everything of diagonal i + 1 that does not depend on csum(i) is loaded / evaluated
            // while the shuffles of diagonal i are in flight
            Shifts sh = load_shifts(dl, 0);
            Cand cur = types(std::integral_constant<int, 0>{}, cb, sh, 0), nxt = cur;
            double cost1 = cb[lane];                              // K == 1: cost of the chunk's first diagonal
            int fcode = oc[0];
            double fval = ov[0];
            for (int g = 0; g < ndiag; g += NH) {
                const bool group_follows = g + NH < kChunk;
                auto step = [&](auto u_tag) {
                    constexpr int u = decltype(u_tag)::value;
                    const int i = g + u;                       // diagonals past the end compute on padding; not stored
                    const bool more = u + 1 < NH || group_follows;          // diagonal i + 1 is in this chunk
                    const int d1 = sh.v[0];
                    if constexpr (K == 1) {
                        // coarse levels, one type (1,1): its predecessor csum(i-2) is still in a register, so the
                        // whole step runs on shuffles - no ring traffic, no __syncwarp
                        const double hp = __dadd_rn(prev, pen);
                        const double p11 = __shfl_sync(0xffffffffu, prev2, lane + (sh.v[2] >> 3) - 1);
                        const double ty = __shfl_sync(0xffffffffu, hp, lane + d1 - 1);
                        const double tx = __shfl_sync(0xffffffffu, hp, lane + d1);
                        const double cost = cost1;
                        int ncode = fcode;
                        double nval = fval;
                        if (more) {
                            sh = load_shifts(dl, i + 1);
                            ncode = oc[(i + 1) * 32];
                            nval = ov[(i + 1) * 32];
                            cost1 = cb[(size_t)(i + 1) * tb + lane];
                        }
                        double best = INFINITY;
                        int code = SVX_BP_NONE;
                        const double t11 = __dadd_rn(p11, cost);
                        if (t11 < best) { best = t11; code = 0; }
                        if (ty < best) { best = ty; code = T; }
                        if (tx < best) { best = tx; code = T + 1; }
                        if (fcode != kNoOvr) { best = fval; code = fcode; }
                        bp_out[i * 32] = (uint8_t)code;
                        cs_out[i * 32] = best;
                        prev2 = prev;
                        prev = best;
                        fcode = ncode; fval = nval;
                        return;
                    }
                    const double hp = __dadd_rn(prev, pen);
                    const double ty = __shfl_sync(0xffffffffu, hp, lane + d1 - 1);   // (0,1): consume y
                    const double tx = __shfl_sync(0xffffffffu, hp, lane + d1);       // (1,0): consume x
                    int ncode = fcode;
                    double nval = fval;
                    if (more) {
                        sh = load_shifts(dl, i + 1);
                        ncode = oc[(i + 1) * 32];
                        nval = ov[(i + 1) * 32];
                        nxt = types(std::integral_constant<int, (u + 1) % NH>{}, cb, sh, i + 1);
                    }
                    double best = cur.best;
                    int code = cur.code;
                    if (ty < best) { best = ty; code = T; }
                    if (tx < best) { best = tx; code = T + 1; }
                    if (fcode != kNoOvr) { best = fval; code = fcode; }
                    bp_out[i * 32] = (uint8_t)code;
                    cs_out[i * 32] = best;
                    ring[u * RS + K + lane] = best;
                    __syncwarp();
                    prev = best;
                    cur = nxt; fcode = ncode; fval = nval;
                };
                svx_static_for(step, std::make_integer_sequence<int, NH>{});
            }
        }
        __syncthreads();
    }
    flush(nchunks - 1, tid, blockDim.x);
    __syncthreads();

    // ---- phase 2: backpointer walk through a shared-memory window ---------------------------------
#ifdef SVX_DP_TIMING
    tm1 = clock64();
#endif
    int *wboff = reinterpret_cast<int *>(smem_raw);
    uint8_t *wbp = reinterpret_cast<uint8_t *>(wboff + ((win_diags + 8 + 3) & ~3));     // 16-byte aligned
    const int cap = job.rec_cap;
    SvxAlignRec *const recs_top = job.recs + (cap - 1);      // in a register: the walk's stores must not reload the descriptor
    // backpointer code -> dx | dy << 4 (types x outer, y inner: row x holds K+1-x types; then (0,1), (1,0))
    constexpr unsigned long long kDecPacked = svx_dec_packed(K, 0), kDecPackedHi = svx_dec_packed(K, 8);
    for (int code = tid; code < T + 2; code += blockDim.x) tdec[code] = (uint8_t)svx_dec_code(K, code);
    if (tid == 0) {
        wst[0] = s0; wst[1] = s1; wst[2] = 0;
        wst[3] = (s0 + s1 >= nodes_a) ? SVX_ST_LEFT_BAND : SVX_ST_OK;
        wst[4] = (wst[3] != SVX_ST_OK) || (s0 == 0 && s1 == 0);
        segq.n = 0;
    }
    __syncthreads();
    while (!wst[4]) {
        const int top = wst[0] + wst[1];
        // window [lo, hi): lo rounded down to a multiple of 8 diagonals so that its first backpointer is
        // 16-byte aligned (B is even) and the bulk of the window moves as uint4
        const int hi = top + 1, lo = max(0, hi - win_diags) & ~7;
        for (int i = tid; i < hi - lo; i += blockDim.x) wboff[i] = svx_boff_out(job.ypath, lo + i, w);
        const uint8_t *bsrc = job.bp + (size_t)lo * B;
        const int nbytes = (hi - lo) * B, n16 = nbytes >> 4;
        for (int i = tid; i < n16; i += blockDim.x)
            reinterpret_cast<uint4 *>(wbp)[i] = __ldcg(reinterpret_cast<const uint4 *>(bsrc) + i);
        for (int i = (n16 << 4) + tid; i < nbytes; i += blockDim.x) wbp[i] = bsrc[i];
        __syncthreads();
        if (tid == 0) {
            // One thread chases the backpointers.  Per step the dependent chain is: LDS backpointer code ->
            // (dx,dy) decode (a shift of a packed constant for T + 2 <= 16, else one LDS of a table) -> band
            // offset of the predecessor's diagonal -> new window index.  A warp does not speculate past a
            // branch, so every way out of the loop shares ONE branch (measured: 6 exits 280 -> 1 exit 220
            // cycles per record at K = 1; `volatile` / hoisted offset loads were slower).
            int x = wst[0], y = wst[1], cnt = wst[2], st = SVX_ST_OK;
            int ai = x + y - lo;                              // window row of the current node
            int b = y - wboff[ai];
            if (b < 0 || b >= B) st = SVX_ST_LEFT_BAND;
            while (st == SVX_ST_OK && (x | y) != 0) {
                const int code = wbp[ai * B + b];
                int wo[NH];
#pragma unroll
                for (int q = 0; q < NH; ++q) wo[q] = wboff[max(ai - 1 - q, 0)];
                int dd;
                if constexpr (T + 2 <= 8) dd = (int)(kDecPacked >> (8 * (code & 7))) & 0xff;
                else if constexpr (T + 2 <= 16) dd = (int)((code < 8 ? kDecPacked : kDecPackedHi) >> (8 * (code & 7))) & 0xff;
                else dd = tdec[code & 127];
                const int dx = dd & 15, dy = dd >> 4;
                const int px = x - dx, py = y - dy;
                const int sft = dx + dy;
                int pw = wo[0];
#pragma unroll
                for (int q = 1; q < NH; ++q) pw = (sft == q + 1) ? wo[q] : pw;
                const int pb = py - pw;
                // one branch for every way out (a warp does not speculate past a branch): classified afterwards
                if ((code > T + 1) | ((px | py) < 0) | (ai - sft < 0) | ((unsigned)pb >= (unsigned)B)) {
                    if (code > T + 1) st = SVX_ST_NO_BACKPTR;
                    else if ((px | py) < 0) st = SVX_ST_LEFT_BAND;           // reference: 'traceback bug'
                    else if (ai - sft >= 0) st = SVX_ST_LEFT_BAND;           // else: window exhausted, reload
                    break;
                }
                if (cnt < cap) {
                    int2 *r = reinterpret_cast<int2 *>(recs_top - cnt);
                    r[0] = make_int2(x, y); r[1] = make_int2(dx, dy);
                } else st |= SVX_ST_OVERFLOW;
                ++cnt;
                x = px; y = py; ai -= sft; b = pb;
            }
            wst[0] = x; wst[1] = y; wst[2] = cnt; wst[3] = st;
            wst[4] = (st != SVX_ST_OK) || (x == 0 && y == 0);
        }
        __syncthreads();
    }

    // ---- phase 3: scores and the next level's search path -------------------------------------------
#ifdef SVX_DP_TIMING
    tm2 = clock64();
#endif
    const int n = min(wst[2], cap), status = wst[3];
    SvxAlignRec *recs = job.recs + (cap - n);                // document order
    for (int i = tid; i < n; i += blockDim.x) {
        const SvxAlignRec r = recs[i];
        const int a = r.x_end + r.y_end, pa = a - r.nx - r.ny;
        const int b = r.y_end - svx_boff_out(job.ypath, a, w);
        const int pb = (r.y_end - r.ny) - svx_boff_out(job.ypath, pa, w);
        double sc = __dsub_rn(job.csum[(size_t)a * B + b], job.csum[(size_t)pa * B + pb]);   // np.diff
        if (sc < 0.0) sc = 0.0;                                  // np.clip(a_min=0); NaN stays NaN
        if (r.nx == 0 || r.ny == 0) sc = 0.0;
        else sc = __ddiv_rn(__ddiv_rn(sc, (double)r.nx), (double)r.ny);
        recs[i].score = sc;
    }
    if (job.next_ypath && status == SVX_ST_OK) {
        int32_t *np_ = job.next_ypath;
        const int plen = job.next_len;
        // extend_alignments (dp_utils.py:228-258) on the upsampled ids
        const int xmax = s0 > 0 ? 2 * s0 - 1 : 0, ymax = s1 > 0 ? 2 * s1 - 1 : 0;
        const int lenx = job.t0 - xmax > 0 ? job.t0 - xmax : 0;
        const int leny = job.t1 - ymax > 0 ? job.t1 - ymax : 0;
        long long ext_x = 0, ext_y = 0;          // joins the trailing deletion run
        if (lenx > 0 && leny > 0) {
            if (tid == 0) emit_segment(&segq, np_, plen, 2LL * s0, 2LL * s1, lenx, leny);
        } else if (lenx == 0) ext_y = leny;
        else ext_x = lenx;
        for (int i = tid - 1; i < n; i += blockDim.x) {
            // i == -1: the deletion run at the document start (finish()); i >= 0: alignment i and the
            // deletion run that follows it in document order
            long long sx = 0, sy = 0;
            if (i >= 0) {
                const SvxAlignRec r = recs[i];
                if (r.nx == 0 || r.ny == 0) continue;
                emit_segment(&segq, np_, plen, 2LL * (r.x_end - r.nx), 2LL * (r.y_end - r.ny), 2LL * r.nx, 2LL * r.ny);
                sx = 2LL * r.x_end; sy = 2LL * r.y_end;
            }
            long long px = 0, py = 0;
            int j = i + 1;
            for (; j < n; ++j) {
                const int nx = recs[j].nx, ny = recs[j].ny;
                if (nx > 0 && ny > 0) break;
                px += 2LL * nx; py += 2LL * ny;
            }
            if (j == n) { px += ext_x; py += ext_y; }
            emit_segment(&segq, np_, plen, sx, sy, px, py);
        }
        __syncthreads();
        const int nq = min(segq.n, kSegQueue);
        for (int qi = 0; qi < nq; ++qi)
            emit_points(np_, plen, segq.xs[qi], segq.ys[qi], segq.xw[qi], segq.yw[qi], 1 + tid, blockDim.x);
        if (tid == 0 && plen > 0) np_[0] = 0;
    }
    if (tid == 0) {
        *job.status_d = status;
        if (job.nrecs) *job.nrecs = wst[2];
    }
#ifdef SVX_DP_TIMING
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0)
        printf("dp_tri<%d> A=%d: recurrence %lld, walk %lld, scores+path %lld cycles (%d records)\n", K, A, tm1 - tm0, tm2 - tm1,
               clock64() - tm2, n);
#endif
}

inline bool is_standard_types(const SvxBandJob &j, int *k_out)
{
    // make_alignment_types(a): x outer 1..a-1, y inner, x+y <= a  -> K = a-1, T = K(K+1)/2
    int K = 0;
    while (K * (K + 1) / 2 < j.ntypes) ++K;
    if (K * (K + 1) / 2 != j.ntypes || K < 1 || K > 9) return false;
    int t = 0;
    for (int x = 1; x <= K; ++x)
        for (int y = 1; x + y <= K + 1; ++y, ++t)
            if (j.xo[t] != x || j.yo[t] != y) return false;
    if (j.k0 < K || j.k1 < K) return false;
    *k_out = K;
    return true;
}

template <int CX, int CY, bool TRI>
int launch_costs(const SvxBandJob *jobs_d, int nj, int max_alen, int band, int dim, int mode, cudaStream_t st)
{
    const int lb = (band + 7) & ~7;
    // anti-diagonals per tile: as many as 384 threads allow, up to 24 (measured: 24 beats 16 by 7 % at K = 5 -
    // the 2B-1 halo rows of a tile are amortised over more diagonals), 16 if the staging limits say so
    const int maxk = CX > CY ? CX : CY;
    int ta = 384 / lb < 24 ? 384 / lb : 24;
    if (ta < 1) return -1;
    // threads = one per band cell, plus staging-only threads when a narrow band leaves too few of them
    // to move the tile's rows in kMaxItems (8) pieces each
    auto threads_for = [&](int t) {
        const size_t items = (size_t)maxk * (t + 2 * band - 1) * (kBC / 4);
        int need = (int)(((items + 7) / 8 + 31) & ~(size_t)31);
        if (need > 384) need = 384;          // wider: the kernel resolves the copy sources per slice (many_items)
        return need > t * lb ? need : t * lb;
    };
    auto fits = [&](int t) {
        return (size_t)2 * maxk * (t + 2 * band - 1) * kBS * sizeof(float) <= 110 * 1024 &&      // two CTAs per SM
               threads_for(t) <= 384;
    };
    if (!fits(ta) && ta > 16) ta = 16;
    while (ta > 1 && (size_t)2 * maxk * (ta + 2 * band - 1) * kBS * sizeof(float) > 220 * 1024) --ta;   // very wide bands
    const int threads = threads_for(ta);
    if (threads > 384) return -1;
    const size_t smem = (size_t)2 * maxk * (ta + 2 * band - 1) * kBS * sizeof(float);   // two slice buffers
    if (smem > 220 * 1024) return -1;
    dim3 grid((max_alen + ta - 1) / ta, nj);
    if (mode == SVX_COST_EXACT) {
        auto kern = k_banded_costs<CX, CY, TRI, true>;
        if (smem > 36 * 1024) SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(jobs_d, dim, ta, lb);
    } else {
        auto kern = k_banded_costs<CX, CY, TRI, false>;
        if (smem > 36 * 1024) SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(jobs_d, dim, ta, lb);
    }
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

}  // namespace

// All jobs of one call must share (band, type list); the host side groups them (level 0 jobs vs
// coarser-level jobs of a batch differ in their type list).
extern "C" int svx_banded_costs(const SvxBandJob *jobs_d, const SvxBandJob *jobs_h, int njobs, int dim, int mode,
                                void *stream)
{
    SVX_REQUIRE(dim > 0 && dim % kBC == 0, SVX_ERR_UNSUPPORTED, "svx_banded_costs: dim %d must be a multiple of %d", dim, kBC);
    if (njobs <= 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const SvxBandJob &j0 = jobs_h[0];
    int K = 0;
    bool standard = is_standard_types(j0, &K);
    int max_alen = 0;
    for (int j = 0; j < njobs; ++j) {
        const SvxBandJob &jb = jobs_h[j];
        SVX_REQUIRE(jb.band == j0.band && jb.ntypes == j0.ntypes && jb.width_over2 == j0.width_over2, SVX_ERR_ARG,
                    "svx_banded_costs: job %d differs from job 0 in band/ntypes", j);
        SVX_REQUIRE(jb.band == 2 * jb.width_over2, SVX_ERR_ARG, "svx_banded_costs: band != 2*width_over2");
        for (int t = 0; t < jb.ntypes; ++t) {
            SVX_REQUIRE(jb.xo[t] == j0.xo[t] && jb.yo[t] == j0.yo[t], SVX_ERR_ARG, "svx_banded_costs: type lists differ");
            // dp_core.pyx:204-209
            SVX_REQUIRE(jb.xo[t] >= 1 && jb.yo[t] >= 1 && jb.xo[t] <= jb.k0 && jb.yo[t] <= jb.k1 && jb.xo[t] <= 64 && jb.yo[t] <= 64,
                        SVX_ERR_ARG, "svx_banded_costs: type (%d,%d) needs more overlaps than provided (%d,%d)",
                        jb.xo[t], jb.yo[t], jb.k0, jb.k1);
        }
        if (jb.a_len > max_alen) max_alen = jb.a_len;
        // the kernels address rows with 32-bit float offsets from v0 / v1
        SVX_REQUIRE((size_t)jb.k0 * jb.s0 * dim < ((size_t)1 << 31) && (size_t)jb.k1 * jb.s1 * dim < ((size_t)1 << 31),
                    SVX_ERR_UNSUPPORTED, "svx_banded_costs: job %d: a side of %d x %d rows exceeds 2^31 floats", j,
                    jb.k0 > jb.k1 ? jb.k0 : jb.k1, jb.s0 > jb.s1 ? jb.s0 : jb.s1);
        int kk;
        if (standard && !(is_standard_types(jb, &kk) && kk == K)) standard = false;
    }
    if (max_alen == 0 || j0.ntypes == 0) return SVX_OK;
    for (int jb0 = 0; jb0 < njobs; jb0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - jb0 < SVX_MAX_GRID_Y ? njobs - jb0 : SVX_MAX_GRID_Y;
        int rc = -1;
        // A/B switches (read once): SVX_COSTS_KERNEL = p2 (default: packed FFMA2 blocks, banded_p2.cu) | blk (scalar
        // 2x2 blocks, K <= 4) | cell (thread per cell); SVX_COSTS_CELL=1 is the older spelling of `cell`
        static const char *which = getenv("SVX_COSTS_KERNEL") ? getenv("SVX_COSTS_KERNEL") : "p2";
        static const bool per_cell = (getenv("SVX_COSTS_CELL") && atoi(getenv("SVX_COSTS_CELL")) != 0) || !strcmp(which, "cell");
        static const bool use_p2 = !per_cell && !strcmp(which, "p2");
        // slice width: 32 floats for K <= 4 (measured 13.6 vs 14.3 ms on config 2), 16 above (a 32-float stage of
        // K >= 5 rows leaves room for two stages only); three ring stages - a fourth one was SLOWER at K = 7
        // (42.6 vs 35.0 ms on config 5: the larger shared-memory carve-out leaves the L1 no room for the 4-byte copies)
        static const int p2_bc_env = getenv("SVX_P2_BC") ? atoi(getenv("SVX_P2_BC")) : 0;
        static const int p2_stages_env = getenv("SVX_P2_STAGES") ? atoi(getenv("SVX_P2_STAGES")) : 0;
        const int p2_bc = p2_bc_env ? p2_bc_env : (K <= 4 ? 32 : 16);
        const int p2_stages = p2_stages_env ? p2_stages_env : 3;
        static const int p2_cw_env = getenv("SVX_P2_CONSUMERS") ? atoi(getenv("SVX_P2_CONSUMERS")) : 0;
        // consumer warps: 12 (three per scheduler, 128 registers) where the accumulators allow it - K <= 4: measured
        // 12.6 vs 13.6 ms on config 2 - else 8 (168 registers)
        const int p2_cw = p2_cw_env ? p2_cw_env : (K <= 4 ? 12 : 8);
        static const int p2_prod = getenv("SVX_P2_PRODUCERS") ? atoi(getenv("SVX_P2_PRODUCERS")) : 4;   // clamped to 12 warps per CTA
        static const int p2_mink = getenv("SVX_P2_MINK") ? atoi(getenv("SVX_P2_MINK")) : 2;   // K = 1 (coarse levels, one type) is copy-bound
        if (standard && use_p2 && K >= p2_mink && K <= 7)
            rc = svx_launch_costs_p2(K, jobs_d + jb0, nj, max_alen, j0.band, dim, mode, p2_bc, p2_stages, p2_prod, p2_cw, st);
        // the scalar register-blocked kernel (round 1): K <= 4 only - at K = 5 its 60 accumulators per thread made
        // the thread-per-cell kernel faster
        if (rc == -1 && standard && !per_cell && K <= 4 && (j0.band & 1) == 0) {
            switch (K) {
#define CASE(KK) case KK: rc = launch_costs_blk<KK>(jobs_d + jb0, nj, max_alen, j0.band, dim, mode, st); break;
                CASE(1) CASE(2) CASE(3) CASE(4)
#undef CASE
                default: break;
            }
        }
        if (rc == -1 && standard) {
            switch (K) {
#define CASE(KK) case KK: rc = launch_costs<KK, KK, true>(jobs_d + jb0, nj, max_alen, j0.band, dim, mode, st); break;
                CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
                default: break;
            }
        }
        // any type list (vecalign.py:165-171 many-to-one lists, hand-written lists): passes over blocks of 4 x 4, or
        // for very wide bands 2 x 2 / 1 x 1, (x overlap, y overlap) pairs with a type lookup table
        if (rc == -1) rc = launch_costs<4, 4, false>(jobs_d + jb0, nj, max_alen, j0.band, dim, mode, st);
        if (rc == -1) rc = launch_costs<2, 2, false>(jobs_d + jb0, nj, max_alen, j0.band, dim, mode, st);
        if (rc == -1) rc = launch_costs<1, 1, false>(jobs_d + jb0, nj, max_alen, j0.band, dim, mode, st);
        SVX_REQUIRE(rc != -1, SVX_ERR_UNSUPPORTED, "svx_banded_costs: band %d too wide for one CTA", j0.band);
        if (rc != SVX_OK) return rc;
    }
    return SVX_OK;
}

static inline int ring_size(int amax)
{
    int r = 2;
    while (r < amax + 1) r <<= 1;
    return r;
}

template <int K>
static int launch_dp_tri(const SvxBandJob *jobs_d, int njobs, int bmax, int amax_len, cudaStream_t st)
{
    constexpr int T = K * (K + 1) / 2;
    const int tb = T * bmax;
    constexpr int NH = K + 1;
    // chunk: as many diagonals as ~96 KB of fp64 cost buffers hold (at most kMaxChunk), a multiple of NH (the
    // recurrence's csum ring has compile-time rows) and even (16-byte staging loads: chunk * tb floats must
    // be a multiple of 4, B is even)
    constexpr int STEP = (NH % 2) ? 2 * NH : NH;
    constexpr int RS = 32 + K + NH + ((32 + K + NH) & 1);
    int chunk = (int)((96 * 1024) / ((size_t)2 * tb * sizeof(double)));
    chunk = chunk > kMaxChunk ? kMaxChunk : chunk;
    chunk = chunk / STEP * STEP;
    if (chunk < STEP) chunk = STEP;
    const size_t dp_bytes = (size_t)2 * (chunk * tb + 32) * sizeof(double) + (size_t)2 * chunk * 32 * sizeof(double) +
                            (size_t)NH * RS * sizeof(double) + (size_t)2 * chunk * (NH + 1) * sizeof(int) + (size_t)2 * chunk * 32 +
                            (size_t)2 * chunk * 32 + (size_t)2 * chunk * 32 * sizeof(double) + (size_t)(chunk + NH) * sizeof(int) + 16;
    if (dp_bytes > 220 * 1024) return -1;
    // walk window: whole job when it fits in ~96 KB, otherwise 96 KB windows
    const int per_diag = bmax + (int)sizeof(int);
    int win = amax_len + 2;
    const int win_cap = (96 * 1024) / per_diag;
    if (win > win_cap) win = win_cap;
    if (win < 64) win = 64;
    const size_t walk_bytes = (size_t)(win + 12) * per_diag + 32;     // + the 8-diagonal alignment slack
    const size_t smem = dp_bytes > walk_bytes ? dp_bytes : walk_bytes;
    auto kern = k_banded_dp_tri<K>;
    if (smem > 40 * 1024) SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<njobs, kDpThreads, smem, st>>>(jobs_d, chunk, win);
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

extern "C" int svx_banded_dp(const SvxBandJob *jobs_d, const SvxBandJob *jobs_h, int njobs, void *stream)
{
    if (njobs <= 0) return SVX_OK;
    int amax = 2, tbmax = 0, bmax = 0, alen_max = 0, K = 0;
    bool standard = true;
    for (int j = 0; j < njobs; ++j) {
        const SvxBandJob &jb = jobs_h[j];
        SVX_REQUIRE(jb.band >= 2 && jb.band <= 256, SVX_ERR_UNSUPPORTED, "svx_banded_dp: band %d must be in [2,256]", jb.band);
        SVX_REQUIRE(jb.ntypes >= 0 && jb.ntypes <= SVX_MAX_TYPES - 2, SVX_ERR_ARG, "svx_banded_dp: too many types");
        SVX_REQUIRE(jb.a_len >= 1, SVX_ERR_ARG, "svx_banded_dp: empty search path");
        for (int t = 0; t < jb.ntypes; ++t) if (jb.xo[t] + jb.yo[t] > amax) amax = jb.xo[t] + jb.yo[t];
        if (jb.ntypes * jb.band > tbmax) tbmax = jb.ntypes * jb.band;
        if (jb.band > bmax) bmax = jb.band;
        if (jb.a_len > alen_max) alen_max = jb.a_len;
        // the register/shuffle kernel needs: the standard type set (same K for the whole launch),
        // one band width, and record storage
        int kk = 0;
        SvxBandJob probe = jb;
        probe.k0 = probe.k1 = 64;          // is_standard_types only checks the list itself here
        if (!(is_standard_types(probe, &kk) && (K == 0 || kk == K) && jb.band == jobs_h[0].band && jb.recs)) standard = false;
        else K = kk;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (standard && bmax + K <= 32) {      // source-lane wrap-around must land on a lane >= band
        int rc = -1;                        // -1: the chunk plan does not fit shared memory (K = 8 with bands >= 20)
        switch (K) {
#define CASE(KK) case KK: rc = launch_dp_tri<KK>(jobs_d, njobs, bmax, alen_max, st); break;
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
#undef CASE
            default: break;
        }
        if (rc != -1) return rc;            // otherwise the generic kernel below handles the shape
    }
    // any other type list: generic kernel (shared-memory ring, run-time type loop)
    const int R = ring_size(amax);
    int chunk = tbmax > 0 ? (int)((96 * 1024) / ((size_t)2 * tbmax * sizeof(float))) : kMaxChunk;
    chunk = chunk > kMaxChunk ? kMaxChunk : (chunk < 2 ? 2 : chunk);
    const size_t smem = (size_t)R * bmax * sizeof(double) + (size_t)2 * chunk * tbmax * sizeof(float) +
                        (size_t)2 * (chunk + R) * sizeof(int) + 16;
    SVX_REQUIRE(smem <= 220 * 1024, SVX_ERR_UNSUPPORTED, "svx_banded_dp: %zu B of shared memory needed", smem);
    if (smem > 48 * 1024)
        SVX_CUDA_OK(cudaFuncSetAttribute(k_banded_dp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_banded_dp<<<njobs, 128, smem, st>>>(jobs_d, R, chunk);
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

// Host twin of k_banded_dp: same node function and walk, serial loops, host pointers.
extern "C" int svx_host_banded_dp(const SvxBandJob *job)
{
    SVX_REQUIRE(job && job->bp && job->csum && job->ypath && job->status_d && job->del_penalty, SVX_ERR_ARG,
                "svx_host_banded_dp: null pointer");
    const int B = job->band, T = job->ntypes, A = job->a_len, w = job->width_over2;
    const double pen = *job->del_penalty;
    for (int aa = 0; aa < A + 2; ++aa) {
        for (int bb = 0; bb < B; ++bb) {
            int bp;
            auto boff = [&](int q) { return svx_boff_out(job->ypath, q, w); };
            auto csum_at = [&](int aq, int bq) { return job->csum[(size_t)aq * B + bq]; };
            auto cost_at = [&](int t) { return job->costs[((size_t)(aa - 2) * T + t) * B + bb]; };
            const double v = svx_band_node(aa, bb, job->s0, job->s1, A, B, T, job->xo, job->yo, pen, boff, csum_at, cost_at, &bp);
            job->csum[(size_t)aa * B + bb] = v;
            job->bp[(size_t)aa * B + bb] = (uint8_t)bp;
        }
    }
    RecWriter rw{job->recs, job->rec_cap, 0, 0, 0};
    int st;
    if (job->next_ypath) {
        SvxSerialEmit emit{job->next_ypath, job->next_len};
        SvxPathBuilder<SvxSerialEmit> pb(emit);
        pb.begin(job->s0, job->s1, job->t0, job->t1, 1);
        OnAlignDev<SvxPathBuilder<SvxSerialEmit>> on{&rw, &pb};
        st = svx_band_walk(job->bp, job->csum, job->ypath, w, job->s0, job->s1, A, B, T, job->xo, job->yo, on);
        if (st == SVX_ST_OK) pb.finish();
        if (job->next_len > 0) job->next_ypath[0] = 0;
    } else {
        OnAlignDev<SvxPathBuilder<SvxSerialEmit>> on{&rw, nullptr};
        st = svx_band_walk(job->bp, job->csum, job->ypath, w, job->s0, job->s1, A, B, T, job->xo, job->yo, on);
    }
    if (rw.overflow) st |= SVX_ST_OVERFLOW;
    *job->status_d = st;
    if (job->nrecs) *job->nrecs = rw.count;
    return SVX_OK;
}
