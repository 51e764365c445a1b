// margin.cu — margin scoring of aligned segment pairs:
//   svx_margin_scores   (svecalign/postprocess/score_align.py:124-161 compute_sim_with_nonflat_idx)
//
// For pairs (x_i, y_i) and the two indexed collections X_base, Y_base (score_align.py searches faiss indexes populated
// by prep_index.py:153-183 with L2-normalised embeddings; the shipped example uses `Flat` indexes, i.e. exact search,
// held by faiss-gpu in fp16 - `--gpu_type fp16-shard`):
//     a_i     = <x_i, y_i>                                   (L2-normalised fp32 rows, faiss.normalize_L2)
//     Avg_xy  = mean of the k smallest |x_i - y|^2 over y in Y_base;  Avg_yx likewise for y_i over X_base
//     b_i     = ((2 - Avg_xy) / 2 + (2 - Avg_yx) / 2) / 2    (inplace_l2_to_cosine)
//     score_i = a_i / b_i  (margin = ratio)  or  a_i - b_i  (distance)
//
// The search is a dense contraction  S = Q . B^T  (n x 1024 x m) followed by a per-row top-k - the one place on this
// path where the 5th-generation tensor cores are fed well:
//   k_margin_prepare  warp per row: L2-normalise in fp32, round to fp16 (the flat index's storage type), squared norm of
//                     the ROUNDED row (what a flat L2 index adds to -2 q.b)
//   k_margin_knn      persistent CTA per 128 query rows, sweeping the base in 256-row tiles.  Warp-specialised:
//                     warp 0 = TMA producer (cp.async.bulk.tensor.2d, SWIZZLE_128B, 4-stage ring of 128x64 + 256x64 fp16
//                     k-slices, full/empty mbarriers); warp 1 = MMA issuer (tcgen05.mma.cta_group::1.kind::f16, M = 128,
//                     N = 256, fp32 accumulators in TMEM, two 256-column accumulators so that tile t + 1 is multiplied
//                     while tile t is being read); warps 2-5 = epilogue: tcgen05.ld 32x32b brings each thread ITS query
//                     row, which keeps a sorted top-16 of  key = q.b - |b|^2 / 2  in registers across the whole sweep
//                     (a value enters only if it beats the current 16th: after the first tiles almost none do).
//   k_margin_finish   a_i, b_i, score_i
#include <stdlib.h>
#include <string.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include "svx_common.cuh"

namespace {

constexpr int kQM = 128;                 // query rows per CTA = UMMA M = TMEM lanes
constexpr int kBN = 256;                 // base rows per tile = UMMA N
constexpr int kBK = 64;                  // fp16 elements per k-slice = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kABytes = kQM * kBK * 2;   // 16 KB
constexpr int kBBytes = kBN * kBK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */;
constexpr int kTmemCols = 512;           // two fp32 accumulators of 256 columns
constexpr int kTopK = 16;
constexpr int kThreads = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *tmap, unsigned bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// The same load delivered to the same shared-memory offset (and mbarrier offset) of every CTA of the cluster in `mask`
__device__ __forceinline__ void tma_load_2d_multicast(unsigned dst, const CUtensorMap *tmap, unsigned bar, int c0, int c1, unsigned short mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;\n" ::
            "r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO = 64 x 16 B), LBO = 1,
// descriptor version 1 (Blackwell), layout type 2 (same geometry as dense_tc.cu: 64 fp16 = 32 tf32 = 128 bytes)
__device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: fp16 x fp16 (a_format = b_format = 0), fp32 accumulate (c_format = 1), both K-major, N = 256, M = 128
constexpr uint32_t kInstrDesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kQM >> 4) << 24);

__device__ __forceinline__ void umma_f16(unsigned tmem_d, uint64_t da, uint64_t db, unsigned accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// arrives on the barrier at this offset in every CTA of `mask` (a stage is free once BOTH CTAs of a pair have read it)
__device__ __forceinline__ void umma_commit_multicast(unsigned bar, unsigned short mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar), "h"(mask)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Rows -> L2-normalised fp16 rows + squared norm of the rounded row.  Warp per row.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool FP16>
__device__ __forceinline__ float load_elem(const void *rows, size_t i)
{
    if (FP16) return __half2float(reinterpret_cast<const __half *>(rows)[i]);
    return reinterpret_cast<const float *>(rows)[i];
}

template <bool FP16>
__global__ void __launch_bounds__(256) k_margin_prepare(const void *rows, int n, int dim, __half *out, float *n2)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const size_t base = (size_t)row * dim;
    float ss = 0.f;
    for (int d = lane; d < dim; d += 32) { const float v = load_elem<FP16>(rows, base + d); ss += v * v; }
    ss = warp_sum(ss);
    const float inv = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;          // faiss.normalize_L2 leaves all-zero rows alone
    float hs = 0.f;
    for (int d = lane; d < dim; d += 32) {
        const __half h = __float2half_rn(load_elem<FP16>(rows, base + d) * inv);
        out[base + d] = h;
        const float f = __half2float(h);
        hs += f * f;
    }
    hs = warp_sum(hs);
    if (lane == 0) n2[row] = hs;
}

// a_i = <x_i / |x_i|, y_i / |y_i|> in fp32 (score_align.py:152 np.dot on the normalised rows)
template <bool FP16>
__global__ void __launch_bounds__(256) k_margin_pair_dot(const void *x, const void *y, int n, int dim, float *dot)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const size_t base = (size_t)row * dim;
    float sx = 0.f, sy = 0.f;
    for (int d = lane; d < dim; d += 32) {
        const float a = load_elem<FP16>(x, base + d), b = load_elem<FP16>(y, base + d);
        sx += a * a; sy += b * b;
    }
    sx = warp_sum(sx); sy = warp_sum(sy);
    const float ix = sx > 0.f ? 1.0f / sqrtf(sx) : 0.f, iy = sy > 0.f ? 1.0f / sqrtf(sy) : 0.f;
    float acc = 0.f;
    for (int d = lane; d < dim; d += 32) acc += (load_elem<FP16>(x, base + d) * ix) * (load_elem<FP16>(y, base + d) * iy);
    acc = warp_sum(acc);
    if (lane == 0) dot[row] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// kNN: avg[i] = mean of the k smallest |q_i - b_j|^2 = qn2[i] - 2 * mean of the k largest (q_i.b_j - bn2[j] / 2)
// ---------------------------------------------------------------------------------------------------------------------
// PAIR: the CTAs run as clusters of two that sweep the base in lockstep; each loads ONE half (128 rows) of every base
// tile and multicasts it into both CTAs' shared memory, so the L2 -> SM traffic per CTA and k-slice drops from 48 KB
// (16 KB of queries + 32 KB of base) to 32 KB.  The MMAs stay per CTA (cta_group::1, own queries x the whole tile).
template <bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
k_margin_knn(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_bh,
             const float *qn2, const float *bn2, int nq, int nb, int dim, int k, float *avg_out)
{
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long bars[2 * kStages + 4];   // full[], empty[], tmem_full[2], tmem_empty[2]
    __shared__ unsigned tmem_slot;
    __shared__ __align__(16) float hb[2][kBN];                         // -bn2 / 2 of the tile's columns (-inf past the end)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *tiles = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const unsigned tiles_u32 = smem_u32(tiles);
    const unsigned bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kStages]);
    const unsigned bar_tfull = smem_u32(&bars[2 * kStages]), bar_tempty = smem_u32(&bars[2 * kStages + 2]);
    const int q0 = blockIdx.x * kQM;
    const int ntiles = (nb + kBN - 1) / kBN, nk = dim / kBK;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, PAIR ? 2 : 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (PAIR) cluster_sync_all();          // both CTAs' barriers exist before either multicasts into the other
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const unsigned tmem_d = tmem_slot;
    const unsigned rank = PAIR ? cluster_ctarank() : 0u;

    if (warp == 0) {
        // ---- TMA producer ------------------------------------------------------------------------------------
        if (lane == 0) {
            int it = 0;
            for (int t = 0; t < ntiles; ++t)
                for (int ks = 0; ks < nk; ++ks, ++it) {
                    const int s = it % kStages;
                    mbar_wait(bar_empty + 8 * s, ((it / kStages) & 1) ^ 1);          // first lap: passes immediately
                    const unsigned base = tiles_u32 + s * kStageBytes;
                    mbar_expect_tx(bar_full + 8 * s, kStageBytes);                   // own queries + both halves of the base tile
                    tma_load_2d(base, &tm_q, bar_full + 8 * s, ks * kBK, q0);
                    if (PAIR)
                        tma_load_2d_multicast(base + kABytes + rank * (kBBytes / 2), &tm_bh, bar_full + 8 * s, ks * kBK,
                                              t * kBN + (int)rank * (kBN / 2), (unsigned short)3);
                    else
                        tma_load_2d(base + kABytes, &tm_b, bar_full + 8 * s, ks * kBK, t * kBN);
                }
        }
    } else if (warp == 1) {
        // ---- MMA issuer --------------------------------------------------------------------------------------
        if (lane == 0) {
            int it = 0;
            for (int t = 0; t < ntiles; ++t) {
                const int buf = t & 1;
                mbar_wait(bar_tempty + 8 * buf, ((t >> 1) & 1) ^ 1);                 // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                const unsigned d = tmem_d + (unsigned)(buf * kBN);
                for (int ks = 0; ks < nk; ++ks, ++it) {
                    const int s = it % kStages;
                    mbar_wait(bar_full + 8 * s, (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                    const unsigned a = tiles_u32 + s * kStageBytes, b = a + kABytes;
#pragma unroll
                    for (int kk = 0; kk < kBK / 16; ++kk)                            // UMMA K = 16 fp16 = 32 bytes along the row
                        umma_f16(d, umma_desc(a + kk * 32), umma_desc(b + kk * 32), (ks | kk) != 0);
                    if (PAIR) umma_commit_multicast(bar_empty + 8 * s, (unsigned short)3);   // ... in both CTAs: the peer writes into this stage too
                    else umma_commit(bar_empty + 8 * s);                             // the stage is free once these MMAs have read it
                }
                umma_commit(bar_tfull + 8 * buf);                                    // the accumulator is complete
            }
        }
    } else {
        // ---- epilogue: thread = query row; TMEM lanes [32 (warp % 4), +32) belong to this warp --------------------
        const int et = tid - 64;                                   // 0..127
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        float top[kTopK];
#pragma unroll
        for (int i = 0; i < kTopK; ++i) top[i] = -INFINITY;
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            for (int c = et; c < kBN; c += 128) {
                const int j = t * kBN + c;
                hb[buf][c] = j < nb ? -0.5f * bn2[j] : -INFINITY;
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");        // hb[buf] is complete (and hb[buf ^ 1] no longer read)
            mbar_wait(bar_tfull + 8 * buf, (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
#pragma unroll 1
            for (int c0 = 0; c0 < kBN; c0 += 32) {
                uint32_t r[32];
                const unsigned taddr = tmem_d + ((unsigned)(quarter * 32) << 16) + (unsigned)(buf * kBN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                // keys of the 32 columns and their maximum: after the first tiles almost no chunk holds a candidate, and the
                // common path is then 8 LDS.128 + 32 FADD + 32 FMNMX and one branch (an insertion inlined per column made the
                // epilogue instruction-fetch bound - ncu: stall_no_inst at every reconvergence point - and the MMA warp
                // waited for its accumulators)
                float v[32];
                float vmax = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 h4 = *reinterpret_cast<const float4 *>(&hb[buf][c0 + j]);
                    v[j] = __uint_as_float(r[j]) + h4.x; v[j + 1] = __uint_as_float(r[j + 1]) + h4.y;
                    v[j + 2] = __uint_as_float(r[j + 2]) + h4.z; v[j + 3] = __uint_as_float(r[j + 3]) + h4.w;
                    vmax = fmaxf(vmax, fmaxf(fmaxf(v[j], v[j + 1]), fmaxf(v[j + 2], v[j + 3])));
                }
                // rare path: take the chunk's maxima one by one while they still beat the current k-th best
                while (vmax > top[kTopK - 1]) {
                    int cnt = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        cnt += v[j] == vmax;
                        v[j] = v[j] == vmax ? -INFINITY : v[j];
                    }
                    for (int c = 0; c < cnt && vmax > top[kTopK - 1]; ++c) {        // equal keys (duplicate vectors) each count
#pragma unroll
                        for (int i = kTopK - 1; i > 0; --i) top[i] = vmax > top[i - 1] ? top[i - 1] : (vmax > top[i] ? vmax : top[i]);
                        top[0] = fmaxf(top[0], vmax);
                    }
                    vmax = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, v[j]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        if (q0 + row < nq) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kTopK; ++i) s += i < k ? top[i] : 0.f;   // descending keys = ascending distances, as faiss returns them
            avg_out[q0 + row] = qn2[q0 + row] - 2.0f * (s / (float)k);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (PAIR) cluster_sync_all();          // the peer may still be multicasting into / arriving on this CTA's shared memory
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(kTmemCols));
}

__global__ void k_margin_finish(const float *dot, const float *avg_xy, const float *avg_yx, int n, int margin, float *scores)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float cxy = (2.0f - avg_xy[i]) / 2.0f, cyx = (2.0f - avg_yx[i]) / 2.0f;      // inplace_l2_to_cosine
    const float b = (cxy + cyx) / 2.0f;
    scores[i] = margin == 0 ? dot[i] / b : dot[i] - b;
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

int make_map(CUtensorMap *m, const __half *base, int rows, int dim, int box_rows)
{
    auto enc = encode_fn();
    SVX_REQUIRE(enc, SVX_ERR_CUDA, "svx_margin_scores: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SVX_REQUIRE(r == CUDA_SUCCESS, SVX_ERR_CUDA, "svx_margin_scores: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SVX_OK;
}

inline int64_t up256(int64_t v) { return (v + 255) / 256 * 256; }

struct Layout {
    int64_t xh, yh, xbh, ybh, xn2, yn2, xbn2, ybn2, dot, axy, ayx, total;
};

Layout layout(int n, int nxb, int nyb, int dim, bool xb_same, bool yb_same)
{
    Layout L;
    int64_t top = 0;
    auto take = [&](int64_t bytes) { const int64_t o = top; top += up256(bytes); return o; };
    L.xh = take((int64_t)n * dim * 2);
    L.yh = take((int64_t)n * dim * 2);
    L.xbh = xb_same ? L.xh : take((int64_t)nxb * dim * 2);
    L.ybh = yb_same ? L.yh : take((int64_t)nyb * dim * 2);
    L.xn2 = take((int64_t)n * 4);
    L.yn2 = take((int64_t)n * 4);
    L.xbn2 = xb_same ? L.xn2 : take((int64_t)nxb * 4);
    L.ybn2 = yb_same ? L.yn2 : take((int64_t)nyb * 4);
    L.dot = take((int64_t)n * 4);
    L.axy = take((int64_t)n * 4);
    L.ayx = take((int64_t)n * 4);
    L.total = top;
    return L;
}

int prepare(const void *rows, int n, int dim, int fp16, __half *out, float *n2, cudaStream_t st)
{
    if (n <= 0) return SVX_OK;
    const int blocks = (n + 7) / 8;
    if (fp16) k_margin_prepare<true><<<blocks, 256, 0, st>>>(rows, n, dim, out, n2);
    else k_margin_prepare<false><<<blocks, 256, 0, st>>>(rows, n, dim, out, n2);
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

int knn(const __half *q, const float *qn2, int nq, const __half *b, const float *bn2, int nb, int dim, int k, float *avg, cudaStream_t st)
{
    CUtensorMap tq, tb, tbh;
    int rc = make_map(&tq, q, nq, dim, kQM);
    if (rc != SVX_OK) return rc;
    if ((rc = make_map(&tb, b, nb, dim, kBN)) != SVX_OK) return rc;
    if ((rc = make_map(&tbh, b, nb, dim, kBN / 2)) != SVX_OK) return rc;          // half tiles: what one CTA of a pair loads
    static const bool pair = !(getenv("SVX_MARGIN_PAIR") && atoi(getenv("SVX_MARGIN_PAIR")) == 0);
    const int blocks = (nq + kQM - 1) / kQM;
    if (pair && blocks >= 2) {
        auto kern = k_margin_knn<true>;
        static bool attr_set = false;
        if (!attr_set) {
            SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
            attr_set = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((blocks + 1) / 2 * 2));      // whole pairs: the odd CTA out multiplies zero rows
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = kSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SVX_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tq, tb, tbh, qn2, bn2, nq, nb, dim, k, avg));
    } else {
        auto kern = k_margin_knn<false>;
        static bool attr_set = false;
        if (!attr_set) {
            SVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
            attr_set = true;
        }
        kern<<<blocks, kThreads, kSmemBytes, st>>>(tq, tb, tbh, qn2, bn2, nq, nb, dim, k, avg);
    }
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}

}  // namespace

extern "C" int svx_margin_workspace_bytes(int n, int n_xbase, int n_ybase, int dim, int64_t *bytes)
{
    SVX_REQUIRE(bytes && n >= 0 && dim > 0, SVX_ERR_ARG, "svx_margin_workspace_bytes: bad arguments");
    *bytes = layout(n, n_xbase > 0 ? n_xbase : n, n_ybase > 0 ? n_ybase : n, dim, n_xbase <= 0, n_ybase <= 0).total;
    return SVX_OK;
}

// x_d / y_d: (n, dim) rows of the aligned pairs; xbase_d / ybase_d: the collections the reference's indexes were populated
// with ((n_xbase, dim), (n_ybase, dim)); NULL = the pairs' own rows (one file scored against itself, the shipped example).
extern "C" int svx_margin_scores(const void *x_d, const void *y_d, int n, const void *xbase_d, int n_xbase, const void *ybase_d,
                                 int n_ybase, int dim, int is_fp16, int k, int margin, float *scores_d, void *workspace_d,
                                 int64_t workspace_bytes, void *stream)
{
    SVX_REQUIRE(n >= 0 && (n == 0 || (x_d && y_d && scores_d && workspace_d)), SVX_ERR_ARG, "svx_margin_scores: null argument");
    SVX_REQUIRE(dim > 0 && dim % kBK == 0, SVX_ERR_UNSUPPORTED, "svx_margin_scores: dim %d must be a multiple of %d", dim, kBK);
    SVX_REQUIRE(k >= 1 && k <= kTopK, SVX_ERR_UNSUPPORTED, "svx_margin_scores: k = %d outside [1, %d]", k, kTopK);
    SVX_REQUIRE(margin == 0 || margin == 1, SVX_ERR_ARG, "Wrong margin type: %d", margin);     // score_align.py:159
    if (n == 0) return SVX_OK;
    const bool xb_same = xbase_d == nullptr || xbase_d == x_d, yb_same = ybase_d == nullptr || ybase_d == y_d;
    const int nxb = xb_same ? n : n_xbase, nyb = yb_same ? n : n_ybase;
    SVX_REQUIRE(nxb >= k && nyb >= k, SVX_ERR_ARG, "svx_margin_scores: fewer than k = %d indexed vectors (%d, %d)", k, nxb, nyb);
    const Layout L = layout(n, nxb, nyb, dim, xb_same, yb_same);
    SVX_REQUIRE(workspace_bytes >= L.total, SVX_ERR_ARG, "svx_margin_scores: workspace of %lld bytes, %lld needed", (long long)workspace_bytes,
                (long long)L.total);
    SVX_REQUIRE(((uintptr_t)workspace_d & 255) == 0, SVX_ERR_ARG, "svx_margin_scores: the workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *w = (char *)workspace_d;
    auto H = [&](int64_t o) { return reinterpret_cast<__half *>(w + o); };
    auto F = [&](int64_t o) { return reinterpret_cast<float *>(w + o); };
    int rc;
    if ((rc = prepare(x_d, n, dim, is_fp16, H(L.xh), F(L.xn2), st)) != SVX_OK) return rc;
    if ((rc = prepare(y_d, n, dim, is_fp16, H(L.yh), F(L.yn2), st)) != SVX_OK) return rc;
    if (!xb_same && (rc = prepare(xbase_d, nxb, dim, is_fp16, H(L.xbh), F(L.xbn2), st)) != SVX_OK) return rc;
    if (!yb_same && (rc = prepare(ybase_d, nyb, dim, is_fp16, H(L.ybh), F(L.ybn2), st)) != SVX_OK) return rc;
    const int blocks = (n + 7) / 8;
    if (is_fp16) k_margin_pair_dot<true><<<blocks, 256, 0, st>>>(x_d, y_d, n, dim, F(L.dot));
    else k_margin_pair_dot<false><<<blocks, 256, 0, st>>>(x_d, y_d, n, dim, F(L.dot));
    SVX_LAUNCH_CHECK();
    // x against the y index, y against the x index (score_align.py:139-142)
    if ((rc = knn(H(L.xh), F(L.xn2), n, H(L.ybh), F(L.ybn2), nyb, dim, k, F(L.axy), st)) != SVX_OK) return rc;
    if ((rc = knn(H(L.yh), F(L.yn2), n, H(L.xbh), F(L.xbn2), nxb, dim, k, F(L.ayx), st)) != SVX_OK) return rc;
    k_margin_finish<<<(n + 255) / 256, 256, 0, st>>>(F(L.dot), F(L.axy), F(L.ayx), n, margin, scores_d);
    SVX_LAUNCH_CHECK();
    return SVX_OK;
}
