// dense_tc.cu — coarsest-level cost matrix on the 5th-generation tensor cores (sm_100a):
//   svx_dense_costs(mode = SVX_COST_TC)   (dp_core.pyx:36-77 make_dense_costs)
//
// costs[x, y] = 2 (1 - v0[x].v1[y]) / (1e-6 + n0[x] + n1[y]) is a true dense (s0 x 1024) . (1024 x s1)
// contraction per document pair, computed in 128 x 128 tiles.
//
// 3xTF32: a = hi + lo with hi = the upper 19 bits of a (exactly a TF32 number) and lo = a - hi (exact in fp32);
// a.b ~ hi.hi + hi.lo + lo.hi (the lo.lo term is < 2^-22 relative).  kind::tf32 reads only the upper 19 bits of a 32-bit
// operand (checked on the device: the results are identical with and without masking the operand in shared memory), so
// the RAW fp32 rows are the hi plane, and only the residual plane lo is materialised - once per level
// (k_dense_residual, SvxDenseJob.lo0 / lo1), not once per tile: the GEMM kernel has no CUDA-core work in its k loop.
//
// Persistent, warp-specialised CTAs (one per SM) walk the launch's 128 x 128 tiles:
//   TMA      (warp 0) four cp.async.bulk.tensor.2d loads per k-slice - 128 rows x 32 floats (128 B, SWIZZLE_128B) of the
//            raw and the residual plane of each operand - into a 3-stage shared-memory ring, completion on the stage's
//            `full` mbarrier; rows past the end of a document are zero-filled by the TMA unit.  Two CUtensorMaps per
//            operand per pair, encoded on the host (svx_dense_tmaps_encode) and shipped with the job descriptors.
//   tcgen05  (warp 1, one thread) per 8-wide k step three kind::tf32 MMAs (hi.hi, hi.lo, lo.hi) accumulating the
//            128 x 128 fp32 tile in TMEM (three 128-column accumulators), and commits to the stage's `free` mbarrier.
//   epilogue (warps 2-9) tcgen05.ld 32x32b brings 32 accumulator rows x 32 columns per instruction to registers (two
//            warps per TMEM lane quarter, 64 columns each); the cost formula is evaluated in double exactly as in the
//            reference and stored with the raw dots.  The accumulators are handed back to the MMA warp as soon as they
//            have been read, before the formula of the last columns; the ring refills during the epilogue.
//
// 3xTF32 tolerance (tests/test_gpu_tensor_core.py): |dot - fp32 sequential dot| <= 4e-6 on unit
// vectors, i.e. costs within 1e-5 absolute; the coarse alignment path is checked to stay identical.
#include <string.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "svx_common.cuh"

namespace {

constexpr int kTM = 128;             // tile rows (x), UMMA M
constexpr int kTN = 128;             // tile cols (y), UMMA N
constexpr int kTK = 32;              // floats per k-slice = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kTileBytes = kTM * kTK * 4;          // 16 KB per operand per stage
constexpr int kStageBytes = 4 * kTileBytes;        // A raw (= hi), B raw, A residual, B residual
constexpr int kEpiWarps = 8;
constexpr int kXposeFloats = 32 * 33;             // per epilogue warp: a 32 x 32 chunk, rows padded to 33 floats
constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kXposeFloats * 4 + 1024 /* alignment slack */;
constexpr int kTmemCols = 512;          // three 128-column fp32 accumulators (allocation must be a power of 2)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(unsigned dst, const void *tmap, unsigned bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO = 64 x 16 B),
// LBO is 1 for swizzled K-major layouts, descriptor version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTN >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, uint64_t da, uint64_t db, unsigned accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

// residual planes of the launch's operands: lo = a - (a with the 13 low mantissa bits cleared), exact in fp32
__global__ void __launch_bounds__(256) k_dense_residual(const SvxDenseJob *jobs, int dim)
{
    const SvxDenseJob &job = jobs[blockIdx.y];
    const long long n0 = (long long)job.s0 * dim / 4, n1 = (long long)job.s1 * dim / 4;
    const float4 *a0 = reinterpret_cast<const float4 *>(job.v0), *a1 = reinterpret_cast<const float4 *>(job.v1);
    float4 *l0 = reinterpret_cast<float4 *>(job.lo0), *l1 = reinterpret_cast<float4 *>(job.lo1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = i < n0 ? __ldg(a0 + i) : __ldg(a1 + (i - n0));
        float4 lo;
        lo.x = __fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
        lo.y = __fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
        lo.z = __fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
        lo.w = __fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
        if (i < n0) l0[i] = lo; else l1[i - n0] = lo;
    }
}

// Warp-specialised, persistent: CTA b takes tiles b, b + gridDim.x, ... of the launch's (job, tile row, tile column)
// grid.  warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue.  Two mbarriers per ring stage: full (the
// bytes of the four operand slices have landed) -> free (tcgen05.commit: the MMAs have read the stage); `acc` hands a
// finished tile to the epilogue, `drained` hands the accumulators back.
constexpr int kTcThreads = 32 * (2 + kEpiWarps);

__global__ void __launch_bounds__(kTcThreads, 1) k_dense_costs_tc(const SvxDenseJob *jobs, int dim, int tiles_x, int tiles_y, int njobs)
{
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long bars[2 * kStages + 2];   // full[], free[], accumulators done, accumulators drained
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 1024-byte alignment (SWIZZLE_128B atoms) by an OFFSET into the dynamic shared array
    const unsigned pad = (1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u;
    const unsigned tiles_u32 = smem_u32(smem_dyn) + pad;
    float *xpose = reinterpret_cast<float *>(smem_dyn + pad + kStages * kStageBytes);     // stays a shared-space pointer
    const unsigned bar_full = smem_u32(&bars[0]), bar_free = smem_u32(&bars[kStages]);
    const unsigned bar_acc = smem_u32(&bars[2 * kStages]), bar_drained = smem_u32(&bars[2 * kStages + 1]);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_free + 8 * s, 1); }
        mbar_init(bar_acc, 1);
        mbar_init(bar_drained, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const unsigned tmem_d = tmem_slot;
    const int nk = dim / kTK;
    const int per_job = tiles_x * tiles_y, ntiles = per_job * njobs;

    // every role walks the same tile list and skips the tiles outside its job's matrix
    auto tile_of = [&](int t, int &j, int &x0, int &y0) {
        j = t / per_job;
        const int r = t - j * per_job;
        x0 = (r / tiles_x) * kTM;
        y0 = (r % tiles_x) * kTN;
        return x0 < jobs[j].s0 && y0 < jobs[j].s1;
    };

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int j, x0, y0;
                if (!tile_of(t, j, x0, y0)) continue;
                const unsigned char *tm0 = static_cast<const unsigned char *>(jobs[j].tmap0);
                const unsigned char *tm1 = static_cast<const unsigned char *>(jobs[j].tmap1);
                for (int ks = 0; ks < nk; ++ks, ++it) {
                    const int s = it % kStages;
                    mbar_wait(bar_free + 8 * s, ((it / kStages) & 1) ^ 1);
                    const unsigned base = tiles_u32 + s * kStageBytes, full = bar_full + 8 * s;
                    mbar_expect_tx(full, 4 * kTileBytes);
                    tma_load_2d(base, tm0, full, ks * kTK, x0);
                    tma_load_2d(base + kTileBytes, tm1, full, ks * kTK, y0);
                    tma_load_2d(base + 2 * kTileBytes, tm0 + 128, full, ks * kTK, x0);
                    tma_load_2d(base + 3 * kTileBytes, tm1 + 128, full, ks * kTK, y0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int it = 0, nt = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int j, x0, y0;
                if (!tile_of(t, j, x0, y0)) continue;
                mbar_wait(bar_drained, (nt & 1) ^ 1);                    // the previous tile's epilogue has read the accumulators
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                for (int ks = 0; ks < nk; ++ks, ++it) {
                    const int s = it % kStages;
                    mbar_wait(bar_full + 8 * s, (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                    const unsigned base = tiles_u32 + s * kStageBytes;
                    const unsigned a_hi = base, b_hi = base + kTileBytes, a_lo = base + 2 * kTileBytes, b_lo = base + 3 * kTileBytes;
                    // The tensor core rounds toward zero each time it adds an MMA result into the fp32
                    // accumulator, so the error grows with the number of MMAs per accumulator.  The hi.hi
                    // products alternate between two accumulators (even / odd k-slices) and the two small
                    // cross terms (2^-11 of the main term) go to a third one; the epilogue adds the three.
                    const unsigned d_main = tmem_d + (unsigned)((ks & 1) * kTN);
                    const unsigned d_cross = tmem_d + (unsigned)(2 * kTN);
#pragma unroll
                    for (int k = 0; k < kTK / 8; ++k) {                // UMMA K = 8 tf32 = 32 bytes along the row
                        const unsigned off = k * 32;
                        umma_tf32(d_main, umma_desc(a_hi + off), umma_desc(b_hi + off), (ks >= 2) || k != 0);
                        umma_tf32(d_cross, umma_desc(a_hi + off), umma_desc(b_lo + off), (ks | k) != 0);
                        umma_tf32(d_cross, umma_desc(a_lo + off), umma_desc(b_hi + off), 1u);
                    }
                    umma_commit(bar_free + 8 * s);                     // arrives when these MMAs have read the stage
                }
                umma_commit(bar_acc);
                ++nt;
            }
        }
    } else {
        // ---- epilogue: TMEM lanes [32 (warp % 4), +32) belong to this warp; warps 2-5 take columns 0-63, 6-9 columns 64-127.
        // tcgen05.ld gives a thread ITS ROW (32 columns); storing from there would write 32 different rows per instruction
        // (ncu: the epilogue, not the MMAs, set the tile time: 55 us per tile).  Each 32 x 32 chunk is transposed through a
        // padded shared-memory buffer so that a lane owns a COLUMN: 128-byte row segments per store, n1[y] loaded once.
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        float *xp = xpose + (warp - 2) * kXposeFloats;
        int nt = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            int j, x0, y0;
            if (!tile_of(t, j, x0, y0)) continue;
            const SvxDenseJob &job = jobs[j];
            const int xr0 = x0 + quarter * 32;                        // first row of this warp
            const float nx_lane = xr0 + lane < job.s0 ? job.n0[xr0 + lane] : 1.0f;
            const int nrows = min(32, job.s0 - xr0);
            mbar_wait(bar_acc, nt & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            constexpr int kCols = kTN / 2;
            const int cbase = chalf * kCols;
            // both 32-column chunks of this warp go to registers first, so that the accumulators return to the MMA warp
            // before any formula runs
            float dot[2][32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int acc = 0; acc < 3; ++acc) {
                    if (acc == 1 && nk < 2) continue;                 // odd-slice accumulator never written
                    uint32_t r[32];
                    const unsigned taddr = tmem_d + ((unsigned)(quarter * 32) << 16) + (unsigned)(acc * kTN + cbase + 32 * h);
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
#pragma unroll
                    for (int q = 0; q < 32; ++q) dot[h][q] = acc == 0 ? __uint_as_float(r[q]) : __fadd_rn(dot[h][q], __uint_as_float(r[q]));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_drained);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int q = 0; q < 32; ++q) xp[lane * 33 + q] = dot[h][q];
                __syncwarp();
                const int y = y0 + cbase + 32 * h + lane;
                const bool yok = y < job.s1;
                const float ny = yok ? job.n1[y] : 1.0f;
                float *crow = job.costs + (size_t)xr0 * job.s1 + y;
                float *drow = job.dots ? job.dots + (size_t)xr0 * job.s1 + y : nullptr;
                // four rows at a time: the fp64 division is a long dependent chain, four independent ones fill the pipe
                for (int r = 0; r < nrows; r += 4) {
                    float d[4], nx[4], c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        d[u] = xp[(r + u) * 33 + lane];            // rows past nrows hold the tile's zero padding
                        nx[u] = __shfl_sync(0xffffffffu, nx_lane, (r + u) & 31);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[u] = svx_dense_cost(d[u], nx[u], ny);
                    if (yok) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (r + u < nrows) {
                                crow[(size_t)(r + u) * job.s1] = c[u];
                                if (drow) drow[(size_t)(r + u) * job.s1] = d[u];
                            }
                        }
                    }
                }
                __syncwarp();
            }
            ++nt;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(kTmemCols));
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

}  // namespace

// Host: two CUtensorMaps (128 B each) per operand per job - the raw rows and the residual plane - at out_host[4 j + ...]:
// {v0, lo0, v1, lo1}.  The caller copies them to the device and points SvxDenseJob.tmap0 at the first pair and tmap1 at
// the second (64-byte aligned).
extern "C" int svx_dense_tmaps_encode(const SvxDenseJob *jobs_h, int njobs, int dim, void *out_host)
{
    SVX_REQUIRE(dim > 0 && dim % kTK == 0, SVX_ERR_UNSUPPORTED, "svx_dense_tmaps_encode: dim %d must be a multiple of %d", dim, kTK);
    SVX_REQUIRE(out_host || njobs <= 0, SVX_ERR_ARG, "svx_dense_tmaps_encode: null output");
    auto enc = encode_fn();
    SVX_REQUIRE(enc, SVX_ERR_CUDA, "svx_dense_tmaps_encode: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap *out = reinterpret_cast<CUtensorMap *>(out_host);
    for (int j = 0; j < njobs; ++j) {
        for (int which = 0; which < 4; ++which) {
            const SvxDenseJob &jb = jobs_h[j];
            const float *base = which == 0 ? jb.v0 : which == 1 ? jb.lo0 : which == 2 ? jb.v1 : jb.lo1;
            const int rows = which < 2 ? jb.s0 : jb.s1;
            CUtensorMap *m = out + 4 * j + which;
            memset(m, 0, sizeof(*m));
            if (rows <= 0) continue;
            SVX_REQUIRE(base, SVX_ERR_ARG, "svx_dense_tmaps_encode: job %d has no %s", j,
                        (which & 1) ? "residual plane (SvxDenseJob.lo0 / lo1)" : "operand rows");
            SVX_REQUIRE(((uintptr_t)base & 15) == 0, SVX_ERR_ARG, "svx_dense_tmaps_encode: operand of job %d is not 16-byte aligned", j);
            cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
            cuuint64_t gstride[1] = {(cuuint64_t)dim * sizeof(float)};
            cuuint32_t box[2] = {(cuuint32_t)kTK, (cuuint32_t)kTM};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            SVX_REQUIRE(r == CUDA_SUCCESS, SVX_ERR_CUDA, "svx_dense_tmaps_encode: cuTensorMapEncodeTiled failed (%d) for job %d", (int)r, j);
        }
    }
    return SVX_OK;
}

int svx_dense_costs_tc_launch(const SvxDenseJob *jobs_d, const SvxDenseJob *jobs_h, int njobs, int dim, cudaStream_t st)
{
    SVX_REQUIRE(dim > 0 && dim % kTK == 0, SVX_ERR_UNSUPPORTED, "svx_dense_costs: dim %d must be a multiple of %d", dim, kTK);
    static bool attr_set = false;
    if (!attr_set) {
        SVX_CUDA_OK(cudaFuncSetAttribute(k_dense_costs_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    for (int j0 = 0; j0 < njobs; j0 += SVX_MAX_GRID_Y) {
        const int nj = njobs - j0 < SVX_MAX_GRID_Y ? njobs - j0 : SVX_MAX_GRID_Y;
        int m0 = 0, m1 = 0;
        for (int j = 0; j < nj; ++j) {
            const SvxDenseJob &jb = jobs_h[j0 + j];
            SVX_REQUIRE(jb.s0 == 0 || jb.s1 == 0 || (jb.tmap0 && jb.tmap1 && jb.lo0 && jb.lo1), SVX_ERR_ARG,
                        "svx_dense_costs: tensor-core mode needs SvxDenseJob.tmap0/tmap1 (svx_dense_tmaps_encode) and lo0/lo1");
            if (jb.s0 > 0 && jb.s1 > 0) { m0 = jb.s0 > m0 ? jb.s0 : m0; m1 = jb.s1 > m1 ? jb.s1 : m1; }
        }
        if (m0 == 0 || m1 == 0) continue;
        const int tx = (m1 + kTN - 1) / kTN, ty = (m0 + kTM - 1) / kTM;
        const long long ntiles = (long long)tx * ty * nj;
        static int sms = 0;
        if (!sms) {
            int devid = 0;
            SVX_CUDA_OK(cudaGetDevice(&devid));
            SVX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid));
        }
        {
            long long rows = (long long)(m0 + m1) * dim / 4 / 256;     // float4s of the largest job / block size
            dim3 rg((unsigned)(rows < 1 ? 1 : rows > 64 ? 64 : rows), nj);
            k_dense_residual<<<rg, 256, 0, st>>>(jobs_d + j0, dim);
            SVX_LAUNCH_CHECK();
        }
        const int grid = (int)(ntiles < sms ? ntiles : sms);          // persistent: one CTA per SM walks the tile list
        k_dense_costs_tc<<<grid, kTcThreads, kSmemBytes, st>>>(jobs_d + j0, dim, tx, ty, nj);
        SVX_LAUNCH_CHECK();
    }
    return SVX_OK;
}
