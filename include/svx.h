/*
 * include/svx.h — C ABI of libsvx.so, the B200 (sm_100a) implementation of Speech-Vecalign's
 * segment-alignment hot path (svecalign/vecalign/dp_core.pyx + the numeric helpers of
 * svecalign/vecalign/dp_utils.py).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - Every pointer named *_d / inside a job struct is a DEVICE pointer owned by the caller;
 *     kernels never allocate.  `stream` is a cudaStream_t passed as void*.
 *   - Every launcher takes an array of job descriptors that lives in DEVICE memory plus the same
 *     array on the HOST (`jobs_h`, used only to size the grid); a single document pair is a batch of
 *     one job.  Work for different jobs is independent: no collective, no cross-job ordering.
 *   - Return value: 0 = OK, otherwise an SVX_ERR_* code; svx_last_error_string() explains it.
 *     Device-side failures (search path leaves the band, "-42" backpointer reached — the
 *     reference's 'traceback bug' / IndexError conditions, dp_utils.py:123-124) are reported per
 *     job in SvxBandJob.status_d / SvxDenseJob.status_d.
 *   - All fp32 dot products of the cost kernels are evaluated in the reference's order
 *     (sequential multiply-then-add over the embedding dimension, no FMA) when mode ==
 *     SVX_COST_EXACT, so costs are bit-identical to dp_core.pyx given identical inputs;
 *     SVX_COST_FAST contracts to FMA (same order) and is within 2e-6 absolute; SVX_COST_TC
 *     additionally computes the coarsest-level matrix as a 3xTF32 tensor-core GEMM (4e-6).
 *
 * Each entry point cites the reference interface it replaces (file:line under /root/reference).
 */
#ifndef SVX_H_
#define SVX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVX_VERSION 100

#if defined(__GNUC__)
#define SVX_API __attribute__((visibility("default")))
#else
#define SVX_API
#endif

/* error codes */
#define SVX_OK 0
#define SVX_ERR_CUDA 1        /* a CUDA runtime call failed */
#define SVX_ERR_ARG 2         /* invalid argument (shape, dim, null pointer)  */
#define SVX_ERR_UNSUPPORTED 3 /* valid in the reference but outside this build's limits */

/* cost arithmetic */
#define SVX_COST_EXACT 0
#define SVX_COST_FAST 1
#define SVX_COST_TC 2         /* FAST + the coarsest-level cost matrix on tcgen05 tensor cores
                                 (3xTF32, TMA-fed); needs SvxDenseJob.tmap0/tmap1              */

/* per-job device status bits (status_d) */
#define SVX_ST_OK 0
#define SVX_ST_LEFT_BAND 1   /* traceback left the band (reference: IndexError / wrap-around)  */
#define SVX_ST_NO_BACKPTR 2  /* reached a node whose backpointer is the -42 sentinel          */
#define SVX_ST_OVERFLOW 4    /* output record capacity exceeded                               */

#define SVX_BP_NONE 255      /* uint8 backpointer for the reference's (-42,-42) sentinel      */
#define SVX_MAX_TYPES 126    /* alignment types per job incl. the two deletion types          */

/* ------------------------------------------------------------------------------------------------
 * Row blocks: a (K*n, dim) fp32 matrix of embedding rows.
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxRows {
    float *ptr;      /* (nrows, dim) row-major, 16-byte aligned                        */
    int64_t nrows;
} SvxRows;

/* Replaces dp_utils.make_norm1 (dp_utils.py:32-40): in place v /= sqrt(sum(v*v)) + 1e-5, fp32,
 * with numpy's pairwise summation order (bit-identical to numpy >= 2.0).  dim in
 * {128,256,512,1024}. */
SVX_API int svx_normalize_rows(const SvxRows *jobs_d, const SvxRows *jobs_h, int njobs, int dim, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Downsample: replaces dp_utils.downsample_vectors (dp_utils.py:362-378).
 * out[o,j,:] = in[o,2j,:] + in[o,2j+1,:]; minus the per-overlap mean row (sequential fp32 row
 * accumulation = np.mean(axis=0)); rows normalised as above.
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxDownJob {
    const float *in;   /* (k, n, dim)                                   */
    float *out;        /* (k, n/2, dim)                                 */
    float *mean;       /* (k, dim) scratch: receives the mean rows      */
    int32_t k, n;
} SvxDownJob;
SVX_API int svx_downsample(const SvxDownJob *jobs_d, const SvxDownJob *jobs_h, int njobs, int dim, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Sample norms: replaces dp_utils.compute_norms (dp_utils.py:326-359).
 * norms[o,i] = 1 - vecs[o,i,:] . mean_s(samples_s), samples = rows `idx[o']` of overlap o' of the
 * OTHER side (host RNG draws, np.random.choice order of :346).  The reference computes the same
 * quantity with an sgemm followed by an fp32 mean (order not reproducible across CPUs); this
 * GEMV form accumulates in fp64 and agrees to <= 2 ulp (tests state the tolerance).
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxNormJob {
    const float *vecs;     /* (k, n, dim) rows to be normed                                 */
    const float *other;    /* (ko, no, dim) side the samples are drawn from                 */
    const int32_t *idx;    /* (ko, per) sampled row ids in [0, no)                          */
    double *mbar;          /* (dim) scratch: mean sample vector                             */
    float *norms;          /* (k, n) output                                                 */
    int32_t k, n, ko, no, per;
} SvxNormJob;
SVX_API int svx_sample_norms(const SvxNormJob *jobs_d, const SvxNormJob *jobs_h, int njobs, int dim, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Overlap tensor on the device: replaces the copy loop of make_doc_embedding
 * (utils/embedding_utils.py:135-203).  out[j, e, :] = rows[table[j, e]] widened to fp32, a zero row
 * for table < 0 (PAD, unknown key, ignored concatenation, position before the document start) and
 * for source rows that contain a NaN (the reference zeroes those, :196-200).  `rows` is the content
 * of the .embed file in its on-disk dtype (fp16 for SpeechLASER / SONAR dumps), so the host->device
 * transfer is the file, not the K-fold expanded fp32 tensor.
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxGatherJob {
    const void *rows;        /* (nrows, dim) fp16 or fp32                                  */
    const int32_t *table;    /* (k, n) source row of every (overlap, position), -1 = zeros;
                                NULL = identity (rows is already (k, n, dim): widen + NaN scrub) */
    float *out;              /* (k, n, dim) fp32                                           */
    int32_t *nan_rows;       /* (1) or NULL: number of output rows zeroed because of NaNs  */
    int32_t k, n, nrows, is_fp16;
} SvxGatherJob;
SVX_API int svx_gather_doc_embedding(const SvxGatherJob *jobs_d, const SvxGatherJob *jobs_h, int njobs, int dim, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused level prologue: everything dp_utils.vecalign does to one side of one level before the
 * costs (dp_utils.py:396-397, 416-418, 423-444), in the fewest passes over HBM:
 *   (1) mean   [levels >= 1]  per-overlap mean row of the un-centred pair sums (np.mean axis=0 order)
 *   (2) mbar   mean sample vector of the OTHER side, its sampled rows centred + unit-normalised on
 *              the fly (they are still raw when this runs)
 *   (3) finish one pass over the rows, a warp per row pair: subtract the mean row, unit-normalise in
 *              place (make_norm1 arithmetic), norms = 1 - row.mbar, and the pair sums
 *              row[2j] + row[2j+1] that are the next coarser level's input.
 * Per-row arithmetic is identical to svx_normalize_rows / svx_downsample / svx_sample_norms; only
 * the number of passes changes.  Both sides of a pair must be in the same call (step 2 of every job
 * runs before step 3 of any job).
 * ---------------------------------------------------------------------------------------------- */
/* Optional source of a level's raw rows: the (k, n, dim) tensor that svx_gather_doc_embedding would have written -
 * rows[table[o, i]] widened to fp32, a zero row for table < 0, a row index outside [0, nrows) or a source row holding a
 * NaN (embedding_utils.py:135-203) - is read THROUGH the table by the prologue's level-0 pass instead of being written
 * to HBM first and read back (for fp16 .embed rows: 2 bytes read per element instead of 2 read + 4 written + 4 read). */
typedef struct SvxRowSource {
    const void *rows;        /* (nrows, dim) fp16 or fp32; NULL = no source: the raw rows are in SvxLevelJob.vecs */
    const int32_t *table;    /* (k, n) source row of every (overlap, position), -1 = zeros; NULL = identity       */
    int32_t *nan_rows;       /* (1) or NULL: += number of (overlap, position) rows zeroed because of NaNs          */
    int32_t nrows, is_fp16;
} SvxRowSource;

typedef struct SvxLevelJob {
    float *vecs;             /* (k, n, dim) raw rows (level 0) or un-centred pair sums; finished in place */
    float *mean;             /* (k, dim) scratch for the mean rows, NULL at level 0 (nothing to subtract) */
    float *next;             /* (k, n/2, dim) pair sums for the next coarser level, or NULL               */
    const float *other;      /* (ko, no, dim) the other side, same level, same (unfinished) state         */
    const float *other_mean; /* (ko, dim) the other side's `mean` (written by step 1 of its job) or NULL  */
    const int32_t *idx;      /* (ko, per) sampled rows of `other`; NULL: leave `norms` untouched          */
    double *mbar;            /* (dim + 1024) scratch: mean sample vector, then 2048 fp32 denominators     */
    float *norms;            /* (k, n) output, or NULL                                                    */
    int32_t k, n, ko, no, per;
    int32_t keep;            /* overlaps [0, keep) get their finished rows and norms stored; the others
                                are only normalised on the fly for the pair sums (levels >= 1 align
                                1-1 only: overlap 0 is all the later kernels read).  keep = k stores all. */
    SvxRowSource src;        /* src.rows != NULL (mean must be NULL): the raw rows come from here and the */
    SvxRowSource osrc;       /*   finished rows go to `vecs`; osrc likewise replaces `other` (its         */
                             /*   nan_rows is not used: the other side's own job counts them)             */
} SvxLevelJob;
SVX_API int svx_level_prologue(const SvxLevelJob *jobs_d, const SvxLevelJob *jobs_h, int njobs, int dim, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Sampled pair scores + deletion penalty: replaces dp_core.score_path (dp_core.pyx:143-161) and
 * dp_utils.DeletionKnob / make_del_knob (dp_utils.py:43-79, 278-323).
 * scores[i] = 2(1 - e[xi].f[yi]) / (ne[xi] + nf[yi])  (fp32 denominator, no 1e-6).
 * If xi == NULL the full ne x nf grid is scored row-major (the reference's no-RNG branch).
 * svx_del_knob reproduces numpy's histogram(1000 bins, density) / cumsum / searchsorted / interp
 * arithmetic bit-for-bit and writes del_penalty (fp64).
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxScoreJob {
    const float *e;        /* (ne, dim) overlap-0 rows of side 0                */
    const float *f;        /* (nf, dim) overlap-0 rows of side 1                */
    const float *norm_e;   /* (ne)                                              */
    const float *norm_f;   /* (nf)                                              */
    const int32_t *xi;     /* (nsamp) or NULL for the full grid                 */
    const int32_t *yi;     /* (nsamp) or NULL                                   */
    float *scores;         /* (nsamp) output                                    */
    double *del_penalty;   /* (1) output of svx_del_knob                        */
    int32_t *perm;         /* (nsamp) scratch or NULL: svx_score_pairs orders the samples by
                              xi here (device counting sort) so that warps share x rows    */
    const float *dots;     /* (ne, nf) or NULL: e.f already computed in the reference's order
                              (SvxDenseJob.dots of the same level); then only gathered      */
    int32_t ne, nf, nsamp;
} SvxScoreJob;
SVX_API int svx_score_pairs(const SvxScoreJob *jobs_d, const SvxScoreJob *jobs_h, int njobs, int dim, int mode,
                    void *stream);
SVX_API int svx_del_knob(const SvxScoreJob *jobs_d, const SvxScoreJob *jobs_h, int njobs, double frac, void *stream);
/* Host twin of svx_del_knob (same arithmetic compiled for the CPU; used by the not-gpu tests). */
SVX_API int svx_host_del_knob(const float *scores, int n, double frac, double *del_penalty);

/* ------------------------------------------------------------------------------------------------
 * Coarsest level: dense costs + 3-way DP + traceback + search path of the next finer level.
 * Replaces dp_core.make_dense_costs (dp_core.pyx:36-77), dp_core.dense_dp (:79-141),
 * dp_utils.dense_traceback (dp_utils.py:146-174) and, for the path, upsample_alignment /
 * extend_alignments / alignment_to_search_path / append_slant (dp_utils.py:177-275).
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxDenseJob {
    const float *v0;           /* (s0, dim) overlap-0 rows                               */
    const float *v1;           /* (s1, dim)                                              */
    const float *n0;           /* (s0)                                                   */
    const float *n1;           /* (s1)                                                   */
    float *costs;              /* (s0, s1) output of svx_dense_costs                     */
    float *dots;               /* (s0, s1) or NULL: the raw dot products v0[x].v1[y]     */
    const void *tmap0;         /* SVX_COST_TC only: device copies (64-byte aligned) of   */
    const void *tmap1;         /*   the TMA descriptor PAIRS {v0, lo0} / {v1, lo1}, 256 B */
                               /*   each (svx_dense_tmaps_encode)                         */
    float *lo0;                /* SVX_COST_TC only: (s0, dim) / (s1, dim) scratch for the */
    float *lo1;                /*   3xTF32 residual planes, written by svx_dense_costs    */
    const double *del_penalty; /* (1); narrowed to fp32 as dense_dp(float pen) does      */
    uint8_t *bp;               /* (s0+1, s1+1) backpointers 0/1/2, 4 at the origin       */
    double *csum;              /* (s0+1, s1+1) or NULL (debug/parity only)               */
    int32_t *ypath;            /* (path_len) y of the search path of the target level    */
    int32_t *status_d;         /* (1)                                                    */
    int32_t s0, s1;
    int32_t t0, t1;            /* sizes of the target (finer) level                      */
    int32_t upsample;          /* 1: target is the next finer level (x2 + extension);    */
                               /* 0: same level (max_depth == 0)                         */
    int32_t path_len;          /* A of the target level (svx_path_len)                   */
} SvxDenseJob;
SVX_API int svx_dense_costs(const SvxDenseJob *jobs_d, const SvxDenseJob *jobs_h, int njobs, int dim, int mode,
                    void *stream);
SVX_API int svx_dense_dp(const SvxDenseJob *jobs_d, const SvxDenseJob *jobs_h, int njobs, void *stream);
/* Host only: encodes the four CUtensorMap descriptors (128 B each; rows x dim fp32, 128-row x 32-float
 * boxes, 128-byte swizzle) of every job's v0, lo0, v1, lo1 into out_host[4*j .. 4*j+3] (lo0 / lo1 must be
 * set).  The caller copies them to the device and points tmap0 at the first pair and tmap1 at the second
 * before calling svx_dense_costs(SVX_COST_TC), which first fills lo0 / lo1 (the 3xTF32 residual planes). */
SVX_API int svx_dense_tmaps_encode(const SvxDenseJob *jobs_h, int njobs, int dim, void *out_host);

/* Length of the search path built from a coarse alignment of (c0,c1) segments for a target level
 * of (t0,t1) segments (upsample=1), or from a same-size dense alignment (upsample=0). */
SVX_API int svx_path_len(int c0, int c1, int t0, int t1, int upsample);

/* ------------------------------------------------------------------------------------------------
 * Banded levels: multi-type costs inside the band, banded DP, traceback.
 * Replaces dp_core.make_sparse_costs (dp_core.pyx:165-267), dp_core.sparse_dp (:269-404),
 * dp_utils.sparse_traceback + process_scores (dp_utils.py:89-143) and the path glue for the
 * next finer level.
 *
 * Layouts (device): costs (A, T, B) fp32 — anti-diagonal major so that one DP step reads one
 * contiguous T*B block (the reference's (T, A, B) is produced by the Python debug shim);
 * bp (A+2, B) uint8 = index into types ++ [(0,1),(1,0)], SVX_BP_NONE for (-42,-42);
 * csum (A+2, B) fp64.
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxAlignRec {   /* one alignment of the traceback                          */
    int32_t x_end, y_end;      /* exclusive ends: x = [x_end-nx, x_end), y likewise       */
    int32_t nx, ny;
    double score;              /* process_scores() value                                  */
} SvxAlignRec;

typedef struct SvxBandJob {
    const float *v0;           /* (k0, s0, dim)                                           */
    const float *v1;           /* (k1, s1, dim)                                           */
    const float *n0;           /* (k0, s0)                                                */
    const float *n1;           /* (k1, s1)                                                */
    const int32_t *ypath;      /* (a_len) search path y per anti-diagonal (x = a - y)     */
    float *costs;              /* (a_len, ntypes, band)                                   */
    const double *del_penalty; /* (1)                                                     */
    uint8_t *bp;               /* (a_len+2, band)                                         */
    double *csum;              /* (a_len+2, band)                                         */
    SvxAlignRec *recs;         /* (rec_cap) alignments, written back-to-front             */
    int32_t *nrecs;            /* (1) number of records: valid = recs[rec_cap-n, rec_cap) */
    int32_t *next_ypath;       /* (next_len) or NULL at level 0                           */
    int32_t *status_d;         /* (1)                                                     */
    int32_t s0, s1, k0, k1;
    int32_t a_len, band, width_over2, ntypes;
    int32_t rec_cap;
    int32_t t0, t1, next_len;  /* finer level sizes / path length (unused at level 0)     */
    int8_t xo[SVX_MAX_TYPES];  /* alignment types (x,y), reference order; the DP appends  */
    int8_t yo[SVX_MAX_TYPES];  /*   (0,1) and (1,0) itself                                */
    int16_t amax;              /* max over types of xo+yo                                 */
} SvxBandJob;
SVX_API int svx_banded_costs(const SvxBandJob *jobs_d, const SvxBandJob *jobs_h, int njobs, int dim, int mode,
                     void *stream);
SVX_API int svx_banded_dp(const SvxBandJob *jobs_d, const SvxBandJob *jobs_h, int njobs, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Host twins of the integer/fp64 DP logic (identical source compiled for the CPU) — they let
 * the not-gpu tests check the cell update, traceback and path builder against the oracle.
 * ---------------------------------------------------------------------------------------------- */
SVX_API int svx_host_banded_dp(const SvxBandJob *job_host_pointers);
SVX_API int svx_host_dense_dp(const SvxDenseJob *job_host_pointers);

/* ------------------------------------------------------------------------------------------------
 * Host-side replay of the reference's RNG draws.  dp_utils.py:301-302,346 call
 * np.random.choice(range(n), size=k) on the global legacy RandomState, which consumes the MT19937
 * stream exactly like randint(0, n, k) (masked rejection on 32-bit outputs).  These two functions
 * generate the same numbers in C: `stream` continues one state (np.random.get_state() key/pos, to be
 * put back with set_state), `seeded` runs independent streams seeded like np.random.seed(seed), in
 * parallel.  dst[c] receives count[c] int32 values of call c.
 * ---------------------------------------------------------------------------------------------- */
SVX_API int svx_host_randint_stream(uint32_t *key624, int32_t *pos, int ncalls, const int32_t *high,
                                    const int64_t *count, int32_t *const *dst);
SVX_API int svx_host_randint_seeded(int nstreams, const uint32_t *seeds, const int64_t *call_begin, const int32_t *high,
                                    const int64_t *count, int32_t *const *dst, int nthreads);

/* Copies nbytes (rounded up to 16) from PINNED host memory to the device with a kernel that reads the
 * host buffer over PCIe (UVA), i.e. without the DMA queue that bulk cudaMemcpyAsync traffic occupies.
 * For the few-MB descriptor block of a batch.  Both pointers 16-byte aligned. */
SVX_API int svx_upload_pinned(void *dst_d, const void *src_pinned_h, long long nbytes, void *stream);

/* Host memcpy on `nthreads` threads (page-aligned slices).  Pageable numpy inputs - what the reference's callers
 * hand to dp_utils.vecalign (dp_utils.py:381) - are staged through pinned buffers with it: one thread of the
 * driver's own pageable path reaches ~10 GB/s, the PCIe link takes 55. */
SVX_API int svx_host_memcpy(void *dst, const void *src, long long nbytes, int nthreads);

/* ------------------------------------------------------------------------------------------------
 * Whole path: one batch of document pairs from overlap tensors to alignment records.
 * Replaces the control flow of dp_utils.vecalign (dp_utils.py:381-537) - level sizes (:403-408), the
 * sampling plan and RNG call order (:288-302, :339-346, :423-455), the coarse-to-fine loop (:457-535) -
 * for MANY pairs per call.  The caller owns two buffers, sized by svx_workspace_bytes / svx_plan_info:
 * the device ARENA (every intermediate and the results) and a host STAGING block (descriptors + RNG
 * draws; pinned memory lets the upload bypass the DMA queue).  A plan can be re-bound and re-run.
 * ---------------------------------------------------------------------------------------------- */
typedef struct SvxAlignParams {
    int32_t k0, k1, dim;           /* overlaps per side, embedding dimension (vecs are (k, n, dim) fp32)     */
    int32_t ntypes;                /* final_alignment_types (vecalign.py:154-171), reference order           */
    int8_t xo[SVX_MAX_TYPES], yo[SVX_MAX_TYPES];
    double del_percentile_frac;
    int32_t width_over2;           /* raised to 3 like dp_utils.py:391-393                                   */
    int32_t max_size_full_dp, costs_sample_size, num_samps_for_norm;
    int32_t cost_mode;             /* SVX_COST_*                                                             */
    int32_t keep_all;              /* 1: keep every overlap of every level + the dense csum (parity/debug)   */
    int32_t unfused_prologue;      /* 1: the three separate prologue launchers (A/B)                         */
    int32_t skip_norms0, skip_norms1; /* 1: the caller writes level-0 norms itself (norms0= / norms1=, dp_utils.py:428-444) */
} SvxAlignParams;

typedef struct SvxPlan SvxPlan;    /* opaque */

typedef struct SvxPlanInfo {
    int32_t npairs, nrecords;      /* records = (pair, level) */
    int32_t max_depth, band, width_over2, per0, per1, fused_prologue, nlaunchers;
    int64_t ndraw_calls;
    int64_t arena_bytes;           /* device workspace                                          */
    int64_t host_bytes;            /* staging block = the host-initialised prefix of the arena  */
    int64_t result_offset;         /* [result_offset, +result_bytes): the level-0 alignment records */
    int64_t result_bytes;
    int64_t counts_offset;         /* [counts_offset, arena_bytes): record counts and status words  */
    double fallback_del_penalty;   /* dp_utils.py:315-321                                       */
} SvxPlanInfo;

/* per-record byte offsets into the arena (svx_plan_array(SVX_PA_OFFSETS + key)) */
enum {
    SVX_PO_IDX0, SVX_PO_IDX1, SVX_PO_XI, SVX_PO_YI, SVX_PO_DELPEN, SVX_PO_TMAPS,
    SVX_PO_NORMS0, SVX_PO_NORMS1, SVX_PO_VEC0, SVX_PO_VEC1, SVX_PO_MEAN0, SVX_PO_MEAN1, SVX_PO_MBAR0, SVX_PO_MBAR1,
    SVX_PO_SCORES, SVX_PO_PERM, SVX_PO_DCOST, SVX_PO_DDOTS, SVX_PO_DBP, SVX_PO_DCSUM, SVX_PO_YPATH, SVX_PO_BCOST,
    SVX_PO_BBP, SVX_PO_BCSUM, SVX_PO_RECS, SVX_PO_NRECS, SVX_PO_STATUS, SVX_PO_DLO0, SVX_PO_DLO1, SVX_PO_COUNT
};
/* int64 arrays of a plan: per pair (FIRST, NLEV, DEPTH, TOP_REC, TGT_REC), per record, per RNG call */
enum {
    SVX_PA_FIRST, SVX_PA_NLEV, SVX_PA_DEPTH, SVX_PA_REC_PAIR, SVX_PA_REC_LEVEL, SVX_PA_RS0, SVX_PA_RS1, SVX_PA_A, SVX_PA_T,
    SVX_PA_BANDED, SVX_PA_REC_CAP, SVX_PA_NSAMP, SVX_PA_HAS_DRAW, SVX_PA_TOP_REC, SVX_PA_TGT_REC,
    SVX_PA_DRAW_PAIR, SVX_PA_DRAW_HIGH, SVX_PA_DRAW_COUNT, SVX_PA_DRAW_OFF, SVX_PA_DRAW_BEGIN,
    SVX_PA_OFFSETS = 64
};

SVX_API int svx_plan_create(const SvxAlignParams *params, int npairs, const int32_t *n0, const int32_t *n1, SvxPlan **out);
SVX_API void svx_plan_destroy(SvxPlan *plan);
SVX_API int svx_plan_info(const SvxPlan *plan, SvxPlanInfo *info);
SVX_API int svx_plan_array(const SvxPlan *plan, int which, const int64_t **ptr, int64_t *count);   /* borrowed */
/* Optional, before svx_plan_bind: per-pair sources of the level-0 rows (arrays of npairs entries, copied; NULL, NULL
 * clears them).  v0_d[p] / v1_d[p] of svx_plan_bind are then pure outputs (the normalised fp32 rows).  Needs the fused
 * prologue (SvxPlanInfo.fused_prologue); SVX_ERR_UNSUPPORTED otherwise - gather with svx_gather_doc_embedding then. */
SVX_API int svx_plan_set_sources(SvxPlan *plan, const SvxRowSource *src0, const SvxRowSource *src1);
/* Writes default penalties, TMA descriptors and every job descriptor for (arena_d, v0_d[p], v1_d[p]) into stage_h. */
SVX_API int svx_plan_bind(SvxPlan *plan, void *arena_d, void *stage_h, const void *const *v0_d, const void *const *v1_d);
/* The reference's np.random draws (np.random.choice(range(n), size=k), dp_utils.py:301-302,346) written into the
 * staging block in the reference's call order: one np.random.seed(seeds[p]) stream per pair, or the caller's global
 * stream continued from np.random.get_state() (key, pos) and handed back for set_state. */
SVX_API int svx_plan_draw_seeded(SvxPlan *plan, const uint32_t *seeds, int nthreads);
SVX_API int svx_plan_draw_stream(SvxPlan *plan, uint32_t *key624, int32_t *pos);
/* staging block -> arena; norms preset to 1.0 (dp_utils.py:356-357), counts and status cleared.  Asynchronous. */
SVX_API int svx_plan_upload(SvxPlan *plan, int stage_is_pinned, void *stream);
/* The same reset from a device copy of the staging block (host_bytes long): for callers that run several plans in one
 * arena and restore each plan's descriptors and draws before its run without touching PCIe. */
SVX_API int svx_plan_restore(SvxPlan *plan, const void *stage_copy_d, void *stream);
SVX_API int svx_plan_launcher_name(const SvxPlan *plan, int launcher, char *buf, int cap);
/* Enqueues launcher `launcher` (-1: the whole chain, in order) for pairs [pair_lo, pair_hi).  Asynchronous. */
SVX_API int svx_plan_enqueue(const SvxPlan *plan, int launcher, int pair_lo, int pair_hi, void *stream);
/* Level-0 alignment records in document order, level-0 deletion penalty and the OR of all status words of every
 * pair, copied to host arrays; pair p owns recs_out[rec_begin[p], rec_begin[p+1]).  Synchronises the stream. */
SVX_API int svx_plan_fetch(const SvxPlan *plan, SvxAlignRec *recs_out, const int64_t *rec_begin, int32_t *nrecs_out,
                           double *del_penalty_out, int32_t *status_out, void *stream);

/* The two calls of a host without a planner of its own: size the buffers, then align a batch (plan, seeded draws,
 * upload, launch chain, fetch; synchronous).  v0_d[p] / v1_d[p]: device (k, n[p], dim) fp32, normalised in place
 * like dp_utils.py:396-397. */
SVX_API int svx_workspace_bytes(const SvxAlignParams *params, int npairs, const int32_t *n0, const int32_t *n1,
                                int64_t *arena_bytes, int64_t *stage_bytes);
SVX_API int svx_align_batch(const SvxAlignParams *params, int npairs, const int32_t *n0, const int32_t *n1,
                            const void *const *v0_d, const void *const *v1_d, const uint32_t *seeds,
                            void *arena_d, int64_t arena_bytes, void *stage_h, int64_t stage_bytes, int stage_is_pinned,
                            SvxAlignRec *recs_out, const int64_t *rec_begin, int32_t *nrecs_out, double *del_penalty_out,
                            int32_t *status_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Margin scoring of aligned segment pairs: replaces compute_sim_with_nonflat_idx
 * (svecalign/postprocess/score_align.py:124-161) with an exact flat k-nearest-neighbour search on the
 * tensor cores (tcgen05 fp16 x fp16 -> fp32, TMA-fed; the reference's faiss-gpu flat index stores fp16):
 *   score_i = a_i / b_i (margin 0, "ratio") or a_i - b_i (1, "distance"),  a_i = <x_i, y_i> on L2-normalised rows,
 *   b_i = mean of (2 - avg_k |x_i - y|^2) / 2 over the k nearest y in ybase and the same for y_i in xbase.
 * xbase_d / ybase_d = NULL: the pairs' own rows are the searched collections (the shipped example).
 * Rows are fp32 or fp16 (is_fp16), dim a multiple of 64, 1 <= k <= 16.  Asynchronous on `stream`.
 * ---------------------------------------------------------------------------------------------- */
SVX_API int svx_margin_workspace_bytes(int n, int n_xbase, int n_ybase, int dim, int64_t *bytes);
SVX_API int svx_margin_scores(const void *x_d, const void *y_d, int n, const void *xbase_d, int n_xbase, const void *ybase_d,
                              int n_ybase, int dim, int is_fp16, int k, int margin, float *scores_d, void *workspace_d,
                              int64_t workspace_bytes, void *stream);

/* Host only: the (max_overlaps, nlines[d]) int32 row tables of make_doc_embedding (utils/embedding_utils.py:106-203,
 * overlap_segments=True as seg_align/align.py:222 passes) for `ndocs` documents, read straight from their segment and
 * concatenation files on `nthreads` threads; ignore_pairs[d] = n_ignore[d] (start, end) pairs (vecalign.py:43-52 files),
 * NULL for none.  tables_out[d] feeds SvxGatherJob.table; nrows_out[d] = rows the concatenation file names. */
SVX_API int svx_host_overlap_tables(int ndocs, const char *const *seg_paths, const char *const *cat_paths,
                                    const int32_t *const *ignore_pairs, const int32_t *n_ignore, int max_overlaps,
                                    int32_t *const *tables_out, const int32_t *nlines, int32_t *nrows_out, int nthreads);

/* misc */
SVX_API int svx_version(void);
SVX_API const char *svx_last_error_string(void);
SVX_API long long svx_launch_count(int reset); /* kernels launched since the last reset */
SVX_API int svx_sizeof_job(int which); /* 0 Rows,1 Down,2 Norm,3 Score,4 Dense,5 Band,6 AlignRec,7 Level,8 Gather,9 AlignParams,10 PlanInfo */

#ifdef __cplusplus
}
#endif
#endif /* SVX_H_ */
