// plan.cu — the whole-path entry points of the C ABI: workspace sizing, arena layout, job descriptors and the
// launch chain of one batch of document pairs.  This is the control flow of dp_utils.vecalign
// (svecalign/vecalign/dp_utils.py:381-537) for MANY pairs at once, with every numeric step a launcher of this
// library; the host language above it (Python here, anything with a C FFI elsewhere) only owns the buffers.
//
//   svx_plan_create   sizes of every level of every pair (dp_utils.py:403-408), search-path lengths, sampling plan
//                     (:288-302, :339-346), arena layout (256-byte aligned bump allocation), launch chain
//   svx_plan_bind     descriptors for a concrete arena / input tensors, written into the caller's staging block
//   svx_plan_draw_*   the reference's np.random draws in the reference's call order (global stream continued from
//                     a get_state() snapshot, or one np.random.seed(seed) stream per pair)
//   svx_plan_upload   staging block -> arena, norms preset to 1.0 (dp_utils.py:356-357), status words cleared
//   svx_plan_enqueue  the launch chain (or one launcher of it) for a range of pairs
//   svx_plan_fetch    alignment records, deletion penalties and status words back to the host
//   svx_workspace_bytes / svx_align_batch   the two calls a non-Python host needs
#include <string.h>
#include <math.h>
#include <algorithm>
#include <string>
#include <vector>
#include "svx_common.cuh"

namespace {

constexpr int64_t kAlign = 256;
constexpr int kMaxFusedSamples = 2048;       // svx_level_prologue keeps the sampled rows' denominators in its scratch

enum LauncherKind { L_LEVEL, L_ROWS, L_DOWN, L_NORM, L_DENSE_COSTS, L_SCORE, L_KNOB, L_DENSE_DP, L_BAND_COSTS, L_BAND_DP };

struct JobArr {
    std::vector<uint8_t> host;     // packed structs (the launchers' jobs_h)
    std::vector<int32_t> pair;     // pair of every job, non-decreasing: a pair range is a job range
    size_t item = 0;
    int64_t off = 0;               // byte offset of the device copy inside the arena (host-initialised region)
    int count() const { return item ? (int)(host.size() / item) : 0; }
};

struct Launcher {
    LauncherKind kind;
    int arr;                       // index into SvxPlanImpl::arrays
    std::string name;
};

}  // namespace

struct SvxPlan {
    SvxAlignParams prm;
    int P = 0, R = 0, lmax = 0, band = 0, w = 0, per0 = 0, per1 = 0;
    bool fused = true;
    std::vector<int64_t> n0, n1, depth, nlev, first;
    std::vector<int64_t> rec_pair, rec_level, rs0, rs1, A, T, banded, rec_cap, nsamp, has_draw, is_top, top_rec, tgt_rec;
    std::vector<int64_t> off[SVX_PO_COUNT];
    int64_t host_bytes = 0, arena_bytes = 0, norms_lo = 0, norms_hi = 0, zero_lo = 0, jobs_off = 0, jobs_cap = 0;
    int64_t result_lo = 0, result_hi = 0;          // level-0 alignment records
    std::vector<int64_t> draw_pair, draw_high, draw_count, draw_off, draw_begin;
    std::vector<JobArr> arrays;
    std::vector<Launcher> chain;
    // binding
    char *arena = nullptr;
    unsigned char *stage = nullptr;
    bool bound = false;
    double fallback_pen = 0.0;
    std::vector<SvxRowSource> src0, src1;          // optional sources of the level-0 rows (svx_plan_set_sources)
};

namespace {

int64_t take(int64_t &top, std::vector<int64_t> &offs, const std::vector<int64_t> &nbytes)
{
    offs.resize(nbytes.size());
    for (size_t i = 0; i < nbytes.size(); ++i) {
        offs[i] = top;
        top += (nbytes[i] + kAlign - 1) / kAlign * kAlign;
    }
    return top;
}

int64_t path_len(int64_t c0, int64_t c1, int64_t t0, int64_t t1, bool upsample)
{
    // consequence of dp_utils.py:228-258 (extend_alignments) + :261-275 (upsample_alignment)
    if (!upsample) return 1 + c0 + c1;
    const int64_t xmax = c0 > 0 ? 2 * c0 - 1 : 0, ymax = c1 > 0 ? 2 * c1 - 1 : 0;
    const int64_t lenx = std::max<int64_t>(t0 - xmax, 0), leny = std::max<int64_t>(t1 - ymax, 0);
    return 1 + 2 * c0 + lenx + 2 * c1 + leny;
}

template <class J>
J *push_job(JobArr &a, int pair)
{
    a.item = sizeof(J);
    a.host.resize(a.host.size() + sizeof(J));
    a.pair.push_back(pair);
    J *j = reinterpret_cast<J *>(a.host.data() + a.host.size() - sizeof(J));
    memset(j, 0, sizeof(J));
    return j;
}

// dp_utils.py:315-321: with an empty side the knob is built from the samples [0, .5, 1] on [0, 1]
double fallback_penalty(double frac)
{
    const float samp[3] = {0.0f, 0.5f, 1.0f};
    double pen = 0.0;
    svx_host_del_knob(samp, 3, frac, &pen);
    return pen;
}

}  // namespace

extern "C" int svx_plan_create(const SvxAlignParams *prm, int npairs, const int32_t *n0, const int32_t *n1, SvxPlan **out)
{
    SVX_REQUIRE(prm && out && npairs >= 0 && (npairs == 0 || (n0 && n1)), SVX_ERR_ARG, "svx_plan_create: null argument");
    SVX_REQUIRE(prm->ntypes >= 0 && prm->ntypes + 2 <= SVX_MAX_TYPES, SVX_ERR_UNSUPPORTED, "svx_plan_create: too many alignment types for this build");
    int mx = 0, my = 0;
    for (int t = 0; t < prm->ntypes; ++t) {
        SVX_REQUIRE(prm->xo[t] > 0 && prm->yo[t] > 0, SVX_ERR_ARG, "svx_plan_create: alignment type (%d,%d) must be positive (dp_core.pyx:28-30)",
                    prm->xo[t], prm->yo[t]);
        mx = std::max(mx, (int)prm->xo[t]);
        my = std::max(my, (int)prm->yo[t]);
    }
    // dp_core.pyx:204-209 (the reference's message, typo included)
    SVX_REQUIRE(mx <= prm->k0, SVX_ERR_ARG, "%d x overlaps requrested (via alignment_types), but vecs0 only has %d", mx, prm->k0);
    SVX_REQUIRE(my <= prm->k1, SVX_ERR_ARG, "%d y overlaps requrested (via alignment_types), but vecs1 only has %d", my, prm->k1);
    SVX_REQUIRE(prm->k0 >= 0 && prm->k1 >= 0 && prm->dim > 0, SVX_ERR_ARG, "svx_plan_create: bad shape");

    SvxPlan *pl = new SvxPlan();
    pl->prm = *prm;
    const int P = pl->P = npairs;
    const int k0 = prm->k0, k1 = prm->k1;
    const int64_t D = prm->dim;
    pl->w = std::max(3, (int)prm->width_over2);                 // dp_utils.py:391-393
    pl->band = 2 * pl->w;
    const int64_t band = pl->band;
    const int64_t S = prm->costs_sample_size;
    pl->per1 = k1 ? (prm->num_samps_for_norm + k1 - 1) / k1 : 0;   // samples per overlap of side 1 (they norm side 0)
    pl->per0 = k0 ? (prm->num_samps_for_norm + k0 - 1) / k0 : 0;
    if (prm->num_samps_for_norm <= 0) pl->per0 = pl->per1 = 0;
    const int per0 = pl->per0, per1 = pl->per1;
    pl->fused = !prm->unfused_prologue && (int64_t)k1 * per1 <= kMaxFusedSamples && (int64_t)k0 * per0 <= kMaxFusedSamples;
    pl->fallback_pen = fallback_penalty(prm->del_percentile_frac);
    const bool keep = prm->keep_all != 0;
    const bool tc = prm->cost_mode == SVX_COST_TC;

    // ---- levels (dp_utils.py:403-408: halve both sides until s0 * s1 <= max_size_full_dp ** 2) ----------------
    const int64_t lim = (int64_t)prm->max_size_full_dp * prm->max_size_full_dp;
    pl->n0.assign(n0, n0 + P);
    pl->n1.assign(n1, n1 + P);
    pl->depth.resize(P); pl->nlev.resize(P); pl->first.resize(P);
    int64_t R = 0;
    for (int p = 0; p < P; ++p) {
        SVX_REQUIRE(n0[p] >= 0 && n1[p] >= 0, SVX_ERR_ARG, "svx_plan_create: negative document length");
        int64_t a = n0[p], b = n1[p], d = 0;
        while (a * b > lim) { a /= 2; b /= 2; ++d; }
        pl->depth[p] = d; pl->nlev[p] = d + 1; pl->first[p] = R;
        R += d + 1;
        pl->lmax = std::max<int>(pl->lmax, (int)d);
    }
    pl->R = (int)R;
    auto &rp = pl->rec_pair; auto &rl = pl->rec_level; auto &rs0 = pl->rs0; auto &rs1 = pl->rs1;
    rp.resize(R); rl.resize(R); rs0.resize(R); rs1.resize(R);
    pl->A.assign(R, 0); pl->T.assign(R, 1); pl->banded.assign(R, 0); pl->rec_cap.assign(R, 0);
    pl->nsamp.assign(R, 0); pl->has_draw.assign(R, 0); pl->is_top.assign(R, 0);
    pl->top_rec.resize(P); pl->tgt_rec.resize(P);
    for (int p = 0; p < P; ++p)
        for (int l = 0; l <= pl->depth[p]; ++l) {
            const int64_t r = pl->first[p] + l;
            rp[r] = p; rl[r] = l; rs0[r] = pl->n0[p] >> l; rs1[r] = pl->n1[p] >> l;
        }
    for (int64_t r = 0; r < R; ++r) {
        const int p = (int)rp[r];
        const bool top = rl[r] == pl->depth[p], l0 = rl[r] == 0;
        pl->is_top[r] = top;
        const bool banded = !top || pl->depth[p] == 0;       // every level below the top, or a one-level pair's only level
        pl->banded[r] = banded;
        if (banded) {
            if (pl->depth[p] == 0) pl->A[r] = path_len(rs0[r], rs1[r], rs0[r], rs1[r], false);
            else pl->A[r] = path_len(rs0[r + 1], rs1[r + 1], rs0[r], rs1[r], true);
            pl->rec_cap[r] = rs0[r] + rs1[r] + 2;
        }
        pl->T[r] = l0 ? prm->ntypes : 1;
        const int64_t prod = rs0[r] * rs1[r];
        const bool any = rs0[r] > 0 && rs1[r] > 0 && S > 0;
        pl->nsamp[r] = any ? std::min(prod, S) : 0;
        pl->has_draw[r] = any && prod >= S;                     // dp_utils.py:288-302: full grid below the sample size
        if (top) { pl->top_rec[p] = r; pl->tgt_rec[p] = pl->depth[p] > 0 ? r - 1 : r; }
    }

    // ---- arena layout ------------------------------------------------------------------------------------------
    int64_t top = 0;
    std::vector<int64_t> nb(R);
    auto lay = [&](int key, auto fn) {
        for (int64_t r = 0; r < R; ++r) nb[r] = fn(r);
        take(top, pl->off[key], nb);
    };
    auto l0 = [&](int64_t r) { return rl[r] == 0; };
    // host-initialised region first (one upload)
    lay(SVX_PO_IDX0, [&](int64_t) { return (int64_t)k1 * per1 * 4; });
    lay(SVX_PO_IDX1, [&](int64_t) { return (int64_t)k0 * per0 * 4; });
    lay(SVX_PO_XI, [&](int64_t r) { return pl->has_draw[r] ? pl->nsamp[r] * 4 : 0; });
    lay(SVX_PO_YI, [&](int64_t r) { return pl->has_draw[r] ? pl->nsamp[r] * 4 : 0; });
    lay(SVX_PO_DELPEN, [&](int64_t) { return (int64_t)8; });
    lay(SVX_PO_TMAPS, [&](int64_t r) { return (pl->is_top[r] && tc) ? (int64_t)512 : 0; });
    {
        const int64_t nstage = std::max(1, pl->lmax);
        pl->jobs_cap = 2 * P * (int64_t)sizeof(SvxRows) + 2 * R * (int64_t)sizeof(SvxDownJob) + 2 * R * (int64_t)sizeof(SvxNormJob) +
                       2 * R * (int64_t)sizeof(SvxLevelJob) + R * (int64_t)sizeof(SvxScoreJob) + P * (int64_t)sizeof(SvxDenseJob) +
                       R * (int64_t)sizeof(SvxBandJob) + 16 * (8 + 2 * (pl->lmax + 1) + 2 * nstage) + 4096;
        pl->jobs_off = top;
        top += (pl->jobs_cap + kAlign - 1) / kAlign * kAlign;
    }
    pl->host_bytes = top;
    // device-only region
    pl->norms_lo = top;
    lay(SVX_PO_NORMS0, [&](int64_t r) { return (int64_t)k0 * rs0[r] * 4; });
    lay(SVX_PO_NORMS1, [&](int64_t r) { return (int64_t)k1 * rs1[r] * 4; });
    pl->norms_hi = top;
    lay(SVX_PO_VEC0, [&](int64_t r) { return l0(r) ? 0 : (int64_t)k0 * rs0[r] * D * 4; });
    lay(SVX_PO_VEC1, [&](int64_t r) { return l0(r) ? 0 : (int64_t)k1 * rs1[r] * D * 4; });
    lay(SVX_PO_MEAN0, [&](int64_t r) { return l0(r) ? 0 : (int64_t)k0 * D * 4; });
    lay(SVX_PO_MEAN1, [&](int64_t r) { return l0(r) ? 0 : (int64_t)k1 * D * 4; });
    lay(SVX_PO_MBAR0, [&](int64_t) { return (D + 1024) * 8; });     // + the sampled rows' denominators (SvxLevelJob.mbar)
    lay(SVX_PO_MBAR1, [&](int64_t) { return (D + 1024) * 8; });
    lay(SVX_PO_SCORES, [&](int64_t r) { return pl->nsamp[r] * 4; });
    lay(SVX_PO_PERM, [&](int64_t r) { return (pl->has_draw[r] && !pl->is_top[r]) ? pl->nsamp[r] * 4 : 0; });
    lay(SVX_PO_DCOST, [&](int64_t r) { return pl->is_top[r] ? rs0[r] * rs1[r] * 4 : 0; });
    lay(SVX_PO_DDOTS, [&](int64_t r) { return pl->is_top[r] ? rs0[r] * rs1[r] * 4 : 0; });
    lay(SVX_PO_DBP, [&](int64_t r) { return pl->is_top[r] ? (rs0[r] + 1) * (rs1[r] + 1) : 0; });
    lay(SVX_PO_DCSUM, [&](int64_t r) { return (pl->is_top[r] && keep) ? (rs0[r] + 1) * (rs1[r] + 1) * 8 : 0; });
    lay(SVX_PO_YPATH, [&](int64_t r) { return pl->A[r] * 4; });
    lay(SVX_PO_BCOST, [&](int64_t r) { return pl->A[r] * pl->T[r] * band * 4; });
    lay(SVX_PO_BBP, [&](int64_t r) { return pl->banded[r] ? (pl->A[r] + 2) * band : 0; });
    lay(SVX_PO_BCSUM, [&](int64_t r) { return pl->banded[r] ? (pl->A[r] + 2) * band * 8 : 0; });
    lay(SVX_PO_DLO0, [&](int64_t r) { return (pl->is_top[r] && tc) ? rs0[r] * D * 4 : 0; });   // 3xTF32 residual planes
    lay(SVX_PO_DLO1, [&](int64_t r) { return (pl->is_top[r] && tc) ? rs1[r] * D * 4 : 0; });
    // alignment records: the level-0 ones (the results) first and contiguous, so that reading the results back moves
    // only them; the coarser levels' records (consumed on the device by the next level's path builder) follow
    pl->off[SVX_PO_RECS].assign(R, 0);
    pl->result_lo = top;
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t r = 0; r < R; ++r)
            if ((rl[r] == 0) == (pass == 0)) {
                pl->off[SVX_PO_RECS][r] = top;
                top += (pl->rec_cap[r] * (int64_t)sizeof(SvxAlignRec) + kAlign - 1) / kAlign * kAlign;
            }
        if (pass == 0) pl->result_hi = top;
    }
    pl->zero_lo = top;
    lay(SVX_PO_NRECS, [&](int64_t) { return (int64_t)4; });
    lay(SVX_PO_STATUS, [&](int64_t) { return (int64_t)8; });       // [0] banded status, [1] dense status
    pl->arena_bytes = top;

    // ---- RNG call list, the reference's order: per pair, for each level ascending the n0 draws
    // (K1 calls over range(size1), dp_utils.py:346), then the n1 draws (K0 calls over range(size0)); then for each
    // level ascending the knob draws x, y (:301-302) iff size0 * size1 >= costs_sample_size -------------------------
    pl->draw_begin.assign(P + 1, 0);
    for (int p = 0; p < P; ++p) {
        pl->draw_begin[p] = (int64_t)pl->draw_pair.size();
        auto call = [&](int64_t high, int64_t count, int64_t off) {
            pl->draw_pair.push_back(p); pl->draw_high.push_back(high); pl->draw_count.push_back(count); pl->draw_off.push_back(off);
        };
        for (int64_t r = pl->first[p]; r < pl->first[p] + pl->nlev[p]; ++r) {
            const bool lvl0 = rl[r] == 0;
            if (!(lvl0 && prm->skip_norms0) && rs1[r] > 0 && per1 > 0)
                for (int o = 0; o < k1; ++o) call(rs1[r], per1, pl->off[SVX_PO_IDX0][r] + (int64_t)o * per1 * 4);
            if (!(lvl0 && prm->skip_norms1) && rs0[r] > 0 && per0 > 0)
                for (int o = 0; o < k0; ++o) call(rs0[r], per0, pl->off[SVX_PO_IDX1][r] + (int64_t)o * per0 * 4);
        }
        for (int64_t r = pl->first[p]; r < pl->first[p] + pl->nlev[p]; ++r)
            if (pl->has_draw[r]) {
                call(rs0[r], S, pl->off[SVX_PO_XI][r]);
                call(rs1[r], S, pl->off[SVX_PO_YI][r]);
            }
    }
    pl->draw_begin[P] = (int64_t)pl->draw_pair.size();
    *out = pl;
    return SVX_OK;
}

extern "C" void svx_plan_destroy(SvxPlan *pl) { delete pl; }

extern "C" int svx_plan_info(const SvxPlan *pl, SvxPlanInfo *info)
{
    SVX_REQUIRE(pl && info, SVX_ERR_ARG, "svx_plan_info: null argument");
    memset(info, 0, sizeof(*info));
    info->npairs = pl->P; info->nrecords = pl->R; info->max_depth = pl->lmax; info->band = pl->band; info->width_over2 = pl->w;
    info->per0 = pl->per0; info->per1 = pl->per1; info->fused_prologue = pl->fused;
    info->nlaunchers = (int32_t)pl->chain.size();
    info->ndraw_calls = (int64_t)pl->draw_pair.size();
    info->arena_bytes = pl->arena_bytes; info->host_bytes = pl->host_bytes;
    info->result_offset = pl->result_lo;
    info->result_bytes = pl->result_hi - pl->result_lo;
    info->counts_offset = pl->zero_lo;
    info->fallback_del_penalty = pl->fallback_pen;
    return SVX_OK;
}

extern "C" int svx_plan_array(const SvxPlan *pl, int which, const int64_t **ptr, int64_t *count)
{
    SVX_REQUIRE(pl && ptr && count, SVX_ERR_ARG, "svx_plan_array: null argument");
    const std::vector<int64_t> *v = nullptr;
    switch (which) {
        case SVX_PA_FIRST: v = &pl->first; break;
        case SVX_PA_NLEV: v = &pl->nlev; break;
        case SVX_PA_DEPTH: v = &pl->depth; break;
        case SVX_PA_REC_PAIR: v = &pl->rec_pair; break;
        case SVX_PA_REC_LEVEL: v = &pl->rec_level; break;
        case SVX_PA_RS0: v = &pl->rs0; break;
        case SVX_PA_RS1: v = &pl->rs1; break;
        case SVX_PA_A: v = &pl->A; break;
        case SVX_PA_T: v = &pl->T; break;
        case SVX_PA_BANDED: v = &pl->banded; break;
        case SVX_PA_REC_CAP: v = &pl->rec_cap; break;
        case SVX_PA_NSAMP: v = &pl->nsamp; break;
        case SVX_PA_HAS_DRAW: v = &pl->has_draw; break;
        case SVX_PA_TOP_REC: v = &pl->top_rec; break;
        case SVX_PA_TGT_REC: v = &pl->tgt_rec; break;
        case SVX_PA_DRAW_PAIR: v = &pl->draw_pair; break;
        case SVX_PA_DRAW_HIGH: v = &pl->draw_high; break;
        case SVX_PA_DRAW_COUNT: v = &pl->draw_count; break;
        case SVX_PA_DRAW_OFF: v = &pl->draw_off; break;
        case SVX_PA_DRAW_BEGIN: v = &pl->draw_begin; break;
        default:
            if (which >= SVX_PA_OFFSETS && which < SVX_PA_OFFSETS + SVX_PO_COUNT) v = &pl->off[which - SVX_PA_OFFSETS];
    }
    SVX_REQUIRE(v, SVX_ERR_ARG, "svx_plan_array: unknown array %d", which);
    *ptr = v->data();
    *count = (int64_t)v->size();
    return SVX_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Binding: descriptors for a concrete arena and concrete input tensors.
// ------------------------------------------------------------------------------------------------------------------
extern "C" int svx_plan_set_sources(SvxPlan *pl, const SvxRowSource *src0, const SvxRowSource *src1)
{
    SVX_REQUIRE(pl, SVX_ERR_ARG, "svx_plan_set_sources: null plan");
    SVX_REQUIRE((src0 == nullptr) == (src1 == nullptr), SVX_ERR_ARG, "svx_plan_set_sources: give the sources of both sides or of none");
    pl->src0.clear();
    pl->src1.clear();
    pl->bound = false;
    if (!src0) return SVX_OK;
    SVX_REQUIRE(pl->fused, SVX_ERR_UNSUPPORTED,
                "svx_plan_set_sources: this plan runs the unfused prologue (more than %d norm samples per side, or requested); "
                "materialise the rows with svx_gather_doc_embedding instead", kMaxFusedSamples);
    for (int p = 0; p < pl->P; ++p)
        SVX_REQUIRE(src0[p].rows && src1[p].rows && src0[p].nrows >= 0 && src1[p].nrows >= 0, SVX_ERR_ARG,
                    "svx_plan_set_sources: pair %d has no source rows", p);
    pl->src0.assign(src0, src0 + pl->P);
    pl->src1.assign(src1, src1 + pl->P);
    return SVX_OK;
}

extern "C" int svx_plan_bind(SvxPlan *pl, void *arena_d, void *stage_h, const void *const *v0_d, const void *const *v1_d)
{
    SVX_REQUIRE(pl && (pl->P == 0 || (arena_d && stage_h && v0_d && v1_d)), SVX_ERR_ARG, "svx_plan_bind: null argument");
    SVX_REQUIRE(((uintptr_t)arena_d & 15) == 0, SVX_ERR_ARG, "svx_plan_bind: the arena must be 16-byte aligned");
    const SvxAlignParams &prm = pl->prm;
    const int P = pl->P, k0 = prm.k0, k1 = prm.k1, per0 = pl->per0, per1 = pl->per1;
    const int64_t R = pl->R;
    char *base = (char *)arena_d;
    pl->arena = base;
    pl->stage = (unsigned char *)stage_h;
    auto &rp = pl->rec_pair; auto &rl = pl->rec_level; auto &rs0 = pl->rs0; auto &rs1 = pl->rs1;
    auto at = [&](int key, int64_t r) -> char * { return base + pl->off[key][r]; };
    auto vec0 = [&](int64_t r) -> float * { return rl[r] == 0 ? (float *)v0_d[rp[r]] : (float *)at(SVX_PO_VEC0, r); };
    auto vec1 = [&](int64_t r) -> float * { return rl[r] == 0 ? (float *)v1_d[rp[r]] : (float *)at(SVX_PO_VEC1, r); };
    const bool keep = prm.keep_all != 0;

    pl->arrays.clear();
    pl->chain.clear();
    auto new_arr = [&]() { pl->arrays.emplace_back(); return (int)pl->arrays.size() - 1; };
    auto add = [&](LauncherKind k, int arr, const char *name) { pl->chain.push_back(Launcher{k, arr, name}); };

    // which records get sampled norms (dp_utils.py:326-359; ones otherwise, preset by svx_plan_upload)
    std::vector<char> want0(R), want1(R);
    for (int64_t r = 0; r < R; ++r) {
        const bool lvl0 = rl[r] == 0;
        want0[r] = rs1[r] > 0 && per1 > 0 && k1 > 0 && !(lvl0 && prm.skip_norms0) && rs0[r] > 0;
        want1[r] = rs0[r] > 0 && per0 > 0 && k0 > 0 && !(lvl0 && prm.skip_norms1) && rs1[r] > 0;
    }

    if (pl->fused) {
        // fused prologue: one job per (record, side), one launch per level (both sides of a pair in the same call)
        for (int lvl = 0; lvl <= pl->lmax; ++lvl) {
            const int ai = new_arr();
            for (int p = 0; p < P; ++p) {
                if (pl->depth[p] < lvl) continue;
                const int64_t r = pl->first[p] + lvl;
                const bool has_next = lvl < pl->depth[p];
                for (int side = 0; side < 2; ++side) {
                    SvxLevelJob *j = push_job<SvxLevelJob>(pl->arrays[ai], p);
                    j->vecs = side ? vec1(r) : vec0(r);
                    j->other = side ? vec0(r) : vec1(r);
                    if (lvl > 0) {
                        j->mean = (float *)at(side ? SVX_PO_MEAN1 : SVX_PO_MEAN0, r);
                        j->other_mean = (float *)at(side ? SVX_PO_MEAN0 : SVX_PO_MEAN1, r);
                    }
                    j->next = has_next ? (side ? vec1(r + 1) : vec0(r + 1)) : nullptr;
                    const bool want = side ? want1[r] : want0[r];
                    j->idx = want ? (const int32_t *)at(side ? SVX_PO_IDX1 : SVX_PO_IDX0, r) : nullptr;
                    j->norms = want ? (float *)at(side ? SVX_PO_NORMS1 : SVX_PO_NORMS0, r) : nullptr;
                    j->mbar = (double *)at(side ? SVX_PO_MBAR1 : SVX_PO_MBAR0, r);
                    j->k = side ? k1 : k0; j->n = (int32_t)(side ? rs1[r] : rs0[r]);
                    j->ko = side ? k0 : k1; j->no = (int32_t)(side ? rs0[r] : rs1[r]);
                    j->per = side ? per0 : per1;
                    // levels >= 1 align 1-1 only: later kernels read overlap 0 (debug keeps everything)
                    j->keep = (lvl == 0 || keep) ? j->k : std::min(1, j->k);
                    if (lvl == 0 && !pl->src0.empty()) {
                        j->src = side ? pl->src1[p] : pl->src0[p];
                        j->osrc = side ? pl->src0[p] : pl->src1[p];
                        j->osrc.nan_rows = nullptr;
                    }
                }
            }
            add(L_LEVEL, ai, "svx_level_prologue");
        }
    } else {
        const int ar = new_arr();
        for (int p = 0; p < P; ++p) {
            SvxRows *a = push_job<SvxRows>(pl->arrays[ar], p);
            a->ptr = (float *)v0_d[p]; a->nrows = (int64_t)k0 * pl->n0[p];
            SvxRows *b = push_job<SvxRows>(pl->arrays[ar], p);
            b->ptr = (float *)v1_d[p]; b->nrows = (int64_t)k1 * pl->n1[p];
        }
        add(L_ROWS, ar, "svx_normalize_rows");
        for (int lvl = 1; lvl <= pl->lmax; ++lvl) {
            const int ai = new_arr();
            for (int p = 0; p < P; ++p) {
                if (pl->depth[p] < lvl) continue;
                const int64_t r = pl->first[p] + lvl;
                for (int side = 0; side < 2; ++side) {
                    SvxDownJob *j = push_job<SvxDownJob>(pl->arrays[ai], p);
                    j->in = side ? vec1(r - 1) : vec0(r - 1);
                    j->out = side ? vec1(r) : vec0(r);
                    j->mean = (float *)at(side ? SVX_PO_MEAN1 : SVX_PO_MEAN0, r);
                    j->k = side ? k1 : k0;
                    j->n = (int32_t)(side ? rs1[r - 1] : rs0[r - 1]);
                }
            }
            add(L_DOWN, ai, "svx_downsample");
        }
        const int an = new_arr();
        for (int p = 0; p < P; ++p)
            for (int side = 0; side < 2; ++side)
                for (int64_t r = pl->first[p]; r < pl->first[p] + pl->nlev[p]; ++r) {
                    if (!(side ? want1[r] : want0[r])) continue;
                    SvxNormJob *j = push_job<SvxNormJob>(pl->arrays[an], p);
                    j->vecs = side ? vec1(r) : vec0(r);
                    j->other = side ? vec0(r) : vec1(r);
                    j->idx = (const int32_t *)at(side ? SVX_PO_IDX1 : SVX_PO_IDX0, r);
                    j->mbar = (double *)at(side ? SVX_PO_MBAR1 : SVX_PO_MBAR0, r);
                    j->norms = (float *)at(side ? SVX_PO_NORMS1 : SVX_PO_NORMS0, r);
                    j->k = side ? k1 : k0; j->n = (int32_t)(side ? rs1[r] : rs0[r]);
                    j->ko = side ? k0 : k1; j->no = (int32_t)(side ? rs0[r] : rs1[r]);
                    j->per = side ? per0 : per1;
                }
        add(L_NORM, an, "svx_sample_norms");
    }

    // coarsest level: dense costs; sampled scores + knob for every level; dense DP
    const int ad = new_arr();
    for (int p = 0; p < P; ++p) {
        const int64_t t = pl->top_rec[p], g = pl->tgt_rec[p];
        SvxDenseJob *j = push_job<SvxDenseJob>(pl->arrays[ad], p);
        j->v0 = vec0(t); j->v1 = vec1(t);
        j->n0 = (const float *)at(SVX_PO_NORMS0, t); j->n1 = (const float *)at(SVX_PO_NORMS1, t);
        j->costs = (float *)at(SVX_PO_DCOST, t); j->dots = (float *)at(SVX_PO_DDOTS, t);
        j->del_penalty = (const double *)at(SVX_PO_DELPEN, t);
        j->bp = (uint8_t *)at(SVX_PO_DBP, t);
        j->csum = keep ? (double *)at(SVX_PO_DCSUM, t) : nullptr;
        j->ypath = (int32_t *)at(SVX_PO_YPATH, g);
        j->status_d = (int32_t *)(at(SVX_PO_STATUS, t) + 4);
        j->s0 = (int32_t)rs0[t]; j->s1 = (int32_t)rs1[t]; j->t0 = (int32_t)rs0[g]; j->t1 = (int32_t)rs1[g];
        j->upsample = pl->depth[p] > 0; j->path_len = (int32_t)pl->A[g];
        if (prm.cost_mode == SVX_COST_TC) {
            j->tmap0 = at(SVX_PO_TMAPS, t);
            j->tmap1 = at(SVX_PO_TMAPS, t) + 256;
            j->lo0 = (float *)at(SVX_PO_DLO0, t);
            j->lo1 = (float *)at(SVX_PO_DLO1, t);
        }
    }
    const int as = new_arr();
    for (int64_t r = 0; r < R; ++r) {
        if (pl->nsamp[r] <= 0) continue;
        SvxScoreJob *j = push_job<SvxScoreJob>(pl->arrays[as], (int)rp[r]);
        j->e = vec0(r); j->f = vec1(r);
        j->norm_e = (const float *)at(SVX_PO_NORMS0, r); j->norm_f = (const float *)at(SVX_PO_NORMS1, r);
        j->xi = pl->has_draw[r] ? (const int32_t *)at(SVX_PO_XI, r) : nullptr;
        j->yi = pl->has_draw[r] ? (const int32_t *)at(SVX_PO_YI, r) : nullptr;
        j->scores = (float *)at(SVX_PO_SCORES, r);
        j->del_penalty = (double *)at(SVX_PO_DELPEN, r);
        // coarsest level: the dense cost kernel has already produced every dot product of the level
        j->dots = pl->is_top[r] ? (const float *)at(SVX_PO_DDOTS, r) : nullptr;
        j->perm = (pl->has_draw[r] && !pl->is_top[r]) ? (int32_t *)at(SVX_PO_PERM, r) : nullptr;
        j->ne = (int32_t)rs0[r]; j->nf = (int32_t)rs1[r]; j->nsamp = (int32_t)pl->nsamp[r];
    }
    add(L_DENSE_COSTS, ad, "svx_dense_costs");
    add(L_SCORE, as, "svx_score_pairs");
    add(L_KNOB, as, "svx_del_knob");
    add(L_DENSE_DP, ad, "svx_dense_dp");

    // banded stages: stage s (1-based) handles level max(depth, 1) - s of every pair that has it; level-0 jobs
    // (the caller's type list) and coarser-level jobs ((1,1) only) are separate launches
    const int nstage = std::max(1, pl->lmax);
    for (int s = 1; s <= nstage; ++s) {
        int arr_id[2] = {-1, -1};            // [0] coarse, [1] level 0
        for (int want_l0 = 0; want_l0 < 2; ++want_l0) {
            int ai = -1;
            for (int p = 0; p < P; ++p) {
                const int64_t lvl = std::max<int64_t>(pl->depth[p], 1) - s;
                if (lvl < 0 || (lvl == 0) != (want_l0 == 1)) continue;
                const int64_t r = pl->first[p] + lvl;
                if (ai < 0) ai = new_arr();
                SvxBandJob *j = push_job<SvxBandJob>(pl->arrays[ai], p);
                j->v0 = vec0(r); j->v1 = vec1(r);
                j->n0 = (const float *)at(SVX_PO_NORMS0, r); j->n1 = (const float *)at(SVX_PO_NORMS1, r);
                j->ypath = (const int32_t *)at(SVX_PO_YPATH, r);
                j->costs = (float *)at(SVX_PO_BCOST, r);
                j->del_penalty = (const double *)at(SVX_PO_DELPEN, r);
                j->bp = (uint8_t *)at(SVX_PO_BBP, r); j->csum = (double *)at(SVX_PO_BCSUM, r);
                j->recs = (SvxAlignRec *)at(SVX_PO_RECS, r); j->nrecs = (int32_t *)at(SVX_PO_NRECS, r);
                j->status_d = (int32_t *)at(SVX_PO_STATUS, r);
                j->s0 = (int32_t)rs0[r]; j->s1 = (int32_t)rs1[r]; j->k0 = k0; j->k1 = k1;
                j->a_len = (int32_t)pl->A[r]; j->band = pl->band; j->width_over2 = pl->w;
                j->rec_cap = (int32_t)pl->rec_cap[r];
                if (want_l0) {
                    j->ntypes = prm.ntypes;
                    int amax = 2;
                    for (int t = 0; t < prm.ntypes; ++t) {
                        j->xo[t] = prm.xo[t]; j->yo[t] = prm.yo[t];
                        amax = std::max(amax, prm.xo[t] + prm.yo[t]);
                    }
                    j->amax = (int16_t)amax;
                    j->next_ypath = nullptr;
                } else {
                    j->ntypes = 1; j->xo[0] = 1; j->yo[0] = 1; j->amax = 2;
                    j->next_ypath = (int32_t *)at(SVX_PO_YPATH, r - 1);
                    j->t0 = (int32_t)rs0[r - 1]; j->t1 = (int32_t)rs1[r - 1]; j->next_len = (int32_t)pl->A[r - 1];
                }
            }
            arr_id[want_l0] = ai;
        }
        for (int g = 0; g < 2; ++g)
            if (arr_id[g] >= 0) add(L_BAND_COSTS, arr_id[g], g ? "svx_banded_costs_level0" : "svx_banded_costs_coarse");
        for (int g = 0; g < 2; ++g)
            if (arr_id[g] >= 0) add(L_BAND_DP, arr_id[g], g ? "svx_banded_dp_level0" : "svx_banded_dp_coarse");
    }

    // ---- staging block: default penalties, TMA descriptors, descriptor arrays ----------------------------------------
    for (int64_t r = 0; r < R; ++r) memcpy(pl->stage + pl->off[SVX_PO_DELPEN][r], &pl->fallback_pen, 8);
    if (prm.cost_mode == SVX_COST_TC && P) {
        std::vector<uint8_t> blobs((size_t)P * 512);
        const JobArr &da = pl->arrays[ad];
        const int rc = svx_dense_tmaps_encode(reinterpret_cast<const SvxDenseJob *>(da.host.data()), P, prm.dim, blobs.data());
        if (rc != SVX_OK) return rc;
        for (int p = 0; p < P; ++p) memcpy(pl->stage + pl->off[SVX_PO_TMAPS][pl->top_rec[p]], blobs.data() + (size_t)p * 512, 512);
    }
    int64_t cur = pl->jobs_off;
    for (auto &a : pl->arrays) {
        cur = (cur + 15) / 16 * 16;
        SVX_REQUIRE(cur + (int64_t)a.host.size() <= pl->jobs_off + pl->jobs_cap, SVX_ERR_ARG, "svx_plan_bind: descriptor region too small");
        a.off = cur;
        if (!a.host.empty()) memcpy(pl->stage + cur, a.host.data(), a.host.size());
        cur += (int64_t)a.host.size();
    }
    pl->bound = true;
    return SVX_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// RNG draws into the staging block.
// ------------------------------------------------------------------------------------------------------------------
static int draw_args(const SvxPlan *pl, std::vector<int32_t> &high, std::vector<int32_t *> &dst)
{
    const size_t n = pl->draw_pair.size();
    high.resize(n); dst.resize(n);
    for (size_t c = 0; c < n; ++c) {
        high[c] = (int32_t)pl->draw_high[c];
        dst[c] = reinterpret_cast<int32_t *>(pl->stage + pl->draw_off[c]);
    }
    return SVX_OK;
}

extern "C" int svx_plan_draw_seeded(SvxPlan *pl, const uint32_t *seeds, int nthreads)
{
    SVX_REQUIRE(pl && pl->bound && (seeds || pl->P == 0), SVX_ERR_ARG, "svx_plan_draw_seeded: plan not bound / no seeds");
    if (pl->draw_pair.empty()) return SVX_OK;
    std::vector<int32_t> high; std::vector<int32_t *> dst;
    draw_args(pl, high, dst);
    return svx_host_randint_seeded(pl->P, seeds, pl->draw_begin.data(), high.data(), pl->draw_count.data(), dst.data(), nthreads);
}

extern "C" int svx_plan_draw_stream(SvxPlan *pl, uint32_t *key624, int32_t *pos)
{
    SVX_REQUIRE(pl && pl->bound, SVX_ERR_ARG, "svx_plan_draw_stream: plan not bound");
    if (pl->draw_pair.empty()) return SVX_OK;
    std::vector<int32_t> high; std::vector<int32_t *> dst;
    draw_args(pl, high, dst);
    return svx_host_randint_stream(key624, pos, (int)high.size(), high.data(), pl->draw_count.data(), dst.data());
}

// ------------------------------------------------------------------------------------------------------------------
// Device side.
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_fill_f32(float *p, long long n, float v)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

extern "C" int svx_plan_upload(SvxPlan *pl, int stage_is_pinned, void *stream)
{
    SVX_REQUIRE(pl && pl->bound, SVX_ERR_ARG, "svx_plan_upload: plan not bound");
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->R == 0) return SVX_OK;
    // norms default to 1.0 (dp_utils.py:356-357); record counts and status words to 0
    const long long nn = (pl->norms_hi - pl->norms_lo) / 4;
    if (nn > 0) {
        const int blocks = (int)std::min<long long>((nn + 1023) / 1024, 148 * 8);
        k_fill_f32<<<blocks, 256, 0, st>>>(reinterpret_cast<float *>(pl->arena + pl->norms_lo), nn, 1.0f);
        SVX_LAUNCH_CHECK();
    }
    SVX_CUDA_OK(cudaMemsetAsync(pl->arena + pl->zero_lo, 0, (size_t)(pl->arena_bytes - pl->zero_lo), st));
    if (stage_is_pinned) return svx_upload_pinned(pl->arena, pl->stage, pl->host_bytes, stream);
    SVX_CUDA_OK(cudaMemcpyAsync(pl->arena, pl->stage, (size_t)pl->host_bytes, cudaMemcpyHostToDevice, st));
    return SVX_OK;
}

// The same reset from a DEVICE copy of the staging block (a caller that runs several plans in one arena keeps each
// plan's host-initialised prefix in device memory and restores it before every run: no PCIe traffic per run).
extern "C" int svx_plan_restore(SvxPlan *pl, const void *stage_copy_d, void *stream)
{
    SVX_REQUIRE(pl && pl->bound && (stage_copy_d || pl->R == 0), SVX_ERR_ARG, "svx_plan_restore: plan not bound / null copy");
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->R == 0) return SVX_OK;
    const long long nn = (pl->norms_hi - pl->norms_lo) / 4;
    if (nn > 0) {
        const int blocks = (int)std::min<long long>((nn + 1023) / 1024, 148 * 8);
        k_fill_f32<<<blocks, 256, 0, st>>>(reinterpret_cast<float *>(pl->arena + pl->norms_lo), nn, 1.0f);
        SVX_LAUNCH_CHECK();
    }
    SVX_CUDA_OK(cudaMemsetAsync(pl->arena + pl->zero_lo, 0, (size_t)(pl->arena_bytes - pl->zero_lo), st));
    SVX_CUDA_OK(cudaMemcpyAsync(pl->arena, stage_copy_d, (size_t)pl->host_bytes, cudaMemcpyDeviceToDevice, st));
    return SVX_OK;
}

extern "C" int svx_plan_launcher_name(const SvxPlan *pl, int i, char *buf, int cap)
{
    SVX_REQUIRE(pl && buf && i >= 0 && i < (int)pl->chain.size(), SVX_ERR_ARG, "svx_plan_launcher_name: bad index");
    snprintf(buf, (size_t)cap, "%s", pl->chain[i].name.c_str());
    return SVX_OK;
}

static int run_launcher(const SvxPlan *pl, const Launcher &L, int pair_lo, int pair_hi, void *stream)
{
    const JobArr &a = pl->arrays[L.arr];
    const int lo = (int)(std::lower_bound(a.pair.begin(), a.pair.end(), pair_lo) - a.pair.begin());
    const int hi = (int)(std::lower_bound(a.pair.begin(), a.pair.end(), pair_hi) - a.pair.begin());
    if (hi <= lo) return SVX_OK;
    const char *jd = pl->arena + a.off + (size_t)lo * a.item;
    const uint8_t *jh = a.host.data() + (size_t)lo * a.item;
    const int n = hi - lo, D = pl->prm.dim, mode = pl->prm.cost_mode;
    switch (L.kind) {
        case L_LEVEL: return svx_level_prologue((const SvxLevelJob *)jd, (const SvxLevelJob *)jh, n, D, stream);
        case L_ROWS: return svx_normalize_rows((const SvxRows *)jd, (const SvxRows *)jh, n, D, stream);
        case L_DOWN: return svx_downsample((const SvxDownJob *)jd, (const SvxDownJob *)jh, n, D, stream);
        case L_NORM: return svx_sample_norms((const SvxNormJob *)jd, (const SvxNormJob *)jh, n, D, stream);
        case L_DENSE_COSTS: return svx_dense_costs((const SvxDenseJob *)jd, (const SvxDenseJob *)jh, n, D, mode, stream);
        case L_SCORE: return svx_score_pairs((const SvxScoreJob *)jd, (const SvxScoreJob *)jh, n, D, mode, stream);
        case L_KNOB: return svx_del_knob((const SvxScoreJob *)jd, (const SvxScoreJob *)jh, n, pl->prm.del_percentile_frac, stream);
        case L_DENSE_DP: return svx_dense_dp((const SvxDenseJob *)jd, (const SvxDenseJob *)jh, n, stream);
        case L_BAND_COSTS: return svx_banded_costs((const SvxBandJob *)jd, (const SvxBandJob *)jh, n, D, mode, stream);
        case L_BAND_DP: return svx_banded_dp((const SvxBandJob *)jd, (const SvxBandJob *)jh, n, stream);
    }
    return SVX_ERR_ARG;
}

extern "C" int svx_plan_enqueue(const SvxPlan *pl, int launcher, int pair_lo, int pair_hi, void *stream)
{
    SVX_REQUIRE(pl && pl->bound, SVX_ERR_ARG, "svx_plan_enqueue: plan not bound");
    SVX_REQUIRE(launcher >= -1 && launcher < (int)pl->chain.size(), SVX_ERR_ARG, "svx_plan_enqueue: launcher %d out of range", launcher);
    if (pair_hi > pl->P) pair_hi = pl->P;
    if (pair_lo < 0) pair_lo = 0;
    if (launcher >= 0) return run_launcher(pl, pl->chain[launcher], pair_lo, pair_hi, stream);
    for (const Launcher &L : pl->chain) {
        const int rc = run_launcher(pl, L, pair_lo, pair_hi, stream);
        if (rc != SVX_OK) return rc;
    }
    return SVX_OK;
}

// Records of every pair (level 0), first deletion penalty, OR of the status words of all levels.  recs_out receives pair
// p's alignments in document order at recs_out[rec_begin[p] ...]; rec_begin[p+1] - rec_begin[p] is the caller's capacity
// for pair p (n0 + n1 + 2 always suffices).  Synchronises `stream`.
extern "C" int svx_plan_fetch(const SvxPlan *pl, SvxAlignRec *recs_out, const int64_t *rec_begin, int32_t *nrecs_out,
                              double *del_penalty_out, int32_t *status_out, void *stream)
{
    SVX_REQUIRE(pl && pl->bound, SVX_ERR_ARG, "svx_plan_fetch: plan not bound");
    if (pl->P == 0) return SVX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // three copies: the level-0 records, the record counts + status words, the penalties
    const int64_t lo = pl->result_lo, n = pl->result_hi - pl->result_lo;
    const int64_t clo = pl->zero_lo, cn = pl->arena_bytes - pl->zero_lo;
    std::vector<unsigned char> blob((size_t)std::max<int64_t>(n, 1)), cnt((size_t)std::max<int64_t>(cn, 1));
    std::vector<unsigned char> pens((size_t)pl->R * kAlign);
    const int64_t plo = pl->off[SVX_PO_DELPEN][0];
    if (n > 0) SVX_CUDA_OK(cudaMemcpyAsync(blob.data(), pl->arena + lo, (size_t)n, cudaMemcpyDeviceToHost, st));
    SVX_CUDA_OK(cudaMemcpyAsync(cnt.data(), pl->arena + clo, (size_t)cn, cudaMemcpyDeviceToHost, st));
    SVX_CUDA_OK(cudaMemcpyAsync(pens.data(), pl->arena + plo, pens.size(), cudaMemcpyDeviceToHost, st));
    SVX_CUDA_OK(cudaStreamSynchronize(st));
    for (int p = 0; p < pl->P; ++p) {
        const int64_t r0 = pl->first[p];
        const int64_t cap = pl->rec_cap[r0];
        int32_t nrec = 0, status = 0;
        memcpy(&nrec, cnt.data() + pl->off[SVX_PO_NRECS][r0] - clo, 4);
        for (int64_t r = r0; r < r0 + pl->nlev[p]; ++r) {
            int32_t s2[2];
            memcpy(s2, cnt.data() + pl->off[SVX_PO_STATUS][r] - clo, 8);
            status |= s2[0] | s2[1];
        }
        const int64_t valid = std::min<int64_t>(nrec, cap);
        if (recs_out && rec_begin) {
            const int64_t room = rec_begin[p + 1] - rec_begin[p];
            if (valid > room) status |= SVX_ST_OVERFLOW;
            const int64_t ncopy = std::min(valid, room);
            memcpy(recs_out + rec_begin[p], blob.data() + pl->off[SVX_PO_RECS][r0] - lo + (size_t)(cap - valid) * sizeof(SvxAlignRec),
                   (size_t)ncopy * sizeof(SvxAlignRec));
        }
        if (nrecs_out) nrecs_out[p] = nrec;
        if (del_penalty_out) memcpy(del_penalty_out + p, pens.data() + pl->off[SVX_PO_DELPEN][r0] - plo, 8);
        if (status_out) status_out[p] = status;
    }
    return SVX_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// The two calls of a host without a planner of its own.
// ------------------------------------------------------------------------------------------------------------------
extern "C" int svx_workspace_bytes(const SvxAlignParams *prm, int npairs, const int32_t *n0, const int32_t *n1,
                                   int64_t *arena_bytes, int64_t *stage_bytes)
{
    SvxPlan *pl = nullptr;
    const int rc = svx_plan_create(prm, npairs, n0, n1, &pl);
    if (rc != SVX_OK) return rc;
    if (arena_bytes) *arena_bytes = pl->arena_bytes;
    if (stage_bytes) *stage_bytes = pl->host_bytes;
    svx_plan_destroy(pl);
    return SVX_OK;
}

extern "C" int svx_align_batch(const SvxAlignParams *prm, int npairs, const int32_t *n0, const int32_t *n1,
                               const void *const *v0_d, const void *const *v1_d, const uint32_t *seeds,
                               void *arena_d, int64_t arena_bytes, void *stage_h, int64_t stage_bytes, int stage_is_pinned,
                               SvxAlignRec *recs_out, const int64_t *rec_begin, int32_t *nrecs_out, double *del_penalty_out,
                               int32_t *status_out, void *stream)
{
    SvxPlan *pl = nullptr;
    int rc = svx_plan_create(prm, npairs, n0, n1, &pl);
    if (rc != SVX_OK) return rc;
    auto fail = [&](int code) { svx_plan_destroy(pl); return code; };
    if (arena_bytes < pl->arena_bytes || stage_bytes < pl->host_bytes) {
        svx_set_error("svx_align_batch: workspace too small (arena %lld < %lld or staging %lld < %lld bytes)", (long long)arena_bytes,
                      (long long)pl->arena_bytes, (long long)stage_bytes, (long long)pl->host_bytes);
        return fail(SVX_ERR_ARG);
    }
    if (!seeds && !pl->draw_pair.empty()) {
        svx_set_error("svx_align_batch: per-pair seeds are required (the reference draws from np.random; use the svx_plan_* calls to continue a global stream)");
        return fail(SVX_ERR_ARG);
    }
    if ((rc = svx_plan_bind(pl, arena_d, stage_h, v0_d, v1_d)) != SVX_OK) return fail(rc);
    if ((rc = svx_plan_draw_seeded(pl, seeds, 8)) != SVX_OK) return fail(rc);
    if ((rc = svx_plan_upload(pl, stage_is_pinned, stream)) != SVX_OK) return fail(rc);
    if ((rc = svx_plan_enqueue(pl, -1, 0, npairs, stream)) != SVX_OK) return fail(rc);
    rc = svx_plan_fetch(pl, recs_out, rec_begin, nrecs_out, del_penalty_out, status_out, stream);
    return fail(rc);
}
