// capi.cu — error string, version and struct-size entry points of the C ABI (include/svx.h).
#include <stdarg.h>
#include <stdio.h>
#include "../../include/svx.h"

static thread_local char g_err[512] = "";

void svx_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void svx_count_launch(void) { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }

// number of kernels launched by this library since the last reset (bench.py's gpu_launches)
extern "C" long long svx_launch_count(int reset)
{
    long long v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_launches, 0, __ATOMIC_RELAXED);
    return v;
}

extern "C" int svx_version(void) { return SVX_VERSION; }
extern "C" const char *svx_last_error_string(void) { return g_err; }
extern "C" int svx_sizeof_job(int which)
{
    switch (which) {
        case 0: return (int)sizeof(SvxRows);
        case 1: return (int)sizeof(SvxDownJob);
        case 2: return (int)sizeof(SvxNormJob);
        case 3: return (int)sizeof(SvxScoreJob);
        case 4: return (int)sizeof(SvxDenseJob);
        case 5: return (int)sizeof(SvxBandJob);
        case 6: return (int)sizeof(SvxAlignRec);
        case 7: return (int)sizeof(SvxLevelJob);
        case 8: return (int)sizeof(SvxGatherJob);
        case 9: return (int)sizeof(SvxAlignParams);
        case 10: return (int)sizeof(SvxPlanInfo);
        case 11: return (int)sizeof(SvxRowSource);
        default: return -1;
    }
}

// ------------------------------------------------------------------------------------------------
// Host-side replay of the reference's RNG consumption (dp_utils.py:301-302,346 call
// np.random.choice(range(n), size=k), which draws exactly like the legacy
// RandomState.randint(0, n, k)): MT19937 + numpy's masked rejection on 32-bit outputs
// (numpy/random/src/distributions/distributions.c random_bounded_uint64_fill, use_masked = true,
// range < 2^32).  Written here so that a batch's ~10^7 draws cost milliseconds and, with per-pair
// seeds, run on all host cores.
// ------------------------------------------------------------------------------------------------
#include <thread>
#include <vector>

namespace {

struct Mt {
    uint32_t key[624];
    int pos;
};

inline void mt_seed(Mt &m, uint32_t seed)   // numpy legacy seeding with an integer: init_genrand
{
    m.key[0] = seed;
    for (int i = 1; i < 624; ++i) m.key[i] = 1812433253u * (m.key[i - 1] ^ (m.key[i - 1] >> 30)) + (uint32_t)i;
    m.pos = 624;
}

inline void mt_refill(Mt &m)
{
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX = 0x9908b0dfu;
    uint32_t *k = m.key;
    int i = 0;
    for (; i < 624 - 397; ++i) {
        const uint32_t y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX);
    }
    for (; i < 623; ++i) {
        const uint32_t y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX);
    }
    const uint32_t y = (k[623] & UPPER) | (k[0] & LOWER);
    k[623] = k[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX);
    m.pos = 0;
}

inline uint32_t mt_next(Mt &m)
{
    if (m.pos == 624) mt_refill(m);
    uint32_t y = m.key[m.pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

inline void randint_fill(Mt &m, int32_t high, int64_t count, int32_t *out)
{
    const uint32_t rng = (uint32_t)(high - 1);          // randint(0, high): closed range [0, high-1]
    if (rng == 0) { for (int64_t i = 0; i < count; ++i) out[i] = 0; return; }
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    for (int64_t i = 0; i < count; ++i) {
        uint32_t v;
        while ((v = (mt_next(m) & mask)) > rng) {}
        out[i] = (int32_t)v;
    }
}

}  // namespace

// One stream: continue from (key[624], pos) — np.random.get_state()[1:3] — through `ncalls` calls
// randint(0, high[c], count[c]) written to dst[c]; key/pos are updated for np.random.set_state.
extern "C" int svx_host_randint_stream(uint32_t *key624, int32_t *pos, int ncalls, const int32_t *high,
                                       const int64_t *count, int32_t *const *dst)
{
    if (!key624 || !pos || *pos < 0 || *pos > 624) { svx_set_error("svx_host_randint_stream: bad state"); return SVX_ERR_ARG; }
    Mt m;
    for (int i = 0; i < 624; ++i) m.key[i] = key624[i];
    m.pos = *pos;
    for (int c = 0; c < ncalls; ++c) {
        if (high[c] < 1) { svx_set_error("svx_host_randint_stream: high %d < 1", high[c]); return SVX_ERR_ARG; }
        randint_fill(m, high[c], count[c], dst[c]);
    }
    for (int i = 0; i < 624; ++i) key624[i] = m.key[i];
    *pos = m.pos;
    return SVX_OK;
}

// Independent streams: stream s is seeded like np.random.seed(seeds[s]) and serves calls
// [call_begin[s], call_begin[s+1]).  Streams run on up to `nthreads` host threads.
extern "C" int svx_host_randint_seeded(int nstreams, const uint32_t *seeds, const int64_t *call_begin, const int32_t *high,
                                       const int64_t *count, int32_t *const *dst, int nthreads)
{
    if (nstreams <= 0) return SVX_OK;
    for (int64_t c = call_begin[0]; c < call_begin[nstreams]; ++c)
        if (high[c] < 1) { svx_set_error("svx_host_randint_seeded: high %d < 1", high[c]); return SVX_ERR_ARG; }
    auto work = [&](int t, int nt) {
        for (int s = t; s < nstreams; s += nt) {
            Mt m;
            mt_seed(m, seeds[s]);
            for (int64_t c = call_begin[s]; c < call_begin[s + 1]; ++c) randint_fill(m, high[c], count[c], dst[c]);
        }
    };
    int nt = nthreads < 1 ? 1 : nthreads;
    if (nt > nstreams) nt = nstreams;
    if (nt == 1) { work(0, 1); return SVX_OK; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, t, nt);
    for (auto &th : pool) th.join();
    return SVX_OK;
}

// ---------------------------------------------------------------------------------------------
// Parallel host memcpy (pageable -> pinned staging of host inputs).
// ---------------------------------------------------------------------------------------------
#include <string.h>
extern "C" int svx_host_memcpy(void *dst, const void *src, long long nbytes, int nthreads)
{
    if (nbytes <= 0) return SVX_OK;
    if (!dst || !src) { svx_set_error("svx_host_memcpy: null pointer"); return SVX_ERR_ARG; }
    if (nthreads < 1) nthreads = 1;
    const long long min_slice = 1 << 20;
    if ((long long)nthreads * min_slice > nbytes) nthreads = (int)((nbytes + min_slice - 1) / min_slice);
    if (nthreads <= 1) { memcpy(dst, src, (size_t)nbytes); return SVX_OK; }
    const long long slice = ((nbytes + nthreads - 1) / nthreads + 4095) & ~4095LL;
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        const long long lo = (long long)t * slice;
        if (lo >= nbytes) break;
        const long long n = nbytes - lo < slice ? nbytes - lo : slice;
        pool.emplace_back([=] { memcpy((char *)dst + lo, (const char *)src + lo, (size_t)n); });
    }
    for (auto &th : pool) th.join();
    return SVX_OK;
}

// ---------------------------------------------------------------------------------------------
// (K, N) row tables of make_doc_embedding for many documents at once (the input contract of the path,
// utils/embedding_utils.py:106-203 with overlap_segments=True, the mode seg_align/align.py:222 uses):
//   table[j, i + j] = row of the embedding file whose key is "<first token of line i> <second token of line i+j>",
//   -1 (a zero row) for keys the concatenation file does not hold, for (i, i+j) pairs in the ignore list and every
//   longer concatenation from the same start (:124-126), and for positions before the document start.
// Lines are stripped like Python's str.strip(); an empty line becomes '[BLANK_LINE]' (:29-35); the concatenation
// file maps a stripped line to the FIRST row that carries it (:97-101).  Documents run on `nthreads` host threads.
// The Python twin is embedding_utils.overlap_row_table (tests compare the two).
// ---------------------------------------------------------------------------------------------
#include <string>
#include <unordered_map>

namespace {

inline bool is_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || c == 0x1c || c == 0x1d || c == 0x1e || c == 0x1f || c == 0x85 || c == 0xa0; }

bool read_lines(const char *path, std::vector<std::string> &out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    std::string buf;
    char chunk[1 << 16];
    size_t n;
    while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.append(chunk, n);
    fclose(f);
    size_t pos = 0;
    while (pos < buf.size()) {
        size_t e = buf.find('\n', pos);
        if (e == std::string::npos) e = buf.size();
        size_t a = pos, b = e;
        while (a < b && (unsigned char)buf[a] < 0x80 && is_ws((unsigned char)buf[a])) ++a;
        while (b > a && (unsigned char)buf[b - 1] < 0x80 && is_ws((unsigned char)buf[b - 1])) --b;
        out.emplace_back(buf, a, b - a);
        pos = e + 1;
    }
    return true;
}

// first / second whitespace-separated token of a line (str.split()); false if the line has fewer than `idx + 1`
bool token(const std::string &s, int idx, std::string &tok)
{
    size_t p = 0;
    for (int t = 0;; ++t) {
        while (p < s.size() && (unsigned char)s[p] < 0x80 && is_ws((unsigned char)s[p])) ++p;
        if (p >= s.size()) return false;
        size_t q = p;
        while (q < s.size() && !((unsigned char)s[q] < 0x80 && is_ws((unsigned char)s[q]))) ++q;
        if (t == idx) { tok.assign(s, p, q - p); return true; }
        p = q;
    }
}

}  // namespace

extern "C" int svx_host_overlap_tables(int ndocs, const char *const *seg_paths, const char *const *cat_paths,
                                       const int32_t *const *ignore_pairs, const int32_t *n_ignore, int max_overlaps,
                                       int32_t *const *tables_out, const int32_t *nlines, int32_t *nrows_out, int nthreads)
{
    if (ndocs <= 0) return SVX_OK;
    if (!seg_paths || !cat_paths || !tables_out || !nlines || max_overlaps < 0) { svx_set_error("svx_host_overlap_tables: null argument"); return SVX_ERR_ARG; }
    std::vector<int> status(ndocs, 0);
    auto work = [&](int t, int nt) {
        for (int d = t; d < ndocs; d += nt) {
            std::vector<std::string> lines, cats;
            if (!read_lines(seg_paths[d], lines) || !read_lines(cat_paths[d], cats)) { status[d] = 1; continue; }
            if ((int)lines.size() != nlines[d]) { status[d] = 2; continue; }
            std::unordered_map<std::string, int> key2row;
            key2row.reserve(cats.size() * 2);
            for (size_t i = 0; i < cats.size(); ++i) key2row.emplace(cats[i], (int)i);     // setdefault: first row wins
            if (nrows_out) nrows_out[d] = (int32_t)cats.size();
            const int n = (int)lines.size();
            std::vector<std::string> t0(n), t1(n);
            std::vector<char> ok0(n), ok1(n);
            for (int i = 0; i < n; ++i) {
                const std::string &ln = lines[i].empty() ? std::string("[BLANK_LINE]") : lines[i];
                ok0[i] = token(ln, 0, t0[i]);
                ok1[i] = token(ln, 1, t1[i]);
            }
            int32_t *tab = tables_out[d];
            for (long long i = 0; i < (long long)max_overlaps * n; ++i) tab[i] = -1;
            const int32_t *ign = ignore_pairs ? ignore_pairs[d] : nullptr;
            const int nign = (ign && n_ignore) ? n_ignore[d] : 0;
            std::string key;
            for (int i = 0; i < n && status[d] == 0; ++i) {
                for (int j = i; j < i + max_overlaps && j < n; ++j) {
                    bool ignored = false;
                    for (int q = 0; q < nign; ++q) if (ign[2 * q] == i && ign[2 * q + 1] == j) { ignored = true; break; }
                    if (ignored) break;                                  // PAD from here on (embedding_utils.py:124-126)
                    if (!ok0[i] || !ok1[j]) { status[d] = 3; break; }    // Python: IndexError on .split()[k]
                    key.assign(t0[i]); key.push_back(' '); key.append(t1[j]);
                    auto it = key2row.find(key);
                    if (it != key2row.end()) tab[(size_t)(j - i) * n + j] = it->second;
                }
            }
        }
    };
    int nt = nthreads < 1 ? 1 : nthreads;
    if (nt > ndocs) nt = ndocs;
    if (nt == 1) work(0, 1);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work, t, nt);
        for (auto &th : pool) th.join();
    }
    for (int d = 0; d < ndocs; ++d)
        if (status[d]) {
            svx_set_error("svx_host_overlap_tables: document %d (%s): %s", d, seg_paths[d],
                          status[d] == 1 ? "cannot read the segment / concatenation file" :
                          status[d] == 2 ? "segment file changed length" : "a segment line has fewer than two fields");
            return SVX_ERR_ARG;
        }
    return SVX_OK;
}
