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
        default: return -1;
    }
}
