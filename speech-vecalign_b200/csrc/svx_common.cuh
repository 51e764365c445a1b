// svx_common.cuh — launch helpers and error plumbing shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../include/svx.h"
#include "svx_math.h"

void svx_set_error(const char *fmt, ...);
void svx_count_launch(void);

#define SVX_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            svx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return SVX_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define SVX_REQUIRE(cond, code, ...)           \
    do {                                       \
        if (!(cond)) {                         \
            svx_set_error(__VA_ARGS__);        \
            return (code);                     \
        }                                      \
    } while (0)

#define SVX_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess) {                                                               \
            svx_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return SVX_ERR_CUDA;                                                               \
        }                                                                                      \
        svx_count_launch();                                                                    \
    } while (0)

static inline bool svx_dim_supported(int dim)
{
    return dim == 128 || dim == 256 || dim == 512 || dim == 1024;
}

// grid.y is limited to 65535: launchers chunk the job list.
#define SVX_MAX_GRID_Y 65535

__device__ __forceinline__ float4 ldg_f4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
