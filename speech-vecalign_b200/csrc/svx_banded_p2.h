// svx_banded_p2.h — launcher of the packed (FFMA2) register-blocked banded cost kernel (banded_p2.cu).
#pragma once
#include "../../include/svx.h"

// segment positions (x and y together) a tile of `ne` block-diagonals can touch: 2 ne + 1 anti-diagonals + 2 (B - 1) + rounding
constexpr int svx_p2_rows_cap(int band, int ne) { return 2 * band + 2 * ne + 6; }

// Returns -1 when the shape is outside the kernel's limits (the caller falls back), else an SVX status.
int svx_launch_costs_p2(int K, const SvxBandJob *jobs_d, int nj, int max_alen, int band, int dim, int mode, int bc,
                        int nstages, int nprod, int cons_warps, void *stream);
