// Deterministic cross-CTA reductions ("parts" buffers, include/abl_b200.h): producers store one row of partial sums
// per CTA, consumers add the rows in a fixed order.  No floating-point atomics on the product path.
#pragma once
#include <stdio.h>

#include "host_util.h"

namespace abl {

// rows a producer of this library may write: pointwise kernels launch at most kPartWaves blocks per SM when they
// produce partial sums (conv / wgrad kernels: one CTA per SM)
constexpr int kPartWaves = 8;

// out[c] (+)= sum_r parts[HEADER + r*C + c], r in a fixed order: block = 4 float4 columns x 64 row groups; thread
// (col, g) adds rows g, g+64, ... and the 64 group sums are combined in group order through shared memory.
// rows_base: first row (for a parts buffer: parts + ABL_PARTS_HEADER, rows = parts_rows(parts)).
__device__ __forceinline__ int parts_rows(const float* parts) { return *reinterpret_cast<const int*>(parts); }
__device__ __forceinline__ void parts_reduce_block(const float* __restrict__ rows_base, int rows, int C,
                                                   float* __restrict__ out, int accumulate, int block, float4 (*red)[4]) {
    const int C4 = C >> 2;
    const int cl = threadIdx.x & 3, g = threadIdx.x >> 2;
    const int c4 = block * 4 + cl;
    const float4* base = reinterpret_cast<const float4*>(rows_base);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < C4) {
#pragma unroll 4
        for (int r = g; r < rows; r += 64) {
            const float4 v = __ldcg(base + static_cast<size_t>(r) * C4 + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    red[g][cl] = s;
    __syncthreads();
    if (g == 0 && c4 < C4) {
        float4* o = reinterpret_cast<float4*>(out) + c4;
        float4 t = accumulate ? *o : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int q = 0; q < 64; ++q) {
            const float4 v = red[q][cl];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        *o = t;
    }
}


static inline int parts_check(const float* parts, long long bytes, int rows, int C, const char* who) {
    if ((reinterpret_cast<uintptr_t>(parts) & 15) != 0 || (C & 3) != 0) {
        char buf[160]; snprintf(buf, sizeof(buf), "%s: parts buffer must be 16-byte aligned and C a multiple of 4", who);
        set_error(buf); return ABL_ERR_ALIGN;
    }
    if (bytes < (static_cast<long long>(ABL_PARTS_HEADER) + static_cast<long long>(rows) * C) * 4) {
        char buf[160]; snprintf(buf, sizeof(buf), "%s: parts buffer too small (see abl_parts_bytes)", who);
        set_error(buf); return ABL_ERR_SHAPE;
    }
    return ABL_OK;
}

}  // namespace abl
