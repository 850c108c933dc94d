// HBM-bound kernels of the cutoffvgg16 path: weight re-layout, max-pool/ReLU backward, global
// average pool (+backward), Dense/softmax/cross-entropy head, Adam / RMSprop, Grad-CAM.
// All are grid-stride / vectorised (16-byte) kernels sized in multiples of the SM count.
#include <stdio.h>

#include "host_util.h"
#include "ptx.cuh"
#include "parts.cuh"

namespace abl {

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// ------------------------------------------------------------------ weight re-layout
__global__ void prep_conv_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_fwd,
                                         __nv_bfloat16* __restrict__ w_dgrad, int Cin, int Cout) {
    const int total = 9 * Cin * Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout;
        const int ci = (i / Cout) % Cin;
        const int t = i / (Cout * Cin);
        const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
        const int kh = t / 3, kw = t - kh * 3;
        // GEMM operands are tap-indexed kw-major (tap' = kw*3 + kh); dgrad uses the 180-degree rotated filter
        if (w_fwd) w_fwd[(static_cast<size_t>(kw * 3 + kh) * Cout + co) * Cin + ci] = v;
        if (w_dgrad) w_dgrad[(static_cast<size_t>((2 - kw) * 3 + (2 - kh)) * Cin + ci) * Cout + co] = v;
    }
}

// All layers in one launch (after every optimizer step): blockIdx.y = layer.
constexpr int kMaxPrepLayers = 16;
struct PrepMulti {
    const float* w[kMaxPrepLayers];
    __nv_bfloat16* w_fwd[kMaxPrepLayers];
    __nv_bfloat16* w_dgrad[kMaxPrepLayers];
    int cin[kMaxPrepLayers], cout[kMaxPrepLayers];
    int tile_end[kMaxPrepLayers];      // running total of 32x32 tiles (9 taps x Cin/32 x Cout/32 per layer)
    int n_layers;
};
__global__ void __launch_bounds__(256)
prep_conv_weights_multi_kernel(const __grid_constant__ PrepMulti m) {
    // One block = one 32 (ci) x 32 (co) tile of one tap of one layer, transposed through shared memory: the fp32 source
    // rows (co fastest) are read coalesced once, the dgrad copy (same orientation) and the forward copy (transposed) are
    // both written coalesced.  (The first version indexed by output element and read the source with a stride of Cout
    // floats: 21-25 us for 3.5 MB; this one is bound by launch latency.)
    __shared__ float tile[32][33];
    int l = 0, t = blockIdx.x;
    while (l + 1 < m.n_layers && t >= m.tile_end[l]) ++l;
    if (l > 0) t -= m.tile_end[l - 1];
    const int Cin = m.cin[l], Cout = m.cout[l];
    const int tiles_co = Cout >> 5, tiles_ci = Cin >> 5;
    const int tap = t / (tiles_ci * tiles_co);
    t -= tap * tiles_ci * tiles_co;
    const int ci0 = (t / tiles_co) << 5, co0 = (t % tiles_co) << 5;
    const int kh = tap / 3, kw = tap - kh * 3;
    const float* __restrict__ w = m.w[l] + static_cast<size_t>(tap) * Cin * Cout;      // HWIO: tap = kh*3 + kw
    __nv_bfloat16* __restrict__ w_fwd = m.w_fwd[l];
    __nv_bfloat16* __restrict__ w_dgrad = m.w_dgrad[l];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const float v = w[static_cast<size_t>(ci0 + r) * Cout + co0 + tx];
        tile[r][tx] = v;
        // w_dgrad[t'][ci][co], t' = (2-kw)*3 + (2-kh): same inner order as the source
        if (w_dgrad)
            w_dgrad[(static_cast<size_t>((2 - kw) * 3 + (2 - kh)) * Cin + ci0 + r) * Cout + co0 + tx] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    if (w_fwd) {        // w_fwd[t'][co][ci], t' = kw*3 + kh
#pragma unroll
        for (int r = ty; r < 32; r += 8)
            w_fwd[(static_cast<size_t>(kw * 3 + kh) * Cout + co0 + r) * Cin + ci0 + tx] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// ------------------------------------------------------------------ MaxPoolGrad + ReluGrad
// One thread per (pooled pixel, 8-channel chunk): 16-byte loads of the 4 window activations, the
// pooled value and dP; writes the 4 dZ chunks.  First maximum in window scan order wins.
// Block-level finish of a fused BiasAddGrad: every thread owns ONE 8-channel chunk for its whole grid-stride
// walk (blockDim % C8 == 0 and the stride is a multiple of blockDim), so it arrives here with 8 partial sums;
// they are reduced over the blockDim/C8 threads sharing the chunk in shared memory (fixed order) and the block
// stores ONE row of a parts buffer (include/abl_b200.h "Partial-sum buffers"): no floating-point atomics, the
// consumer (abl_conv3x3_wgrad / abl_parts_reduce) adds the rows in a fixed order -> bit-reproducible gradients.
__device__ __forceinline__ void bias_partial_flush(const float (&acc)[8], float* __restrict__ parts, int C8, float* red) {
    const int chunk = threadIdx.x % C8, rg = threadIdx.x / C8, rgs = blockDim.x / C8, C = C8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rg * C + chunk * 8 + j] = acc[j];
    __syncthreads();
    float* row = parts + ABL_PARTS_HEADER + static_cast<size_t>(blockIdx.x) * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < rgs; ++g) s += red[g * C + c];
        row[c] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<int*>(parts) = static_cast<int>(gridDim.x);
}

__global__ void __launch_bounds__(256)
maxpool_bwd_relu_kernel(const uint4* __restrict__ dp, const uint4* __restrict__ a,
                                        const uint4* __restrict__ pooled, uint4* __restrict__ dz, int B, int H,
                                        int W, int C8, float* __restrict__ db) {
    extern __shared__ float red[];       // [256] x 8 floats when db != nullptr
    float bacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int PH = H >> 1, PW = W >> 1;
    const long long total = static_cast<long long>(B) * PH * PW * C8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C8);
        long long r = i / C8;
        const int pw = static_cast<int>(r % PW);
        r /= PW;
        const int ph = static_cast<int>(r % PH);
        const int n = static_cast<int>(r / PH);
        const uint4 g = dp[i];
        const uint4 pm = pooled[i];
        const long long base = ((static_cast<long long>(n) * H + 2 * ph) * W + 2 * pw) * C8 + c;
        const long long offs[4] = {base, base + C8, base + static_cast<long long>(W) * C8,
                                   base + static_cast<long long>(W) * C8 + C8};
        uint4 av[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) av[k] = a[offs[k]];
        const uint32_t gv[4] = {g.x, g.y, g.z, g.w};
        const uint32_t pv[4] = {pm.x, pm.y, pm.z, pm.w};
        uint32_t out[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t taken_lo = 0, taken_hi = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = (j == 0 ? av[k].x : j == 1 ? av[k].y : j == 2 ? av[k].z : av[k].w);
                // activations are post-ReLU (>= 0, never -0 from fmaxf(.,0)): bit equality == value equality
                const bool lo = !taken_lo && ((x & 0xFFFFu) == (pv[j] & 0xFFFFu)) && ((x & 0x7FFFu) != 0);
                const bool hi = !taken_hi && ((x >> 16) == (pv[j] >> 16)) && ((x & 0x7FFF0000u) != 0);
                const bool lo_eq = ((x & 0xFFFFu) == (pv[j] & 0xFFFFu));
                const bool hi_eq = ((x >> 16) == (pv[j] >> 16));
                out[k][j] = (lo ? (gv[j] & 0xFFFFu) : 0u) | (hi ? (gv[j] & 0xFFFF0000u) : 0u);
                taken_lo |= lo_eq;   // the first maximum claims the window even if it is 0 (then masked)
                taken_hi |= hi_eq;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dz[offs[k]] = make_uint4(out[k][0], out[k][1], out[k][2], out[k][3]);
        if (db != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    bacc[2 * j] += bf16_lo(out[k][j]);
                    bacc[2 * j + 1] += bf16_hi(out[k][j]);
                }
        }
    }
    if (db != nullptr) bias_partial_flush(bacc, db, C8, red);
}

// MaxPoolGrad from the 2-bit routing codes the forward epilogue stored (abl_conv3x3_igemm pool_code): the
// full-resolution activation is neither kept nor re-read.  dp must already carry the ReLU mask of the pooled value
// (the dgrad that produces it runs with ABL_EPI_RELU_MASK against the pooled tensor: relu'(max) == relu' of the
// routed element).  Traffic per full-resolution element: 2 B written + 0.56 B read (was 2 + 3 B).
__global__ void __launch_bounds__(256)
maxpool_bwd_code_kernel(const uint4* __restrict__ dp, const uint16_t* __restrict__ code, uint4* __restrict__ dz, int B,
                        int H, int W, int C8, float* __restrict__ db) {
    extern __shared__ float red[];
    float bacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int PH = H >> 1, PW = W >> 1;
    const long long total = static_cast<long long>(B) * PH * PW * C8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C8);
        long long r = i / C8;
        const int pw = static_cast<int>(r % PW);
        r /= PW;
        const int ph = static_cast<int>(r % PH);
        const int n = static_cast<int>(r / PH);
        const uint4 g = __ldg(dp + i);
        const uint32_t cd = __ldg(code + i);
        const uint32_t gv[4] = {g.x, g.y, g.z, g.w};
        const long long base = ((static_cast<long long>(n) * H + 2 * ph) * W + 2 * pw) * C8 + c;
        const long long offs[4] = {base, base + C8, base + static_cast<long long>(W) * C8,
                                   base + static_cast<long long>(W) * C8 + C8};
        uint32_t out[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t lo = (cd >> (4 * j)) & 3u, hi = (cd >> (4 * j + 2)) & 3u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                out[k][j] = (lo == static_cast<uint32_t>(k) ? (gv[j] & 0xFFFFu) : 0u) |
                            (hi == static_cast<uint32_t>(k) ? (gv[j] & 0xFFFF0000u) : 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dz[offs[k]] = make_uint4(out[k][0], out[k][1], out[k][2], out[k][3]);
        if (db != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { bacc[2 * j] += bf16_lo(gv[j]); bacc[2 * j + 1] += bf16_hi(gv[j]); }
        }
    }
    if (db != nullptr) bias_partial_flush(bacc, db, C8, red);
}

// ------------------------------------------------------------------ global average pool
// grid = (B, C/64): each block reduces HW rows x 64 channels (128 B per row) of one frame.
__global__ void __launch_bounds__(256) gap_fwd_kernel(const uint4* __restrict__ a, float* __restrict__ gap,
                                                      int HW, int C) {
    __shared__ float red[32][65];
    const int n = blockIdx.x, cb = blockIdx.y;
    const int chunk = threadIdx.x & 7;        // 8 x 16 B = 64 channels
    const int rg = threadIdx.x >> 3;          // 32 row groups
    const int C8 = C >> 3;
    const uint4* base = a + static_cast<size_t>(n) * HW * C8 + cb * 8 + chunk;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = rg; r < HW; r += 32) {
        const uint4 v = __ldg(base + static_cast<size_t>(r) * C8);
        acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x);
        acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
        acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z);
        acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rg][chunk * 8 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
        gap[static_cast<size_t>(n) * C + cb * 64 + threadIdx.x] = s / static_cast<float>(HW);
    }
}

// gap[n][c] = (1/HW) * sum_p part[n*parts + p][c], p in fixed order: finishes the GAP partial sums that
// abl_conv3x3_igemm_gap's epilogue wrote per 8x16-pixel rectangle (deterministic, 8 KiB read per frame).
__global__ void gap_reduce_kernel(const float4* __restrict__ part, float4* __restrict__ gap, int B, int parts, int C4,
                                  float inv_hw) {
    const int total = B * C4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / C4, c = i - n * C4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < parts; ++q) {
            const float4 v = __ldg(part + (static_cast<size_t>(n) * parts + q) * C4 + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        gap[i] = make_float4(s.x * inv_hw, s.y * inv_hw, s.z * inv_hw, s.w * inv_hw);
    }
}

// Index arithmetic: the block's 256 threads cover 256 / C8 whole pixels per iteration, so a thread's channel chunk is
// fixed and the frame index is one 32-bit division per pixel.
__global__ void __launch_bounds__(256)
gap_bwd_relu_kernel(const float* __restrict__ dgap, const uint4* __restrict__ a,
                                    uint4* __restrict__ dz, int B, int HW, int C8, float inv_hw,
                                    float* __restrict__ db) {
    extern __shared__ float red[];
    float bacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned rgs = 256u / C8, c = threadIdx.x % C8, npix = static_cast<unsigned>(B) * HW;
    const unsigned stride = gridDim.x * rgs;
    constexpr int kU = 4;       // four activation loads in flight per thread (see gap_bwd_bits_kernel)
    // (C8 not dividing 256: the last 256 % C8 threads idle)
    for (unsigned pix0 = threadIdx.x / C8 < rgs ? blockIdx.x * rgs + threadIdx.x / C8 : npix; pix0 < npix;
         pix0 += kU * stride) {
        uint4 av4[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const unsigned pix = pix0 + u * stride;
            av4[u] = pix < npix ? a[static_cast<size_t>(pix) * C8 + c] : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const unsigned pix = pix0 + u * stride;
            if (pix >= npix) break;
            const unsigned n = pix / static_cast<unsigned>(HW);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(dgap + (static_cast<size_t>(n) * C8 + c) * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(dgap + (static_cast<size_t>(n) * C8 + c) * 8) + 1);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const uint32_t av[4] = {av4[u].x, av4[u].y, av4[u].z, av4[u].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float lo = bf16_lo(av[j]) > 0.f ? g[2 * j] * inv_hw : 0.f;
                const float hi = bf16_hi(av[j]) > 0.f ? g[2 * j + 1] * inv_hw : 0.f;
                o[j] = pack_bf16x2(lo, hi);
                if (db != nullptr) { bacc[2 * j] += bf16_lo(o[j]); bacc[2 * j + 1] += bf16_hi(o[j]); }   // sum of the ROUNDED dz
            }
            dz[static_cast<size_t>(pix) * C8 + c] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if (db != nullptr) bias_partial_flush(bacc, db, C8, red);
}

// The same dZ from the 1-bit ReLU mask the last conv wrote instead of its activation (abl_conv3x3_igemm_gap relu_bits):
// 1 bit instead of 16 read per element, and the forward never wrote the activation.  One thread = 8 channels of a pixel.
// A block works on `slice_px` pixels of ONE frame at a time (work items = frames x slices, strided over the grid), so
// the eight rounded values dgap[n][c] / HW of a thread are loop-invariant: per pixel the thread loads one mask byte,
// expands it to four bf16x2 AND-masks (multiply the nibble by 0x10204080: bit k lands on the top bit of byte k; PRMT
// replicates a byte's top bit over a half word) and stores 16 bytes -- ~30 instructions instead of the ~100 of the
// select / multiply / convert form.  The bias gradient is value x count: the
// mask bits of a slice are counted in byte lanes (<= 255 pixels per thread and slice), and n * q is exact in fp32.
// prmt.b32 in its default mode: a selector nibble 8 + k replicates the TOP BIT of source byte k over the result byte
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t selector) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(selector));
    return d;
}
__global__ void __launch_bounds__(256)
gap_bwd_bits_kernel(const float* __restrict__ dgap, const unsigned char* __restrict__ bits,
                    uint4* __restrict__ dz, int B, int HW, int C8, float inv_hw, int slices, int slice_px,
                    float* __restrict__ db) {
    extern __shared__ float red[];
    float bacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned rgs = 256u / C8, c = threadIdx.x % C8, prow = threadIdx.x / C8;
    const int items = B * slices;
    constexpr int kPx = 8;                           // pixels per thread and item (the host sizes slice_px = kPx * rgs)
    // the mask bytes of the NEXT item are in flight while this one is expanded and stored (ncu on the first form: the
    // kernel waited on one 1-byte load per 16-byte store -- 22 stalled warps per issue on `long_scoreboard`)
    auto fetch = [&](int item, uint32_t (&on)[kPx]) {
        const int n = item / slices, sl = item - n * slices;
        const int p0 = sl * slice_px, p1 = min(HW, p0 + slice_px);
#pragma unroll
        for (int u = 0; u < kPx; ++u) {
            const int px = p0 + static_cast<int>(prow + u * rgs);
            // little-endian words: byte i = channels 8 i .. 8 i + 7 of the flattened [pixel][C/8] array
            on[u] = (item < items && prow < rgs && px < p1) ? __ldg(bits + (static_cast<size_t>(n) * HW + px) * C8 + c) : 0u;
        }
    };
    uint32_t cur[kPx], nxt[kPx];
    fetch(blockIdx.x, cur);
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        fetch(item + gridDim.x, nxt);
        const int n = item / slices, sl = item - n * slices;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(dgap + (static_cast<size_t>(n) * C8 + c) * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(dgap + (static_cast<size_t>(n) * C8 + c) * 8) + 1);
        const uint32_t q[4] = {pack_bf16x2(g0.x * inv_hw, g0.y * inv_hw), pack_bf16x2(g0.z * inv_hw, g0.w * inv_hw),
                               pack_bf16x2(g1.x * inv_hw, g1.y * inv_hw), pack_bf16x2(g1.z * inv_hw, g1.w * inv_hw)};
        uint32_t cnt_lo = 0u, cnt_hi = 0u;           // byte k: pixels of this slice whose channel k (4 + k) is on
        const int p0 = sl * slice_px, p1 = min(HW, p0 + slice_px);
#pragma unroll
        for (int u = 0; u < kPx; ++u) {
            const int px = p0 + static_cast<int>(prow + u * rgs);
            if (prow < rgs && px < p1) {              // (C8 not dividing 256: the last 256 % C8 threads idle)
                const uint32_t on = cur[u];
                const uint32_t x = (on & 0xFu) * 0x10204080u, y = (on >> 4) * 0x10204080u;
                dz[(static_cast<size_t>(n) * HW + px) * C8 + c] =
                    make_uint4(q[0] & prmt_b32(x, 0x9988u), q[1] & prmt_b32(x, 0xBBAAu), q[2] & prmt_b32(y, 0x9988u),
                               q[3] & prmt_b32(y, 0xBBAAu));
                cnt_lo += (x >> 7) & 0x01010101u;
                cnt_hi += (y >> 7) & 0x01010101u;
            }
        }
        if (db != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bacc[2 * j] += bf16_lo(q[j]) * static_cast<float>((j < 2 ? cnt_lo >> (16 * j) : cnt_hi >> (16 * (j - 2))) & 0xFFu);
                bacc[2 * j + 1] += bf16_hi(q[j]) *
                                   static_cast<float>((j < 2 ? cnt_lo >> (16 * j + 8) : cnt_hi >> (16 * (j - 2) + 8)) & 0xFFu);
            }
        }
#pragma unroll
        for (int u = 0; u < kPx; ++u) cur[u] = nxt[u];
    }
    if (db != nullptr) bias_partial_flush(bacc, db, C8, red);
}

// ------------------------------------------------------------------ head
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
    // splitmix64 finaliser -> 32 uniform bits (counter-based: reproducible across ranks/steps)
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<uint32_t>(z >> 32);
}

constexpr int kMaxClasses = 8;

// One warp per frame: lanes stride over channels, shuffle-reduce the n logits.
// Loss: every block adds its (<= 4) frames in warp order and stores ONE partial into loss_ws; the block that arrives
// last (integer ticket in loss_ws word 0, reset for the next launch) adds all partials in block order and does the
// single read-modify-write of *loss_sum -> the loss is bit-reproducible (no floating-point atomics).
__global__ void __launch_bounds__(128)
head_fwd_kernel(float* __restrict__ gap, const float* __restrict__ gap_part, int gap_parts, float gap_inv_hw,
                const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ logits, float* __restrict__ probs, const float* __restrict__ y,
                const float* __restrict__ cw, float inv_batch, float* __restrict__ loss_sum,
                float* __restrict__ dlogits, unsigned char* __restrict__ keep_mask, int keep_mask_in, int training,
                float dropout, unsigned long long seed, unsigned int step, unsigned int rank, int B, int C, int n,
                unsigned int* __restrict__ step_dev, float* __restrict__ loss_ws, int loss_accumulate) {
    __shared__ float wloss[4];
    __shared__ float red[128];
    __shared__ int is_last;
    if (step_dev != nullptr) step = *step_dev;      // CUDA-graph replays: the counter lives on the device
    const int lane = threadIdx.x & 31;
    const int frame = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    float my_loss = 0.f;
    if (frame < B) {
    const float keep_scale = 1.f / (1.f - dropout);
    const uint32_t thresh = static_cast<uint32_t>(fminf(dropout, 0.99999994f) * 4294967296.0);
    float z[kMaxClasses];
#pragma unroll
    for (int j = 0; j < kMaxClasses; ++j) z[j] = 0.f;
#pragma unroll 2
    for (int c = lane; c < C; c += 32) {
        float h;
        if (gap_part != nullptr) {
            // GlobalAveragePooling finish folded in (abl_gap_reduce's arithmetic, rectangle order): the frame's mean is
            // stored for head_bwd
            float sum = 0.f;
#pragma unroll 8
            for (int q = 0; q < gap_parts; ++q) sum += __ldg(gap_part + (static_cast<size_t>(frame) * gap_parts + q) * C + c);
            h = sum * gap_inv_hw;
            gap[static_cast<size_t>(frame) * C + c] = h;
        } else {
            h = gap[static_cast<size_t>(frame) * C + c];
        }
        if (training) {
            unsigned char keep;
            const size_t mi = static_cast<size_t>(frame) * C + c;
            if (keep_mask_in) {
                keep = keep_mask[mi];
            } else {
                const uint64_t ctr = (static_cast<uint64_t>(step) << 40) ^ (static_cast<uint64_t>(rank) << 32) ^ mi;
                keep = mix32(seed ^ (ctr * 0xD1342543DE82EF95ull)) >= thresh ? 1 : 0;
                keep_mask[mi] = keep;
            }
            h = keep ? h * keep_scale : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kMaxClasses; ++j)
            if (j < n) z[j] += h * __ldg(w + c * n + j);
    }
#pragma unroll
    for (int j = 0; j < kMaxClasses; ++j)
        if (j < n) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) z[j] += __shfl_xor_sync(0xffffffffu, z[j], o);
            z[j] += b[j];
        }
    if (lane == 0) {
        float mx = z[0];
        for (int j = 1; j < n; ++j) mx = fmaxf(mx, z[j]);
        float e[kMaxClasses], s = 0.f;
        for (int j = 0; j < n; ++j) { e[j] = expf(z[j] - mx); s += e[j]; }
        const float inv = 1.f / s;
        float wi = 0.f, zy = 0.f;
        for (int j = 0; j < n; ++j) {
            logits[frame * n + j] = z[j];
            probs[frame * n + j] = e[j] * inv;
            if (y) {
                const float yj = y[frame * n + j];
                wi += yj * (cw ? cw[j] : 1.f);
                zy += yj * z[j];
            }
        }
        if (y) {
            const float lse = mx + logf(s);
            my_loss = wi * (lse - zy) * inv_batch;
            if (dlogits)
                for (int j = 0; j < n; ++j) dlogits[frame * n + j] = wi * inv_batch * (e[j] * inv - y[frame * n + j]);
        }
    }
    }
    if (loss_sum == nullptr || y == nullptr) return;          // uniform for the whole grid
    if (lane == 0) wloss[threadIdx.x >> 5] = my_loss;
    __syncthreads();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(loss_ws);
    float* part = loss_ws + ABL_PARTS_HEADER;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = (wloss[0] + wloss[1]) + (wloss[2] + wloss[3]);
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;     // integer ticket: order-independent
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float s = 0.f;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(part + i);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *loss_sum = loss_accumulate ? *loss_sum + red[0] : red[0];
        *ticket = 0u;
        // CUDA-graph replays: every block has read the dropout step by now (its ticket came after its frames), so the
        // last block advances the device-side counters: [0] dropout step, [1] optimizer step (read by abl_adam_flat_tab
        // later in the same step)
        if (step_dev != nullptr) { step_dev[0] = step + 1u; step_dev[1] += 1u; }
    }
}

// Deterministic Dense backward without partial buffers: block = 8 channels x 32 frame lanes (one warp per channel,
// lane l walks frames l, l+32, ...), so every (frame, channel) element is visited by exactly one thread (dgap) and a
// channel's dw[c][:] is one fixed-shape shuffle tree over its warp.  db[j] = sum_f dlogits[f][j] is block 0's job.
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ gap, const float* __restrict__ w, const float* __restrict__ dlogits,
                const unsigned char* __restrict__ keep_mask, float dropout, float* __restrict__ dw,
                float* __restrict__ db, float* __restrict__ dgap, int B, int C, int n) {
    __shared__ float red[256];
    const float keep_scale = 1.f / (1.f - dropout);
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c < C) {
        float wrow[kMaxClasses], acc[kMaxClasses];
#pragma unroll
        for (int j = 0; j < kMaxClasses; ++j) { wrow[j] = j < n ? w[c * n + j] : 0.f; acc[j] = 0.f; }
        for (int f = lane; f < B; f += 32) {
            const size_t idx = static_cast<size_t>(f) * C + c;
            const float scale = keep_mask ? (keep_mask[idx] ? keep_scale : 0.f) : 1.f;
            const float h = gap[idx] * scale;
            float dg = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxClasses; ++j)
                if (j < n) {
                    const float d = __ldg(dlogits + f * n + j);
                    acc[j] += h * d;
                    dg += d * wrow[j];
                }
            dgap[idx] = dg * scale;
        }
#pragma unroll
        for (int j = 0; j < kMaxClasses; ++j)
            if (j < n) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
                if (lane == 0) dw[c * n + j] = acc[j];
            }
    }
    if (blockIdx.x == 0) {
        for (int j = 0; j < n; ++j) {
            float s = 0.f;
            for (int f = threadIdx.x; f < B; f += blockDim.x) s += __ldg(dlogits + f * n + j);
            red[threadIdx.x] = s;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) db[j] = red[0];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ optimizers (28 / 20 B per param)
// One element of Keras 2.9 Adam / RMSprop.  Shared by the local kernels and the peer-memory data-parallel kernel and
// written with explicit roundings (no FMA contraction left to the compiler), so both paths produce the SAME bits from
// the same gradient.
__device__ __forceinline__ void adam_update(float& t, float& m, float& v, float g, float alpha, float b1, float b2,
                                            float ob1, float ob2, float eps) {
    m = __fadd_rn(__fmul_rn(b1, m), __fmul_rn(ob1, g));
    v = __fadd_rn(__fmul_rn(b2, v), __fmul_rn(ob2, __fmul_rn(g, g)));
    t = __fsub_rn(t, __fdiv_rn(__fmul_rn(alpha, m), __fadd_rn(__fsqrt_rn(v), eps)));
}
__device__ __forceinline__ void rmsprop_update(float& t, float& r, float g, float lr, float rho, float orho, float eps) {
    r = __fadd_rn(__fmul_rn(rho, r), __fmul_rn(orho, __fmul_rn(g, g)));
    t = __fsub_rn(t, __fdiv_rn(__fmul_rn(lr, g), __fadd_rn(__fsqrt_rn(r), eps)));
}

__global__ void adam_flat_kernel(float4* __restrict__ theta, const float4* __restrict__ g, float4* __restrict__ m,
                                 float4* __restrict__ v, long long n4, float* theta_tail, const float* g_tail,
                                 float* m_tail, float* v_tail, int tail, float alpha, float b1, float b2, float eps,
                                 const float* __restrict__ alpha_tab, const int* __restrict__ t_dev, int tab_len) {
    // CUDA-graph replays: the step count lives on the device and alpha_t = lr*sqrt(1-b2^t)/(1-b1^t) comes from a table
    // the host computed with the same arithmetic as the eager path (entries beyond the table equal its last one)
    if (alpha_tab != nullptr) alpha = alpha_tab[min(max(*t_dev, 1), tab_len) - 1];
    const float ob1 = 1.f - b1, ob2 = 1.f - b2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 gg = g[i];
        float4 mm = m[i], vv = v[i], tt = theta[i];
        adam_update(tt.x, mm.x, vv.x, gg.x, alpha, b1, b2, ob1, ob2, eps);
        adam_update(tt.y, mm.y, vv.y, gg.y, alpha, b1, b2, ob1, ob2, eps);
        adam_update(tt.z, mm.z, vv.z, gg.z, alpha, b1, b2, ob1, ob2, eps);
        adam_update(tt.w, mm.w, vv.w, gg.w, alpha, b1, b2, ob1, ob2, eps);
        m[i] = mm; v[i] = vv; theta[i] = tt;
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) {
        const int i = threadIdx.x;
        float tt = theta_tail[i], mm = m_tail[i], vv = v_tail[i];
        adam_update(tt, mm, vv, g_tail[i], alpha, b1, b2, ob1, ob2, eps);
        theta_tail[i] = tt; m_tail[i] = mm; v_tail[i] = vv;
    }
}

__global__ void rmsprop_flat_kernel(float4* __restrict__ theta, const float4* __restrict__ g,
                                    float4* __restrict__ rms, long long n4, float* theta_tail, const float* g_tail,
                                    float* rms_tail, int tail, float lr, float rho, float eps) {
    const float orho = 1.f - rho;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 gg = g[i];
        float4 rr = rms[i], tt = theta[i];
        rmsprop_update(tt.x, rr.x, gg.x, lr, rho, orho, eps);
        rmsprop_update(tt.y, rr.y, gg.y, lr, rho, orho, eps);
        rmsprop_update(tt.z, rr.z, gg.z, lr, rho, orho, eps);
        rmsprop_update(tt.w, rr.w, gg.w, lr, rho, orho, eps);
        rms[i] = rr; theta[i] = tt;
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) {
        const int i = threadIdx.x;
        float tt = theta_tail[i], rr = rms_tail[i];
        rmsprop_update(tt, rr, g_tail[i], lr, rho, orho, eps);
        theta_tail[i] = tt; rms_tail[i] = rr;
    }
}

// ------------------------------------------------------------------ nearest-neighbour resize of raw uint8 frames
// tf.image.resize(method=NEAREST_NEIGHBOR) as the reference's _parse_fn applies it (preprocessor.py:76-80): TF2 uses
// half-pixel centres, src = min(floor((dst + 0.5) * in/out), in - 1), evaluated in fp32 like TF's kernel.  One thread
// per output pixel (3 bytes); the result feeds block1_conv1's uint8 path, which applies vgg16.preprocess_input.
__global__ void resize_nearest_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int B, int h,
                                         int w, int H, int W) {
    const float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
    const long long total = static_cast<long long>(B) * H * W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % W);
        const int oy = static_cast<int>((i / W) % H);
        const long long n = i / (static_cast<long long>(W) * H);
        const int iy = min(static_cast<int>(floorf((oy + 0.5f) * sy)), h - 1);
        const int ix = min(static_cast<int>(floorf((ox + 0.5f) * sx)), w - 1);
        const uint8_t* s = src + ((n * h + iy) * w + ix) * 3;
        uint8_t* d = dst + i * 3;
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
    }
}

// ------------------------------------------------------------------ clip-level aggregation (predict.py:159-166)
// One warp per clip over that clip's rows of the frame-probability table.  Per clip it produces everything the three
// reference rules need: the class-wise mean ('average', predict.py:164), the longest run of frames with
// p[b] > threshold ('contiguous', predict.py:283-308) and the highest mean of p[b] over `window` consecutive frames
// ('sliding_window', predict.py:365-376; 0 when the clip is shorter than the window, as in the reference's loop).
// Each lane scans a contiguous chunk of frames; chunks are combined with the run-length monoid
// (len, all_ones, prefix_run, suffix_run, best_run) through shuffles.  Window sums use a lane-local sliding sum.
struct RunState { int len, pre, suf, best; };
__device__ __forceinline__ RunState run_combine(const RunState a, const RunState b) {
    RunState r;
    r.len = a.len + b.len;
    r.pre = (a.pre == a.len) ? a.len + b.pre : a.pre;
    r.suf = (b.suf == b.len) ? b.len + a.suf : b.suf;
    r.best = max(max(a.best, b.best), a.suf + b.pre);
    return r;
}

__global__ void __launch_bounds__(128)
clip_aggregate_kernel(const float* __restrict__ probs, const long long* __restrict__ offsets, int n_clips, int n,
                      int b_idx, float threshold, int window, float* __restrict__ mean_out, int* __restrict__ run_out,
                      float* __restrict__ win_out) {
    const int lane = threadIdx.x & 31;
    const int clip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (clip >= n_clips) return;
    const long long f0 = offsets[clip], f1 = offsets[clip + 1];
    const int nf = static_cast<int>(f1 - f0);
    const float* p = probs + f0 * n;
    // class-wise means: lanes stride over (frame, class) pairs of each class
    for (int j = 0; j < n; ++j) {
        float s = 0.f;
        for (int f = lane; f < nf; f += 32) s += p[static_cast<long long>(f) * n + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) mean_out[static_cast<long long>(clip) * n + j] = nf > 0 ? s / static_cast<float>(nf) : 0.f;
    }
    // longest run of positives: contiguous chunk per lane, then an ordered combine over lanes
    const int per = (nf + 31) / 32;
    const int c0 = min(nf, lane * per), c1 = min(nf, c0 + per);
    RunState st = {c1 - c0, 0, 0, 0};
    {
        int cur = 0;
        bool prefix = true;
        for (int f = c0; f < c1; ++f) {
            const bool on = p[static_cast<long long>(f) * n + b_idx] > threshold;
            cur = on ? cur + 1 : 0;
            if (prefix) { if (on) st.pre = cur; else prefix = false; }
            st.best = max(st.best, cur);
        }
        st.suf = cur;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {      // tree combine keeping lane order: lane l absorbs lane l + o
        RunState other;
        other.len = __shfl_down_sync(0xffffffffu, st.len, o);
        other.pre = __shfl_down_sync(0xffffffffu, st.pre, o);
        other.suf = __shfl_down_sync(0xffffffffu, st.suf, o);
        other.best = __shfl_down_sync(0xffffffffu, st.best, o);
        if ((lane & (2 * o - 1)) == 0 && lane + o < 32) st = run_combine(st, other);
    }
    if (lane == 0) run_out[clip] = st.best;
    // sliding window: windows starting at i in [0, nf - window]; lanes stride over start positions
    float best = 0.f;
    if (window > 0 && nf >= window) {
        const float inv = 1.f / static_cast<float>(window);
        for (int i = lane; i + window <= nf; i += 32) {
            float s = 0.f;
            for (int k = 0; k < window; ++k) s += p[static_cast<long long>(i + k) * n + b_idx];
            best = fmaxf(best, s * inv);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) win_out[clip] = best;
}

// ------------------------------------------------------------------ data-parallel optimizer over NVLink peer memory
// ONE kernel = reduce-scatter(grad) -> Adam / RMSprop on this rank's shard -> all-gather(theta), all through peer
// pointers of a symmetric allocation (every rank maps every rank's flat theta / grad buffer):
//   g     = sum_r grad_r[i]         peer LOADS over NVLink, fixed rank order (every rank would compute the same bits)
//   theta = update(theta, g, m, v)  optimizer state is touched for this rank's 1/world shard only
//   theta_r[i] = theta  for all r   peer STORES: every rank ends the step with identical weights
// so a step moves (world-1)/world x 4 bytes/param in each direction per GPU instead of a ring all-reduce plus a
// separate optimizer pass, and no NCCL kernel competes with the persistent conv kernels for SMs.  The caller
// brackets the launch with two device-side barriers (all gradients final / all weights delivered).
constexpr int kMaxPeers = 8;
struct PeerBufs {
    float4* theta[kMaxPeers];
    const float4* grad[kMaxPeers];
};

template <int kKind>   // 0 = Adam (p0 = alpha_t, p1 = beta1, p2 = beta2), 1 = RMSprop (p0 = lr, p1 = rho)
__global__ void opt_flat_dp_kernel(const __grid_constant__ PeerBufs pb, float4* __restrict__ m, float4* __restrict__ v,
                                   long long lo4, long long hi4, int world, int rank, float p0, float p1, float p2,
                                   float eps, const float* __restrict__ alpha_tab, const int* __restrict__ t_dev,
                                   int tab_len) {
    // CUDA-graph replays (Adam): alpha_t from the host-computed table, indexed by the device-side step count
    if (kKind == 0 && alpha_tab != nullptr) p0 = alpha_tab[min(max(*t_dev, 1), tab_len) - 1];
    for (long long i = lo4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < hi4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int r = 0; r < world; ++r) {
            const float4 q = __ldcv(pb.grad[r] + i);      // not cached: written by a peer's kernels this step
            g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
        }
        float4 tt = pb.theta[rank][i];
        if (kKind == 0) {
            const float ob1 = 1.f - p1, ob2 = 1.f - p2;
            float4 mm = m[i], vv = v[i];
            adam_update(tt.x, mm.x, vv.x, g.x, p0, p1, p2, ob1, ob2, eps);
            adam_update(tt.y, mm.y, vv.y, g.y, p0, p1, p2, ob1, ob2, eps);
            adam_update(tt.z, mm.z, vv.z, g.z, p0, p1, p2, ob1, ob2, eps);
            adam_update(tt.w, mm.w, vv.w, g.w, p0, p1, p2, ob1, ob2, eps);
            m[i] = mm; v[i] = vv;
        } else {
            const float orho = 1.f - p1;
            float4 rr = m[i];
            rmsprop_update(tt.x, rr.x, g.x, p0, p1, orho, eps);
            rmsprop_update(tt.y, rr.y, g.y, p0, p1, orho, eps);
            rmsprop_update(tt.z, rr.z, g.z, p0, p1, orho, eps);
            rmsprop_update(tt.w, rr.w, g.w, p0, p1, orho, eps);
            m[i] = rr;
        }
#pragma unroll 1
        for (int r = 0; r < world; ++r) pb.theta[r][i] = tt;
    }
}

// ------------------------------------------------------------------ Grad-CAM
// One block of up to 1024 threads per frame.  Eight lanes share a pixel (8 x 16 B = one 128-byte line per load, four
// pixels per warp load: fully coalesced), each lane contracts its channels against alpha[] in shared memory and a
// 3-step shuffle finishes the pixel.  (v1: 256 threads, one pixel per warp at a time, a 5-step reduction per 512
// bytes loaded -> 27 % of HBM bandwidth.)  heat (H*W floats) lives in dynamic shared memory.
__global__ void __launch_bounds__(1024)
gradcam_kernel(const uint4* __restrict__ a, const float* __restrict__ probs, const float* __restrict__ w,
               float* __restrict__ heat_out, int H, int W, int C, int n, int out_h, int out_w) {
    extern __shared__ float sm[];
    float* alpha = sm;            // [C]
    float* heat = sm + C;         // [H*W]
    __shared__ float wmax[32];
    const int frame = blockIdx.x;
    const int HW = H * W, C8 = C >> 3;
    const float* p = probs + static_cast<size_t>(frame) * n;
    int cls = 0;
    float pc = p[0];
    for (int j = 1; j < n; ++j)
        if (p[j] > pc) { pc = p[j]; cls = j; }
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < n; ++j) s += pc * ((j == cls ? 1.f : 0.f) - p[j]) * w[k * n + j];
        alpha[k] = s / static_cast<float>(HW);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4* base = a + static_cast<size_t>(frame) * HW * C8;
    float lmax = 0.f;
    // 8 lanes x 16 B = one 128-byte line of one pixel: a warp load touches 4 pixels x 1 full line (coalesced)
    const int sub = lane >> 3, ch8 = lane & 7;
    const int nwarps = blockDim.x >> 5;
    for (int px0 = warp * 4; px0 < HW; px0 += nwarps * 4) {
        const int px = px0 + sub;
        const bool ok = px < HW;
        const uint4* src = base + static_cast<size_t>(ok ? px : 0) * C8;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
        for (int c = ch8; c < C8; c += 8) {
            const uint4 v = __ldg(src + c);
            const float4 a0 = *reinterpret_cast<const float4*>(alpha + c * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(alpha + c * 8 + 4);
            s0 += a0.x * bf16_lo(v.x) + a0.y * bf16_hi(v.x) + a0.z * bf16_lo(v.y) + a0.w * bf16_hi(v.y);
            s1 += a1.x * bf16_lo(v.z) + a1.y * bf16_hi(v.z) + a1.z * bf16_lo(v.w) + a1.w * bf16_hi(v.w);
        }
        float s = s0 + s1;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s = fmaxf(s / static_cast<float>(C), 0.f);
        if (ok && ch8 == 0) heat[px] = s;
        if (ok) lmax = fmaxf(lmax, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0) wmax[warp] = lmax;
    __syncthreads();
    float mx = 0.f;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) mx = fmaxf(mx, wmax[i]);
    const float inv = mx > 0.f ? 1.f / mx : 0.f;
    float* out = heat_out + static_cast<size_t>(frame) * out_h * out_w;
    const float sy = static_cast<float>(H) / out_h, sx = static_cast<float>(W) / out_w;
    for (int i = threadIdx.x; i < out_h * out_w; i += blockDim.x) {
        const int oy = i / out_w, ox = i - oy * out_w;
        // cv2.resize INTER_LINEAR: half-pixel centres, edge replicate
        const float fy = (oy + 0.5f) * sy - 0.5f, fx = (ox + 0.5f) * sx - 0.5f;
        const int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
        const float wy = fy - y0, wx = fx - x0;
        const int y0c = min(max(y0, 0), H - 1), y1c = min(max(y0 + 1, 0), H - 1);
        const int x0c = min(max(x0, 0), W - 1), x1c = min(max(x0 + 1, 0), W - 1);
        const float top = heat[y0c * W + x0c] * (1.f - wx) + heat[y0c * W + x1c] * wx;
        const float bot = heat[y1c * W + x0c] * (1.f - wx) + heat[y1c * W + x1c] * wx;
        out[i] = (top * (1.f - wy) + bot * wy) * inv;
    }
}

__global__ void __launch_bounds__(256)
parts_reduce_kernel(const float* __restrict__ parts, float* __restrict__ out, int C, int accumulate) {
    __shared__ float4 red[64][4];
    parts_reduce_block(parts + ABL_PARTS_HEADER, parts_rows(parts), C, out, accumulate, blockIdx.x, red);
}

static inline int grid_for(long long work_items, int threads, int waves) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = static_cast<long long>(sm_count()) * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace abl

using namespace abl;

extern "C" int abl_prep_conv_weights(const float* w_hwio, void* w_fwd, void* w_dgrad, int Cin, int Cout,
                                     void* stream) {
    if (!w_hwio || (!w_fwd && !w_dgrad)) { set_error("abl_prep_conv_weights: null pointer"); return ABL_ERR_NULL; }
    if (Cin <= 0 || Cout <= 0) { set_error("abl_prep_conv_weights: bad shape"); return ABL_ERR_SHAPE; }
    prep_conv_weights_kernel<<<grid_for(9LL * Cin * Cout, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w_hwio, static_cast<__nv_bfloat16*>(w_fwd), static_cast<__nv_bfloat16*>(w_dgrad), Cin, Cout);
    return check_launch();
}

extern "C" int abl_prep_conv_weights_multi(const float* const* w_hwio, void* const* w_fwd, void* const* w_dgrad,
                                           const int* cin, const int* cout, int n_layers, void* stream) {
    if (!w_hwio || !w_fwd || !w_dgrad || !cin || !cout) { set_error("abl_prep_conv_weights_multi: null pointer"); return ABL_ERR_NULL; }
    if (n_layers <= 0 || n_layers > kMaxPrepLayers) { set_error("abl_prep_conv_weights_multi: 1..16 layers"); return ABL_ERR_SHAPE; }
    PrepMulti m;
    int tiles = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (!w_hwio[l] || (!w_fwd[l] && !w_dgrad[l]) || cin[l] <= 0 || cout[l] <= 0 || (cin[l] & 31) || (cout[l] & 31)) {
            set_error("abl_prep_conv_weights_multi: bad layer entry (Cin, Cout multiples of 32)"); return ABL_ERR_SHAPE;
        }
        m.w[l] = w_hwio[l];
        m.w_fwd[l] = static_cast<__nv_bfloat16*>(w_fwd[l]);
        m.w_dgrad[l] = static_cast<__nv_bfloat16*>(w_dgrad[l]);
        m.cin[l] = cin[l]; m.cout[l] = cout[l];
        tiles += 9 * (cin[l] >> 5) * (cout[l] >> 5);
        m.tile_end[l] = tiles;
    }
    m.n_layers = n_layers;
    prep_conv_weights_multi_kernel<<<tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(m);
    return check_launch();
}

extern "C" int abl_bias_grad(const void* dz, float* db_parts, long long db_parts_bytes, long long rows, int C, void* stream);

extern "C" long long abl_parts_bytes(int C) {
    if (C <= 0) return -1;
    return (static_cast<long long>(ABL_PARTS_HEADER) + static_cast<long long>(sm_count()) * kPartWaves * C) * 4;
}

extern "C" int abl_parts_reduce(const float* parts, float* out, int C, int accumulate, void* stream) {
    if (!parts || !out) { set_error("abl_parts_reduce: null pointer"); return ABL_ERR_NULL; }
    if (C <= 0 || (C & 3) || ((reinterpret_cast<uintptr_t>(parts) | reinterpret_cast<uintptr_t>(out)) & 15)) {
        set_error("abl_parts_reduce: C must be a multiple of 4 and the buffers 16-byte aligned"); return ABL_ERR_ALIGN;
    }
    parts_reduce_kernel<<<(C / 4 + 3) / 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(parts, out, C, accumulate);
    return check_launch();
}

extern "C" int abl_maxpool_bwd_relu(const void* dp, const void* a, const void* pooled, void* dz, float* db_parts,
                                    long long db_parts_bytes, int B, int H, int W, int C, void* stream) {
    if (!dp || !a || !pooled || !dz) { set_error("abl_maxpool_bwd_relu: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || H < 2 || W < 2 || (C & 7)) { set_error("abl_maxpool_bwd_relu: need H, W >= 2 and C % 8 == 0"); return ABL_ERR_SHAPE; }
    if (((H | W) & 1) && cudaMemsetAsync(dz, 0, static_cast<size_t>(B) * H * W * C * 2, static_cast<cudaStream_t>(stream)) != cudaSuccess)
        return check_launch();      // (see abl_maxpool_bwd_code)
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
    const bool fuse = db_parts != nullptr && (256 % (C / 8)) == 0;      // a thread keeps one channel chunk
    const int grid = grid_for(total, 256, fuse ? kPartWaves : 16);
    if (fuse) { int r = parts_check(db_parts, db_parts_bytes, grid, C, "abl_maxpool_bwd_relu"); if (r != ABL_OK) return r; }
    maxpool_bwd_relu_kernel<<<grid, 256, fuse ? 256 * 8 * sizeof(float) : 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(dp), static_cast<const uint4*>(a), static_cast<const uint4*>(pooled),
        static_cast<uint4*>(dz), B, H, W, C / 8, fuse ? db_parts : nullptr);
    int r = check_launch();
    if (r == ABL_OK && db_parts != nullptr && !fuse)
        r = abl_bias_grad(dz, db_parts, db_parts_bytes, static_cast<long long>(B) * H * W, C, stream);
    return r;
}

extern "C" int abl_maxpool_bwd_code(const void* dp, const unsigned short* pool_code, void* dz, float* db_parts,
                                    long long db_parts_bytes, int B, int H, int W, int C, void* stream) {
    if (!dp || !pool_code || !dz) { set_error("abl_maxpool_bwd_code: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || H < 2 || W < 2 || (C & 7)) { set_error("abl_maxpool_bwd_code: need H, W >= 2 and C % 8 == 0"); return ABL_ERR_SHAPE; }
    // odd H or W (Keras 'valid' pooling drops the last row / column): those elements belong to no window, their
    // gradient is zero -- the kernel only writes whole windows, so the tensor is cleared first
    if (((H | W) & 1) && cudaMemsetAsync(dz, 0, static_cast<size_t>(B) * H * W * C * 2, static_cast<cudaStream_t>(stream)) != cudaSuccess)
        return check_launch();
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
    const bool fuse = db_parts != nullptr && (256 % (C / 8)) == 0;
    const int grid = grid_for(total, 256, fuse ? kPartWaves : 16);
    if (fuse) { int r = parts_check(db_parts, db_parts_bytes, grid, C, "abl_maxpool_bwd_code"); if (r != ABL_OK) return r; }
    maxpool_bwd_code_kernel<<<grid, 256, fuse ? 256 * 8 * sizeof(float) : 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(dp), pool_code, static_cast<uint4*>(dz), B, H, W, C / 8, fuse ? db_parts : nullptr);
    int r = check_launch();
    if (r == ABL_OK && db_parts != nullptr && !fuse)
        r = abl_bias_grad(dz, db_parts, db_parts_bytes, static_cast<long long>(B) * H * W, C, stream);
    return r;
}

extern "C" int abl_gap_fwd(const void* a, float* gap, int B, int HW, int C, void* stream) {
    if (!a || !gap) { set_error("abl_gap_fwd: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || HW <= 0 || (C % 64) != 0) { set_error("abl_gap_fwd: need C % 64 == 0"); return ABL_ERR_SHAPE; }
    gap_fwd_kernel<<<dim3(B, C / 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(a), gap,
                                                                                  HW, C);
    return check_launch();
}

extern "C" int abl_gap_reduce(const float* gap_part, float* gap, int B, int parts, int C, int HW, void* stream) {
    if (!gap_part || !gap) { set_error("abl_gap_reduce: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || parts <= 0 || C <= 0 || (C & 3) || HW <= 0) { set_error("abl_gap_reduce: bad shape (C % 4 == 0)"); return ABL_ERR_SHAPE; }
    gap_reduce_kernel<<<grid_for(static_cast<long long>(B) * (C / 4), 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(gap_part), reinterpret_cast<float4*>(gap), B, parts, C / 4,
        1.f / static_cast<float>(HW));
    return check_launch();
}

extern "C" int abl_gap_bwd_relu(const float* dgap, const void* a, void* dz, float* db_parts, long long db_parts_bytes,
                                int B, int HW, int C, void* stream) {
    if (!dgap || !a || !dz) { set_error("abl_gap_bwd_relu: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || HW <= 0 || (C & 7)) { set_error("abl_gap_bwd_relu: need C % 8 == 0"); return ABL_ERR_SHAPE; }
    const long long total = static_cast<long long>(B) * HW * (C / 8);
    const bool fuse = db_parts != nullptr && (256 % (C / 8)) == 0;
    const int grid = grid_for(total, 256, fuse ? kPartWaves : 16);
    if (fuse) { int r = parts_check(db_parts, db_parts_bytes, grid, C, "abl_gap_bwd_relu"); if (r != ABL_OK) return r; }
    gap_bwd_relu_kernel<<<grid, 256, fuse ? 256 * 8 * sizeof(float) : 0, static_cast<cudaStream_t>(stream)>>>(
        dgap, static_cast<const uint4*>(a), static_cast<uint4*>(dz), B, HW, C / 8, 1.f / static_cast<float>(HW),
        fuse ? db_parts : nullptr);
    int r = check_launch();
    if (r == ABL_OK && db_parts != nullptr && !fuse)
        r = abl_bias_grad(dz, db_parts, db_parts_bytes, static_cast<long long>(B) * HW, C, stream);
    return r;
}

extern "C" int abl_gap_bwd_bits(const float* dgap, const unsigned int* relu_bits, void* dz, float* db_parts,
                                long long db_parts_bytes, int B, int HW, int C, void* stream) {
    if (!dgap || !relu_bits || !dz) { set_error("abl_gap_bwd_bits: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || HW <= 0 || (C & 31)) { set_error("abl_gap_bwd_bits: need C % 32 == 0"); return ABL_ERR_SHAPE; }
    const bool fuse = db_parts != nullptr && (256 % (C / 8)) == 0;
    // Work items = frames x 64-pixel slices of a frame (a thread sees <= 255 pixels of a slice: byte-lane counters),
    // strided over a grid of exactly the blocks that are resident at once.  (The first version launched 8 blocks per SM
    // regardless of the kernel's registers: 5-6 fit, so a second, mostly empty wave ran -- ncu: 1.6 waves.)
    const int C8 = C / 8, rgs = 256 / C8 > 0 ? 256 / C8 : 1;
    const size_t dyn_smem = fuse ? 256 * 8 * sizeof(float) : 0;
    static int occ_cache[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ = occ_cache[dev & 63][fuse ? 1 : 0];
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gap_bwd_bits_kernel, 256, dyn_smem) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            occ = 1;
        }
        if (occ > kPartWaves) occ = kPartWaves;                   // (parts buffers hold kPartWaves rows per SM)
    }
    int slice_px = 8 * rgs;                                       // eight pixels per thread and item
    if (slice_px > HW) slice_px = HW;
    const int slices = (HW + slice_px - 1) / slice_px;
    const long long items = static_cast<long long>(B) * slices;
    const long long cap = static_cast<long long>(sm_count()) * occ;
    const int grid = static_cast<int>(items < cap ? items : cap);
    if (fuse) { int r = parts_check(db_parts, db_parts_bytes, grid, C, "abl_gap_bwd_bits"); if (r != ABL_OK) return r; }
    gap_bwd_bits_kernel<<<grid, 256, dyn_smem, static_cast<cudaStream_t>(stream)>>>(
        dgap, reinterpret_cast<const unsigned char*>(relu_bits), static_cast<uint4*>(dz), B, HW, C8,
        1.f / static_cast<float>(HW), slices, slice_px, fuse ? db_parts : nullptr);
    int r = check_launch();
    if (r == ABL_OK && db_parts != nullptr && !fuse)
        r = abl_bias_grad(dz, db_parts, db_parts_bytes, static_cast<long long>(B) * HW, C, stream);
    return r;
}

static int head_fwd_impl(float* gap, const float* gap_part, int gap_parts, int gap_hw, const float* w, const float* b,
                         float* logits, float* probs,
                         const float* y_onehot, const float* class_weight, float inv_batch, float* loss_sum,
                         float* dlogits, unsigned char* keep_mask, int keep_mask_in, int training, float dropout,
                         unsigned long long seed, unsigned int step, unsigned int rank, int B, int C, int n,
                         unsigned int* step_dev, int loss_accumulate, float* loss_ws, void* stream) {
    if (!gap || !w || !b || !logits || !probs) { set_error("abl_head_fwd: null pointer"); return ABL_ERR_NULL; }
    if (loss_sum && y_onehot && !loss_ws) { set_error("abl_head_fwd: the loss needs loss_ws (abl_head_loss_ws_bytes)"); return ABL_ERR_NULL; }
    if (training && !keep_mask) { set_error("abl_head_fwd: training needs a keep_mask buffer"); return ABL_ERR_NULL; }
    if (B <= 0 || C <= 0 || n <= 0 || n > kMaxClasses) { set_error("abl_head_fwd: need 1 <= n <= 8"); return ABL_ERR_SHAPE; }
    if (training && !(dropout >= 0.f && dropout < 1.f)) { set_error("abl_head_fwd: dropout must be in [0,1)"); return ABL_ERR_SHAPE; }
    if (gap_part && (gap_parts <= 0 || gap_hw <= 0)) { set_error("abl_head_fwd: gap_part needs gap_parts > 0 and gap_hw > 0"); return ABL_ERR_SHAPE; }
    head_fwd_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        gap, gap_part, gap_parts, gap_part ? 1.f / static_cast<float>(gap_hw) : 0.f, w, b, logits, probs, y_onehot, class_weight, inv_batch, loss_sum, dlogits, keep_mask, keep_mask_in,
        training, training ? dropout : 0.f, seed, step, rank, B, C, n, step_dev, loss_ws, loss_accumulate);
    return check_launch();
}

extern "C" long long abl_head_loss_ws_bytes(int B) {
    if (B <= 0) return -1;
    return (static_cast<long long>(ABL_PARTS_HEADER) + (B + 3) / 4) * 4;
}

extern "C" int abl_head_fwd(float* gap, const float* w, const float* b, float* logits, float* probs,
                            const float* y_onehot, const float* class_weight, float inv_batch, float* loss_sum,
                            float* dlogits, unsigned char* keep_mask, int keep_mask_in, int training, float dropout,
                            unsigned long long seed, unsigned int step, unsigned int rank, int B, int C, int n,
                            int loss_accumulate, float* loss_ws, const float* gap_part, int gap_parts, int gap_hw,
                            void* stream) {
    return head_fwd_impl(gap, gap_part, gap_parts, gap_hw, w, b, logits, probs, y_onehot, class_weight, inv_batch, loss_sum, dlogits, keep_mask,
                         keep_mask_in, training, dropout, seed, step, rank, B, C, n, nullptr, loss_accumulate, loss_ws, stream);
}

extern "C" int abl_head_fwd_stepdev(float* gap, const float* w, const float* b, float* logits, float* probs,
                                    const float* y_onehot, const float* class_weight, float inv_batch, float* loss_sum,
                                    float* dlogits, unsigned char* keep_mask, int training, float dropout,
                                    unsigned long long seed, unsigned int* step_dev, unsigned int rank, int B,
                                    int C, int n, int loss_accumulate, float* loss_ws, const float* gap_part,
                                    int gap_parts, int gap_hw, void* stream) {
    if (!step_dev) { set_error("abl_head_fwd_stepdev: step_dev is null"); return ABL_ERR_NULL; }
    if (!loss_sum || !y_onehot) { set_error("abl_head_fwd_stepdev: needs y_onehot and loss_sum (the counters advance with the loss)"); return ABL_ERR_NULL; }
    return head_fwd_impl(gap, gap_part, gap_parts, gap_hw, w, b, logits, probs, y_onehot, class_weight, inv_batch, loss_sum,
                         dlogits, keep_mask, 0, training, dropout, seed, 0u, rank, B, C, n, step_dev, loss_accumulate, loss_ws,
                         stream);
}

extern "C" int abl_head_bwd(const float* gap, const float* w, const float* dlogits, const unsigned char* keep_mask,
                            float dropout, float* dw, float* db, float* dgap, int B, int C, int n, void* stream) {
    if (!gap || !w || !dlogits || !dw || !db || !dgap) { set_error("abl_head_bwd: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || C <= 0 || n <= 0 || n > kMaxClasses) { set_error("abl_head_bwd: need 1 <= n <= 8"); return ABL_ERR_SHAPE; }
    head_bwd_kernel<<<(C + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        gap, w, dlogits, keep_mask, keep_mask ? dropout : 0.f, dw, db, dgap, B, C, n);
    return check_launch();
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int adam_flat_impl(float* theta, const float* g, float* m, float* v, long long n, float alpha_t, float b1, float b2,
                          float eps, const float* alpha_tab, const int* t_dev, int tab_len, void* stream);

extern "C" int abl_adam_flat(float* theta, const float* g, float* m, float* v, long long n, float alpha_t, float b1,
                             float b2, float eps, void* stream) {
    return adam_flat_impl(theta, g, m, v, n, alpha_t, b1, b2, eps, nullptr, nullptr, 0, stream);
}

extern "C" int abl_adam_flat_tab(float* theta, const float* g, float* m, float* v, long long n, const float* alpha_tab,
                                 const int* t_dev, int tab_len, float b1, float b2, float eps, void* stream) {
    if (!alpha_tab || !t_dev || tab_len <= 0) { set_error("abl_adam_flat_tab: table / step pointer missing"); return ABL_ERR_NULL; }
    return adam_flat_impl(theta, g, m, v, n, 0.f, b1, b2, eps, alpha_tab, t_dev, tab_len, stream);
}

static int adam_flat_impl(float* theta, const float* g, float* m, float* v, long long n, float alpha_t, float b1, float b2,
                          float eps, const float* alpha_tab, const int* t_dev, int tab_len, void* stream) {
    if (!theta || !g || !m || !v) { set_error("abl_adam_flat: null pointer"); return ABL_ERR_NULL; }
    if (n <= 0) return ABL_OK;
    if (!aligned16(theta) || !aligned16(g) || !aligned16(m) || !aligned16(v)) {
        set_error("abl_adam_flat: buffers must be 16-byte aligned"); return ABL_ERR_ALIGN;
    }
    const long long n4 = n / 4;
    const int tail = static_cast<int>(n - n4 * 4);
    adam_flat_kernel<<<grid_for(n4 > 0 ? n4 : 1, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4*>(theta), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), n4, theta + n4 * 4, g + n4 * 4, m + n4 * 4, v + n4 * 4, tail, alpha_t, b1, b2, eps,
        alpha_tab, t_dev, tab_len);
    return check_launch();
}

extern "C" int abl_rmsprop_flat(float* theta, const float* g, float* rms, long long n, float lr, float rho, float eps,
                                void* stream) {
    if (!theta || !g || !rms) { set_error("abl_rmsprop_flat: null pointer"); return ABL_ERR_NULL; }
    if (n <= 0) return ABL_OK;
    if (!aligned16(theta) || !aligned16(g) || !aligned16(rms)) {
        set_error("abl_rmsprop_flat: buffers must be 16-byte aligned"); return ABL_ERR_ALIGN;
    }
    const long long n4 = n / 4;
    const int tail = static_cast<int>(n - n4 * 4);
    rmsprop_flat_kernel<<<grid_for(n4 > 0 ? n4 : 1, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4*>(theta), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(rms), n4,
        theta + n4 * 4, g + n4 * 4, rms + n4 * 4, tail, lr, rho, eps);
    return check_launch();
}

extern "C" int abl_resize_nearest_u8(const unsigned char* src, unsigned char* dst, int B, int h, int w, int H, int W,
                                     void* stream) {
    if (!src || !dst) { set_error("abl_resize_nearest_u8: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) { set_error("abl_resize_nearest_u8: bad shape"); return ABL_ERR_SHAPE; }
    const long long total = static_cast<long long>(B) * H * W;
    resize_nearest_u8_kernel<<<grid_for(total, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, B, h, w, H, W);
    return check_launch();
}

extern "C" int abl_clip_aggregate(const float* probs, const long long* clip_offsets, int n_clips, int n_classes,
                                  int b_idx, float threshold, int window, float* mean_out, int* run_out,
                                  float* win_out, void* stream) {
    if (!probs || !clip_offsets || !mean_out || !run_out || !win_out) { set_error("abl_clip_aggregate: null pointer"); return ABL_ERR_NULL; }
    if (n_clips <= 0 || n_classes <= 0 || b_idx < 0 || b_idx >= n_classes || window < 0) {
        set_error("abl_clip_aggregate: bad arguments"); return ABL_ERR_SHAPE;
    }
    clip_aggregate_kernel<<<(n_clips + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        probs, clip_offsets, n_clips, n_classes, b_idx, threshold, window, mean_out, run_out, win_out);
    return check_launch();
}

extern "C" int abl_opt_flat_dp(int kind, float* const* theta_peers, const float* const* grad_peers, float* m_or_rms,
                               float* v, long long offset, long long n, int rank, int world, float lr_or_alpha,
                               float b1_or_rho, float b2, float eps, const float* alpha_tab, const int* t_dev,
                               int tab_len, void* stream) {
    if (alpha_tab && (!t_dev || tab_len <= 0)) { set_error("abl_opt_flat_dp: alpha_tab needs t_dev and tab_len"); return ABL_ERR_NULL; }
    if (!theta_peers || !grad_peers || !m_or_rms || (kind == 0 && !v)) { set_error("abl_opt_flat_dp: null pointer"); return ABL_ERR_NULL; }
    if (kind != 0 && kind != 1) { set_error("abl_opt_flat_dp: kind must be 0 (Adam) or 1 (RMSprop)"); return ABL_ERR_SHAPE; }
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) { set_error("abl_opt_flat_dp: need 1 <= world <= 8"); return ABL_ERR_SHAPE; }
    if ((offset & 3) || (n & 3) || n < 0) { set_error("abl_opt_flat_dp: offset and n must be multiples of 4 floats"); return ABL_ERR_SHAPE; }
    if (n == 0) return ABL_OK;
    PeerBufs pb;
    for (int r = 0; r < kMaxPeers; ++r) { pb.theta[r] = nullptr; pb.grad[r] = nullptr; }
    for (int r = 0; r < world; ++r) {
        if (!theta_peers[r] || !grad_peers[r] || !aligned16(theta_peers[r]) || !aligned16(grad_peers[r])) {
            set_error("abl_opt_flat_dp: peer buffers must be non-null and 16-byte aligned"); return ABL_ERR_ALIGN;
        }
        pb.theta[r] = reinterpret_cast<float4*>(theta_peers[r] + offset);
        pb.grad[r] = reinterpret_cast<const float4*>(grad_peers[r] + offset);
    }
    const long long n4 = n / 4;
    const long long lo4 = n4 * rank / world, hi4 = n4 * (rank + 1) / world;     // this rank's shard of the run
    if (hi4 <= lo4) return ABL_OK;
    float4* m4 = reinterpret_cast<float4*>(m_or_rms + offset);
    float4* v4 = v ? reinterpret_cast<float4*>(v + offset) : nullptr;
    const int grid = grid_for(hi4 - lo4, 256, 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (kind == 0)
        opt_flat_dp_kernel<0><<<grid, 256, 0, st>>>(pb, m4, v4, lo4, hi4, world, rank, lr_or_alpha, b1_or_rho, b2, eps,
                                                    alpha_tab, t_dev, tab_len);
    else
        opt_flat_dp_kernel<1><<<grid, 256, 0, st>>>(pb, m4, v4, lo4, hi4, world, rank, lr_or_alpha, b1_or_rho, b2, eps,
                                                    nullptr, nullptr, 0);
    return check_launch();
}

extern "C" int abl_gradcam(const void* a, const float* probs, const float* w, float* heat, int B, int H, int W, int C,
                           int n, int out_h, int out_w, void* stream) {
    if (!a || !probs || !w || !heat) { set_error("abl_gradcam: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || H <= 0 || W <= 0 || (C & 7) || n <= 0 || n > kMaxClasses || out_h <= 0 || out_w <= 0) {
        set_error("abl_gradcam: bad shape (need C % 8 == 0, 1 <= n <= 8)"); return ABL_ERR_SHAPE;
    }
    const size_t smem = (static_cast<size_t>(C) + static_cast<size_t>(H) * W) * sizeof(float);
    if (smem > 200 * 1024) { set_error("abl_gradcam: activation map too large for shared memory"); return ABL_ERR_SHAPE; }
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(gradcam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return check_launch();
    }
    int threads = ((H * W + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    if (threads < 64) threads = 64;
    gradcam_kernel<<<B, threads, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(a), probs, w, heat,
                                                                           H, W, C, n, out_h, out_w);
    return check_launch();
}
