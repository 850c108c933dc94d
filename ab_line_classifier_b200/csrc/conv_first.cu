// block1_conv1 (Cin = 3, K = 27): Keras Conv2D+BiasAdd+Relu on the caller's fp32 frame
// (reference: first VGG16 layer reused at src/models/models.py:409-413) and its filter gradient.
// K = 27 is not tensor-core shaped and the layer is HBM-bound (0.2 MB in, 2 MB out per frame), so it
// runs on the CUDA cores in full fp32 (no bf16 rounding of the +-150-range input), writing bf16.
#include "host_util.h"
#include "ptx.cuh"

namespace abl {

constexpr int kFirstCout = 64;

// thread = one output pixel x 64 channels; weights [27][64] + bias broadcast from shared memory.
__global__ void __launch_bounds__(256)
conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      __nv_bfloat16* __restrict__ out, int B, int H, int W) {
    __shared__ __align__(16) float ws[27 * kFirstCout];
    __shared__ __align__(16) float bs[kFirstCout];
    for (int i = threadIdx.x; i < 27 * kFirstCout; i += blockDim.x) ws[i] = w[i];
    if (threadIdx.x < kFirstCout) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const long long total = static_cast<long long>(B) * H * W;
    for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; pix < total;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int wq = static_cast<int>(pix % W);
        const int hq = static_cast<int>((pix / W) % H);
        const long long n = pix / (static_cast<long long>(W) * H);
        float xin[27];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hh = hq + kh - 1, ww = wq + kw - 1;
                const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
                const float* px = x + ((n * H + hh) * W + ww) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) xin[(kh * 3 + kw) * 3 + c] = ok ? __ldg(px + c) : 0.f;
            }
        uint4* dst = reinterpret_cast<uint4*>(out + pix * kFirstCout);
#pragma unroll
        for (int cb = 0; cb < kFirstCout; cb += 16) {
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = bs[cb + j];
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                const float4* wr = reinterpret_cast<const float4*>(ws + k * kFirstCout + cb);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w4 = wr[j];
                    acc[4 * j + 0] = fmaf(xin[k], w4.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(xin[k], w4.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(xin[k], w4.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(xin[k], w4.w, acc[4 * j + 3]);
                }
            }
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
            dst[cb / 8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[cb / 8 + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
    }
}

// Filter gradient: dw[k][co] += sum_pix patch[pix][k] * dz[pix][co], k = (kh*3+kw)*3+c.
// Persistent blocks; each block walks strips of 64 pixels of one image row, staging the 3-row x halo and the
// dz strip in shared memory.  Warp g owns channels [8g, 8g+8); lane k < 27 owns filter tap k, so a warp's dz
// read is one broadcast and its x reads are 27 distinct words.  Accumulators stay in registers for the whole
// kernel; one atomicAdd per (k, co) per block at the end.
constexpr int kStrip = 64;
__global__ void __launch_bounds__(256)
conv_first_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dz, float* __restrict__ dw,
                        float* __restrict__ db, int B, int H, int W) {
    __shared__ float xs[3][(kStrip + 2) * 3];
    __shared__ __align__(16) __nv_bfloat16 dzs[kStrip][kFirstCout];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int k = lane < 27 ? lane : 26;
    const int kh = k / 9, kw = (k / 3) % 3, kc = k % 3;
    float acc[8], bacc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j] = 0.f; bacc[j] = 0.f; }
    const int strips_per_row = W / kStrip;
    const long long nstrips = static_cast<long long>(B) * H * strips_per_row;
    for (long long s = blockIdx.x; s < nstrips; s += gridDim.x) {
        const int sw = static_cast<int>(s % strips_per_row);
        const int hq = static_cast<int>((s / strips_per_row) % H);
        const long long n = s / (static_cast<long long>(strips_per_row) * H);
        const int w0 = sw * kStrip;
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * (kStrip + 2) * 3; i += blockDim.x) {
            const int r = i / ((kStrip + 2) * 3), rem = i % ((kStrip + 2) * 3);
            const int ww = w0 - 1 + rem / 3, hh = hq - 1 + r;
            const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
            xs[r][rem] = ok ? __ldg(x + ((n * H + hh) * W + ww) * 3 + rem % 3) : 0.f;
        }
        const uint4* src = reinterpret_cast<const uint4*>(dz + ((n * H + hq) * W + w0) * kFirstCout);
        uint4* dsts = reinterpret_cast<uint4*>(&dzs[0][0]);
        for (int i = threadIdx.x; i < kStrip * kFirstCout / 8; i += blockDim.x) dsts[i] = __ldg(src + i);
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < kStrip; ++p) {
            const float xv = xs[kh][(p + kw) * 3 + kc];
            const uint4 d = *reinterpret_cast<const uint4*>(&dzs[p][g * 8]);
            const float dv[8] = {__uint_as_float(d.x << 16), __uint_as_float(d.x & 0xFFFF0000u),
                                 __uint_as_float(d.y << 16), __uint_as_float(d.y & 0xFFFF0000u),
                                 __uint_as_float(d.z << 16), __uint_as_float(d.z & 0xFFFF0000u),
                                 __uint_as_float(d.w << 16), __uint_as_float(d.w & 0xFFFF0000u)};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[j] = fmaf(xv, dv[j], acc[j]);
                bacc[j] += dv[j];
            }
        }
    }
    if (lane < 27) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw + k * kFirstCout + g * 8 + j, acc[j]);
    }
    if (lane == 31 && db != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(db + g * 8 + j, bacc[j]);
    }
}

}  // namespace abl

using namespace abl;

extern "C" int abl_conv_first_fwd(const float* x, const float* w_hwio, const float* bias, void* out, int B, int H,
                                  int W, int Cout, void* stream) {
    if (!x || !w_hwio || !bias || !out) { set_error("abl_conv_first_fwd: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || H <= 0 || W <= 0 || Cout != kFirstCout) { set_error("abl_conv_first_fwd: Cout must be 64"); return ABL_ERR_SHAPE; }
    if (reinterpret_cast<uintptr_t>(out) & 15) { set_error("abl_conv_first_fwd: out must be 16-byte aligned"); return ABL_ERR_ALIGN; }
    const long long total = static_cast<long long>(B) * H * W;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    conv_first_fwd_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w_hwio, bias, static_cast<__nv_bfloat16*>(out), B, H, W);
    return check_launch();
}

extern "C" int abl_conv_first_wgrad(const float* x, const void* dz, float* dw, float* db, int B, int H, int W,
                                    int Cout, void* stream) {
    if (!x || !dz || !dw) { set_error("abl_conv_first_wgrad: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || H <= 0 || W <= 0 || (W % kStrip) != 0 || Cout != kFirstCout) {
        set_error("abl_conv_first_wgrad: need Cout == 64 and W % 64 == 0"); return ABL_ERR_SHAPE;
    }
    if (reinterpret_cast<uintptr_t>(dz) & 15) { set_error("abl_conv_first_wgrad: dz must be 16-byte aligned"); return ABL_ERR_ALIGN; }
    const long long nstrips = static_cast<long long>(B) * H * (W / kStrip);
    long long blocks = static_cast<long long>(sm_count()) * 4;
    if (blocks > nstrips) blocks = nstrips;
    conv_first_wgrad_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<const __nv_bfloat16*>(dz), dw, db, B, H, W);
    return check_launch();
}
