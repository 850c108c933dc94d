// Either side of the network: training-time data augmentation in front of block1_conv1 and
// the Keras metric accumulators behind the softmax head.  Both are small HBM/latency-bound kernels.
//
// augment_kernel: the reference's `tf.keras.Sequential([RandomZoom, RandomTranslation, RandomRotation, RandomFlip,
// RandomBrightness])` (src/data/preprocessor.py:25-32 with its own RandomBrightness layer at 83-104; applied to every
// training batch at preprocessor.py:56-57) followed
// by `vgg16.preprocess_input` (preprocessor.py:60-63).  The three geometric layers are three successive
// ImageProjectiveTransformV3 resamplings (bilinear, fill_mode='constant', fill value 0), each cropping to the frame;
// the kernel reproduces them one after the other -- NOT as one composed warp, which would interpolate once and keep
// content that an earlier stage pushed out of the frame.  One CTA owns one channel of one frame: a 128x128 fp32
// plane is 64 KB, so two ping-pong planes live in the CTA's shared memory (a B200 SM has 227 KB) and the frame is read
// from HBM once and written once.  Larger frames ping-pong through a caller-provided workspace instead.
// Every product/sum is rounded separately (no FMA contraction) so that the result is bit-identical to the NumPy
// restatement of TF's kernel in oracle/augment.py.
#include <stdint.h>

#include "host_util.h"

namespace abl {

constexpr int kAugParams = 20;     // per frame: 3 x [a0 a1 a2 b0 b1 b2] (zoom, translation, rotation), flip, delta

// One ImageProjectiveTransformV3 pass (BILINEAR, CONSTANT) over an H x W plane: out(oy,ox) = in(b.(ox,oy,1), a.(ox,oy,1)).
// TF's read_with_fill_value (image_ops.h) returns the fill value (0) for a neighbour outside the frame; the four
// neighbours share two row and two column validity tests, made on the integer coordinates (floor() is integral, and
// float->int conversion saturates, so they are the float tests of the oracle).  The kernel is instruction-bound
// (~60 instructions per pixel and pass): pixel coordinates advance incrementally instead of by division.
template <int kStride, typename T, typename Store>
__device__ __forceinline__ void aug_stage(const T* __restrict__ img, int H, int W, const float* m, Store store) {
    const float a0 = m[0], a1 = m[1], a2 = m[2], b0 = m[3], b1 = m[4], b2 = m[5];
    const int step_y = blockDim.x / W, step_x = blockDim.x - step_y * W;
    int oy = threadIdx.x / W, ox = threadIdx.x - oy * W;
    const int row = W * kStride;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const float fx = static_cast<float>(ox), fy = static_cast<float>(oy);
        const float x = __fadd_rn(__fadd_rn(__fmul_rn(a0, fx), __fmul_rn(a1, fy)), a2);
        const float y = __fadd_rn(__fadd_rn(__fmul_rn(b0, fx), __fmul_rn(b1, fy)), b2);
        const float xf = floorf(x), yf = floorf(y);
        const float xc = __fadd_rn(xf, 1.f), yc = __fadd_rn(yf, 1.f);
        const float wx0 = __fsub_rn(xc, x), wx1 = __fsub_rn(x, xf);
        const unsigned ix = static_cast<unsigned>(__float2int_rz(xf)), iy = static_cast<unsigned>(__float2int_rz(yf));
        const bool vx0 = ix < static_cast<unsigned>(W), vx1 = ix + 1u < static_cast<unsigned>(W);
        const bool vy0 = iy < static_cast<unsigned>(H), vy1 = iy + 1u < static_cast<unsigned>(H);
        // the offset wraps for coordinates far outside; dereferenced only where both tests of that neighbour passed
        const T* p = img + static_cast<int>((iy * static_cast<unsigned>(W) + ix) * static_cast<unsigned>(kStride));
        const T* q = p + row;
        const float v00 = (vy0 && vx0) ? static_cast<float>(p[0]) : 0.f;
        const float v01 = (vy0 && vx1) ? static_cast<float>(p[kStride]) : 0.f;
        const float v10 = (vy1 && vx0) ? static_cast<float>(q[0]) : 0.f;
        const float v11 = (vy1 && vx1) ? static_cast<float>(q[kStride]) : 0.f;
        const float v_lo = __fadd_rn(__fmul_rn(wx0, v00), __fmul_rn(wx1, v01));
        const float v_hi = __fadd_rn(__fmul_rn(wx0, v10), __fmul_rn(wx1, v11));
        store(i, oy, ox, __fadd_rn(__fmul_rn(__fsub_rn(yc, y), v_lo), __fmul_rn(__fsub_rn(y, yf), v_hi)));
        ox += step_x;
        oy += step_y;
        if (ox >= W) { ox -= W; ++oy; }
    }
}

// kSmem: the two ping-pong planes are the CTA's dynamic shared memory (LDS/STS); otherwise a slice of the workspace.
template <typename TSrc, bool kSmem>
__global__ void __launch_bounds__(1024)
augment_kernel(const TSrc* __restrict__ src, const float* __restrict__ params, float* __restrict__ dst,
               float* __restrict__ workspace, int H, int W, int preprocess) {
    extern __shared__ float aug_smem[];
    const int n = blockIdx.x / 3, c = blockIdx.x % 3;
    const int plane = H * W;
    float* buf0 = kSmem ? aug_smem : workspace + (static_cast<long long>(blockIdx.x) * 2) * plane;
    float* buf1 = buf0 + plane;
    const float* p = params + static_cast<long long>(n) * kAugParams;
    const TSrc* img = src + static_cast<long long>(n) * plane * 3 + c;
    asm volatile("" : "+l"(img));          // keep the frame base in registers (else it is rebuilt for every load)
    // RandomZoom: frame (interleaved RGB in HBM) -> plane 0
    aug_stage<3>(img, H, W, p, [&](int i, int, int, float v) { buf0[i] = v; });
    __syncthreads();
    // RandomTranslation: plane 0 -> plane 1
    aug_stage<1>(buf0, H, W, p + 6, [&](int i, int, int, float v) { buf1[i] = v; });
    __syncthreads();
    // RandomRotation, then RandomFlip('horizontal') as a mirrored store, the reference's own RandomBrightness layer
    // (preprocessor.py:83-104: tf.image.adjust_brightness of a float image = "+ delta", no clipping) and
    // vgg16.preprocess_input (caffe mode: RGB -> BGR, minus the ImageNet channel mean, no scaling)
    const bool flip = p[18] != 0.f;
    const float delta = p[19];
    const float mean = preprocess ? (c == 0 ? 123.68f : (c == 1 ? 116.779f : 103.939f)) : 0.f;
    float* out = dst + static_cast<long long>(n) * plane * 3 + (preprocess ? 2 - c : c);
    asm volatile("" : "+l"(out));
    aug_stage<1>(buf1, H, W, p + 12, [&](int, int oy, int ox, float v) {
        v = __fsub_rn(__fadd_rn(v, delta), mean);
        out[(oy * W + (flip ? W - 1 - ox : ox)) * 3] = v;
    });
}

// ---------------------------------------------------------------------------------------------------- metrics
// Keras metric state for the list the reference builds at src/train.py:109-115 -- 'accuracy' (categorical), AUC (ROC,
// num_thresholds, labels and predictions flattened over the class axis), Precision / Recall per class_id at one
// threshold -- accumulated on the device from (y one-hot, probs) after every batch, so neither fit() nor evaluate()
// collects the probabilities or labels on the host; a single read of a few hundred counters per epoch replaces it.
// Counters (unsigned 64-bit; integer atomics, so the result does not depend on scheduling):
//   [0] frames  [1] argmax(y) == argmax(p)  [2+3c..] tp_c, fp_c, fn_c of class c at class_thr[c]
//   [2+3C + k]         number of positive (y != 0) flattened entries whose p exceeds exactly k of the AUC thresholds
//   [2+3C + (T+1) + k] the same for negative entries.        tp(t) = sum_{k > t} pos[k], fp(t) likewise.
__global__ void __launch_bounds__(256)
metrics_accumulate_kernel(const float* __restrict__ probs, const float* __restrict__ y, int B, int C,
                          const float* __restrict__ auc_thr, int T, const float* __restrict__ class_thr,
                          unsigned long long* __restrict__ counters) {
    extern __shared__ unsigned int met_smem[];        // [2 + 3C + 2(T+1)] counters, then T thresholds
    const int n_cnt = 2 + 3 * C + 2 * (T + 1);
    float* thr = reinterpret_cast<float*>(met_smem + n_cnt);
    for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) met_smem[i] = 0u;
    for (int i = threadIdx.x; i < T; i += blockDim.x) thr[i] = auc_thr[i];
    __syncthreads();
    unsigned int* hist = met_smem + 2 + 3 * C;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < B; n += gridDim.x * blockDim.x) {
        const float* pn = probs + static_cast<long long>(n) * C;
        const float* yn = y + static_cast<long long>(n) * C;
        int arg_p = 0, arg_y = 0;
        float best_p = pn[0], best_y = yn[0];
        for (int c = 0; c < C; ++c) {
            const float pc = pn[c], yc = yn[c];
            if (pc > best_p) { best_p = pc; arg_p = c; }          // first maximum wins, as tf.argmax
            if (yc > best_y) { best_y = yc; arg_y = c; }
            const bool pos = yc != 0.f;                            // Keras casts y_true to bool
            const bool hit = pc > class_thr[c];
            if (pos && hit) atomicAdd(&met_smem[2 + 3 * c], 1u);
            if (!pos && hit) atomicAdd(&met_smem[2 + 3 * c + 1], 1u);
            if (pos && !hit) atomicAdd(&met_smem[2 + 3 * c + 2], 1u);
            int lo = 0, hi = T;                                    // k = number of thresholds with thr < p (thr ascending)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (pc > thr[mid]) lo = mid + 1; else hi = mid;
            }
            atomicAdd(&hist[(pos ? 0 : T + 1) + lo], 1u);
        }
        atomicAdd(&met_smem[0], 1u);
        if (arg_p == arg_y) atomicAdd(&met_smem[1], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_cnt; i += blockDim.x)
        if (met_smem[i]) atomicAdd(&counters[i], static_cast<unsigned long long>(met_smem[i]));
}

static constexpr int kAugMaxSmem = 227 * 1024;

}  // namespace abl

using namespace abl;

extern "C" long long abl_augment_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return -1;
    const long long planes = 2LL * H * W * static_cast<long long>(sizeof(float));
    return planes <= kAugMaxSmem ? 0 : planes * 3 * B;
}

extern "C" int abl_augment_frames(const void* src, int src_is_u8, const float* params, float* dst, void* workspace,
                                  long long workspace_bytes, int B, int H, int W, int preprocess, void* stream) {
    if (!src || !params || !dst) { set_error("abl_augment_frames: null pointer"); return ABL_ERR_NULL; }
    if (B <= 0 || H <= 0 || W <= 0 || static_cast<long long>(H) * W > (1LL << 24)) {
        set_error("abl_augment_frames: bad shape"); return ABL_ERR_SHAPE;
    }
    const long long need = abl_augment_workspace_bytes(B, H, W);
    if (need > 0 && (!workspace || workspace_bytes < need)) {
        set_error("abl_augment_frames: frame planes exceed shared memory and the workspace is missing or too small");
        return ABL_ERR_SHAPE;
    }
    const size_t smem = need > 0 ? 0 : 2 * static_cast<size_t>(H) * W * sizeof(float);
    float* ws = need > 0 ? static_cast<float*>(workspace) : nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    const uint8_t* s8 = static_cast<const uint8_t*>(src);
    const float* s32 = static_cast<const float*>(src);
    if (need == 0) {
        e = src_is_u8 ? cudaFuncSetAttribute(augment_kernel<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAugMaxSmem)
                      : cudaFuncSetAttribute(augment_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAugMaxSmem);
        if (e == cudaSuccess) {
            if (src_is_u8) augment_kernel<uint8_t, true><<<B * 3, 1024, smem, st>>>(s8, params, dst, nullptr, H, W, preprocess);
            else augment_kernel<float, true><<<B * 3, 1024, smem, st>>>(s32, params, dst, nullptr, H, W, preprocess);
        }
    } else if (src_is_u8) {
        augment_kernel<uint8_t, false><<<B * 3, 1024, 0, st>>>(s8, params, dst, ws, H, W, preprocess);
    } else {
        augment_kernel<float, false><<<B * 3, 1024, 0, st>>>(s32, params, dst, ws, H, W, preprocess);
    }
    if (e != cudaSuccess) { set_error("abl_augment_frames: cudaFuncSetAttribute failed"); return ABL_ERR_CUDA; }
    return check_launch();
}

extern "C" int abl_metrics_counter_count(int n_classes, int num_thresholds) {
    if (n_classes <= 0 || num_thresholds <= 0) return -1;
    return 2 + 3 * n_classes + 2 * (num_thresholds + 1);
}

extern "C" int abl_metrics_accumulate(const float* probs, const float* y_true, int B, int n_classes,
                                      const float* auc_thresholds, int num_thresholds, const float* class_thresholds,
                                      unsigned long long* counters, void* stream) {
    if (B < 0 || n_classes <= 0 || n_classes > 1024 || num_thresholds <= 0 || num_thresholds > 4096) {
        set_error("abl_metrics_accumulate: bad shape"); return ABL_ERR_SHAPE;
    }
    if (B == 0) return ABL_OK;          // an empty batch adds nothing (its pointers may be null)
    if (!probs || !y_true || !auc_thresholds || !class_thresholds || !counters) {
        set_error("abl_metrics_accumulate: null pointer"); return ABL_ERR_NULL;
    }
    const int n_cnt = abl_metrics_counter_count(n_classes, num_thresholds);
    const size_t smem = static_cast<size_t>(n_cnt) * sizeof(unsigned int) + static_cast<size_t>(num_thresholds) * sizeof(float);
    if (smem > 48 * 1024) { set_error("abl_metrics_accumulate: n_classes x num_thresholds too large for one CTA's counters"); return ABL_ERR_SHAPE; }
    int grid = (B + 255) / 256;
    if (grid > 4 * sm_count()) grid = 4 * sm_count();
    metrics_accumulate_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        probs, y_true, B, n_classes, auc_thresholds, num_thresholds, class_thresholds, counters);
    return check_launch();
}
