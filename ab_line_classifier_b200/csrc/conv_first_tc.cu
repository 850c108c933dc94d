// block1_conv1 (Cin = 3, K = 27) forward and filter gradient on tcgen05 via an in-kernel im2col.
// Replaces Keras Conv2D+BiasAdd+Relu of the first VGG16 layer (reference src/models/models.py:409-413)
// and its Conv2DBackpropFilter/BiasAddGrad under model.fit.
//
// The layer is HBM-bound (0.2 MB in, 2 MB out per frame) and K = 27 is far too small for a TMA-fed
// implicit GEMM, so producer warps build the A operand themselves: one thread per output pixel gathers
// its 3x3x3 fp32 patch and writes ONE 128-byte K-major row (64 bf16 "K slots", SWIZZLE_128B) per k-block:
//     k-block 0: slots  0..26 = hi(x)   slots 27..53 = lo(x)   slot 54 = 1   slots 55..63 = 0
//     k-block 1: slots  0..26 = hi(x)   rest 0                                   (forward only)
// with hi(x) = bf16(x), lo(x) = bf16(x - hi(x)).  The forward B operand pairs them with hi(W) / lo(W):
//     y = hi(x)*hi(W) + lo(x)*hi(W) + hi(x)*lo(W) + b        (error ~2^-16: fp32-like, fp32 accumulate)
// so the +-150-range input is NOT rounded to bf16.  The same A tile, read MN-major (K = pixels), is the
// A operand of the filter gradient: D[slot][co] = sum_pix A[pix][slot] * dZ[pix][co], rows 0..26 + 27..53
// give dW, row 54 (the ones slot) gives the bias gradient.
// A tile = 128 consecutive pixels of the flattened (n,h,w) raster (contiguous 16 KiB of NHWC output).
#include <stdio.h>
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace abl {

constexpr int kF_Cout = 64;
constexpr int kF_TileM = 128;
constexpr int kF_ABlock = kF_TileM * 128;     // 16 KiB: [128 px][64 slots] bf16
constexpr int kF_Threads = 512;               // warps 0-3 control, 4-15 = 3 worker groups of 4 warps
constexpr int kF_Stages = 4;
constexpr int kF_ProdGroups = 3;              // worker group g owns the CTA's tiles g, g+3, g+6, ...: it builds the im2col
                                              // rows and (forward) drains the accumulator of its previous tile
constexpr int kF_TsCols = 48;                // tensor-memory columns of one A stage (96 bf16 slots per row)
constexpr int kF_AccBufs = 4;                 // forward TMEM accumulators (64 columns each)

struct FirstArgs {
    const float* x;        // [B,H,W,3] fp32 (already BGR / mean-subtracted), or
    const uint8_t* x8;     // [B,H,W,3] uint8 RGB: preprocessing fused into the gather (kU8 kernels)
    const float* w;        // [27][64] fp32 (HWIO flattened)
    const float* bias;     // [64]
    float* part;           // wgrad: per-CTA partial sums [gridDim.x][55][64] (rows 0..26 hi, 27..53 lo, 54 = bias)
    int H, W;
    long long total_pix;   // B*H*W
    int num_tiles;         // ceil(total_pix / 128)
    // staged-patch geometry (0 = tile shape not rectangular: per-thread global gather instead)
    int wt_log2;           // a warp's 32 pixels = (32 >> wt_log2) rows x (1 << wt_log2) pixels (patch path)
    int patch_ok;
    int w_log2, h_log2;    // log2(W), log2(H) when both are powers of two, else -1: shifts instead of per-tile divisions
};

constexpr int kF_PartRows = 55;                                // hi rows 0..26, lo rows 27..53, ones row 54
constexpr int kF_PatchRegs = 10;                               // ceil(3 * 34 * 3 / 32) elements per producer thread
constexpr int kF_WarpPatch = kF_PatchRegs * 32;                // one warp's haloed patch (1.25 KiB of packed hi|lo words)
constexpr int kF_PatchFloats = 4 * kF_WarpPatch;               // the four warps of a group (5 KiB)

// The input patch of a WARP's 32 pixels ((rows + 2) x (wtw + 2) pixels x 3 channels, zero padded, already preprocessed;
// wtw = min(tile width, 32), rows = 32 / wtw) is fetched with COALESCED loads one tile ahead into kF_PatchRegs registers
// per producer thread, parked in shared memory, and the 27-tap gather then reads shared memory (lane stride 3 words:
// conflict-free).  The v1 gather issued 27 dependent-latency global loads per pixel with only 8 producer warps per SM to
// hide them; ncu showed the kernel at 30 % of HBM bandwidth with every pipe idle.  The patch belongs to the warp (not
// to the 128-thread group, as in the first version: its named barrier per tile was the kernel's top stall, 1.7 stalled
// warps per issue slot in the round-2 ncu capture), so a __syncwarp is all the synchronisation the gather needs; the
// halo columns two neighbouring warps both load cost 5 % more (L1/L2-served) load instructions.
// Per-thread, tile-invariant description of patch element e = lane + 32 k, packed into one register:
//   bits 0-1 channel, bit 2 used, bit 3 first patch row, bit 4 last patch row, bit 5 first column, bit 6 last column,
//   bits 8.. offset (in input elements) of the value relative to the warp's first pixel, biased by kF_RelBias.
constexpr int kF_RelBias = 1 << 22;
template <bool kU8>
__device__ __forceinline__ int patch_code(const FirstArgs& p, int e) {
    const int wtw_log2 = p.wt_log2 < 5 ? p.wt_log2 : 5;
    const int pcols = (1 << wtw_log2) + 2, nrows = (32 >> wtw_log2) + 2;
    const bool used = e < nrows * pcols * 3;
    const int prow = e / (pcols * 3);
    const int rem = e - prow * (pcols * 3);
    const int pcol = rem / 3;
    const int ch = rem - pcol * 3;
    const int rel = ((prow - 1) * p.W + (pcol - 1)) * 3 + (kU8 ? 2 - ch : ch);      // uint8 frames are RGB: BGR = reversed
    return ch | (used ? 4 : 0) | (prow == 0 ? 8 : 0) | (prow == nrows - 1 ? 16 : 0) | (pcol == 0 ? 32 : 0) |
           (pcol == pcols - 1 ? 64 : 0) | ((rel + kF_RelBias) << 8);
}

// Issue this lane's loads of the patch of warp `wq` of tile `tile` (values land in r[]; nothing waits here).
template <bool kU8>
__device__ __forceinline__ void patch_fetch(const FirstArgs& p, int tile, int wq, const int (&code)[kF_PatchRegs],
                                            uint32_t (&r)[kF_PatchRegs]) {
    const long long pix0 = static_cast<long long>(tile) * kF_TileM + 32 * wq;
    int c0, h0;
    if (p.w_log2 >= 0) {       // ncu: the two 64-bit divisions per tile and thread were ~10 % of the kernel's instructions
        c0 = static_cast<int>(pix0 & (p.W - 1));
        h0 = static_cast<int>((pix0 >> p.w_log2) & (p.H - 1));
    } else {
        c0 = static_cast<int>(pix0 % p.W);
        h0 = static_cast<int>((pix0 / p.W) % p.H);
    }
    const int wtw_log2 = p.wt_log2 < 5 ? p.wt_log2 : 5;
    // which borders of the frame this warp's rectangle touches: an element on the matching patch border is zero padding
    // (bit 2 = "used" of every element: set for a warp beyond the last pixel of a short last tile -> all zeros)
    const int edge = (h0 == 0 ? 8 : 0) | (h0 + (32 >> wtw_log2) == p.H ? 16 : 0) | (c0 == 0 ? 32 : 0) |
                     (c0 + (1 << wtw_log2) == p.W ? 64 : 0) | (pix0 < p.total_pix ? 0 : 4);
    const long long base = pix0 * 3 - kF_RelBias;
#pragma unroll
    for (int k = 0; k < kF_PatchRegs; ++k) {
        const int cd = code[k];
        const bool ok = (cd & 4) && !(cd & edge & ~4) && !(edge & 4);
        const long long off = base + (cd >> 8);
        // uint8 frames: the RAW byte stays in the register (bit 8 = zero padding) and patch_value converts it when the
        // patch is written -- a conversion here would make every load of the prefetch wait for its data at once
        if (kU8) r[k] = ok ? static_cast<uint32_t>(__ldg(p.x8 + off)) : 0x100u;
        else r[k] = ok ? __float_as_uint(__ldg(p.x + off)) : 0u;
    }
}
// The value patch_fetch left in a register, as the float the convolution sees.
template <bool kU8>
__device__ __forceinline__ float patch_value(uint32_t raw, int ch) {
    if (!kU8) return __uint_as_float(raw);
    const float mean = ch == 0 ? 103.939f : (ch == 1 ? 116.779f : 123.68f);      // BGR means
    return (raw & 0x100u) ? 0.f : static_cast<float>(raw) - mean;
}

// Write one pixel's A row(s).  Only chunks that can hold data are stored: k-block 0 chunks 0..6 (slots 0..55) and, in
// forward, k-block 1 chunks 0..3 (hi(x) in slots 0..26, the constant 1 of the bias row in slot 27); every other chunk
// of the stage was zero-filled once at kernel start and is never written again.
// kTS (forward only): the row goes to TENSOR memory instead -- lane = tile row, 48 columns = the 64 slots of k-block 0
// followed by the 32 slots of k-block 1 -- and the MMA takes its A operand from there (tcgen05.mma [d], [a], b-desc):
// no swizzled shared-memory A tile is written or read at all (it was 45 % of the kernel's shared-memory traffic).
// a_tmem already carries the warp's lane base.
template <bool kFwd, bool kTS>
__device__ __forceinline__ void store_a_row(const uint32_t (&s0)[32], int row, uint8_t* a_stage, uint32_t a_tmem) {
    if (kTS) {
        tmem_st_32x32b_x32(a_tmem, s0);
        uint32_t s1[16];
#pragma unroll
        for (int j = 0; j < 13; ++j) s1[j] = s0[j];
        s1[13] = (s0[13] & 0xFFFFu) | 0x3F800000u;       // hi(x[26]), 1.0 (pairs with lo(bias) in the B operand)
        s1[14] = 0u; s1[15] = 0u;
        tmem_st_32x32b_x16(a_tmem + 32, s1);
        return;
    }
    uint8_t* r0 = a_stage + row * 128;
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 7; ++c)
        *reinterpret_cast<uint4*>(r0 + ((c ^ sw) << 4)) = make_uint4(s0[4 * c], s0[4 * c + 1], s0[4 * c + 2], s0[4 * c + 3]);
    if (kFwd) {
        uint8_t* r1 = r0 + kF_ABlock;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            *reinterpret_cast<uint4*>(r1 + ((c ^ sw) << 4)) = make_uint4(s0[4 * c], s0[4 * c + 1], s0[4 * c + 2], s0[4 * c + 3]);
        // slots 24, 25 | 26, 27 = hi(x[26]), 1.0 (pairs with lo(bias) in the B operand) | zeros
        *reinterpret_cast<uint4*>(r1 + ((3 ^ sw) << 4)) = make_uint4(s0[12], (s0[13] & 0xFFFFu) | 0x3F800000u, 0, 0);
    }
}

// One thread = one pixel: gather the 27-value patch (zero padded) and write the swizzled K-major row(s).
// kU8: the frame is raw uint8 RGB and Keras' vgg16 preprocess_input (caffe mode: RGB->BGR, subtract the
// ImageNet BGR means; reference preprocessor.py:60-63 / models.py:13) is applied on the fly; the SAME zero
// padding applies to the PREPROCESSED tensor, so out-of-frame taps stay exactly 0.
template <bool kFwd, bool kU8, bool kTS>
__device__ __forceinline__ void im2col_row(const FirstArgs& p, long long pix, int row, uint8_t* a_stage, uint32_t a_tmem) {
    const int wq = static_cast<int>(pix % p.W);
    const int hq = static_cast<int>((pix / p.W) % p.H);
    float v[27];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const bool hok = pix < p.total_pix && (hq + kh - 1) >= 0 && (hq + kh - 1) < p.H;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const bool ok = hok && (wq + kw - 1) >= 0 && (wq + kw - 1) < p.W;
            const long long off = (pix + (kh - 1) * p.W + (kw - 1)) * 3;
            if (kU8) {
                const float mean_bgr[3] = {103.939f, 116.779f, 123.68f};
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[(kh * 3 + kw) * 3 + c] = ok ? static_cast<float>(__ldg(p.x8 + off + (2 - c))) - mean_bgr[c] : 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[(kh * 3 + kw) * 3 + c] = ok ? __ldg(p.x + off + c) : 0.f;
            }
        }
    }
    // 64 slots of k-block 0 as 32 packed bf16x2 words
    uint32_t s0[32];
    __nv_bfloat16 hi[27], lo[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        hi[k] = __float2bfloat16_rn(v[k]);
        lo[k] = __float2bfloat16_rn(v[k] - __bfloat162float(hi[k]));
    }
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int k0 = 2 * j, k1 = 2 * j + 1;
        const __nv_bfloat16 e0 = k0 < 27 ? hi[k0] : (k0 < 54 ? lo[k0 - 27] : (k0 == 54 ? one : zero));
        const __nv_bfloat16 e1 = k1 < 27 ? hi[k1] : (k1 < 54 ? lo[k1 - 27] : (k1 == 54 ? one : zero));
        s0[j] = static_cast<uint32_t>(__bfloat16_as_ushort(e0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(e1)) << 16);
    }
    store_a_row<kFwd, kTS>(s0, row, a_stage, a_tmem);
}

// hi | lo << 16 of one input value: converted ONCE per input element when the patch is parked in shared memory
// (the v1 gather converted every element 9 times, once per tap that touches it: 54 F2F per pixel).
__device__ __forceinline__ uint32_t split_hi_lo(float v) {
    // the two-operand pack (F2FP.BF16.F32.PACK_AB) runs on the ALU pipe; the scalar F2F.BF16.F32 conversion it replaces
    // was the kernel's top stall in the round-2 ncu source view.  Same round-to-nearest-even result.
    const uint32_t hi = pack_bf16x2(v, 0.f) & 0xFFFFu;
    const uint32_t lo = pack_bf16x2(v - __uint_as_float(hi << 16), 0.f) & 0xFFFFu;
    return hi | (lo << 16);
}

// Same A rows as im2col_row, gathered from the packed shared-memory patch: 27 LDS + 27 byte-permutes per pixel.
template <bool kFwd, bool kTS>
__device__ __forceinline__ void im2col_row_patch(const uint32_t* patch, int wt_log2, int row, uint8_t* a_stage,
                                                 uint32_t a_tmem) {
    // patch = this WARP's patch; the lane's pixel sits at (lane >> wtw_log2, lane & (wtw - 1)) of the warp's rectangle
    const int wtw_log2 = wt_log2 < 5 ? wt_log2 : 5, lane = row & 31;
    const int pcols = (1 << wtw_log2) + 2;
    const uint32_t* base = patch + ((lane >> wtw_log2) * pcols + (lane & ((1 << wtw_log2) - 1))) * 3;
    uint32_t u[28];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int j = 0; j < 9; ++j) u[kh * 9 + j] = base[kh * pcols * 3 + j];
    u[27] = 0;
    uint32_t s0[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j <= 12) s0[j] = __byte_perm(u[2 * j], u[2 * j + 1], 0x5410);                   // hi[2j], hi[2j+1]
        else if (j == 13) s0[j] = __byte_perm(u[26], u[0], 0x7610);                         // hi[26], lo[0]
        else if (j <= 26) s0[j] = __byte_perm(u[2 * j - 27], u[2 * j - 26], 0x7632);        // lo[2j-27], lo[2j-26]
        else s0[j] = j == 27 ? 0x00003F80u : 0u;                                            // slot 54 = 1.0
    }
    store_a_row<kFwd, kTS>(s0, row, a_stage, a_tmem);
}

// im2col producer loop of one 128-thread group (group g builds the CTA's tiles g, g+2, ...): shared by forward and
// wgrad.  kBlocks = 16 KiB k-blocks per A stage (2 forward, 1 wgrad).
// kTS: the A rows go to tensor memory (a_tmem0 = stage 0, this warp's lanes; kF_TsCols columns per stage).
template <bool kFwd, bool kU8, bool kTS, typename AfterTile>
__device__ __forceinline__ void producer_loop(const FirstArgs& p, uint8_t* smem_a, uint32_t a_tmem0, float* patch,
                                              uint64_t* afull, uint64_t* aempty, int row, int grp, AfterTile after_tile) {
    constexpr int kStageBytes = (kFwd ? 2 : 1) * kF_ABlock;
    int code[kF_PatchRegs];
    uint32_t nxt[kF_PatchRegs];
    const int lane = row & 31, wq = row >> 5;
    const int ch0 = lane % 3;                 // channel of patch element e = lane + 32 k: e % 3 = (lane % 3 + 2 k) % 3
    if (p.patch_ok) {
#pragma unroll
        for (int k = 0; k < kF_PatchRegs; ++k) code[k] = patch_code<kU8>(p, lane + 32 * k);
        if (blockIdx.x + static_cast<long long>(grp) * gridDim.x < p.num_tiles)
            patch_fetch<kU8>(p, blockIdx.x + grp * gridDim.x, wq, code, nxt);
    }
    int pbuf = 0;
    for (int it = grp; blockIdx.x + static_cast<long long>(it) * gridDim.x < p.num_tiles; it += kF_ProdGroups) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int stage = it % kF_Stages;
        uint32_t* pcur = reinterpret_cast<uint32_t*>(patch) + pbuf * kF_PatchFloats + wq * kF_WarpPatch;
        if (p.patch_ok) {
            // the warp's patch is double buffered: this tile's values go to the buffer the tile BEFORE the previous one
            // used (every lane passed the previous tile's __syncwarp only after all lanes had finished that older
            // gather), so one __syncwarp per tile is enough
#pragma unroll
            for (int k = 0; k < kF_PatchRegs; ++k)
                pcur[lane + 32 * k] = split_hi_lo(patch_value<kU8>(nxt[k], (ch0 + 2 * k) % 3));
            if (blockIdx.x + static_cast<long long>(it + kF_ProdGroups) * gridDim.x < p.num_tiles)
                patch_fetch<kU8>(p, tile + kF_ProdGroups * gridDim.x, wq, code, nxt);     // next tile: in flight
            __syncwarp();                                 // patch complete
            pbuf ^= 1;
        }
        mbar_wait(&aempty[stage], ((it / kF_Stages) & 1) ^ 1);
        if (kTS) tc_fence_after();                        // the MMAs that read this stage's columns have completed
        if (p.patch_ok)
            im2col_row_patch<kFwd, kTS>(pcur, p.wt_log2, row, smem_a + stage * kStageBytes, a_tmem0 + stage * kF_TsCols);
        else
            im2col_row<kFwd, kU8, kTS>(p, static_cast<long long>(tile) * kF_TileM + row, row, smem_a + stage * kStageBytes,
                                       a_tmem0 + stage * kF_TsCols);
        if (kTS) {
            tmem_st_wait();
            tc_fence_before();
        } else {
            fence_proxy_async_smem();
        }
        mbar_arrive(&afull[stage]);
        if (it >= kF_ProdGroups) after_tile(it - kF_ProdGroups);      // e.g. epilogue of this group's previous tile
    }
}

// ------------------------------------------------------------------------------------------ forward
// kTS: A operand in tensor memory (store_a_row); ABL_FIRST_SMEM_A=1 selects the shared-memory form.
template <bool kU8, bool kTS>
__global__ void __launch_bounds__(kF_Threads, 1)
conv_first_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_out, const FirstArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kASmem = kTS ? 0 : kF_Stages * 2 * kF_ABlock;
    constexpr int kTmemCols = kTS ? 512 : kF_AccBufs * 64;        // accumulators [0, 256) + kF_Stages A stages from 256
    uint8_t* smem_a = smem;                                       // kF_Stages x 2 k-blocks x 16 KiB (not kTS)
    uint8_t* smem_b = smem + kASmem;                              // 2 k-blocks x [64 co][64 slots] = 16 KiB
    uint8_t* staging = smem_b + 2 * 64 * 128;                     // one 16 KiB swizzled output tile per worker group
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kF_ProdGroups * kF_ABlock);
    uint64_t* afull = bars;
    uint64_t* aempty = bars + kF_Stages;
    uint64_t* tfull = bars + 2 * kF_Stages;
    uint64_t* tempty = tfull + kF_AccBufs;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kF_AccBufs);
    float* patches = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    const int warp = threadIdx.x >> 5;

    if (warp == 1 && elect_one()) {
        for (int s = 0; s < kF_Stages; ++s) { mbar_init(&afull[s], 128); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < kF_AccBufs; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
        fence_barrier_init();
        tma_prefetch_desc(&map_out);
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    for (int i = threadIdx.x; i < kASmem / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);       // see store_a_row
    // B operand, built once per CTA: row = co (128 B, swizzled), slot k of block 0 = hi(W[k%27]) for k < 54,
    // block 1 = lo(W[k]) for k < 27.  The bias rides along: block 0 slot 54 (A holds 1.0 there) = hi(b[co]) and
    // block 1 slot 27 (A: 1.0) = lo(b[co]), so the accumulator already contains conv + bias in fp32-like precision.
    for (int i = threadIdx.x; i < 2 * 64 * 64; i += blockDim.x) {
        const int blk = i >> 12, co = (i >> 6) & 63, slot = i & 63;
        float val = 0.f;
        if (blk == 0 && slot < 54) {
            val = __bfloat162float(__float2bfloat16_rn(p.w[(slot % 27) * kF_Cout + co]));
        } else if (blk == 1 && slot < 27) {
            const float wv = p.w[slot * kF_Cout + co];
            val = wv - __bfloat162float(__float2bfloat16_rn(wv));
        } else if (blk == 0 && slot == 54) {
            val = __bfloat162float(__float2bfloat16_rn(p.bias[co]));
        } else if (blk == 1 && slot == 27) {
            val = p.bias[co] - __bfloat162float(__float2bfloat16_rn(p.bias[co]));
        }
        uint8_t* rowp = smem_b + blk * (64 * 128) + co * 128;
        *reinterpret_cast<__nv_bfloat16*>(rowp + ((((slot >> 3) ^ (co & 7)) << 4) | ((slot & 7) << 1))) =
            __float2bfloat16_rn(val);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a), 16);
            const uint32_t b_lo0 = desc_lo_sw128(smem_u32(smem_b), 16);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int as = it % kF_AccBufs;
                mbar_wait(&tempty[as], ((it / kF_AccBufs) & 1) ^ 1);
                mbar_wait(&afull[stage], phase);
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + stage * ((2 * kF_ABlock) >> 4);
                // k-block 0 holds 55 slots (4 K = 16 steps), k-block 1 only 28 (2 steps): its upper half is all zeros
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < (kb == 0 ? 4 : 2); ++k) {
                        if (kTS)      // 8 columns = 16 bf16 per K step; k-block 1 starts at column 32 of the stage
                            umma_bf16_ts(tmem_base + as * 64, tmem_base + kF_AccBufs * 64 + stage * kF_TsCols + kb * 32 + k * 8,
                                         b_lo0 + kb * ((64 * 128) >> 4) + k * 2, dhi, idesc, (kb | k) ? 1u : 0u);
                        else
                            umma_bf16_lohi(tmem_base + as * 64, a_lo + kb * (kF_ABlock >> 4) + k * 2, dhi,
                                           b_lo0 + kb * ((64 * 128) >> 4) + k * 2, dhi, idesc, (kb | k) ? 1u : 0u);
                    }
                umma_commit(&aempty[stage]);
                umma_commit(&tfull[as]);
                if (++stage == kF_Stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- worker groups (3 x 128 threads)
        // Group g: for each of its tiles, build the im2col rows, then drain the accumulator of its PREVIOUS tile
        // (whose MMAs ran meanwhile) -> bias + ReLU -> bf16 -> swizzled staging -> TMA store.  With three groups the
        // epilogue latency of one tile (a single warp per scheduler: ~1.8k clocks) overlaps two other tiles' work;
        // the v1 layout had ONE epilogue group and was bound by exactly that latency (221 tiles x 1.8k clocks).
        const int row = (threadIdx.x - 128) & 127;       // tile row == TMEM lane (group base warp is a multiple of 4)
        const int grp = (threadIdx.x - 128) >> 7;
        uint8_t* stg = staging + grp * kF_ABlock;
        // Every warp drains its own 32 accumulator rows = 32 consecutive raster pixels: private 4 KiB staging slice, its own
        // TMA store (box of 32 pixels), warp-level synchronisation only (the group-wide named barriers of the first
        // version showed up as the kernel's largest stall class).
        const int lane = threadIdx.x & 31, wq = row >> 5;
        uint8_t* stg_w = stg + wq * 4096;
        auto epilogue = [&](int it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int as = it % kF_AccBufs;
            mbar_wait(&tfull[as], (it / kF_AccBufs) & 1);
            tc_fence_after();
            if (lane == 0) tma_store_wait_read<0>();     // this warp's previous store has finished reading its slice
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(row & ~31) << 16) + as * 64 + c * 32, r);
                tmem_ld_wait();
                uint32_t packed[16];       // the accumulator already holds conv + bias (bias row of the B operand)
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    packed[j] = pack_bf16x2_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stg_w + lane * 128 + (((c * 4 + j) ^ (lane & 7)) << 4)) =
                        make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
            tc_fence_before();
            mbar_arrive(&tempty[as]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&map_out, stg_w, 0, tile * kF_TileM + wq * 32);
                tma_store_commit();
            }
        };
        producer_loop<true, kU8, kTS>(p, smem_a, tmem_base + (static_cast<uint32_t>(row & ~31) << 16) + kF_AccBufs * 64,
                                      patches + grp * 2 * kF_PatchFloats, afull, aempty, row, grp, epilogue);
        // the last tile of this group has not been drained yet
        int n_here = 0;
        for (int it = grp; blockIdx.x + static_cast<long long>(it) * gridDim.x < p.num_tiles; it += kF_ProdGroups) ++n_here;
        if (n_here > 0) epilogue(grp + (n_here - 1) * kF_ProdGroups);
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------ wgrad
// Persistent CTAs; each accumulates D[slot][co] over its tiles in TMEM, then stores rows 0..54 as its partial sums.
// kM64: the MMA is M = 64 (the 64 slots) instead of M = 128 with a shared block of zeros as rows 64..127: the tensor
// core's issue cost is the same, but each MMA reads 4 KiB instead of 6 KiB of shared memory, and this kernel is bound by
// shared-memory bandwidth.  An M = 64 accumulator sits in 16 lanes of each 32-lane quarter (row r -> lane r % 16 + 32 *
// (r / 16)).  ABL_FIRST_WGRAD_M128=1 selects the M = 128 form.
template <bool kU8, bool kM64>
__global__ void __launch_bounds__(kF_Threads, 1)
conv_first_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dz, const FirstArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;                                   // kF_Stages x 16 KiB (k-block 0 rows)
    uint8_t* smem_z = smem + kF_Stages * kF_ABlock;           // 16 KiB of zeros: M rows 64..127
    uint8_t* smem_b = smem_z + kF_ABlock;                     // kF_Stages x 16 KiB dZ tiles [128 px][64 co]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kF_Stages * kF_ABlock);
    uint64_t* afull = bars;                 // 128 producer arrivals + 1 expect_tx arrival (TMA bytes)
    uint64_t* aempty = bars + kF_Stages;
    uint64_t* done = bars + 2 * kF_Stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    float* patches = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    const int warp = threadIdx.x >> 5;

    if (warp == 1 && elect_one()) {
        for (int s = 0; s < kF_Stages; ++s) { mbar_init(&afull[s], 129); mbar_init(&aempty[s], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
        tma_prefetch_desc(&map_dz);
    }
    if (warp == 2) tmem_alloc<64>(tmem_slot);
    for (int i = threadIdx.x; i < (kF_Stages + 1) * kF_ABlock / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);       // A stages (see store_a_row) + the zero block
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ntiles_here = blockIdx.x < p.num_tiles ? (p.num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp == 0) {
        if (elect_one()) {      // dZ tile loader
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&aempty[stage], phase ^ 1);
                mbar_expect_tx(&afull[stage], kF_ABlock);
                tma_load_2d(&map_dz, &afull[stage], smem_b + stage * kF_ABlock, 0, tile * kF_TileM);
                if (++stage == kF_Stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kM64 ? 64 : 128, 64, 1, 1);      // both operands MN-major (K = pixels)
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            int stage = 0;
            uint32_t phase = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&afull[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + stage * kF_ABlock);
                // M block 0 = this stage's im2col rows, M block 1 (LBO away) = the shared zero block
                const uint32_t a_lo = desc_lo_sw128(a_addr, smem_u32(smem_z) - a_addr);
                const uint32_t b_lo = desc_lo_sw128(smem_u32(smem_b + stage * kF_ABlock), 0);
#pragma unroll
                for (int k = 0; k < kF_TileM / 16; ++k)
                    umma_bf16_lohi(tmem_base, a_lo + k * (2048 >> 4), dhi, b_lo + k * (2048 >> 4), dhi, idesc,
                                   (first && k == 0) ? 0u : 1u);
                first = false;
                umma_commit(&aempty[stage]);
                if (++stage == kF_Stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done);
        }
    } else if (warp >= 4) {
        {
            const int row = (threadIdx.x - 128) & 127;
            const int grp = (threadIdx.x - 128) >> 7;
            producer_loop<false, kU8, false>(p, smem_a, 0u, patches + grp * 2 * kF_PatchFloats, afull, aempty, row, grp, [](int) {});
        }
        // D rows 0..63 live in TMEM lanes 0..63: warps 4 and 5 (lane quarters 0 and 1) drain them at the end
        // (M = 64: in lanes 0..15 of all four quarters, warps 4..7)
        const int tl = threadIdx.x - 128;                  // TMEM lane of this thread (first worker group only)
        const int row = kM64 ? ((tl & 31) < 16 ? (tl >> 5) * 16 + (tl & 15) : kF_PartRows) : tl;
        if (warp < (kM64 ? 8 : 6) && ntiles_here > 0) {
            mbar_wait(done, 0);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(tl & ~31) << 16) + c * 32, r);
                tmem_ld_wait();
                if (row < kF_PartRows) {
                    // plain stores of this CTA's sums; conv_first_wgrad_reduce_kernel adds the CTAs in a fixed order
                    float4* dst = reinterpret_cast<float4*>(p.part + (static_cast<size_t>(blockIdx.x) * kF_PartRows + row) * kF_Cout + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                             __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<64>(tmem_base);
    }
}

// dw[k][co] = sum_cta (part[cta][k][co] + part[cta][27 + k][co]), db[co] = sum_cta part[cta][54][co], CTAs in a fixed
// order: block = 4 float4 columns x 64 CTA groups (thread (col, g) adds CTAs g, g+64, ...; groups combined in order).
__global__ void __launch_bounds__(256)
conv_first_wgrad_reduce_kernel(const float4* __restrict__ part, int n_cta, float4* __restrict__ dw, float4* __restrict__ db) {
    __shared__ float4 red[64][4];
    constexpr int kRow4 = kF_Cout / 4, kDw4 = 27 * kRow4, kOut4 = kDw4 + kRow4, kCta4 = kF_PartRows * kRow4;
    const int cl = threadIdx.x & 3, g = threadIdx.x >> 2;
    const int o4 = blockIdx.x * 4 + cl;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o4 < kOut4) {
        const int src = o4 < kDw4 ? o4 : 54 * kRow4 + (o4 - kDw4);
        for (int c = g; c < n_cta; c += 64) {
            const float4 v = __ldcg(part + static_cast<size_t>(c) * kCta4 + src);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            if (o4 < kDw4) {
                const float4 l = __ldcg(part + static_cast<size_t>(c) * kCta4 + src + kDw4);
                s.x += l.x; s.y += l.y; s.z += l.z; s.w += l.w;
            }
        }
    }
    red[g][cl] = s;
    __syncthreads();
    if (g == 0 && o4 < kOut4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int q = 0; q < 64; ++q) {
            const float4 v = red[q][cl];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        if (o4 < kDw4) dw[o4] = t;
        else if (db != nullptr) db[o4 - kDw4] = t;
    }
}

static int first_args(FirstArgs* a, const void* x, bool u8, int B, int H, int W, int Cout, const char* who) {
    if (B <= 0 || H <= 0 || W <= 0 || Cout != kF_Cout) {
        char buf[160];
        snprintf(buf, sizeof(buf), "%s: need B, H, W > 0 and Cout == 64", who);
        set_error(buf);
        return ABL_ERR_SHAPE;
    }
    a->x = u8 ? nullptr : static_cast<const float*>(x);
    a->x8 = u8 ? static_cast<const uint8_t*>(x) : nullptr;
    a->H = H; a->W = W;
    a->total_pix = static_cast<long long>(B) * H * W;
    // the last tile may be short (gather path only): its missing pixels read as zeros, TMA zero-fills dz rows beyond
    // total_pix and clips the forward store
    a->num_tiles = static_cast<int>((a->total_pix + kF_TileM - 1) / kF_TileM);
    a->w = nullptr; a->bias = nullptr; a->part = nullptr;
    // staged patch: a WARP's 32 raster pixels must be a rectangle inside one frame -- a 32-pixel piece of one image row
    // (W a multiple of 32: 96, 128, 160, 224 ...) or 32 / W whole rows of a narrow frame (W = 1, 2, ... 16)
    const bool wide = W % 32 == 0;
    const bool narrow = W < 32 && (W & (W - 1)) == 0 && (static_cast<long long>(H) * W) % 32 == 0;
    a->patch_ok = ((wide || narrow) && getenv("ABL_FIRST_GATHER") == nullptr) ? 1 : 0;
    a->wt_log2 = 5;            // log2 of the warp rectangle's width
    if (!wide) { a->wt_log2 = 0; while ((1 << a->wt_log2) < W && a->wt_log2 < 5) ++a->wt_log2; }
    a->w_log2 = a->h_log2 = -1;
    if ((W & (W - 1)) == 0 && (H & (H - 1)) == 0) {
        a->w_log2 = a->h_log2 = 0;
        while ((1 << a->w_log2) < W) ++a->w_log2;
        while ((1 << a->h_log2) < H) ++a->h_log2;
    }
    return ABL_OK;
}

static int pixel_map(CUtensorMap* m, const void* base, long long total_pix, int box_rows) {
    const uint64_t dims[2] = {64, (uint64_t)total_pix};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, (uint32_t)box_rows};
    return encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <bool kU8, bool kTS>
static int first_fwd(const void* x, const float* w_hwio, const float* bias, void* out, int B, int H, int W, int Cout,
                     void* stream, const char* who) {
    if (!x || !w_hwio || !bias || !out) { set_error("abl_conv_first_fwd: null pointer"); return ABL_ERR_NULL; }
    if ((reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15)) {
        set_error("abl_conv_first_fwd: out and bias must be 16-byte aligned"); return ABL_ERR_ALIGN;
    }
    FirstArgs a;
    int r = first_args(&a, x, kU8, B, H, W, Cout, who);
    if (r != ABL_OK) return r;
    a.w = w_hwio; a.bias = bias;
    CUtensorMap mo;
    r = pixel_map(&mo, out, a.total_pix, 32);           // one store per epilogue warp: 32 pixels x 64 channels
    if (r != ABL_OK) return r;
    constexpr int smem_rest = 2 * 64 * 128 + kF_ProdGroups * kF_ABlock + 256 + kF_ProdGroups * 2 * kF_PatchFloats * 4;
    constexpr int smem_bytes = smem_rest + (kTS ? 0 : kF_Stages * 2 * kF_ABlock);
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(conv_first_fwd_tc_kernel<kU8, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
            return check_launch();
        attr_done[dev & 63] = true;
    }
    const int grid = a.num_tiles < sm_count() ? a.num_tiles : sm_count();
    conv_first_fwd_tc_kernel<kU8, kTS><<<grid, kF_Threads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(mo, a);
    return check_launch();
}

template <bool kU8, bool kM64>
static int first_wgrad(const void* x, const void* dz, float* dw, float* db, void* workspace, long long workspace_bytes,
                       int B, int H, int W, int Cout, void* stream, const char* who) {
    if (!x || !dz || !dw || !workspace) { set_error("abl_conv_first_wgrad: null pointer"); return ABL_ERR_NULL; }
    if ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(db) |
         reinterpret_cast<uintptr_t>(workspace)) & 15) {
        set_error("abl_conv_first_wgrad: dz, dw, db and workspace must be 16-byte aligned"); return ABL_ERR_ALIGN;
    }
    FirstArgs a;
    int r = first_args(&a, x, kU8, B, H, W, Cout, who);
    if (r != ABL_OK) return r;
    const int grid = a.num_tiles < sm_count() ? a.num_tiles : sm_count();
    if (workspace_bytes < static_cast<long long>(grid) * kF_PartRows * kF_Cout * 4) {
        set_error("abl_conv_first_wgrad: workspace too small (abl_conv_first_wgrad_workspace_bytes)"); return ABL_ERR_SHAPE;
    }
    a.part = static_cast<float*>(workspace);
    CUtensorMap mz;
    r = pixel_map(&mz, dz, a.total_pix, kF_TileM);
    if (r != ABL_OK) return r;
    constexpr int smem_bytes = (2 * kF_Stages + 1) * kF_ABlock + 256 + kF_ProdGroups * 2 * kF_PatchFloats * 4;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(conv_first_wgrad_tc_kernel<kU8, kM64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
            return check_launch();
        attr_done[dev & 63] = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    conv_first_wgrad_tc_kernel<kU8, kM64><<<grid, kF_Threads, smem_bytes, st>>>(mz, a);
    r = check_launch();
    if (r != ABL_OK) return r;
    constexpr int kOut4 = 28 * (kF_Cout / 4);
    conv_first_wgrad_reduce_kernel<<<(kOut4 + 3) / 4, 256, 0, st>>>(static_cast<const float4*>(workspace), grid,
                                                                   reinterpret_cast<float4*>(dw), reinterpret_cast<float4*>(db));
    return check_launch();
}

}  // namespace abl

using namespace abl;

extern "C" int abl_conv_first_fwd(const float* x, const float* w_hwio, const float* bias, void* out, int B, int H,
                                  int W, int Cout, void* stream) {
    const char* sa = getenv("ABL_FIRST_SMEM_A");       // A/B switch: the shared-memory A operand of round 1
    if (sa && sa[0] == '1') return first_fwd<false, false>(x, w_hwio, bias, out, B, H, W, Cout, stream, "abl_conv_first_fwd");
    return first_fwd<false, true>(x, w_hwio, bias, out, B, H, W, Cout, stream, "abl_conv_first_fwd");
}
extern "C" int abl_conv_first_fwd_u8(const unsigned char* frames_rgb, const float* w_hwio, const float* bias, void* out,
                                     int B, int H, int W, int Cout, void* stream) {
    const char* sa = getenv("ABL_FIRST_SMEM_A");
    if (sa && sa[0] == '1')
        return first_fwd<true, false>(frames_rgb, w_hwio, bias, out, B, H, W, Cout, stream, "abl_conv_first_fwd_u8");
    return first_fwd<true, true>(frames_rgb, w_hwio, bias, out, B, H, W, Cout, stream, "abl_conv_first_fwd_u8");
}
extern "C" long long abl_conv_first_wgrad_workspace_bytes(void) {
    return static_cast<long long>(sm_count()) * kF_PartRows * kF_Cout * 4;
}
extern "C" int abl_conv_first_wgrad(const float* x, const void* dz, float* dw, float* db, void* workspace,
                                    long long workspace_bytes, int B, int H, int W, int Cout, void* stream) {
    const char* m = getenv("ABL_FIRST_WGRAD_M128");
    if (m && m[0] == '1')
        return first_wgrad<false, false>(x, dz, dw, db, workspace, workspace_bytes, B, H, W, Cout, stream, "abl_conv_first_wgrad");
    return first_wgrad<false, true>(x, dz, dw, db, workspace, workspace_bytes, B, H, W, Cout, stream, "abl_conv_first_wgrad");
}
extern "C" int abl_conv_first_wgrad_u8(const unsigned char* frames_rgb, const void* dz, float* dw, float* db,
                                       void* workspace, long long workspace_bytes, int B, int H, int W, int Cout,
                                       void* stream) {
    const char* m = getenv("ABL_FIRST_WGRAD_M128");
    if (m && m[0] == '1')
        return first_wgrad<true, false>(frames_rgb, dz, dw, db, workspace, workspace_bytes, B, H, W, Cout, stream,
                                        "abl_conv_first_wgrad_u8");
    return first_wgrad<true, true>(frames_rgb, dz, dw, db, workspace, workspace_bytes, B, H, W, Cout, stream,
                                   "abl_conv_first_wgrad_u8");
}
