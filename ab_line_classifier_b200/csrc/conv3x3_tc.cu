// 3x3 / stride 1 / SAME convolution as a tcgen05 implicit GEMM (sm_100a), NHWC bf16 in, fp32
// accumulation in TMEM.  One kernel serves
//   * forward  (Keras Conv2D+BiasAdd+Relu [+MaxPool 2x2], reference call site models.py:412-413):
//       Y[n,h,w,co] = relu(b[co] + sum_{kh,kw,ci} X[n,h+kh-1,w+kw-1,ci] * Wf[kh*3+kw][co][ci])
//   * dgrad    (Conv2DBackpropInput + ReluGrad of the layer below): the same contraction with
//       X := dZ (Cout channels) and the tap-flipped, transposed filter Wd, epilogue = ReLU mask.
//
// GEMM view per CTA tile: M = 128 output pixels (a box_h x box_w rectangle of one frame),
// N = NTILE output channels, K = 9 taps x Cin.  A k-block is one tap x 64 input channels:
//   A (K-major, 128 rows x 128 B, SWIZZLE_128B) = TMA box {64 ch, box_w, box_h, 1} of the NHWC
//     tensor at (c0, w0+kw-1, h0+kh-1, n); the SAME zero padding is TMA out-of-bounds fill.
//   B (K-major, NTILE rows x 128 B)             = TMA box {64 ci, NTILE co, 1 tap} of Wf.
// Warp roles (256 threads, persistent, 1 CTA/SM): warp 0 TMA producer, warp 1 MMA issuer
// (single thread), warp 2 TMEM allocator, warps 4-7 epilogue.  The accumulator is double
// buffered in TMEM (2 x NTILE columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "host_util.h"
#include "ptx.cuh"
#include "../../include/abl_b200.h"

namespace abl {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // bf16 elements per k-block row = 128 bytes
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KiB
constexpr int kEpiThreads = 128;
constexpr int kThreads = 256;

struct ConvArgs {
    int B, H, W, Cin, Cout;
    int box_h, box_w, tiles_h, tiles_w;
    int num_m_tiles, num_n_tiles, kchunks;
    int flags;
    const float* bias;              // [Cout] fp32 (ABL_EPI_BIAS_RELU)
    const __nv_bfloat16* mask_src;  // [B,H,W,Cout] (ABL_EPI_RELU_MASK)
    __nv_bfloat16* out_full;        // [B,H,W,Cout] or null
    __nv_bfloat16* out_pool;        // [B,H/2,W/2,Cout] or null
};

template <int NTILE>
struct ConvCfg {
    static constexpr int kStages = 4;
    static constexpr int kBBytes = NTILE * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiN = NTILE < 128 ? NTILE : 128;        // columns staged per epilogue pass
    static constexpr int kStageRow = kEpiN * 2 + 16;               // padded staging row (bank-conflict free)
    static constexpr int kStagingBytes = kTileM * kStageRow;
    static constexpr int kTmemCols = 2 * NTILE;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 128;
};

template <int NTILE>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const ConvArgs p) {
    using Cfg = ConvCfg<NTILE>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kStages * kABytes;
    uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::kStages;
    uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int kblocks = 9 * p.kchunks;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], kEpiThreads);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.num_n_tiles;
                int m_tile = tile / p.num_n_tiles;
                const int tw = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th = m_tile % p.tiles_h;
                const int img = m_tile / p.tiles_h;
                const int h0 = th * p.box_h, w0 = tw * p.box_w, n0 = n_tile * NTILE;
                for (int tap = 0; tap < 9; ++tap) {
                    const int kh = tap / 3, kw = tap - kh * 3;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        tma_load_4d(&map_a, &full_bar[stage], smem_a + stage * kABytes, kc * kBlockK,
                                    w0 + kw - 1, h0 + kh - 1, img);
                        tma_load_3d(&map_b, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kc * kBlockK, n0,
                                    tap);
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kTileM, NTILE, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * NTILE;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
                    const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);          // accumulator ready for the epilogue
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (128 threads)
        const int et = threadIdx.x - (kThreads - kEpiThreads);   // 0..127 == tile row == TMEM lane
        const int ewarp = et >> 5;
        constexpr int kChunks16 = Cfg::kEpiN / 8;                // 16-byte chunks per staged row
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int n_tile = tile % p.num_n_tiles;
            int m_tile = tile / p.num_n_tiles;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int img = m_tile / p.tiles_h;
            const int h0 = th * p.box_h, w0 = tw * p.box_w, n0 = n_tile * NTILE;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;

            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();

#pragma unroll 1
            for (int half = 0; half < NTILE / Cfg::kEpiN; ++half) {
                named_bar_sync(1, kEpiThreads);   // staging buffer free
                uint8_t* my_row = staging + et * Cfg::kStageRow;
#pragma unroll 1
                for (int c = 0; c < Cfg::kEpiN / 32; ++c) {
                    uint32_t r[32];
                    const uint32_t col = as * NTILE + half * Cfg::kEpiN + c * 32;
                    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + col, r);
                    tmem_ld_wait();
                    uint32_t packed[16];
                    if (p.flags & ABL_EPI_BIAS_RELU) {
                        // warp-uniform address: one broadcast transaction per load, L1-resident
                        const float4* bs = reinterpret_cast<const float4*>(p.bias + n0 + half * Cfg::kEpiN + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(bs + j);
                            packed[2 * j] = pack_bf16x2(fmaxf(__uint_as_float(r[4 * j]) + b4.x, 0.f),
                                                        fmaxf(__uint_as_float(r[4 * j + 1]) + b4.y, 0.f));
                            packed[2 * j + 1] = pack_bf16x2(fmaxf(__uint_as_float(r[4 * j + 2]) + b4.z, 0.f),
                                                            fmaxf(__uint_as_float(r[4 * j + 3]) + b4.w, 0.f));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            packed[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    }
                    uint4* dst = reinterpret_cast<uint4*>(my_row + c * 64);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
                if (half == NTILE / Cfg::kEpiN - 1) {
                    tc_fence_before();
                    mbar_arrive(&tempty_bar[as]);   // TMEM buffer may be overwritten by the next tile
                }
                named_bar_sync(1, kEpiThreads);     // staging filled

                const int cbase = n0 + half * Cfg::kEpiN;
                if (p.out_full != nullptr) {
                    for (int q = et; q < kTileM * kChunks16; q += kEpiThreads) {
                        const int row = q / kChunks16, cc = q - row * kChunks16;
                        const int h = h0 + row / p.box_w, w = w0 + row % p.box_w;
                        const size_t off = ((static_cast<size_t>(img) * p.H + h) * p.W + w) * p.Cout + cbase + cc * 8;
                        uint4 v = *reinterpret_cast<const uint4*>(staging + row * Cfg::kStageRow + cc * 16);
                        if (p.flags & ABL_EPI_RELU_MASK) {
                            const uint4 m = *reinterpret_cast<const uint4*>(p.mask_src + off);
                            const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
                            uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                // post-ReLU activations are >= 0: "> 0" <=> any magnitude bit set, sign clear
                                const uint32_t lo_on = ((mm[j] & 0x7FFFu) != 0 && !(mm[j] & 0x8000u)) ? 0xFFFFu : 0u;
                                const uint32_t hi_on =
                                    ((mm[j] & 0x7FFF0000u) != 0 && !(mm[j] & 0x80000000u)) ? 0xFFFF0000u : 0u;
                                vv[j] &= (lo_on | hi_on);
                            }
                            v = make_uint4(vv[0], vv[1], vv[2], vv[3]);
                        }
                        *reinterpret_cast<uint4*>(p.out_full + off) = v;
                    }
                }
                if (p.out_pool != nullptr) {
                    const int pw_n = p.box_w >> 1;
                    const int pooled = (kTileM / 4) * kChunks16;
                    for (int q = et; q < pooled; q += kEpiThreads) {
                        const int prow = q / kChunks16, cc = q - prow * kChunks16;
                        const int ph = prow / pw_n, pw = prow - ph * pw_n;
                        const int r00 = (2 * ph) * p.box_w + 2 * pw;
                        const uint8_t* s = staging + cc * 16;
                        const uint4 a = *reinterpret_cast<const uint4*>(s + r00 * Cfg::kStageRow);
                        const uint4 b = *reinterpret_cast<const uint4*>(s + (r00 + 1) * Cfg::kStageRow);
                        const uint4 c2 = *reinterpret_cast<const uint4*>(s + (r00 + p.box_w) * Cfg::kStageRow);
                        const uint4 d = *reinterpret_cast<const uint4*>(s + (r00 + p.box_w + 1) * Cfg::kStageRow);
                        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                        const uint32_t cv[4] = {c2.x, c2.y, c2.z, c2.w}, dv[4] = {d.x, d.y, d.z, d.w};
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __nv_bfloat162 m0 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&av[j]),
                                                              *reinterpret_cast<const __nv_bfloat162*>(&bv[j]));
                            const __nv_bfloat162 m1 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&cv[j]),
                                                              *reinterpret_cast<const __nv_bfloat162*>(&dv[j]));
                            const __nv_bfloat162 m2 = __hmax2(m0, m1);
                            o[j] = *reinterpret_cast<const uint32_t*>(&m2);
                        }
                        const int h = (h0 >> 1) + ph, w = (w0 >> 1) + pw;
                        const size_t off =
                            ((static_cast<size_t>(img) * (p.H >> 1) + h) * (p.W >> 1) + w) * p.Cout + cbase + cc * 8;
                        *reinterpret_cast<uint4*>(p.out_pool + off) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// Pick the tile rectangle: box_w = largest of {64,32,16,8} with W % box_w == 0 and
// H % (128 / box_w) == 0 (box_w >= 8 keeps a tile row a whole number of 1024-byte swizzle atoms).
static int pick_box(int H, int W, int* box_h, int* box_w) {
    for (int bw = 64; bw >= 8; bw >>= 1) {
        const int bh = kTileM / bw;
        if (W % bw == 0 && H % bh == 0) {
            *box_w = bw;
            *box_h = bh;
            return ABL_OK;
        }
    }
    return ABL_ERR_SHAPE;
}

template <int NTILE>
static int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const ConvArgs& a, cudaStream_t st) {
    using Cfg = ConvCfg<NTILE>;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(conv3x3_tc_kernel<NTILE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes) != cudaSuccess)
            return check_launch();
        attr_done[dev & 63] = true;
    }
    const int tiles = a.num_m_tiles * a.num_n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    conv3x3_tc_kernel<NTILE><<<grid, kThreads, Cfg::kSmemBytes, st>>>(ma, mb, a);
    return check_launch();
}

}  // namespace abl

using namespace abl;

// See include/abl_b200.h for the contract.
extern "C" int abl_conv3x3_igemm(const void* x, const void* w_kmajor, const float* bias, const void* mask_src,
                                 void* out_full, void* out_pool, int B, int H, int W, int Cin, int Cout,
                                 int flags, void* stream) {
    if (!x || !w_kmajor) { set_error("abl_conv3x3_igemm: null input"); return ABL_ERR_NULL; }
    if (!out_full && !out_pool) { set_error("abl_conv3x3_igemm: no output requested"); return ABL_ERR_NULL; }
    if ((flags & ABL_EPI_BIAS_RELU) && !bias) { set_error("abl_conv3x3_igemm: bias missing"); return ABL_ERR_NULL; }
    if ((flags & ABL_EPI_RELU_MASK) && (!mask_src || !out_full)) {
        set_error("abl_conv3x3_igemm: relu-mask epilogue needs mask_src and out_full");
        return ABL_ERR_NULL;
    }
    if (B <= 0 || Cin % 64 != 0 || Cout % 64 != 0 || Cin > 512 || Cout > 512) {
        set_error("abl_conv3x3_igemm: need B>0, Cin and Cout multiples of 64 and <= 512");
        return ABL_ERR_SHAPE;
    }
    if (out_pool && ((H & 1) || (W & 1))) { set_error("abl_conv3x3_igemm: pooled output needs even H, W"); return ABL_ERR_SHAPE; }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_kmajor) |
         reinterpret_cast<uintptr_t>(out_full) | reinterpret_cast<uintptr_t>(out_pool) |
         reinterpret_cast<uintptr_t>(mask_src)) & 15) {
        set_error("abl_conv3x3_igemm: pointers must be 16-byte aligned");
        return ABL_ERR_ALIGN;
    }
    ConvArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    if (pick_box(H, W, &a.box_h, &a.box_w) != ABL_OK) {
        set_error("abl_conv3x3_igemm: H x W cannot be tiled into 128-pixel rectangles (need W % 8 == 0 etc.)");
        return ABL_ERR_SHAPE;
    }
    if (out_pool && (a.box_h & 1)) { set_error("abl_conv3x3_igemm: pooled output needs an even tile height"); return ABL_ERR_SHAPE; }
    const int ntile = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
    a.tiles_h = H / a.box_h; a.tiles_w = W / a.box_w;
    a.num_m_tiles = B * a.tiles_h * a.tiles_w;
    a.num_n_tiles = Cout / ntile;
    a.kchunks = Cin / kBlockK;
    a.flags = flags;
    a.bias = bias;
    a.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
    a.out_full = static_cast<__nv_bfloat16*>(out_full);
    a.out_pool = static_cast<__nv_bfloat16*>(out_pool);

    CUtensorMap ma, mb;
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)a.box_w, (uint32_t)a.box_h, 1};
        int r = encode_tiled(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != ABL_OK) return r;
    }
    {
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
        const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
        const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)ntile, 1};
        int r = encode_tiled(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w_kmajor, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != ABL_OK) return r;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (ntile) {
        case 256: return launch_conv<256>(ma, mb, a, st);
        case 128: return launch_conv<128>(ma, mb, a, st);
        default: return launch_conv<64>(ma, mb, a, st);
    }
}
