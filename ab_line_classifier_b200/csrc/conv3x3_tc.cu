// 3x3 / stride 1 / SAME convolution as a tcgen05 implicit GEMM (sm_100a), NHWC bf16 in, fp32
// accumulation in TMEM.  One kernel serves
//   * forward  (Keras Conv2D+BiasAdd+Relu [+MaxPool 2x2], reference call site models.py:412-413):
//       Y[n,h,w,co] = relu(b[co] + sum_{kh,kw,ci} X[n,h+kh-1,w+kw-1,ci] * Wf[kw*3+kh][co][ci])
//   * dgrad    (Conv2DBackpropInput + ReluGrad of the layer below): the same contraction with
//       X := dZ (Cout channels) and the tap-flipped, transposed filter Wd, epilogue = ReLU mask.
//
// Two main loops share one epilogue: conv3x3_halo_kernel (default: haloed activation tiles, stacked sub-tiles, split
// A/B rings; see the comment above it) and conv3x3_tc_kernel (v1, described next: the fallback for frame sizes the
// 16-pixel-wide tiles do not divide).  Entry points: abl_conv3x3_igemm (forward / dgrad), abl_conv3x3_igemm_gap (last
// conv + fused GlobalAveragePooling sums), abl_conv3x3_dgrad_unpool (dgrad + ReluGrad + MaxPoolGrad in the epilogue).
//
// GEMM view per CTA tile: M = 128 output pixels (a box_h x box_w rectangle of one frame),
// N = NTILE output channels, K = 9 taps x Cin.  A k-block is one tap x 64 input channels:
//   A (K-major, 128 rows x 128 B, SWIZZLE_128B) = TMA box {64 ch, box_w, box_h, 1} of the NHWC
//     tensor at (c0, w0+kw-1, h0+kh-1, n); the SAME zero padding is TMA out-of-bounds fill.
//   B (K-major, NTILE rows x 128 B)             = TMA box {64 ci, NTILE co, 1 tap} of Wf.
// Warp roles (384 threads, persistent, 1 CTA/SM): warp 0 TMA producer, warp 1 MMA issuer
// (single thread), warp 2 TMEM allocator, warps 4-11 epilogue.  The accumulator is double
// buffered in TMEM (2 x NTILE columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
// the epilogue writes full-resolution tiles with TMA stores from a swizzled staging buffer.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"
#include "parts.cuh"
#include "../../include/abl_b200.h"

namespace abl {

constexpr int ABL_MAIN_WS_PAIRS = 1 << 8;     // internal: weight-stationary MMAs in the N = 64 main loop
constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // bf16 elements per k-block row = 128 bytes
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KiB
constexpr int kEpiThreads = 256;                 // 8 epilogue warps
constexpr int kThreads = 128 + kEpiThreads;      // + producer, MMA, TMEM-alloc, spare warps

struct ConvArgs {
    int B, H, W, Cin, Cout;
    int box_h, box_w, tiles_h, tiles_w;
    int num_m_tiles, num_n_tiles, kchunks;
    int flags;
    const float* bias;              // [Cout] fp32 (ABL_EPI_BIAS_RELU)
    const __nv_bfloat16* mask_src;  // [B,H,W,Cout] (ABL_EPI_RELU_MASK)
    __nv_bfloat16* out_full;        // [B,H,W,Cout] or null
    __nv_bfloat16* out_pool;        // [B,H/2,W/2,Cout] or null
    float* colsum;                  // parts buffer (include/abl_b200.h): row blockIdx.x = this CTA's fp32 column sums of
                                    // the values it wrote (fused BiasAddGrad of the layer below) or null
    uint16_t* pool_code;            // [B,H/2,W/2,Cout/8]: 2-bit index of the window's first maximum per channel, or null
    // dgrad through a max-pool: route each output value straight to the 2x2 window element named by the forward's
    // routing code (MaxPoolGrad fused into the epilogue; the pooled-resolution gradient never reaches HBM)
    __nv_bfloat16* out_unpool;      // [B,2H,2W,Cout] or null
    const uint16_t* unpool_code;    // [B,H,W,Cout/8]
    // fused GlobalAveragePooling partial sums: row (8x16-pixel rectangle index) x Cout fp32, plain stores (each value
    // has exactly one writer -> deterministic); abl_gap_reduce adds the rectangles of a frame in fixed order
    float* gap_part;                // [B * tiles_h * MSUB * tiles_w][Cout] or null (tiles rounded UP)
    // 1-bit ReLU mask of the (bias + ReLU) output, [B,H,W,Cout/32] words, bit k of word g = channel 32 g + k is > 0; the
    // last conv of a training step writes THIS instead of its activation (nothing but the mask of it is used afterwards)
    uint32_t* relu_bits;            // or null (abl_conv3x3_igemm_gap only)
    // the frame is not a whole number of tiles (haloed kernels): TMA zero-fills the loads and clips the stores; the
    // epilogue zeroes accumulator rows that lie outside the frame and guards its non-TMA loads / stores
    int partial;
};

// 2x2 max-pool of 8 channels (four packed bf16x2 words per window element) + optional routing code: for every
// channel the 2-bit index (window scan order a, b, c, d) of the FIRST element equal to the maximum -- the element
// TF's MaxPoolGrad routes the gradient to.  Channel k of the chunk occupies bits [2k, 2k+1] of the 16-bit code.
// Values are post-ReLU (>= +0, never -0), so bit equality is value equality.
__device__ __forceinline__ void pool8(const uint4& a, const uint4& b, const uint4& c, const uint4& d, uint32_t (&o)[4],
                                      uint32_t& code, bool want_code) {
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    const uint32_t cv[4] = {c.x, c.y, c.z, c.w}, dv[4] = {d.x, d.y, d.z, d.w};
    code = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 m0 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&av[j]),
                                          *reinterpret_cast<const __nv_bfloat162*>(&bv[j]));
        const __nv_bfloat162 m1 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&cv[j]),
                                          *reinterpret_cast<const __nv_bfloat162*>(&dv[j]));
        const __nv_bfloat162 m2 = __hmax2(m0, m1);
        o[j] = *reinterpret_cast<const uint32_t*>(&m2);
        if (want_code) {
            const uint32_t ea = __vcmpeq2(av[j], o[j]), eb = __vcmpeq2(bv[j], o[j]), ec = __vcmpeq2(cv[j], o[j]);
            const uint32_t bit0 = ~ea & (eb | ~ec);      // index 1 or 3
            const uint32_t bit1 = ~ea & ~eb;             // index 2 or 3
            const uint32_t w = (bit0 & 0x00010001u) | (bit1 & 0x00020002u);   // lo channel: bits 0-1, hi channel: bits 16-17
            code |= ((w & 3u) | ((w >> 14) & 0xCu)) << (4 * j);
        }
    }
}

// Inverse of pool8 for gradients: 8 channels of a pooled gradient g (already masked by relu'(pooled)) go to the window
// element named by the 2-bit codes, zeros to the other three; writes the four full-resolution 16-byte chunks.
__device__ __forceinline__ void unpool8_store(const uint4& g, uint32_t cd, __nv_bfloat16* base, size_t right,
                                              size_t down) {
    const uint32_t gv[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t lo = (cd >> (4 * j)) & 3u, hi = (cd >> (4 * j + 2)) & 3u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k][j] = (lo == static_cast<uint32_t>(k) ? (gv[j] & 0xFFFFu) : 0u) |
                      (hi == static_cast<uint32_t>(k) ? (gv[j] & 0xFFFF0000u) : 0u);
    }
    *reinterpret_cast<uint4*>(base) = make_uint4(o[0][0], o[0][1], o[0][2], o[0][3]);
    *reinterpret_cast<uint4*>(base + right) = make_uint4(o[1][0], o[1][1], o[1][2], o[1][3]);
    *reinterpret_cast<uint4*>(base + down) = make_uint4(o[2][0], o[2][1], o[2][2], o[2][3]);
    *reinterpret_cast<uint4*>(base + down + right) = make_uint4(o[3][0], o[3][1], o[3][2], o[3][3]);
}

// ---------------------------------------------------------------------------------------------
// Epilogue of ONE 128-pixel x NTILE accumulator (shared by both main-loop variants), run by the 8
// epilogue warps (256 threads).  Warp e reads TMEM lanes 32*(e%4).. (hardware lane quarter of its warp id)
// and the column half e/4 of each pass; thread = one output pixel (tile row) x 32-column chunks.  Values are
// packed to bf16 into a 128B-swizzled staging tile ([64-ch box][128 rows][128 B], the layout of a TMA box),
// from which ONE elected thread issues the full-resolution TMA store while all threads compute the 2x2
// max-pool (if requested) out of the same staging tile.
//   tmem_acc : TMEM address (lane 0) of the accumulator's first column
//   (img,h0,w0): origin of the box_h x box_w pixel rectangle; n0 = first output channel
//   wait_bar : accumulator-full barrier to wait on before the first TMEM read (nullptr: already waited)
//   arrive_bar: accumulator-empty barrier to arrive on after the last TMEM read (nullptr: none)
//   next_origin: element offset in mask_src of the (h0, w0, channel 0) corner of the NEXT rectangle this thread group
//              will process, or -1: the ReLU-mask line each thread needs for its next pass is prefetched into L2 one
//              pass ahead (ncu: the mask load's HBM latency was ~10 % of the dgrad epilogue's stall samples)
//   gap_row  : global fp32 [Cout] row receiving this rectangle's column sums (fused GAP partial sums) or nullptr
//   csum     : shared-memory fp32 [NTILE] column sums of the bf16 values written (fused BiasAddGrad of the
//              tensor this kernel produces; the caller flushes it to global memory once per CTA) or nullptr
template <int NTILE, int kEpiN, int kStgBufs>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& p, const CUtensorMap* map_out, uint8_t* staging,
                                              uint32_t tmem_acc, int img, int h0, int w0, int n0, int box_w,
                                              int et, uint32_t& pass_ctr, uint64_t* wait_bar, uint32_t wait_phase,
                                              uint64_t* arrive_bar, float* csum = nullptr,
                                              long long next_origin = -1, float* gap_row = nullptr) {
    constexpr int kStagingBytes = (kEpiN / 64) * kABytes;
    constexpr int kPasses = NTILE / kEpiN;
    constexpr int kColsPerWarp = kEpiN / 2;      // 32 or 64
    constexpr int kChunks = kColsPerWarp / 32;   // 1 or 2
    const int lane_row = ((et >> 5) & 3) * 32 + (et & 31);   // tile row == TMEM lane
    const int hsel = et >> 7;                                // which half of the pass' columns
    const int ph_row = lane_row / box_w, pw_col = lane_row - ph_row * box_w;
    const size_t pix_off = ((static_cast<size_t>(img) * p.H + h0 + ph_row) * p.W + w0 + pw_col) * p.Cout;

#pragma unroll 1
    for (int pass = 0; pass < kPasses; ++pass, ++pass_ctr) {
        uint8_t* stg = staging + (pass_ctr % kStgBufs) * kStagingBytes;
        const int col0 = pass * kEpiN + hsel * kColsPerWarp;     // first column of this warp in the tile
        // ReLU-mask prefetch (dgrad): this thread's pixel, its columns -- issued before any wait
        uint4 mk[kChunks * 4];
        if (p.flags & ABL_EPI_RELU_MASK) {
            const uint4* mp = reinterpret_cast<const uint4*>(p.mask_src + pix_off + n0 + col0);
#pragma unroll
            for (int j = 0; j < kChunks * 4; ++j) mk[j] = __ldg(mp + j);
            // next pass: same pixel, next column block; after the last pass: same thread position in the next rectangle
            const __nv_bfloat16* nx = nullptr;
            if (pass + 1 < kPasses) nx = p.mask_src + pix_off + n0 + col0 + kEpiN;
            else if (next_origin >= 0)
                nx = p.mask_src + next_origin + (static_cast<size_t>(ph_row) * p.W + pw_col) * p.Cout + n0 + hsel * kColsPerWarp;
            if (nx != nullptr) {
#pragma unroll
                for (int j = 0; j < kChunks; ++j)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + j * 32));
            }
        }
        if (pass == 0 && wait_bar != nullptr) {
            mbar_wait(wait_bar, wait_phase);
            tc_fence_after();
        }
        // staging buffer free: the TMA store issued kStgBufs passes ago has finished READING it
        if (et == 0) tma_store_wait_read<kStgBufs - 1>();
        named_bar_sync(1, kEpiThreads);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            uint32_t r[32];
            const int col = col0 + c * 32;
            tmem_ld_32x32b_x32(tmem_acc + (static_cast<uint32_t>(lane_row & ~31) << 16) + col, r);
            tmem_ld_wait();
            uint32_t packed[16];
            if (p.flags & ABL_EPI_BIAS_RELU) {
                // warp-uniform address: one broadcast transaction per load, L1-resident
                const float4* bs = reinterpret_cast<const float4*>(p.bias + n0 + col);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = __ldg(bs + j);
                    packed[2 * j] = pack_bf16x2_relu(__uint_as_float(r[4 * j]) + b4.x,
                                                __uint_as_float(r[4 * j + 1]) + b4.y);
                    packed[2 * j + 1] = pack_bf16x2_relu(__uint_as_float(r[4 * j + 2]) + b4.z,
                                                    __uint_as_float(r[4 * j + 3]) + b4.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    packed[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            if (p.flags & ABL_EPI_RELU_MASK) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t mm[4] = {mk[c * 4 + j].x, mk[c * 4 + j].y, mk[c * 4 + j].z, mk[c * 4 + j].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // post-ReLU activations are >= 0: "> 0" <=> magnitude bits set, sign clear
                        const uint32_t lo_on = ((mm[k] & 0x7FFFu) != 0 && !(mm[k] & 0x8000u)) ? 0xFFFFu : 0u;
                        const uint32_t hi_on =
                            ((mm[k] & 0x7FFF0000u) != 0 && !(mm[k] & 0x80000000u)) ? 0xFFFF0000u : 0u;
                        packed[4 * j + k] &= (lo_on | hi_on);
                    }
                }
            }
            // swizzled store: 16-byte chunk index XOR (row & 7) inside the 64-channel box
            const int pcol = col - pass * kEpiN;              // column within the pass
            uint8_t* box = stg + (pcol >> 6) * kABytes + lane_row * 128;
            const int cc0 = (pcol & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(box + (((cc0 + j) ^ (lane_row & 7)) << 4)) =
                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        if (pass == kPasses - 1 && arrive_bar != nullptr) {
            tc_fence_before();
            mbar_arrive(arrive_bar);   // TMEM buffer may be overwritten by the next tile
        }
        const int cbase = n0 + pass * kEpiN;
        // un-pool: request this thread's routing codes of the pass now, so that the 2-byte loads overlap the barrier
        // instead of sitting one by one on the critical path of the loop below (ncu: its first use was the kernel's top
        // stall).  Slot k: q = et + 256 k -> tile row q / kCh16, chunk q % kCh16; two 16-bit codes per register.
        constexpr int kUnpoolSlots = kTileM * (kEpiN / 8) / kEpiThreads;      // 4 (kEpiN 64) or 8 (kEpiN 128)
        constexpr bool kCanUnpool = NTILE <= 128;         // the host never routes a 256-column tile through un-pool
        uint32_t ucd[kUnpoolSlots / 2];
        if (kCanUnpool && p.out_unpool != nullptr) {
#pragma unroll
            for (int k = 0; k < kUnpoolSlots; ++k) {
                const int q = et + kEpiThreads * k;
                const int r = q / (kEpiN / 8), cc = q - r * (kEpiN / 8);
                const int ph = r / box_w, pw = r - ph * box_w;
                const size_t pix = (static_cast<size_t>(img) * p.H + h0 + ph) * p.W + w0 + pw;
                const uint32_t c16 = __ldg(p.unpool_code + pix * (p.Cout >> 3) + ((cbase >> 3) + cc));
                ucd[k >> 1] = (k & 1) ? (ucd[k >> 1] | (c16 << 16)) : c16;
            }
        }
        fence_proxy_async_smem();           // generic-proxy smem writes -> visible to the TMA store
        named_bar_sync(1, kEpiThreads);     // staging filled

        if (p.out_full != nullptr && et == 0) {
#pragma unroll
            for (int j = 0; j < kEpiN / 64; ++j)
                tma_store_4d(map_out, stg + j * kABytes, cbase + j * 64, w0, h0, img);
            tma_store_commit();
        }
        if (kCanUnpool && p.out_unpool != nullptr) {
            // thread = (tile pixel, 8-channel chunk): consecutive threads cover one pixel's channels -> coalesced
            constexpr int kCh16 = kEpiN / 8;
            const size_t right = p.Cout, down = static_cast<size_t>(2 * p.W) * p.Cout;
#pragma unroll
            for (int k = 0; k < kUnpoolSlots; ++k) {
                const int q = et + kEpiThreads * k;
                const int r = q / kCh16, cc = q - r * kCh16;
                const int ph = r / box_w, pw = r - ph * box_w;
                const uint4 g = *reinterpret_cast<const uint4*>(stg + (cc >> 3) * kABytes + r * 128 + (((cc & 7) ^ (r & 7)) << 4));
                const uint32_t cd = (ucd[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                __nv_bfloat16* dst = p.out_unpool +
                    ((static_cast<size_t>(img) * 2 * p.H + 2 * (h0 + ph)) * (2 * p.W) + 2 * (w0 + pw)) * p.Cout + cbase + cc * 8;
                unpool8_store(g, cd, dst, right, down);
            }
        }
        if (csum != nullptr || gap_row != nullptr) {
            // Column sums of the staged bf16 tile.  Warp w OWNS the 16-byte chunks [w*kCh/8, (w+1)*kCh/8) (8 channels
            // each) for the whole kernel, so after a shuffle reduction over the lanes holding other row groups one
            // lane adds into csum with a plain read-modify-write: no shared-memory atomics (a float atomicAdd on
            // shared memory is a CAS loop and measured 3.5 retries per add here).
            constexpr int kCh = kEpiN / 8;                  // chunks per staged row: 16 or 8
            constexpr int kCpw = kCh / 8;                   // chunks per warp: 2 or 1
            constexpr int kRowsPer = kTileM / (32 / kCpw);  // rows per lane: 8 or 4
            const int lane = et & 31, w8 = et >> 5;
            const int chunk = w8 * kCpw + (lane & (kCpw - 1)), rg = lane / kCpw;
            const uint8_t* bx = stg + (chunk >> 3) * kABytes;
            const int c8 = chunk & 7;
            float s8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < kRowsPer; ++r) {
                const int row = rg * kRowsPer + r;
                const uint4 v = *reinterpret_cast<const uint4*>(bx + row * 128 + ((c8 ^ (row & 7)) << 4));
                s8[0] += __uint_as_float(v.x << 16); s8[1] += __uint_as_float(v.x & 0xFFFF0000u);
                s8[2] += __uint_as_float(v.y << 16); s8[3] += __uint_as_float(v.y & 0xFFFF0000u);
                s8[4] += __uint_as_float(v.z << 16); s8[5] += __uint_as_float(v.z & 0xFFFF0000u);
                s8[6] += __uint_as_float(v.w << 16); s8[7] += __uint_as_float(v.w & 0xFFFF0000u);
            }
#pragma unroll
            for (int o = 16; o >= kCpw; o >>= 1)
#pragma unroll
                for (int j = 0; j < 8; ++j) s8[j] += __shfl_xor_sync(0xffffffffu, s8[j], o);
            if (lane < kCpw) {
                if (csum != nullptr) {
                    float* dst = csum + pass * kEpiN + chunk * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] += s8[j];
                }
                if (gap_row != nullptr) {
                    float4* dst = reinterpret_cast<float4*>(gap_row + cbase + chunk * 8);
                    dst[0] = make_float4(s8[0], s8[1], s8[2], s8[3]);
                    dst[1] = make_float4(s8[4], s8[5], s8[6], s8[7]);
                }
            }
        }
        if (p.out_pool != nullptr) {
            constexpr int kChunks16 = kEpiN / 8;
            const int pw_n = box_w >> 1;
#pragma unroll
            for (int q = et; q < (kTileM / 4) * kChunks16; q += kEpiThreads) {
                const int prow = q / kChunks16, cc = q - prow * kChunks16;
                const int ph = prow / pw_n, pw = prow - ph * pw_n;
                const int r00 = (2 * ph) * box_w + 2 * pw;          // even row: r00 & 7 in {0,2,4,6}
                const uint8_t* bx = stg + (cc >> 3) * kABytes;
                const int c8 = cc & 7;
                const int r01 = r00 + 1, r10 = r00 + box_w, r11 = r10 + 1;
                const uint4 a = *reinterpret_cast<const uint4*>(bx + r00 * 128 + ((c8 ^ (r00 & 7)) << 4));
                const uint4 b = *reinterpret_cast<const uint4*>(bx + r01 * 128 + ((c8 ^ (r01 & 7)) << 4));
                const uint4 c2 = *reinterpret_cast<const uint4*>(bx + r10 * 128 + ((c8 ^ (r10 & 7)) << 4));
                const uint4 d = *reinterpret_cast<const uint4*>(bx + r11 * 128 + ((c8 ^ (r11 & 7)) << 4));
                uint32_t o[4], code;
                pool8(a, b, c2, d, o, code, p.pool_code != nullptr);
                const int h = (h0 >> 1) + ph, w = (w0 >> 1) + pw;
                const size_t off =
                    ((static_cast<size_t>(img) * (p.H >> 1) + h) * (p.W >> 1) + w) * p.Cout + cbase + cc * 8;
                *reinterpret_cast<uint4*>(p.out_pool + off) = make_uint4(o[0], o[1], o[2], o[3]);
                if (p.pool_code != nullptr) p.pool_code[off >> 3] = static_cast<uint16_t>(code);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-local epilogue of ONE 8 x 16-pixel sub-tile (128 accumulator rows x NTILE columns) for the haloed kernels with
// N = 128 / 256.  Warp e of the 8 epilogue warps owns TMEM lanes 32*(e%4).. = two image rows of the sub-tile, and the
// 64-column half e/4 of each 128-column pass: a 32-row x 128-byte slice of the swizzled staging area (4 KiB, exactly the
// layout of a TMA box {64 channels, 16 pixels, 2 rows}).  The warp packs its slice, issues ITS OWN TMA store and does
// the 2x2 max-pool / the un-pool scatter / the column sums of those two rows by itself -- a pooling window or an
// un-pool target never leaves two image rows -- so the two 256-thread named barriers per pass of epilogue_tile (ncu
// round 2: 17-22 % of the stall samples of the N = 128 kernels) become warp-level synchronisation.
//   wsum   : this lane's running column sums (channels 2*lane, 2*lane+1 of the warp's 64 columns, per pass) for the
//            fused bias gradient; the caller combines the four row-pair warps of a column in fixed order at the end
//   gsum   : shared fp32 [4][128] exchange for the fused GlobalAveragePooling sums (the only cross-warp step: one
//            named barrier pair per pass, in the last conv's forward only)
template <int NTILE, bool PARTIAL>
__device__ __forceinline__ void epilogue_tile_wl(const ConvArgs& p, const CUtensorMap* map_out, uint8_t* staging,
                                                 uint32_t tmem_acc, int img, int h0, int w0, int n0, int et,
                                                 uint64_t* wait_bar, uint32_t wait_phase, uint64_t* arrive_bar,
                                                 float (&wsum)[NTILE / 128][2], long long next_origin, float* gap_row,
                                                 float* gsum) {
    constexpr int kEpiN = 128, kPasses = NTILE / kEpiN, kBoxW = 16;
    const int lane = et & 31, quarter = (et >> 5) & 3, hsel = et >> 7;
    const int lane_row = quarter * 32 + lane;                 // tile row == TMEM lane
    const int ph_row = lane_row >> 4, pw_col = lane_row & 15;
    const int hw0 = h0 + 2 * quarter;                         // first image row of this warp
    const size_t pix_off = ((static_cast<size_t>(img) * p.H + h0 + ph_row) * p.W + w0 + pw_col) * p.Cout;
    uint8_t* stg_w = staging + hsel * kABytes + quarter * 4096;       // this warp's 32 rows of 64-channel box `hsel`
    const bool valid = !PARTIAL || (h0 + ph_row < p.H && w0 + pw_col < p.W);      // this lane's pixel lies in the frame

#pragma unroll 1
    for (int pass = 0; pass < kPasses; ++pass) {
        const int cw0 = n0 + pass * kEpiN + hsel * 64;        // first output channel of this warp in this pass
        uint4 mk[8];
        if (p.flags & ABL_EPI_RELU_MASK) {
            const uint4* mp = reinterpret_cast<const uint4*>(p.mask_src + pix_off + cw0);
#pragma unroll
            for (int j = 0; j < 8; ++j) mk[j] = valid ? __ldg(mp + j) : make_uint4(0u, 0u, 0u, 0u);
            const __nv_bfloat16* nx = nullptr;       // next pass: same pixel, next column block; then the next rectangle
            if (pass + 1 < kPasses) nx = p.mask_src + pix_off + cw0 + kEpiN;
            else if (next_origin >= 0)
                nx = p.mask_src + next_origin + (static_cast<size_t>(ph_row) * p.W + pw_col) * p.Cout + n0 + hsel * 64;
            if (nx != nullptr && !PARTIAL) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32));
            }
        }
        if (pass == 0 && wait_bar != nullptr) {
            mbar_wait(wait_bar, wait_phase);
            tc_fence_after();
        }
        if (lane == 0) tma_store_wait_read<0>();     // this warp's previous store has finished reading its slice
        __syncwarp();                                // ... and every lane is done with the slice
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_acc + (static_cast<uint32_t>(lane_row & ~31) << 16) + pass * kEpiN + hsel * 64 + c * 32, r);
            tmem_ld_wait();
            uint32_t packed[16];
            if (NTILE == 256 && (p.flags & ABL_EPI_BIAS_RELU) && p.relu_bits != nullptr) {      // (N = 256 kernels only)
                // as below, and the signs are collected on the way: u | (u - 1) has its top bit set iff the float with
                // bits u is <= 0 (either zero included); one funnel shift per channel appends it, channel 0 first
                const float4* bs = reinterpret_cast<const float4*>(p.bias + cw0 + c * 32);
                uint32_t off = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = __ldg(bs + j);
                    const float v[4] = {__uint_as_float(r[4 * j]) + b4.x, __uint_as_float(r[4 * j + 1]) + b4.y,
                                        __uint_as_float(r[4 * j + 2]) + b4.z, __uint_as_float(r[4 * j + 3]) + b4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t u = __float_as_uint(v[k]);
                        off = __funnelshift_l(u | (u - 1u), off, 1);
                    }
                    packed[2 * j] = pack_bf16x2_relu(v[0], v[1]);
                    packed[2 * j + 1] = pack_bf16x2_relu(v[2], v[3]);
                }
                if (valid) p.relu_bits[(pix_off + cw0 + c * 32) >> 5] = ~__brev(off);
            } else if (p.flags & ABL_EPI_BIAS_RELU) {
                const float4* bs = reinterpret_cast<const float4*>(p.bias + cw0 + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = __ldg(bs + j);
                    packed[2 * j] = pack_bf16x2_relu(__uint_as_float(r[4 * j]) + b4.x,
                                                __uint_as_float(r[4 * j + 1]) + b4.y);
                    packed[2 * j + 1] = pack_bf16x2_relu(__uint_as_float(r[4 * j + 2]) + b4.z,
                                                    __uint_as_float(r[4 * j + 3]) + b4.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    packed[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            if (p.flags & ABL_EPI_RELU_MASK) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t mm[4] = {mk[c * 4 + j].x, mk[c * 4 + j].y, mk[c * 4 + j].z, mk[c * 4 + j].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t lo_on = ((mm[k] & 0x7FFFu) != 0 && !(mm[k] & 0x8000u)) ? 0xFFFFu : 0u;
                        const uint32_t hi_on =
                            ((mm[k] & 0x7FFF0000u) != 0 && !(mm[k] & 0x80000000u)) ? 0xFFFF0000u : 0u;
                        packed[4 * j + k] &= (lo_on | hi_on);
                    }
                }
            } else if (!valid) {
                // a row outside the frame holds sums over real neighbours: zero it, so that pooling, column sums and the
                // GAP sums see nothing (the TMA store clips it anyway)
#pragma unroll
                for (int j = 0; j < 16; ++j) packed[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(stg_w + lane * 128 + (((c * 4 + j) ^ (lane & 7)) << 4)) =
                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        if (pass == kPasses - 1 && arrive_bar != nullptr) {
            tc_fence_before();
            mbar_arrive(arrive_bar);   // TMEM buffer may be overwritten by the next tile
        }
        // un-pool: all 8 routing codes of this lane requested before the first is used (slot k: slice row
        // r = (lane >> 3) + 4 k, 8-channel chunk cc = lane & 7 of the warp's 64 columns)
        uint32_t ucd[4] = {0u, 0u, 0u, 0u};
        if (NTILE <= 128 && p.out_unpool != nullptr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = (lane >> 3) + 4 * k;
                if (PARTIAL && (hw0 + (r >> 4) >= p.H || w0 + (r & 15) >= p.W)) continue;
                const size_t pix = (static_cast<size_t>(img) * p.H + hw0 + (r >> 4)) * p.W + w0 + (r & 15);
                const uint32_t c16 = __ldg(p.unpool_code + pix * (p.Cout >> 3) + ((cw0 >> 3) + (lane & 7)));
                ucd[k >> 1] |= c16 << (16 * (k & 1));
            }
        }
        fence_proxy_async_smem();           // generic-proxy smem writes -> visible to the TMA store
        __syncwarp();                       // slice filled
        if (p.out_full != nullptr && lane == 0) {
            tma_store_4d(map_out, stg_w, cw0, w0, hw0, img);       // box = 64 channels x 16 pixels x 2 rows
            tma_store_commit();
        }
        if (NTILE <= 128 && p.out_unpool != nullptr) {
            const size_t right = p.Cout, down = static_cast<size_t>(2 * p.W) * p.Cout;
#pragma unroll 2
            for (int k = 0; k < 8; ++k) {
                const int r = (lane >> 3) + 4 * k, cc = lane & 7;
                const uint4 g = *reinterpret_cast<const uint4*>(stg_w + r * 128 + ((cc ^ (r & 7)) << 4));
                const int hh = hw0 + (r >> 4), ww = w0 + (r & 15);
                if (PARTIAL && (hh >= p.H || ww >= p.W)) continue;
                const uint32_t cw = k < 4 ? (k < 2 ? ucd[0] : ucd[1]) : (k < 6 ? ucd[2] : ucd[3]);
                const uint32_t cd = (cw >> (16 * (k & 1))) & 0xFFFFu;
                __nv_bfloat16* dst = p.out_unpool +
                    ((static_cast<size_t>(img) * 2 * p.H + 2 * hh) * (2 * p.W) + 2 * ww) * p.Cout + cw0 + cc * 8;
                unpool8_store(g, cd, dst, right, down);
            }
        }
        if (p.colsum != nullptr || gap_row != nullptr) {
            // column sums of this warp's 32 rows: lane = one packed word (2 channels) of every row (conflict-free: the
            // 32 lanes read the 32 words of one 128-byte row)
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(stg_w + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
                s0 += __uint_as_float(v << 16);
                s1 += __uint_as_float(v & 0xFFFF0000u);
            }
            if (p.colsum != nullptr) { wsum[pass][0] += s0; wsum[pass][1] += s1; }
            if (gap_row != nullptr) {
                // GlobalAveragePooling partial sums of the whole 8 x 16 rectangle: the four row-pair warps of a column
                // meet in shared memory and are added in fixed order (one writer per value -> deterministic)
                gsum[quarter * 128 + hsel * 64 + 2 * lane] = s0;
                gsum[quarter * 128 + hsel * 64 + 2 * lane + 1] = s1;
                named_bar_sync(1, kEpiThreads);
                if (et < kEpiN)
                    gap_row[n0 + pass * kEpiN + et] = (gsum[et] + gsum[128 + et]) + (gsum[256 + et] + gsum[384 + et]);
                named_bar_sync(1, kEpiThreads);
            }
        }
        if (p.out_pool != nullptr) {
            // the warp's two image rows hold 8 complete 2x2 windows x 8 chunks of 8 channels: 2 items per lane
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int q = lane + 32 * k;
                const int pw = q >> 3, c8 = q & 7;
                const int r00 = 2 * pw;
                const int r01 = r00 + 1, r10 = r00 + kBoxW, r11 = r10 + 1;
                const uint4 a = *reinterpret_cast<const uint4*>(stg_w + r00 * 128 + ((c8 ^ (r00 & 7)) << 4));
                const uint4 b = *reinterpret_cast<const uint4*>(stg_w + r01 * 128 + ((c8 ^ (r01 & 7)) << 4));
                const uint4 c2 = *reinterpret_cast<const uint4*>(stg_w + r10 * 128 + ((c8 ^ (r10 & 7)) << 4));
                const uint4 d = *reinterpret_cast<const uint4*>(stg_w + r11 * 128 + ((c8 ^ (r11 & 7)) << 4));
                uint32_t o[4], code;
                pool8(a, b, c2, d, o, code, p.pool_code != nullptr);
                if (PARTIAL && (hw0 + 1 >= p.H || w0 + 2 * pw + 1 >= p.W)) continue;      // only complete windows (Keras 'valid' pooling floors odd sides)
                const size_t off = ((static_cast<size_t>(img) * (p.H >> 1) + (hw0 >> 1)) * (p.W >> 1) + (w0 >> 1) + pw) * p.Cout +
                                   cw0 + c8 * 8;
                *reinterpret_cast<uint4*>(p.out_pool + off) = make_uint4(o[0], o[1], o[2], o[3]);
                if (p.pool_code != nullptr) p.pool_code[off >> 3] = static_cast<uint16_t>(code);
            }
        }
    }
}

// TAPS = taps consumed per pipeline stage.  TAPS == 1: stage = one tap x 64 channels (A 16 KiB + B).
// TAPS == 3 (used for NTILE == 64, where 4 MMAs per stage cannot amortise the issuing thread's barrier wait
// and the TMA instruction rate): tile = 4 x 32 pixels, stage = one kw x 64 channels; A is the tile plus one
// halo row above and below (6 x 32 pixels, 24 KiB) and serves kh = 0,1,2 through descriptor start offsets of
// kh * 32 rows (4 KiB, a multiple of the 1 KiB swizzle atom); B = the 3 kh taps of that kw (filter operand
// is stored kw-major: tap' = kw*3 + kh).
template <int NTILE, int TAPS>
struct ConvCfg {
    static constexpr int kAStage = TAPS == 3 ? 6 * 32 * 128 : kABytes;
    static constexpr int kBTap = NTILE * kBlockK * 2;
    static constexpr int kBStage = TAPS * kBTap;
    static constexpr int kStageBytes = kAStage + kBStage;
    static constexpr int kEpiN = NTILE < 128 ? NTILE : 128;        // columns staged per epilogue pass
    static constexpr int kStagingBytes = (kEpiN / 64) * kABytes;   // [kEpiN/64 boxes][128 rows][128 B], swizzled
    static constexpr int kStgBufs = NTILE == 256 ? 1 : 2;          // double-buffered where shared memory allows
    static constexpr int kStages = (227 * 1024 - 256 - kStgBufs * kStagingBytes) / kStageBytes > 6
                                       ? 6 : (227 * 1024 - 256 - kStgBufs * kStagingBytes) / kStageBytes;
    static constexpr int kTmemCols = 2 * NTILE;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStgBufs * kStagingBytes + 256;
};

template <int NTILE, int TAPS>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_out, const ConvArgs p) {
    using Cfg = ConvCfg<NTILE, TAPS>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kStages * Cfg::kAStage;
    uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStgBufs * Cfg::kStagingBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::kStages;
    uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int kblocks = (9 / TAPS) * p.kchunks;     // pipeline stages per tile

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
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
                for (int t = 0; t < 9 / TAPS; ++t) {
                    // filter taps are stored kw-major (tap' = kw*3 + kh): TAPS == 3 -> t = kw, all kh at once
                    const int kw = TAPS == 3 ? t : t / 3;
                    const int kh = TAPS == 3 ? 0 : t - kw * 3;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        tma_load_4d(&map_a, &full_bar[stage], smem_a + stage * Cfg::kAStage, kc * kBlockK,
                                    w0 + kw - 1, h0 + kh - 1, img);
                        tma_load_3d(&map_b, &full_bar[stage], smem_b + stage * Cfg::kBStage, kc * kBlockK, n0,
                                    t * TAPS);
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kTileM, NTILE, 0, 0);
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a), 16);
            const uint32_t b_lo0 = desc_lo_sw128(smem_u32(smem_b), 16);
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
                    const uint32_t a_lo = a_lo0 + stage * (Cfg::kAStage >> 4);
                    const uint32_t b_lo = b_lo0 + stage * (Cfg::kBStage >> 4);
#pragma unroll
                    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            // kh shift = 32 tile rows = 4096 B; K advance of 16 elements = 32 B (>> 4 in the lo word)
                            umma_bf16_lohi(tmem_d, a_lo + t * (4096 >> 4) + k * 2, dhi,
                                           b_lo + t * (Cfg::kBTap >> 4) + k * 2, dhi, idesc,
                                           (kb | t | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);          // accumulator ready for the epilogue
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (8 warps, 256 threads)
        const int et = threadIdx.x - 128;                 // 0..255
        int it = 0;
        uint32_t pass_ctr = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int n_tile = tile % p.num_n_tiles;
            int m_tile = tile / p.num_n_tiles;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int img = m_tile / p.tiles_h;
            const int as = it & 1;
            epilogue_tile<NTILE, Cfg::kEpiN, Cfg::kStgBufs>(p, &map_out, staging, tmem_base + as * NTILE, img,
                                                            th * p.box_h, tw * p.box_w, n_tile * NTILE, p.box_w, et,
                                                            pass_ctr, &tfull_bar[as], (it >> 1) & 1, &tempty_bar[as]);
        }
        if (et == 0) tma_store_wait_all();   // all output tiles written before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// "Halo" main loop (v2).  ncu on the v1 kernel above showed every variant saturating the L2 -> SM
// fabric (l1tex__m_xbar2l1tex_read_bytes ~17 TB/s) with the tensor pipe 50-85 % busy: each CTA re-fetched
// its activation tile once per tap (9x) and the filter slab once per 128 pixels.  This variant cuts both:
//   * CTA tile = MSUB vertically stacked sub-tiles of 8 rows x 16 pixels (M = 128 each) that share every
//     filter slab, accumulated in MSUB TMEM accumulators (x2 buffers: 2*MSUB*NTILE <= 512 columns);
//   * per (kw, 64-channel chunk) ONE TMA box brings the tile plus a one-row halo above and below
//     ((8*MSUB+2) x 16 pixels); the three kh taps of every sub-tile are descriptor start offsets of
//     (8*sub + kh) image rows = multiples of 2 KiB (two 1 KiB swizzle atoms), so A is fetched 3x, not 9x;
//   * A and B travel through separate rings with separate producer threads (warp 0: A, warp 3: B), so the
//     small filter slabs prefetch deep while few big A stages are in flight.
// L2 -> SM bytes per MMA clock: 96 -> 75 (N=256), 128 -> 55 (N=128, MSUB=2), 128 -> 78 (N=64, MSUB=2).
// BRES (N = 64, Cin = 64 only): the whole filter (9 taps x 64 x 64 bf16 = 72 KiB) stays resident in shared memory
// for the life of the CTA instead of cycling through the ring: zero filter traffic after the first 3 boxes.
template <int NTILE, int MSUB, bool BRES = false>
struct HaloCfg {
    static constexpr int kSubH = 8, kTileW = 16;
    static constexpr int kARows = (kSubH * MSUB + 2) * kTileW;          // haloed pixel rows of one A stage
    static constexpr int kAStage = kARows * 128;
    static constexpr int kTPB = NTILE == 64 ? 3 : 1;                    // taps per B stage
    static constexpr int kBTap = NTILE * kBlockK * 2;
    static constexpr int kBStage = kTPB * kBTap;
    static constexpr int kEpiN = NTILE < 128 ? NTILE : 128;
    static constexpr int kStagingBytes = (kEpiN / 64) * kABytes;
    static constexpr int kStgBufs = NTILE == 64 ? 2 : 1;
    static constexpr int kAStages = 3;
    static constexpr int kBarBytes = 4096;     // barriers (first 512 B) + fp32 [4][128] exchange of the fused GAP sums
    static constexpr int kBRoom = (227 * 1024 - kBarBytes - kStgBufs * kStagingBytes - kAStages * kAStage) / kBStage;
    static constexpr int kBStages = BRES ? 3 : (kBRoom > 8 ? 8 : kBRoom);     // BRES: 3 "stages" = the 3 kw groups
    static_assert(!BRES || (NTILE == 64 && kBRoom >= 3), "resident filter needs N = 64");
    static constexpr int kTmemCols = 2 * MSUB * NTILE;
    static constexpr int kSmemBytes = kAStages * kAStage + kBStages * kBStage + kStgBufs * kStagingBytes + kBarBytes;
    static_assert(kBStages >= 3, "filter ring too shallow");
    static_assert(kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns must be a power of two");
};

// PARTIAL: the frame is not a whole number of tiles (ConvArgs::partial).  A template parameter, not a run-time branch:
// with the guards compiled into the exact-tiling kernels the two un-pooling dgrads ran 15 % slower (131 -> 154 us and
// 196 -> 225 us at batch 256).
template <int NTILE, int MSUB, bool BRES, bool PARTIAL>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_out, const ConvArgs p) {
    using Cfg = HaloCfg<NTILE, MSUB, BRES>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kAStages * Cfg::kAStage;
    uint8_t* staging = smem_b + Cfg::kBStages * Cfg::kBStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStgBufs * Cfg::kStagingBytes);
    uint64_t* afull_bar = bars;
    uint64_t* aempty_bar = afull_bar + Cfg::kAStages;
    uint64_t* bfull_bar = aempty_bar + Cfg::kAStages;
    uint64_t* bempty_bar = bfull_bar + Cfg::kBStages;
    uint64_t* tfull_bar = bempty_bar + Cfg::kBStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* csum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

    const int warp = threadIdx.x >> 5;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int kwkc = 3 * p.kchunks;                 // A stages per tile: (kw, 64-channel chunk)
    constexpr int kTileH = Cfg::kSubH * MSUB;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < Cfg::kAStages; ++s) {
            mbar_init(&afull_bar[s], 1);
            mbar_init(&aempty_bar[s], 1);
        }
        for (int s = 0; s < Cfg::kBStages; ++s) {
            mbar_init(&bfull_bar[s], 1);
            mbar_init(&bempty_bar[s], 1);
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
        // ------------------------------------------------------------ A producer (haloed activation tiles)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_tile = tile / p.num_n_tiles;
                const int tw = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th = m_tile % p.tiles_h;
                const int img = m_tile / p.tiles_h;
                const int h0 = th * kTileH, w0 = tw * Cfg::kTileW;
                for (int kw = 0; kw < 3; ++kw) {
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&aempty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&afull_bar[stage], Cfg::kAStage);
                        tma_load_4d(&map_a, &afull_bar[stage], smem_a + stage * Cfg::kAStage, kc * kBlockK,
                                    w0 + kw - 1, h0 - 1, img);
                        if (++stage == Cfg::kAStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ B producer (filter slabs, kw-major taps)
        if (BRES) {
            // resident filter: kchunks == 1, one N tile -> stage kw holds taps kw*3 .. kw*3+2, loaded once
            if (elect_one() && blockIdx.x < num_tiles) {
                for (int kw = 0; kw < 3; ++kw) {
                    mbar_expect_tx(&bfull_bar[kw], Cfg::kBStage);
                    tma_load_3d(&map_b, &bfull_bar[kw], smem_b + kw * Cfg::kBStage, 0, 0, kw * 3);
                }
            }
        } else if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n0 = (tile % p.num_n_tiles) * NTILE;
                for (int kw = 0; kw < 3; ++kw) {
                    for (int kc = 0; kc < p.kchunks; ++kc) {
#pragma unroll
                        for (int g = 0; g < 3 / Cfg::kTPB; ++g) {
                            mbar_wait(&bempty_bar[stage], phase ^ 1);
                            mbar_expect_tx(&bfull_bar[stage], Cfg::kBStage);
                            tma_load_3d(&map_b, &bfull_bar[stage], smem_b + stage * Cfg::kBStage, kc * kBlockK, n0,
                                        kw * 3 + g * Cfg::kTPB);
                            if (++stage == Cfg::kBStages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kTileM, NTILE, 0, 0);
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a), 16);
            const uint32_t b_lo0 = desc_lo_sw128(smem_u32(smem_b), 16);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * (MSUB * NTILE);
                for (int s = 0; s < kwkc; ++s) {
                    mbar_wait(&afull_bar[sa], pa);
                    const uint32_t a_lo = a_lo0 + sa * (Cfg::kAStage >> 4);
#pragma unroll
                    for (int g = 0; g < 3 / Cfg::kTPB; ++g) {
                        if (BRES) {
                            if (it == 0) mbar_wait(&bfull_bar[s], 0);     // s == kw (kchunks == 1): loaded once
                            sb = s;
                        } else {
                            mbar_wait(&bfull_bar[sb], pb);
                        }
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + sb * (Cfg::kBStage >> 4);
                        if (NTILE <= 128 && MSUB == 2 && (p.flags & ABL_MAIN_WS_PAIRS)) {
                            // weight-stationary MMA pairs: the two sub-tiles share one shared-memory read of B
#pragma unroll
                            for (int t = 0; t < Cfg::kTPB; ++t) {
                                const int kh = g * Cfg::kTPB + t;
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k) {
                                    umma_bf16_lohi_ws<0>(tmem_d, a_lo + kh * (2048 >> 4) + k * 2, dhi,
                                                         b_lo + t * (Cfg::kBTap >> 4) + k * 2, dhi, idesc,
                                                         (s | kh | k) != 0 ? 1u : 0u);
                                    umma_bf16_lohi_ws<1>(tmem_d + NTILE, a_lo + (Cfg::kSubH + kh) * (2048 >> 4) + k * 2, dhi,
                                                         b_lo + t * (Cfg::kBTap >> 4) + k * 2, dhi, idesc,
                                                         (s | kh | k) != 0 ? 1u : 0u);
                                }
                            }
                        } else {
#pragma unroll
                        for (int t = 0; t < Cfg::kTPB; ++t) {
                            const int kh = g * Cfg::kTPB + t;
#pragma unroll
                            for (int sub = 0; sub < MSUB; ++sub) {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k) {
                                    // (8*sub + kh) image rows of 16 pixels = 2048 B each; K advance of 16 elements = 32 B
                                    umma_bf16_lohi(tmem_d + sub * NTILE,
                                                   a_lo + (Cfg::kSubH * sub + kh) * (2048 >> 4) + k * 2, dhi,
                                                   b_lo + t * (Cfg::kBTap >> 4) + k * 2, dhi, idesc,
                                                   (s | kh | k) != 0 ? 1u : 0u);
                                }
                            }
                        }
                        }
                        if (!BRES) {
                            umma_commit(&bempty_bar[sb]);     // filter slab free when these MMAs retire
                            if (++sb == Cfg::kBStages) { sb = 0; pb ^= 1; }
                        }
                    }
                    umma_commit(&aempty_bar[sa]);         // haloed activation tile free
                    if (++sa == Cfg::kAStages) { sa = 0; pa ^= 1; }
                }
                umma_commit(&tfull_bar[as]);              // all MSUB accumulators ready for the epilogue
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (8 warps, 256 threads)
        const int et = threadIdx.x - 128;
        if constexpr (NTILE == 64 && MSUB == 2) {
            // N = 64 is epilogue-bound (1152 MMA clocks per sub-tile; ncu round 2: 23 % of the dgrad kernel's stall
            // samples sat on the two per-tile named barriers + proxy fence of a 4-warp group, 24 % of the un-pool
            // kernel's on the dependent 2-byte routing-code loads).  So every WARP drains its own 32 accumulator rows
            // = 2 image rows x 16 pixels of a sub-tile on its own: private 4 KiB staging slice, its own TMA store
            // (box of 2 image rows), warp-level synchronisation only -- a 2x2 pooling window and an un-pool target
            // never leave those two rows.  Warps 0-3 take sub-tile 0, warps 4-7 sub-tile 1.  The ReLU mask of the
            // NEXT tile is prefetched while this one is processed; the routing codes of a tile are all requested
            // before the first one is used.
            const int grp = et >> 7, gt = et & 127;
            const int lane = et & 31, wq = gt >> 5;              // warp within the group = image-row pair of the sub-tile
            uint8_t* stg = staging + grp * Cfg::kStagingBytes + wq * 4096;      // this warp's 32 rows x 128 B
            uint4 mk[8];
            auto tile_origin = [&](int tile, int& img, int& h0, int& w0, int& n0) {
                n0 = (tile % p.num_n_tiles) * NTILE;
                int m_tile = tile / p.num_n_tiles;
                w0 = (m_tile % p.tiles_w) * Cfg::kTileW;
                m_tile /= p.tiles_w;
                h0 = (m_tile % p.tiles_h) * kTileH + grp * Cfg::kSubH + 2 * wq;     // first image row of this WARP
                img = m_tile / p.tiles_h;
            };
            auto load_mask = [&](int tile) {
                int img, h0, w0, n0;
                tile_origin(tile, img, h0, w0, n0);
                const bool ok = !PARTIAL || (h0 + (lane >> 4) < p.H && w0 + (lane & 15) < p.W);
                const size_t off = ((static_cast<size_t>(img) * p.H + h0 + (lane >> 4)) * p.W + w0 + (lane & 15)) * p.Cout + n0;
                const uint4* mp = reinterpret_cast<const uint4*>(p.mask_src + off);
#pragma unroll
                for (int j = 0; j < 8; ++j) mk[j] = ok ? __ldg(mp + j) : make_uint4(0u, 0u, 0u, 0u);
            };
            const bool masked = (p.flags & ABL_EPI_RELU_MASK) != 0;
            float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // fused bias gradient of the un-pooled tensor
            if (masked && blockIdx.x < num_tiles) load_mask(blockIdx.x);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                int img, h0, w0, n0;
                tile_origin(tile, img, h0, w0, n0);
                const int as = it & 1;
                mbar_wait(&tfull_bar[as], (it >> 1) & 1);
                tc_fence_after();
                if (lane == 0) tma_store_wait_read<0>();    // this warp's previous store has finished reading its slice
                __syncwarp();                               // ... and every lane is done with the previous tile's slice
                const uint32_t tacc = tmem_base + (as * MSUB + grp) * NTILE + (static_cast<uint32_t>(gt & ~31) << 16);
                uint32_t packed[2][16];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tacc + c * 32, r);
                    tmem_ld_wait();
                    if (p.flags & ABL_EPI_BIAS_RELU) {
                        const float4* bs = reinterpret_cast<const float4*>(p.bias + n0 + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(bs + j);
                            packed[c][2 * j] = pack_bf16x2_relu(__uint_as_float(r[4 * j]) + b4.x,
                                                           __uint_as_float(r[4 * j + 1]) + b4.y);
                            packed[c][2 * j + 1] = pack_bf16x2_relu(__uint_as_float(r[4 * j + 2]) + b4.z,
                                                               __uint_as_float(r[4 * j + 3]) + b4.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            packed[c][j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[as]);               // both accumulators of this buffer drained by 256 arrivals
                if (masked) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t mm[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t lo_on = ((mm[k] & 0x7FFFu) != 0 && !(mm[k] & 0x8000u)) ? 0xFFFFu : 0u;
                            const uint32_t hi_on =
                                ((mm[k] & 0x7FFF0000u) != 0 && !(mm[k] & 0x80000000u)) ? 0xFFFF0000u : 0u;
                            packed[j >> 2][(j & 3) * 4 + k] &= (lo_on | hi_on);
                        }
                    }
                    if (tile + static_cast<int>(gridDim.x) < num_tiles) load_mask(tile + gridDim.x);   // next tile's mask
                } else if (PARTIAL && !(h0 + (lane >> 4) < p.H && w0 + (lane & 15) < p.W)) {
                    // a row outside the frame: zero it for the pooling below (the TMA store clips it anyway)
#pragma unroll
                    for (int j = 0; j < 16; ++j) { packed[0][j] = 0u; packed[1][j] = 0u; }
                }
                uint8_t* row = stg + lane * 128;            // slice row = lane: pixel (lane >> 4, lane & 15) of the row pair
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(row + ((j ^ (lane & 7)) << 4)) =
                        make_uint4(packed[j >> 2][(j & 3) * 4], packed[j >> 2][(j & 3) * 4 + 1],
                                   packed[j >> 2][(j & 3) * 4 + 2], packed[j >> 2][(j & 3) * 4 + 3]);
                // un-pool: request all 8 routing codes of this lane now (`packed` is dead: no extra register pressure)
                // -- slot k: slice row r = (lane >> 3) + 4 k, chunk cc = lane & 7 (the lane's chunk for the whole kernel)
                uint32_t ucd[4] = {0u, 0u, 0u, 0u};
                if (p.out_unpool != nullptr) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int r = (lane >> 3) + 4 * k;
                        if (PARTIAL && (h0 + (r >> 4) >= p.H || w0 + (r & 15) >= p.W)) continue;
                        const size_t pix = (static_cast<size_t>(img) * p.H + h0 + (r >> 4)) * p.W + w0 + (r & 15);
                        const uint32_t c16 = __ldg(p.unpool_code + pix * (p.Cout >> 3) + ((n0 >> 3) + (lane & 7)));
                        ucd[k >> 1] |= c16 << (16 * (k & 1));
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (p.out_full != nullptr && lane == 0) {
                    tma_store_4d(&map_out, stg, n0, w0, h0, img);      // box = 64 channels x 16 pixels x 2 rows
                    tma_store_commit();
                }
                if (p.out_unpool != nullptr) {
                    const size_t right = p.Cout, down = static_cast<size_t>(2 * p.W) * p.Cout;
#pragma unroll 2
                    for (int k = 0; k < 8; ++k) {
                        const int r = (lane >> 3) + 4 * k, cc = lane & 7;
                        const uint4 g = *reinterpret_cast<const uint4*>(stg + r * 128 + ((cc ^ (r & 7)) << 4));
                        const int hh = h0 + (r >> 4), ww = w0 + (r & 15);
                        if (PARTIAL && (hh >= p.H || ww >= p.W)) continue;
                        const uint32_t cw = k < 4 ? (k < 2 ? ucd[0] : ucd[1]) : (k < 6 ? ucd[2] : ucd[3]);
                        const uint32_t cd = (cw >> (16 * (k & 1))) & 0xFFFFu;
                        __nv_bfloat16* dst = p.out_unpool +
                            ((static_cast<size_t>(img) * 2 * p.H + 2 * hh) * (2 * p.W) + 2 * ww) * p.Cout + n0 + cc * 8;
                        unpool8_store(g, cd, dst, right, down);
                        if (p.colsum != nullptr) {
                            bacc[0] += __uint_as_float(g.x << 16); bacc[1] += __uint_as_float(g.x & 0xFFFF0000u);
                            bacc[2] += __uint_as_float(g.y << 16); bacc[3] += __uint_as_float(g.y & 0xFFFF0000u);
                            bacc[4] += __uint_as_float(g.z << 16); bacc[5] += __uint_as_float(g.z & 0xFFFF0000u);
                            bacc[6] += __uint_as_float(g.w << 16); bacc[7] += __uint_as_float(g.w & 0xFFFF0000u);
                        }
                    }
                }
                if (p.out_pool != nullptr) {
                    // the warp's two image rows hold 8 complete 2x2 windows: (window, chunk) = 64 items, 2 per lane
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int q = lane + 32 * k;
                        const int pw = q >> 3, c8 = q & 7;
                        const int r00 = 2 * pw;
                        const int r01 = r00 + 1, r10 = r00 + 16, r11 = r10 + 1;
                        const uint4 a = *reinterpret_cast<const uint4*>(stg + r00 * 128 + ((c8 ^ (r00 & 7)) << 4));
                        const uint4 b = *reinterpret_cast<const uint4*>(stg + r01 * 128 + ((c8 ^ (r01 & 7)) << 4));
                        const uint4 c2 = *reinterpret_cast<const uint4*>(stg + r10 * 128 + ((c8 ^ (r10 & 7)) << 4));
                        const uint4 d = *reinterpret_cast<const uint4*>(stg + r11 * 128 + ((c8 ^ (r11 & 7)) << 4));
                        uint32_t o[4], code;
                        pool8(a, b, c2, d, o, code, p.pool_code != nullptr);
                        if (PARTIAL && (h0 + 1 >= p.H || w0 + 2 * pw + 1 >= p.W)) continue;      // only complete windows
                        const size_t off = ((static_cast<size_t>(img) * (p.H >> 1) + (h0 >> 1)) * (p.W >> 1) +
                                            (w0 >> 1) + pw) * p.Cout + n0 + c8 * 8;
                        *reinterpret_cast<uint4*>(p.out_pool + off) = make_uint4(o[0], o[1], o[2], o[3]);
                        if (p.pool_code != nullptr) p.pool_code[off >> 3] = static_cast<uint16_t>(code);
                    }
                }
            }
            if (lane == 0) tma_store_wait_all();
            if (p.colsum != nullptr) {
                // chunk sums of all 256 threads -> this CTA's row of the parts buffer, in a FIXED order: thread t parks
                // its 8 sums in the (now idle) staging area, then channel c adds the 32 threads t with (t & 7) == c / 8
                // in increasing t.  No atomics: the bias gradient is bit-reproducible.
                named_bar_sync(3, kEpiThreads);      // ids 1 and 2 belong to the two groups; all stores have been read
                float* park = reinterpret_cast<float*>(staging);
#pragma unroll
                for (int j = 0; j < 8; ++j) park[et * 8 + j] = bacc[j];
                named_bar_sync(3, kEpiThreads);
                if (et < NTILE) {
                    float sacc = 0.f;
                    for (int t = et >> 3; t < kEpiThreads; t += 8) sacc += park[t * 8 + (et & 7)];
                    p.colsum[ABL_PARTS_HEADER + static_cast<size_t>(blockIdx.x) * NTILE + et] = sacc;
                }
                if (blockIdx.x == 0 && et == 0) *reinterpret_cast<int*>(p.colsum) = static_cast<int>(gridDim.x);
            }
        } else {
        int it = 0;
        float wsum[NTILE / 128][2];
#pragma unroll
        for (int q = 0; q < NTILE / 128; ++q) { wsum[q][0] = 0.f; wsum[q][1] = 0.f; }
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int n_tile = tile % p.num_n_tiles;
            int m_tile = tile / p.num_n_tiles;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int img = m_tile / p.tiles_h;
            const int as = it & 1;
#pragma unroll 1
            for (int sub = 0; sub < MSUB; ++sub) {
                long long next_origin = -1;           // where this thread's next ReLU-mask line lives (L2 prefetch)
                if (p.flags & ABL_EPI_RELU_MASK) {
                    if (sub + 1 < MSUB) {
                        next_origin = ((static_cast<long long>(img) * p.H + th * kTileH + (sub + 1) * Cfg::kSubH) * p.W +
                                       tw * Cfg::kTileW) * p.Cout;
                    } else if (tile + static_cast<int>(gridDim.x) < num_tiles) {
                        int m2 = (tile + gridDim.x) / p.num_n_tiles;
                        const int tw2 = m2 % p.tiles_w;
                        m2 /= p.tiles_w;
                        const int th2 = m2 % p.tiles_h, img2 = m2 / p.tiles_h;
                        next_origin = ((static_cast<long long>(img2) * p.H + th2 * kTileH) * p.W + tw2 * Cfg::kTileW) * p.Cout;
                    }
                }
                epilogue_tile_wl<NTILE, PARTIAL>(
                    p, &map_out, staging, tmem_base + (as * MSUB + sub) * NTILE, img, th * kTileH + sub * Cfg::kSubH,
                    tw * Cfg::kTileW, n_tile * NTILE, et, sub == 0 ? &tfull_bar[as] : nullptr, (it >> 1) & 1,
                    sub == MSUB - 1 ? &tempty_bar[as] : nullptr, wsum, next_origin,
                    p.gap_part != nullptr
                        ? p.gap_part + ((static_cast<size_t>(img) * (p.tiles_h * MSUB) + th * MSUB + sub) * p.tiles_w + tw) * p.Cout
                        : nullptr,
                    csum);
            }
        }
        if ((et & 31) == 0) tma_store_wait_all();   // every warp's stores have been written before the CTA exits
        if (p.colsum != nullptr) {
            // one row per CTA (host guarantees a single N tile).  Each warp holds the sums of its row pair: they are
            // parked in the (now idle) staging area and the four row-pair warps of a column are added in fixed order.
            named_bar_sync(1, kEpiThreads);
            float* park = reinterpret_cast<float*>(staging);          // [pass][quarter][128 columns of the pass]
            const int lane = et & 31, quarter = (et >> 5) & 3, hsel = et >> 7;
#pragma unroll
            for (int q = 0; q < NTILE / 128; ++q) {
                park[(q * 4 + quarter) * 128 + hsel * 64 + 2 * lane] = wsum[q][0];
                park[(q * 4 + quarter) * 128 + hsel * 64 + 2 * lane + 1] = wsum[q][1];
            }
            named_bar_sync(1, kEpiThreads);
            for (int i = et; i < NTILE; i += kEpiThreads) {
                const float* c4 = park + (i >> 7) * 512 + (i & 127);
                p.colsum[ABL_PARTS_HEADER + static_cast<size_t>(blockIdx.x) * NTILE + i] = (c4[0] + c4[128]) + (c4[256] + c4[384]);
            }
            if (blockIdx.x == 0 && et == 0) *reinterpret_cast<int*>(p.colsum) = static_cast<int>(gridDim.x);
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

// ---------------------------------------------------------------------------------------------
// "Blocked" main loop for the pooled forward of a Cin = Cout = 64 layer (block1_conv2: bias + ReLU + 2x2 max-pool
// [+ routing codes], no full-resolution output).  An SS-mode MMA reads its A tile (4 KiB) and its B tile from shared
// memory at 128 B/clock, so an M128.N64.K16 MMA cannot retire faster than every 48 clocks although the math needs 32
// (tools/mma_rate.py on B200: 49.7 with independent accumulators, 94 back to back into one accumulator).  Here four
// horizontally adjacent pixels form ONE GEMM row: the NHWC tensor [B,H,W,64] is read as [B,H,W/4,4*64] (the same
// bytes), and the 3x3 convolution becomes, per kh, a banded [4*64] x [4*64] product over pixel slots q (input) and p
// (output) with kw = q - p + 1:
//     q=0 -> p in {0,1}   q=1 -> p in {0,1,2}   q=2 -> p in {1,2,3}   q=3 -> p in {2,3}
// plus the two neighbours across the block boundary (left block's q=3 -> p=0, right block's q=0 -> p=3).  With the
// filter rows of one kh stored as S = [W(kw=2); W(kw=1); W(kw=0)] every band is a CONTIGUOUS row range of S and a
// contiguous column range of the accumulator, so one activation tile feeds N = 128 or 192 output columns per MMA
// (N = 64 only for the two boundary terms): 72 MMA issues worth 5376 tensor clocks per 512 output pixels instead of
// 144 N = 64 issues.  No FLOPs are wasted (the bands hold exactly the 9 taps) and no layout changes.
//   A: per pixel slot q one TMA box {64 ch, q, W/4 blocks, rows+2, 1} of the 5-D view [C, 4, W/4, H, B] (a plane of
//      every 4th pixel, haloed by one image row above/below: kh = descriptor row offset of W/4 rows), plus the same
//      box shifted by -1 / +1 block for the boundary terms (out-of-range blocks and rows are TMA zero fill = SAME
//      padding).  Six 24 KiB boxes per tile travel through a 5-slot ring.
//   B: the whole filter stays resident (3 kh x 24 KiB).
//   D: 128 GEMM rows x 256 columns fp32 in TMEM, double buffered (all 512 columns).
//   Epilogue: one GEMM row holds four adjacent pixels, so the horizontal half of the max-pool is a max between column
//      groups inside a thread's registers; only the h-pooled values cross shared memory for the vertical half.
// Measured on B200, batch 256 (tools/blk64_check.py): 237 us against 288 us for the haloed N = 64 kernel.  The same
// main loop with the generic full-resolution epilogue (forward with out_full, dgrad + ReLU mask) was built and
// measured in round 1/2 and was NOT faster than the haloed kernel there (267 vs 288 us forward, 463 vs 444 us dgrad:
// its epilogue moves twice the bytes through shared memory), so those variants were removed.
struct Blk64bCfg {
    static constexpr int kSlots = 5;
    static constexpr int kMaxARows = 192;                       // (rows + 2) * (W/4): 6 x 32 (W = 128), 10 x 16 (W = 64)
    static constexpr int kASlot = kMaxARows * 128;              // 24 KiB
    static constexpr int kBKh = 3 * 64 * 128;                   // S of one kh: 192 rows x 128 B
    static constexpr int kBBytes = 3 * kBKh;                    // S0 | S1 | S2 (no zero rows: see the MMA issuer)
    static constexpr int kStagingBytes = 2 * kABytes;           // kEpiN = 128
    static constexpr int kBarBytes = 3072;                      // barriers, then csum [256] fp32 OR the pooled epilogue's h-bits (2 KiB)
    static constexpr int kSmemBytes = kSlots * kASlot + kBBytes + kStagingBytes + kBarBytes;
    static_assert(kSmemBytes <= 227 * 1024, "blocked conv: shared memory");
};

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_blk64_pool_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_out, const ConvArgs p) {
    using Cfg = Blk64bCfg;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kSlots * Cfg::kASlot;
    uint8_t* staging = smem_b + Cfg::kBBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
    uint64_t* afull_bar = bars;
    uint64_t* aempty_bar = afull_bar + Cfg::kSlots;
    uint64_t* bfull_bar = aempty_bar + Cfg::kSlots;
    uint64_t* tfull_bar = bfull_bar + 1;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int wb = p.W;                             // blocks per image row (blocked view: W/4)
    const int rows = p.box_h;                       // image rows per tile: 128 / wb
    const int num_tiles = p.num_m_tiles;
    const uint32_t a_bytes = static_cast<uint32_t>((rows + 2) * wb * 128);
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < Cfg::kSlots; ++s) {
            mbar_init(&afull_bar[s], 1);
            mbar_init(&aempty_bar[s], 1);
        }
        mbar_init(bfull_bar, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], kEpiThreads);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    auto s_off = [](int kh) { return kh * Blk64bCfg::kBKh; };      // byte offset of S(kh) in the filter area

    if (warp == 0) {
        // ------------------------------------------------------------ A producer: 6 haloed pixel-slot planes per tile
        if (elect_one()) {
            int slot = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int img = tile / p.tiles_h;
                const int h0 = (tile - img * p.tiles_h) * rows;
#pragma unroll 1
                for (int j = 0; j < 6; ++j) {
                    // consumption order: q = 1 first (initialises columns 0-191), then 0, 2 (initialises 192-255), 3, left, right
                    const int q = j == 0 ? 1 : (j == 1 ? 0 : (j == 2 ? 2 : (j == 5 ? 0 : 3)));
                    const int xb = j == 4 ? -1 : (j == 5 ? 1 : 0);
                    mbar_wait(&aempty_bar[slot], phase ^ 1);
                    mbar_expect_tx(&afull_bar[slot], a_bytes);
                    tma_load_5d(&map_a, &afull_bar[slot], smem_a + slot * Cfg::kASlot, 0, q, xb, h0 - 1, img);
                    if (++slot == Cfg::kSlots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ filter: loaded once, S(kh) = [W(kw=2); W(kw=1); W(kw=0)]
        if (elect_one() && blockIdx.x < num_tiles) {
            mbar_expect_tx(bfull_bar, 9 * 64 * 128);
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw)        // w_kmajor taps are kw-major: tap' = kw*3 + kh
                    tma_load_3d(&map_b, bfull_bar, smem_b + s_off(kh) + (2 - kw) * (64 * 128), 0, 0, kw * 3 + kh);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (elect_one()) {
            constexpr uint32_t id64 = make_idesc_bf16(kTileM, 64, 0, 0), id128 = make_idesc_bf16(kTileM, 128, 0, 0);
            constexpr uint32_t id192 = make_idesc_bf16(kTileM, 192, 0, 0);
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a), 16);
            const uint32_t b_lo0 = desc_lo_sw128(smem_u32(smem_b), 16);
            const uint32_t kh_rows = static_cast<uint32_t>(wb * 128) >> 4;      // descriptor units per image row
            int slot = 0;
            uint32_t phase = 0;
            int it = 0;
            mbar_wait(bfull_bar, 0);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * 256;
#pragma unroll 1
                for (int j = 0; j < 6; ++j) {
                    mbar_wait(&afull_bar[slot], phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + slot * (Cfg::kASlot >> 4);
                    // band of this plane: filter rows [brow, brow + n) of S, accumulator columns [dcol, dcol + n)
                    //   j: 0 (q=1)  1 (q=0)  2 (q=2)  3 (q=3)  4 (left, q=3 -> p=0, kw=0)  5 (right, q=0 -> p=3, kw=2)
                    const int brow = j == 1 ? 64 : (j == 4 ? 128 : 0);
                    const int dcol = j == 2 ? 64 : (j == 3 ? 128 : (j == 5 ? 192 : 0));
                    const uint32_t idesc = (j == 0 || j == 2) ? id192 : (j >= 4 ? id64 : id128);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint32_t b_lo = b_lo0 + ((s_off(kh) + brow * 128) >> 4);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            // First touches: plane q = 1 opens columns 0-191 (accumulate off on its very first MMA);
                            // columns 192-255 are opened by the first k-step of plane q = 2, issued as N = 128
                            // (columns 64-191, accumulating) + N = 64 (columns 192-255, accumulate off).
                            if (j == 2 && (kh | k) == 0) {
                                umma_bf16_lohi(tmem_d + 64, a_lo, dhi, b_lo, dhi, id128, 1u);
                                umma_bf16_lohi(tmem_d + 192, a_lo, dhi, b_lo + ((128 * 128) >> 4), dhi, id64, 0u);
                            } else {
                                umma_bf16_lohi(tmem_d + dcol, a_lo + kh * kh_rows + k * 2, dhi, b_lo + k * 2, dhi, idesc,
                                               (j | kh | k) == 0 ? 0u : 1u);
                            }
                        }
                    }
                    umma_commit(&aempty_bar[slot]);
                    if (++slot == Cfg::kSlots) { slot = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (8 warps): a 256-channel conv on [B, H, W/4]
        const int et = threadIdx.x - 128;
        {
            // Pooled-only forward epilogue (training/inference forward of block1_conv2: bias + ReLU + 2x2 max-pool
            // [+ routing codes], no full-resolution output).  One GEMM row holds four adjacent pixels, so the
            // horizontal half of the pool is a max between column groups inside the thread's registers; only the
            // h-pooled values (half of the generic staging) cross shared memory for the vertical half.
            //   thread = (TMEM lane = GEMM row, hsel): hsel 0 -> pixels p0,p1 (pooled column 2*xb), hsel 1 -> p2,p3
            //   write phase: staging[hsel][row][128 B swizzled] = h-max of 64 channels, hbits[(hsel*128+row)*2 + c] =
            //                bit ch set where the RIGHT pixel is strictly greater (first maximum wins, as TF)
            //   read phase : (row pair, block, hsel, 8-channel chunk): v-max of the two rows, routing code =
            //                top >= bottom ? hbit_top : 2 + hbit_bottom  (index of the first maximum in scan order)
            const int lane_row = ((et >> 5) & 3) * 32 + (et & 31);
            const int hsel = et >> 7;
            uint32_t* hbits = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 512);
            auto bf2_max = [](uint32_t a, uint32_t b) {
                const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a),
                                                 *reinterpret_cast<const __nv_bfloat162*>(&b));
                return *reinterpret_cast<const uint32_t*>(&m);
            };
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int img = tile / p.tiles_h;
                const int h0 = (tile - img * p.tiles_h) * rows;
                const int as = it & 1;
                mbar_wait(&tfull_bar[as], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * 256 + hsel * 128 + (static_cast<uint32_t>(lane_row & ~31) << 16);
                uint8_t* hrow = staging + hsel * kABytes + lane_row * 128;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(tacc + c * 32, ra);            // left pixel of the pair, channels 32c .. 32c+31
                    tmem_ld_32x32b_x32(tacc + 64 + c * 32, rb);       // right pixel
                    tmem_ld_wait();
                    const float4* bs = reinterpret_cast<const float4*>(p.bias + c * 32);
                    uint32_t hm[16], bits = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = __ldg(bs + j);
                        const uint32_t a0 = pack_bf16x2_relu(__uint_as_float(ra[4 * j]) + b4.x,
                                                        __uint_as_float(ra[4 * j + 1]) + b4.y);
                        const uint32_t a1 = pack_bf16x2_relu(__uint_as_float(ra[4 * j + 2]) + b4.z,
                                                        __uint_as_float(ra[4 * j + 3]) + b4.w);
                        const uint32_t c0 = pack_bf16x2_relu(__uint_as_float(rb[4 * j]) + b4.x,
                                                        __uint_as_float(rb[4 * j + 1]) + b4.y);
                        const uint32_t c1 = pack_bf16x2_relu(__uint_as_float(rb[4 * j + 2]) + b4.z,
                                                        __uint_as_float(rb[4 * j + 3]) + b4.w);
                        hm[2 * j] = bf2_max(a0, c0);
                        hm[2 * j + 1] = bf2_max(a1, c1);
                        // post-ReLU values are >= +0: bit equality is value equality; the right pixel wins only if the
                        // left one is not the maximum
                        const uint32_t n0 = ~__vcmpeq2(a0, hm[2 * j]), n1 = ~__vcmpeq2(a1, hm[2 * j + 1]);
                        bits |= ((n0 & 1u) | ((n0 >> 15) & 2u) | ((n1 & 1u) << 2) | ((n1 >> 13) & 8u)) << (4 * j);
                    }
                    if (c == 1) {               // all TMEM reads of this thread are done
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[as]);
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        *reinterpret_cast<uint4*>(hrow + (((c * 4 + cc) ^ (lane_row & 7)) << 4)) =
                            make_uint4(hm[4 * cc], hm[4 * cc + 1], hm[4 * cc + 2], hm[4 * cc + 3]);
                    hbits[(hsel * 128 + lane_row) * 2 + c] = bits;
                }
                // The vertical half of the pool pairs GEMM rows m and m + wb: only the 2*wb threads that hold one pair of
                // image rows of one pixel-pair half have to meet (two warps at W = 128, one at W = 64) -- eight small
                // named barriers instead of one over all 256 epilogue threads (ncu: 20 % barrier stalls).
                const int gthreads = 2 * wb;                        // 64 or 32
                const int gq = lane_row / gthreads, tg = lane_row - gq * gthreads;
                const uint32_t bar_id = 1 + hsel * 4 + gq;
                named_bar_sync(bar_id, gthreads);                   // h-pooled rows and bits of this row pair are in shared memory
                const uint8_t* hb = staging + hsel * kABytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {                       // (block, 8-channel chunk) of row pair gq: wb * 8 items
                    const int q = tg + gthreads * k;
                    const int c8 = q & 7, xb = q >> 3, rp = gq, pp = hsel;
                    const int m0 = (2 * rp) * wb + xb, m1 = m0 + wb;
                    const uint4 t4 = *reinterpret_cast<const uint4*>(hb + m0 * 128 + ((c8 ^ (m0 & 7)) << 4));
                    const uint4 b4 = *reinterpret_cast<const uint4*>(hb + m1 * 128 + ((c8 ^ (m1 & 7)) << 4));
                    const uint32_t tb = (hbits[(pp * 128 + m0) * 2 + (c8 >> 2)] >> ((c8 & 3) * 8)) & 0xFFu;
                    const uint32_t bb = (hbits[(pp * 128 + m1) * 2 + (c8 >> 2)] >> ((c8 & 3) * 8)) & 0xFFu;
                    const uint32_t tv[4] = {t4.x, t4.y, t4.z, t4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
                    uint32_t o[4], code = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o[j] = bf2_max(tv[j], bv[j]);
                        const uint32_t et2 = __vcmpeq2(tv[j], o[j]);          // 0xFFFF where the top row holds the maximum
                        const uint32_t k0 = 2 * j, k1 = 2 * j + 1;
                        const uint32_t c_lo = (et2 & 1u) ? ((tb >> k0) & 1u) : (2u + ((bb >> k0) & 1u));
                        const uint32_t c_hi = (et2 & 0x10000u) ? ((tb >> k1) & 1u) : (2u + ((bb >> k1) & 1u));
                        code |= (c_lo << (2 * k0)) | (c_hi << (2 * k1));
                    }
                    const size_t off = ((static_cast<size_t>(img) * (p.H >> 1) + (h0 >> 1) + rp) * (2 * wb) + 2 * xb + pp) * 64 + c8 * 8;
                    *reinterpret_cast<uint4*>(p.out_pool + off) = make_uint4(o[0], o[1], o[2], o[3]);
                    if (p.pool_code != nullptr) p.pool_code[off >> 3] = static_cast<uint16_t>(code);
                }
                named_bar_sync(bar_id, gthreads);                   // the exchange rows may be overwritten by the next tile
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
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

template <int NTILE, int TAPS>
static int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const ConvArgs& a,
                       cudaStream_t st) {
    using Cfg = ConvCfg<NTILE, TAPS>;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(conv3x3_tc_kernel<NTILE, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes) != cudaSuccess)
            return check_launch();
        attr_done[dev & 63] = true;
    }
    const int tiles = a.num_m_tiles * a.num_n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    conv3x3_tc_kernel<NTILE, TAPS><<<grid, kThreads, Cfg::kSmemBytes, st>>>(ma, mb, mo, a);
    return check_launch();
}

template <int NTILE, int MSUB, bool BRES, bool PARTIAL>
static int launch_halo_inst(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const ConvArgs& a,
                            cudaStream_t st) {
    using Cfg = HaloCfg<NTILE, MSUB, BRES>;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(conv3x3_halo_kernel<NTILE, MSUB, BRES, PARTIAL>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
            return check_launch();
        attr_done[dev & 63] = true;
    }
    const int tiles = a.num_m_tiles * a.num_n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    conv3x3_halo_kernel<NTILE, MSUB, BRES, PARTIAL><<<grid, kThreads, Cfg::kSmemBytes, st>>>(ma, mb, mo, a);
    return check_launch();
}

template <int NTILE, int MSUB, bool BRES = false>
static int launch_halo(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const ConvArgs& a,
                       cudaStream_t st) {
    return a.partial ? launch_halo_inst<NTILE, MSUB, BRES, true>(ma, mb, mo, a, st)
                     : launch_halo_inst<NTILE, MSUB, BRES, false>(ma, mb, mo, a, st);
}

}  // namespace abl

using namespace abl;

// See include/abl_b200.h for the contract.
extern "C" int abl_bias_grad(const void* dz, float* db_parts, long long db_parts_bytes, long long rows, int C, void* stream);

static int conv_igemm_impl(const void* x, const void* w_kmajor, const float* bias, const void* mask_src,
                           void* out_full, void* out_pool, unsigned short* pool_code, float* colsum, long long colsum_bytes,
                           void* out_unpool, const unsigned short* unpool_code, float* gap_part,
                           unsigned int* relu_bits, int B, int H, int W,
                           int Cin, int Cout, int flags, void* stream) {
    if (colsum && !out_full && !out_unpool) { set_error("abl_conv3x3_igemm: colsum needs out_full"); return ABL_ERR_NULL; }
    if (pool_code && !out_pool) { set_error("abl_conv3x3_igemm: pool_code needs out_pool"); return ABL_ERR_NULL; }
    if (out_unpool && (!unpool_code || out_full || out_pool)) {
        set_error("abl_conv3x3_dgrad_unpool: needs routing codes and no other output"); return ABL_ERR_NULL;
    }
    if (!x || !w_kmajor) { set_error("abl_conv3x3_igemm: null input"); return ABL_ERR_NULL; }
    if (!out_full && !out_pool && !out_unpool && !gap_part) { set_error("abl_conv3x3_igemm: no output requested"); return ABL_ERR_NULL; }
    if ((flags & ABL_EPI_BIAS_RELU) && !bias) { set_error("abl_conv3x3_igemm: bias missing"); return ABL_ERR_NULL; }
    if ((flags & ABL_EPI_RELU_MASK) && (!mask_src || (!out_full && !out_unpool))) {
        set_error("abl_conv3x3_igemm: relu-mask epilogue needs mask_src and out_full");
        return ABL_ERR_NULL;
    }
    if (B <= 0 || Cin % 64 != 0 || Cout % 64 != 0 || Cin > 512 || Cout > 512) {
        set_error("abl_conv3x3_igemm: need B>0, Cin and Cout multiples of 64 and <= 512");
        return ABL_ERR_SHAPE;
    }
    if (out_pool && (H < 2 || W < 2)) { set_error("abl_conv3x3_igemm: pooled output needs H, W >= 2"); return ABL_ERR_SHAPE; }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_kmajor) |
         reinterpret_cast<uintptr_t>(out_full) | reinterpret_cast<uintptr_t>(out_pool) |
         reinterpret_cast<uintptr_t>(mask_src)) & 15) {
        set_error("abl_conv3x3_igemm: pointers must be 16-byte aligned");
        return ABL_ERR_ALIGN;
    }
    ConvArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    // v1 / blocked kernels need an exact tiling into 128-pixel rectangles; the haloed kernels below take any frame
    const bool boxed = pick_box(H, W, &a.box_h, &a.box_w) == ABL_OK && !(out_pool && ((a.box_h | H | W) & 1));
    {
        // Blocked main loop (four pixels per GEMM row) for the pooled-only forward of a Cin = Cout = 64 layer
        // (block1_conv2 in training and inference).  ABL_BLOCKED64=0 falls back to the haloed kernel (A/B runs, tests).
        const char* blk = getenv("ABL_BLOCKED64");
        const bool shape_ok = Cin == 64 && Cout == 64 && (W == 128 || W == 64) && H % (512 / W) == 0 && (512 / W) % 2 == 0;
        if (shape_ok && out_pool && !out_full && !colsum && !out_unpool && !gap_part && flags == ABL_EPI_BIAS_RELU &&
            !(blk && blk[0] == '0')) {
            const int wb = W / 4, rows = 512 / W;
            a.W = wb; a.Cout = 256;
            a.box_w = wb; a.box_h = rows; a.tiles_h = H / rows; a.tiles_w = 1;
            a.num_m_tiles = B * a.tiles_h; a.num_n_tiles = 1; a.kchunks = 1;
            a.flags = flags;
            a.bias = bias;
            a.mask_src = nullptr;
            a.out_full = nullptr;
            a.out_pool = static_cast<__nv_bfloat16*>(out_pool);
            a.pool_code = pool_code;
            a.colsum = nullptr;
            a.out_unpool = nullptr; a.unpool_code = nullptr; a.gap_part = nullptr; a.partial = 0; a.relu_bits = nullptr;
            CUtensorMap ma, mb;
            {
                const uint64_t dims[5] = {64, 4, (uint64_t)wb, (uint64_t)H, (uint64_t)B};
                const uint64_t strides[4] = {64 * 2, 4 * 64 * 2, (uint64_t)W * 64 * 2, (uint64_t)H * W * 64 * 2};
                const uint32_t box[5] = {64, 1, (uint32_t)wb, (uint32_t)(rows + 2), 1};
                int r = encode_tiled(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
                if (r != ABL_OK) return r;
            }
            {
                const uint64_t dims[3] = {64, 64, 9};
                const uint64_t strides[2] = {64 * 2, 64 * 64 * 2};
                const uint32_t box[3] = {64, 64, 1};
                int r = encode_tiled(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w_kmajor, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
                if (r != ABL_OK) return r;
            }
            static bool attr_done[64] = {false};
            int dev = 0;
            cudaGetDevice(&dev);
            if (!attr_done[dev & 63]) {
                if (cudaFuncSetAttribute(conv3x3_blk64_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Blk64bCfg::kSmemBytes) != cudaSuccess)
                    return check_launch();
                attr_done[dev & 63] = true;
            }
            const int grid = a.num_m_tiles < sm_count() ? a.num_m_tiles : sm_count();
            conv3x3_blk64_pool_kernel<<<grid, kThreads, Blk64bCfg::kSmemBytes, static_cast<cudaStream_t>(stream)>>>(ma, mb, ma, a);
            return check_launch();
        }
    }
    const int ntile = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
    // v2 "halo" main loop: tiles of (8*MSUB) rows x 16 pixels (MSUB = 1 for N = 256, else 2).  ABL_IGEMM_V1=1
    // (environment, read per call) forces the v1 kernel for A/B measurements and its own parity tests.
    const int msub = ntile == 256 ? 1 : 2;
    const char* force_v1 = getenv("ABL_IGEMM_V1");
    // N = 64 MMAs are shared-memory-read-bound in SS mode (ncu: tensor pipe 50 % active = one M128 N64 K16 MMA per
    // 64 clocks); with Cin = 128 the v1 4x32 tile measured 7 % faster than the haloed tile (134 vs 144 us, B = 256)
    // -- without a ReLU-mask epilogue only: with one, the haloed kernel's next-tile mask prefetch wins (150 vs 157 us)
    const bool v1_better = ntile == 64 && Cin != 64 && !(flags & ABL_EPI_RELU_MASK) && (W % 32 == 0) && (H % 4 == 0) &&
                           !out_unpool && !gap_part && !(force_v1 && force_v1[0] == '0');
    // The haloed kernels take ANY frame size: tiles are counted rounding up, TMA zero-fills loads and clips stores beyond
    // the frame, and the epilogue guards everything else (ConvArgs::partial).  Pooled / un-pooled outputs need even H, W.
    // Frames v1 tiles exactly keep using it when the haloed tile does not fit (same kernels as before).
    const bool exact = W % 16 == 0 && H % (8 * msub) == 0;
    const bool want_halo = exact ? ((!(force_v1 && force_v1[0] == '1') || out_unpool || gap_part) && !v1_better)
                                 : (out_unpool || gap_part || !boxed || (force_v1 && force_v1[0] == '0'));
    if (want_halo) {
        a.box_w = 16; a.box_h = 8 * msub;
        a.tiles_h = (H + a.box_h - 1) / a.box_h; a.tiles_w = (W + 15) / 16;
        a.partial = exact ? 0 : 1;
        a.num_m_tiles = B * a.tiles_h * a.tiles_w;
        a.num_n_tiles = Cout / ntile;
        a.kchunks = Cin / kBlockK;
        a.flags = flags;
        {
            // Weight-stationary MMA pairs (tcgen05.mma.ws, collector b0 fill -> lastuse): the two sub-tiles of a tile
            // share one shared-memory read of each filter slice.  ABL_N64_WS=0 disables.
            // Same-box A/B on B200 (batch 256): N = 64 +4 %; N = 128: block2_conv1 fwd 122 -> 114 us, block2_conv2 dgrad
            // 226 -> 210 us, but block2_conv2 fwd 200 -> 206 us, so N = 128 uses it for Cin = 64 and masked Cin = 128.
            const char* ws = getenv("ABL_N64_WS");
            // (round 2, second A/B with ABL_N128_WS_ALL=1: the N = 128 dgrad + un-pool with K = 256 gains too, 130 -> 127 us;
            // the un-masked Cin = 128 forward still loses, 201 -> 206 us)
            const char* ws_all = getenv("ABL_N128_WS_ALL");
            const bool ws_shape = ntile == 64 || (ntile == 128 && (Cin == 64 || (Cin == 128 && (flags & ABL_EPI_RELU_MASK)) ||
                                                                   out_unpool != nullptr || (ws_all && ws_all[0] == '1')));
            if (ws_shape && !(ws && ws[0] == '0')) a.flags |= ABL_MAIN_WS_PAIRS;
        }
        a.bias = bias;
        a.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
        a.out_full = static_cast<__nv_bfloat16*>(out_full);
        a.out_pool = static_cast<__nv_bfloat16*>(out_pool);
        a.pool_code = pool_code;
        // fused BiasAddGrad: the 8-warp epilogue with a single N tile; otherwise a separate pass over out_full
        const bool fuse_colsum = colsum != nullptr && a.num_n_tiles == 1 && (ntile != 64 || out_unpool != nullptr);
        a.colsum = fuse_colsum ? colsum : nullptr;
        a.out_unpool = static_cast<__nv_bfloat16*>(out_unpool);
        a.unpool_code = unpool_code;
        a.gap_part = gap_part;
        a.relu_bits = relu_bits;
        if (relu_bits && ntile != 256) { set_error("abl_conv3x3_igemm_gap: relu_bits needs Cout a multiple of 256"); return ABL_ERR_SHAPE; }
        if (gap_part && ntile == 64) { set_error("abl_conv3x3_igemm_gap: needs Cout a multiple of 128"); return ABL_ERR_SHAPE; }
        if (out_unpool && (ntile == 256 || !fuse_colsum && colsum != nullptr)) {
            set_error("abl_conv3x3_dgrad_unpool: supported for Cout <= 128 (single N tile)"); return ABL_ERR_SHAPE;
        }
        CUtensorMap ma, mb;
        {
            const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
            const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
            const uint32_t box[4] = {(uint32_t)kBlockK, 16u, (uint32_t)(a.box_h + 2), 1};
            int r = encode_tiled(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (r != ABL_OK) return r;
        }
        {
            const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
            const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
            const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)ntile, ntile == 64 ? 3u : 1u};
            int r = encode_tiled(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w_kmajor, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (r != ABL_OK) return r;
        }
        CUtensorMap mo = ma;   // valid placeholder when only the pooled output is requested (never dereferenced)
        if (out_full) {
            const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
            const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
            // every epilogue warp stores its own 2 image rows x 16 pixels x 64 channels (see the kernels)
            const uint32_t box[4] = {64, 16u, 2u, 1};
            int r = encode_tiled(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out_full, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (r != ABL_OK) return r;
        }
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        int r;
        if (fuse_colsum) {
            const int tiles = a.num_m_tiles * a.num_n_tiles;
            r = parts_check(colsum, colsum_bytes, tiles < sm_count() ? tiles : sm_count(), Cout, "abl_conv3x3_igemm");
            if (r != ABL_OK) return r;
        }
        switch (ntile) {
            case 256: r = launch_halo<256, 1>(ma, mb, mo, a, st); break;
            case 128: r = launch_halo<128, 2>(ma, mb, mo, a, st); break;
            default: r = (Cin == 64 && Cout == 64) ? launch_halo<64, 2, true>(ma, mb, mo, a, st)
                                                    : launch_halo<64, 2, false>(ma, mb, mo, a, st);
        }
        if (r == ABL_OK && colsum != nullptr && !fuse_colsum)
            r = abl_bias_grad(out_full, colsum, colsum_bytes, static_cast<long long>(B) * H * W, Cout, stream);
        return r;
    }
    if (!boxed) { set_error("abl_conv3x3_igemm: internal: the exact-tile kernel was chosen for a frame it cannot tile"); return ABL_ERR_SHAPE; }
    a.partial = 0;
    a.colsum = nullptr;
    a.out_unpool = nullptr;
    a.unpool_code = nullptr;
    a.gap_part = nullptr;
    a.relu_bits = nullptr;
    // 3-taps-per-stage main loop for the narrow tile: needs the 4 x 32 pixel rectangle
    const bool taps3 = ntile == 64 && (W % 32 == 0) && (H % 4 == 0);
    if (taps3) { a.box_w = 32; a.box_h = 4; }
    a.tiles_h = H / a.box_h; a.tiles_w = W / a.box_w;
    a.num_m_tiles = B * a.tiles_h * a.tiles_w;
    a.num_n_tiles = Cout / ntile;
    a.kchunks = Cin / kBlockK;
    a.flags = flags;
    a.bias = bias;
    a.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
    a.out_full = static_cast<__nv_bfloat16*>(out_full);
    a.out_pool = static_cast<__nv_bfloat16*>(out_pool);
        a.pool_code = pool_code;

    CUtensorMap ma, mb;
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)a.box_w, (uint32_t)(a.box_h + (taps3 ? 2 : 0)), 1};
        int r = encode_tiled(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != ABL_OK) return r;
    }
    {
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
        const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
        const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)ntile, taps3 ? 3u : 1u};
        int r = encode_tiled(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w_kmajor, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != ABL_OK) return r;
    }
    CUtensorMap mo = ma;   // valid placeholder when only the pooled output is requested (never dereferenced)
    if (out_full) {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
        const uint32_t box[4] = {64, (uint32_t)a.box_w, (uint32_t)a.box_h, 1};
        int r = encode_tiled(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out_full, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != ABL_OK) return r;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int r;
    switch (ntile) {
        case 256: r = launch_conv<256, 1>(ma, mb, mo, a, st); break;
        case 128: r = launch_conv<128, 1>(ma, mb, mo, a, st); break;
        default: r = taps3 ? launch_conv<64, 3>(ma, mb, mo, a, st) : launch_conv<64, 1>(ma, mb, mo, a, st);
    }
    if (r == ABL_OK && colsum != nullptr)
        r = abl_bias_grad(out_full, colsum, colsum_bytes, static_cast<long long>(B) * H * W, Cout, stream);
    return r;
}

extern "C" int abl_conv3x3_igemm(const void* x, const void* w_kmajor, const float* bias, const void* mask_src,
                                 void* out_full, void* out_pool, unsigned short* pool_code, float* colsum_parts,
                                 long long colsum_parts_bytes, int B, int H, int W, int Cin, int Cout, int flags,
                                 void* stream) {
    return conv_igemm_impl(x, w_kmajor, bias, mask_src, out_full, out_pool, pool_code, colsum_parts, colsum_parts_bytes,
                           nullptr, nullptr, nullptr, nullptr, B, H, W, Cin, Cout, flags, stream);
}

// See include/abl_b200.h: forward conv + bias + ReLU with the GlobalAveragePooling partial sums fused in.
extern "C" int abl_conv3x3_igemm_gap(const void* x, const void* w_kmajor, const float* bias, void* out_full,
                                     float* gap_part, unsigned int* relu_bits, int B, int H, int W, int Cin, int Cout,
                                     void* stream) {
    if (!gap_part) { set_error("abl_conv3x3_igemm_gap: gap_part is null"); return ABL_ERR_NULL; }
    if (reinterpret_cast<uintptr_t>(relu_bits) & 3) { set_error("abl_conv3x3_igemm_gap: relu_bits must be 4-byte aligned"); return ABL_ERR_ALIGN; }
    return conv_igemm_impl(x, w_kmajor, bias, nullptr, out_full, nullptr, nullptr, nullptr, 0, nullptr, nullptr, gap_part,
                           relu_bits, B, H, W, Cin, Cout, ABL_EPI_BIAS_RELU, stream);
}

// See include/abl_b200.h: dgrad through a max-pool in one kernel.
extern "C" int abl_conv3x3_dgrad_unpool(const void* dz, const void* w_dgrad, const void* pooled,
                                        const unsigned short* pool_code, void* dz_out, float* db_parts,
                                        long long db_parts_bytes, int B, int H, int W, int Cin, int Cout, void* stream) {
    if (!pooled || !pool_code || !dz_out) { set_error("abl_conv3x3_dgrad_unpool: null pointer"); return ABL_ERR_NULL; }
    return conv_igemm_impl(dz, w_dgrad, nullptr, pooled, nullptr, nullptr, nullptr, db_parts, db_parts_bytes, dz_out,
                           pool_code, nullptr, nullptr, B, H, W, Cin, Cout, ABL_EPI_RELU_MASK, stream);
}
