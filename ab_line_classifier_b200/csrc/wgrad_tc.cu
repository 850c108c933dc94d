// Filter gradient of a 3x3/s1/SAME convolution (TF Conv2DBackpropFilter run by model.fit's tape,
// reference call site src/models/models.py:427,432) as a tcgen05 GEMM whose REDUCTION dimension is
// the pixel index:
//     dW[kh][kw][ci][co] = sum_{n,h,w} X[n,h+kh-1,w+kw-1,ci] * dZ[n,h,w,co]
// Both operands are read straight from the NHWC tensors by TMA (SWIZZLE_128B boxes of 64 channels):
// in shared memory a box is "pixel rows x 128 B", i.e. the MN-major canonical layout with K = pixel
// rows, so no transpose pass exists anywhere.
//
// Work decomposition.  A CTA owns one UNIT and one slice of the pixel tiles (4 rows x 16 pixels each):
//   unit = (kw, 128 input channels, NT output channels)            [Cin >= 128]
//        = ({kw0,kw1} or {kw2,-}, 64 input channels, NT)           [Cin == 64: two kw share UMMA M = 128]
// Per pipeline stage ONE 5-D TMA box brings the X tile with a one-row halo above and below
// (6 x 16 pixels x 128 channels, 24 KiB) and ONE box brings the dZ tile (4 x 16 pixels x NT, <= 16 KiB).
// The three kh taps are three TMEM accumulators fed from the SAME X tile through descriptor start
// offsets of kh*16 rows (2 KiB, a multiple of the 1 KiB swizzle atom): 12 MMAs per two TMA instructions.
// Partial sums of each pixel slice go to a workspace with plain vector stores; a second kernel reduces
// the slices in a fixed order (deterministic) into the fp32 HWIO gradient.
#include <stdlib.h>

#include <mutex>

#include "host_util.h"
#include "ptx.cuh"
#include "parts.cuh"

namespace abl {

constexpr int kWgBw = 16, kWgBh = 4;                       // pixel tile (GEMM K = 64 per stage)
constexpr int kWgARows = (kWgBh + 2) * kWgBw;              // 96 haloed pixel rows
constexpr int kWgABox = kWgARows * 128;                    // 12 KiB per 64-channel block
constexpr int kWgBBox = kWgBh * kWgBw * 128;               // 8 KiB per 64-channel block
constexpr int kWgThreads = 256;
constexpr int kWgMaxStages = 8;

struct WgradArgs {
    int B, H, W, Cin, Cout;
    int tiles_h, tiles_w, num_ptiles;
    int mode;            // 0: Cin >= 128 (M = 2 channel blocks of one kw); 1: Cin == 64 (M = 2 kw of one block)
    int ci_blocks;       // mode 0: Cin/128
    int co_blocks;       // Cout/NT
    int units, splits, stages, stage_bytes;
    float* partial;      // [splits][9][Cin][Cout] fp32
    float* db_rows;      // wgrad64 only: bias-gradient partial rows [splits][Cout] (ones row of the GEMM) or null
    // Slice reduction INSIDE the GEMM kernel (tickets != null; wgrad_unit_finish).  tickets: 2 * units zeroed words
    // (arrivals, departures) that the kernel leaves zeroed again.
    unsigned int* tickets;
    float* dw;           // [9][Cin][Cout] fp32 result
    int accumulate;      // dw += instead of dw =
    float* db;           // bias gradient [Cout] or null
    const float* db_src; // a parts buffer (db_nrows < 0: row count in its header) finished by the first db_blocks CTAs
    int db_nrows, db_blocks;
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Fixed-order sum of the pixel slices, done by the GEMM kernel itself (round 2 ran a separate reduction kernel after every
// wgrad GEMM: 8-15 us each, six per step).  Called by ALL 256 threads of CTA (unit, split) after the __syncthreads that
// follows the epilogue's stores of the CTA's slice sums into p.partial:
//   1. arrive on the unit's ticket and wait until all `splits` CTAs of the unit have arrived -- they are all resident
//      (grid <= #SM, one CTA per SM), so the wait cannot deadlock; a bounded spin traps instead of hanging;
//   2. this CTA owns rows [split * R / splits, ...) of the unit's region of dw.  Every output float4 is the sum of its
//      `splits` slice values; G = 256 / outputs threads share one output, thread g of them adds a contiguous range of
//      slices in slice order (16 L2 loads in flight: __ldcg, other SMs wrote the data) and the G sums are added in g
//      order -- the same order on every launch, so the result is bit-reproducible;
//   3. the last CTA to leave zeroes the unit's two tickets for the next launch.
// row_off(r): offset (floats) of row r of the unit's region inside a [9][Cin][Cout] image; ncol floats per row.
template <typename RowOff>
__device__ __forceinline__ void wgrad_unit_finish(const WgradArgs& p, int unit, int split, int nrows, int ncol,
                                                  RowOff row_off) {
    const int t = threadIdx.x;
    __threadfence();
    __syncthreads();
    unsigned int* arrive = p.tickets + unit;
    unsigned int* leave = p.tickets + p.units + unit;
    if (t == 0) {
        atomicAdd(arrive, 1u);
        if (ld_acquire_u32(arrive) < static_cast<unsigned>(p.splits)) {
            const uint64_t t0 = global_timer_ns();
            uint32_t spins = 0;
            while (ld_acquire_u32(arrive) < static_cast<unsigned>(p.splits))
                if ((++spins & 0xFFF) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();
        }
    }
    __syncthreads();
    const int c4n = ncol >> 2;                       // float4 per row (16 or 32)
    const int per = (nrows + p.splits - 1) / p.splits;
    const int r0 = split * per, r1 = min(nrows, r0 + per);
    const int n_out = max(r1 - r0, 0) * c4n;         // float4 outputs of this CTA
    int G = 1;                                       // threads per output: a power of two, lanes of one warp
    while (G < 32 && n_out * (G * 2) <= 256) G *= 2;
    const int g = t & (G - 1);
    const int kper = (p.splits + G - 1) / G;         // slices per thread
    const int k0 = g * kper, k1 = min(p.splits, k0 + kper);
    const size_t sstride4 = static_cast<size_t>(9) * p.Cin * p.Cout >> 2;
    const int passes = (n_out * G + 255) / 256;
    for (int pass = 0; pass < passes; ++pass) {      // (uniform trip count: the shuffles below need whole warps)
        const int o = (pass * 256 + t) / G;
        const bool live = o < n_out;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        size_t off = 0;
        if (live) {
            off = row_off(r0 + o / c4n) + static_cast<size_t>(o % c4n) * 4;
            const float4* src = reinterpret_cast<const float4*>(p.partial + off);
            for (int k = k0; k < k1; k += 16) {
                float4 v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = k + j < k1 ? __ldcg(src + static_cast<size_t>(k + j) * sstride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 16; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
            }
        }
        float4 tot = acc;
        for (int q = 1; q < G; ++q) {                // thread g == 0 collects the other groups' sums in g order
            const int srcl = (t & 31 & ~(G - 1)) + q;
            tot.x += __shfl_sync(0xffffffffu, acc.x, srcl);
            tot.y += __shfl_sync(0xffffffffu, acc.y, srcl);
            tot.z += __shfl_sync(0xffffffffu, acc.z, srcl);
            tot.w += __shfl_sync(0xffffffffu, acc.w, srcl);
        }
        if (live && g == 0) {
            float4* dst = reinterpret_cast<float4*>(p.dw + off);
            if (p.accumulate) { const float4 d = *dst; tot.x += d.x; tot.y += d.y; tot.z += d.z; tot.w += d.w; }
            *dst = tot;
        }
    }
    // wgrad64: the ones row of every slice -> db of this unit's 64 output channels (slice order)
    if (p.db_rows != nullptr && p.db != nullptr && split == 0 && t < 16) {
        const float4* src = reinterpret_cast<const float4*>(p.db_rows + unit * 64) + t;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < p.splits; ++k) {
            const float4 v = __ldcg(src + (static_cast<size_t>(k) * p.Cout >> 2));
            s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
        }
        reinterpret_cast<float4*>(p.db + unit * 64)[t] = s4;
    }
    if (t == 0 && atomicAdd(leave, 1u) == static_cast<unsigned>(p.splits) - 1u) {
        *arrive = 0u;           // every CTA of the unit has passed its wait: safe to re-arm for the next launch
        *leave = 0u;
    }
}

// BiasAddGrad finish from a parts buffer inside the GEMM kernel: the first db_blocks CTAs each add the rows of 16 channels
// (all 256 threads, after the CTA's own work; `red` = 4 KiB of the drained pipeline's shared memory).
__device__ __forceinline__ void wgrad_bias_finish(const WgradArgs& p, uint8_t* smem) {
    if (p.db_src == nullptr || static_cast<int>(blockIdx.x) >= p.db_blocks) return;
    const float* rows = p.db_nrows < 0 ? p.db_src + ABL_PARTS_HEADER : p.db_src;
    const int nrows = p.db_nrows < 0 ? parts_rows(p.db_src) : p.db_nrows;
    for (int blk = blockIdx.x; blk < p.db_blocks; blk += gridDim.x) {       // (a tiny problem has fewer CTAs than blocks)
        parts_reduce_block(rows, nrows, p.Cout, p.db, 0, blk, reinterpret_cast<float4(*)[4]>(smem));
        __syncthreads();
    }
}

template <int NT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz,
                const WgradArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kABytes = 2 * kWgABox;                   // M = 128 = two 64-wide blocks
    constexpr int kBBytes = (NT / 64) * kWgBBox;
    uint8_t* zero_blk = smem + p.stages * p.stage_bytes;   // mode 1, unit {kw2,-}: the missing M block
    uint64_t* bars = reinterpret_cast<uint64_t*>(zero_blk + kWgABox);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgMaxStages;
    uint64_t* done_bar = bars + 2 * kWgMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int unit = blockIdx.x % p.units;
    const int split = blockIdx.x / p.units;
    int kw0, kw1, ci0, co0;          // kw of M block 0 / 1 (kw1 < 0: zero block), first input / output channel
    if (p.mode == 0) {
        kw0 = kw1 = unit % 3;
        ci0 = ((unit / 3) % p.ci_blocks) * 128;
        co0 = (unit / (3 * p.ci_blocks)) * NT;
    } else {
        const int pair = unit & 1;
        kw0 = pair * 2;
        kw1 = pair == 0 ? 1 : -1;
        ci0 = 0;
        co0 = (unit >> 1) * NT;
    }
    const int per = (p.num_ptiles + p.splits - 1) / p.splits;
    const int pt0 = split * per;
    const int pt1 = min(p.num_ptiles, pt0 + per);

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_dz);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    if (p.mode == 1)
        for (int i = threadIdx.x; i < kWgABox / 16; i += blockDim.x) reinterpret_cast<uint4*>(zero_blk)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (p.mode == 0 || kw1 >= 0 ? kABytes : kWgABox) + kBBytes;
            for (int pt = pt0; pt < pt1; ++pt) {
                int t = pt;
                const int tw = t % p.tiles_w;
                t /= p.tiles_w;
                const int th = t % p.tiles_h;
                const int img = t / p.tiles_h;
                const int h0 = th * kWgBh, w0 = tw * kWgBw;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], tx);
                uint8_t* sa = smem + stage * p.stage_bytes;
                tma_load_5d(&map_dz, &full_bar[stage], sa + kABytes, 0, w0, h0, co0 / 64, img);
                if (p.mode == 0) {
                    tma_load_5d(&map_x, &full_bar[stage], sa, 0, w0 + kw0 - 1, h0 - 1, ci0 / 64, img);
                } else {
                    tma_load_5d(&map_x, &full_bar[stage], sa, 0, w0 + kw0 - 1, h0 - 1, 0, img);
                    if (kw1 >= 0) tma_load_5d(&map_x, &full_bar[stage], sa + kWgABox, 0, w0 + kw1 - 1, h0 - 1, 0, img);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);   // both operands MN-major
            // MN-major SWIZZLE_128B: 8-row (pixel) groups 1024 B apart (SBO), 64-channel blocks LBO apart
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            const uint32_t smem0 = smem_u32(smem);
            const uint32_t zero_addr = smem_u32(zero_blk);
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem0 + stage * p.stage_bytes;
                const uint32_t a_lbo = (p.mode == 1 && kw1 < 0) ? (zero_addr - a_addr) : kWgABox;
                const uint32_t a_lo = desc_lo_sw128(a_addr, a_lbo);
                const uint32_t b_lo = desc_lo_sw128(a_addr + kABytes, kWgBBox);
                const uint32_t acc_flag = pt > pt0 ? 1u : 0u;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int k = 0; k < (kWgBh * kWgBw) / 16; ++k) {
                        // kh shift = 16 pixel rows = 2048 B; K advance of 16 pixel rows = 2048 B
                        umma_bf16_lohi(tmem_base + kh * NT, a_lo + (kh + k) * (2048 >> 4), dhi, b_lo + k * (2048 >> 4),
                                       dhi, idesc, k > 0 ? 1u : acc_flag);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);
        }
    } else if (warp >= 4) {
        const int row = threadIdx.x - 128;        // TMEM lane = D row
        const int half = row >> 6;
        float* part = p.partial + static_cast<size_t>(split) * 9 * p.Cin * p.Cout;
        const bool have_work = pt1 > pt0;
        if (have_work) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        const int kw = half == 0 ? kw0 : kw1;
        const int ci = p.mode == 0 ? ci0 + row : (row & 63);
        const bool valid = kw >= 0;
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh) {
            float* dst = part + (static_cast<size_t>(kh * 3 + (valid ? kw : 0)) * p.Cin + ci) * p.Cout + co0;
#pragma unroll 1
            for (int c = 0; c < NT / 32; ++c) {
                uint32_t r[32];
                if (have_work) {
                    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(row & ~31) << 16) + kh * NT + c * 32, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                if (valid) {
                    float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        d4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
    if (p.tickets != nullptr) {
        // the unit's region of dw: (kh, M half, 64 or 128 input channels) x NT output channels
        const int rows_half = p.mode == 0 ? 128 : 64, halves = p.mode == 0 ? 1 : (kw1 >= 0 ? 2 : 1);
        wgrad_unit_finish(p, unit, split, 3 * halves * rows_half, NT, [&](int r) {
            const int hf = r / (3 * rows_half), rr = r - hf * 3 * rows_half;
            const int kh_r = rr / rows_half, ci_r = ci0 + rr % rows_half;
            return (static_cast<size_t>(kh_r * 3 + (hf == 0 ? kw0 : kw1)) * p.Cin + ci_r) * p.Cout + co0;
        });
    }
    wgrad_bias_finish(p, smem);
}

// ---------------------------------------------------------------------------------------------
// Cin == 64 variant.  The generic kernel above pairs two kw taps on UMMA M (2 x 64 input channels) and leaves half
// of the second unit's rows on a zero block (75 % useful MMAs), with N = Cout-block of 64..128.  For N = 64 an SS-mode
// MMA is shared-memory-read-bound (A 4 KiB + B 2 KiB per 32 tensor clocks; ncu: tensor pipe 50 % active).  Here the
// three kh taps move to the N side instead: the dZ tile is loaded with a one-row halo above and below and the B
// operand is THREE overlapping row-shifted windows of it (N = 3 x 64, "leading byte offset" = one image row = 2 KiB):
//     D[(kw, ci)][(n, co)] = sum_p X[p + (0, kw-1)][ci] * dZ[p + (n-1, 0)][co]   ==   dW[kh = 2-n][kw][ci][co]
// (substitute q = p + (n-1, 0); rows outside the frame are TMA zero fill on both operands).  Two M128 x N192 MMAs per
// 16 pixels cover all 9 taps (rows {kw0,kw1} and {kw2, zero}): 192 tensor clocks instead of 6 x ~58.
constexpr int kW64XBox = kWgBh * kWgBw * 128;             // 8 KiB: 4 x 16 pixels x 64 channels (no halo)
constexpr int kW64ZBox = (kWgBh + 2) * kWgBw * 128;       // 12 KiB: haloed dZ tile
constexpr int kW64Stage = 3 * kW64XBox + kW64ZBox;        // 36 KiB
constexpr int kW64Stages = 6;

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad64_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz,
                  const WgradArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* zero_blk = smem + kW64Stages * kW64Stage;      // (holds ones, see below)
    uint64_t* bars = reinterpret_cast<uint64_t*>(zero_blk + kW64XBox);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgMaxStages;
    uint64_t* done_bar = bars + 2 * kWgMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int unit = blockIdx.x % p.units;          // 64-wide block of output channels
    const int split = blockIdx.x / p.units;
    const int co0 = unit * 64;
    const int per = (p.num_ptiles + p.splits - 1) / p.splits;
    const int pt0 = split * per;
    const int pt1 = min(p.num_ptiles, pt0 + per);

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_dz);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < kW64Stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    // rows 64..127 of the second MMA pair with no kw tap: instead of zeros they hold bf16 1.0, so those accumulator
    // rows become sum_pixels dZ = the bias gradient (BiasAddGrad for free: the MMAs were issued anyway)
    for (int i = threadIdx.x; i < kW64XBox / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(zero_blk)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                int t = pt;
                const int tw = t % p.tiles_w;
                t /= p.tiles_w;
                const int th = t % p.tiles_h;
                const int img = t / p.tiles_h;
                const int h0 = th * kWgBh, w0 = tw * kWgBw;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], kW64Stage);
                uint8_t* sa = smem + stage * kW64Stage;
                tma_load_5d(&map_dz, &full_bar[stage], sa + 3 * kW64XBox, 0, w0, h0 - 1, co0 / 64, img);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
                    tma_load_5d(&map_x, &full_bar[stage], sa + kw * kW64XBox, 0, w0 + kw - 1, h0, 0, img);
                if (++stage == kW64Stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 192, 1, 1);   // both operands MN-major (K = pixels)
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            const uint32_t smem0 = smem_u32(smem);
            const uint32_t zero_addr = smem_u32(zero_blk);
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem0 + stage * kW64Stage;
                const uint32_t a1_lo = desc_lo_sw128(sa, kW64XBox);                              // {kw0, kw1}
                const uint32_t a2_lo = desc_lo_sw128(sa + 2 * kW64XBox, zero_addr - (sa + 2 * kW64XBox));   // {kw2, zero}
                const uint32_t b_lo = desc_lo_sw128(sa + 3 * kW64XBox, 2048);                    // 3 windows, one image row apart
                const uint32_t acc_flag = pt > pt0 ? 1u : 0u;
#pragma unroll
                for (int k = 0; k < (kWgBh * kWgBw) / 16; ++k) {
                    // K advance of 16 pixel rows = 2048 B on both operands
                    umma_bf16_lohi(tmem_base, a1_lo + k * (2048 >> 4), dhi, b_lo + k * (2048 >> 4), dhi, idesc,
                                   k > 0 ? 1u : acc_flag);
                    umma_bf16_lohi(tmem_base + 192, a2_lo + k * (2048 >> 4), dhi, b_lo + k * (2048 >> 4), dhi, idesc,
                                   k > 0 ? 1u : acc_flag);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == kW64Stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);
        }
    } else if (warp >= 4) {
        const int row = threadIdx.x - 128;        // TMEM lane = D row = (kw block, ci)
        float* part = p.partial + static_cast<size_t>(split) * 9 * p.Cin * p.Cout;
        const bool have_work = pt1 > pt0;
        if (have_work) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        const int ci = row & 63;
#pragma unroll 1
        for (int acc = 0; acc < 2; ++acc) {
            const int kw = acc == 0 ? (row >> 6) : 2;
            const bool valid = acc == 0 || row < 64;
#pragma unroll 1
            for (int n = 0; n < 3; ++n) {             // N block n <-> kh = 2 - n
                float* dst = part + (static_cast<size_t>((2 - n) * 3 + kw) * p.Cin + ci) * p.Cout + co0;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    if (have_work) {
                        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(row & ~31) << 16) + acc * 192 + n * 64 + c * 32, r);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = 0u;
                    }
                    if (valid) {
                        float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            d4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    } else if (row == 64 && n == 1 && p.db_rows != nullptr) {
                        // ones row x un-shifted dZ window: every pixel of this CTA's slice exactly once.  One row per
                        // pixel slice (zeros from a CTA without work); wgrad_reduce_kernel adds the slices in order.
                        float4* d4 = reinterpret_cast<float4*>(p.db_rows + static_cast<size_t>(split) * p.Cout + co0 + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            d4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
    if (p.tickets != nullptr)      // the unit's region of dw: all 9 taps x 64 input channels x its 64 output channels
        wgrad_unit_finish(p, unit, split, 9 * 64, 64, [&](int r) { return static_cast<size_t>(r) * p.Cout + co0; });
    wgrad_bias_finish(p, smem);
}

// dw[i] (+)= sum_s partial[s][i], fixed summation order (deterministic).  Block = 32 float4 columns x 8 split groups:
// thread (c, g) adds splits g, g+8, g+16, ... (several loads in flight), then the 8 group sums are combined in group
// order through shared memory.  (v1 gave one thread a whole column: up to 148 dependent loads in only 36 blocks,
// 38 us for 22 MB; this form is bandwidth-bound.)
// The last `bias_blocks` blocks finish the layer's BiasAddGrad the same way: db = sum of the partial rows that the
// kernel producing dZ (or this wgrad's own ones row) left behind -- db_rows/db_nrows, or a parts buffer whose header
// holds the row count (db_nrows < 0).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float4* __restrict__ partial, float4* __restrict__ dw, long long n4, int splits,
                    int accumulate, int w_blocks, const float* __restrict__ db_rows, int db_nrows, float* __restrict__ db,
                    int Cout) {
    __shared__ float4 red[8][32];
    if (static_cast<int>(blockIdx.x) >= w_blocks) {
        const float* rows = db_nrows < 0 ? db_rows + ABL_PARTS_HEADER : db_rows;
        const int nrows = db_nrows < 0 ? parts_rows(db_rows) : db_nrows;
        parts_reduce_block(rows, nrows, Cout, db, 0, blockIdx.x - w_blocks, reinterpret_cast<float4(*)[4]>(red));
        return;
    }
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    for (long long base = static_cast<long long>(blockIdx.x) * 32; base < n4; base += static_cast<long long>(w_blocks) * 32) {
        const long long i = base + c;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n4) {
#pragma unroll 4
            for (int k = g; k < splits; k += 8) {
                const float4 v = __ldg(partial + k * n4 + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        red[g][c] = s;
        __syncthreads();
        if (g == 0 && i < n4) {
            float4 t = accumulate ? dw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = red[q][c];
                t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
            }
            dw[i] = t;
        }
        __syncthreads();
    }
}

// Partial rows of db[c] = sum_rows dz[row][c] (BiasAddGrad): thread = 8-channel chunk, block-level reduce in shared
// memory (fixed order), ONE parts row per block -- the stand-alone producer for the paths that cannot fuse the column
// sums into the kernel that writes dz.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const uint4* __restrict__ dz, float* __restrict__ parts, long long rows, int C8) {
    extern __shared__ float red[];   // [256/C8 row groups][C8*8]
    const int chunk = threadIdx.x % C8;
    const int rg = threadIdx.x / C8;
    const int rgs = blockDim.x / C8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long r = blockIdx.x * static_cast<long long>(rgs) + rg; r < rows;
         r += static_cast<long long>(gridDim.x) * rgs) {
        const uint4 v = __ldg(dz + r * C8 + chunk);
        acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xFFFF0000u);
        acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xFFFF0000u);
        acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xFFFF0000u);
        acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xFFFF0000u);
    }
    const int C = C8 * 8;
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

static int wgrad_plan(int B, int H, int W, int Cin, int Cout, WgradArgs* a, int* nt_out) {
    if (B <= 0 || Cin % 64 != 0 || Cout % 64 != 0 || Cin > 512 || Cout > 512) return ABL_ERR_SHAPE;
    if (Cin != 64 && Cin % 128 != 0) return ABL_ERR_SHAPE;
    if (H <= 0 || W <= 0) return ABL_ERR_SHAPE;
    a->B = B; a->H = H; a->W = W; a->Cin = Cin; a->Cout = Cout;
    const int nt = (Cout % 128 == 0) ? 128 : 64;            // 3 accumulators x NT <= 512 TMEM columns
    // any frame size: tiles are counted rounding up; TMA zero-fills the pixels of x and dz beyond the frame, and a zero
    // pixel adds nothing to either sum
    a->tiles_h = (H + kWgBh - 1) / kWgBh; a->tiles_w = (W + kWgBw - 1) / kWgBw;
    a->num_ptiles = B * a->tiles_h * a->tiles_w;
    a->mode = Cin == 64 ? 1 : 0;
    a->ci_blocks = Cin == 64 ? 1 : Cin / 128;
    a->co_blocks = Cout / nt;
    a->units = (a->mode == 0 ? 3 * a->ci_blocks : 2) * a->co_blocks;
    const char* generic = getenv("ABL_WGRAD_GENERIC");
    if (Cin == 64 && !(generic && generic[0] == '1')) {     // wgrad64_tc_kernel: unit = 64 output channels
        a->mode = 2;
        a->units = Cout / 64;
    }
    int splits = sm_count() / a->units;
    if (splits < 1) splits = 1;
    if (splits > a->num_ptiles) splits = a->num_ptiles;
    a->splits = splits;
    a->stage_bytes = a->mode == 2 ? kW64Stage : 2 * kWgABox + (nt / 64) * kWgBBox;
    int stages = (227 * 1024 - 512 - kWgABox) / a->stage_bytes;
    if (stages > kWgMaxStages) stages = kWgMaxStages;
    if (stages < 2) return ABL_ERR_SHAPE;
    a->stages = stages;
    *nt_out = nt;
    return ABL_OK;
}

template <int NT>
static int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mdz, const WgradArgs& a, cudaStream_t st) {
    const int smem_bytes = a.stages * a.stage_bytes + kWgABox + 512;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess)
            return check_launch();
        attr_done[dev & 63] = true;
    }
    wgrad_tc_kernel<NT><<<a.units * a.splits, kWgThreads, smem_bytes, st>>>(mx, mdz, a);
    return check_launch();
}

// 5-D view of an NHWC bf16 tensor: (64 ch, W, H, C/64, N); the box stacks `blocks` 64-channel blocks.
static int nhwc_block_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_h, int blocks) {
    const uint64_t dims[5] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)(C / 64), (uint64_t)B};
    const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, 128, (uint64_t)H * W * C * 2};
    const uint32_t box[5] = {64, (uint32_t)kWgBw, (uint32_t)box_h, (uint32_t)blocks, 1};
    return encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace abl

using namespace abl;

// Tickets of the in-kernel slice reduction: zero-initialised module memory, one slot per (device, stream) that has
// called abl_conv3x3_wgrad, so that wgrads on different streams never share counters (two launches on ONE stream are
// ordered; a captured graph keeps the slot of the stream it was captured on).  No free slot -> null -> the caller
// falls back to the separate reduction kernel.
constexpr int kTicketSlots = 16, kTicketWords = 128;            // 2 * units <= 96 words used
__device__ unsigned int g_wgrad_tickets[kTicketSlots][kTicketWords];
static unsigned int* wgrad_ticket_slot(cudaStream_t st) {
    struct Slot { int dev; cudaStream_t st; };
    static Slot slots[kTicketSlots];
    static int used = 0;
    static unsigned int* base[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (base[dev & 63] == nullptr) {
        void* ptr = nullptr;
        if (cudaGetSymbolAddress(&ptr, g_wgrad_tickets) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        base[dev & 63] = static_cast<unsigned int*>(ptr);
    }
    int idx = -1, per_dev = 0;
    for (int i = 0; i < used; ++i) {
        if (slots[i].dev != dev) continue;
        if (slots[i].st == st) { idx = per_dev; break; }
        ++per_dev;
    }
    if (idx < 0) {
        if (used == kTicketSlots) return nullptr;
        slots[used].dev = dev; slots[used].st = st; ++used;
        idx = per_dev;
    }
    return base[dev & 63] + static_cast<size_t>(idx) * kTicketWords;
}

// workspace layout: [splits][9][Cin][Cout] fp32 slice sums | [splits][Cout] fp32 bias rows of the Cin == 64 kernel's ones
// row | a parts buffer for the stand-alone bias pass (db requested, no db_parts given, Cin != 64)
static long long wgrad_ws_dw_bytes(const WgradArgs& a) { return static_cast<long long>(a.splits) * 9 * a.Cin * a.Cout * 4; }
static long long wgrad_ws_rows_bytes(const WgradArgs& a) { return static_cast<long long>(a.splits) * a.Cout * 4; }

extern "C" long long abl_parts_bytes(int C);

extern "C" long long abl_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    WgradArgs a;
    int nt;
    if (wgrad_plan(B, H, W, Cin, Cout, &a, &nt) != ABL_OK) return -1;
    return wgrad_ws_dw_bytes(a) + wgrad_ws_rows_bytes(a) + abl_parts_bytes(Cout);
}

extern "C" int abl_bias_grad(const void* dz, float* db_parts, long long db_parts_bytes, long long rows, int C, void* stream) {
    if (!dz || !db_parts) { set_error("abl_bias_grad: null pointer"); return ABL_ERR_NULL; }
    if (rows <= 0 || (C % 8) != 0 || C > 2048 || (256 % (C / 8)) != 0) {
        set_error("abl_bias_grad: need C % 8 == 0 and C/8 dividing 256"); return ABL_ERR_SHAPE;
    }
    const int C8 = C / 8, rgs = 256 / C8;
    long long blocks = (rows + rgs - 1) / rgs;
    const long long cap = static_cast<long long>(sm_count()) * kPartWaves;
    if (blocks > cap) blocks = cap;
    int r = parts_check(db_parts, db_parts_bytes, static_cast<int>(blocks), C, "abl_bias_grad");
    if (r != ABL_OK) return r;
    bias_grad_kernel<<<static_cast<int>(blocks), 256, rgs * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(dz), db_parts, rows, C8);
    return check_launch();
}

extern "C" int abl_conv3x3_wgrad(const void* x, const void* dz, float* dw, float* db, const float* db_parts,
                                 void* workspace, long long workspace_bytes, int accumulate, int B, int H, int W, int Cin,
                                 int Cout, void* stream) {
    if (!x || !dz || !dw || !workspace) { set_error("abl_conv3x3_wgrad: null pointer"); return ABL_ERR_NULL; }
    if (db_parts && !db) { set_error("abl_conv3x3_wgrad: db_parts given without db"); return ABL_ERR_NULL; }
    WgradArgs a;
    int nt = 0;
    if (wgrad_plan(B, H, W, Cin, Cout, &a, &nt) != ABL_OK) {
        set_error("abl_conv3x3_wgrad: unsupported shape (Cin 64 or a multiple of 128, Cout a multiple of 64, both <= 512)");
        return ABL_ERR_SHAPE;
    }
    const long long need = wgrad_ws_dw_bytes(a) + wgrad_ws_rows_bytes(a) + abl_parts_bytes(Cout);
    if (workspace_bytes < need) { set_error("abl_conv3x3_wgrad: workspace too small"); return ABL_ERR_SHAPE; }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dw) |
         reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(db) | reinterpret_cast<uintptr_t>(db_parts)) & 15) {
        set_error("abl_conv3x3_wgrad: pointers must be 16-byte aligned"); return ABL_ERR_ALIGN;
    }
    a.partial = static_cast<float*>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // slice reduction inside the GEMM kernel (default) or by wgrad_reduce_kernel (ABL_WGRAD_FUSED_REDUCE=0, or no free
    // ticket slot for this stream)
    // Measured on B200 (batch 256, same box): the Cin >= 128 layers gain 1-4 us each (176 / 182 / 106 / 173 us against
    // 179 / 186 / 108 / 175 us with the separate kernel) and lose a launch; the Cin = 64 kernel, whose 74-148 slices per
    // unit leave every CTA only a few rows, was 3-9 us SLOWER, so it keeps the reduction kernel (ABL_WGRAD_FUSED_REDUCE=2
    // forces the in-kernel form there too).
    const char* fr = getenv("ABL_WGRAD_FUSED_REDUCE");
    const bool fuse_reduce = fr ? (fr[0] == '2' || (fr[0] != '0' && a.mode != 2)) : a.mode != 2;
    a.tickets = fuse_reduce ? wgrad_ticket_slot(st) : nullptr;
    a.dw = dw; a.accumulate = accumulate; a.db = db; a.db_src = nullptr; a.db_nrows = 0; a.db_blocks = 0;
    float* ws_rows = reinterpret_cast<float*>(static_cast<char*>(workspace) + wgrad_ws_dw_bytes(a));
    float* ws_parts = reinterpret_cast<float*>(static_cast<char*>(workspace) + wgrad_ws_dw_bytes(a) + wgrad_ws_rows_bytes(a));
    // where db comes from: partial rows the producer of dz left (db_parts), the ones row of the Cin == 64 GEMM, or a
    // stand-alone pass over dz into this call's workspace
    const bool own_rows = db != nullptr && db_parts == nullptr && a.mode == 2;
    a.db_rows = own_rows ? ws_rows : nullptr;
    CUtensorMap mx, mdz;
    int r;
    const float* rows_src = nullptr;
    int nrows = 0;
    if (db != nullptr) {
        if (db_parts != nullptr) { rows_src = db_parts; nrows = -1; }
        else if (own_rows) { rows_src = ws_rows; nrows = a.splits; }
        else {      // before the GEMM: with the fused reduction its first CTAs finish the bias gradient at their end
            r = abl_bias_grad(dz, ws_parts, abl_parts_bytes(Cout), static_cast<long long>(B) * H * W, Cout, stream);
            if (r != ABL_OK) return r;
            rows_src = ws_parts; nrows = -1;
        }
    }
    const int bias_blocks = db != nullptr ? (Cout / 4 + 3) / 4 : 0;
    if (a.tickets != nullptr && db != nullptr && !own_rows) { a.db_src = rows_src; a.db_nrows = nrows; a.db_blocks = bias_blocks; }
    if (a.tickets == nullptr) a.db = nullptr;            // (the kernels' own-rows finish belongs to the fused path)
    if (a.mode == 2) {
        r = nhwc_block_map(&mx, x, B, H, W, Cin, kWgBh, 1);              // plain 4 x 16 pixel tile per kw
        if (r != ABL_OK) return r;
        r = nhwc_block_map(&mdz, dz, B, H, W, Cout, kWgBh + 2, 1);       // haloed dZ tile, one 64-channel block
        if (r != ABL_OK) return r;
        constexpr int smem_bytes = kW64Stages * kW64Stage + kW64XBox + 512;
        static bool attr_done[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr_done[dev & 63]) {
            if (cudaFuncSetAttribute(wgrad64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
                return check_launch();
            attr_done[dev & 63] = true;
        }
        wgrad64_tc_kernel<<<a.units * a.splits, kWgThreads, smem_bytes, st>>>(mx, mdz, a);
        r = check_launch();
    } else {
        r = nhwc_block_map(&mx, x, B, H, W, Cin, kWgBh + 2, a.mode == 0 ? 2 : 1);
        if (r != ABL_OK) return r;
        r = nhwc_block_map(&mdz, dz, B, H, W, Cout, kWgBh, nt / 64);
        if (r != ABL_OK) return r;
        r = (nt == 128) ? launch_wgrad<128>(mx, mdz, a, st) : launch_wgrad<64>(mx, mdz, a, st);
    }
    if (r != ABL_OK || a.tickets != nullptr) return r;
    const long long n4 = 9LL * Cin * Cout / 4;
    long long blocks = (n4 + 31) / 32;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    wgrad_reduce_kernel<<<static_cast<int>(blocks) + bias_blocks, 256, 0, st>>>(
        static_cast<const float4*>(workspace), reinterpret_cast<float4*>(dw), n4, a.splits, accumulate,
        static_cast<int>(blocks), rows_src, nrows, db, Cout);
    return check_launch();
}
