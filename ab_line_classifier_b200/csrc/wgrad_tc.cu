// Filter gradient of a 3x3/s1/SAME convolution (TF Conv2DBackpropFilter run by model.fit's tape,
// reference call site src/models/models.py:427,432) as a tcgen05 GEMM whose REDUCTION dimension is
// the pixel index:
//     dW[tap][ci][co] = sum_{n,h,w} X[n,h+kh-1,w+kw-1,ci] * dZ[n,h,w,co]
// Both operands are read straight from the NHWC tensors with TMA boxes {64 ch, box_w, box_h, 1}
// (SWIZZLE_128B): in shared memory a box is "pixels x 64 channels", i.e. the MN-major canonical
// layout with K = pixel rows of 128 B, so no transpose pass is needed:
//     A = shifted X box  (M = 64 ci per box, two boxes stacked -> UMMA M = 128, LBO = box bytes)
//     B = dZ boxes       (N = NT co = NT/64 boxes, LBO = box bytes)
// A CTA owns one "unit" = up to ACC accumulators (A box pairs = (tap, ci-chunk) pairs) x one co
// block of NT columns, all fed by the SAME dZ tile, and one slice of the pixel range.  Partial
// sums go to a workspace with plain vector stores; a second kernel reduces the slices in a fixed
// order (deterministic) into the fp32 HWIO gradient.
#include "host_util.h"
#include "ptx.cuh"

namespace abl {

constexpr int kWgKPix = 32;                       // pixels (GEMM K) per pipeline stage
constexpr int kWgBoxBytes = kWgKPix * 128;        // 4 KiB: one {64 ch x 32 px} box
constexpr int kWgThreads = 256;
constexpr int kWgMaxStages = 8;

struct WgradArgs {
    int B, H, W, Cin, Cout;
    int box_h, box_w, tiles_h, tiles_w;   // pixel-tile geometry (box_h*box_w == kWgKPix)
    int num_ptiles;                       // B*tiles_h*tiles_w
    int chunks;                           // Cin/64
    int nboxes;                           // 9*chunks
    int npairs;                           // ceil(nboxes/2)
    int acc;                              // accumulators (pairs) per unit
    int pair_groups;                      // ceil(npairs/acc)
    int co_blocks;                        // Cout/NT
    int units;                            // pair_groups*co_blocks
    int splits;                           // pixel slices per unit; grid = units*splits
    int stages;
    int stage_bytes;
    float* partial;                       // [splits][9][Cin][Cout] fp32
};

template <int NT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz,
                const WgradArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgMaxStages;
    uint64_t* done_bar = bars + 2 * kWgMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int unit = blockIdx.x % p.units;
    const int split = blockIdx.x / p.units;
    const int co_blk = unit % p.co_blocks;
    const int pgroup = unit / p.co_blocks;
    const int pair0 = pgroup * p.acc;
    const int npair_here = min(p.acc, p.npairs - pair0);
    const int co0 = co_blk * NT;
    // contiguous slice of pixel tiles for this split
    const int per = (p.num_ptiles + p.splits - 1) / p.splits;
    const int pt0 = split * per;
    const int pt1 = min(p.num_ptiles, pt0 + per);
    const int a_bytes = 2 * p.acc * kWgBoxBytes;   // A region of a stage (fixed stride, even if fewer pairs)

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
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const int nbox_here = min(2 * npair_here, p.nboxes - 2 * pair0);
            const uint32_t tx = static_cast<uint32_t>(nbox_here + NT / 64) * kWgBoxBytes;
            for (int pt = pt0; pt < pt1; ++pt) {
                int t = pt;
                const int tw = t % p.tiles_w;
                t /= p.tiles_w;
                const int th = t % p.tiles_h;
                const int img = t / p.tiles_h;
                const int h0 = th * p.box_h, w0 = tw * p.box_w;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], tx);
                uint8_t* sa = smem + stage * p.stage_bytes;
                uint8_t* sb = sa + a_bytes;
                for (int j = 0; j < NT / 64; ++j)
                    tma_load_4d(&map_dz, &full_bar[stage], sb + j * kWgBoxBytes, co0 + j * 64, w0, h0, img);
                for (int b = 0; b < nbox_here; ++b) {
                    const int box = 2 * pair0 + b;
                    const int tap = box / p.chunks, chunk = box - tap * p.chunks;
                    const int kh = tap / 3, kw = tap - kh * 3;
                    tma_load_4d(&map_x, &full_bar[stage], sa + b * kWgBoxBytes, chunk * 64, w0 + kw - 1, h0 + kh - 1,
                                img);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);   // both operands MN-major
            // MN-major SWIZZLE_128B: 8-row (pixel) groups 1024 B apart (SBO), 64-channel blocks one box apart (LBO)
            constexpr uint32_t dhi = desc_hi_sw128(1024);
            const uint32_t lo0 = desc_lo_sw128(smem_u32(smem), kWgBoxBytes);
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_lo = lo0 + stage * (p.stage_bytes >> 4);
                const uint32_t b_lo = a_lo + (a_bytes >> 4);
                const uint32_t acc_flag = pt > pt0 ? 1u : 0u;
                for (int a = 0; a < npair_here; ++a) {
#pragma unroll
                    for (int k = 0; k < kWgKPix / 16; ++k) {
                        // K advance of 16 pixel rows = 2048 B
                        umma_bf16_lohi(tmem_base + a * NT, a_lo + a * (2 * kWgBoxBytes >> 4) + k * (2048 >> 4), dhi,
                                       b_lo + k * (2048 >> 4), dhi, idesc, k > 0 ? 1u : acc_flag);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);
        }
    } else if (warp >= 4) {
        const int row = threadIdx.x - 128;        // TMEM lane = D row
        const int ewarp = row >> 5;
        float* part = p.partial + static_cast<size_t>(split) * 9 * p.Cin * p.Cout;
        const bool have_work = pt1 > pt0;
        if (have_work) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        for (int a = 0; a < npair_here; ++a) {
            const int box = 2 * (pair0 + a) + (row >> 6);
            const bool valid = box < p.nboxes;
            const int tap = valid ? box / p.chunks : 0;
            const int chunk = valid ? box - tap * p.chunks : 0;
            const int ci = chunk * 64 + (row & 63);
            float* dst = part + (static_cast<size_t>(tap) * p.Cin + ci) * p.Cout + co0;
#pragma unroll 1
            for (int c = 0; c < NT / 32; ++c) {
                uint32_t r[32];
                if (have_work) {
                    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + a * NT + c * 32, r);
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
}

// dw[i] (+)= sum_s partial[s][i], fixed summation order.
__global__ void wgrad_reduce_kernel(const float4* __restrict__ partial, float4* __restrict__ dw, long long n4,
                                    int splits, int accumulate) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 s = accumulate ? dw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < splits; ++k) {
            const float4 v = __ldg(partial + k * n4 + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        dw[i] = s;
    }
}

// db[c] += sum_rows dz[row][c]  (BiasAddGrad); thread = 8-channel chunk, block-level smem reduce, atomics.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const uint4* __restrict__ dz, float* __restrict__ db, long long rows, int C8) {
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
    if (rg < rgs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[rg * C + chunk * 8 + j] = acc[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < rgs; ++g) s += red[g * C + c];
        atomicAdd(db + c, s);
    }
}

static int pick_kbox(int H, int W, int* bh, int* bw) {
    for (int w = kWgKPix; w >= 8; w >>= 1) {
        const int h = kWgKPix / w;
        if (W % w == 0 && H % h == 0) { *bw = w; *bh = h; return ABL_OK; }
    }
    return ABL_ERR_SHAPE;
}

static int wgrad_plan(int B, int H, int W, int Cin, int Cout, WgradArgs* a, int* nt_out) {
    if (B <= 0 || Cin % 64 != 0 || Cout % 64 != 0 || Cin > 512 || Cout > 512) return ABL_ERR_SHAPE;
    a->B = B; a->H = H; a->W = W; a->Cin = Cin; a->Cout = Cout;
    if (pick_kbox(H, W, &a->box_h, &a->box_w) != ABL_OK) return ABL_ERR_SHAPE;
    const int nt = (Cout % 128 == 0) ? 128 : 64;
    a->tiles_h = H / a->box_h; a->tiles_w = W / a->box_w;
    a->num_ptiles = B * a->tiles_h * a->tiles_w;
    a->chunks = Cin / 64;
    a->nboxes = 9 * a->chunks;
    a->npairs = (a->nboxes + 1) / 2;
    const int max_acc = 512 / nt;                                   // TMEM columns
    a->pair_groups = (a->npairs + max_acc - 1) / max_acc;
    a->acc = (a->npairs + a->pair_groups - 1) / a->pair_groups;     // balance pairs over groups
    a->co_blocks = Cout / nt;
    a->units = a->pair_groups * a->co_blocks;
    int splits = sm_count() / a->units;
    if (splits < 1) splits = 1;
    if (splits > a->num_ptiles) splits = a->num_ptiles;
    a->splits = splits;
    a->stage_bytes = (2 * a->acc + nt / 64) * kWgBoxBytes;
    int stages = (227 * 1024 - 1024) / a->stage_bytes;
    if (stages > kWgMaxStages) stages = kWgMaxStages;
    if (stages < 2) return ABL_ERR_SHAPE;
    a->stages = stages;
    *nt_out = nt;
    return ABL_OK;
}

template <int NT>
static int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mdz, const WgradArgs& a, cudaStream_t st) {
    const int smem_bytes = a.stages * a.stage_bytes + 256;
    static int attr_bytes[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_bytes[dev & 63] < smem_bytes) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess)
            return check_launch();
        attr_bytes[dev & 63] = 227 * 1024;
    }
    wgrad_tc_kernel<NT><<<a.units * a.splits, kWgThreads, smem_bytes, st>>>(mx, mdz, a);
    return check_launch();
}

}  // namespace abl

using namespace abl;

extern "C" long long abl_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    WgradArgs a;
    int nt;
    if (wgrad_plan(B, H, W, Cin, Cout, &a, &nt) != ABL_OK) return -1;
    return static_cast<long long>(a.splits) * 9 * Cin * Cout * 4;
}

extern "C" int abl_bias_grad(const void* dz, float* db, long long rows, int C, void* stream) {
    if (!dz || !db) { set_error("abl_bias_grad: null pointer"); return ABL_ERR_NULL; }
    if (rows <= 0 || (C % 8) != 0 || C > 2048 || (256 % (C / 8)) != 0) {
        set_error("abl_bias_grad: need C % 8 == 0 and C/8 dividing 256"); return ABL_ERR_SHAPE;
    }
    const int C8 = C / 8, rgs = 256 / C8;
    long long blocks = (rows + rgs - 1) / rgs;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    bias_grad_kernel<<<static_cast<int>(blocks), 256, rgs * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(dz), db, rows, C8);
    return check_launch();
}

extern "C" int abl_conv3x3_wgrad(const void* x, const void* dz, float* dw, float* db, void* workspace,
                                 long long workspace_bytes, int accumulate, int B, int H, int W, int Cin, int Cout,
                                 void* stream) {
    if (!x || !dz || !dw || !workspace) { set_error("abl_conv3x3_wgrad: null pointer"); return ABL_ERR_NULL; }
    WgradArgs a;
    int nt = 0;
    if (wgrad_plan(B, H, W, Cin, Cout, &a, &nt) != ABL_OK) {
        set_error("abl_conv3x3_wgrad: unsupported shape (Cin, Cout multiples of 64 <= 512; W % 8 == 0)");
        return ABL_ERR_SHAPE;
    }
    const long long need = static_cast<long long>(a.splits) * 9 * Cin * Cout * 4;
    if (workspace_bytes < need) { set_error("abl_conv3x3_wgrad: workspace too small"); return ABL_ERR_SHAPE; }
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dw) |
         reinterpret_cast<uintptr_t>(workspace)) & 15) {
        set_error("abl_conv3x3_wgrad: pointers must be 16-byte aligned"); return ABL_ERR_ALIGN;
    }
    a.partial = static_cast<float*>(workspace);
    CUtensorMap mx, mdz;
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box[4] = {64, (uint32_t)a.box_w, (uint32_t)a.box_h, 1};
        int r = encode_tiled(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != ABL_OK) return r;
    }
    {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
        const uint32_t box[4] = {64, (uint32_t)a.box_w, (uint32_t)a.box_h, 1};
        int r = encode_tiled(&mdz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dz, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != ABL_OK) return r;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int r = (nt == 128) ? launch_wgrad<128>(mx, mdz, a, st) : launch_wgrad<64>(mx, mdz, a, st);
    if (r != ABL_OK) return r;
    const long long n4 = 9LL * Cin * Cout / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    wgrad_reduce_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const float4*>(workspace),
                                                                reinterpret_cast<float4*>(dw), n4, a.splits, accumulate);
    r = check_launch();
    if (r != ABL_OK) return r;
    if (db) return abl_bias_grad(dz, db, static_cast<long long>(B) * H * W, Cout, stream);
    return ABL_OK;
}
