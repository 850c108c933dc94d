// Developer probe, not part of the reference-facing path: how many SM clocks does one tcgen05.mma (M = 128, K = 16,
// bf16, operands in shared memory) take as a function of N and of the issue pattern?  The conv kernels' ceilings are
// set by this number (ncu: N = 64 issues once per ~64 clocks although the math needs 32), so the blocked main loop
// (conv3x3_blk64_kernel) was designed around it; this kernel measures it directly instead of inferring it from
// whole-kernel timings.  One thread issues `iters` groups of four K-steps back to back from fixed operand tiles,
// commits to an mbarrier, and the CTA reports clock64() deltas.
//   mode 0: plain SS MMAs into one accumulator
//   mode 1: weight-stationary pairs (collector b0 fill -> lastuse): two A tiles / two accumulators share each B read
//   mode 2: like 0, but every MMA of a group uses a different 8 KiB-aligned window of B (no operand reuse at all)
//   mode 3: plain SS MMAs issued round-robin over FOUR accumulators (N <= 128), two alternating A tiles, one B
//   mode 4: plain SS MMAs issued round-robin over TWO accumulators (same pattern as mode 1 without .ws)
//   mode 5: weight-stationary quads (collector b0 fill -> use -> use -> lastuse) over four accumulators (N <= 128)
#include <stdint.h>

#include "host_util.h"
#include "ptx.cuh"

namespace abl {

constexpr int kProbeA = 2 * 128 * 128;        // two A tiles [128 rows][128 B]
constexpr int kProbeB = 4 * 256 * 128;        // four B tiles [256 rows][128 B]
constexpr int kProbeSmem = kProbeA + kProbeB + 1024;

__global__ void __launch_bounds__(128, 1)
mma_rate_probe_kernel(int n, int mode, int iters, unsigned long long* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kProbeA;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_b + kProbeB);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (kProbeA + kProbeB) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (warp == 1 && elect_one()) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, static_cast<uint32_t>(n), 0, 0);
        constexpr uint32_t dhi = desc_hi_sw128(1024);
        const uint32_t a_lo = desc_lo_sw128(smem_u32(smem_a), 16);
        const uint32_t b_lo = desc_lo_sw128(smem_u32(smem_b), 16);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode == 3 || mode == 4) {
                    const int nacc = mode == 3 ? 4 : 2;
                    for (int q = 0; q < nacc; ++q)
                        umma_bf16_lohi(tmem_base + q * (512 / nacc), a_lo + (q & 1) * ((128 * 128) >> 4) + k * 2, dhi, b_lo + k * 2,
                                       dhi, idesc, (it | k) != 0 ? 1u : 0u);
                } else if (mode == 5) {
                    umma_bf16_lohi_ws<0>(tmem_base, a_lo + k * 2, dhi, b_lo + k * 2, dhi, idesc, (it | k) != 0 ? 1u : 0u);
                    umma_bf16_lohi_ws<2>(tmem_base + 128, a_lo + ((128 * 128) >> 4) + k * 2, dhi, b_lo + k * 2, dhi, idesc,
                                         (it | k) != 0 ? 1u : 0u);
                    umma_bf16_lohi_ws<2>(tmem_base + 256, a_lo + k * 2, dhi, b_lo + k * 2, dhi, idesc, (it | k) != 0 ? 1u : 0u);
                    umma_bf16_lohi_ws<1>(tmem_base + 384, a_lo + ((128 * 128) >> 4) + k * 2, dhi, b_lo + k * 2, dhi, idesc,
                                         (it | k) != 0 ? 1u : 0u);
                } else if (mode == 1) {
                    umma_bf16_lohi_ws<0>(tmem_base, a_lo + k * 2, dhi, b_lo + k * 2, dhi, idesc, (it | k) != 0 ? 1u : 0u);
                    umma_bf16_lohi_ws<1>(tmem_base + 256, a_lo + ((128 * 128) >> 4) + k * 2, dhi, b_lo + k * 2, dhi, idesc,
                                         (it | k) != 0 ? 1u : 0u);
                } else {
                    const uint32_t boff = mode == 2 ? static_cast<uint32_t>(((it * 4 + k) & 3) * ((256 * 128) >> 4)) : 0u;
                    umma_bf16_lohi(tmem_base, a_lo + k * 2, dhi, b_lo + boff + k * 2, dhi, idesc, (it | k) != 0 ? 1u : 0u);
                }
            }
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        out[blockIdx.x * 2] = static_cast<unsigned long long>(t1 - t0);        // issue loop (back-pressure included)
        out[blockIdx.x * 2 + 1] = static_cast<unsigned long long>(t2 - t0);    // until the last MMA retired
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}


// ---------------------------------------------------------------------------------------------- HBM write probe
// How fast can SM-issued stores fill a buffer?  Every HBM-bound kernel of the path writes at least as much as it reads
// (block1_conv1 forward: 92 % of its traffic), and the copy peak of MEASURED_PEAKS.json is half reads.  Modes:
//   0 st.global.v4 (default policy)   1 st.global.cs.v4 (streaming)   2 st.global.wt.v4 (write-through)
//   3 st.global.v4 + L2 evict_first policy   4 st.global.v8 (256-bit)   5 TMA bulk store, 4 KiB per warp from shared
//   memory (what the conv epilogues do)   6 as 5 with an L2 evict_first hint
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__global__ void __launch_bounds__(512, 1) write_probe_kernel(uint4* __restrict__ dst, long long n16, int mode) {
    extern __shared__ __align__(1024) uint8_t psm[];
    const uint4 val = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    if (mode <= 3) {
        const uint64_t pol = policy_evict_first();
        for (long long i = tid; i < n16; i += nthreads) {
            if (mode == 0) dst[i] = val;
            else if (mode == 1) __stcs(dst + i, val);
            else if (mode == 2) __stwt(dst + i, val);
            else
                asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst + i), "r"(val.x),
                             "r"(val.y), "r"(val.z), "r"(val.w), "l"(pol)
                             : "memory");
        }
    } else if (mode == 4) {
        for (long long i = tid; 2 * i + 1 < n16; i += nthreads)
            asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(dst + 2 * i), "r"(val.x) : "memory");
    } else {
        // every warp owns a 4 KiB shared-memory slice and streams it out with bulk stores, two in flight
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        uint4* slice = reinterpret_cast<uint4*>(psm + warp * 4096);
        for (int j = lane; j < 256; j += 32) slice[j] = val;
        fence_proxy_async_smem();
        __syncwarp();
        const uint64_t pol = policy_evict_first();
        const long long nchunks = n16 / 256, nwarps = nthreads >> 5;
        if (lane == 0) {
            for (long long c = tid >> 5; c < nchunks; c += nwarps) {
                if (mode == 5)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(dst + c * 256),
                                 "r"(smem_u32(slice))
                                 : "memory");
                else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], 4096, %2;" ::"l"(
                                     dst + c * 256),
                                 "r"(smem_u32(slice)), "l"(pol)
                                 : "memory");
                tma_store_commit();
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            tma_store_wait_all();
        }
    }
}
}  // namespace abl

using namespace abl;

extern "C" int abl_dev_mma_rate_probe(int n, int mode, int iters, int ctas, unsigned long long* out_cycles, void* stream) {
    if (!out_cycles) { set_error("abl_dev_mma_rate_probe: null pointer"); return ABL_ERR_NULL; }
    if (n < 16 || n > 256 || (n & 15) || mode < 0 || mode > 5 || iters <= 0 || ctas <= 0) {
        set_error("abl_dev_mma_rate_probe: need N in 16..256 (multiple of 16), mode 0..5, iters > 0, ctas > 0");
        return ABL_ERR_SHAPE;
    }
    if ((mode == 1 || mode == 5) && n != 64 && n != 128 && n != 256) { set_error("abl_dev_mma_rate_probe: .ws needs N in {64,128,256}"); return ABL_ERR_SHAPE; }
    if ((mode == 3 || mode == 5) && n > 128) { set_error("abl_dev_mma_rate_probe: four accumulators need N <= 128"); return ABL_ERR_SHAPE; }
    if (cudaFuncSetAttribute(mma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kProbeSmem) != cudaSuccess)
        return check_launch();
    mma_rate_probe_kernel<<<ctas, 128, kProbeSmem, static_cast<cudaStream_t>(stream)>>>(n, mode, iters, out_cycles);
    return check_launch();
}

extern "C" int abl_dev_write_probe(void* dst, long long bytes, int mode, int ctas_per_sm, void* stream) {
    if (!dst) { set_error("abl_dev_write_probe: null pointer"); return ABL_ERR_NULL; }
    if (bytes <= 0 || (bytes & 4095) || mode < 0 || mode > 6 || ctas_per_sm < 1 || ctas_per_sm > 4 ||
        (reinterpret_cast<uintptr_t>(dst) & 31)) {
        set_error("abl_dev_write_probe: need bytes a multiple of 4096, mode 0..6, 1..4 CTAs per SM, 32-byte alignment");
        return ABL_ERR_SHAPE;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = mode >= 5 ? 16 * 4096 : 0;
    if (cudaFuncSetAttribute(write_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4096) != cudaSuccess)
        return check_launch();
    write_probe_kernel<<<sms * ctas_per_sm, 512, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<uint4*>(dst),
                                                                                           bytes / 16, mode);
    return check_launch();
}
