// Issue-rate probe: cycles per tcgen05.mma (M = 128, N = 256, A in tensor memory) for kind::f16 (K = 16) and kind::f8f6f4
// (K = 32), one CTA, operands are whatever shared / tensor memory holds.  Also N = 128.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../ssp_bayesopt_b200/csrc/sspbo_tc.cuh"
using namespace sspbo_tc;

__device__ __forceinline__ void umma_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_f8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_rate(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 128 / 4 + 128 * 128 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // B tile (256 x 128 B) + A tile
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_slot;
    {   // A operand region: columns [256, 320)
        uint32_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = 0x3c003c00u;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int g = 0; g < 8; ++g) tmem_st8(lane_addr + 256 + g * 8, w);
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0 && elect_one()) {
        const uint64_t db = make_desc_sw128(smem_u32(smem)), da = make_desc_sw128(smem_u32(smem + 256 * 128));
        uint32_t phase = 0;
        int slot = 0;
        for (int N = 256; N >= 64; N >>= 1) {
            const uint32_t idesc = make_idesc(128, N);
            for (int kind = 0; kind < 6; ++kind) {
                if (kind >= 4 && N == 256) { out[slot++] = 0; continue; }      // two / four accumulators need N <= 128 / 64 columns each
                if (kind == 5 && N == 128) { out[slot++] = 0; continue; }
                long long t0 = clock64();
                for (int i = 0; i < iters; ++i) {
                    const uint32_t a = tmem + 256 + 8 * (i & 3);
                    // kinds 4, 5: consecutive MMAs accumulate into DIFFERENT accumulators (2 or 4 of them): is the floor of
                    // ~106 cycles per MMA at N <= 128 the latency of the accumulate chain?
                    if (kind == 4) umma_f16_ts(tmem + (uint32_t)N * (i & 1), a, db + 2 * (i & 3), idesc, 1);
                    else if (kind == 5) umma_f16_ts(tmem + (uint32_t)N * (i & 3), a, db + 2 * (i & 3), idesc, 1);
                    else if (kind == 0) umma_f16_ts(tmem, a, db + 2 * (i & 3), idesc, 1);
                    else if (kind == 1) umma_f8_ts(tmem, a, db + 2 * (i & 3), idesc, 1);
                    else if (kind == 2) umma_f16(tmem, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
                    else umma_f8_ss(tmem, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
                }
                umma_commit(smem_u32(&bar));
                mbar_wait(smem_u32(&bar), phase);
                phase ^= 1;
                long long t1 = clock64();
                out[slot++] = t1 - t0;
            }
        }
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)); }
}

int main() {
    long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
    const int iters = 2048;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k_rate<<<1, 128, 100 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[6] = {"f16 K=16 A:tmem", "e4m3 K=32 A:tmem", "f16 K=16 A:smem", "e4m3 K=32 A:smem",
                            "f16 A:tmem, 2 accumulators in turn", "f16 A:tmem, 4 accumulators in turn"};
    int s = 0;
    for (int N = 256; N >= 64; N >>= 1)
        for (int k = 0; k < 6; ++k, ++s) if (h[s]) printf("N=%3d %-36s %.1f cycles per MMA\n", N, names[k], (double)h[s] / iters);
    return 0;
}
