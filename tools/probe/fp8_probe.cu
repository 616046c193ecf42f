// Probe of the operand conventions the fp16 + fp8 scorer format relies on (run once on a B200, see profiles/):
//   (1) tcgen05.mma kind::f8f6f4 with the A operand in tensor memory: 4 e4m3 per 32-bit column, K index 0 in the low byte;
//   (2) K-major SWIZZLE_128B shared-memory B operand of bytes: 128 K-elements per row, + 32 bytes per K = 32 step;
//   (3) kind::f16 and kind::f8f6f4 MMAs accumulating into the same float32 accumulator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o fp8_probe fp8_probe.cu ; prints max |error| per mode.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../../ssp_bayesopt_b200/csrc/sspbo_tc.cuh"
using namespace sspbo_tc;

constexpr int M = 128, N = 64, KH = 64, K8 = 128;

__device__ __forceinline__ void umma_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}

// mode bit 0: fp16 MMAs, bit 1: fp8 MMAs
__global__ void __launch_bounds__(128, 1)
k_probe(const __half* Ah, const uint8_t* A8, const __half* Bh, const uint8_t* B8, float* out, int mode) {
    __shared__ __align__(1024) uint8_t bh_s[N * 128];
    __shared__ __align__(1024) uint8_t b8_s[N * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, row = threadIdx.x;
    // B tiles in the layout TMA writes with SWIZZLE_128B: 16-byte chunk c of row r lands at chunk c ^ (r % 8)
    for (int i = threadIdx.x; i < N * 8; i += 128) {
        const int r = i >> 3, c = i & 7, pc = c ^ (r & 7);
        *(uint4*)(bh_s + r * 128 + pc * 16) = *(const uint4*)((const uint8_t*)Bh + r * 128 + c * 16);
        *(uint4*)(b8_s + r * 128 + pc * 16) = *(const uint4*)(B8 + r * 128 + c * 16);
    }
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    // A operand: row = TMEM lane; fp16 pairs in columns [64, 96), e4m3 quadruples in columns [96, 128)
    {
        uint32_t w[8];
        for (int g = 0; g < 4; ++g) {
            for (int i = 0; i < 8; ++i) w[i] = ((const uint32_t*)(Ah + (size_t)row * KH))[g * 8 + i];
            tmem_st8(lane_addr + 64 + g * 8, w);
        }
        for (int g = 0; g < 4; ++g) {
            for (int i = 0; i < 8; ++i) w[i] = ((const uint32_t*)(A8 + (size_t)row * K8))[g * 8 + i];
            tmem_st8(lane_addr + 96 + g * 8, w);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc(M, N);
            const uint64_t dh = make_desc_sw128(smem_u32(bh_s)), d8 = make_desc_sw128(smem_u32(b8_s));
            uint32_t acc = 0;
            if (mode & 1)
                for (int kk = 0; kk < 4; ++kk) { umma_f16_ts(tmem, tmem + 64 + 8 * kk, dh + 2 * kk, idesc, acc); acc = 1; }
            if (mode & 2)
                for (int kk = 0; kk < 4; ++kk) { umma_f8_ts(tmem, tmem + 96 + 8 * kk, d8 + 2 * kk, idesc, acc); acc = 1; }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int g = 0; g < N; g += 16) {
        uint32_t v[16];
        tmem_ld16(lane_addr + g, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(size_t)row * N + g + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
    }
}

static double e4m3_value(uint8_t b) {
    const int s = b >> 7, e = (b >> 3) & 15, m = b & 7;
    const double v = e == 0 ? m * ldexp(1.0, -9) : (1.0 + m / 8.0) * ldexp(1.0, e - 7);
    return s ? -v : v;
}

int main() {
    std::vector<__half> Ah(M * KH), Bh(N * KH);
    std::vector<uint8_t> A8(M * K8), B8(N * K8);
    srand(1);
    auto rnd = []() { return (rand() % 2001 - 1000) / 1000.0f; };
    for (auto& v : Ah) v = __float2half(rnd());
    for (auto& v : Bh) v = __float2half(rnd() * 8.f);
    auto rb = []() { uint8_t b; do { b = (uint8_t)(rand() & 0xff); } while ((b & 0x7f) == 0x7f || ((b >> 3) & 15) > 10); return b; };
    for (auto& v : A8) v = rb();
    for (auto& v : B8) v = rb();
    __half *dAh, *dBh; uint8_t *dA8, *dB8; float* dout;
    cudaMalloc(&dAh, Ah.size() * 2); cudaMalloc(&dBh, Bh.size() * 2); cudaMalloc(&dA8, A8.size()); cudaMalloc(&dB8, B8.size());
    cudaMalloc(&dout, M * N * 4);
    cudaMemcpy(dAh, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dBh, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dA8, A8.data(), A8.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB8, B8.data(), B8.size(), cudaMemcpyHostToDevice);
    int bad = 0;
    for (int mode = 1; mode <= 3; ++mode) {
        cudaMemset(dout, 0, M * N * 4);
        k_probe<<<1, 128>>>(dAh, dA8, dBh, dB8, dout, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
        std::vector<float> out(M * N);
        cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                if (mode & 1) for (int k = 0; k < KH; ++k) ref += (double)__half2float(Ah[m * KH + k]) * __half2float(Bh[n * KH + k]);
                if (mode & 2) for (int k = 0; k < K8; ++k) ref += e4m3_value(A8[m * K8 + k]) * e4m3_value(B8[n * K8 + k]);
                maxerr = fmax(maxerr, fabs(ref - out[m * N + n])); maxref = fmax(maxref, fabs(ref));
            }
        printf("mode %d (%s%s): max |err| = %.3e, max |ref| = %.3e -> %s\n", mode, mode & 1 ? "f16 " : "", mode & 2 ? "e4m3" : "",
               maxerr, maxref, maxerr <= 2e-5 * maxref ? "OK" : "MISMATCH");
        bad += !(maxerr <= 2e-5 * maxref);
    }
    return bad ? 1 : 0;
}
