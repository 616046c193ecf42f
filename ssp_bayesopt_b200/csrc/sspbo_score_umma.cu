// K2, tcgen05/TMEM engine: fused encode + BLR acquisition + argmax for sm_100a.
//
// Per CTA (persistent, one per SM) and per tile of 128 candidates:
//   generator warps  theta_f = A_f . x in float64 turns, exact range reduction, MUFU sin/cos -> the psi tile
//                    (A operand, 128 x 64 per k-block) written straight into 128B-swizzled shared memory as a
//                    split-fp16 pair (hi, lo); psi never exists in HBM.  The first pass also accumulates
//                    mu = m' . psi in float32.
//   TMA warp         streams the matching 64-column slabs of the split-fp16 factor R (hi, lo) from L2.
//   MMA warp         one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16, fp16 x fp16 -> fp32 in TMEM):
//                    Y += psi_hi R_hi^T + psi_hi R_lo^T + psi_lo R_hi^T   (3-term split: ~2^-22 operand precision)
//   epilogue warps   tcgen05.ld the accumulator chunk, q += sum_j Y_j^2, then var = 1/beta + q / r_scale^2,
//                    acq = lam var / (sqrt(var+gamma) + sqrt(gamma)), score = mu + acq, warp-shuffle argmax.
// The d output rows are processed in n_chunks chunks of BN rows with two TMEM accumulators (2 x 256 columns) so the
// epilogue of chunk c overlaps the MMAs of chunk c+1; the psi k-block is regenerated per chunk (it is cheaper than
// the 3 x MMA time it overlaps with, and 128 x 1024 x 2 x fp16 does not fit on chip).
//
// Replaces SSPAgent.eval (agents/ssp_agent.py:125-137) + np.argmax (experiments/demo_blr_gif.py:129-136).
#include "sspbo_internal.cuh"
#include <cuda.h>

namespace {

constexpr int BM = 128;                 // candidates per tile (UMMA M)
constexpr int BK = 64;                  // operand columns per k-block = one 128-byte swizzle atom of fp16
constexpr int GEN_WARPS = 16;                     // 512 threads: (row, quarter) -> 8 frequencies of a k-block each
constexpr int EPI_WARPS = 4;
constexpr int WARP_TMA = 0, WARP_MMA = 1, WARP_EPI0 = 2, WARP_GEN0 = WARP_EPI0 + EPI_WARPS;
constexpr int THREADS = 32 * (WARP_GEN0 + GEN_WARPS);            // 704
constexpr int GEN_THREADS = 32 * GEN_WARPS;                      // 512
constexpr int A_TILE_BYTES = BM * BK * 2;                        // 16 KiB
constexpr int MAX_N = 8;                                         // domain dimension held in registers
constexpr uint32_t TMEM_COLS = 512;
constexpr int ACC_COLS = 256;                                    // columns per accumulator buffer

struct Params {
    const void* x; int x_f64; long long N; long long index0;
    const double* At_op; const float* mp_op;
    int n, K, KC_pad, BN, n_chunks, n_kb, stages;
    float inv_beta, lam, gamma_t, inv_rscale2;
    float *mu, *var, *acq;
    unsigned long long* best;
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
// same, with a short sleep between polls: for roles that wait for a long time (epilogue, producers)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        __nanosleep(100);
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 ; [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1) ; [32,46) stride byte
//   offset >> 4 (1024 B between 8-row groups) ; [46,48) version = 1 ; [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = f32 (bits [4,6) = 1), A = B = f16 (0), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Pipe {                            // ring position shared by every role (all iterate the same sequence)
    int stage = 0; uint32_t phase = 0;
    __device__ __forceinline__ void advance(int stages) { if (++stage == stages) { stage = 0; phase ^= 1; } }
};

// ------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(THREADS, 1)
k_score_umma(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle; offset arithmetic (not integer casts) keeps the address space known
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_tile_bytes = p.BN * BK * 2;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
    uint8_t* tables = smem + (size_t)p.stages * stage_bytes;
    double* At_s = (double*)tables;                                         // (KC_pad/2) x n
    float* mp_s = (float*)(At_s + (size_t)(p.KC_pad / 2) * p.n);            // KC_pad
    float* mu_part = mp_s + p.KC_pad;                                       // [2 tile parities][4 quarters][128]
    uint64_t* bars = (uint64_t*)(mu_part + 2 * 4 * BM);                     // offset is a multiple of 8 bytes (see table_bytes)
    // barrier map: full_a[s], full_b[s], empty[s] for s < stages ; tmem_full[2] ; tmem_empty[2]
    const uint32_t bar0 = smem_u32(bars);
    auto full_a = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto full_b = [&](int s) { return bar0 + 8u * (uint32_t)(p.stages + s); };
    auto empty_ = [&](int s) { return bar0 + 8u * (uint32_t)(2 * p.stages + s); };
    auto tmem_full = [&](int b) { return bar0 + 8u * (uint32_t)(3 * p.stages + b); };
    auto tmem_empty = [&](int b) { return bar0 + 8u * (uint32_t)(3 * p.stages + 2 + b); };
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * p.stages + 4);

    // ---- one-time setup ----
    for (int i = threadIdx.x; i < (p.KC_pad / 2) * p.n; i += THREADS) At_s[i] = p.At_op[i];
    for (int i = threadIdx.x; i < p.KC_pad; i += THREADS) mp_s[i] = p.mp_op[i];
    if (warp == WARP_TMA && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_a(s), GEN_THREADS); mbar_init(full_b(s), 1); mbar_init(empty_(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), EPI_WARPS * 32); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long n_tiles = (p.N + BM - 1) / BM;

    if (warp == WARP_TMA) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            Pipe pipe;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int c = 0; c < p.n_chunks; ++c)
                    for (int kb = 0; kb < p.n_kb; ++kb) {
                        mbar_wait_relaxed(empty_(pipe.stage), pipe.phase ^ 1);
                        uint8_t* st = smem + (size_t)pipe.stage * stage_bytes;
                        mbar_expect_tx(full_b(pipe.stage), 2u * (uint32_t)b_tile_bytes);
                        tma_load_2d(smem_u32(st + 2 * A_TILE_BYTES), &map_hi, full_b(pipe.stage), kb * BK, c * p.BN);
                        tma_load_2d(smem_u32(st + 2 * A_TILE_BYTES + b_tile_bytes), &map_lo, full_b(pipe.stage), kb * BK, c * p.BN);
                        pipe.advance(p.stages);
                    }
        }
    } else if (warp == WARP_MMA) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, p.BN);
            Pipe pipe;
            uint32_t acc_phase[2] = {0, 0};
            int acc_iter = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int c = 0; c < p.n_chunks; ++c, ++acc_iter) {
                    const int buf = acc_iter & 1;
                    mbar_wait(tmem_empty(buf), acc_phase[buf] ^ 1);       // epilogue drained this accumulator
                    acc_phase[buf] ^= 1;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS);
                    for (int kb = 0; kb < p.n_kb; ++kb) {
                        mbar_wait(full_a(pipe.stage), pipe.phase);
                        mbar_wait(full_b(pipe.stage), pipe.phase);
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + (size_t)pipe.stage * stage_bytes);
                        const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + A_TILE_BYTES);
                        const uint64_t b_hi = make_desc_sw128(st + 2 * A_TILE_BYTES);
                        const uint64_t b_lo = make_desc_sw128(st + 2 * A_TILE_BYTES + b_tile_bytes);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            const uint64_t off = (uint64_t)((kk * 32) >> 4);        // 16 fp16 = 32 bytes along K
                            umma_f16(d_tmem, a_hi + off, b_hi + off, idesc, (kb | kk) != 0);
                            umma_f16(d_tmem, a_hi + off, b_lo + off, idesc, 1);
                            umma_f16(d_tmem, a_lo + off, b_hi + off, idesc, 1);
                        }
                        umma_commit(empty_(pipe.stage));                  // frees the stage when these MMAs retire
                        if (kb == p.n_kb - 1) umma_commit(tmem_full(buf));
                        pipe.advance(p.stages);
                    }
                }
        }
    } else if (warp < WARP_GEN0) {
        // ================================ epilogue ================================
        const int quad = warp & 3;                                        // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        uint32_t acc_phase[2] = {0, 0};
        int acc_iter = 0, tile_iter = 0;
        unsigned long long my_best = 0ull;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
            float q = 0.f;
            for (int c = 0; c < p.n_chunks; ++c, ++acc_iter) {
                const int buf = acc_iter & 1;
                mbar_wait_relaxed(tmem_full(buf), acc_phase[buf]);
                acc_phase[buf] ^= 1;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * ACC_COLS);
                for (int g = 0; g < p.BN; g += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + (uint32_t)g, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) { float y = __uint_as_float(v[i]); q = fmaf(y, y, q); }
                }
                tc_fence_before();
                mbar_arrive(tmem_empty(buf));
            }
            const long long gi = tile * BM + row;
            if (gi < p.N) {
                const float* mpart = mu_part + (tile_iter & 1) * 4 * BM;
                float mu = (mpart[row] + mpart[BM + row]) + (mpart[2 * BM + row] + mpart[3 * BM + row]);
                float var = p.inv_beta + q * p.inv_rscale2;                           // blr.py:96
                float acq = p.lam * var / (sqrtf(var + p.gamma_t) + sqrtf(p.gamma_t)); // ssp_agent.py:136, cancellation-free
                float score = mu + acq;
                if (p.mu) p.mu[gi] = mu;
                if (p.var) p.var[gi] = var;
                if (p.acq) p.acq[gi] = acq;
                unsigned long long key = sspbo_pack(score, (unsigned long long)(p.index0 + gi));
                if (key > my_best) my_best = key;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, my_best, o);
            if (other > my_best) my_best = other;
        }
        if (lane == 0 && my_best && p.best) atomicMax(p.best, my_best);
    } else {
        // ================================ psi generators ================================
        const int gt = threadIdx.x - WARP_GEN0 * 32;
        const int row = gt & (BM - 1), quarter = gt >> 7;                 // quarter q: frequencies 8q .. 8q+7 of the k-block
        const int n = p.n;
        const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t sw = (uint32_t)(row & 7);
        // one 16-byte chunk of cos (chunk q) and one of sin (chunk 4+q) per stage, XOR-swizzled by the row
        const uint32_t off_cos = row_off + (((uint32_t)quarter ^ sw) << 4);
        const uint32_t off_sin = row_off + (((uint32_t)(quarter + 4) ^ sw) << 4);
        Pipe pipe;
        int tile_iter = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
            double x[MAX_N];
            const long long gi = tile * BM + row;
#pragma unroll
            for (int a = 0; a < MAX_N; ++a) {
                x[a] = 0.0;
                if (a < n && gi < p.N)
                    x[a] = p.x_f64 ? ((const double*)p.x)[gi * n + a] : (double)((const float*)p.x)[gi * n + a];
            }
            float mu = 0.f;
            for (int c = 0; c < p.n_chunks; ++c)
                for (int kb = 0; kb < p.n_kb; ++kb) {
                    const double* Ablk = At_s + (size_t)(kb * 32 + quarter * 8) * n;
                    float cs[8], sn[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const double* Af = Ablk + (size_t)i * n;
                        double t = 0.0;
#pragma unroll
                        for (int a = 0; a < MAX_N; ++a)
                            if (a < n) t = fma(Af[a], x[a], t);
                        float tr = (float)(t - rint(t));                  // exact reduction to [-1/2, 1/2] turns
                        float ang = tr * 6.283185307179586f;
                        cs[i] = __cosf(ang);
                        sn[i] = __sinf(ang);
                    }
                    if (c == 0) {
                        const float* mcos = mp_s + kb * BK + quarter * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            mu = fmaf(mcos[i], cs[i], mu);
                            mu = fmaf(mcos[32 + i], sn[i], mu);
                        }
                    }
                    uint32_t ch[4], cl[4], sh[4], sl[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __half2 h = __floats2half2_rn(cs[2 * i], cs[2 * i + 1]);
                        float2 hf = __half22float2(h);
                        __half2 l = __floats2half2_rn(cs[2 * i] - hf.x, cs[2 * i + 1] - hf.y);
                        ch[i] = *reinterpret_cast<uint32_t*>(&h); cl[i] = *reinterpret_cast<uint32_t*>(&l);
                        h = __floats2half2_rn(sn[2 * i], sn[2 * i + 1]);
                        hf = __half22float2(h);
                        l = __floats2half2_rn(sn[2 * i] - hf.x, sn[2 * i + 1] - hf.y);
                        sh[i] = *reinterpret_cast<uint32_t*>(&h); sl[i] = *reinterpret_cast<uint32_t*>(&l);
                    }
                    // values are ready in registers before the slot is claimed: the wait overlaps the math above
                    mbar_wait(empty_(pipe.stage), pipe.phase ^ 1);
                    uint8_t* st = smem + (size_t)pipe.stage * stage_bytes;
                    *reinterpret_cast<uint4*>(st + off_cos) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
                    *reinterpret_cast<uint4*>(st + off_sin) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
                    *reinterpret_cast<uint4*>(st + A_TILE_BYTES + off_cos) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
                    *reinterpret_cast<uint4*>(st + A_TILE_BYTES + off_sin) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
                    if (c == 0 && kb == p.n_kb - 1) mu_part[(tile_iter & 1) * 4 * BM + quarter * BM + row] = mu;
                    fence_proxy_async();                                   // generic-proxy stores -> visible to the tensor core
                    mbar_arrive(full_a(pipe.stage));
                    pipe.advance(p.stages);
                }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------ host side
struct UmmaState {
    CUtensorMap map_hi, map_lo;
    const void* rh = nullptr; const void* rl = nullptr;
    int KC_pad = 0, NJ_pad = 0, BN = 0;
    int max_smem = 0;
    bool attr_set = false;
};

size_t table_bytes(const sspbo_ctx* ctx) {
    return (size_t)(ctx->KC_pad / 2) * ctx->n * sizeof(double) + (size_t)ctx->KC_pad * sizeof(float) +
           2 * 4 * BM * sizeof(float) + (3 * 4 + 4) * 8 + 16;
}
size_t stage_bytes_of(const sspbo_ctx* ctx) { return 2 * (size_t)A_TILE_BYTES + 2 * (size_t)ctx->BN * BK * 2; }

int make_map(sspbo_ctx* ctx, CUtensorMap* map, const void* base) {
    cuuint64_t gdim[2] = {(cuuint64_t)ctx->KC_pad, (cuuint64_t)ctx->NJ_pad};
    cuuint64_t gstride[1] = {(cuuint64_t)ctx->KC_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)ctx->BN};
    cuuint32_t estr[2] = {1, 1};
    // resolved through the runtime so libsspbo.so has no link-time dependency on libcuda.so.1 (absent on CPU-only hosts)
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            SSPBO_FAIL(ctx, SSPBO_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (encode_fn)fn;
    }
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SSPBO_FAIL(ctx, SSPBO_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SSPBO_OK;
}

}  // namespace

int sspbo_score_umma_supported(const sspbo_ctx* ctx) {
    if (!ctx->has_factor || ctx->L != 1 || ctx->n > MAX_N || ctx->K < 1) return 0;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, ctx->device) != cudaSuccess || major != 10) return 0;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    if (table_bytes(ctx) + 2 * stage_bytes_of(ctx) + 1024 > (size_t)max_smem) return 0;
    return 1;
}

void sspbo_umma_release(sspbo_ctx* ctx) {
    if (ctx->umma) { delete (UmmaState*)ctx->umma; ctx->umma = nullptr; }
}

int sspbo_score_umma_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, int64_t index0, float lam, float gamma_t,
                            float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st) {
    UmmaState* us = (UmmaState*)ctx->umma;
    if (!us) { us = new UmmaState(); ctx->umma = us; }
    if (us->rh != ctx->Rh || us->rl != ctx->Rl || us->KC_pad != ctx->KC_pad || us->NJ_pad != ctx->NJ_pad || us->BN != ctx->BN) {
        int rc;
        if ((rc = make_map(ctx, &us->map_hi, ctx->Rh)) || (rc = make_map(ctx, &us->map_lo, ctx->Rl))) return rc;
        us->rh = ctx->Rh; us->rl = ctx->Rl; us->KC_pad = ctx->KC_pad; us->NJ_pad = ctx->NJ_pad; us->BN = ctx->BN;
    }
    if (!us->max_smem) cudaDeviceGetAttribute(&us->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    const size_t tb = table_bytes(ctx), sb = stage_bytes_of(ctx);
    int stages = (int)(((size_t)us->max_smem - 1024 - tb) / sb);
    if (stages > 4) stages = 4;
    if (stages < 2) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "tcgen05 engine: shared memory does not fit two stages");
    const size_t smem = 1024 + (size_t)stages * sb + tb;
    if (!us->attr_set) {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_score_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, us->max_smem));
        us->attr_set = true;
    }
    Params p;
    p.x = x; p.x_f64 = (x_dtype == SSPBO_F64); p.N = N; p.index0 = index0;
    p.At_op = ctx->At_op; p.mp_op = ctx->mp_op;
    p.n = ctx->n; p.K = ctx->K; p.KC_pad = ctx->KC_pad; p.BN = ctx->BN; p.n_chunks = ctx->n_chunks;
    p.n_kb = ctx->KC_pad / BK; p.stages = stages;
    p.inv_beta = (float)(1.0 / ctx->beta); p.lam = lam; p.gamma_t = gamma_t;
    p.inv_rscale2 = (float)(1.0 / (ctx->r_scale * ctx->r_scale));
    p.mu = mu; p.var = var; p.acq = acq; p.best = best;
    const long long tiles = (N + BM - 1) / BM;
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    k_score_umma<<<grid, THREADS, smem, st>>>(us->map_hi, us->map_lo, p);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}
