// K2, tcgen05/TMEM engine: fused encode + BLR acquisition + argmax for sm_100a.
//
// Per CTA (persistent, one per SM) and per tile of 128 candidates:
//   generator warps  theta_f = A_f . x in 64-bit fixed point: A and x are Q31.32 integers, the low 64 bits of the
//                    sum of products are the phase in turns modulo 1 (range reduction is the integer wrap-around,
//                    exact for any |x| < 2^31; resolution 2^-33 turns per unit of A or x), then MUFU sin/cos -> the
//                    psi k-block (A operand, 128 x 64) written straight into 128B-swizzled shared memory as a
//                    split-fp16 pair (hi, lo); psi never exists in HBM.  Each thread owns 2 candidate rows x 8
//                    frequencies, so one shared-memory read of a phase-matrix entry feeds two multiply-add chains.
//                    The first pass also accumulates mu = m' . psi in float32.
//   TMA warp         streams the 64-column slabs of the split-fp16 factor R (hi, lo) from L2 into the B ring.
//   MMA warp         one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16, fp16 x fp16 -> fp32 in TMEM):
//                    Y += psi_hi R_hi^T + psi_hi R_lo^T + psi_lo R_hi^T   (3-term split: ~2^-22 operand precision)
//   epilogue warps   tcgen05.ld the accumulator, q += sum_j Y_j^2, then var = 1/beta + q / r_scale^2,
//                    acq = lam var / (sqrt(var+gamma) + sqrt(gamma)), score = mu + acq, warp-shuffle argmax.
// The d output rows are processed in chunks of BN rows, two chunks (= both 256-column TMEM accumulators) per pass
// over the contraction dimension, so every generated psi k-block is consumed by two MMAs groups before it is
// regenerated for the next pair of chunks (128 x 1024 x 2 fp16 does not fit on chip).  The factor is upper
// triangular (Cholesky of the permuted psi-space covariance), so chunk c only needs k-blocks >= kb_start[c]:
// 40 of the 64 (chunk, k-block) MMA groups and 24 of the 32 psi k-blocks at d = 1023.
//
// Fast phase path (P32): A in Q.fa and x in Q.fx 32-bit fixed point with fa + fx = 55: one IMAD.WIDE per axis
// accumulates the phase in a 64-bit register whose bits [32, 55) are the top 23 bits of the phase in turns.
//
// Replaces SSPAgent.eval (agents/ssp_agent.py:125-137) + np.argmax (experiments/demo_blr_gif.py:129-136).
#include "sspbo_internal.cuh"
#include "sspbo_tc.cuh"
#include <stdlib.h>

namespace {
using namespace sspbo_tc;

constexpr int BM = 128;                 // candidates per tile (UMMA M)
constexpr int BK = 64;                  // operand columns per k-block = one 128-byte swizzle atom of fp16
#ifndef SSPBO_GEN_ROWS
#define SSPBO_GEN_ROWS 2
#endif
constexpr int GEN_ROWS = SSPBO_GEN_ROWS;   // candidate rows per generator thread (1 or 2)
constexpr int ROW_GROUPS = 4 / GEN_ROWS;   // generator warps per quarter of a k-block
constexpr int GEN_WARPS = 4 * ROW_GROUPS;  // thread (row group, quarter) -> GEN_ROWS rows x 8 frequencies of a k-block
constexpr int EPI_WARPS = 4;
constexpr int WARP_TMA = 0, WARP_MMA = 1, WARP_EPI0 = 2, WARP_GEN0 = WARP_EPI0 + EPI_WARPS;
constexpr int THREADS = 32 * (WARP_GEN0 + GEN_WARPS);            // 448 (GEN_ROWS = 2) or 704
constexpr int A_TILE_BYTES = BM * BK * 2;                        // 16 KiB (one of hi / lo)
constexpr int A_SLOT_BYTES = 2 * A_TILE_BYTES;
constexpr int NA_MIN = 2, NA_MAX = 4;
// tiles of mu in flight (power of two): 8 when a tile has fewer than 8 k-blocks, else 2
static inline int mu_depth_for(int n_kb) { return n_kb < 8 ? 8 : 2; }                            // A ring slots (as many as fit after the B ring)
constexpr int MAX_NB = 6;                                        // B ring slots (as many as fit)
constexpr int MAX_N = 8;                                         // domain dimension (compile-time specialised)
constexpr uint32_t TMEM_COLS = 512;
constexpr int ACC_COLS = 256;                                    // columns per accumulator buffer

struct Params {
    const void* x; int x_f64; long long N; long long index0;
    const long long* Aq_op; const float* mp_op;
    // trajectory mode (TRAJ): candidates pre-converted to Q31.32 (N x L x n), per-timestep phase offsets Q0.64 (L x KC_pad/2)
    const unsigned long long* xq; const unsigned long long* tauq; int L;
    int K, KC_pad, BN, n_chunks, n_kb, nb;
    int kb_start[8];                    // first k-block holding non-zeros of the (upper triangular) factor, per output chunk
    short nz_rows[32];                  // factor rows that can be non-zero in k-block kb (= components with column < 64 (kb+1))
    float inv_beta, lam, gamma_t, inv_rscale2;
    float *mu, *var, *acq;
    unsigned long long* best;
    int na;                             // A ring slots
    int mu_depth;                       // tiles of mu in flight between generator and epilogue
    float* enc_out; int enc_pitch; float enc_inv_scale;   // encode mode: the B operand is the inverse-DFT basis, Y = phi is stored
    unsigned long long* stats;          // device counters (out-of-bound candidates of the 32-bit phase path)
    // 32-bit fixed-point phase path
    const int* A32_op; double x_scale;
    int debug;                          // SSPBO_DEBUG experiments (sspbo_internal.cuh); constant 0 in product builds
};

// ------------------------------------------------------------------------------------------ kernel
template <int NDIM, bool TRAJ, bool P32>
__global__ void __launch_bounds__(THREADS, 1)
k_score_umma(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle; offset arithmetic (not integer casts) keeps the address space known
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform: role branches and the MMA issuer loop stay on the uniform datapath
    const int b_tile_bytes = p.BN * BK * 2;
    const int b_slot_bytes = 2 * b_tile_bytes;
    const int NA = p.na;
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + NA * A_SLOT_BYTES;
    uint8_t* tables = b_ring + (size_t)p.nb * b_slot_bytes;
    long long* Aq_s = (long long*)tables;                                   // (KC_pad/2) x NDIM, Q31.32 turns per unit
    float* mp_s = (float*)(Aq_s + (size_t)(p.KC_pad / 2) * NDIM);           // KC_pad
    // mu of a tile travels generator -> epilogue through shared memory.  The generators run up to NA k-blocks ahead of
    // the MMAs and the MMAs up to two tiles ahead of the epilogue, so with few k-blocks per tile several tiles are in
    // flight: MU_DEPTH buffers (the generator of tile t + MU_DEPTH can only start after the epilogue of tile t).
    float* mu_part = mp_s + p.KC_pad;                                       // [MU_DEPTH tiles][4 quarters][128]
    const int MU_DEPTH = p.mu_depth;
    uint64_t* bars = (uint64_t*)(mu_part + MU_DEPTH * 4 * BM);
    // barrier map: a_full[NA] a_empty[NA] b_full[nb] b_empty[nb] tmem_full[2] tmem_empty[2]
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (uint32_t)(NA + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (uint32_t)(2 * NA + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * NA + p.nb + s); };
    auto tmem_full = [&](int b) { return bar0 + 8u * (uint32_t)(2 * NA + 2 * p.nb + b); };
    auto tmem_empty = [&](int b) { return bar0 + 8u * (uint32_t)(2 * NA + 2 * p.nb + 2 + b); };
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * NA + 2 * p.nb + 4);

    // ---- one-time setup ----
    if constexpr (P32) {
        int* A32_s = (int*)tables;
        for (int i = threadIdx.x; i < (p.KC_pad / 2) * NDIM; i += THREADS) A32_s[i] = p.A32_op[i];
    } else {
        for (int i = threadIdx.x; i < (p.KC_pad / 2) * NDIM; i += THREADS) Aq_s[i] = p.Aq_op[i];
    }
    for (int i = threadIdx.x; i < p.KC_pad; i += THREADS) mp_s[i] = p.mp_op[i];
    if (warp == WARP_TMA && lane == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(a_full(s), GEN_WARPS); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.nb; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
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
    const int n_pairs = (p.n_chunks + 1) / 2;
    // accumulator of chunk cc of a pair: with a single chunk per tile the two buffers alternate by tile instead
    auto acc_of = [&](int tile_iter, int cc) { return p.n_chunks == 1 ? (tile_iter & 1) : cc; };

    if (warp == WARP_TMA) {
        // ================================ TMA producer (B ring) ================================
        if (lane == 0) {
            Ring rb;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int pr = 0; pr < n_pairs; ++pr)
                    for (int kb = p.kb_start[2 * pr]; kb < p.n_kb; ++kb)
                        for (int cc = 0; cc < 2 && 2 * pr + cc < p.n_chunks; ++cc) {
                            if (kb < p.kb_start[2 * pr + cc]) continue;   // zero block of the triangular factor
                            mbar_wait_relaxed(b_empty(rb.slot), rb.phase ^ 1);
                            uint8_t* st = b_ring + (size_t)rb.slot * b_slot_bytes;
                            mbar_expect_tx(b_full(rb.slot), (uint32_t)b_slot_bytes);
                            tma_load_2d(smem_u32(st), &map_hi, b_full(rb.slot), kb * BK, (2 * pr + cc) * p.BN);
                            tma_load_2d(smem_u32(st + b_tile_bytes), &map_lo, b_full(rb.slot), kb * BK, (2 * pr + cc) * p.BN);
                            rb.advance(p.nb);
                        }
        }
    } else if (warp == WARP_MMA) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            Ring ra, rb;
            uint32_t acc_phase[2] = {0, 0};
            int tile_iter = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter)
                for (int pr = 0; pr < n_pairs; ++pr) {
                    const int ncc = (2 * pr + 1 < p.n_chunks) ? 2 : 1;
                    for (int kb = p.kb_start[2 * pr]; kb < p.n_kb; ++kb) {
                        mbar_wait(a_full(ra.slot), ra.phase);
                        const uint32_t sa = smem_u32(a_ring + (size_t)ra.slot * A_SLOT_BYTES);
                        const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + A_TILE_BYTES);
                        for (int cc = 0; cc < ncc; ++cc) {
                            const int kb_first = p.kb_start[2 * pr + cc];
                            if (kb < kb_first) continue;                  // zero block of the triangular factor
                            if (kb == kb_first) {                         // first use of this accumulator in the pair:
                                const int buf = acc_of(tile_iter, cc);    // the epilogue must have drained it (waited for
                                mbar_wait(tmem_empty(buf), acc_phase[buf] ^ 1);   // lazily, so the other accumulator's
                                acc_phase[buf] ^= 1;                      // epilogue overlaps these MMAs)
                            }
                            mbar_wait(b_full(rb.slot), rb.phase);
                            tc_fence_after();
                            const uint32_t sb = smem_u32(b_ring + (size_t)rb.slot * b_slot_bytes);
                            const uint64_t b_hi = make_desc_sw128(sb), b_lo = make_desc_sw128(sb + b_tile_bytes);
                            const uint32_t d_tmem = tmem_base + (uint32_t)(acc_of(tile_iter, cc) * ACC_COLS);
                            // Rows of this chunk beyond the triangle are zero in this k-block: shrink N (multiple of 16).
                            // The first k-block of a chunk runs at full N with accumulate = 0 so every column is initialised.
                            int n_eff = p.BN;
                            if (kb != kb_first) {
                                int nz = (int)p.nz_rows[kb] - (2 * pr + cc) * p.BN;
                                nz = (nz + 15) & ~15;
                                n_eff = nz < p.BN ? nz : p.BN;
                            }
                            const uint32_t idesc = make_idesc(BM, n_eff);
#pragma unroll
                            for (int kk = 0; kk < BK / 16 && !(SSPBO_DBG_BITS(p) & 1); ++kk) {
                                const uint64_t off = (uint64_t)((kk * 32) >> 4);    // 16 fp16 = 32 bytes along K
                                umma_f16(d_tmem, a_hi + off, b_hi + off, idesc, (kb != kb_first) || (kk != 0));
                                umma_f16(d_tmem, a_hi + off, b_lo + off, idesc, 1);
                                umma_f16(d_tmem, a_lo + off, b_hi + off, idesc, 1);
                            }
                            umma_commit(b_empty(rb.slot));                // frees the B slot when these MMAs retire
                            if (kb == p.n_kb - 1) umma_commit(tmem_full(acc_of(tile_iter, cc)));
                            rb.advance(p.nb);
                        }
                        umma_commit(a_empty(ra.slot));                    // frees the A slot
                        ra.advance(NA);
                    }
                }
        }
    } else if (warp < WARP_GEN0) {
        // ================================ epilogue ================================
        const int quad = warp & 3;                                        // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        uint32_t acc_phase[2] = {0, 0};
        int tile_iter = 0;
        unsigned long long my_best = 0ull;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
            float q = 0.f;
            for (int c = 0; c < p.n_chunks; ++c) {
                const int buf = acc_of(tile_iter, c & 1);
                mbar_wait_relaxed(tmem_full(buf), acc_phase[buf]);
                acc_phase[buf] ^= 1;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * ACC_COLS);
                for (int g = 0; g < p.BN; g += 64) {                      // BN % 16 == 0; up to 64 columns in flight
                    uint32_t v[64];
                    const int nld = min(4, (p.BN - g) >> 4);
#pragma unroll
                    for (int l = 0; l < 4; ++l)
                        if (l < nld) tmem_ld16(taddr + (uint32_t)(g + 16 * l), v + 16 * l);
                    tmem_ld_wait();
                    if (p.enc_out) {                                      // encode mode: phi row -> HBM (float32, 64 B per load)
                        const long long ge = tile * BM + row;
                        if (ge < p.N) {
                            float* orow = p.enc_out + (size_t)ge * p.enc_pitch + c * p.BN + g;
#pragma unroll
                            for (int l = 0; l < 4; ++l)
                                if (l < nld) {
#pragma unroll
                                    for (int i = 0; i < 16; i += 4)
                                        *reinterpret_cast<float4*>(orow + 16 * l + i) =
                                            make_float4(__uint_as_float(v[16 * l + i]) * p.enc_inv_scale,
                                                        __uint_as_float(v[16 * l + i + 1]) * p.enc_inv_scale,
                                                        __uint_as_float(v[16 * l + i + 2]) * p.enc_inv_scale,
                                                        __uint_as_float(v[16 * l + i + 3]) * p.enc_inv_scale);
                                }
                        }
                        continue;
                    }
                    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                    for (int l = 0; l < 4; ++l)
                        if (l < nld) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                float y0 = __uint_as_float(v[16 * l + i]), y1 = __uint_as_float(v[16 * l + i + 1]);
                                float y2 = __uint_as_float(v[16 * l + i + 2]), y3 = __uint_as_float(v[16 * l + i + 3]);
                                q0 = fmaf(y0, y0, q0); q1 = fmaf(y1, y1, q1); q2 = fmaf(y2, y2, q2); q3 = fmaf(y3, y3, q3);
                            }
                        }
                    q += (q0 + q1) + (q2 + q3);
                }
                tc_fence_before();
                mbar_arrive(tmem_empty(buf));
            }
            const long long gi = tile * BM + row;
            if (gi < p.N && !p.enc_out) {
                const float* mpart = mu_part + (tile_iter & (MU_DEPTH - 1)) * 4 * BM;
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
        // ================================ psi generators (A ring) ================================
        // Thread (row group, quarter q) produces frequencies 8q .. 8q+7 of GEN_ROWS rows for every k-block (one
        // shared-memory read of a phase-matrix entry feeds GEN_ROWS multiply-add chains).  Software pipeline per
        // k-block: probe the slot's "empty" barrier without blocking, do the math into registers, publish the PREVIOUS
        // k-block (its shared-memory stores have long completed, so the proxy fence is cheap), then store this one.
        const int gw = warp - WARP_GEN0;
        const int r0 = (gw % ROW_GROUPS) * 32 + lane, quarter = gw / ROW_GROUPS;
        constexpr uint32_t ROW_STEP = 32 * ROW_GROUPS;                    // rows r0, r0 + ROW_STEP, ...; ROW_STEP % 8 == 0
        const uint32_t row_off = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128);
        const uint32_t sw = (uint32_t)(r0 & 7);
        // one 16-byte chunk of cos (chunk q) and one of sin (chunk 4+q) per row and stage, XOR-swizzled by the row
        const uint32_t off_cos = row_off + (((uint32_t)quarter ^ sw) << 4);
        const uint32_t off_sin = row_off + (((uint32_t)(quarter + 4) ^ sw) << 4);
        const int n_kb = p.n_kb, L = p.L;
        const long long N = p.N;
        Ring ra;
        int tile_iter = 0;
        bool pending = false; uint32_t pending_bar = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
            unsigned long long x0[GEN_ROWS][NDIM];                        // Q31.32 candidates (two's complement)
            int xq0[GEN_ROWS][NDIM];                                      // Q.fx candidates of the 32-bit phase path
            long long g[GEN_ROWS];
#pragma unroll
            for (int r = 0; r < GEN_ROWS; ++r) g[r] = tile * BM + r0 + r * ROW_STEP;
            if constexpr (P32) {
                bool oor = false;
#pragma unroll
                for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                    for (int a = 0; a < NDIM; ++a) {
                        double v0 = 0.0;
                        if (g[r] < N) v0 = p.x_f64 ? ((const double*)p.x)[g[r] * NDIM + a] : (double)((const float*)p.x)[g[r] * NDIM + a];
                        v0 *= p.x_scale;
                        oor |= !(fabs(v0) < 2147483647.0);
                        xq0[r][a] = __double2int_rn(v0);
                    }
                if (oor && quarter == 0) atomicAdd(p.stats + SSPBO_STAT_OOR, 1ull);     // outside the declared |x| bound
            } else if constexpr (!TRAJ) {
#pragma unroll
                for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                    for (int a = 0; a < NDIM; ++a) {
                        double v0 = 0.0;
                        if (g[r] < N) v0 = p.x_f64 ? ((const double*)p.x)[g[r] * NDIM + a] : (double)((const float*)p.x)[g[r] * NDIM + a];
                        x0[r][a] = (unsigned long long)__double2ll_rn(v0 * 4294967296.0);
                    }
            }
            if constexpr (!TRAJ) {                                        // next tile's rows: pull them towards the SM now
                if (quarter == 0) {
#pragma unroll
                    for (int r = 0; r < GEN_ROWS; ++r) {
                        const long long gn = g[r] + (long long)gridDim.x * BM;
                        if (gn < N) {
                            const char* nx = (const char*)p.x + (size_t)gn * NDIM * (p.x_f64 ? 8 : 4);
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                        }
                    }
                }
            }
            float mu0[GEN_ROWS];
#pragma unroll
            for (int r = 0; r < GEN_ROWS; ++r) mu0[r] = 0.f;
            for (int pr = 0; pr < n_pairs; ++pr)
                for (int kb = p.kb_start[2 * pr]; kb < n_kb; ++kb) {
                    const uint32_t empty_bar = a_empty(ra.slot);
                    const bool slot_free = mbar_test_wait(empty_bar, ra.phase ^ 1);         // early, non-blocking probe
                    float cs[GEN_ROWS][8], sn[GEN_ROWS][8];
                    if (SSPBO_DBG_BITS(p) & 2) {
#pragma unroll
                        for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                            for (int i = 0; i < 8; ++i) { cs[r][i] = 0.25f * (float)(i + kb); sn[r][i] = 0.125f * (float)(r + kb); }
                    } else if constexpr (P32) {
                        const int* A32blk = (const int*)tables + (size_t)(kb * 32 + quarter * 8) * NDIM;
                        long long t[GEN_ROWS][8];                         // phase in turns at bit 55
#pragma unroll
                        for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                            for (int i = 0; i < 8; ++i) t[r][i] = 0ll;
#pragma unroll
                        for (int a = 0; a < NDIM; ++a) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int Afa = A32blk[i * NDIM + a];     // one broadcast read feeds every row
#pragma unroll
                                for (int r = 0; r < GEN_ROWS; ++r) t[r][i] += (long long)Afa * (long long)xq0[r][a];   // IMAD.WIDE
                            }
                        }
#pragma unroll
                        for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                // bits [32, 55) = top 23 bits of the phase mod 1; flipping the top one adds 1/2 turn, and
                                // with the exponent of 1.0 they are the mantissa of f in [1, 2): (f - 1.5) 2 pi is the
                                // phase in radians in [-pi, pi).  One LOP3: (hi & 0x7fffff) ^ 0x3fc00000.
                                const float f0 = __uint_as_float(lop3_and_xor((unsigned int)((unsigned long long)t[r][i] >> 32), 0x7fffffu, 0x3fc00000u));
                                const float a0 = fmaf(f0, 6.283185307179586f, -9.42477796076938f);
                                cs[r][i] = __cosf(a0); sn[r][i] = __sinf(a0);
                            }
                    } else if constexpr (!TRAJ) {
                        const unsigned long long* Ablk = (const unsigned long long*)Aq_s + (size_t)(kb * 32 + quarter * 8) * NDIM;
                        unsigned long long t[GEN_ROWS][8];                // phase in turns, Q0.64, wraps modulo 1
#pragma unroll
                        for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                            for (int i = 0; i < 8; ++i) t[r][i] = 0ull;
#pragma unroll
                        for (int a = 0; a < NDIM; ++a) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const unsigned long long Afa = Ablk[i * NDIM + a];
#pragma unroll
                                for (int r = 0; r < GEN_ROWS; ++r) t[r][i] += Afa * x0[r][a];
                            }
                        }
#pragma unroll
                        for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                // top 32 bits as a signed fraction of a turn in [-1/2, 1/2) -> radians in [-pi, pi)
                                float a0 = (float)(int)(t[r][i] >> 32) * 1.4629180792671596e-9f;   // 2 pi / 2^32
                                cs[r][i] = __cosf(a0); sn[r][i] = __sinf(a0);
                            }
                    } else {
                        // trajectory: the spectrum is the sum over the L waypoints of unit phasors
                        // exp(i (theta_f(x_j) + tau_jf))  (agents/ssp_traj_agent.py:145-163 in the Fourier domain)
                        const unsigned long long* Ablk = (const unsigned long long*)Aq_s + (size_t)(kb * 32 + quarter * 8) * NDIM;
#pragma unroll
                        for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                            for (int i = 0; i < 8; ++i) { cs[r][i] = 0.f; sn[r][i] = 0.f; }
                        const int nf = p.KC_pad >> 1;
                        for (int j = 0; j < L; ++j) {
#pragma unroll
                            for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                                for (int a = 0; a < NDIM; ++a)
                                    x0[r][a] = (g[r] < N) ? __ldg(p.xq + ((size_t)g[r] * L + j) * NDIM + a) : 0ull;
                            const unsigned long long* tq = p.tauq + (size_t)j * nf + kb * 32 + quarter * 8;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const unsigned long long tau = __ldg(tq + i);
#pragma unroll
                                for (int r = 0; r < GEN_ROWS; ++r) {
                                    unsigned long long t0 = tau;
#pragma unroll
                                    for (int a = 0; a < NDIM; ++a) t0 += Ablk[i * NDIM + a] * x0[r][a];
                                    float a0 = (float)(int)(t0 >> 32) * 1.4629180792671596e-9f;
                                    cs[r][i] += __cosf(a0); sn[r][i] += __sinf(a0);
                                }
                            }
                        }
                    }
                    if (pr == 0) {
                        const float* mcos = mp_s + kb * BK + quarter * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float mc = mcos[i], ms = mcos[32 + i];
#pragma unroll
                            for (int r = 0; r < GEN_ROWS; ++r) { mu0[r] = fmaf(mc, cs[r][i], mu0[r]); mu0[r] = fmaf(ms, sn[r][i], mu0[r]); }
                        }
                    }
                    uint32_t w[GEN_ROWS][16];                             // [cos hi x4, cos lo x4, sin hi x4, sin lo x4]
#pragma unroll
                    for (int r = 0; r < GEN_ROWS; ++r)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            split_pair(cs[r][2 * i], cs[r][2 * i + 1], w[r][i], w[r][4 + i]);
                            split_pair(sn[r][2 * i], sn[r][2 * i + 1], w[r][8 + i], w[r][12 + i]);
                        }
                    if (pending) {                                        // publish the previous k-block: every thread
                        fence_proxy_async();                              // fences its own stores towards the tensor core,
                        __syncwarp();                                     // one lane per warp arrives
                        if (lane == 0) mbar_arrive(pending_bar);
                    }
                    if (!slot_free) mbar_wait(empty_bar, ra.phase ^ 1);
                    uint8_t* st = a_ring + (size_t)ra.slot * A_SLOT_BYTES;
#pragma unroll
                    for (int r = 0; r < GEN_ROWS; ++r) {
                        uint8_t* sr = st + r * (ROW_STEP * 128);          // 128 bytes per row inside a tile
                        *reinterpret_cast<uint4*>(sr + off_cos) = make_uint4(w[r][0], w[r][1], w[r][2], w[r][3]);
                        *reinterpret_cast<uint4*>(sr + A_TILE_BYTES + off_cos) = make_uint4(w[r][4], w[r][5], w[r][6], w[r][7]);
                        *reinterpret_cast<uint4*>(sr + off_sin) = make_uint4(w[r][8], w[r][9], w[r][10], w[r][11]);
                        *reinterpret_cast<uint4*>(sr + A_TILE_BYTES + off_sin) = make_uint4(w[r][12], w[r][13], w[r][14], w[r][15]);
                    }
                    if (pr == 0 && kb == n_kb - 1) {
                        float* mpart = mu_part + (tile_iter & (MU_DEPTH - 1)) * 4 * BM + quarter * BM;
#pragma unroll
                        for (int r = 0; r < GEN_ROWS; ++r) mpart[r0 + r * ROW_STEP] = mu0[r];
                    }
                    pending = true; pending_bar = a_full(ra.slot);
                    ra.advance(NA);
                }
        }
        if (pending) { fence_proxy_async(); __syncwarp(); if (lane == 0) mbar_arrive(pending_bar); }
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
struct MapSet {                           // tensor maps of one operand pair (hi, lo) and the geometry they were built for
    CUtensorMap hi, lo;
    const void* bh = nullptr; const void* bl = nullptr;
    int KC_pad = 0, rows = 0, BN = 0;
};
struct UmmaState {
    MapSet full, enc;
    int max_smem = 0;
    bool attr_set[2][2][MAX_N + 1] = {};  // [traj][p32][n]
    unsigned long long* xq = nullptr; size_t xq_cap = 0;      // scratch of the trajectory mode
};

size_t table_bytes(const sspbo_ctx* ctx) {
    return (size_t)(ctx->KC_pad / 2) * ctx->n * sizeof(long long) + (size_t)ctx->KC_pad * sizeof(float) +
           mu_depth_for(ctx->KC_pad / BK) * 4 * BM * sizeof(float) + (2 * NA_MAX + 2 * MAX_NB + 4) * 8 + 16;
}

typedef void (*kernel_fn)(const CUtensorMap, const CUtensorMap, const Params);
constexpr int MAX_TRAJ_N = 3;           // waypoint dimension of the trajectory instantiations
kernel_fn kernel_for(int n, bool traj, bool p32) {
    if (traj) {
        switch (n) {
            case 1: return k_score_umma<1, true, false>; case 2: return k_score_umma<2, true, false>;
            case 3: return k_score_umma<3, true, false>;
        }
        return nullptr;
    }
    if (p32) {
        switch (n) {
            case 1: return k_score_umma<1, false, true>; case 2: return k_score_umma<2, false, true>;
            case 3: return k_score_umma<3, false, true>; case 4: return k_score_umma<4, false, true>;
            case 5: return k_score_umma<5, false, true>; case 6: return k_score_umma<6, false, true>;
            case 7: return k_score_umma<7, false, true>; case 8: return k_score_umma<8, false, true>;
        }
        return nullptr;
    }
    switch (n) {
        case 1: return k_score_umma<1, false, false>; case 2: return k_score_umma<2, false, false>;
        case 3: return k_score_umma<3, false, false>; case 4: return k_score_umma<4, false, false>;
        case 5: return k_score_umma<5, false, false>; case 6: return k_score_umma<6, false, false>;
        case 7: return k_score_umma<7, false, false>; case 8: return k_score_umma<8, false, false>;
    }
    return nullptr;
}

// candidates -> Q31.32 two's complement (trajectory mode reads each waypoint once per k-block)
template <typename XT>
__global__ void k_to_q32(const XT* __restrict__ x, long long count, unsigned long long* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        out[i] = (unsigned long long)__double2ll_rn((double)x[i] * 4294967296.0);
}

// K-major fp16 operand [rows x KC_pad], box = 64 columns x BN rows, 128-byte swizzle
int make_map(sspbo_ctx* ctx, CUtensorMap* map, const void* base, int KC_pad, int rows, int BN) {
    cuuint64_t gdim[2] = {(cuuint64_t)KC_pad, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)KC_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
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

int ensure_maps(sspbo_ctx* ctx, MapSet* ms, const void* bh, const void* bl, int KC_pad, int rows, int BN) {
    if (ms->bh == bh && ms->bl == bl && ms->KC_pad == KC_pad && ms->rows == rows && ms->BN == BN) return SSPBO_OK;
    int rc;
    if ((rc = make_map(ctx, &ms->hi, bh, KC_pad, rows, BN)) || (rc = make_map(ctx, &ms->lo, bl, KC_pad, rows, BN))) return rc;
    ms->bh = bh; ms->bl = bl; ms->KC_pad = KC_pad; ms->rows = rows; ms->BN = BN;
    return SSPBO_OK;
}

int launch(sspbo_ctx* ctx, float* enc_out, int enc_pitch, const void* x, int x_dtype, int64_t N, int64_t index0, float lam, float gamma_t,
           float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st) {
    UmmaState* us = (UmmaState*)ctx->umma;
    if (!us) { us = new UmmaState(); ctx->umma = us; }
    if (!us->max_smem) cudaDeviceGetAttribute(&us->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    Params p;
    memset(&p, 0, sizeof(p));
    MapSet* ms;
    p.BN = ctx->BN; p.n_chunks = ctx->n_chunks;
    p.stats = ctx->stats;
    if (enc_out) {                                       // B operand = inverse-DFT basis (dense: every k-block, every row)
        for (int i = 0; i < 8; ++i) p.kb_start[i] = 0;
        for (int i = 0; i < 32; ++i) p.nz_rows[i] = 32767;
        ms = &us->enc;
        int rc = ensure_maps(ctx, ms, ctx->Bh_enc, ctx->Bl_enc, ctx->KC_pad, ctx->NJ_pad, ctx->BN);
        if (rc) return rc;
        p.enc_out = enc_out; p.enc_pitch = enc_pitch; p.enc_inv_scale = (float)(1.0 / ctx->enc_scale);
    } else {
        for (int i = 0; i < 8; ++i) p.kb_start[i] = ctx->kb_start[i];
        for (int i = 0; i < 32; ++i) p.nz_rows[i] = ctx->nz_rows[i];
        ms = &us->full;
        int rc = ensure_maps(ctx, ms, ctx->Rh, ctx->Rl, ctx->KC_pad, ctx->NJ_pad, ctx->BN);
        if (rc) return rc;
    }
    const size_t tb = table_bytes(ctx), sb = 2 * (size_t)p.BN * BK * 2;
    const size_t avail = (size_t)us->max_smem - 1024 - tb;
    // a deep A ring first (it decouples the generator warps from each other and from the MMA latency), three B
    // slots cover the TMA latency, the rest goes to B
    int na = NA_MAX, nb = 3;
    while (na > NA_MIN && (size_t)na * A_SLOT_BYTES + (size_t)nb * sb > avail) --na;
    while (nb > 2 && (size_t)na * A_SLOT_BYTES + (size_t)nb * sb > avail) --nb;
    if ((size_t)na * A_SLOT_BYTES + (size_t)nb * sb > avail)
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "tcgen05 engine: shared memory does not fit two B slots");
    while (nb < MAX_NB && (size_t)na * A_SLOT_BYTES + (size_t)(nb + 1) * sb <= avail) ++nb;
    const size_t smem = 1024 + (size_t)na * A_SLOT_BYTES + (size_t)nb * sb + tb;
    const bool traj = ctx->L != 1;
    const bool p32 = !traj && ctx->p32_ok && !ctx->p32_disabled && ctx->stats;
    kernel_fn kern = kernel_for(ctx->n, traj, p32);
    if (!kern) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "tcgen05 engine: domain dimension %d not instantiated", ctx->n);
    if (!us->attr_set[traj][p32][ctx->n]) {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, us->max_smem));
        us->attr_set[traj][p32][ctx->n] = true;
    }
    p.x = x; p.x_f64 = (x_dtype == SSPBO_F64); p.N = N; p.index0 = index0;
    p.Aq_op = ctx->Aq_op; p.mp_op = ctx->mp_op;
    p.A32_op = ctx->A32_op; p.x_scale = p32 ? ldexp(1.0, ctx->p32_fx) : 0.0;
    p.K = ctx->K; p.KC_pad = ctx->KC_pad;
    p.n_kb = ctx->KC_pad / BK; p.nb = nb; p.na = na; p.mu_depth = mu_depth_for(p.n_kb);
    p.inv_beta = (float)(1.0 / ctx->beta); p.lam = lam; p.gamma_t = gamma_t;
    p.inv_rscale2 = (float)(1.0 / (ctx->r_scale * ctx->r_scale));
    p.mu = mu; p.var = var; p.acq = acq; p.best = best;
    p.xq = nullptr; p.tauq = (const unsigned long long*)ctx->tauq_op; p.L = ctx->L;
    p.debug = sspbo_debug_env();
    if (traj) {
        const size_t count = (size_t)N * ctx->L * ctx->n;
        if (count > us->xq_cap) {
            if (us->xq) { SSPBO_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(us->xq); us->xq = nullptr; us->xq_cap = 0; }
            SSPBO_CUDA(ctx, cudaMalloc((void**)&us->xq, count * sizeof(unsigned long long)));
            us->xq_cap = count;
        }
        const int cb = (int)((count + 255) / 256 < (size_t)ctx->sm_count * 16 ? (count + 255) / 256 : (size_t)ctx->sm_count * 16);
        if (x_dtype == SSPBO_F64) k_to_q32<double><<<cb, 256, 0, st>>>((const double*)x, (long long)count, us->xq);
        else k_to_q32<float><<<cb, 256, 0, st>>>((const float*)x, (long long)count, us->xq);
        SSPBO_LAUNCH_CHECK(ctx);
        p.xq = us->xq;
    }
    const long long tiles = (N + BM - 1) / BM;
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    kern<<<grid, THREADS, smem, st>>>(ms->hi, ms->lo, p);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

}  // namespace

int sspbo_make_map_f16(sspbo_ctx* ctx, void* map, const void* base, int KC_pad, int rows, int BN) {
    return make_map(ctx, (CUtensorMap*)map, base, KC_pad, rows, BN);
}

int sspbo_score_umma_supported(const sspbo_ctx* ctx) {
    if (!ctx->has_factor || ctx->n > MAX_N || ctx->K < 1 || !ctx->umma_range_ok || ctx->n_chunks > 8) return 0;
    if (ctx->L != 1 && (ctx->n > MAX_TRAJ_N || ctx->L > 64 || !ctx->tauq_op)) return 0;
    if (ctx->normalize || ctx->n_cls != 1) return 0;   // |phi| is not formed by the dense-factor kernel; one frequency table
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, ctx->device) != cudaSuccess || major != 10) return 0;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    if (table_bytes(ctx) + NA_MIN * A_SLOT_BYTES + 2 * (2 * (size_t)ctx->BN * BK * 2) + 1024 > (size_t)max_smem) return 0;
    return 1;
}

void sspbo_umma_release(sspbo_ctx* ctx) {
    if (ctx->umma) {
        UmmaState* us = (UmmaState*)ctx->umma;
        if (us->xq) cudaFree(us->xq);
        delete us; ctx->umma = nullptr;
    }
}

int sspbo_score_umma_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, int64_t index0, float lam, float gamma_t,
                            float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st) {
    return launch(ctx, nullptr, 0, x, x_dtype, N, index0, lam, gamma_t, mu, var, acq, best, st);
}

// K1 on the tensor cores: phi = psi B^T with the split-fp16 inverse-DFT basis as the B operand (float32 out, pitch = NJ_pad)
int sspbo_encode_umma_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, float* out, int pitch, cudaStream_t st) {
    return launch(ctx, out, pitch, x, x_dtype, N, 0, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, st);
}
