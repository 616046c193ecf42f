// K2, low-rank tcgen05 engine: fused encode + BLR acquisition + argmax while the posterior has rank r <= 255.
//
//   Sp = Sp0 - V^T V   (one row of V per observation, sspbo_core.cu:k_lr_append), psi^T Sp0 psi = |phi|^2 / alpha = 1 / alpha
//   var = 1/beta + 1/alpha - |V psi|^2,   mu = m' . psi
//
// so the contraction is 128 candidates x 1024 (psi) x r instead of x 1023: at r = 64 sixteen times fewer tensor flops
// than the dense quadratic form, and the kernel is bound by generating psi.  Everything about psi is therefore kept
// off shared memory: the generator warps write the split-fp16 psi k-block (A operand) straight into TENSOR MEMORY
// with tcgen05.st, and tcgen05.mma reads A from TMEM and only the V slab (B operand, TMA-fed) from shared memory.
//
// Per CTA (persistent, one per SM), per tile of 128 candidates:
//   generator warps (8)  thread <-> candidate row (TMEM lane = row, the warp's lane quadrant is warp % 4), 16 of the 32
//                        frequencies of a k-block each.  theta_f = A_f . x in fixed point (32-bit operands with
//                        fa + fx = 55 -> one IMAD.WIDE per axis, or Q31.32 x Q31.32), the phase bits become the
//                        mantissa of a float in [1, 2), MUFU sin/cos, split into fp16 hi/lo, tcgen05.st into the
//                        A ring (4-6 slots x 64 columns: 32 hi + 32 lo).
//   TMA warp             streams V_hi / V_lo slabs (64 columns x BN rows) through the B ring.
//   MMA warp             per k-block and chunk: 4 x 3 tcgen05.mma (M=128, N=BN, K=16): psi_hi V_hi^T + psi_hi V_lo^T +
//                        psi_lo V_hi^T into a float32 accumulator in TMEM.
//   (row 0 of the B operand is m' with its own scale, so output column 0 is mu = m' . psi: no FMA in the generator)
//   epilogue warps (4)   q = sum_j Y_j^2, var = 1/beta + 1/alpha - q, acquisition, packed-key argmax; candidates with
//                        var < SSPBO_LR_FLAG_FRACTION / alpha are appended to the flag list instead (the exact float64
//                        engine re-scores them right after this kernel, sspbo_core.cu:k_exact_*).
// TMEM map (512 columns): accumulator buffer(s) first ([0,128) / [0,256)), the A ring behind them.
// Rank: r_eff = r + 1 rows (row 0 = m'), one chunk of BN = ceil16(r_eff) <= 256 rows.  BN <= 128: the B slot's hi and lo
// tiles are read through ONE descriptor as a 2 BN-row operand (two MMAs per K = 16 step instead of three; the accumulator
// is 2 BN wide and the epilogue adds its halves); two accumulator buffers alternate by tile while BN <= 64, a single
// 256-column accumulator beyond (the epilogue then sits between two tiles).
//
// Replaces SSPAgent.eval (agents/ssp_agent.py:125-137) + np.argmax (experiments/demo_blr_gif.py:129-136).
#include "sspbo_lr_common.cuh"
#include <stdlib.h>

namespace {
using namespace sspbo_tc;
using namespace sspbo_lr;

// PH: phase arithmetic of the generator: 0 = Q31.32 x Q31.32 (any range), 1 = 32-bit fixed point (fa + fx = 55),
// 2 = float32 FMA chain in radians (|theta| <= 32 pi by the declared bounds: rounding < 2e-5 rad)
// TRAJ: a candidate is a row of L waypoints; psi sums L unit phasors per frequency (agents/ssp_traj_agent.py:145-163 in
// the Fourier domain) and |phi|^2 = psi^T D psi is no longer 1: the generators accumulate it for the epilogue.
// GROUPED: one launch scores the shared candidate set for `n_runs` independent runs (RunScore, one per run): CTA b takes all
// tiles of runs b, b + gridDim.x, ...; the pipeline state (rings, barrier phases, accumulators) simply runs on from one run
// into the next.  The per-run launches of config C4 cost ~5 us of host time each, several times what the kernel needs.
template <int NDIM, int PH, bool TRAJ, bool GROUPED = false>
__global__ void __launch_bounds__(THREADS, 1)
k_score_lr(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const Params p,
           const RunScore* __restrict__ runs, const int n_runs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform: role branches and the MMA issuer loop stay on the uniform datapath
    const int NA = p.na;
    const uint32_t ACC_COLS = (uint32_t)p.acc_cols, A_RING_COL0 = (uint32_t)(p.n_bufs * p.acc_cols);
    const int b_tile_bytes = p.BN * BK * 2;
    const int b_slot_bytes = 2 * b_tile_bytes;
    uint8_t* b_ring = smem;
    uint8_t* tables = b_ring + (size_t)p.nb * b_slot_bytes;
    long long* Aq_s = (long long*)tables;                                   // (KC_pad/2) x NDIM (int32 in the P32 path)
    // |phi|^2 partial sums of the tiles in flight, [NRM_DEPTH tiles][GEN_W warps of a quadrant][128 rows] (trajectory mode)
    float* nrm_part = (float*)(Aq_s + (size_t)p.n_cls * (p.KC_pad / 2) * NDIM);
    uint64_t* bars = (uint64_t*)(nrm_part + (TRAJ ? NRM_DEPTH * GEN_W * BM : 0));
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
    if constexpr (!GROUPED) load_tables<NDIM, PH>(p, tables, THREADS);
    if constexpr (TRAJ) {
        for (int i = threadIdx.x; i < NRM_DEPTH * GEN_W * BM; i += THREADS) nrm_part[i] = 0.f;
    }
    if (warp == WARP_TMA && lane == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(a_full(s), 4); mbar_init(a_empty(s), 1); }
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
    const int n_kb = p.n_kb, n_chunks = p.n_chunks;
    // this CTA's tile sequence: tiles blockIdx.x, blockIdx.x + gridDim.x, ... of the one candidate set, or every tile of its runs
    const long long my_units = GROUPED ? ((n_runs > (int)blockIdx.x) ? (n_runs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0)
                                       : ((n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
    const long long my_tiles_all = GROUPED ? my_units * n_tiles : my_units;
    auto run_of = [&](long long tile_iter) { return GROUPED ? (int)(blockIdx.x + (tile_iter / n_tiles) * gridDim.x) : 0; };
    auto tile_of = [&](long long tile_iter) { return GROUPED ? tile_iter % n_tiles : (long long)blockIdx.x + tile_iter * gridDim.x; };
    // accumulator of chunk c: with a single chunk per tile the two buffers alternate by tile instead
    auto acc_of = [&](int tile_iter, int c) { return n_chunks == 1 ? (tile_iter & (p.n_bufs - 1)) : c; };

    if (warp == WARP_TMA) {
        // ================================ TMA producer (B ring) ================================
        if (lane == 0) {
            Ring rb;
            for (long long ti = 0; ti < my_tiles_all; ++ti) {
                const CUtensorMap* mh = GROUPED ? &runs[run_of(ti)].map_hi : &map_hi;
                const CUtensorMap* ml = GROUPED ? &runs[run_of(ti)].map_lo : &map_lo;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait_relaxed<64>(b_empty(rb.slot), rb.phase ^ 1);
                    uint8_t* st = b_ring + (size_t)rb.slot * b_slot_bytes;
                    mbar_expect_tx(b_full(rb.slot), (uint32_t)b_slot_bytes);
                    tma_load_2d(smem_u32(st), mh, b_full(rb.slot), kb * BK, 0);
                    tma_load_2d(smem_u32(st + b_tile_bytes), ml, b_full(rb.slot), kb * BK, 0);
                    rb.advance(p.nb);
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ================================ MMA issuer ================================
        // The whole warp walks the pipeline (warp-uniform control flow keeps ring state and descriptors in uniform
        // registers); one elected lane issues.  This loop is the consumer side of every k-block and runs on ONE warp,
        // so it is kept short: one wait per ring, one fence, one elected section with the MMAs and all commits.
        Ring ra, rb;
        uint32_t acc_phase[2] = {0, 0};
        const uint32_t idesc = make_idesc(BM, p.BN), idesc2 = make_idesc(BM, 2 * p.BN);
        const bool no_mma = (SSPBO_DBG_BITS(p) & 1) != 0;
        const uint64_t b_desc0 = make_desc_sw128(smem_u32(b_ring));       // slot s: start address + s * slot bytes (16 B units)
        const uint32_t b_slot_units = (uint32_t)(b_slot_bytes >> 4), b_lo_units = (uint32_t)(b_tile_bytes >> 4);
        const uint32_t a_ring0 = tmem_base + A_RING_COL0;
        const bool merged = p.merged != 0;
        if (p.split_issue) {
            // Software pipeline: the waits for k-block i + 1 sit BETWEEN the two halves of the MMAs of k-block i, so the
            // tensor pipe always has queued work while this warp polls barriers (its instruction queue is only a few
            // MMAs deep: issuing a whole k-block and then waiting would let it run dry).
            const long long total_blocks = my_tiles_all * n_kb;
            if (total_blocks > 0) {
                mbar_wait_relaxed<40>(a_full(ra.slot), ra.phase);
                mbar_wait(b_full(rb.slot), rb.phase);
            }
            int tile_iter = 0, kb = 0;
            for (long long g = 0; g < total_blocks; ++g) {
                const int buf = tile_iter & (p.n_bufs - 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * ACC_COLS;
                if (kb == 0) {
                    mbar_wait(tmem_empty(buf), acc_phase[buf] ^ 1);           // the epilogue has drained this accumulator
                    acc_phase[buf] ^= 1;
                }
                tc_fence_after();
                const uint32_t a_hi = a_ring0 + (uint32_t)(ra.slot * A_SLOT_COLS);
                const uint64_t b_hi = b_desc0 + (uint64_t)((uint32_t)rb.slot * b_slot_units);
                const uint64_t b_lo = b_hi + b_lo_units;
                const uint32_t a_bar = a_empty(ra.slot), b_bar = b_empty(rb.slot);
                // 16 fp16 along K = 32 bytes in shared memory (2 descriptor units of 16 B) = 8 TMEM columns.  The first MMA
                // of a tile overwrites the accumulator.  Merged: the lo tile follows the hi tile in the B slot, so ONE
                // descriptor with N = 2 BN covers [V_hi; V_lo]: accumulator columns [0, BN) get psi_hi V_hi^T, [BN, 2 BN)
                // psi_hi V_lo^T; a second MMA adds psi_lo V_hi^T to the first half; the epilogue adds the halves.
                if (elect_one() && !no_mma) {
                    if (merged) {
                        umma_f16_ts(d_tmem, a_hi, b_hi, idesc2, kb != 0);
                        umma_f16_ts_acc(d_tmem, a_hi + 32, b_hi, idesc);
                        umma_f16_ts_acc(d_tmem, a_hi + 8, b_hi + 2, idesc2);
                        umma_f16_ts_acc(d_tmem, a_hi + 32 + 8, b_hi + 2, idesc);
                    } else {
                        umma_f16_ts(d_tmem, a_hi, b_hi, idesc, kb != 0);
                        umma_f16_ts_acc(d_tmem, a_hi, b_lo, idesc);
                        umma_f16_ts_acc(d_tmem, a_hi + 32, b_hi, idesc);
                        umma_f16_ts_acc(d_tmem, a_hi + 8, b_hi + 2, idesc);
                        umma_f16_ts_acc(d_tmem, a_hi + 8, b_lo + 2, idesc);
                        umma_f16_ts_acc(d_tmem, a_hi + 32 + 8, b_hi + 2, idesc);
                    }
                }
                __syncwarp();
                rb.advance(p.nb);
                ra.advance(NA);
                if (g + 1 < total_blocks) {                                   // next k-block's operands, while the pipe works
                    mbar_wait_relaxed<40>(a_full(ra.slot), ra.phase);
                    mbar_wait(b_full(rb.slot), rb.phase);
                }
                if (elect_one()) {
                    if (!no_mma) {
                        if (merged) {
                            umma_f16_ts_acc(d_tmem, a_hi + 16, b_hi + 4, idesc2);
                            umma_f16_ts_acc(d_tmem, a_hi + 32 + 16, b_hi + 4, idesc);
                            umma_f16_ts_acc(d_tmem, a_hi + 24, b_hi + 6, idesc2);
                            umma_f16_ts_acc(d_tmem, a_hi + 32 + 24, b_hi + 6, idesc);
                        } else {
                            umma_f16_ts_acc(d_tmem, a_hi + 16, b_hi + 4, idesc);
                            umma_f16_ts_acc(d_tmem, a_hi + 16, b_lo + 4, idesc);
                            umma_f16_ts_acc(d_tmem, a_hi + 32 + 16, b_hi + 4, idesc);
                            umma_f16_ts_acc(d_tmem, a_hi + 24, b_hi + 6, idesc);
                            umma_f16_ts_acc(d_tmem, a_hi + 24, b_lo + 6, idesc);
                            umma_f16_ts_acc(d_tmem, a_hi + 32 + 24, b_hi + 6, idesc);
                        }
                    }
                    umma_commit(b_bar);                                       // frees the B slot when these MMAs retire
                    umma_commit(a_bar);                                       // ... and the A slot
                    if (kb == n_kb - 1) umma_commit(tmem_full(buf));
                }
                __syncwarp();
                if (++kb == n_kb) { kb = 0; ++tile_iter; }
            }
        } else {
            // short MMAs (BN <= 128): one elected section per k-block keeps the loop at a minimum of instructions
            for (int tile_iter = 0; tile_iter < (int)my_tiles_all; ++tile_iter) {
                const int buf = tile_iter & (p.n_bufs - 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * ACC_COLS;
                mbar_wait_inline(tmem_empty(buf), acc_phase[buf] ^ 1);        // the epilogue has drained this accumulator
                acc_phase[buf] ^= 1;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait_inline(a_full(ra.slot), ra.phase);
                    mbar_wait_inline(b_full(rb.slot), rb.phase);
                    tc_fence_after();
                    const uint32_t a_hi = a_ring0 + (uint32_t)(ra.slot * A_SLOT_COLS);
                    const uint64_t b_hi = b_desc0 + (uint64_t)((uint32_t)rb.slot * b_slot_units);
                    if (elect_one()) {
                        if (!no_mma) {
                            // 16 fp16 along K = 32 bytes in shared memory (2 descriptor units of 16 B) = 8 TMEM columns.
                            // The first MMA of a tile overwrites the accumulator.
                            if (merged) {
                                // The lo tile follows the hi tile in the B slot, so ONE descriptor with N = 2 BN covers
                                // [V_hi; V_lo]: accumulator columns [0, BN) get psi_hi V_hi^T, [BN, 2 BN) psi_hi V_lo^T; a
                                // second MMA adds psi_lo V_hi^T to the first half.  The epilogue adds the halves.
                                umma_f16_ts(d_tmem, a_hi, b_hi, idesc2, kb != 0);
                                umma_f16_ts_acc(d_tmem, a_hi + 32, b_hi, idesc);
    #pragma unroll
                                for (int kk = 1; kk < BK / 16; ++kk) {
                                    umma_f16_ts_acc(d_tmem, a_hi + 8 * kk, b_hi + 2 * kk, idesc2);
                                    umma_f16_ts_acc(d_tmem, a_hi + 32 + 8 * kk, b_hi + 2 * kk, idesc);
                                }
                            } else {
                                const uint64_t b_lo = b_hi + b_lo_units;
                                umma_f16_ts(d_tmem, a_hi, b_hi, idesc, kb != 0);
                                umma_f16_ts_acc(d_tmem, a_hi, b_lo, idesc);
                                umma_f16_ts_acc(d_tmem, a_hi + 32, b_hi, idesc);
    #pragma unroll
                                for (int kk = 1; kk < BK / 16; ++kk) {
                                    umma_f16_ts_acc(d_tmem, a_hi + 8 * kk, b_hi + 2 * kk, idesc);
                                    umma_f16_ts_acc(d_tmem, a_hi + 8 * kk, b_lo + 2 * kk, idesc);
                                    umma_f16_ts_acc(d_tmem, a_hi + 32 + 8 * kk, b_hi + 2 * kk, idesc);
                                }
                            }
                        }
                        umma_commit(b_empty(rb.slot));                        // frees the B slot when these MMAs retire
                        umma_commit(a_empty(ra.slot));                        // ... and the A slot
                        if (kb == n_kb - 1) umma_commit(tmem_full(buf));
                    }
                    __syncwarp();
                    rb.advance(p.nb);
                    ra.advance(NA);
                }
            }
        }
    } else if (warp < WARP_GEN0) {
        // ================================ epilogue ================================
        const int quad = warp & 3;                                        // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        uint32_t acc_phase[2] = {0, 0};
        unsigned long long my_best = 0ull;
        float inv_mscale = GROUPED ? 0.f : __ldg(p.mscale);
        // what the epilogue takes from the run (GROUPED) or from the launch parameters
        float e_inv_beta = p.inv_beta, e_lam = p.lam, e_gamma = p.gamma_t, e_inv_rs2 = p.inv_rscale2, e_prior = p.prior_var,
              e_flag_below = p.flag_below;
        unsigned long long* e_best = p.best; unsigned int* e_flag_idx = p.flag_idx; unsigned long long* e_stats = p.stats;
        int cur_run = -1;
        for (int tile_iter = 0; tile_iter < (int)my_tiles_all; ++tile_iter) {
            const long long tile = tile_of(tile_iter);
            if constexpr (GROUPED) {
                const int run = run_of(tile_iter);
                if (run != cur_run) {                                       // first tile of a run: publish the last run's key, load this run's scalars
                    if (cur_run >= 0) {
                        for (int o = 16; o > 0; o >>= 1) {
                            unsigned long long other = __shfl_xor_sync(0xffffffffu, my_best, o);
                            if (other > my_best) my_best = other;
                        }
                        if (lane == 0 && my_best && e_best) atomicMax(e_best, my_best);
                        my_best = 0ull;
                    }
                    cur_run = run;
                    const RunScore& rs = runs[run];
                    inv_mscale = __ldg(rs.mscale);
                    e_inv_beta = rs.inv_beta; e_lam = rs.lam; e_gamma = rs.gamma_t; e_inv_rs2 = rs.inv_rscale2;
                    e_prior = rs.prior_var; e_flag_below = rs.flag_below;
                    e_best = rs.best; e_flag_idx = rs.flag_idx; e_stats = rs.stats;
                }
            }
            float q = 0.f, mu_raw = 0.f;
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = acc_of(tile_iter, c);
                mbar_wait_relaxed<100>(tmem_full(buf), acc_phase[buf]);
                acc_phase[buf] ^= 1;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * ACC_COLS;
                if (p.merged) {
                    // columns [0, BN): psi_hi V_hi^T + psi_lo V_hi^T, columns [BN, 2 BN): psi_hi V_lo^T
                    for (int g = 0; g < p.BN; g += 16) {                  // BN % 16 == 0
                        uint32_t va[16], vb[16];
                        tmem_ld16(taddr + (uint32_t)g, va);
                        tmem_ld16(taddr + (uint32_t)(p.BN + g), vb);
                        tmem_ld_wait();
                        float q0 = 0.f, q1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            float y0 = __uint_as_float(va[i]) + __uint_as_float(vb[i]);
                            const float y1 = __uint_as_float(va[i + 1]) + __uint_as_float(vb[i + 1]);
                            if (i == 0 && g == 0) { mu_raw = y0; y0 = 0.f; }       // output row 0 is m' . psi, not a row of V
                            q0 = fmaf(y0, y0, q0); q1 = fmaf(y1, y1, q1);
                        }
                        q += q0 + q1;
                    }
                } else
                for (int g = 0; g < p.BN; g += 64) {                      // BN % 16 == 0; up to 64 columns in flight
                    uint32_t v[64];
                    const int nld = min(4, (p.BN - g) >> 4);
#pragma unroll
                    for (int l = 0; l < 4; ++l)
                        if (l < nld) tmem_ld16(taddr + (uint32_t)(g + 16 * l), v + 16 * l);
                    tmem_ld_wait();
                    if (c == 0 && g == 0) { mu_raw = __uint_as_float(v[0]); v[0] = 0u; }   // output row 0 is m' . psi, not a row of V
                    // sum of squares on packed pairs (FFMA2): the drain of a 256-column accumulator holds up the next tile's MMAs,
                    // and these warps share their schedulers with the generators
                    f32x2 q01 = 0ull, q23 = 0ull;
#pragma unroll
                    for (int l = 0; l < 4; ++l)
                        if (l < nld) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const f32x2 y01 = f2_pack(__uint_as_float(v[16 * l + i]), __uint_as_float(v[16 * l + i + 1]));
                                const f32x2 y23 = f2_pack(__uint_as_float(v[16 * l + i + 2]), __uint_as_float(v[16 * l + i + 3]));
                                q01 = f2_fma(y01, y01, q01); q23 = f2_fma(y23, y23, q23);
                            }
                        }
                    float q0, q1, q2, q3;
                    f2_unpack(q01, q0, q1); f2_unpack(q23, q2, q3);
                    q += (q0 + q1) + (q2 + q3);
                }
                tc_fence_before();
                mbar_arrive(tmem_empty(buf));
            }
            const long long gi = tile * BM + row;
            if (gi < p.N) {
                float mu = mu_raw * inv_mscale;
                float prior = e_prior, flag_below = e_flag_below;
                if constexpr (TRAJ) {                                                 // |phi|^2 / alpha instead of 1 / alpha
                    float* np_ = nrm_part + (size_t)(tile_iter & (NRM_DEPTH - 1)) * GEN_W * BM + row;
                    float ssum = 0.f;
#pragma unroll
                    for (int w = 0; w < GEN_W; ++w) { ssum += np_[w * BM]; np_[w * BM] = 0.f; }   // warps without a k-block in a tile write nothing
                    const float nrm2 = fmaf(p.nrm_scale, ssum, -p.nrm_offset);
                    if (p.normalize) { const float inv = 1.f / nrm2; mu *= sqrtf(inv); q *= inv; }   // phi / |phi|
                    else { prior *= nrm2; flag_below *= nrm2; }
                }
                const float rem = prior - q * e_inv_rs2;                              // |phi|^2 / alpha - |V psi|^2
                const float var = e_inv_beta + rem;                                   // blr.py:96
                bool flagged = false;
                if (var < flag_below) {                                             // too close to the observations for
                    const unsigned long long slot = atomicAdd(e_stats + SSPBO_STAT_CUR, 1ull);   // the cancelling form
                    if (slot < (unsigned long long)p.flag_cap) {
                        e_flag_idx[slot] = (unsigned int)gi; flagged = true;
                        e_flag_idx[p.flag_cap + slot] = __float_as_uint(flagged_score_bound(mu, var, prior, e_lam, e_gamma));
                    }
                    else atomicAdd(e_stats + SSPBO_STAT_OVERFLOW, 1ull);
                }
                if (!flagged) {
                    const float acq = e_lam * var / (sqrtf(var + e_gamma) + sqrtf(e_gamma));     // ssp_agent.py:136, cancellation-free
                    const float score = mu + acq;
                    if (p.mu) p.mu[gi] = mu;
                    if (p.var) p.var[gi] = var;
                    if (p.acq) p.acq[gi] = acq;
                    const unsigned long long key = sspbo_pack(score, (unsigned long long)(p.index0 + gi));
                    if (key > my_best) my_best = key;
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, my_best, o);
            if (other > my_best) my_best = other;
        }
        if (lane == 0 && my_best && e_best) atomicMax(e_best, my_best);
    } else {
        // ================================ psi generators (A ring in tensor memory) ================================
        // GEN_W warps per TMEM lane quadrant (thread <-> candidate row).  Warp j of a quadrant produces the k-blocks
        // g = j, j + GEN_W, ... of the CTA's k-block sequence (tiles x n_kb), both frequency halves of each, so GEN_W
        // consecutive k-blocks are in production at once and every scheduler has GEN_W generator warps to hide the
        // fixed-latency stalls of each other (the instruction stream of one warp issues only every ~4 cycles).
        const int gw = warp - WARP_GEN0;
        const int quad = warp & 3, member = gw >> 2;                      // TMEM lane quadrant, position in the rotation
        const int r0 = quad * 32 + lane;                                  // candidate row of this thread = TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + A_RING_COL0;
        const long long N = p.N;
        const long long total_blocks = my_tiles_all * n_kb;
        const uint8_t* gen_tab = tables;                                  // GROUPED: the run's table in global memory
        unsigned long long* oor_stats = p.stats;
        int slot = member % NA;
        uint32_t phase = (uint32_t)((member / NA) & 1);
        int kb = member % n_kb;
        long long tile_iter = member / n_kb, loaded_iter = -1;
        unsigned long long x0[NDIM];                                      // Q31.32 candidate (two's complement)
        int xq0[NDIM];                                                    // Q.fx candidate of the 32-bit phase path
        float nrm_acc = 0.f;                                              // sum of C^2 + S^2 over this warp's k-blocks of the tile
        long long g0 = 0;
        bool row_nan = false;                                             // a coordinate of this row is NaN: its score must be NaN (np.argmax semantics)
        for (long long g = member; g < total_blocks; g += GEN_W) {
            if (tile_iter != loaded_iter) {                               // first k-block of a tile for this warp: its row
                loaded_iter = tile_iter;
                nrm_acc = 0.f;
                g0 = tile_of(tile_iter) * BM + r0;
                if constexpr (GROUPED) { const RunScore& rs = runs[run_of(tile_iter)]; gen_tab = (const uint8_t*)rs.tab; oor_stats = rs.stats; }
                if constexpr (!TRAJ) {
                const bool oor = load_candidate<NDIM, PH>(p, g0, x0, xq0, row_nan);
                if (oor && member == 0) atomicAdd(oor_stats + SSPBO_STAT_OOR, 1ull);  // outside the declared |x| bound
                const long long gn = GROUPED ? tile_of(tile_iter + 1) * BM + r0 : g0 + (long long)gridDim.x * BM;   // next tile's row: pull it into L1 now (with few k-blocks
                if (gn < N && p.cand.mode == 0) {                         // per tile every warp loads a row per k-block or two)
                    const char* nx = (const char*)p.cand.x + (size_t)gn * NDIM * (p.cand.x_f64 ? 8 : 4);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
                }
                }
            }
            const uint32_t empty_bar = a_empty(slot);
            bool slot_free = mbar_test_wait(empty_bar, phase ^ 1);        // early, non-blocking probe
            const uint32_t slot_addr = lane_addr + (uint32_t)(slot * A_SLOT_COLS);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float cs[NF], sn[NF];
                gen_half<NDIM, PH, TRAJ>(p, gen_tab, g0, kb, half, x0, xq0, cs, sn, nrm_acc);
                if (!TRAJ && row_nan) cs[0] = __int_as_float(0x7fc00000);      // poisons mu and |V psi|^2 of the row through the MMA
                uint32_t wch[NF / 2], wcl[NF / 2], wsh[NF / 2], wsl[NF / 2];     // cos hi / cos lo / sin hi / sin lo, fp16 pairs
#pragma unroll
                for (int i = 0; i < NF / 2; ++i) {
                    split_pair_trunc2(cs[2 * i], cs[2 * i + 1], wch[i], wcl[i]);
                    split_pair_trunc2(sn[2 * i], sn[2 * i + 1], wsh[i], wsl[i]);
                }
                if (!slot_free) { mbar_wait(empty_bar, phase ^ 1); slot_free = true; }
                tc_fence_after();
                // operand order inside a k-block: K index i < 32 is cos(f = i), i >= 32 is sin(f = i - 32); two fp16 per
                // TMEM column, hi part in columns [0, 32), lo part in [32, 64) of the slot
                const uint32_t sa = slot_addr + (uint32_t)(half * (NF / 2));
                tmem_st8(sa, wch);
                tmem_st8(sa + 16, wsh);
                tmem_st8(sa + 32, wcl);
                tmem_st8(sa + 48, wsl);
            }
            if constexpr (TRAJ) {                                         // this warp's last k-block of the tile: its share of |phi|^2
                if (kb + GEN_W >= n_kb)
                    nrm_part[(size_t)(tile_iter & (NRM_DEPTH - 1)) * GEN_W * BM + member * BM + r0] = nrm_acc;
            }
            tmem_st_wait();                                               // publish the k-block
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(slot));
            slot += GEN_W; while (slot >= NA) { slot -= NA; phase ^= 1; }
            kb += GEN_W; while (kb >= n_kb) { kb -= n_kb; ++tile_iter; }
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
struct LrState {
    CUtensorMap hi, lo;
    const void* bh = nullptr; const void* bl = nullptr;
    int KC_pad = 0, BN = 0;
    int max_smem = 0;
    bool attr_set[2][3][MAX_N + 1] = {};
    bool runs_attr_set[MAX_N + 1] = {};
};

size_t table_bytes(const sspbo_ctx* ctx) {
    return (size_t)ctx->n_cls * (ctx->KC_pad / 2) * ctx->n * sizeof(long long) + (ctx->L != 1 ? NRM_DEPTH * GEN_W * BM * sizeof(float) : 0) +
           (2 * NA_MAX + 2 * MAX_NB + 4) * 8 + 16;
}

typedef void (*kernel_fn)(const CUtensorMap, const CUtensorMap, const Params, const RunScore*, const int);
template <int PH>
kernel_fn kernel_for_n(int n) {
    switch (n) {
        case 1: return k_score_lr<1, PH, false>; case 2: return k_score_lr<2, PH, false>; case 3: return k_score_lr<3, PH, false>;
        case 4: return k_score_lr<4, PH, false>; case 5: return k_score_lr<5, PH, false>; case 6: return k_score_lr<6, PH, false>;
        case 7: return k_score_lr<7, PH, false>; case 8: return k_score_lr<8, PH, false>;
    }
    if constexpr (PH != 1) {                             // 9 .. 16 dimensions: float32 chain or Q31.32 only
        switch (n) {
            case 9: return k_score_lr<9, PH, false>; case 10: return k_score_lr<10, PH, false>;
            case 11: return k_score_lr<11, PH, false>; case 12: return k_score_lr<12, PH, false>;
            case 13: return k_score_lr<13, PH, false>; case 14: return k_score_lr<14, PH, false>;
            case 15: return k_score_lr<15, PH, false>; case 16: return k_score_lr<16, PH, false>;
        }
    }
    return nullptr;
}
template <int PH>
kernel_fn kernel_for_traj(int n) {
    switch (n) {
        case 1: return k_score_lr<1, PH, true>; case 2: return k_score_lr<2, PH, true>; case 3: return k_score_lr<3, PH, true>;
    }
    return nullptr;
}
kernel_fn kernel_for(int n, int ph, bool traj) {
    if (traj) return ph == 2 ? kernel_for_traj<2>(n) : kernel_for_traj<0>(n);
    return ph == 2 ? kernel_for_n<2>(n) : (ph == 1 ? kernel_for_n<1>(n) : kernel_for_n<0>(n));
}
// grouped scoring of many runs: float32 phase chain, up to 4 domain dimensions
kernel_fn kernel_for_runs(int n) {
    switch (n) {
        case 1: return k_score_lr<1, 2, false, true>; case 2: return k_score_lr<2, 2, false, true>;
        case 3: return k_score_lr<3, 2, false, true>; case 4: return k_score_lr<4, 2, false, true>;
    }
    return nullptr;
}

}  // namespace

// ranks this engine covers in its own operand format: one chunk of <= 256 rows
int sspbo_score_lr_supported(const sspbo_ctx* ctx) {
    if (ctx->L != 1 && (ctx->n > MAX_TRAJ_N || ctx->L > 64 || !ctx->tauq_op)) return 0;
    if (ctx->cc_major != 10 || !ctx->has_factor) return 0;
    return ctx->lr_valid && ctx->lr_rank >= 1 && ctx->lr_rank + 1 <= 2 * (int)ACC_COLS_MAX && ctx->n <= MAX_N &&
           ctx->umma_range_ok && ctx->stats;
}

void sspbo_lr_release(sspbo_ctx* ctx) {
    if (ctx->lr_state) { delete (LrState*)ctx->lr_state; ctx->lr_state = nullptr; }
}

// phase path: float32 FMA chain when the declared bounds keep |theta| small, 32-bit fixed point when they allow it,
// Q31.32 otherwise (or after candidates outside the declared bound were seen)
int sspbo_lr_phase_path(const sspbo_ctx* ctx) {
    int ph = 0;
    const bool traj = ctx->L != 1;
    if (!ctx->p32_disabled) ph = (ctx->pf32_ok && !ctx->pf32_disabled) ? 2 : (ctx->p32_ok ? 1 : 0);
    if ((traj || ctx->n > 8) && ph == 1) ph = 0;         // trajectories and n > 8: float32 chain or Q31.32
    if (sspbo_debug_env() & 8) ph = 0;                   // timing experiments only (compiled out of product builds)
    if (sspbo_debug_env() & 16) ph = ctx->p32_ok ? 1 : 0;
    return ph;
}

// the launch parameters that do not depend on the operand format
void sspbo_lr_fill_common(const sspbo_ctx* ctx, const CandDesc& cand, int64_t N, int64_t index0, float lam, float gamma_t,
                          float* mu, float* var, float* acq, unsigned long long* best, int ph, Params* pp) {
    Params& p = *pp;
    memset(&p, 0, sizeof(p));
    p.cand = cand; p.N = N; p.index0 = index0;
    p.Aq_op = ctx->Aq_op; p.A32_op = ctx->A32_op; p.Af_op = ctx->Af_op; p.mscale = ctx->mscale;
    p.x_scale = ph == 1 ? ldexp(1.0, ctx->p32_fx) : 0.0; p.x_absmax = (float)(ctx->x_absmax * 1.000001);
    p.KC_pad = ctx->KC_pad; p.n_kb = ctx->KC_pad / BK;
    p.n_chunks = 1; p.bpt = p.n_kb;
    p.inv_beta = (float)(1.0 / ctx->beta); p.lam = lam; p.gamma_t = gamma_t;
    p.inv_rscale2 = (float)(1.0 / (ctx->r_scale * ctx->r_scale));
    p.prior_var = (float)(1.0 / ctx->alpha);
    p.flag_below = (float)(SSPBO_LR_FLAG_FRACTION / ctx->alpha);
    p.mu = mu; p.var = var; p.acq = acq; p.best = best;
    p.flag_idx = ctx->flag_idx; p.stats = ctx->stats; p.flag_cap = SSPBO_FLAG_CAP;
    p.L = ctx->L; p.n_cls = ctx->n_cls; p.tauq = (const unsigned long long*)ctx->tauq_op; p.tauf = ctx->tauf_op;
    {   // |phi|^2 = (1/d) (psi_0^2 + 2 sum_{k>=1} (C_k^2 + S_k^2)); the generators sum C^2 + S^2 over every operand frequency,
        // including the DC term (= dc^2, weight 1 not 2) and the padding frequencies (= L^2 each, weight 0)
        const double d = (double)ctx->d, L2 = (double)ctx->L * ctx->L;
        const int n_pad = ctx->KC_pad / 2 - 1 - ctx->K;
        p.nrm_scale = (float)(2.0 / d); p.nrm_offset = (float)((ctx->dc * ctx->dc + 2.0 * n_pad * L2) / d);
        p.normalize = ctx->normalize ? 1 : 0;
    }
    p.debug = sspbo_debug_env();
}

// tensor maps of the context's operand images for chunks of bn rows (rebuilt when the images or bn changed)
static int lr_ensure_maps(sspbo_ctx* ctx, int bn, LrState** out) {
    LrState* ls = (LrState*)ctx->lr_state;
    if (!ls) { ls = new LrState(); ctx->lr_state = ls; }
    if (!ls->max_smem) cudaDeviceGetAttribute(&ls->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    if (ls->bh != ctx->Vh || ls->bl != ctx->Vl || ls->KC_pad != ctx->KC_pad || ls->BN != bn) {
        int rc;
        if ((rc = sspbo_make_map_f16(ctx, &ls->hi, ctx->Vh, ctx->KC_pad, SSPBO_LR_MAX + 1, bn)) ||
            (rc = sspbo_make_map_f16(ctx, &ls->lo, ctx->Vl, ctx->KC_pad, SSPBO_LR_MAX + 1, bn)))
            return rc;
        ls->bh = ctx->Vh; ls->bl = ctx->Vl; ls->KC_pad = ctx->KC_pad; ls->BN = bn;
    }
    *out = ls;
    return SSPBO_OK;
}

// layout of one launch for chunks of bn operand rows: B ring depth, shared memory, accumulators and A ring in tensor memory
static int lr_configure(sspbo_ctx* ctx, const LrState* ls, int bn, Params* p, size_t* smem_out) {
    const size_t tb = table_bytes(ctx), sb = 2 * (size_t)bn * BK * 2;
    int nb = (int)(((size_t)ls->max_smem - 1024 - tb) / sb);
    if (nb > MAX_NB) nb = MAX_NB;
    if (nb < 2) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "low-rank tcgen05 engine: shared memory does not fit two B slots");
    *smem_out = 1024 + (size_t)nb * sb + tb;
    p->BN = bn; p->nb = nb;
    // TMEM: accumulators first, the A ring behind them.  BN <= 128: three BN-wide MMAs per K step (hi hi, hi lo, lo hi) into
    // ONE BN-wide accumulator, two of which alternate by tile, so the epilogue of a tile overlaps the MMAs of the next one
    // and reads BN columns.  (Round 1 merged the hi|lo tiles into one 2 BN-row operand to issue two MMAs per K step instead
    // of three: the issue cost ~100 cycles per MMA then.  With the issuer's loop on the uniform datapath the issue is cheap,
    // the tensor-pipe time of the two layouts is the same, and the merged one pays a 2 BN-wide accumulator and an epilogue
    // that adds the halves: measured after the change, d = 1023 ranks 10..127: +1..3 % unmerged; d = 151 rank 70: 0.107 ->
    // 0.100 ms per 1e6 candidates.  SSPBO_LR_MERGED=1 brings the merged layout back for comparison.)
    // BN > 128: one 256-column accumulator, the waits of the next k-block between the two halves of the MMAs.
    p->merged = 0;
    p->split_issue = bn > 128;
    if (bn <= 64) { p->acc_cols = 64; p->n_bufs = 2; }
    else if (bn <= 128) { p->acc_cols = 128; p->n_bufs = 2; }
    else { p->acc_cols = 256; p->n_bufs = 1; }
    {
        static const int merged_env = getenv("SSPBO_LR_MERGED") ? atoi(getenv("SSPBO_LR_MERGED")) : -1;
        if (merged_env == 1 && bn <= 128) {
            p->merged = 1;
            if (bn <= 64) { p->acc_cols = 128; p->n_bufs = 2; }
            else if (4 * bn + 3 * A_SLOT_COLS <= (int)TMEM_COLS) { p->acc_cols = 2 * bn; p->n_bufs = 2; }
            else { p->acc_cols = 256; p->n_bufs = 1; }
        }
    }
    p->na = (int)((TMEM_COLS - p->n_bufs * p->acc_cols) / A_SLOT_COLS);
    if (p->na > NA_MAX) p->na = NA_MAX;
    if (p->na < GEN_W)                                    // every generator warp of a quadrant owns a slot of the rotation
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "low-rank tcgen05 engine: %d A slots for %d generator warps per quadrant", p->na, GEN_W);
    return SSPBO_OK;
}

int sspbo_score_lr_launch(sspbo_ctx* ctx, const CandDesc& cand, int64_t N, int64_t index0, float lam, float gamma_t,
                          float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st) {
    // One chunk of BN = ceil16(rank + 1) <= 256 rows (row 0 of the operand is m')
    const int bn = (ctx->lr_rank + 1 + 15) / 16 * 16;
    LrState* ls;
    int rc = lr_ensure_maps(ctx, bn, &ls);
    if (rc) return rc;
    const int ph = sspbo_lr_phase_path(ctx);
    const bool traj = ctx->L != 1;
    if (traj && cand.mode != 0) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "low-rank tcgen05 engine: generated candidates need L == 1");
    kernel_fn kern = kernel_for(ctx->n, ph, traj);
    if (!kern) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "low-rank tcgen05 engine: domain dimension %d not instantiated", ctx->n);
    if (!ls->attr_set[traj][ph][ctx->n]) {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ls->max_smem));
        ls->attr_set[traj][ph][ctx->n] = true;
    }
    Params p;
    sspbo_lr_fill_common(ctx, cand, N, index0, lam, gamma_t, mu, var, acq, best, ph, &p);
    size_t smem;
    if ((rc = lr_configure(ctx, ls, bn, &p, &smem))) return rc;
    const long long tiles = (N + BM - 1) / BM;
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    kern<<<grid, THREADS, smem, st>>>(ls->hi, ls->lo, p, nullptr, 0);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

// Grouped scoring of R independent runs over one explicit candidate set: fills runs_host[r] (the caller uploads the array to
// runs_dev and has it resident when the launch executes) and enqueues ONE launch.  SSPBO_EUNSUPPORTED, with nothing enqueued,
// when the runs do not share a layout (same rank <= 255, table size and |x| bound; split-fp16 operands; float32 phase chain;
// n <= 4; plain spaces).
int sspbo_score_lr_runs_prepare(sspbo_ctx* const* ctxs, int R, int64_t N, const double* lam, const double* gamma_t,
                                unsigned long long* keys_dev, RunScore* runs_host) {
    sspbo_ctx* c0 = ctxs[0];
    const int bn = (c0->lr_rank + 1 + 15) / 16 * 16;
    if (bn > 256 || N < 1024) return SSPBO_EUNSUPPORTED;    // (rank <= 255: one chunk of split-fp16 operand rows)
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        if (!c || !sspbo_score_lr_supported(c) || c->lr_disabled || c->operand_format >= 1 || c->L != 1 || c->normalize ||
            sspbo_lr_phase_path(c) != 2 || !c->Af_gen || c->n != c0->n || c->n > 4 || c->KC_pad != c0->KC_pad ||
            c->lr_rank != c0->lr_rank || c->x_absmax != c0->x_absmax || c->device != c0->device || !c->has_beta || !c->blr_ready)
            return SSPBO_EUNSUPPORTED;
    }
    if (!kernel_for_runs(c0->n)) return SSPBO_EUNSUPPORTED;
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        LrState* ls;
        int rc = lr_ensure_maps(c, bn, &ls);
        if (rc) return rc;
        RunScore& rs = runs_host[r];
        rs.map_hi = ls->hi; rs.map_lo = ls->lo;
        rs.tab = c->Af_gen; rs.mscale = c->mscale;
        rs.inv_beta = (float)(1.0 / c->beta); rs.lam = (float)lam[r]; rs.gamma_t = (float)gamma_t[r];
        rs.inv_rscale2 = (float)(1.0 / (c->r_scale * c->r_scale));
        rs.prior_var = (float)(1.0 / c->alpha); rs.flag_below = (float)(SSPBO_LR_FLAG_FRACTION / c->alpha);
        rs.best = keys_dev + r; rs.flag_idx = c->flag_idx; rs.stats = c->stats;
    }
    return SSPBO_OK;
}

int sspbo_score_lr_runs_launch(sspbo_ctx* const* ctxs, int R, const CandDesc& cand, int64_t N, const RunScore* runs_dev, cudaStream_t st) {
    sspbo_ctx* c0 = ctxs[0];
    const int bn = (c0->lr_rank + 1 + 15) / 16 * 16;
    LrState* ls;
    int rc = lr_ensure_maps(c0, bn, &ls);
    if (rc) return rc;
    kernel_fn kern = kernel_for_runs(c0->n);
    if (!ls->runs_attr_set[c0->n]) {
        SSPBO_CUDA(c0, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ls->max_smem));
        ls->runs_attr_set[c0->n] = true;
    }
    Params p;
    sspbo_lr_fill_common(c0, cand, N, 0, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, 2, &p);
    size_t smem;
    if ((rc = lr_configure(c0, ls, bn, &p, &smem))) return rc;
    const int grid = R < c0->sm_count ? R : c0->sm_count;
    kern<<<grid, THREADS, smem, st>>>(ls->hi, ls->lo, p, runs_dev, R);
    SSPBO_LAUNCH_CHECK(c0);
    for (int r = 0; r < R; ++r) ctxs[r]->last_engine = SSPBO_ENGINE_UMMA_LR;
    return SSPBO_OK;
}
