// K2, tcgen05 engine with fp16 + e4m3 operands and several output chunks per tile: the tensor-bound regime of the fused
// encode + BLR acquisition + argmax (posterior rank above ~100, up to the full triangular factor).
//
// sspbo_score_lr.cu splits both operands of Y = psi W^T into fp16 (hi, lo) and issues three fp16 products per K step
// (hi hi + hi lo + lo hi): from rank ~100 on that kernel is bound by the tensor pipe, 2/3 of whose work is the two
// correction products.  Those only need ~2^-4 relative accuracy each (they are 2^-11 of the result), so here they run as
// ONE e4m3 product over a K dimension of 128 per k-block:
//     Y = psi_hi W_hi^T                                   kind::f16     K = 64 per k-block   (4 MMAs)
//       + [e4m3(psi) | e4m3(psi_lo 2^10)] [e4m3(W_lo) | e4m3(W_hi 2^-10)]^T   kind::f8f6f4  K = 128 per k-block (4 MMAs)
// into the same float32 accumulator: 8 MMA issue slots of M = 128, N = BN per k-block instead of 12 (tools/probe/
// fp8_probe.cu checks the three hardware conventions this needs; tools/emulate_fp8lo.py the accuracy: var within 1.3e-5
// for every candidate the low-rank form keeps, 6.5e-5 in factor form even on top of an observation).
//
// The operand W is either V of the low-rank form (var = 1/beta + |phi|^2/alpha - |V psi|^2; ranks up to 511 in two chunks
// of 256 rows) or R, the triangular factor of the psi-space covariance (var = 1/beta + |R psi|^2; d rows in chunks of
// <= 256, chunk c skipping the k-blocks before its first non-zero column).  Each chunk is one pass of the generators over
// psi; psi still never leaves tensor memory.  The mean mu = m' . psi needs more than the e4m3 corrections give (a candidate
// on top of an observation has a small score, and 1e-4 of it is ~1e-6 absolute), so the generators accumulate it in float32
// while they produce chunk 0 and hand it to the epilogue through shared memory.  Roles, rings and the A operand in tensor memory are those
// of sspbo_score_lr.cu; the accumulator is BN columns wide (two buffers while BN <= 128, one of 256 columns beyond).
//
// Replaces SSPAgent.eval (agents/ssp_agent.py:125-137) + np.argmax (experiments/demo_blr_gif.py:129-136).
#include "sspbo_lr_common.cuh"
#include <cuda_fp8.h>
#include <stdlib.h>

namespace {
using namespace sspbo_tc;
using namespace sspbo_lr;

constexpr int F8_SHIFT = 10;            // psi_lo is scaled by 2^10 and W_hi by 2^-10 before their e4m3 rounding
constexpr int NA_F8 = 4;                // A ring slots of 64 columns behind 256 accumulator columns

__device__ __forceinline__ void umma_f8_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
}

// float32 pair -> e4m3 pair (first operand -> low byte), one rounding
__device__ __forceinline__ uint32_t e4m3x2_of_floats(float lo, float hi) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return (uint32_t)r;
}

// Operand words of one pair of values (a, b): fp16 hi pair, e4m3 pair of the values, e4m3 pair of the scaled remainders.
// Seven issue slots per pair: two LOP3, FADD2, F2FP, FMUL2 and the two float32 -> e4m3x2 conversions (the remainders used to
// travel through half2: one conversion and one HMUL2 more, and a second rounding).
__device__ __forceinline__ void split_pair_f8(float a, float b, uint32_t& hi, uint32_t& v8, uint32_t& l8) {
    const float ha = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
    const float hb = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
    __half2 h = __floats2half2_rn(ha, hb);
    hi = *reinterpret_cast<uint32_t*>(&h);
    float la, lb;
    const float sc = (float)(1 << F8_SHIFT);
    f2_unpack(f2_mul(f2_sub(f2_pack(a, b), f2_pack(ha, hb)), f2_pack(sc, sc)), la, lb);
    v8 = e4m3x2_of_floats(a, b);
    l8 = e4m3x2_of_floats(la, lb);
}

// PAIR: two CTAs of a cluster (one TPC) share every MMA through cta_group::2: a tile is 256 candidates, 128 per CTA (each its
// own generators, A ring and accumulator in its own tensor memory); each CTA streams only HALF of the operand rows of B into
// its shared memory and the tensor cores of both SMs read both halves.  At 256 operand rows a single CTA has to pull 64 KB
// per k-block from L2 and the MMA warp ends up waiting for B (ncu: profiles/r2_f8_r255_*); the pair pulls 32 KB per SM.  The
// even CTA issues the MMAs; its a_full / b_full / tmem_empty barriers collect the arrivals of both CTAs, and every commit
// is multicast to the a_empty / b_empty / tmem_full barriers of both.
template <int NDIM, int PH, bool TRAJ, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
k_score_f8(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_8, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform: role branches and the MMA issuer loop stay on the uniform datapath
    constexpr int NA = NA_F8;
    const uint32_t ACC_COLS = (uint32_t)p.acc_cols, A_RING_COL0 = (uint32_t)(p.n_bufs * p.acc_cols);
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    constexpr int TILE_M = PAIR ? 2 * BM : BM;                              // candidates per tile
    const int bn_cta = PAIR ? p.BN / 2 : p.BN;                              // operand rows this CTA holds in shared memory
    const int b_tile_bytes = bn_cta * BK * 2;                               // fp16 tile: 64 K x 2 B; e4m3 tile: 128 K x 1 B
    const int b_slot_bytes = 2 * b_tile_bytes;
    uint8_t* b_ring = smem;
    uint8_t* tables = b_ring + (size_t)p.nb * b_slot_bytes;
    float* mp_s = (float*)((long long*)tables + (size_t)p.n_cls * (p.KC_pad / 2) * NDIM);   // m' in operand order (KC_pad)
    // partial sums of the tiles in flight, [NRM_DEPTH tiles][GEN_W warps of a quadrant][128 rows]: mu, and |phi|^2 in trajectory mode
    float* mu_part = mp_s + p.KC_pad;
    float* nrm_part = mu_part + NRM_DEPTH * GEN_W * BM;
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
    load_tables<NDIM, PH>(p, tables, THREADS);
    for (int i = threadIdx.x; i < p.KC_pad; i += THREADS) mp_s[i] = p.mp_op[i];
    for (int i = threadIdx.x; i < (TRAJ ? 2 : 1) * NRM_DEPTH * GEN_W * BM; i += THREADS) mu_part[i] = 0.f;
    if (warp == WARP_TMA && lane == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(a_full(s), PAIR ? 8 : 4); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.nb; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), PAIR ? 2 * EPI_WARPS : EPI_WARPS * 32); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();          // the partner's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long n_tiles = (p.N + TILE_M - 1) / TILE_M;
    const int n_kb = p.n_kb, n_chunks = p.n_chunks, bpt = p.bpt;
    // tile walkers: a single CTA, or the pair (both CTAs walk the same tiles, 128 rows each)
    const long long walker = PAIR ? blockIdx.x >> 1 : blockIdx.x, n_walkers = PAIR ? gridDim.x >> 1 : gridDim.x;
    const long long my_tiles = (n_tiles > walker) ? (n_tiles - walker + n_walkers - 1) / n_walkers : 0;
    const long long total_blocks = my_tiles * bpt;
    const uint32_t buf_mask = (uint32_t)(p.n_bufs - 1);

    if (warp == WARP_TMA) {
        // ================================ TMA producer (B ring) ================================
        if (lane == 0) {
            Ring rb;
            for (long long t = 0; t < my_tiles; ++t)
                for (int c = 0; c < n_chunks; ++c)
                    for (int kb = p.kb0[c]; kb < n_kb; ++kb) {
                        mbar_wait_relaxed<64>(b_empty(rb.slot), rb.phase ^ 1);
                        uint8_t* st = b_ring + (size_t)rb.slot * b_slot_bytes;
                        if constexpr (PAIR) {                             // my half of the chunk's rows; bytes of both halves land on the even CTA's barrier
                            if (cta_rank == 0) mbar_expect_tx(b_full(rb.slot), 2u * (uint32_t)b_slot_bytes);
                            const int row = p.b_row0 + c * p.BN + (int)cta_rank * bn_cta;
                            tma_load_2d_pair(smem_u32(st), &map_hi, b_full(rb.slot), kb * BK, row);
                            tma_load_2d_pair(smem_u32(st + b_tile_bytes), &map_8, b_full(rb.slot), kb * 2 * BK, row);
                        } else {
                            mbar_expect_tx(b_full(rb.slot), (uint32_t)b_slot_bytes);
                            tma_load_2d(smem_u32(st), &map_hi, b_full(rb.slot), kb * BK, p.b_row0 + c * p.BN);
                            tma_load_2d(smem_u32(st + b_tile_bytes), &map_8, b_full(rb.slot), kb * 2 * BK, p.b_row0 + c * p.BN);
                        }
                        rb.advance(p.nb);
                    }
        }
    } else if (warp == WARP_MMA && cta_rank == 0) {
        // ================================ MMA issuer (the even CTA of a pair) ================================
        // Warp-uniform walk over the CTA's k-block sequence (tiles x chunks x k-blocks); the waits for k-block i + 1 sit
        // between the two halves of the MMAs of k-block i, so the tensor pipe has queued work while this warp polls.
        Ring ra, rb;
        uint32_t acc_phase[2] = {0, 0};
        const uint32_t idesc = make_idesc(TILE_M, p.BN);
        const bool no_mma = (SSPBO_DBG_BITS(p) & 1) != 0;
        const uint64_t b_desc0 = make_desc_sw128(smem_u32(b_ring));       // slot s: start address + s * slot bytes (16 B units)
        const uint32_t b_slot_units = (uint32_t)(b_slot_bytes >> 4), b_8_units = (uint32_t)(b_tile_bytes >> 4);
        const uint32_t a_ring0 = tmem_base + A_RING_COL0;
        if (total_blocks > 0 && p.split_issue) {
            mbar_wait_relaxed<40>(a_full(ra.slot), ra.phase);
            mbar_wait(b_full(rb.slot), rb.phase);
        }
        int c = 0, kb = p.kb0[0];
        uint32_t cc = 0;                                                  // running chunk counter: accumulator buffer = cc & mask
        if (!p.split_issue) {
            // One thread walks the sequence: wait for the operands of a k-block, issue its eight MMAs and the commits.  With
            // 200+ operand rows the MMAs of one k-block keep the tensor pipe busy for ~1000 cycles, longer than this loop
            // takes, so nothing is gained by interleaving the waits with the issue (the loop below), and its second elected
            // section, reconvergence points and register -> uniform-register moves are all on the critical path of the
            // generator -> MMA -> free-slot cycle.
            if (lane == 0) {
                uint32_t sa = 0, pa = 0, sb = 0, pb = 0, accph = 0;
                int kb_first = kb;
                for (uint32_t g = 0, gn = (uint32_t)total_blocks; g < gn; ++g) {
                    const uint32_t buf = cc & buf_mask;
                    const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
                    const bool first = kb == kb_first, last = kb == n_kb - 1;
                    if (first) {
                        mbar_wait(tmem_empty(buf), ((accph >> buf) & 1u) ^ 1u);
                        accph ^= 1u << buf;
                    }
                    mbar_wait(a_full((int)sa), pa);
                    mbar_wait(b_full((int)sb), pb);
                    tc_fence_after();
                    const uint32_t a_hi = a_ring0 + sa * (uint32_t)A_SLOT_COLS, a_8 = a_hi + 32;
                    const uint64_t b_hi = b_desc0 + (uint64_t)(sb * b_slot_units);
                    const uint64_t b_8 = b_hi + b_8_units;
                    if constexpr (PAIR) {
                        umma_f16_ts_pair(d_tmem, a_hi, b_hi, idesc, first ? 0u : 1u);
#pragma unroll
                        for (int k = 1; k < 4; ++k) umma_f16_ts_pair(d_tmem, a_hi + 8 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f8_ts_pair(d_tmem, a_8 + 8 * k, b_8 + 2 * k, idesc, 1u);
                        umma_commit_pair(b_empty((int)sb));
                        umma_commit_pair(a_empty((int)sa));
                        if (last) umma_commit_pair(tmem_full(buf));
                    } else {
                        umma_f16_ts(d_tmem, a_hi, b_hi, idesc, first ? 0u : 1u);
#pragma unroll
                        for (int k = 1; k < 4; ++k) umma_f16_ts_acc(d_tmem, a_hi + 8 * k, b_hi + 2 * k, idesc);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f8_ts_acc(d_tmem, a_8 + 8 * k, b_8 + 2 * k, idesc);
                        umma_commit(b_empty((int)sb));
                        umma_commit(a_empty((int)sa));
                        if (last) umma_commit(tmem_full(buf));
                    }
                    if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1u; }
                    if (++sb == (uint32_t)p.nb) { sb = 0; pb ^= 1u; }
                    if (++kb == n_kb) { ++cc; if (++c == n_chunks) c = 0; kb = kb_first = p.kb0[c]; }
                }
            }
            __syncwarp();
        } else
        for (long long g = 0; g < total_blocks; ++g) {
            const uint32_t buf = cc & buf_mask;
            const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
            const bool first = kb == p.kb0[c], last = kb == n_kb - 1;
            if (first) {
                mbar_wait(tmem_empty(buf), acc_phase[buf] ^ 1);           // the epilogue has drained this accumulator
                acc_phase[buf] ^= 1;
            }
            tc_fence_after();
            // fp16 part: 16 values along K = 32 bytes of shared memory (2 descriptor units) = 8 TMEM columns;
            // e4m3 part: 32 values along K = 32 bytes = 8 TMEM columns, tile behind the fp16 tile of the B slot
            const uint32_t a_hi = a_ring0 + (uint32_t)(ra.slot * A_SLOT_COLS), a_8 = a_hi + 32;
            const uint64_t b_hi = b_desc0 + (uint64_t)((uint32_t)rb.slot * b_slot_units);
            const uint64_t b_8 = b_hi + b_8_units;
            const uint32_t a_bar = a_empty(ra.slot), b_bar = b_empty(rb.slot);
            if (elect_one() && !no_mma) {
                if constexpr (PAIR) {
                    umma_f16_ts_pair(d_tmem, a_hi, b_hi, idesc, first ? 0u : 1u);
                    umma_f16_ts_pair(d_tmem, a_hi + 8, b_hi + 2, idesc, 1u);
                    umma_f8_ts_pair(d_tmem, a_8, b_8, idesc, 1u);
                    umma_f8_ts_pair(d_tmem, a_8 + 8, b_8 + 2, idesc, 1u);
                } else {
                    umma_f16_ts(d_tmem, a_hi, b_hi, idesc, first ? 0u : 1u);  // the first MMA of a chunk overwrites the accumulator
                    umma_f16_ts_acc(d_tmem, a_hi + 8, b_hi + 2, idesc);
                    umma_f8_ts_acc(d_tmem, a_8, b_8, idesc);
                    umma_f8_ts_acc(d_tmem, a_8 + 8, b_8 + 2, idesc);
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
                if constexpr (PAIR) {
                    if (!no_mma) {
                        umma_f16_ts_pair(d_tmem, a_hi + 16, b_hi + 4, idesc, 1u);
                        umma_f16_ts_pair(d_tmem, a_hi + 24, b_hi + 6, idesc, 1u);
                        umma_f8_ts_pair(d_tmem, a_8 + 16, b_8 + 4, idesc, 1u);
                        umma_f8_ts_pair(d_tmem, a_8 + 24, b_8 + 6, idesc, 1u);
                    }
                    umma_commit_pair(b_bar);                              // both CTAs' B slots, A slots, accumulators
                    umma_commit_pair(a_bar);
                    if (last) umma_commit_pair(tmem_full(buf));
                } else {
                    if (!no_mma) {
                        umma_f16_ts_acc(d_tmem, a_hi + 16, b_hi + 4, idesc);
                        umma_f16_ts_acc(d_tmem, a_hi + 24, b_hi + 6, idesc);
                        umma_f8_ts_acc(d_tmem, a_8 + 16, b_8 + 4, idesc);
                        umma_f8_ts_acc(d_tmem, a_8 + 24, b_8 + 6, idesc);
                    }
                    umma_commit(b_bar);                                   // frees the B slot when these MMAs retire
                    umma_commit(a_bar);                                   // ... and the A slot
                    if (last) umma_commit(tmem_full(buf));
                }
            }
            __syncwarp();
            if (++kb == n_kb) { ++cc; if (++c == n_chunks) c = 0; kb = p.kb0[c]; }
        }
    } else if (warp >= WARP_EPI0 && warp < WARP_GEN0) {
        // ================================ epilogue ================================
        const int quad = warp & 3;                                        // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        uint32_t acc_phase[2] = {0, 0};
        uint32_t cc = 0;
        unsigned long long my_best = 0ull;
        for (long long t = 0; t < my_tiles; ++t) {
            const long long tile = walker + t * n_walkers;
            float q = 0.f;
            for (int c = 0; c < n_chunks; ++c, ++cc) {
                const uint32_t buf = cc & buf_mask;
                mbar_wait_relaxed<100>(tmem_full(buf), acc_phase[buf]);
                acc_phase[buf] ^= 1;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * ACC_COLS;
                for (int g = 0; g < p.BN; g += 64) {                      // BN % 16 == 0; up to 64 columns in flight
                    uint32_t v[64];
                    const int nld = min(4, (p.BN - g) >> 4);
#pragma unroll
                    for (int l = 0; l < 4; ++l)
                        if (l < nld) tmem_ld16(taddr + (uint32_t)(g + 16 * l), v + 16 * l);
                    tmem_ld_wait();
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
                if constexpr (PAIR) { __syncwarp(); if (lane == 0) mbar_arrive_leader(tmem_empty(buf)); }
                else mbar_arrive(tmem_empty(buf));
            }
            const long long gi = tile * TILE_M + (long long)cta_rank * BM + row;
            float mu = 0.f;
            {   // the generators' shares of mu = m' . psi (warps without a k-block in chunk 0 of a tile write nothing)
                float* mp_ = mu_part + (size_t)(t & (NRM_DEPTH - 1)) * GEN_W * BM + row;
#pragma unroll
                for (int w = 0; w < GEN_W; ++w) { mu += mp_[w * BM]; mp_[w * BM] = 0.f; }
            }
            if (gi < p.N) {
                float prior = p.prior_var, flag_below = p.flag_below;
                q *= p.inv_rscale2;
                if constexpr (TRAJ) {                                                 // |phi|^2 / alpha instead of 1 / alpha
                    float* np_ = nrm_part + (size_t)(t & (NRM_DEPTH - 1)) * GEN_W * BM + row;
                    float ssum = 0.f;
#pragma unroll
                    for (int w = 0; w < GEN_W; ++w) { ssum += np_[w * BM]; np_[w * BM] = 0.f; }   // warps without a k-block in a tile write nothing
                    const float nrm2 = fmaf(p.nrm_scale, ssum, -p.nrm_offset);
                    if (p.normalize) { const float inv = 1.f / nrm2; mu *= sqrtf(inv); q *= inv; }   // phi / |phi|
                    else { prior *= nrm2; flag_below *= nrm2; }
                }
                bool flagged = false;
                float var;
                if (p.form == 0) {
                    var = p.inv_beta + (prior - q);                                   // blr.py:96 with S = I/alpha - V^T V
                    if (var < flag_below) {                                         // too close to the observations for
                        const unsigned long long slot = atomicAdd(p.stats + SSPBO_STAT_CUR, 1ull);   // the cancelling form
                        if (slot < (unsigned long long)p.flag_cap) {
                            p.flag_idx[slot] = (unsigned int)gi; flagged = true;
                            p.flag_idx[p.flag_cap + slot] = __float_as_uint(flagged_score_bound(mu, var, prior, p.lam, p.gamma_t));
                        }
                        else atomicAdd(p.stats + SSPBO_STAT_OVERFLOW, 1ull);
                    }
                } else {
                    var = p.inv_beta + q;                                             // blr.py:96 with psi^T Sp psi = |R psi|^2
                    if (var < flag_below) {                                         // nearly on top of an observation: the e4m3
                        const unsigned long long slot = atomicAdd(p.stats + SSPBO_STAT_CUR, 1ull);   // corrections are too coarse
                        if (slot < (unsigned long long)p.flag_cap) {
                            p.flag_idx[slot] = (unsigned int)gi; flagged = true;
                            p.flag_idx[p.flag_cap + slot] = __float_as_uint(flagged_score_bound(mu, var, prior, p.lam, p.gamma_t));
                        }
                        else atomicAdd(p.stats + SSPBO_STAT_OVERFLOW, 1ull);
                    }
                }
                if (!flagged) {
                    const unsigned long long key = finish_candidate(p, gi, mu, var);
                    if (key > my_best) my_best = key;
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, my_best, o);
            if (other > my_best) my_best = other;
        }
        if (lane == 0 && my_best && p.best) atomicMax(p.best, my_best);
    } else if (warp >= WARP_GEN0) {
        // ================================ psi generators (A ring in tensor memory) ================================
        // GEN_W warps per TMEM lane quadrant (thread <-> candidate row); warp j of a quadrant produces the k-blocks
        // g = j, j + GEN_W, ... of the CTA's k-block sequence.  psi is regenerated for every chunk of a tile.
        const int gw = warp - WARP_GEN0;
        const int quad = warp & 3, member = gw >> 2;                      // TMEM lane quadrant, position in the rotation
        const int r0 = quad * 32 + lane;                                  // candidate row of this thread = TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + A_RING_COL0;
        int slot = member % NA;
        uint32_t phase = (uint32_t)((member / NA) & 1);
        int pos = member % bpt;                                           // position in the tile's k-block sequence
        long long tile_iter = member / bpt, loaded_iter = -1;
        unsigned long long x0[NDIM];                                      // Q31.32 candidate (two's complement)
        int xq0[NDIM];                                                    // Q.fx / float32 candidate of the 32-bit phase paths
        float nrm_acc = 0.f, mu_acc = 0.f;                                // sum of C^2 + S^2 / of m' psi over this warp's k-blocks of chunk 0
        long long g0 = 0;
        bool row_nan = false;                                             // a coordinate of this row is NaN: its score must be NaN (np.argmax semantics)
        for (long long g = member; g < total_blocks; g += GEN_W) {
            if (tile_iter != loaded_iter) {                               // first k-block of a tile for this warp: its row
                loaded_iter = tile_iter;
                nrm_acc = 0.f; mu_acc = 0.f;
                g0 = (walker + tile_iter * n_walkers) * TILE_M + (long long)cta_rank * BM + r0;
                if constexpr (!TRAJ) {
                    const bool oor = load_candidate<NDIM, PH>(p, g0, x0, xq0, row_nan);
                    if (oor && member == 0) atomicAdd(p.stats + SSPBO_STAT_OOR, 1ull);  // outside the declared |x| bound
                    const long long gn = g0 + n_walkers * TILE_M;       // next tile's row: pull it towards the SM now
                    if (member == 0 && gn < p.N && p.cand.mode == 0) {
                        const char* nx = (const char*)p.cand.x + (size_t)gn * NDIM * (p.cand.x_f64 ? 8 : 4);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                    }
                }
            }
            int c = 0;
#pragma unroll
            for (int i = 1; i < MAX_CHUNKS; ++i) if (i < n_chunks && pos >= p.pos0[i]) c = i;
            const int kb = p.kb0[c] + (pos - p.pos0[c]);
            const uint32_t empty_bar = a_empty(slot);
            bool slot_free = mbar_test_wait(empty_bar, phase ^ 1);        // early, non-blocking probe
            const uint32_t slot_addr = lane_addr + (uint32_t)(slot * A_SLOT_COLS);
            float nrm_blk = 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float cs[NF], sn[NF];
                gen_half<NDIM, PH, TRAJ>(p, tables, g0, kb, half, x0, xq0, cs, sn, nrm_blk);
                if (!TRAJ && row_nan) cs[0] = __int_as_float(0x7fc00000);      // poisons mu and |V psi|^2 of the row through the MMA
                if (c == 0) {                                             // mu = m' . psi in float32 (m' in operand order: cos | sin per k-block)
                    const f32x2* mc2 = (const f32x2*)(mp_s + kb * BK + half * NF);      // two frequencies per FFMA2
                    f32x2 m0 = 0ull, m1 = 0ull;
#pragma unroll
                    for (int i = 0; i < NF / 2; ++i) {
                        m0 = f2_fma(f2_pack(cs[2 * i], cs[2 * i + 1]), mc2[i], m0);
                        m1 = f2_fma(f2_pack(sn[2 * i], sn[2 * i + 1]), mc2[16 + i], m1);
                    }
                    float a0, a1, b0, b1;
                    f2_unpack(m0, a0, a1); f2_unpack(m1, b0, b1);
                    mu_acc += (a0 + a1) + (b0 + b1);
                }
                // fp16 hi pairs (8 columns each) and e4m3 quadruples (4 columns each): values, then scaled remainders
                uint32_t wch[NF / 2], wsh[NF / 2], wc8[NF / 4], ws8[NF / 4], wcl[NF / 4], wsl[NF / 4];
#pragma unroll
                for (int i = 0; i < NF / 4; ++i) {
                    uint32_t v0, v1, l0, l1;
                    split_pair_f8(cs[4 * i], cs[4 * i + 1], wch[2 * i], v0, l0);
                    split_pair_f8(cs[4 * i + 2], cs[4 * i + 3], wch[2 * i + 1], v1, l1);
                    wc8[i] = v0 | (v1 << 16); wcl[i] = l0 | (l1 << 16);
                    split_pair_f8(sn[4 * i], sn[4 * i + 1], wsh[2 * i], v0, l0);
                    split_pair_f8(sn[4 * i + 2], sn[4 * i + 3], wsh[2 * i + 1], v1, l1);
                    ws8[i] = v0 | (v1 << 16); wsl[i] = l0 | (l1 << 16);
                }
                if (!slot_free) { mbar_wait(empty_bar, phase ^ 1); slot_free = true; }
                tc_fence_after();
                // operand order inside a k-block: K index i < 32 is cos(f = i), i >= 32 is sin(f = i - 32).  fp16 part: two
                // per column in columns [0, 32) of the slot.  e4m3 part, four per column in columns [32, 64): K8 index
                // j < 64 = e4m3(psi) in the same order, j >= 64 = e4m3(psi_lo 2^10)
                const uint32_t sa = slot_addr + (uint32_t)(half * (NF / 2));
                const uint32_t s8 = slot_addr + 32u + (uint32_t)(half * (NF / 4));
                tmem_st8(sa, wch);
                tmem_st8(sa + 16, wsh);
                tmem_st4(s8, wc8);
                tmem_st4(s8 + 8, ws8);
                tmem_st4(s8 + 16, wcl);
                tmem_st4(s8 + 24, wsl);
            }
            if (c == 0) {                                                 // chunk 0 carries mu and |phi|^2; this warp's last k-block of it
                nrm_acc += nrm_blk;
                if (kb + GEN_W >= n_kb) {
                    const size_t o = (size_t)(tile_iter & (NRM_DEPTH - 1)) * GEN_W * BM + member * BM + r0;
                    mu_part[o] = mu_acc;
                    if constexpr (TRAJ) nrm_part[o] = nrm_acc;
                }
            }
            tmem_st_wait();                                               // publish the k-block
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(a_full(slot)); else mbar_arrive(a_full(slot)); }
            slot += GEN_W; while (slot >= NA) { slot -= NA; phase ^= 1; }
            pos += GEN_W; while (pos >= bpt) { pos -= bpt; ++tile_iter; }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();          // the partner has finished with my barriers and tensor memory
    if (warp == WARP_MMA) {
        __syncwarp();
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------ operand images
__device__ __forceinline__ uint8_t e4m3_of(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }

// e4m3 image of rows [row0, row1) of a split-fp16 operand (hi, lo: rows x KC_pad): per k-block of 64 columns one 128-byte
// group, bytes [0, 64) = e4m3(lo), bytes [64, 128) = e4m3(hi 2^-10)
__global__ void k_build_8_rows(const __half* __restrict__ hi, const __half* __restrict__ lo, int KC_pad, int row0, int row1,
                               uint8_t* __restrict__ out) {
    const int row = row0 + blockIdx.x;
    if (row >= row1) return;
    const float sc = 1.f / (float)(1 << F8_SHIFT);
    for (int c = threadIdx.x; c < KC_pad; c += blockDim.x) {
        const size_t o = (size_t)row * 2 * KC_pad + (size_t)(c >> 6) * 128 + (c & 63);
        out[o] = e4m3_of(__half2float(lo[(size_t)row * KC_pad + c]));
        out[o + 64] = e4m3_of(__half2float(hi[(size_t)row * KC_pad + c]) * sc);
    }
}

// Operand images of the factor form: row j = column j of the lower Cholesky factor L of the psi-space covariance,
// F[j][col_of_rank[kk]] = L[kk][j] r_scale for kk >= j; everything else stays zero (cleared once at allocation: the pattern
// never changes).
__global__ void k_build_factor_images(const double* __restrict__ L, const int* __restrict__ col_of_rank, int d, int KC_pad,
                                      double r_scale, __half* __restrict__ Fh, uint8_t* __restrict__ F8) {
    __shared__ double tile[32][33];
    const int kk0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const float sc = 1.f / (float)(1 << F8_SHIFT);
    if (j0 > kk0 + 31) return;                                            // strictly upper tile of L: structurally zero
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int kk = kk0 + r, j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (kk < d && j <= kk) ? L[(size_t)kk * d + j] * r_scale : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int j = j0 + r, kk = kk0 + threadIdx.x;
        if (j < d && kk < d) {
            const double v = tile[threadIdx.x][r];
            const __half h = __double2half(v);
            const int c = col_of_rank[kk];
            Fh[(size_t)j * KC_pad + c] = h;
            const size_t o = (size_t)j * 2 * KC_pad + (size_t)(c >> 6) * 128 + (c & 63);
            F8[o] = e4m3_of((float)(v - (double)__half2float(h)));
            F8[o + 64] = e4m3_of(__half2float(h) * sc);
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
struct MapPair {
    CUtensorMap hi, b8;
    const void* bh = nullptr; const void* b8p = nullptr;
    int KC_pad = 0, BN = 0, rows = 0;
};
struct F8State {
    MapPair lr[2], fac[2];                                                // [single CTA, CTA pair (boxes of BN / 2 rows)]
    int max_smem = 0;
    bool attr_set[2][2][3][MAX_N + 1] = {};
};

size_t table_bytes(const sspbo_ctx* ctx) {
    return (size_t)ctx->n_cls * (ctx->KC_pad / 2) * ctx->n * sizeof(long long) + (size_t)ctx->KC_pad * sizeof(float) +
           (ctx->L != 1 ? 2 : 1) * NRM_DEPTH * GEN_W * BM * sizeof(float) + (2 * NA_F8 + 2 * MAX_NB + 4) * 8 + 16;
}

typedef void (*kernel_fn)(const CUtensorMap, const CUtensorMap, const Params);
// instantiated for the domain dimensions the 1-chunk engine does not already serve well: every n <= 8 with the two 32-bit
// phase paths, plus Q31.32 as the any-range fallback; trajectories with waypoints of up to 3 dimensions
template <int PH, bool PAIR>
kernel_fn kernel_for_n(int n) {
    switch (n) {
        case 1: return k_score_f8<1, PH, false, PAIR>; case 2: return k_score_f8<2, PH, false, PAIR>;
        case 3: return k_score_f8<3, PH, false, PAIR>; case 4: return k_score_f8<4, PH, false, PAIR>;
        case 5: return k_score_f8<5, PH, false, PAIR>; case 6: return k_score_f8<6, PH, false, PAIR>;
        case 7: return k_score_f8<7, PH, false, PAIR>; case 8: return k_score_f8<8, PH, false, PAIR>;
    }
    if constexpr (PH != 1 && !PAIR) {
        switch (n) {
            case 9: return k_score_f8<9, PH, false, false>; case 10: return k_score_f8<10, PH, false, false>;
            case 11: return k_score_f8<11, PH, false, false>; case 12: return k_score_f8<12, PH, false, false>;
            case 13: return k_score_f8<13, PH, false, false>; case 14: return k_score_f8<14, PH, false, false>;
            case 15: return k_score_f8<15, PH, false, false>; case 16: return k_score_f8<16, PH, false, false>;
        }
    }
    return nullptr;
}
template <int PH, bool PAIR>
kernel_fn kernel_for_traj(int n) {
    switch (n) {
        case 1: return k_score_f8<1, PH, true, PAIR>; case 2: return k_score_f8<2, PH, true, PAIR>; case 3: return k_score_f8<3, PH, true, PAIR>;
    }
    return nullptr;
}
template <bool PAIR>
kernel_fn kernel_for_p(int n, int ph, bool traj) {
    if (traj) return ph == 2 ? kernel_for_traj<2, PAIR>(n) : kernel_for_traj<0, PAIR>(n);
    return ph == 2 ? kernel_for_n<2, PAIR>(n) : (ph == 1 ? kernel_for_n<1, PAIR>(n) : kernel_for_n<0, PAIR>(n));
}
kernel_fn kernel_for(int n, int ph, bool traj, bool pair) {
    return pair ? kernel_for_p<true>(n, ph, traj) : kernel_for_p<false>(n, ph, traj);
}

int make_map_u8(sspbo_ctx* ctx, CUtensorMap* map, const void* base, int KC_pad, int rows, int BN) {
    cuuint64_t gdim[2] = {(cuuint64_t)KC_pad * 2, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)KC_pad * 2};
    cuuint32_t box[2] = {128u, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
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
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SSPBO_FAIL(ctx, SSPBO_ECUDA, "cuTensorMapEncodeTiled (uint8) failed (%d)", (int)r);
    return SSPBO_OK;
}

// BN here = rows of one TMA box (the chunk's rows, or half of them when a CTA pair shares the chunk)
int ensure_maps(sspbo_ctx* ctx, MapPair* mp, const void* bh, const void* b8, int KC_pad, int rows, int BN) {
    if (mp->bh == bh && mp->b8p == b8 && mp->KC_pad == KC_pad && mp->rows == rows && mp->BN == BN) return SSPBO_OK;
    int rc;
    if ((rc = sspbo_make_map_f16(ctx, &mp->hi, bh, KC_pad, rows, BN)) || (rc = make_map_u8(ctx, &mp->b8, b8, KC_pad, rows, BN))) return rc;
    mp->bh = bh; mp->b8p = b8; mp->KC_pad = KC_pad; mp->rows = rows; mp->BN = BN;
    return SSPBO_OK;
}

// CTA pairs pay off once a single CTA would wait for its B operand (BN > 128); operand format 1 keeps single CTAs, 2 pairs
// wherever the shape allows (BN / 2 a multiple of 8 rows, the candidate set at least a few tiles per pair)
bool want_pair(const sspbo_ctx* ctx, int bn, int64_t N) {
    if (ctx->n > 8 || bn % 16 != 0 || ctx->sm_count < 2) return false;
    if (ctx->operand_format == 1) return false;
    // measured on C3 (tools/pair_check.py): never faster than single CTAs with the one-thread issue loop -- at 256 operand rows
    // both sit at the rate the generators deliver psi (0.92-0.98e9 candidates/s), below that the pair's cross-CTA barriers cost
    // 5-15 % -- so the pair is only run on request
    (void)N;
    return ctx->operand_format == 2;
}

// chunks of the factor-form operand: d rows in n_chunks chunks of BN rows
void factor_geometry(const sspbo_ctx* ctx, int* bn, int* nc) {
    const int rows = ctx->d;
    *nc = (rows + 255) / 256;
    *bn = ((rows + *nc - 1) / *nc + 15) / 16 * 16;
}

}  // namespace

void sspbo_f8_release(sspbo_ctx* ctx) {
    if (ctx->f8_state) { delete (F8State*)ctx->f8_state; ctx->f8_state = nullptr; }
    if (ctx->V8) { cudaFree(ctx->V8); ctx->V8 = nullptr; }
    if (ctx->Fh) { cudaFree(ctx->Fh); ctx->Fh = nullptr; }
    if (ctx->F8) { cudaFree(ctx->F8); ctx->F8 = nullptr; }
    ctx->v8_rank = 0; ctx->v8_row0_dirty = true; ctx->f8_factor_dirty = true;
}

// form 0 (low rank): every rank the rows of V cover, up to SSPBO_LR_MAX; form 1 (factor): any posterior over an SSP space
int sspbo_score_f8_supported(const sspbo_ctx* ctx, int form) {
    if (ctx->L != 1 && (ctx->n > MAX_TRAJ_N || ctx->L > 64 || !ctx->tauq_op)) return 0;
    if (ctx->cc_major != 10 || !ctx->has_factor || ctx->n > MAX_N || !ctx->umma_range_ok || !ctx->stats) return 0;
    if (form == 0) return ctx->lr_valid && ctx->lr_rank >= 1 && ctx->lr_rank <= SSPBO_LR_MAX;
    int bn, nc;
    factor_geometry(ctx, &bn, &nc);
    return nc <= MAX_CHUNKS;
}

// Estimated seconds of one scoring call of N candidates in the low-rank (0) or the factor form (1): the cost model
// SSPBO_ENGINE_AUTO compares the two with.  A k-block pass over a tile costs max(rows x 3.9 ns, 0.5 us) -- the tensor pipe
// (measured: 256 rows x 16 k-blocks = 15.8 us at rank 255, d = 1023) or the generators, whichever binds -- and every SM
// sees ceil(N / 128 / SMs) tiles.  The low-rank form pays for its float64 pass over flagged candidates; the factor form pays
// for refreshing the factor when the posterior has changed since the last call (the rows since folded into the psi-space
// covariance, a Cholesky factorisation, the operand images: ~0.1 ms of small launches + d^3/3 float64 flops), which at small d
// outweighs any difference between the scorers (config C2 late in its run: d = 151, rank > 118).
double sspbo_f8_cost(const sspbo_ctx* ctx, int form, int64_t N) {
    const int n_kb = ctx->KC_pad / BK;
    const double tiles = ceil((double)N / (double)BM / (double)(ctx->sm_count > 0 ? ctx->sm_count : 148));
    auto pass = [](int rows) { const double t = rows * 3.9e-9; return t > 0.5e-6 ? t : 0.5e-6; };
    if (form == 0) {
        const int rows = ctx->lr_rank, nc = (rows + 255) / 256;
        const int bn = ((rows + nc - 1) / nc + 15) / 16 * 16;
        // plus the float64 pass over the flagged candidates: one of its candidate-rows costs ~36 tensor-pass candidate-rows
        // (measured at rank 511, 1.4 % flagged: 4.3 ms of 12.9 ms per 4.19M candidates)
        return tiles * nc * n_kb * pass(bn) * (1.0 + 36.0 * ctx->lr_flag_frac);
    }
    int bn, nc;
    factor_geometry(ctx, &bn, &nc);
    double cost = 0.0;
    for (int c = 0; c < nc; ++c) cost += (double)(n_kb - ctx->host_col_of_rank[c * bn] / 64) * pass(bn);
    cost *= tiles;
    if (ctx->factor_dirty) cost += 1.0e-4 + (double)ctx->d * ctx->d * ctx->d / 3.0 / 4.0e11;
    return cost;
}

int sspbo_score_f8_launch(sspbo_ctx* ctx, const CandDesc& cand, int64_t N, int64_t index0, float lam, float gamma_t,
                          float* mu, float* var, float* acq, unsigned long long* best, int form, cudaStream_t st) {
    F8State* fs = (F8State*)ctx->f8_state;
    if (!fs) { fs = new F8State(); ctx->f8_state = fs; }
    if (!fs->max_smem) cudaDeviceGetAttribute(&fs->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    const int KC = ctx->KC_pad, n_kb = KC / BK;
    const int ph = sspbo_lr_phase_path(ctx);
    const bool traj = ctx->L != 1;
    if (traj && cand.mode != 0) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "fp16+e4m3 tcgen05 engine: generated candidates need L == 1");
    Params p;
    sspbo_lr_fill_common(ctx, cand, N, index0, lam, gamma_t, mu, var, acq, best, ph, &p);
    p.form = form; p.mp_op = ctx->mp_op;

    if (form == 1) p.flag_below = (float)(SSPBO_F8_FLAG_FRACTION / ctx->alpha);
    MapPair* mp;
    int rc;
    bool pair = false;
    if (form == 0) {
        const int rows = ctx->lr_rank;                                    // rows 1 .. rank of the images (row 0 holds m')
        const int nc = (rows + 255) / 256;
        const int bn = ((rows + nc - 1) / nc + 15) / 16 * 16;
        p.b_row0 = 1;
        if (!ctx->V8) {
            SSPBO_CUDA(ctx, cudaMalloc((void**)&ctx->V8, (size_t)(SSPBO_LR_MAX + 1) * 2 * KC));
            SSPBO_CUDA(ctx, cudaMemsetAsync(ctx->V8, 0, (size_t)(SSPBO_LR_MAX + 1) * 2 * KC, st));
            ctx->v8_rank = 0; ctx->v8_row0_dirty = true;
        }
        if (ctx->v8_rank < ctx->lr_rank) {                                // rows appended since the last call
            k_build_8_rows<<<ctx->lr_rank - ctx->v8_rank, 256, 0, st>>>(ctx->Vh, ctx->Vl, KC, ctx->v8_rank + 1, ctx->lr_rank + 1, ctx->V8);
            SSPBO_LAUNCH_CHECK(ctx);
            ctx->v8_rank = ctx->lr_rank;
        }
        pair = want_pair(ctx, bn, N);
        mp = &fs->lr[pair];
        if ((rc = ensure_maps(ctx, mp, ctx->Vh, ctx->V8, KC, SSPBO_LR_MAX + 1, pair ? bn / 2 : bn))) return rc;
        p.BN = bn; p.n_chunks = nc;
        for (int c = 0; c < nc; ++c) { p.kb0[c] = 0; p.pos0[c] = c * n_kb; }
        p.bpt = nc * n_kb;
    } else {
        int bn, nc;
        factor_geometry(ctx, &bn, &nc);
        if (nc > MAX_CHUNKS) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "fp16+e4m3 tcgen05 engine: d = %d needs %d chunks", ctx->d, nc);
        const size_t rows = (size_t)nc * bn;
        if (!ctx->Fh) {
            SSPBO_CUDA(ctx, cudaMalloc((void**)&ctx->Fh, rows * KC * sizeof(__half)));
            SSPBO_CUDA(ctx, cudaMalloc((void**)&ctx->F8, rows * 2 * KC));
            SSPBO_CUDA(ctx, cudaMemsetAsync(ctx->Fh, 0, rows * KC * sizeof(__half), st));
            SSPBO_CUDA(ctx, cudaMemsetAsync(ctx->F8, 0, rows * 2 * KC, st));
            ctx->f8_factor_dirty = true;
        }
        if (ctx->f8_factor_dirty) {
            if ((rc = sspbo_refresh_factor(ctx, st))) return rc;
            const int tiles = (ctx->d + 31) / 32;
            k_build_factor_images<<<dim3(tiles, tiles), dim3(32, 8), 0, st>>>(ctx->R, ctx->col_of_rank, ctx->d, KC, ctx->r_scale,
                                                                              ctx->Fh, ctx->F8);
            SSPBO_LAUNCH_CHECK(ctx);
            ctx->f8_factor_dirty = false;
        }
        pair = want_pair(ctx, bn, N);
        mp = &fs->fac[pair];
        if ((rc = ensure_maps(ctx, mp, ctx->Fh, ctx->F8, KC, (int)rows, pair ? bn / 2 : bn))) return rc;
        p.BN = bn; p.n_chunks = nc;
        int pos = 0;
        for (int c = 0; c < nc; ++c) {
            p.kb0[c] = ctx->host_col_of_rank[c * bn] / 64;               // first factor column held by the chunk
            p.pos0[c] = pos;
            pos += n_kb - p.kb0[c];
        }
        p.bpt = pos;
    }
    // MMA issue loop: the interleaved one while a k-block's MMAs are short (measured on C3: BN = 208 1.22e9 against 1.10e9
    // candidates/s; BN = 256 0.84e9 against 0.92e9; factor form at d = 1023 0.33-0.44e9 against 0.40-0.48e9)
    { static const int split_env = getenv("SSPBO_F8_SPLIT_ISSUE") ? atoi(getenv("SSPBO_F8_SPLIT_ISSUE")) : -1;
      p.split_issue = split_env >= 0 ? split_env : (p.BN < 240 ? 1 : 0); }
    const size_t tb = table_bytes(ctx), sb = 2 * (size_t)(pair ? p.BN / 2 : p.BN) * BK * 2;
    int nb = (int)(((size_t)fs->max_smem - 1024 - tb) / sb);
    if (nb > MAX_NB) nb = MAX_NB;
    if (nb < 2) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "fp16+e4m3 tcgen05 engine: shared memory does not fit two B slots");
    p.nb = nb;
    const size_t smem = 1024 + (size_t)nb * sb + tb;
    kernel_fn kern = kernel_for(ctx->n, ph, traj, pair);
    if (!kern) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "fp16+e4m3 tcgen05 engine: domain dimension %d not instantiated", ctx->n);
    if (!fs->attr_set[pair][traj][ph][ctx->n]) {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fs->max_smem));
        fs->attr_set[pair][traj][ph][ctx->n] = true;
    }
    // TMEM: accumulator buffer(s) of BN columns first (two while BN <= 128), the four A slots behind them
    if (p.BN <= 128) { p.acc_cols = 128; p.n_bufs = 2; } else { p.acc_cols = 256; p.n_bufs = 1; }
    p.na = NA_F8;
    if (pair) {
        // clusters of two CTAs (one TPC): tiles of 256 candidates, one pair per two SMs
        const long long tiles = (N + 2 * BM - 1) / (2 * BM);
        const int pairs = (int)(tiles < ctx->sm_count / 2 ? tiles : ctx->sm_count / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        SSPBO_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, mp->hi, mp->b8, p));
        ++ctx->launches;
        return SSPBO_OK;
    }
    const long long tiles = (N + BM - 1) / BM;
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    kern<<<grid, THREADS, smem, st>>>(mp->hi, mp->b8, p);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}
