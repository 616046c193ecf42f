// Shared by the two tensor-memory scorers (sspbo_score_lr.cu: split-fp16 operands; sspbo_score_f8.cu: fp16 + e4m3
// operands, several output chunks): launch parameters, the candidate descriptor (explicit rows / counter-based uniform
// stream / n-D grid generated in the kernel) and the psi generator (phase arithmetic + MUFU sin/cos) of one half k-block.
#pragma once
#include "sspbo_internal.cuh"
#include "sspbo_tc.cuh"

namespace sspbo_lr {
using namespace sspbo_tc;

constexpr int BM = 128;                 // candidates per tile (UMMA M)
constexpr int BK = 64;                  // operand columns per k-block
#ifndef SSPBO_LR_GEN_W
#define SSPBO_LR_GEN_W 3
#endif
constexpr int NF = 16;                  // frequencies per generator thread and pass (two passes per k-block)
constexpr int GEN_W = SSPBO_LR_GEN_W;   // generator warps per TMEM lane quadrant (= per SM sub-partition), k-blocks in rotation
constexpr int GEN_WARPS = 4 * GEN_W, EPI_WARPS = 4;
constexpr int WARP_TMA = 0, WARP_MMA = 1, WARP_EPI0 = 2, WARP_GEN0 = WARP_EPI0 + EPI_WARPS;
constexpr int THREADS = 32 * (WARP_GEN0 + GEN_WARPS);            // 576 with three generator warps per quadrant
constexpr int NA_MAX = 6;               // A ring slots in TMEM: 4 behind 128-column accumulators, 6 behind 64-column ones
constexpr int A_SLOT_COLS = 64;         // 32 columns hi + 32 columns lo (two fp16 / four e4m3 per column)
constexpr uint32_t TMEM_COLS = 512, ACC_COLS_MAX = 128;
constexpr int MAX_NB = 8;
constexpr int NRM_DEPTH = 8;            // tiles in flight between generator and epilogue (power of two)
constexpr int MAX_N = 16;
constexpr int MAX_CHUNKS = 8;           // output chunks of up to 256 rows per tile (sspbo_score_f8.cu)
constexpr int MAX_TRAJ_N = 3;           // waypoint dimension of the trajectory instantiations

// Where the candidates of a scoring call come from.  Generated sets cost no HBM (nor PCIe) byte per candidate:
//   mode 0  explicit rows x (N x n*L, float32 / float64)
//   mode 1  counter-based uniform(0, 1) float32 stream: element e = (global row) * n + axis of the splitmix64 stream of
//           sspbo_fill_uniform (bit-identical values)
//   mode 2  n-D grid in the reference's order (SSPSpace.get_sample_points('grid'), sspspace.py:488-495): row -> per-axis
//           indices (meshgrid('xy'): axis 0 fastest within axis 1, then axes n-1 .. 2), values from the host's np.linspace
//           tables, so the winning row is bit-identical to the reference's sample_points[idx]
struct CandDesc {
    const void* x; int x_f64;
    int mode;
    unsigned long long seed;
    const double* axes;
    unsigned int counts[MAX_N]; int offs[MAX_N];
};

struct Params {
    CandDesc cand; long long N; long long index0;
    const long long* Aq_op; const int* A32_op; const float* Af_op; const float* mscale; double x_scale; float x_absmax;
    // trajectory mode: L waypoints per candidate; per-timestep phase offsets (L x KC_pad/2): Q0.64 turns / float32 radians
    int L; const unsigned long long* tauq; const float* tauf; float nrm_scale, nrm_offset;   // |phi|^2 = nrm_scale * sum(C^2 + S^2) - nrm_offset
    int n_cls;                          // frequency tables staged back to back; waypoint j reads table j % n_cls (one x space per agent)
    int KC_pad, BN, n_chunks, n_kb, nb;
    int na, acc_cols, n_bufs;           // A ring slots; columns per accumulator buffer; buffers alternating (1 or 2)
    int split_issue;                    // BN > 128: next k-block's barrier waits between the two halves of the MMAs
    int merged;                         // BN <= 128: hi|lo tiles of the B slot read as ONE 2 BN-row operand (two MMAs per K step)
    // several chunks per tile (sspbo_score_f8.cu): chunk c covers operand rows [c BN, (c + 1) BN) and the k-blocks
    // kb0[c] .. n_kb - 1 (a triangular factor has no entries before kb0[c]); pos0[c] = position of its first k-block in
    // the tile's k-block sequence, bpt = k-blocks per tile
    int kb0[MAX_CHUNKS], pos0[MAX_CHUNKS], bpt;
    int b_row0;                         // first operand row of chunk 0 in the B images (1 when row 0 of the image holds m')
    const float* mp_op;                 // m' in operand order (KC_pad float32): mu accumulated by the generators (sspbo_score_f8.cu)
    int form;                           // 0: var = 1/beta + |phi|^2/alpha - |V psi|^2 (low rank), 1: var = 1/beta + |R psi|^2 (factor)
    float inv_beta, lam, gamma_t, inv_rscale2, prior_var, flag_below;
    int normalize;                      // trajectory mode: score phi / |phi| (multi-agent trajectories)
    float *mu, *var, *acq;
    unsigned long long* best;
    unsigned int* flag_idx; unsigned long long* stats; unsigned int flag_cap;
    int debug;
};

// Grouped scoring of many independent runs in one launch (BASELINE config C4: thousands of runs, one posterior each, a shared
// candidate set): what differs from run to run.  CTA b scores ALL tiles of runs b, b + gridDim.x, ...; the operand images
// come through the run's own tensor maps (in global memory), the frequency table is read from global memory.
struct alignas(64) RunScore {
    CUtensorMap map_hi, map_lo;
    const float* tab;                   // generator frequency table of the run's space (ctx->Af_gen)
    const float* mscale;
    float inv_beta, lam, gamma_t, inv_rscale2, prior_var, flag_below;
    unsigned long long* best; unsigned int* flag_idx; unsigned long long* stats;
};

// splitmix64 finaliser of the counter-based candidate stream (sspbo_core.cu:k_fill_uniform, oracle counter_uniform)
__host__ __device__ __forceinline__ float counter_uniform_f32(unsigned long long seed, unsigned long long elem) {
    unsigned long long z = elem + (seed << 40) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * 5.9604644775390625e-8f;                                     // 24 bits * 2^-24
}

// per-axis indices of grid row `row` (meshgrid('xy') flattened in C order: shape (N1, N0, N2, ..., N_{n-1})); row < 2^32
template <int NDIM>
__device__ __forceinline__ void grid_indices(const CandDesc& c, unsigned int row, unsigned int (&idx)[NDIM]) {
    unsigned int p = row;
#pragma unroll
    for (int a = NDIM - 1; a >= 2; --a) { const unsigned int q = p / c.counts[a]; idx[a] = p - q * c.counts[a]; p = q; }
    const unsigned int q = p / c.counts[0];
    idx[0] = p - q * c.counts[0];
    if (NDIM >= 2) idx[NDIM >= 2 ? 1 : 0] = q;
}

// coordinates of one candidate (global index gidx = index0 + local row lrow) as T = float / double; rows >= N give zeros
template <int NDIM, typename T>
__device__ __forceinline__ void cand_row(const CandDesc& c, long long gidx, long long lrow, bool valid, T (&out)[NDIM]) {
    if (!valid) {
#pragma unroll
        for (int a = 0; a < NDIM; ++a) out[a] = (T)0;
        return;
    }
    if (c.mode == 0) {
#pragma unroll
        for (int a = 0; a < NDIM; ++a)
            out[a] = c.x_f64 ? (T)((const double*)c.x)[lrow * NDIM + a] : (T)((const float*)c.x)[lrow * NDIM + a];
    } else if (c.mode == 1) {
#pragma unroll
        for (int a = 0; a < NDIM; ++a) out[a] = (T)counter_uniform_f32(c.seed, (unsigned long long)gidx * NDIM + a);
    } else {
        unsigned int idx[NDIM];
        grid_indices<NDIM>(c, (unsigned int)gidx, idx);
#pragma unroll
        for (int a = 0; a < NDIM; ++a) out[a] = (T)__ldg(c.axes + c.offs[a] + idx[a]);
    }
}

// same for a run-time dimension n (float64 pass over flagged candidates), generated modes only
__device__ __forceinline__ void cand_row_rt(const CandDesc& c, int n, long long gidx, double* out) {
    if (c.mode == 1) {
        for (int a = 0; a < n; ++a) out[a] = (double)counter_uniform_f32(c.seed, (unsigned long long)gidx * n + a);
    } else {
        unsigned int p = (unsigned int)gidx;
        for (int a = n - 1; a >= 2; --a) { const unsigned int q = p / c.counts[a]; out[a] = __ldg(c.axes + c.offs[a] + (p - q * c.counts[a])); p = q; }
        const unsigned int q = p / c.counts[0];
        out[0] = __ldg(c.axes + c.offs[0] + (p - q * c.counts[0]));
        if (n >= 2) out[1] = __ldg(c.axes + c.offs[1] + q);
    }
}

// The candidate of a generator thread in the form its phase path wants it (PH 2: float32, PH 1: Q.fx integers, PH 0: Q31.32).
// Returns true when a coordinate is outside the declared |x| bound of the fast phase paths.  `is_nan`: a coordinate is NaN --
// the reference's encode then yields a NaN row, its score is NaN and np.argmax treats NaN as the maximum; the integer phase
// paths would silently turn NaN into 0, so the generators poison psi of such a row instead (gen_half's caller).
template <int NDIM, int PH>
__device__ __forceinline__ bool load_candidate(const Params& p, long long g0, unsigned long long (&x0)[NDIM], int (&xq0)[NDIM],
                                               bool& is_nan) {
    bool oor = false;
    is_nan = false;
    const bool valid = g0 < p.N;
    if constexpr (PH == 2) {
        float v[NDIM];
        cand_row<NDIM, float>(p.cand, p.index0 + g0, g0, valid, v);
#pragma unroll
        for (int a = 0; a < NDIM; ++a) { is_nan |= v[a] != v[a]; oor |= fabsf(v[a]) > p.x_absmax; xq0[a] = __float_as_int(v[a]); }
    } else {
        double v[NDIM];
        cand_row<NDIM, double>(p.cand, p.index0 + g0, g0, valid, v);
        if constexpr (PH == 1) {
#pragma unroll
            for (int a = 0; a < NDIM; ++a) {
                const double s = v[a] * p.x_scale;
                is_nan |= v[a] != v[a];
                oor |= fabs(s) >= 2147483647.0;
                xq0[a] = __double2int_rn(s);
            }
        } else {
#pragma unroll
            for (int a = 0; a < NDIM; ++a) { is_nan |= v[a] != v[a]; x0[a] = (unsigned long long)__double2ll_rn(v[a] * 4294967296.0); }
        }
    }
    return oor;
}

// cos / sin of the NF frequencies kb * 32 + half * NF ... of candidate row g0 (trajectory mode: sums over its L waypoints,
// and sum(C^2 + S^2) added to nrm_acc).  `tables`: the frequency table in shared memory ([block of NF][axis][NF] for the
// 32-bit paths, [frequency][axis] Q31.32 otherwise).
//   PH 0: Q31.32 x Q31.32 (any range), 1: 32-bit fixed point (fa + fx = 55), 2: float32 FMA chain in radians
template <int NDIM, int PH, bool TRAJ>
__device__ __forceinline__ void gen_half(const Params& p, const uint8_t* tables, long long g0, int kb, int half,
                                         unsigned long long (&x0)[NDIM], const int (&xq0)[NDIM], float (&cs)[NF], float (&sn)[NF],
                                         float& nrm_acc) {
    const long long N = p.N;
    const long long* Aq_s = (const long long*)tables;
    if constexpr (TRAJ) {
        // sum over the L waypoints of exp(i (A_f . x_j + tau_jf)); waypoints are re-read per k-block (L1 / L2)
#pragma unroll
        for (int i = 0; i < NF; ++i) { cs[i] = 0.f; sn[i] = 0.f; }
        const int nf = p.KC_pad >> 1, fbase = kb * 32 + half * NF;
        const void* xp = p.cand.x; const int x_f64 = p.cand.x_f64;
        for (int j = 0, cl = 0; j < p.L; ++j) {
            const int fcls = cl * nf + fbase;                 // first frequency of this block in the waypoint's table
            if (++cl == p.n_cls) cl = 0;
            if constexpr (PH == 2) {
                float xf[NDIM];
#pragma unroll
                for (int a = 0; a < NDIM; ++a) {
                    const size_t o = ((size_t)g0 * p.L + j) * NDIM + a;
                    xf[a] = (g0 < N) ? (x_f64 ? (float)__ldg((const double*)xp + o) : __ldg((const float*)xp + o)) : 0.f;
                }
                if (kb == 0 && half == 0) {                   // once per row: waypoints outside the declared |x| bound
                    bool oor = false;
#pragma unroll
                    for (int a = 0; a < NDIM; ++a) oor |= !(fabsf(xf[a]) <= p.x_absmax);
                    if (oor) atomicAdd(p.stats + SSPBO_STAT_OOR, 1ull);
                }
                const float* Afblk = (const float*)tables + (size_t)fcls * NDIM;       // [axis][NF]
                const float* tf = p.tauf + (size_t)j * nf + fbase;
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                    float t = __ldg(tf + i);
#pragma unroll
                    for (int a = 0; a < NDIM; ++a) t = fmaf(Afblk[a * NF + i], xf[a], t);
                    float s_, c_;
                    __sincosf(t, &s_, &c_);
                    cs[i] += c_; sn[i] += s_;
                }
            } else {
#pragma unroll
                for (int a = 0; a < NDIM; ++a) {
                    const size_t o = ((size_t)g0 * p.L + j) * NDIM + a;
                    double v0 = 0.0;
                    if (g0 < N) v0 = x_f64 ? __ldg((const double*)xp + o) : (double)__ldg((const float*)xp + o);
                    x0[a] = (unsigned long long)__double2ll_rn(v0 * 4294967296.0);
                }
                const unsigned long long* Ablk = (const unsigned long long*)Aq_s + (size_t)fcls * NDIM;
                const unsigned long long* tq = p.tauq + (size_t)j * nf + fbase;
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                    unsigned long long t = __ldg(tq + i);
#pragma unroll
                    for (int a = 0; a < NDIM; ++a) t += Ablk[i * NDIM + a] * x0[a];
                    const float a0 = (float)(int)(t >> 32) * 1.4629180792671596e-9f;   // 2 pi / 2^32
                    cs[i] += __cosf(a0); sn[i] += __sinf(a0);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) nrm_acc = fmaf(cs[i], cs[i], fmaf(sn[i], sn[i], nrm_acc));
    } else if (SSPBO_DBG_BITS(p) & 2) {
#pragma unroll
        for (int i = 0; i < NF; ++i) { cs[i] = 0.25f * (float)(i + kb); sn[i] = 0.125f * (float)kb; }
    } else if constexpr (PH == 2) {
        const float* Afblk = (const float*)tables + (size_t)(kb * 32 + half * NF) * NDIM;   // [axis][NF], radians per unit
        // two frequencies per FFMA2: the chain costs NDIM / 2 issue slots per frequency instead of NDIM
        const f32x2* A2 = (const f32x2*)Afblk;
        f32x2 t2[NF / 2];
        {
            const f32x2 xx = f2_pack(__int_as_float(xq0[0]), __int_as_float(xq0[0]));
#pragma unroll
            for (int i = 0; i < NF / 2; ++i) t2[i] = f2_mul(A2[i], xx);
        }
#pragma unroll
        for (int a = 1; a < NDIM; ++a) {
            const f32x2 xx = f2_pack(__int_as_float(xq0[a]), __int_as_float(xq0[a]));
#pragma unroll
            for (int i = 0; i < NF / 2; ++i) t2[i] = f2_fma(A2[a * (NF / 2) + i], xx, t2[i]);
        }
#pragma unroll
        for (int i = 0; i < NF / 2; ++i) {
            float t0, t1;
            f2_unpack(t2[i], t0, t1);
            __sincosf(t0, &sn[2 * i], &cs[2 * i]);                          // one RZ multiply by 1/2pi, then MUFU.SIN / MUFU.COS (the MUFU reduces the range itself)
            __sincosf(t1, &sn[2 * i + 1], &cs[2 * i + 1]);
        }
    } else if constexpr (PH == 1) {
        const int* A32blk = (const int*)tables + (size_t)(kb * 32 + half * NF) * NDIM;   // [axis][NF]
        long long t[NF];                                      // phase in turns at bit 55
#pragma unroll
        for (int i = 0; i < NF; ++i) t[i] = (long long)A32blk[i] * (long long)xq0[0];
#pragma unroll
        for (int a = 1; a < NDIM; ++a) {
#pragma unroll
            for (int i = 0; i < NF; ++i) t[i] += (long long)A32blk[a * NF + i] * (long long)xq0[a];   // IMAD.WIDE
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            // bits [32, 55) = top 23 bits of the phase mod 1; flipping the top one adds 1/2 turn, and with
            // the exponent of 1.0 they are the mantissa of f in [1, 2): (f - 1.5) 2 pi is the phase in
            // radians in [-pi, pi).  One LOP3: (hi & 0x7fffff) ^ 0x3fc00000.
            const float f0 = __uint_as_float(lop3_and_xor((unsigned int)((unsigned long long)t[i] >> 32), 0x7fffffu, 0x3fc00000u));
            const float a0 = fmaf(f0, 6.283185307179586f, -9.42477796076938f);
            cs[i] = __cosf(a0); sn[i] = __sinf(a0);
        }
    } else {
        const unsigned long long* Ablk = (const unsigned long long*)Aq_s + (size_t)(kb * 32 + half * NF) * NDIM;
        unsigned long long t[NF];                             // phase in turns, Q0.64, wraps modulo 1
#pragma unroll
        for (int i = 0; i < NF; ++i) t[i] = 0ull;
#pragma unroll
        for (int a = 0; a < NDIM; ++a) {
#pragma unroll
            for (int i = 0; i < NF; ++i) t[i] += Ablk[i * NDIM + a] * x0[a];
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            // top 32 bits as a signed fraction of a turn in [-1/2, 1/2) -> radians in [-pi, pi)
            const float a0 = (float)(int)(t[i] >> 32) * 1.4629180792671596e-9f;   // 2 pi / 2^32
            cs[i] = __cosf(a0); sn[i] = __sinf(a0);
        }
    }
}

// frequency table -> shared memory in the layout gen_half reads
template <int NDIM, int PH>
__device__ __forceinline__ void load_tables(const Params& p, uint8_t* tables, int nthreads) {
    if constexpr (PH == 1 || PH == 2) {
        // axis-major inside every block of NF frequencies: one LDS.128 feeds four independent multiply-add chains
        int* A32_s = (int*)tables;
        const int* src = PH == 1 ? p.A32_op : (const int*)p.Af_op;
        for (int i = threadIdx.x; i < p.n_cls * (p.KC_pad / 2) * NDIM; i += nthreads) {
            const int f = i / NDIM, a = i - f * NDIM;         // (KC_pad / 2 is a multiple of NF: blocks never straddle tables)
            A32_s[((f / NF) * NDIM + a) * NF + (f % NF)] = src[i];
        }
    } else {
        long long* Aq_s = (long long*)tables;
        for (int i = threadIdx.x; i < p.n_cls * (p.KC_pad / 2) * NDIM; i += nthreads) Aq_s[i] = p.Aq_op[i];
    }
}

// Upper bound of the true score of a candidate the tensor pass flags (its variance is too small for the pass's own accuracy):
// the approximate mean plus the acquisition at an inflated variance.  The error of the pass on the variance is ~2e-5 of the
// prior variance (float32 accumulation inside the tensor core; e4m3 corrections in factor form: ~3e-5); 1e-4 of it is added.
// The float64 pass skips a flagged candidate whose bound, with a further 1e-3 relative margin, stays below the best key found
// so far: a candidate close to the observations has little variance left and rarely competes for the arg-max, and its float64
// re-score costs as much as ~36 tensor-pass candidates per operand row.  Stored behind the index list (flag_idx[cap + slot]).
__device__ __forceinline__ float flagged_score_bound(float mu, float var, float prior, float lam, float gamma_t) {
    if (!(lam >= 0.f)) return __int_as_float(0x7f800000);                 // (a negative weight would need a lower bound: no pruning)
    const float v = fmaxf(var, 0.f) + 1e-4f * prior;
    return mu + lam * v / (sqrtf(v + gamma_t) + sqrtf(gamma_t));
}

// acquisition of one scored candidate: packs the key, writes the optional outputs
__device__ __forceinline__ unsigned long long finish_candidate(const Params& p, long long gi, float mu, float var) {
    const float acq = p.lam * var / (sqrtf(var + p.gamma_t) + sqrtf(p.gamma_t));   // ssp_agent.py:136, cancellation-free
    if (p.mu) p.mu[gi] = mu;
    if (p.var) p.var[gi] = var;
    if (p.acq) p.acq[gi] = acq;
    return sspbo_pack(mu + acq, (unsigned long long)(p.index0 + gi));
}

}  // namespace sspbo_lr

// host side shared by the two engines (sspbo_score_lr.cu)
void sspbo_lr_fill_common(const sspbo_ctx* ctx, const sspbo_lr::CandDesc& cand, int64_t N, int64_t index0, float lam, float gamma_t,
                          float* mu, float* var, float* acq, unsigned long long* best, int ph, sspbo_lr::Params* p);
int sspbo_lr_phase_path(const sspbo_ctx* ctx);
// split-fp16 operands, one chunk (rank <= 255); fp16 + e4m3 operands, form 0 = low rank (rank <= 511), 1 = triangular factor
int sspbo_score_lr_launch(sspbo_ctx* ctx, const sspbo_lr::CandDesc& cand, int64_t N, int64_t index0, float lam, float gamma_t,
                          float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st);
int sspbo_score_lr_runs_prepare(sspbo_ctx* const* ctxs, int R, int64_t N, const double* lam, const double* gamma_t,
                                unsigned long long* keys_dev, sspbo_lr::RunScore* runs_host);
int sspbo_score_lr_runs_launch(sspbo_ctx* const* ctxs, int R, const sspbo_lr::CandDesc& cand, int64_t N, const sspbo_lr::RunScore* runs_dev,
                               cudaStream_t st);
int sspbo_score_f8_launch(sspbo_ctx* ctx, const sspbo_lr::CandDesc& cand, int64_t N, int64_t index0, float lam, float gamma_t,
                          float* mu, float* var, float* acq, unsigned long long* best, int form, cudaStream_t st);
