// Internal definitions shared by the translation units of libsspbo.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/sspbo.h"

struct sspbo_ctx {
    int device = 0;
    int sm_count = 148;
    int cc_major = 0;                // compute capability major (the tcgen05 engines need 10)
    mutable char err[512] = {0};
    int64_t launches = 0;
    int last_engine = 0;
    // kernel timing on request (sspbo_set_timing): CUDA events on the launching stream around the scorer kernel and around
    // the float64 pass over flagged candidates of every sspbo_score call; read and summed by sspbo_timing_read
    bool timing = false;
    void* timing_events = nullptr;   // std::vector<cudaEvent_t>, three per call: before scorer, after scorer, after flagged pass
    int timing_used = 0;

    // ---- space ----
    int K = 0, n = 0, d = 0, L = 1;
    int n_cls = 1;                   // frequency tables (one x space per agent of a multi-agent encoder): waypoint j reads table j % n_cls
    int dp = 0;                      // d padded to a multiple of 64 (operand leading dimension)
    double* A_turns = nullptr;       // n_cls x K x n : A+ * inv_ls / (2 pi)  -> theta in turns (float64)
    float*  A_turns_f32 = nullptr;   // same, float32
    double* tau_turns = nullptr;     // L x K : trajectory time phases in turns (nullptr when L == 1)
    float*  tau_turns_f32 = nullptr;
    double2* twiddle = nullptr;      // d entries (cos, sin)(2 pi i / d), float64
    bool umma_range_ok = true;       // every |A_turns| < 2^31 (fixed-point phase of the tcgen05 scorer)
    double max_turns = 0.0;          // host bound on |theta| in turns per unit |x|_inf (sum_a |A_turns[k][a]|, max over k)

    // ---- BLR (phi-space, float64) ----
    bool blr_ready = false, has_beta = false;
    bool has_factor = false;         // psi-space factor maintained (needs a space: d = 2K+1)
    double alpha = 0.01, beta = 0.0;
    double *S = nullptr, *Sinv = nullptr, *b = nullptr, *m = nullptr;
    // ---- scoring state (psi-space) ----
    // Sp = P (B^T S B) P^T : psi-space covariance with components permuted into operand-column order (perm below).
    // R holds the LOWER Cholesky factor L of Sp (row-major, only j <= i meaningful); the scoring factor is the upper
    // triangular L^T:  psi^T Sp psi = |L^T psi|^2, and output row j of the scorer is column j of L.
    double* Sp = nullptr;            // d x d float64
    double* R = nullptr;             // d x d float64, L (lower)
    int* perm = nullptr;             // d : psi natural index of the component with rank r (operand-column order)
    int* inv_perm = nullptr;         // d : rank of psi natural index
    int* col_of_rank = nullptr;      // d : operand column of rank r
    short nz_rows[32] = {0};         // components whose operand column is < 64 (kb + 1), per k-block
    int kb_start[8] = {0};           // first k-block with non-zero factor entries, per output chunk (triangular skip)
    bool factor_dirty = true;        // L must be recomputed from Sp
    double* mp = nullptr;            // m' = B^T m (d)
    float*  Rt_f32 = nullptr;        // dp x dp float32, Rt[k*dp + j] = R[j][k]      (SIMT scorer)
    float*  mp_f32 = nullptr;        // dp
    // tcgen05 scorer operands, "operand order" (see sspbo_umma_dims): NJ_pad x KC_pad fp16, K-major.
    //   column c -> block b = c/64, i = c%64, frequency f = 32 b + (i%32) (f = 0 is the DC term);
    //   i < 32 : cos(theta_f) column, i >= 32 : sin(theta_f) column; f > K and sin(DC) columns are zero.
    __half* Rh = nullptr;            // hi(R[j][psi_index(c)] * r_scale)
    __half* Rl = nullptr;            // lo part (value - hi)
    float*  mp_op = nullptr;         // m' in operand order (KC_pad)
    float* tauf_op = nullptr;        // L x (KC_pad/2) float32 radians in [-pi, pi): timestep phase of frequency f (float32 phase path)
    unsigned long long* tauq_op = nullptr;   // L x (KC_pad/2), Q0.64 turns: timestep phase of frequency f (trajectory mode)
    long long* Aq_op = nullptr;      // (KC_pad/2) x n, Q31.32 fixed point: A_turns row of frequency f (row 0 and f > K are zero)
    int KC_pad = 0, BN = 0, n_chunks = 0, NJ_pad = 0;
    double r_scale = 1.0;
    bool operands_dirty = true;      // Rt_f32 / Rh / Rl / mp_f32 need a refresh from R, m
    // ---- low-rank scoring state: Sp = Sp0 - V^T V with one row of V per observation (rank = #observations) ----
    //   psi^T Sp psi = psi^T Sp0 psi - |V psi|^2 ;  row r of V = y sqrt(beta / (1 + beta psi.y)), y = Sp psi (before the update)
    // Split-fp16 image in operand order: (SSPBO_LR_MAX + 1) x KC_pad, K-major; row 0 = m' (own scale, see mscale), row j + 1 =
    // observation j, rows beyond the rank are zero.
    __half* Vh = nullptr;
    __half* Vl = nullptr;
    float* mscale = nullptr;         // device scalar: 1 / (power-of-two scale of row 0 = m')
    double* Vd = nullptr;            // float64 rows of V in rank order (SSPBO_LR_MAX x d): exact pass over flagged candidates
    int lr_rank = 0;
    double export_rank = -1.0;       // staging scalar of sspbo_blr_export
    bool lr_valid = false;           // every update since blr_init appended its row (false after import / rank overflow)
    // Low-rank updates (k_blr_update_lr, sspbo_update.cu) touch only the rows of V, b_psi and m'; the d x d state and the
    // phi-space vectors follow when a consumer asks (sspbo_ensure_*): Sp lacks the rows [sp_applied, lr_rank) of V
    // (Sp -= row^T row), S the rows [s_applied, lr_rank) (S -= u u^T, u = B D^-1 row), S_inv the sinv_pending observations
    // whose psi sits in psi_ring (S_inv += beta phi phi^T, phi = B psi), b and m are stale when bm_stale (b = B b_psi,
    // m = B D^-1 m').  The eager kernels keep all of it current and leave b_psi behind (bpsi_stale: b_psi = D^-1 B^T b).
    int sp_applied = 0, s_applied = 0;
    double* psi_ring = nullptr;      // SSPBO_SINV_RING x d, natural order
    int sinv_pending = 0;
    double* bpsi = nullptr;          // d, rank order: beta sum_t t psi_t
    bool bm_stale = false, bpsi_stale = false;
    // checkpoint of the low-rank posterior (sspbo_blr_checkpoint / _rollback): b_psi and m' (2 d doubles), m' in operand order
    // (KC_pad floats), row 0 of the operand images (2 KC_pad halves), the scale of that row; and the bookkeeping at that moment
    double* ck_vec = nullptr; float* ck_mp_op = nullptr; __half* ck_row0 = nullptr; float* ck_mscale = nullptr;
    bool ck_valid = false; int ck_rank = 0, ck_sinv_pending = 0;
    double* lr_gh = nullptr;         // scratch of k_blr_update_lr: 2 SSPBO_LR_MAX + 32 doubles
    double* syn_rows = nullptr;      // scratch: SSPBO_SINV_RING x d rows synthesised into phi-space
    bool lr_disabled = false;        // policy: too many candidates needed the exact pass, use the full factor instead
    bool f8_disabled = false;        // policy: the factor form's own flag list overflowed, AUTO uses the engines that never flag
    bool p32_disabled = false;       // policy: candidates outside the declared |x| bound were seen
    // 32-bit fixed-point phase tables (fast generator of the tcgen05 scorer), valid when p32_ok
    int* A32_op = nullptr;           // (KC_pad/2) x n, Q(31-fa).fa turns per unit
    bool p32_ok = false;
    float* Af_gen = nullptr;         // Af_op in the generators' layout ([block of 16 frequencies][axis][16]), for the grouped scorer
    float* Af_op = nullptr;          // (KC_pad/2) x n float32, radians per unit (float32 phase path)
    bool pf32_disabled = false;      // policy override: never use the float32 chain
    bool pf32_ok = false;            // declared bounds keep |theta| <= 16 pi: float32 FMA chain is accurate to < 1e-5 rad
    int p32_fa = 0, p32_fx = 0;      // fractional bits of A32_op and of the candidate coordinates
    double x_absmax = 0.0;           // declared bound on |x| (0 = unknown)
    double a_absmax = 0.0;           // max |A_turns|
    std::vector<double> host_A_turns;            // host copy of A_turns (K x n) for rebuilding the tables
    // exact float64 scorer (flagged candidates of the low-rank pass, tiny candidate sets)
    unsigned int* flag_idx = nullptr;            // SSPBO_FLAG_CAP local row indices, then SSPBO_FLAG_CAP float32 score bounds
    unsigned long long* stats = nullptr;         // device counters, see SSPBO_STAT_*
    double* ex_part = nullptr;                   // partial quadratic forms / means of the exact scorer
    size_t ex_part_cap = 0;
    // scratch d-vectors (float64)
    double *v = nullptr, *psi = nullptr, *y = nullptr, *z = nullptr, *phi_tmp = nullptr;
    double* scal = nullptr;          // a few device scalars
    // tensor-core encoder: split-fp16 inverse-DFT basis in operand order (NJ_pad x KC_pad), built lazily
    __half* Bh_enc = nullptr; __half* Bl_enc = nullptr; double enc_scale = 1.0; int enc_K = 0;
    // waypoint sums (L > 1): value of the DC component (L, or sum of the +-1 DC signs of bound tags) and whether phi is
    // divided by its norm (multi-agent trajectories, agents/ssp_multi_agent.py:166-186)
    double dc = 1.0; bool normalize = false;
    // staging of sspbo_blr_update_x: pinned x, device x, device phi rows, event guarding the pinned buffer
    void* obs_pin = nullptr; void* obs_x = nullptr; size_t obs_bytes = 0; double* obs_phi = nullptr; size_t obs_phi_cap = 0; cudaEvent_t obs_event = nullptr;
    double* obs_out = nullptr;       // device: predictive (mu, var) at the observed point before the update (sspbo_observe)
    bool want_obs = false;
    bool unfused_update = false;     // posterior updates go through the separate kernels of sspbo_core.cu
    int fused_update_ok = -1;        // cooperative single-launch posterior update usable on this device (-1: not probed)
    // pinned host staging and device staging of the small synchronous transfers of a BO step (sspbo_score_finish, sspbo_eval_host)
    void* pin = nullptr; size_t pin_bytes = 0; void* eval_dev = nullptr; size_t eval_dev_bytes = 0;
    // workspace of the phi-space acquisition ascent (sspbo_ascent.cu): Y = S Phi and per-CTA partial sums, double-buffered
    double* aa_work = nullptr; size_t aa_cap = 0;
    // opaque state of the tcgen05 engines (tensor maps etc.)
    void* umma = nullptr;
    void* lr_state = nullptr;
    void* f8_state = nullptr;
    void* dist_state = nullptr;      // peer-memory key exchange (sspbo_dist.cu)
    // fp16 + e4m3 operand format (sspbo_score_f8.cu), built lazily: V8 = e4m3 companion of (Vh, Vl), rows 1 .. v8_rank
    // current; (Fh, F8) = images of [m'; L^T] for the factor form
    uint8_t* V8 = nullptr; int v8_rank = 0; bool v8_row0_dirty = true;
    __half* Fh = nullptr; uint8_t* F8 = nullptr; bool f8_factor_dirty = true;
    // parameter block + points of sspbo_update_batch's single launch (owned by the first context of the batch)
    void* batch_pin = nullptr; void* batch_dev = nullptr; size_t batch_bytes = 0; cudaEvent_t batch_event = nullptr;
    // ... and of sspbo_score_batch's two grouped launches (RunScore[R] | ExactArgs[R] | status[R])
    void* sbatch_pin = nullptr; void* sbatch_dev = nullptr; size_t sbatch_bytes = 0; cudaEvent_t sbatch_event = nullptr;
    int batch_status_valid = 0;      // runs whose status words the grouped float64 pass left in sbatch_dev
    double* exact_out64 = nullptr;   // set by sspbo_eval_host around its exact-engine launch: 3 x 64 float64 outputs
    double lr_flag_frac = 0.0;       // flagged / scored of the last read-back of a low-rank pass (cost model of ENGINE_AUTO)
    int operand_format = -1;         // policy: -1 = by rank, 0 = split fp16 only (rank <= 255), 1 / 2 = fp16 + e4m3 whenever usable (single CTAs / CTA pairs)
    std::vector<int> host_col_of_rank;           // host copy of col_of_rank (chunk geometry of the factor form)
    // scratch rows for generated candidate sets scored by an engine that reads explicit rows
    void* gen_rows = nullptr; size_t gen_rows_bytes = 0;
};

#define SSPBO_FAIL(ctx, code, ...)                                         \
    do {                                                                   \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);             \
        return (code);                                                     \
    } while (0)

#define SSPBO_CUDA(ctx, expr)                                                              \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s -> %s", __FILE__, __LINE__, \
                     #expr, cudaGetErrorString(e__));                                      \
            (void)cudaGetLastError(); /* reported here: must not resurface at a later launch check */ \
            return SSPBO_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

#define SSPBO_LAUNCH_CHECK(ctx)                      \
    do {                                             \
        (ctx)->launches++;                           \
        SSPBO_CUDA(ctx, cudaGetLastError());         \
    } while (0)

constexpr int SSPBO_LR_MAX = 511;               // largest rank scored in low-rank form (m' + 511 rows = two 256-row MMA operands)
// The tensor core accumulates in float32 with truncation: |V psi|^2 carries a relative error of ~2e-5 (measured), i.e.
// 2e-5 / alpha absolute on the variance.  Candidates whose variance is below this fraction of the prior variance
// 1/alpha would exceed the 1e-4 tolerance after the subtraction and are re-scored by the exact float64 engine.
constexpr double SSPBO_LR_FLAG_FRACTION = 0.3;
// operand rows (BN, a multiple of 16) from which the low-rank form runs with fp16 + e4m3 operands (sspbo_score_f8.cu): below,
// the kernel is bound by generating psi and the split-fp16 format (cheaper to generate, two MMAs per K step through the
// merged hi|lo operand up to 128 rows) is faster: measured on C3, rank 100 / 127: 1.57 / 1.47e9 against 1.33 / 1.35e9
// candidates/s; from 144 rows on the split format needs three MMAs per K step and loses (rank 128: 1.27 against 1.32e9)
constexpr int SSPBO_F8_MIN_BN = 144;
// Factor form with fp16 + e4m3 operands: the e4m3 corrections leave ~2^-15 of the LARGEST products in |R psi|^2, which
// reaches 1e-4 of the sum for candidates almost on top of an observation; candidates whose variance is below this fraction
// of the prior variance |phi|^2 / alpha are re-scored by the float64 pass (measured on B200: 1.3e-4 at 0.3 % of the prior).
constexpr double SSPBO_F8_FLAG_FRACTION = 0.1;
constexpr int SSPBO_CLUSTER_UPDATE_MAX_D = 320;  // single observations up to this dimension take the clustered one-launch update
constexpr unsigned SSPBO_FLAG_CAP = 1u << 17;    // flagged candidates per sspbo_score call that get the exact pass
// device counters (unsigned long long each)
enum { SSPBO_STAT_CUR = 0,       // flagged candidates of the call in flight (reset by its exact pass)
       SSPBO_STAT_FLAGGED = 1,   // candidates re-scored exactly since the last reset
       SSPBO_STAT_OVERFLOW = 2,  // flagged candidates beyond SSPBO_FLAG_CAP (kept their low-rank score)
       SSPBO_STAT_OOR = 3,       // candidates outside the declared |x| bound of the 32-bit phase path
       SSPBO_STAT_SCORED = 4,    // candidates scored by the low-rank pass since the last reset
       SSPBO_STAT_RESCORED = 5,  // flagged candidates the float64 pass really re-scored (the others could not win the arg-max)
       SSPBO_STAT_PEER = 6,      // multi-GPU key exchange (sspbo_dist.cu): bit 0 = some rank asks for a rerun, bit 1 = a peer never arrived
       SSPBO_STAT_COUNT = 8 };

// Timing experiments (tools/kbench.py): built only with -DSSPBO_EXPERIMENTS (`make EXPERIMENTS=1`).  SSPBO_DEBUG bits:
// 1 = no MMAs issued, 2 = generators skip the phase / sincos math, 4 = skip the float64 pass over flagged candidates,
// 8 = force the Q31.32 phase path.  In product builds the switches compile to constants.
#ifdef SSPBO_EXPERIMENTS
#include <stdlib.h>
#define SSPBO_DBG_BITS(p) ((p).debug)
static inline int sspbo_debug_env() { const char* e = getenv("SSPBO_DEBUG"); return e ? atoi(e) : 0; }
#else
#define SSPBO_DBG_BITS(p) 0
static inline int sspbo_debug_env() { return 0; }
#endif

static inline int sspbo_pad64(int x) { return (x + 63) / 64 * 64; }

// Operand geometry of the tcgen05 scorer for a space with K positive frequencies (d = 2K+1):
//   KC_pad   contraction length: (K+1) frequencies in blocks of 32 -> 64 columns (cos|sin) per block
//   n_chunks / BN  the d output rows j are processed in n_chunks chunks of BN rows (BN % 16 == 0, BN <= 256)
static inline void sspbo_umma_dims(int K, int* KC_pad, int* BN, int* n_chunks, int* NJ_pad) {
    int d = 2 * K + 1;
    int kb = (K + 1 + 31) / 32;
    int nc = (d + 255) / 256;
    int bn = ((d + nc - 1) / nc + 15) / 16 * 16;
    *KC_pad = kb * 64; *BN = bn; *n_chunks = nc; *NJ_pad = nc * bn;
}

// ---- packed (score, index) key: max(key) == max score, lowest index on ties; NaN ranks highest ----
__host__ __device__ inline unsigned int sspbo_orderable(float f) {
#ifdef __CUDA_ARCH__
    unsigned int u = __float_as_uint(f);
#else
    unsigned int u; memcpy(&u, &f, 4);
#endif
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;      // NaN counts as the maximum (np.argmax)
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline unsigned long long sspbo_pack(float score, unsigned long long idx) {
    return ((unsigned long long)sspbo_orderable(score) << 32) | (unsigned long long)(0xffffffffu - (unsigned int)idx);
}

// engines implemented in other translation units
int sspbo_score_simt_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, int64_t index0, float lam,
                            float gamma_t, float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st);
int sspbo_score_umma_supported(const sspbo_ctx* ctx);
int sspbo_score_umma_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, int64_t index0, float lam,
                            float gamma_t, float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st);
void sspbo_umma_release(sspbo_ctx* ctx);
int sspbo_encode_umma_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, float* out, int pitch, cudaStream_t st);
// one observation of the posterior update in one cooperative launch (sspbo_update.cu); SSPBO_EUNSUPPORTED -> unfused kernels
int sspbo_blr_update_runs(sspbo_ctx* const* ctxs, int R, const double* x_host, const double* ts_host, cudaStream_t st, int* failed,
                          double* obs_out = nullptr);
bool sspbo_fused_update_available(sspbo_ctx* ctx);
int sspbo_blr_update_fused(sspbo_ctx* ctx, const double* phi, double t, bool finalize, __half* vh_row, __half* vl_row,
                           double* vd_row, double* obs_out, cudaStream_t st);
// low-rank engines with the A operand in tensor memory: prototypes in sspbo_lr_common.cuh (they take a candidate descriptor)
int sspbo_score_lr_supported(const sspbo_ctx* ctx);
void sspbo_lr_release(sspbo_ctx* ctx);
void sspbo_f8_release(sspbo_ctx* ctx);
int sspbo_score_f8_supported(const sspbo_ctx* ctx, int form);
double sspbo_f8_cost(const sspbo_ctx* ctx, int form, int64_t N);
// L = chol(Sp) into ctx->R when the psi-space covariance changed since the last factorisation
int sspbo_refresh_factor(sspbo_ctx* ctx, cudaStream_t st);
void sspbo_dist_release(sspbo_ctx* ctx);
constexpr int SSPBO_SINV_RING = 32;
// what the low-rank updates left out, folded in (no-ops when nothing is pending): Sp; S and m; S_inv; b; everything
int sspbo_ensure_sp(sspbo_ctx* ctx, cudaStream_t st);
int sspbo_ensure_s(sspbo_ctx* ctx, cudaStream_t st);
int sspbo_ensure_sinv(sspbo_ctx* ctx, cudaStream_t st);
int sspbo_ensure_bm(sspbo_ctx* ctx, cudaStream_t st);
int sspbo_ensure_all(sspbo_ctx* ctx, cudaStream_t st);
// before a low-rank update: scratch buffers, b_psi current, *ring_row = where the kernel keeps psi for S_inv
int sspbo_lr_update_prepare(sspbo_ctx* ctx, cudaStream_t st, double** ring_row);
// after an eager update (which keeps the d x d state current itself)
void sspbo_eager_update_done(sspbo_ctx* ctx);
int sspbo_blr_update_lr(sspbo_ctx* const* ctxs, int R, const double* x_host, const double* const* phi_dev, const double* ts_host,
                        cudaStream_t st, int* failed, double* obs_out);
// CUtensorMap (passed as void*) of a K-major fp16 operand [rows x KC_pad], box 64 columns x BN rows, 128-byte swizzle
int sspbo_make_map_f16(sspbo_ctx* ctx, void* map, const void* base, int KC_pad, int rows, int BN);
int sspbo_refresh_operands(sspbo_ctx* ctx, cudaStream_t st);
// exact float64 scorer: rows `list[0..*count)` of x (list == nullptr: rows 0..N-1)
int sspbo_score_exact_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, bool from_flag_list, int64_t index0,
                             double lam, double gamma_t, float* mu, float* var, float* acq, unsigned long long* best,
                             cudaStream_t st);

// ---- device helpers shared by the scorers: theta (in turns) of frequency k for candidate row x ----
// fast path: float32 FMA chain (valid when the host bound max_turns*|x| keeps the error below 2^-20 turns)
// precise path: float64
template <typename XT>
__device__ __forceinline__ float sspbo_theta_turns(const XT* __restrict__ x, const float* __restrict__ Af,
                                                   const double* __restrict__ Ad, int n, bool precise) {
    if (!precise) {
        float t = 0.f;
        for (int a = 0; a < n; ++a) t = fmaf(Af[a], (float)x[a], t);
        return t - rintf(t);
    } else {
        double t = 0.0;
        for (int a = 0; a < n; ++a) t = fma(Ad[a], (double)x[a], t);
        return (float)(t - rint(t));
    }
}
