/*
 * sspbo.h - C ABI of the B200-native SSP-BO hot path.
 *
 * The reference (ctn-waterloo/ssp-bayesopt) is pure Python/numpy and has no FFI of
 * its own; its boundary for this path is the Python method surface.  Each entry
 * point below names the reference method (file:line under ssp_bayes_opt/) whose
 * arithmetic it replaces; `ssp_bayesopt_b200/` re-hosts those methods on top of
 * this ABI with ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types cross the ABI.
 *   - every function returns 0 on success, a negative SSPBO_E* code on failure;
 *     sspbo_last_error(ctx) gives the message.  No exception crosses the ABI.
 *   - pointers named *_dev are caller-owned DEVICE pointers on ctx's device,
 *     row-major; *_host are host pointers.  `stream` is a cudaStream_t passed as
 *     void* (NULL = legacy default stream).  All device work is stream-ordered
 *     and asynchronous unless the function returns a host value.
 *   - one ctx per (space, posterior) on one device - any number of them may share a device (independent runs,
 *     sspbo_score_batch); a ctx is not thread-safe; distinct ctxs are independent.
 *   - psi-space: d = 2K+1, psi = [DC, cos(theta_1..K), sin(theta_1..K)],
 *     phi = B psi with B the real inverse-DFT matrix.
 */
#ifndef SSPBO_H
#define SSPBO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSPBO_OK            0
#define SSPBO_EINVAL       -1   /* bad argument / shape                      */
#define SSPBO_ECUDA        -2   /* CUDA runtime error (message has details) */
#define SSPBO_ESTATE       -3   /* call order (space/blr not initialised)   */
#define SSPBO_EUNSUPPORTED -4   /* shape outside what the kernels cover     */

#define SSPBO_F32 0
#define SSPBO_F64 1

/* scoring engines for sspbo_score */
#define SSPBO_ENGINE_AUTO 0     /* N <= 64: EXACT; N >= 1024: the tensor-memory tcgen05 engines, low-rank form (UMMA_LR) or
                                   factor form (UMMA_F8), whichever costs less per tile: operand rows x k-blocks, plus, for
                                   the low-rank form, the float64 pass over the share of candidates it flagged at the last
                                   read-back (low rank needs rank <= 511 and at most 1 % of flagged candidates; a factor form
                                   whose own flag list overflows hands over to UMMA, which never flags); else SIMT */
#define SSPBO_ENGINE_SIMT 1     /* fp32 CUDA-core kernel (any d)                */
#define SSPBO_ENGINE_UMMA 2     /* tcgen05/TMEM split-fp16 kernel, full triangular factor, psi staged in shared memory */
#define SSPBO_ENGINE_EXACT 3    /* float64 quadratic form on the psi-space covariance (small N)  */
#define SSPBO_ENGINE_UMMA_LR 4  /* tcgen05 kernels with psi in tensor memory on the low-rank form S = I/alpha - V^T V
                                   (one row of V per observation, rank <= 511: split-fp16 operands up to 128 operand
                                   rows, fp16 + e4m3 operands beyond, see sspbo_set_operand_format) followed by a float64
                                   pass over the candidates it flags (var < 0.3 / alpha)          */
#define SSPBO_ENGINE_UMMA_F8 5  /* same kernel family on the triangular factor of the psi-space covariance
                                   (var = 1/beta + |R psi|^2, any rank): fp16 + e4m3 operands, d + 1 operand rows in
                                   chunks of <= 256, each chunk skipping the k-blocks before its first non-zero */

/* candidate sets generated inside the scorer (sspbo_score_generated) */
#define SSPBO_CAND_UNIFORM 1    /* counter-based uniform(0,1) float32 stream of sspbo_fill_uniform */
#define SSPBO_CAND_GRID 2       /* grid of sspbo_grid_points (SSPSpace.get_sample_points('grid')) */

typedef struct sspbo_ctx sspbo_ctx;

int  sspbo_abi_version(void);
int  sspbo_ctx_create(int device, sspbo_ctx** out);
void sspbo_ctx_destroy(sspbo_ctx* ctx);
const char* sspbo_last_error(const sspbo_ctx* ctx);

/* ---- SSP space ---------------------------------------------------------------
 * Replaces the state of SSPSpace.__init__/update_lengthscale (sspspace.py:216-252).
 * A_plus_host: the K positive-frequency rows phase_matrix[1:K+1] (K x n, float64);
 * inv_ls_host: 1/length_scale per axis (n).  d = 2K+1 (conjsym, sspspace.py:821-827).
 * Re-setting the space keeps the BLR state only if K is unchanged. */
int sspbo_space_set(sspbo_ctx* ctx, const double* A_plus_host, const double* inv_ls_host, int K, int n);
/* One phase matrix per agent (SSPMultiAgent._set_ssp_space with same_agt_x_space=False or a list of distinct ssp_x_spaces,
 * agents/ssp_multi_agent.py:51-66): n_cls tables of one shape, A_plus_host n_cls x K x n, inv_ls_host n_cls x n.  Waypoint j of
 * a candidate row is encoded with table j % n_cls -- rows laid out (traj_len, n_agents, x_dim) as SSPMultiAgent.encode reshapes
 * them (ssp_multi_agent.py:178) -- by every consumer (encode, the SIMT / low-rank / fp16+e4m3 scorers, the float64 passes);
 * the dense-factor tcgen05 engine and sspbo_encode_f32 report SSPBO_EUNSUPPORTED.  Until sspbo_space_set_waypoints is called
 * (L a multiple of n_cls) a row is n_cls waypoints with no time phases.  1 <= n_cls <= 16. */
int sspbo_space_set_classes(sspbo_ctx* ctx, const double* A_plus_host, const double* inv_ls_host, int K, int n, int n_cls);

/* Trajectory mode (SSPTrajectoryAgent.encode, agents/ssp_traj_agent.py:145-163):
 * candidates are rows of L waypoints (L*n values); the spectrum is the sum over
 * waypoints of exp(i(theta_k(x_j) + tau[j][k])).  tau_host: L x K float64, radians
 * (phase of timestep j at frequency k, i.e. A_t[k]*t_j/ls_t).  L=1 with tau=NULL
 * restores the plain space. */
int sspbo_space_set_traj(sspbo_ctx* ctx, const double* tau_host, int L);
/* Multi-agent trajectories (SSPMultiAgent.encode, agents/ssp_multi_agent.py:166-186, agents sharing one x space):
 * phi = unit( sum_i a_i (*) sum_j T_j (*) phi_x(x_ji) ).  Bound tags a_i are unitary, so in the Fourier domain the whole
 * encoder is again a sum of unit phasors over the L = traj_len * n_agents waypoints (j, i), candidate layout
 * (traj_len, n_agents, x_dim) flattened, with tau[(j, i)][k] = phase of T_j plus phase of a_i at frequency k.
 * dc_sign_host (nullable): +1 / -1 per waypoint, the DC coefficient of its tag (a unitary real vector has DC +-1).
 * normalize != 0: every consumer (encode, score) uses phi / |phi| (ssp_multi_agent.py:186); the dense-factor tcgen05
 * engine and sspbo_encode_f32 do not form |phi| and report SSPBO_EUNSUPPORTED, AUTO never picks them then. */
int sspbo_space_set_waypoints(sspbo_ctx* ctx, const double* tau_host, const int* dc_sign_host, int L, int normalize);

int sspbo_space_dims(const sspbo_ctx* ctx, int* d, int* K, int* n, int* L);

/* Declares a bound on |x| over every candidate coordinate (max |SSPSpace.domain_bounds|).  With it the tcgen05
 * scorer evaluates theta = A x in 32-bit fixed point (one integer multiply-add per axis) when the worst-case
 * operand rounding stays below 2^-20 turns; candidates outside the bound are counted (sspbo_score_stats) and
 * switch the context back to the 64-bit phase path.  0 = unknown (64-bit path). */
int sspbo_space_set_xbound(sspbo_ctx* ctx, double x_absmax);

/* ---- K1: encode ---------------------------------------------------------------
 * SSPSpace.encode (sspspace.py:254-276): phi = Re IFFT(exp(i A (x/l)^T))^T.
 * x_dev: N x (n*L), dtype SSPBO_F32|SSPBO_F64.  phi_out_dev: N x d float64. */
int sspbo_encode(sspbo_ctx* ctx, const void* x_dev, int x_dtype, int64_t N, double* phi_out_dev, void* stream);
/* SSPSpace.encode_and_deriv (sspspace.py:278-306): phi_out_dev N x d float64 and grad_out_dev, n planes of N x d float64,
 * grad[a][s][j] = d phi_j(x_s) / d x_a = Re IDFT_k( i (A_ka / l_a) exp(i A_k . x_s / l) )_j (the reference's expression at
 * :303-304 multiplies a (d, n) by a (d, N) matrix and raises for n != d; this is the derivative its docstring describes,
 * the one SSPSpace.decode('direct-optim') needs).  Plain spaces only. */
int sspbo_encode_deriv(sspbo_ctx* ctx, const void* x_dev, int x_dtype, int64_t N, double* phi_out_dev, double* grad_out_dev,
                       void* stream);
/* Feature map of the reference's random-Fourier-feature baseline agent (agents/rff_agent.py:153-156):
 * phi_out_dev[s][j] = scale * cos(W_j . x_s + B_j), W_dev m x n, B_dev m float64, x_dev N x n (F32 | F64), phi_out_dev N x m
 * float64.  The BLR over these features is a raw-feature context (sspbo_blr_init_dim / sspbo_blr_update / sspbo_blr_predict);
 * ctx only names the device. */
int sspbo_rff_encode(sspbo_ctx* ctx, const double* W_dev, const double* B_dev, int m, int n, double scale, const void* x_dev,
                     int x_dtype, int64_t N, double* phi_out_dev, void* stream);

/* K1 on the tensor cores, float32 result: phi = psi B^T with the split-fp16 inverse-DFT basis as the MMA operand; psi is
 * generated on chip as in the scorer.  Error ~1e-6 of max |phi| (inside the 1e-4 the path is specified to), ~50x the
 * rate of the float64 kernel.  phi_out_dev: N x sspbo_encode_pitch(ctx) float32 (row pitch = d rounded up to the MMA
 * tiling, columns >= d are zero). */
int sspbo_encode_pitch(sspbo_ctx* ctx);
int sspbo_encode_f32(sspbo_ctx* ctx, const void* x_dev, int x_dtype, int64_t N, float* phi_out_dev, void* stream);

/* ---- K3: Bayesian linear regression (blr.py:27-97) --------------------------------
 * State on device, float64: S (d x d), S_inv, b = beta * sum(phi_i t_i), m = S b, plus
 * the psi-space scoring factor R (R^T R = B^T S B) and its split-fp16 image. */
int sspbo_blr_init(sspbo_ctx* ctx, double alpha);            /* blr.py:27-33: S = I/alpha, m = 0 */
/* BLR over raw features of any dimension (no SSP space on this ctx): update/predict/get only. */
int sspbo_blr_init_dim(sspbo_ctx* ctx, int d, double alpha);
int sspbo_blr_set_beta(sspbo_ctx* ctx, double beta);         /* blr.py:54-59 computed by the host mirror */
int sspbo_blr_get_beta(const sspbo_ctx* ctx, double* beta, int* is_set);
/* blr.py:60-77: S_inv += beta Phi^T Phi; for each row: Sherman-Morrison on S; m = S b.
 * phis_dev: k x d float64 (device); ts_host: k float64 (host).
 * Over an SSP space, while every observation so far has appended its row to the low-rank form (up to 511 of them), an
 * observation is applied in that form -- psi-space, O(rank d), one launch (k_blr_update_lr) -- and S, S_inv, b, m and the
 * psi-space covariance are brought up to date when something reads them (sspbo_blr_get / _predict / _export, the ascent, the
 * factor-form scorers); the values are those of the reference's update to float64 rounding.  Otherwise (and for raw-feature
 * posteriors) the dense O(d^2) update runs. */
int sspbo_blr_update(sspbo_ctx* ctx, const double* phis_dev, const double* ts_host, int k, void* stream);
/* Posterior update path: fused != 0 (default) = one launch per observation (sspbo_update.cu: the low-rank form, else the
 * cooperative dense kernel) where the device supports it, 0 = the separate GEMV / rank-one kernels (sspbo_core.cu; also the
 * automatic fallback). */
int sspbo_set_update_path(sspbo_ctx* ctx, int fused);

/* SSPAgent.update (agents/ssp_agent.py:187-206) for k <= 64 HOST points: x_host k x (n*L) float64 is staged through pinned
 * memory, encoded on the device (K1) and applied as k sequential observations (K3); asynchronous. */
int sspbo_blr_update_x(sspbo_ctx* ctx, const double* x_host, const double* ts_host, int k, void* stream);
/* One BO observation for one HOST point x_t: (*mu_host, *var_host) = BLR predictive mean / variance at phi(x_t) under the
 * CURRENT posterior (what SSPAgent.eval returns, agents/ssp_agent.py:133-135, float64), then the posterior update with
 * target t (SSPAgent.update, :187-206).  Both come out of the same update launch; synchronises once. */
int sspbo_observe(sspbo_ctx* ctx, const double* x_host, double t, double* mu_host, double* var_host, void* stream);
/* copies of the posterior (attributes BLR.m (d), BLR.S (d x d), BLR.S_inv (d x d)); any pointer may be NULL */
int sspbo_blr_get(sspbo_ctx* ctx, double* m_dev, double* S_dev, double* Sinv_dev, void* stream);
/* device-to-device adoption of a replicated posterior (multi-GPU broadcast target):
 * S, S_inv, b, m, the psi-space covariance, m', and the rows of the low-rank form (so the importing ctx keeps scoring
 * with the low-rank engine), as produced by sspbo_blr_export on another ctx with the same space and alpha. */
int64_t sspbo_blr_export_size(const sspbo_ctx* ctx);          /* number of float64 in the packed state */
int sspbo_blr_export(sspbo_ctx* ctx, double* packed_dev, void* stream);
int sspbo_blr_import(sspbo_ctx* ctx, const double* packed_dev, double beta, void* stream);
/* Checkpoint / rollback of a posterior in low-rank form (fantasy observations of batch BO, a posterior pinned at one rank):
 * sspbo_blr_checkpoint remembers the state after the observations so far; sspbo_blr_rollback undoes every observation made
 * since -- one small kernel: b_psi, m' and the scorer's copies back, the appended operand rows zeroed -- and reports
 * SSPBO_EUNSUPPORTED, leaving everything as it is, when the d x d state has meanwhile absorbed a later observation (S, S_inv
 * or the psi-space covariance was read, or more than 32 observations were made): restore an exported state then. */
int sspbo_blr_checkpoint(sspbo_ctx* ctx, void* stream);
int sspbo_blr_rollback(sspbo_ctx* ctx, void* stream);
/* rank word of the packed state sspbo_blr_export would write now: number of valid low-rank rows, -1 when they are not valid */
int sspbo_blr_rank(const sspbo_ctx* ctx, int* lr_rank);
/* sspbo_blr_import for a caller that knows that word (the exporting context is in the same process: a posterior restored
 * every step, a drift guard): skips the synchronising read-back of it, so the import is purely stream-ordered */
int sspbo_blr_import_ranked(sspbo_ctx* ctx, const double* packed_dev, double beta, int lr_rank, void* stream);
/* blr.py:84-97: var = 1/beta + diag(phi S phi^T), mu = m^T phi^T.  phi_dev N x d f64 -> mu (N), var (N) f64 */
int sspbo_blr_predict(sspbo_ctx* ctx, const double* phi_dev, int64_t N, double* mu_dev, double* var_dev, void* stream);

/* ---- K2: fused encode + acquisition + argmax ---------------------------------------
 * Composition SSPAgent.eval (agents/ssp_agent.py:125-137) -> score = mu + acq -> np.argmax
 * (experiments/demo_blr_gif.py:129-136) without materialising phi.
 *   score_i = m^T phi_i + lam * ( sqrt(1/beta + phi_i^T S phi_i + gamma_t) - sqrt(gamma_t) )
 * x_dev: N x (n*L) candidates (F32|F64); index0: global index of x_dev[0] (for sharded
 * candidate sets); outputs mu/var/acq (float32, N each) are optional (NULL to skip);
 * best_key_dev: one uint64, atomically max-combined with
 *   key = orderable_u32(score) << 32 | (0xFFFFFFFF - global_index)
 * so that max(key) = highest score, lowest index on ties (np.argmax).  The caller zeroes
 * *best_key_dev before the first call of a step; global_index must be < 2^32 (index0 + N above that is refused with
 * SSPBO_EINVAL).  Larger candidate sets are scored shard by shard with indices local to the shard and the shards' keys are
 * combined as (key, shard start) pairs: ssp_bayesopt_b200/dist.py:allgather_best, 16 bytes per rank. */
int sspbo_score(sspbo_ctx* ctx, const void* x_dev, int x_dtype, int64_t N, int64_t index0,
                double lam, double gamma_t, float* mu_dev, float* var_dev, float* acq_dev,
                unsigned long long* best_key_dev, int engine, void* stream);
/* The same scoring over a candidate set GENERATED inside the kernel: rows [index0, index0 + N) of the counter-based
 * uniform stream (kind SSPBO_CAND_UNIFORM, `seed`; axes/counts ignored) or of the n-D grid in the reference's order
 * (kind SSPBO_CAND_GRID; axes_dev / counts_host as for sspbo_grid_points, i.e. the host's np.linspace values, so the
 * winning row is bit-identical to SSPSpace.get_sample_points('grid')[index], sspspace.py:488-495).  No HBM or PCIe byte
 * per candidate is read; keys carry the global row index.  The tensor-memory engines generate the rows in their
 * generator warps; any other engine gets them written to a scratch buffer first.  Not for trajectory contexts. */
int sspbo_score_generated(sspbo_ctx* ctx, int kind, uint64_t seed, const double* axes_dev, const int* counts_host,
                          int64_t N, int64_t index0, double lam, double gamma_t, float* mu_dev, float* var_dev,
                          float* acq_dev, unsigned long long* best_key_dev, int engine, void* stream);
/* Operand format of SSPBO_ENGINE_UMMA_LR: -1 = by rank (default), 0 = split fp16 only (three fp16 products per K step,
 * rank <= 255), 1 = fp16 + e4m3 (one fp16 and one e4m3 product per K step) at every rank, 2 = the same with CTA pairs
 * (tcgen05 cta_group::2: tiles of 256 candidates, each CTA of a cluster of two streams half of the operand rows; results are
 * bit-identical to format 1; n <= 8; also applies to SSPBO_ENGINE_UMMA_F8). */
int sspbo_set_operand_format(sspbo_ctx* ctx, int format);
/* Counters of the low-rank engine since the last reset (synchronises `stream`): candidates scored, re-scored by the
 * exact pass, flagged beyond the list capacity, outside the declared |x| bound.  *rerun != 0 means the keys/outputs
 * accumulated since the last reset must be recomputed (the context has already switched its policy, so simply call
 * sspbo_score again).  Any pointer may be NULL. */
int sspbo_score_stats(sspbo_ctx* ctx, int64_t* scored, int64_t* flagged, int64_t* overflow, int64_t* out_of_range,
                      int* rerun, int reset, void* stream);
/* sspbo_score_stats plus the read-back of the packed key in the same synchronisation (the one host wait of a selection
 * step): *key_host = *key_dev.  key_dev / key_host may both be NULL (then this is sspbo_score_stats). */
int sspbo_score_finish(sspbo_ctx* ctx, const unsigned long long* key_dev, unsigned long long* key_host, int64_t* scored,
                       int64_t* flagged, int64_t* overflow, int64_t* out_of_range, int* rerun, int reset, void* stream);
/* SSPAgent.eval (agents/ssp_agent.py:125-137) for N <= 64 HOST points (x_t of a BO step): x_host N x (n*L) float64;
 * mu/var/acq_host N float64 each, float64 precision like the reference's eval.  Exact float64 engine; one pinned upload, one
 * pinned read-back, one synchronisation. */
int sspbo_eval_host(sspbo_ctx* ctx, const double* x_host, int N, double lam, double gamma_t, double* mu_host,
                    double* var_host, double* acq_host, void* stream);
/* state of the low-rank form: rows of V appended since blr_init, whether it is usable, policy switches */
int sspbo_lowrank_info(const sspbo_ctx* ctx, int* rank, int* valid, int* disabled, int* phase32);
/* policy overrides: lowrank_enabled 1 = enable, 0 = disable, -1 = leave (SSPBO_ENGINE_AUTO); phase32_enabled selects the
 * phase arithmetic of the tensor-core scorers: 0 = Q31.32 only, 1 = fastest path the declared bounds allow (float32 FMA
 * chain, else 32-bit fixed point, else Q31.32), 2 = fixed point only, -1 = leave */
int sspbo_set_policy(sspbo_ctx* ctx, int lowrank_enabled, int phase32_enabled);
/* host helpers for the packed key */
void sspbo_key_unpack(unsigned long long key, float* score, int64_t* index);
unsigned long long sspbo_key_pack(float score, int64_t index);
/* engine the last sspbo_score call used (an SSPBO_ENGINE_* code other than AUTO) and kernel launches so far */
int sspbo_last_engine(const sspbo_ctx* ctx);
int64_t sspbo_launch_count(const sspbo_ctx* ctx);
/* Measurement aid (bench.py's roofline figure; no reference counterpart): with timing enabled every sspbo_score /
 * sspbo_score_generated call records CUDA events on its stream before the scorer kernel, after it and after the float64 pass
 * over flagged candidates.  sspbo_timing_read synchronises the stream and returns the summed durations (ms) and the number of
 * calls since the last reset (at most 4096 calls are kept). */
int sspbo_set_timing(sspbo_ctx* ctx, int enable);
int sspbo_timing_read(sspbo_ctx* ctx, double* scorer_ms, double* flagged_ms, int64_t* calls, int reset, void* stream);

/* ---- K4: from-set decode (sspspace.py:352-356) ------------------------------------
 * unit = q / max(||q||, 1e-6); idx[j] = argmax_i sample_ssps[i] . unit[j], first index on ties.
 * sample_ssps_dev: M x d f64; queries_dev: q x d f64; idx_out_dev: q int64. */
int sspbo_from_set_argmax(sspbo_ctx* ctx, const double* sample_ssps_dev, int64_t M, const double* queries_dev,
                          int q, int64_t* idx_out_dev, void* stream);

/* ---- K4b: direct-optim decode (sspspace.py:358-375) --------------------------------------
 * For every query row: x* = argmax_x phi(x) . (q / max(||q||, 1e-6)) over the box `bounds_host` (n x 2 float64, row =
 * [low, high]; NULL = unbounded), started at x0 (q x n, normally the from-set answer).  The reference minimises the
 * negated objective with scipy's L-BFGS-B and numerical gradients; here the objective, its gradient and its Hessian are
 * evaluated analytically in the Fourier domain and a projected, damped Newton iteration runs on the device (float64,
 * one CTA per query).  queries_dev: q x d; x0_dev, x_out_dev: q x n float64. */
int sspbo_decode_optim(sspbo_ctx* ctx, const double* queries_dev, int q, const double* x0_dev, const double* bounds_host,
                       double* x_out_dev, void* stream);

/* ---- K6: phi-space acquisition ascent (agents/ssp_agent.py:156-185, bayesian_optimization.py:249-287) ----------
 * Maximises g(phi) = m.phi + lam sqrt(gamma_t + 1/beta + phi^T S phi) - sqrt(gamma_t) over the ball |phi| <= margin (the
 * reference clamps at margin = 4) from R starting points at once.  The reference runs scipy's L-BFGS-B per restart; here
 * every restart follows the monotone fixed-point step phi <- margin grad g / |grad g| (g is convex, so no step size is
 * needed and g never decreases) in one cooperative kernel, float64.  A restart stops when its last step moved it by at
 * most tol * margin, or after max_iter steps.  phi0_dev, phi_out_dev: R x d float64 (starting points outside the ball are
 * clamped first, like the reference's objective does); val_out_dev: R acquisition values at phi_out; iters_out_dev
 * (nullable): R step counts.  Works on any context with a posterior (SSP space or sspbo_blr_init_dim). */
int sspbo_acq_ascent(sspbo_ctx* ctx, const double* phi0_dev, int R, double lam, double gamma_t, double margin, int max_iter,
                     double tol, double* phi_out_dev, double* val_out_dev, int* iters_out_dev, void* stream);

/* ---- K5: binding (sspspace.py:551-560) ------------------------------------------------
 * Circular convolution out[r] = a[ra] (*) b[rb] over the last axis, float64, evaluated directly in the time
 * domain (the reference goes through fft/ifft; results agree to ~1e-15).  a_dev: qa x d, b_dev: qb x d with
 * qa == qb, or either equal to 1 (numpy broadcasting); out_dev: max(qa, qb) x d.  d is the context's ssp_dim. */
int sspbo_bind(sspbo_ctx* ctx, const double* a_dev, int qa, const double* b_dev, int qb, double* out_dev, void* stream);

/* ---- length-scale search support (agents/ssp_traj_agent.py:113-143) -----------------------
 * G = Phi Phi^T of n encoded rows (phi_dev: n x d float64, gram_out_dev: n x n float64).  With the Gram matrix of the
 * initial design, every fold of the reference's K-fold search is an n x n problem in the BLR's dual form
 * (mu = k^T (G_tr + alpha/beta I)^-1 t, var = 1/beta + (k** - k^T (G_tr + alpha/beta I)^-1 k) / alpha), which the host
 * mirror solves; encoding at the trial length scales and this product are the only O(d) steps. */
int sspbo_gram(sspbo_ctx* ctx, const double* phi_dev, int n, double* gram_out_dev, void* stream);

/* ---- multi-GPU winner selection over peer memory (one process per GPU on one NVLink / NVSwitch node) ----
 * The all_reduce(MAX) of the packed key without a collective library inside the BO step: ONE small kernel per step writes
 * this rank's key (and whether its scoring pass asks for a rerun) straight into every peer's slot array and waits for the
 * peers' words.  Set-up, once per context: sspbo_dist_init returns a 64-byte handle of this rank's slot array; the caller
 * gathers the handles of all ranks (torch.distributed / MPI / files -- not this library's business) and hands them, in rank
 * order, to sspbo_dist_connect.  Per step: sspbo_dist_exchange (asynchronous, every rank the same number of times) replaces
 * *best_dev by the maximum over the ranks; the following sspbo_score_finish reports rerun != 0 when ANY rank asks for it (then
 * every rank recomputes its shard and exchanges again) and an error when a peer did not arrive within ~10 s.
 * Replaces: dist.allreduce_best (NCCL) of this repo; in the reference the step is np.argmax over one array
 * (bayesian_optimization.py:291). */
int sspbo_dist_init(sspbo_ctx* ctx, int rank, int world, unsigned char* handle_out /* 64 bytes */);
int sspbo_dist_connect(sspbo_ctx* ctx, const unsigned char* handles /* world x 64 bytes, rank order */);
/* ranks inside one process (IPC handles cannot be opened by their exporter): exchange the slot arrays themselves */
int sspbo_dist_slots(sspbo_ctx* ctx, void** slots_out);
int sspbo_dist_connect_ptrs(sspbo_ctx* ctx, void* const* slot_ptrs /* world pointers, rank order */);
int sspbo_dist_ready(const sspbo_ctx* ctx);
int sspbo_dist_exchange(sspbo_ctx* ctx, unsigned long long* best_dev, void* stream);

/* ---- lock-step of many independent runs (BASELINE config C4: 4096 runs, one context each) ---------------------
 * sspbo_score_batch: for r < R: sspbo_score(ctxs[r], x_dev, ..., lam[r], gamma_t[r], key = keys_dev + r, AUTO engine); all
 * runs share the candidate set x_dev (N rows) and must share its row width n*L.  keys_dev (R packed keys) must be zeroed by
 * the caller.  sspbo_update_batch: for r < R: sspbo_blr_update_x(ctxs[r], row r of x_host (R x n*L float64), ts_host[r]).
 * Both only enqueue work.  The caller orders every stream of `streams` after the work that precedes the call and waits for
 * all of them afterwards.  Runs that share a layout (plain SSP spaces of one dimension n <= 4 and table size, the same
 * posterior rank <= 255 as in a lock-step, the float32 phase path) are served by GROUPED kernels on streams[0]: one launch of
 * the fused scorer over (run, tile), one of the float64 pass over (run, flagged group), one for the status words; and one
 * launch for all updates (a thread-block cluster per run encodes the run's point, applies the float64 rank-one update and
 * refreshes the run's scoring operands).  Otherwise the per-run calls go to streams[r % n_streams].  On error the first
 * failing code is returned and *failed (nullable) is the run index. */
int sspbo_score_batch(sspbo_ctx* const* ctxs, int R, const void* x_dev, int x_dtype, int64_t N, const double* lam,
                      const double* gamma_t, unsigned long long* keys_dev, void* const* streams, int n_streams, int* failed);
/* After sspbo_score_batch, on the same streams: status_dev[r] = bit 0 (the flag list of run r's tensor pass overflowed) |
 * bit 1 (candidates outside the declared |x| bound met the fast phase path); non-zero means the key of run r must be
 * recomputed with the safe engines (sspbo_set_policy(ctx, 0, 0), then sspbo_score).  Clears the runs' counters. */
int sspbo_score_batch_status(sspbo_ctx* const* ctxs, int R, unsigned long long* status_dev, void* const* streams,
                             int n_streams, int* failed);
int sspbo_update_batch(sspbo_ctx* const* ctxs, int R, const double* x_host, const double* ts_host, void* const* streams,
                       int n_streams, int* failed);

/* ---- candidate generators ---------------------------------------------------------
 * Counter-based uniform(0,1) float32 stream (same integer hash as
 * oracle/ssp_oracle.py:counter_uniform, bit-exact): rows [start, start+count) of an
 * (infinite x n) stream.  out_dev: count x n float32. */
int sspbo_fill_uniform(sspbo_ctx* ctx, uint64_t seed, int64_t start, int64_t count, int n, float* out_dev, void* stream);
/* Candidate grid in the reference's order (SSPSpace.get_sample_points('grid'),
 * sspspace.py:488-495: np.meshgrid(indexing='xy') flattened C-order).  axes_dev: the
 * per-axis linspace values concatenated (sum(counts) float64, computed by the host with
 * np.linspace so points are bit-identical); rows [start, start+count) -> out_dev (count x n f64). */
int sspbo_grid_points(sspbo_ctx* ctx, const double* axes_dev, const int* counts_host, int n, int64_t start,
                      int64_t count, double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSPBO_H */
