// libsspbo.so K3, fused: one observation of BayesianLinearRegression.update (blr.py:39-82) in ONE cooperative launch.
//
// The unfused path (sspbo_core.cu) issues ~15 small dependent kernels per observation; at d = 151 .. 1023 each of them is
// pure launch latency, and the posterior update is on the critical path of every BO step.  Here every CTA owns a slice of
// the rows of S, S_inv and Sp (one warp per row) and the phases are separated by grid barriers:
//   P0  psi = D^-1 B^T phi (analysis DFT, my frequencies)          v = S phi (my rows)
//   P1  y = Sp psi (my rows)
//   P2  scalars, redundantly per CTA in a fixed order: a = phi.v, a' = psi.y, coef = beta / (1 + beta a), ...
//       S -= coef v v^T,  S_inv += beta phi phi^T,  Sp -= coef' y y^T (my rows);  b += beta t phi;  row of V appended;
//       m = S b for my rows from the updated values in the same pass
//   P3  m' = B^T m (my frequencies)
// Float64 throughout, no atomics: the result is bitwise reproducible, so replicated contexts (one per GPU) stay identical.
// S is updated in the symmetric form v v^T (the reference forms (S phi)(phi^T S), which differs from it by the O(1e-13)
// asymmetry its S accumulates; both are the same rank-one update of the same posterior).
#include "sspbo_internal.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace {

constexpr int UP_THREADS = 256;
constexpr int UP_WARPS = UP_THREADS / 32;
constexpr int LR_CB = 128;             // components per pass of the second sweep over V in k_blr_update_lr

struct UpdateParams {
    int d, K, has_factor, finalize;
    double beta, t, r_scale;
    const double* phi;
    double *S, *Sinv, *b, *m;
    double *Sp, *mp;                       // psi-space covariance (rank order), m' (natural order)
    const double2* twiddle; const int* inv_perm; const int* col_of_rank;
    double *v, *psi, *y;                   // global scratch d-vectors
    __half *vh_row, *vl_row; double* vd_row;   // row of V to append (nullptr: none)
    double* obs_out;                       // nullable: [0] = m . phi, [1] = 1/beta + phi^T S phi under the posterior BEFORE this update
};

__device__ __forceinline__ double up_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sums of NV values per thread, fixed order, totals in tot[0..NV)
template <int NV>
__device__ __forceinline__ void up_block_reduce(double (&a)[NV], double* red, double* tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) a[i] = up_warp_sum(a[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[warp * NV + i] = a[i];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int w = 0; w < UP_WARPS; ++w) t += red[w * NV + threadIdx.x];
        tot[threadIdx.x] = t;
    }
    __syncthreads();
}

// out_k = sum_j in[j] (cos, sin)(2 pi j k / d) for the frequencies k = gw, gw + nw, ... (one warp each); in[] in shared memory
template <typename F>
__device__ __forceinline__ void up_analysis(const double* in, int d, int K, const double2* __restrict__ twiddle, int gw, int nw,
                                            int lane, F&& store) {
    for (int k = gw; k <= K; k += nw) {
        double c = 0.0, s = 0.0;
        int idx = (int)(((long long)lane * k) % d);
        const int step = (int)((32LL * k) % d);
        for (int j = lane; j < d; j += 32) {
            const double2 tw = __ldg(twiddle + idx);
            c = fma(in[j], tw.x, c); s = fma(in[j], tw.y, s);
            idx += step; if (idx >= d) idx -= d;
        }
        c = up_warp_sum(c); s = up_warp_sum(s);
        if (lane == 0) store(k, c, s);
    }
}

// The CTAs that share one update: the whole cooperative grid (one observation per launch), or one thread-block cluster per
// run of a batch of independent runs (k_blr_update_runs: BASELINE config C4)
struct GridTeam {
    cg::grid_group g = cg::this_grid();
    __device__ __forceinline__ int rank() const { return blockIdx.x; }
    __device__ __forceinline__ int size() const { return gridDim.x; }
    __device__ __forceinline__ void sync() { g.sync(); }
};
struct ClusterTeam {
    cg::cluster_group c = cg::this_cluster();
    __device__ __forceinline__ int rank() const { return (int)c.block_rank(); }
    __device__ __forceinline__ int size() const { return (int)c.num_blocks(); }
    __device__ __forceinline__ void sync() { __threadfence(); c.sync(); }     // global-memory scratch (v, psi, y, m) crosses CTAs
};

// psi_from_x: psi (rank order, global) was already written by the caller from the point itself, so P0 skips the analysis DFT
template <typename Team>
__device__ __forceinline__ void blr_update_body(const UpdateParams& p, Team& team, double* sm, bool psi_given) {
    const int d = p.d, K = p.K, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = team.rank(), ncta = team.size();
    const int gw = cta * UP_WARPS + warp, nw = ncta * UP_WARPS;
    double* phi = sm;                 // d
    double* vec = phi + d;            // d : psi, later v
    double* yv = vec + d;             // d : y
    double* bp = yv + d;              // d : b + beta t phi, later m
    double* red = bp + d;             // UP_WARPS x 4
    double* tot = red + UP_WARPS * 4; // 4

    for (int j = tid; j < d; j += UP_THREADS) {
        const double f = __ldcg(p.phi + j);                             // (written by the other CTAs of the cluster in the batched kernel)
        phi[j] = f;
        bp[j] = fma(p.beta * p.t, f, p.b[j]);                           // blr.py:70: b += beta t phi (read before any CTA writes b)
    }
    __syncthreads();
    if (p.obs_out && cta == 0) {                                         // predictive mean at phi under the old posterior
        double mu[1] = {0.0};
        for (int j = tid; j < d; j += UP_THREADS) mu[0] = fma(p.m[j], phi[j], mu[0]);
        up_block_reduce<1>(mu, red, tot);
        if (tid == 0) p.obs_out[0] = tot[0];
        __syncthreads();
    }

    // ---- P0: psi (rank order) and v = S phi ----
    if (p.has_factor && !psi_given)
        up_analysis(phi, d, K, p.twiddle, gw, nw, lane, [&](int k, double c, double s) {
            if (k == 0) __stcg(p.psi + p.inv_perm[0], c);
            else { __stcg(p.psi + p.inv_perm[k], c); __stcg(p.psi + p.inv_perm[K + k], -s); }
        });
    for (int i = gw; i < d; i += nw) {
        const double* row = p.S + (size_t)i * d;
        double acc = 0.0;
#pragma unroll 4
        for (int j = lane; j < d; j += 32) acc = fma(row[j], phi[j], acc);       // (several 256-byte requests per warp in flight)
        acc = up_warp_sum(acc);
        if (lane == 0) __stcg(p.v + i, acc);
    }
    team.sync();

    // ---- P1: y = Sp psi ----
    if (p.has_factor) {
        for (int j = tid; j < d; j += UP_THREADS) vec[j] = __ldcg(p.psi + j);
        __syncthreads();
        for (int i = gw; i < d; i += nw) {
            const double* row = p.Sp + (size_t)i * d;
            double acc = 0.0;
#pragma unroll 4
            for (int j = lane; j < d; j += 32) acc = fma(row[j], vec[j], acc);
            acc = up_warp_sum(acc);
            if (lane == 0) __stcg(p.y + i, acc);
        }
        team.sync();
    }

    // ---- P2: scalars (identical in every CTA), rank-one updates of my rows, b, the row of V, m = S b ----
    double part[4] = {0.0, 0.0, 0.0, 0.0};            // phi.v, psi.y
    for (int j = tid; j < d; j += UP_THREADS) {
        const double vj = __ldcg(p.v + j);
        part[0] = fma(phi[j], vj, part[0]);
        if (p.has_factor) {
            const double yj = __ldcg(p.y + j);
            part[1] = fma(vec[j], yj, part[1]);                          // vec still holds psi
            yv[j] = yj;
        }
    }
    up_block_reduce<4>(part, red, tot);
    const double coef = p.beta / (1.0 + p.beta * tot[0]);               // blr.py:64-67 (Sherman-Morrison)
    const double coef2 = p.beta / (1.0 + p.beta * tot[1]);
    if (p.obs_out && cta == 0 && tid == 0) p.obs_out[1] = 1.0 / p.beta + tot[0];   // blr.py:96 before the update
    __syncthreads();                                                    // everyone has read psi from vec
    for (int j = tid; j < d; j += UP_THREADS) vec[j] = __ldcg(p.v + j);
    __syncthreads();
    for (int i = gw; i < d; i += nw) {
        double* srow = p.S + (size_t)i * d;
        double* irow = p.Sinv + (size_t)i * d;
        const double cv = coef * vec[i], bphi = p.beta * phi[i];
        double acc = 0.0;
#pragma unroll 4
        for (int j = lane; j < d; j += 32) {
            const double s = fma(-cv, vec[j], srow[j]);
            srow[j] = s;
            acc = fma(s, bp[j], acc);
            irow[j] = fma(bphi, phi[j], irow[j]);                       // blr.py:60: S_inv += beta phi phi^T
        }
        if (p.has_factor) {
            double* prow = p.Sp + (size_t)i * d;
            const double cy = coef2 * yv[i];
#pragma unroll 4
            for (int j = lane; j < d; j += 32) prow[j] = fma(-cy, yv[j], prow[j]);
        }
        if (p.finalize) {
            acc = up_warp_sum(acc);
            if (lane == 0) __stcg(p.m + i, acc);                        // blr.py:75: m = S (S_inv0 m0 + beta Phi^T t) = S b
        }
    }
    {   // b and the appended row of V: slices by global thread
        const int gt = cta * UP_THREADS + tid, nt = ncta * UP_THREADS;
        const double sq = sqrt(coef2);
        for (int j = gt; j < d; j += nt) {
            p.b[j] = bp[j];
            if (p.vd_row) {
                const double v0 = yv[j] * sq;
                p.vd_row[j] = v0;
                const double vs = v0 * p.r_scale;
                const __half h = __double2half(vs);
                const int c = p.col_of_rank[j];
                p.vh_row[c] = h;
                p.vl_row[c] = __double2half(vs - (double)__half2float(h));
            }
        }
    }
    if (!p.finalize || !p.has_factor) return;
    team.sync();

    // ---- P3: m' = B^T m (natural psi order) ----
    for (int j = tid; j < d; j += UP_THREADS) bp[j] = __ldcg(p.m + j);
    __syncthreads();
    const double s0 = 1.0 / d, sk = 2.0 / d;
    up_analysis(bp, d, K, p.twiddle, gw, nw, lane, [&](int k, double c, double s) {
        if (k == 0) p.mp[0] = s0 * c;
        else { p.mp[k] = sk * c; p.mp[K + k] = -sk * s; }
    });
}

__global__ void __launch_bounds__(UP_THREADS, 1) k_blr_update_fused(const UpdateParams p) {
    extern __shared__ __align__(16) double sm[];
    GridTeam team;
    blr_update_body(p, team, sm, false);
}

// ---- one observation for each of many independent runs in ONE launch (BASELINE config C4) ----
// One thread-block cluster per run.  The cluster encodes the run's point itself -- psi from the phases (float64 sincospi),
// phi = B psi by the synthesis sum, every CTA a slice of the components -- applies the update above with cluster barriers in
// place of grid barriers, and finishes with the operand-order copy of m' that the scorer reads (k_mp_op of the per-run path).
// The per-run path costs a point upload, k_encode, a cooperative launch (which wants the whole machine, so runs serialise)
// and k_mp_op per run: 48 us of kernel time per run at d = 511 (ncu), 70 % of a lock-step of config C4.
struct RunUpdate {
    UpdateParams up;
    const double* x;                      // this run's point (n float64, device)
    const double* A_turns; int n; double dc;
    double* phi_out;                      // d float64 scratch of the run (= up.phi)
    const int* perm; float* mp_op; __half *vh0, *vl0; float* mscale;     // tail: m' in operand order, row 0 of the operand images
};

__global__ void __launch_bounds__(UP_THREADS, 4) k_blr_update_runs(const RunUpdate* __restrict__ runs) {
    extern __shared__ __align__(16) double sm[];
    ClusterTeam team;
    const int run = blockIdx.x / team.size();
    const RunUpdate& r = runs[run];
    const UpdateParams& p = r.up;
    const int d = p.d, K = p.K, tid = threadIdx.x;
    const int cta = team.rank(), ncta = team.size();
    // psi in natural order [dc, cos_1..K, sin_1..K] (every CTA computes all of it: K sincospi)
    double* psi_n = sm;                   // d doubles; the body reuses this memory afterwards
    for (int k = tid; k < K; k += UP_THREADS) {
        double t = 0.0;
        for (int a = 0; a < r.n; ++a) t = fma(r.A_turns[(size_t)k * r.n + a], r.x[a], t);
        t -= rint(t);
        double sn, cs;
        sincospi(2.0 * t, &sn, &cs);
        psi_n[1 + k] = cs; psi_n[1 + K + k] = sn;
    }
    if (tid == 0) psi_n[0] = r.dc;
    __syncthreads();
    if (p.has_factor && cta == 0)
        for (int k = tid; k < d; k += UP_THREADS) __stcg(p.psi + p.inv_perm[k], psi_n[k]);
    // phi_j = (psi_0 + 2 sum_k (C_k cos(2 pi j k / d) - S_k sin(2 pi j k / d))) / d for my slice of j (sspspace.py:275)
    const int j0 = (int)((long long)d * cta / ncta), j1 = (int)((long long)d * (cta + 1) / ncta);
    for (int j = j0 + tid; j < j1; j += UP_THREADS) {
        double acc = 0.0;
        int idx = 0;
        for (int k = 1; k <= K; ++k) {
            idx += j; if (idx >= d) idx -= d;
            const double2 tw = __ldg(p.twiddle + idx);
            acc = fma(psi_n[k], tw.x, fma(-psi_n[K + k], tw.y, acc));
        }
        __stcg(r.phi_out + j, (psi_n[0] + 2.0 * acc) / (double)d);
    }
    team.sync();
    blr_update_body(p, team, sm, true);
    if (!p.finalize || !p.has_factor) return;
    team.sync();                                                         // m' complete
    if (cta != 0) return;
    // tail (k_mp_op): m' scaled by a power of two into row 0 of the operand images, and in operand order as float32
    double* red = sm;
    double mx = 0.0;
    for (int k = tid; k < d; k += UP_THREADS) mx = fmax(mx, fabs(__ldcg(p.mp + k)));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    double m_all = 0.0;
    for (int w = 0; w < UP_WARPS; ++w) m_all = fmax(m_all, red[w]);
    const double sc = (m_all > 0.0 && isfinite(m_all)) ? exp2(floor(log2(1024.0 / m_all))) : 1.0;
    if (tid == 0) *r.mscale = (float)(1.0 / sc);
    for (int kk = tid; kk < d; kk += UP_THREADS) {
        const double m = __ldcg(p.mp + r.perm[kk]);
        const int c = p.col_of_rank[kk];
        r.mp_op[c] = (float)m;
        const double v = m * sc;
        const __half h = __double2half(v);
        r.vh0[c] = h;
        r.vl0[c] = __double2half(v - (double)__half2float(h));
    }
}

// ---- the same observation in low-rank form: O(rank x d) instead of O(d^2) ---------------------------------------------------
// While every observation has appended its row to V (Sp = Sp0 - V^T V, Sp0 the diagonal prior), the whole update can be done
// in psi-space from the rows of V alone:
//     psi = D^-1 B^T phi (or straight from the point),   g = V psi,   h = V b_psi'        (b_psi' = b_psi + beta t psi)
//     y = Sp psi = Sp0 psi - V^T g,      a = psi . y,      coef = beta / (1 + beta a),      new row of V = sqrt(coef) y
//     m' = Sp_new b_psi' = (Sp0 b_psi' - V^T h) - coef (y . b_psi') y
// -- the identities behind it: phi = B psi, Sp = B^T S B, b = B b_psi, m' = B^T m = B^T S b = Sp b_psi, m . phi = m' . psi.
// Nothing of the d x d state (S, S_inv, Sp) nor the phi-space b and m is read or written; they are brought up to date when a
// consumer asks for them (sspbo_ensure_* in sspbo_core.cu: S -= u u^T with u = B D^-1 row, Sp -= row^T row, S_inv from the ring
// of psi, b = B b_psi, m = B D^-1 m').  One thread-block cluster per run; V is read twice (the second time from L2).
// Float64, fixed summation orders: replicated contexts stay bit-identical.
struct RunLR {
    int d, K, n, rank, KC_pad;
    double beta, t, alpha, dc, r_scale;
    const double* x;                      // the run's point (n float64, device) -- or
    const double* phi_in;                 // its encoding (d float64, device) when x is null
    const double* A_turns; const double2* twiddle;
    const int *perm, *inv_perm, *col_of_rank;
    double* Vd;                           // rows 0 .. rank-1 in rank order; row `rank` is written
    __half *Vh, *Vl;                      // operand images: row 0 = m', row r + 1 = observation r
    double *bpsi, *mp;                    // b_psi (rank order), m' (natural order)
    double* psi_ring_row;                 // psi of this observation in natural order, kept for S_inv
    double* gh;                           // scratch: g[LR_MAX], h[LR_MAX], cluster partial sums [2][16]
    double* psi_x;                        // scratch d: psi in rank order (phi_in path)
    float* mp_op; float* mscale;
    double* obs_out;                      // nullable: (m . phi, 1/beta + phi^T S phi) under the posterior before this update
};

__global__ void __launch_bounds__(UP_THREADS, 4) k_blr_update_lr(const RunLR* __restrict__ runs) {
    extern __shared__ __align__(16) double sm[];
    ClusterTeam team;
    const int run = blockIdx.x / team.size();
    const RunLR& r = runs[run];
    const int d = r.d, K = r.K, rank = r.rank, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = team.rank(), ncta = team.size();
    const int gw = cta * UP_WARPS + warp, nw = ncta * UP_WARPS;
    double* psi = sm;                     // d, rank order
    double* bps = psi + d;                // d, b_psi + beta t psi
    double* ys = bps + d;                 // d: y (my slice), before that phi_in
    double* ms = ys + d;                  // d: Sp0 b_psi' - V^T h (my slice)
    double* gs = ms + d;                  // SSPBO_LR_MAX
    double* hs = gs + SSPBO_LR_MAX;       // SSPBO_LR_MAX
    double* red = hs + SSPBO_LR_MAX;      // UP_WARPS x 4
    double* tot = red + UP_WARPS * 4;     // 4

    // ---- psi (rank order) ----
    if (r.x) {                            // from the point: every CTA evaluates all K phases (sspspace.py:270-275)
        for (int k = tid; k < K; k += UP_THREADS) {
            double t = 0.0;
            for (int a = 0; a < r.n; ++a) t = fma(r.A_turns[(size_t)k * r.n + a], r.x[a], t);
            t -= rint(t);
            double sn, cs;
            sincospi(2.0 * t, &sn, &cs);
            psi[r.inv_perm[1 + k]] = cs; psi[r.inv_perm[1 + K + k]] = sn;
        }
        if (tid == 0) psi[r.inv_perm[0]] = r.dc;
    } else {                              // from phi: psi = D^-1 B^T phi, the frequencies shared out over the cluster
        for (int j = tid; j < d; j += UP_THREADS) ys[j] = r.phi_in[j];
        __syncthreads();
        up_analysis(ys, d, K, r.twiddle, gw, nw, lane, [&](int k, double c, double s) {
            if (k == 0) __stcg(r.psi_x + r.inv_perm[0], c);
            else { __stcg(r.psi_x + r.inv_perm[k], c); __stcg(r.psi_x + r.inv_perm[K + k], -s); }
        });
        team.sync();
        for (int j = tid; j < d; j += UP_THREADS) psi[j] = __ldcg(r.psi_x + j);
    }
    __syncthreads();
    for (int j = tid; j < d; j += UP_THREADS) bps[j] = fma(r.beta * r.t, psi[j], r.bpsi[j]);
    if (cta == 0) {
        for (int j = tid; j < d; j += UP_THREADS) r.psi_ring_row[r.perm[j]] = psi[j];
        if (r.obs_out) {                  // predictive mean under the old posterior: m . phi = m' . psi
            double mu[1] = {0.0};
            for (int j = tid; j < d; j += UP_THREADS) mu[0] = fma(r.mp[r.perm[j]], psi[j], mu[0]);
            up_block_reduce<1>(mu, red, tot);
            if (tid == 0) r.obs_out[0] = tot[0];
        }
    }
    __syncthreads();

    // ---- g = V psi, h = V b_psi' : one warp per row ----
    for (int i = gw; i < rank; i += nw) {
        const double* row = r.Vd + (size_t)i * d;
        double g = 0.0, h = 0.0;
#pragma unroll 4
        for (int j = lane; j < d; j += 32) { const double v = row[j]; g = fma(v, psi[j], g); h = fma(v, bps[j], h); }
        g = up_warp_sum(g); h = up_warp_sum(h);
        if (lane == 0) { __stcg(r.gh + i, g); __stcg(r.gh + SSPBO_LR_MAX + i, h); }
    }
    team.sync();
    for (int i = tid; i < rank; i += UP_THREADS) { gs[i] = __ldcg(r.gh + i); hs[i] = __ldcg(r.gh + SSPBO_LR_MAX + i); }
    __syncthreads();

    // ---- my slice of y = Sp0 psi - V^T g and of Sp0 b_psi' - V^T h ; partial sums of psi . y and y . b_psi' ----
    // The warps share out the rows of V (row i to warp i mod 8, lanes along the components: coalesced), 128 components of the
    // slice per pass; their partial sums meet in shared memory and are added in warp order.
    const int j0 = (int)((long long)d * cta / ncta), j1 = (int)((long long)d * (cta + 1) / ncta);
    const double p0 = 1.0 / ((double)d * r.alpha), pk = 2.0 / ((double)d * r.alpha);      // Sp0 = B^T B / alpha (diagonal)
    double* py = tot + 4;                 // UP_WARPS x LR_CB partial sums of -V^T g
    double* pm = py + UP_WARPS * LR_CB;   // ... and of -V^T h
    double part[2] = {0.0, 0.0};
    for (int q0 = j0; q0 < j1; q0 += LR_CB) {
        double ay[LR_CB / 32], am[LR_CB / 32];
#pragma unroll
        for (int q = 0; q < LR_CB / 32; ++q) { ay[q] = 0.0; am[q] = 0.0; }
        for (int i = warp; i < rank; i += UP_WARPS) {
            const double g = gs[i], h = hs[i];
            const double* row = r.Vd + (size_t)i * d + q0;
#pragma unroll
            for (int q = 0; q < LR_CB / 32; ++q) {
                const int jj = q * 32 + lane;
                if (q0 + jj < j1) { const double v = row[jj]; ay[q] = fma(-g, v, ay[q]); am[q] = fma(-h, v, am[q]); }
            }
        }
        __syncthreads();                                                 // (the previous pass has read py / pm)
#pragma unroll
        for (int q = 0; q < LR_CB / 32; ++q) { py[warp * LR_CB + q * 32 + lane] = ay[q]; pm[warp * LR_CB + q * 32 + lane] = am[q]; }
        __syncthreads();
        const int j = q0 + tid;
        if (tid < LR_CB && j < j1) {
            const double sp0 = (r.perm[j] == 0) ? p0 : pk;
            double y = sp0 * psi[j], mm = sp0 * bps[j];
            for (int w = 0; w < UP_WARPS; ++w) { y += py[w * LR_CB + tid]; mm += pm[w * LR_CB + tid]; }
            ys[j] = y; ms[j] = mm;
            part[0] = fma(psi[j], y, part[0]); part[1] = fma(y, bps[j], part[1]);
        }
    }
    up_block_reduce<2>(part, red, tot);
    if (tid < 2) __stcg(r.gh + 2 * SSPBO_LR_MAX + tid * 16 + cta, tot[tid]);
    team.sync();
    double a = 0.0, yb = 0.0;
    for (int c = 0; c < ncta; ++c) { a += __ldcg(r.gh + 2 * SSPBO_LR_MAX + c); yb += __ldcg(r.gh + 2 * SSPBO_LR_MAX + 16 + c); }
    const double coef = r.beta / (1.0 + r.beta * a);                    // blr.py:64-67 (Sherman-Morrison)
    if (r.obs_out && cta == 0 && tid == 0) r.obs_out[1] = 1.0 / r.beta + a;   // blr.py:96 before the update

    // ---- the new row of V, m', b_psi (my slice) ----
    {
        const double sq = sqrt(coef), cyb = coef * yb;
        double* vd_row = r.Vd + (size_t)rank * d;
        __half* vh_row = r.Vh + (size_t)(rank + 1) * r.KC_pad;
        __half* vl_row = r.Vl + (size_t)(rank + 1) * r.KC_pad;
        for (int j = j0 + tid; j < j1; j += UP_THREADS) {
            const double y = ys[j];
            const double v0 = y * sq;
            vd_row[j] = v0;
            const double vs = v0 * r.r_scale;
            const __half hh = __double2half(vs);
            const int c = r.col_of_rank[j];
            vh_row[c] = hh;
            vl_row[c] = __double2half(vs - (double)__half2float(hh));
            __stcg(r.mp + r.perm[j], fma(-cyb, y, ms[j]));
            r.bpsi[j] = bps[j];
        }
    }
    team.sync();                                                         // m' complete
    if (cta != 0) return;
    // tail (k_mp_op): m' scaled by a power of two into row 0 of the operand images, and in operand order as float32
    double mx = 0.0;
    for (int k = tid; k < d; k += UP_THREADS) mx = fmax(mx, fabs(__ldcg(r.mp + k)));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    double m_all = 0.0;
    for (int w = 0; w < UP_WARPS; ++w) m_all = fmax(m_all, red[w]);
    const double sc = (m_all > 0.0 && isfinite(m_all)) ? exp2(floor(log2(1024.0 / m_all))) : 1.0;
    if (tid == 0) *r.mscale = (float)(1.0 / sc);
    for (int kk = tid; kk < d; kk += UP_THREADS) {
        const double m = __ldcg(r.mp + r.perm[kk]);
        const int c = r.col_of_rank[kk];
        r.mp_op[c] = (float)m;
        const double v = m * sc;
        const __half h = __double2half(v);
        r.Vh[c] = h;
        r.Vl[c] = __double2half(v - (double)__half2float(h));
    }
}

}  // namespace

// whether the single-launch updates (cooperative grid / clusters) can run on this device at this dimension; probed once
bool sspbo_fused_update_available(sspbo_ctx* ctx) {
    const size_t smem = ((size_t)4 * ctx->d + UP_WARPS * 4 + 4) * sizeof(double);
    if (ctx->fused_update_ok < 0) {
        int smem_max = 0, coop = 0;
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        ctx->fused_update_ok = (coop && smem <= (size_t)smem_max && smem <= 200 * 1024) ? 1 : 0;
    }
    return ctx->fused_update_ok > 0;
}

// One observation.  Returns SSPBO_EUNSUPPORTED (without touching the state) when the device cannot run it; the caller then
// falls back to the unfused kernels.
int sspbo_blr_update_fused(sspbo_ctx* ctx, const double* phi, double t, bool finalize, __half* vh_row, __half* vl_row,
                           double* vd_row, double* obs_out, cudaStream_t st) {
    const int d = ctx->d;
    const size_t smem = ((size_t)4 * d + UP_WARPS * 4 + 4) * sizeof(double);
    if (!sspbo_fused_update_available(ctx)) return SSPBO_EUNSUPPORTED;
    SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_blr_update_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UpdateParams p;
    p.d = d; p.K = ctx->K; p.has_factor = ctx->has_factor ? 1 : 0; p.finalize = finalize ? 1 : 0;
    p.beta = ctx->beta; p.t = t; p.r_scale = ctx->r_scale;
    p.phi = phi; p.S = ctx->S; p.Sinv = ctx->Sinv; p.b = ctx->b; p.m = ctx->m; p.Sp = ctx->Sp; p.mp = ctx->mp;
    p.twiddle = ctx->twiddle; p.inv_perm = ctx->inv_perm; p.col_of_rank = ctx->col_of_rank;
    p.v = ctx->v; p.psi = ctx->psi; p.y = ctx->y;
    p.vh_row = vh_row; p.vl_row = vl_row; p.vd_row = vd_row; p.obs_out = obs_out;
    int G = (d + UP_WARPS - 1) / UP_WARPS;
    if (G > ctx->sm_count) G = ctx->sm_count;
    void* args[] = {(void*)&p};
    SSPBO_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_blr_update_fused, dim3(G), dim3(UP_THREADS), args, smem, st));
    ctx->launches++;
    return SSPBO_OK;
}

// One observation for each of R independent runs (one context each, all on one device) in one launch.  Returns
// SSPBO_EUNSUPPORTED -- nothing enqueued, no state touched -- when a run is not a plain SSP-space posterior with the fused
// update available (trajectory / normalised contexts, raw-feature BLRs); the caller then takes the per-run path.
int sspbo_blr_update_runs(sspbo_ctx* const* ctxs, int R, const double* x_host, const double* ts_host, cudaStream_t st, int* failed,
                          double* obs_out) {
    if (R <= 0) return SSPBO_OK;
    {   // every run can append its row of V: the low-rank form, O(rank d) per run
        const int rc = sspbo_blr_update_lr(ctxs, R, x_host, nullptr, ts_host, st, failed, obs_out);
        if (rc != SSPBO_EUNSUPPORTED) return rc;
    }
    sspbo_ctx* c0 = ctxs[0];
    int d_max = 0, n_max = 0;
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        if (!c || c->K == 0 || !c->blr_ready || !c->has_beta || c->L != 1 || c->normalize || !c->has_factor || c->unfused_update ||
            c->device != c0->device || c->n != c0->n)
            return SSPBO_EUNSUPPORTED;
        d_max = c->d > d_max ? c->d : d_max;
        n_max = c->n > n_max ? c->n : n_max;
    }
    int cluster_ok = 0;
    cudaDeviceGetAttribute(&cluster_ok, cudaDevAttrClusterLaunch, c0->device);
    const size_t smem = ((size_t)4 * d_max + UP_WARPS * 4 + 4) * sizeof(double);
    if (!cluster_ok || smem > 200 * 1024) return SSPBO_EUNSUPPORTED;
    SSPBO_CUDA(c0, cudaSetDevice(c0->device));
    // parameter block and points: pinned staging + device copy, owned by the first context
    const size_t bytes = (size_t)R * sizeof(RunUpdate) + (size_t)R * n_max * sizeof(double);
    if (c0->batch_pin) SSPBO_CUDA(c0, cudaEventSynchronize(c0->batch_event));   // the previous upload has left the pinned buffer
    if (bytes > c0->batch_bytes) {
        if (c0->batch_pin) { SSPBO_CUDA(c0, cudaStreamSynchronize(st)); cudaFreeHost(c0->batch_pin); cudaFree(c0->batch_dev); c0->batch_pin = nullptr; c0->batch_dev = nullptr; }
        else SSPBO_CUDA(c0, cudaEventCreateWithFlags(&c0->batch_event, cudaEventDisableTiming));
        c0->batch_bytes = 0;
        SSPBO_CUDA(c0, cudaMallocHost(&c0->batch_pin, bytes));
        SSPBO_CUDA(c0, cudaMalloc(&c0->batch_dev, bytes));
        c0->batch_bytes = bytes;
    }
    RunUpdate* hp = (RunUpdate*)c0->batch_pin;
    double* hx = (double*)((char*)c0->batch_pin + (size_t)R * sizeof(RunUpdate));
    const RunUpdate* dp = (const RunUpdate*)c0->batch_dev;
    const double* dx = (const double*)((const char*)c0->batch_dev + (size_t)R * sizeof(RunUpdate));
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        const int d = c->d;
        if ((size_t)d > c->obs_phi_cap) {
            if (c->obs_phi) { SSPBO_CUDA(c, cudaStreamSynchronize(st)); cudaFree(c->obs_phi); c->obs_phi = nullptr; c->obs_phi_cap = 0; }
            if (cudaMalloc((void**)&c->obs_phi, (size_t)d * sizeof(double)) != cudaSuccess) { if (failed) *failed = r; SSPBO_FAIL(c, SSPBO_ECUDA, "update_batch: out of device memory"); }
            c->obs_phi_cap = (size_t)d;
        }
        {   // this kernel keeps the d x d state current itself: fold in what low-rank updates left pending
            const int rc = sspbo_ensure_all(c, st);
            if (rc) { if (failed) *failed = r; return rc; }
        }
        RunUpdate& u = hp[r];
        UpdateParams& p = u.up;
        p.d = d; p.K = c->K; p.has_factor = 1; p.finalize = 1;
        p.beta = c->beta; p.t = ts_host[r]; p.r_scale = c->r_scale;
        p.phi = c->obs_phi; p.S = c->S; p.Sinv = c->Sinv; p.b = c->b; p.m = c->m; p.Sp = c->Sp; p.mp = c->mp;
        p.twiddle = c->twiddle; p.inv_perm = c->inv_perm; p.col_of_rank = c->col_of_rank;
        p.v = c->v; p.psi = c->psi; p.y = c->y;
        p.vh_row = nullptr; p.vl_row = nullptr; p.vd_row = nullptr; p.obs_out = (R == 1) ? obs_out : nullptr;
        if (c->lr_valid) {                                                 // the row of V this observation appends
            if (c->lr_rank >= SSPBO_LR_MAX) c->lr_valid = false;
            else {
                const size_t off = (size_t)(c->lr_rank + 1) * c->KC_pad;
                p.vh_row = c->Vh + off; p.vl_row = c->Vl + off; p.vd_row = c->Vd + (size_t)c->lr_rank * d;
            }
        }
        u.x = dx + (size_t)r * n_max; u.A_turns = c->A_turns; u.n = c->n; u.dc = c->dc; u.phi_out = c->obs_phi;
        u.perm = c->perm; u.mp_op = c->mp_op; u.vh0 = c->Vh; u.vl0 = c->Vl; u.mscale = c->mscale;
        for (int a = 0; a < c->n; ++a) hx[(size_t)r * n_max + a] = x_host[(size_t)r * c->n + a];
    }
    SSPBO_CUDA(c0, cudaMemcpyAsync(c0->batch_dev, c0->batch_pin, bytes, cudaMemcpyHostToDevice, st));
    SSPBO_CUDA(c0, cudaEventRecord(c0->batch_event, st));
    // clusters of 4 CTAs (8 from d = 768 on): one run per cluster, as many clusters in flight as the machine holds
    const int G = d_max >= 768 ? 8 : 4;
    SSPBO_CUDA(c0, cudaFuncSetAttribute(k_blr_update_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(R * G)); cfg.blockDim = dim3(UP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    SSPBO_CUDA(c0, cudaLaunchKernelEx(&cfg, k_blr_update_runs, dp));
    c0->launches++;
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        if (hp[r].up.vd_row) c->lr_rank++;
        sspbo_eager_update_done(c);
        c->v8_row0_dirty = true;
        c->operands_dirty = true; c->factor_dirty = true; c->f8_factor_dirty = true;
    }
    return SSPBO_OK;
}

// One observation for each of R runs in low-rank form (k_blr_update_lr), one launch.  x_host (R rows of n, nullable) or
// phi_dev (R device pointers to d-vectors).  SSPBO_EUNSUPPORTED -- nothing enqueued, no state touched -- unless every run is a
// posterior over an SSP space whose observations have all appended their row to V and that has room for one more.
int sspbo_blr_update_lr(sspbo_ctx* const* ctxs, int R, const double* x_host, const double* const* phi_dev, const double* ts_host,
                        cudaStream_t st, int* failed, double* obs_out) {
    if (R <= 0) return SSPBO_OK;
    sspbo_ctx* c0 = ctxs[0];
    int d_max = 0, n_max = 0;
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        if (!c || c->K == 0 || !c->blr_ready || !c->has_beta || !c->has_factor || c->unfused_update || !c->lr_valid ||
            c->lr_rank >= SSPBO_LR_MAX || c->device != c0->device)
            return SSPBO_EUNSUPPORTED;
        if (x_host && (c->L != 1 || c->normalize || c->n != c0->n)) return SSPBO_EUNSUPPORTED;
        d_max = c->d > d_max ? c->d : d_max;
        n_max = c->n > n_max ? c->n : n_max;
    }
    int cluster_ok = 0;
    cudaDeviceGetAttribute(&cluster_ok, cudaDevAttrClusterLaunch, c0->device);
    const size_t smem = ((size_t)4 * d_max + 2 * SSPBO_LR_MAX + UP_WARPS * 4 + 4 + 2 * UP_WARPS * LR_CB) * sizeof(double);
    if (!cluster_ok || smem > 100 * 1024) return SSPBO_EUNSUPPORTED;
    SSPBO_CUDA(c0, cudaSetDevice(c0->device));
    const size_t bytes = (size_t)R * sizeof(RunLR) + (size_t)R * n_max * sizeof(double);
    if (c0->batch_pin) SSPBO_CUDA(c0, cudaEventSynchronize(c0->batch_event));   // the previous upload has left the pinned buffer
    if (bytes > c0->batch_bytes) {
        if (c0->batch_pin) { SSPBO_CUDA(c0, cudaStreamSynchronize(st)); cudaFreeHost(c0->batch_pin); cudaFree(c0->batch_dev); c0->batch_pin = nullptr; c0->batch_dev = nullptr; }
        else SSPBO_CUDA(c0, cudaEventCreateWithFlags(&c0->batch_event, cudaEventDisableTiming));
        c0->batch_bytes = 0;
        SSPBO_CUDA(c0, cudaMallocHost(&c0->batch_pin, bytes));
        SSPBO_CUDA(c0, cudaMalloc(&c0->batch_dev, bytes));
        c0->batch_bytes = bytes;
    }
    RunLR* hp = (RunLR*)c0->batch_pin;
    double* hx = (double*)((char*)c0->batch_pin + (size_t)R * sizeof(RunLR));
    const RunLR* dp = (const RunLR*)c0->batch_dev;
    const double* dx = (const double*)((const char*)c0->batch_dev + (size_t)R * sizeof(RunLR));
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        double* ring_row = nullptr;
        const int rc = sspbo_lr_update_prepare(c, st, &ring_row);           // scratch, b_psi current, room in the ring of psi
        if (rc) { if (failed) *failed = r; return rc; }
        RunLR& u = hp[r];
        u.d = c->d; u.K = c->K; u.n = c->n; u.rank = c->lr_rank; u.KC_pad = c->KC_pad;
        u.beta = c->beta; u.t = ts_host[r]; u.alpha = c->alpha; u.dc = c->dc; u.r_scale = c->r_scale;
        u.x = x_host ? dx + (size_t)r * n_max : nullptr; u.phi_in = x_host ? nullptr : phi_dev[r];
        u.A_turns = c->A_turns; u.twiddle = c->twiddle; u.perm = c->perm; u.inv_perm = c->inv_perm; u.col_of_rank = c->col_of_rank;
        u.Vd = c->Vd; u.Vh = c->Vh; u.Vl = c->Vl; u.bpsi = c->bpsi; u.mp = c->mp; u.psi_ring_row = ring_row;
        u.gh = c->lr_gh; u.psi_x = c->psi; u.mp_op = c->mp_op; u.mscale = c->mscale;
        u.obs_out = (R == 1) ? obs_out : nullptr;
        if (x_host) for (int a = 0; a < c->n; ++a) hx[(size_t)r * n_max + a] = x_host[(size_t)r * c->n + a];
    }
    SSPBO_CUDA(c0, cudaMemcpyAsync(c0->batch_dev, c0->batch_pin, bytes, cudaMemcpyHostToDevice, st));
    SSPBO_CUDA(c0, cudaEventRecord(c0->batch_event, st));
    // one cluster per run: 4 CTAs (8 from d = 768 on) share the rows of V and the components of y
    const int G = d_max >= 768 ? 8 : 4;
    SSPBO_CUDA(c0, cudaFuncSetAttribute(k_blr_update_lr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(R * G)); cfg.blockDim = dim3(UP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    SSPBO_CUDA(c0, cudaLaunchKernelEx(&cfg, k_blr_update_lr, dp));
    c0->launches++;
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* c = ctxs[r];
        c->lr_rank++; c->sinv_pending++; c->bm_stale = true;
        c->v8_row0_dirty = true;
        c->operands_dirty = true; c->factor_dirty = true; c->f8_factor_dirty = true;
    }
    return SSPBO_OK;
}
