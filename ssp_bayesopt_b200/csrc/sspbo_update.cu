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

__global__ void __launch_bounds__(UP_THREADS, 1) k_blr_update_fused(const UpdateParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double sm[];
    const int d = p.d, K = p.K, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gw = blockIdx.x * UP_WARPS + warp, nw = gridDim.x * UP_WARPS;
    double* phi = sm;                 // d
    double* vec = phi + d;            // d : psi, later v
    double* yv = vec + d;             // d : y
    double* bp = yv + d;              // d : b + beta t phi, later m
    double* red = bp + d;             // UP_WARPS x 4
    double* tot = red + UP_WARPS * 4; // 4

    for (int j = tid; j < d; j += UP_THREADS) {
        const double f = p.phi[j];
        phi[j] = f;
        bp[j] = fma(p.beta * p.t, f, p.b[j]);                           // blr.py:70: b += beta t phi (read before any CTA writes b)
    }
    __syncthreads();
    if (p.obs_out && blockIdx.x == 0) {                                  // predictive mean at phi under the old posterior
        double mu[1] = {0.0};
        for (int j = tid; j < d; j += UP_THREADS) mu[0] = fma(p.m[j], phi[j], mu[0]);
        up_block_reduce<1>(mu, red, tot);
        if (tid == 0) p.obs_out[0] = tot[0];
        __syncthreads();
    }

    // ---- P0: psi (rank order) and v = S phi ----
    if (p.has_factor)
        up_analysis(phi, d, K, p.twiddle, gw, nw, lane, [&](int k, double c, double s) {
            if (k == 0) __stcg(p.psi + p.inv_perm[0], c);
            else { __stcg(p.psi + p.inv_perm[k], c); __stcg(p.psi + p.inv_perm[K + k], -s); }
        });
    for (int i = gw; i < d; i += nw) {
        const double* row = p.S + (size_t)i * d;
        double acc = 0.0;
        for (int j = lane; j < d; j += 32) acc = fma(row[j], phi[j], acc);
        acc = up_warp_sum(acc);
        if (lane == 0) __stcg(p.v + i, acc);
    }
    grid.sync();

    // ---- P1: y = Sp psi ----
    if (p.has_factor) {
        for (int j = tid; j < d; j += UP_THREADS) vec[j] = __ldcg(p.psi + j);
        __syncthreads();
        for (int i = gw; i < d; i += nw) {
            const double* row = p.Sp + (size_t)i * d;
            double acc = 0.0;
            for (int j = lane; j < d; j += 32) acc = fma(row[j], vec[j], acc);
            acc = up_warp_sum(acc);
            if (lane == 0) __stcg(p.y + i, acc);
        }
        grid.sync();
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
    if (p.obs_out && blockIdx.x == 0 && tid == 0) p.obs_out[1] = 1.0 / p.beta + tot[0];   // blr.py:96 before the update
    __syncthreads();                                                    // everyone has read psi from vec
    for (int j = tid; j < d; j += UP_THREADS) vec[j] = __ldcg(p.v + j);
    __syncthreads();
    for (int i = gw; i < d; i += nw) {
        double* srow = p.S + (size_t)i * d;
        double* irow = p.Sinv + (size_t)i * d;
        const double cv = coef * vec[i], bphi = p.beta * phi[i];
        double acc = 0.0;
        for (int j = lane; j < d; j += 32) {
            const double s = fma(-cv, vec[j], srow[j]);
            srow[j] = s;
            acc = fma(s, bp[j], acc);
            irow[j] = fma(bphi, phi[j], irow[j]);                       // blr.py:60: S_inv += beta phi phi^T
        }
        if (p.has_factor) {
            double* prow = p.Sp + (size_t)i * d;
            const double cy = coef2 * yv[i];
            for (int j = lane; j < d; j += 32) prow[j] = fma(-cy, yv[j], prow[j]);
        }
        if (p.finalize) {
            acc = up_warp_sum(acc);
            if (lane == 0) __stcg(p.m + i, acc);                        // blr.py:75: m = S (S_inv0 m0 + beta Phi^T t) = S b
        }
    }
    {   // b and the appended row of V: slices by global thread
        const int gt = blockIdx.x * UP_THREADS + tid, nt = gridDim.x * UP_THREADS;
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
    grid.sync();

    // ---- P3: m' = B^T m (natural psi order) ----
    for (int j = tid; j < d; j += UP_THREADS) bp[j] = __ldcg(p.m + j);
    __syncthreads();
    const double s0 = 1.0 / d, sk = 2.0 / d;
    up_analysis(bp, d, K, p.twiddle, gw, nw, lane, [&](int k, double c, double s) {
        if (k == 0) p.mp[0] = s0 * c;
        else { p.mp[k] = sk * c; p.mp[K + k] = -sk * s; }
    });
}

}  // namespace

// One observation.  Returns SSPBO_EUNSUPPORTED (without touching the state) when the device cannot run it; the caller then
// falls back to the unfused kernels.
int sspbo_blr_update_fused(sspbo_ctx* ctx, const double* phi, double t, bool finalize, __half* vh_row, __half* vl_row,
                           double* vd_row, double* obs_out, cudaStream_t st) {
    const int d = ctx->d;
    const size_t smem = ((size_t)4 * d + UP_WARPS * 4 + 4) * sizeof(double);
    if (ctx->fused_update_ok < 0) {                       // probe once
        int smem_max = 0, coop = 0;
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        ctx->fused_update_ok = (coop && smem <= (size_t)smem_max) ? 1 : 0;
    }
    if (!ctx->fused_update_ok || smem > 200 * 1024) return SSPBO_EUNSUPPORTED;
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
