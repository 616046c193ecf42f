// libsspbo.so K6: phi-space ascent of the acquisition function, all restarts advanced together on the device.
//
// Reference route being replaced (ctn-waterloo/ssp-bayesopt, ssp_bayes_opt/):
//   objective / gradient     agents/ssp_agent.py:156-185   g(phi) = m.phi + lam sqrt(gamma + 1/beta + phi^T S phi) - sqrt(gamma),
//                                                          phi clamped to the ball |phi| <= margin (= 4)
//   multi-restart loop       bayesian_optimization.py:249-287 (scipy L-BFGS-B per restart from blr.sample())
//
// g is convex, so its maximum over the ball sits on the sphere and every step
//     phi <- margin * grad g(phi) / |grad g(phi)|,    grad g = m + lam S phi / sqrt(gamma + 1/beta + phi^T S phi)
// maximises the tangent plane over the ball and therefore never decreases g (the generalised power method); the fixed
// points are exactly the stationary points L-BFGS-B looks for.  No step size, no line search, float64 throughout.
//
// One cooperative persistent kernel per tile of AA_RT restarts: every CTA keeps all iterates and m in shared memory, owns
// a slice of the rows of S (S stays L2-resident: 8.4 MB at d = 1023), publishes its rows of Y = S Phi and its partial
// phi.y, m.y and y.y sums (they give phi^T S phi and |grad g|^2 = |m|^2 + 2 c m.y + c^2 y.y without another pass over Y), and after ONE grid barrier per iteration every CTA rebuilds the full update redundantly and bit-identically
// (fixed summation orders), so the stopping decision is uniform without a second barrier.  Y and the partials are
// double-buffered across iterations.
#include "sspbo_internal.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace {

constexpr int AA_THREADS = 256;
constexpr int AA_WARPS = AA_THREADS / 32;
constexpr int AA_RT = 8;                    // restarts advanced per launch

struct AscentParams {
    const double* S; const double* m; const double* phi0;
    double* phi_out; double* val_out; int* it_out;
    double* ybuf;                           // 2 x d x AA_RT
    double* part;                           // 2 x gridDim.x x AA_RT
    int d, R, max_iter;
    double lam, c, sqrt_gamma, margin, tol;
};

// Iterates in shared memory: AA_RT / 2 planes of [d] double2 (restart pairs), so a warp reading 32 consecutive columns of
// one plane touches 512 contiguous bytes (no bank conflicts); element (column i, restart r):
__device__ __forceinline__ int aa_at(int d, int i, int r) { return (((r >> 1) * d + i) << 1) | (r & 1); }

__device__ __forceinline__ double aa_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums of NV per-thread values, fixed order; totals land in tot[0..NV) (visible to every thread on return).
template <int NV>
__device__ __forceinline__ void aa_block_reduce(double (&a)[NV], double* red, double* tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) a[v] = aa_warp_sum(a[v]);
    __syncthreads();                                        // previous readers of red / tot are done
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) red[warp * NV + v] = a[v];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int w = 0; w < AA_WARPS; ++w) t += red[w * NV + threadIdx.x];
        tot[threadIdx.x] = t;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(AA_THREADS, 1) k_acq_ascent(const AscentParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double sm[];
    const int d = p.d, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, G = gridDim.x;
    double* phi = sm;                                       // iterates, see aa_at()
    double* mv = phi + (size_t)d * AA_RT;                   // [d]
    double* red = mv + d;                                   // [AA_WARPS][3 AA_RT]
    double* tot = red + AA_WARPS * 3 * AA_RT;               // [3 AA_RT]
    double* mphi = tot + 3 * AA_RT;                         // [AA_RT]  m . phi_r of the current iterates
    double* moved = mphi + AA_RT;                           // [AA_RT]  |phi_r - previous phi_r|
    double* coef = moved + AA_RT;                           // [AA_RT]  lam / sqrt(c + q_r)
    double* gscale = coef + AA_RT;                          // [AA_RT]  margin / |grad g(phi_r)|
    double* mm_s = gscale + AA_RT;                          // [1]      m . m
    __shared__ int first_it[AA_RT];

    // ---- load m and the starting points; `_clamp` of ssp_agent.py:160-163 pulls them into the ball ----
    {
        double acc[2 * AA_RT + 1];
#pragma unroll
        for (int v = 0; v < 2 * AA_RT + 1; ++v) acc[v] = 0.0;
        for (int j = tid; j < d; j += AA_THREADS) {
            const double mj = p.m[j];
            mv[j] = mj;
            acc[2 * AA_RT] = fma(mj, mj, acc[2 * AA_RT]);
#pragma unroll
            for (int r = 0; r < AA_RT; ++r) {
                const int src = r < p.R ? r : p.R - 1;      // unused lanes of the tile shadow the last restart
                const double v = p.phi0[(size_t)src * d + j];
                phi[aa_at(d, j, r)] = v;
                acc[r] = fma(v, v, acc[r]);
                acc[AA_RT + r] = fma(mj, v, acc[AA_RT + r]);
            }
        }
        aa_block_reduce<2 * AA_RT + 1>(acc, red, tot);
        double scale[AA_RT];
#pragma unroll
        for (int r = 0; r < AA_RT; ++r) {
            const double nrm = sqrt(tot[r]);
            scale[r] = nrm > p.margin ? p.margin / nrm : 1.0;
        }
        for (int j = tid; j < d; j += AA_THREADS) {
#pragma unroll
            for (int r = 0; r < AA_RT; ++r) phi[aa_at(d, j, r)] *= scale[r];
        }
        if (tid < AA_RT) {
            mphi[tid] = tot[AA_RT + tid] * scale[tid];
            moved[tid] = INFINITY;
            first_it[tid] = -1;
        }
        if (tid == 0) mm_s[0] = tot[2 * AA_RT];
        __syncthreads();
    }

    for (int k = 0;; ++k) {
        double* yb = p.ybuf + (size_t)(k & 1) * d * AA_RT;
        double* pb = p.part + (size_t)(k & 1) * G * 3 * AA_RT;   // per CTA: phi.y, m.y, y.y of its rows, per restart

        // ---- my rows of Y = S Phi, one warp per row, and the partial phi_r . y_r over those rows ----
        double pq[AA_RT], pmy[AA_RT], pyy[AA_RT];
#pragma unroll
        for (int r = 0; r < AA_RT; ++r) { pq[r] = 0.0; pmy[r] = 0.0; pyy[r] = 0.0; }
        for (int j = blockIdx.x * AA_WARPS + warp; j < d; j += G * AA_WARPS) {
            const double* __restrict__ row = p.S + (size_t)j * d;
            double acc[AA_RT];
#pragma unroll
            for (int r = 0; r < AA_RT; ++r) acc[r] = 0.0;
#pragma unroll 8
            for (int i = lane; i < d; i += 32) {
                const double s = __ldg(row + i);
#pragma unroll
                for (int pr = 0; pr < AA_RT / 2; ++pr) {
                    const double2 v = reinterpret_cast<const double2*>(phi)[pr * d + i];
                    acc[2 * pr] = fma(s, v.x, acc[2 * pr]);
                    acc[2 * pr + 1] = fma(s, v.y, acc[2 * pr + 1]);
                }
            }
            double mine = 0.0;
#pragma unroll
            for (int r = 0; r < AA_RT; ++r) {
                acc[r] = aa_warp_sum(acc[r]);
                pq[r] = fma(phi[aa_at(d, j, r)], acc[r], pq[r]);
                pmy[r] = fma(mv[j], acc[r], pmy[r]);
                pyy[r] = fma(acc[r], acc[r], pyy[r]);
                if (lane == r) mine = acc[r];
            }
            if (lane < AA_RT) __stcg(yb + (size_t)j * AA_RT + lane, mine);
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < AA_RT; ++r) {
                red[warp * 3 * AA_RT + r] = pq[r];
                red[warp * 3 * AA_RT + AA_RT + r] = pmy[r];
                red[warp * 3 * AA_RT + 2 * AA_RT + r] = pyy[r];
            }
        }
        __syncthreads();
        if (tid < 3 * AA_RT) {
            double t = 0.0;
            for (int w = 0; w < AA_WARPS; ++w) t += red[w * 3 * AA_RT + tid];
            __stcg(pb + (size_t)blockIdx.x * 3 * AA_RT + tid, t);
        }

        grid.sync();

        // ---- every CTA: q_r = phi_r^T S phi_r in CTA order, the value at the current iterate, the stopping rule ----
        {   // 32 interleaved sub-sums per restart (independent loads), folded in a fixed order
            const int r = lane & (AA_RT - 1), sub = warp * (32 / AA_RT) + lane / AA_RT;
            double t[3] = {0.0, 0.0, 0.0};
#pragma unroll 2
            for (int b = sub; b < G; b += AA_WARPS * (32 / AA_RT)) {
#pragma unroll
                for (int v = 0; v < 3; ++v) t[v] += __ldcg(pb + (size_t)b * 3 * AA_RT + v * AA_RT + r);
            }
#pragma unroll
            for (int v = 0; v < 3; ++v) {
#pragma unroll
                for (int o = AA_RT; o < 32; o <<= 1) t[v] += __shfl_xor_sync(0xffffffffu, t[v], o);
                if (lane < AA_RT) red[warp * 3 * AA_RT + v * AA_RT + lane] = t[v];
            }
        }
        __syncthreads();
        if (tid < AA_RT) {
            double q = 0.0, my = 0.0, yy = 0.0;
            for (int w = 0; w < AA_WARPS; ++w) {
                q += red[w * 3 * AA_RT + tid];
                my += red[w * 3 * AA_RT + AA_RT + tid];
                yy += red[w * 3 * AA_RT + 2 * AA_RT + tid];
            }
            const double s = sqrt(p.c + q);
            const double cfv = p.lam / s;
            tot[tid] = mphi[tid] + p.lam * s - p.sqrt_gamma;            // acquisition value (ssp_agent.py:167-172)
            coef[tid] = cfv;
            // |grad g|^2 = |m + cf y|^2 from the same partial sums: no separate pass over Y for the norm
            gscale[tid] = p.margin / sqrt(fma(cfv, fma(cfv, yy, 2.0 * my), mm_s[0]));
            if (first_it[tid] < 0 && (moved[tid] <= p.tol * p.margin || k == p.max_iter)) first_it[tid] = k;
        }
        __syncthreads();
        bool all_done = true;
        for (int r = 0; r < p.R; ++r) all_done = all_done && (first_it[r] >= 0);
        if (all_done) {
            if (blockIdx.x == 0) {
                for (int j = tid; j < d; j += AA_THREADS)
                    for (int r = 0; r < p.R; ++r) p.phi_out[(size_t)r * d + j] = phi[aa_at(d, j, r)];
                if (tid < p.R) { p.val_out[tid] = tot[tid]; if (p.it_out) p.it_out[tid] = first_it[tid]; }
            }
            break;
        }

        // ---- phi+ = margin grad / |grad|, grad_j = m_j + coef_r y_jr (one pass over Y) ----
        double cf[AA_RT], sc[AA_RT];
#pragma unroll
        for (int r = 0; r < AA_RT; ++r) { cf[r] = coef[r]; sc[r] = gscale[r]; }
        {
            double acc[2 * AA_RT];
#pragma unroll
            for (int v = 0; v < 2 * AA_RT; ++v) acc[v] = 0.0;
            for (int j = tid; j < d; j += AA_THREADS) {
                const double mj = mv[j];
#pragma unroll
                for (int r = 0; r < AA_RT; ++r) {
                    const double nw = sc[r] * fma(cf[r], __ldcg(yb + (size_t)j * AA_RT + r), mj);
                    const double df = nw - phi[aa_at(d, j, r)];
                    phi[aa_at(d, j, r)] = nw;
                    acc[r] = fma(df, df, acc[r]);
                    acc[AA_RT + r] = fma(mj, nw, acc[AA_RT + r]);
                }
            }
            aa_block_reduce<2 * AA_RT>(acc, red, tot);
        }
        if (tid < AA_RT) { moved[tid] = sqrt(tot[tid]); mphi[tid] = tot[AA_RT + tid]; }
        __syncthreads();
    }
}

static size_t ascent_smem(int d) {
    return ((size_t)d * AA_RT + d + AA_WARPS * 3 * AA_RT + 3 * AA_RT + 4 * AA_RT + 1) * sizeof(double);
}

}  // namespace

extern "C" int sspbo_acq_ascent(sspbo_ctx* ctx, const double* phi0, int R, double lam, double gamma_t, double margin,
                                int max_iter, double tol, double* phi_out, double* val_out, int* iters_out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready || !ctx->has_beta) SSPBO_FAIL(ctx, SSPBO_ESTATE, "acq_ascent before the first blr_update");
    if (R < 0 || (R > 0 && (!phi0 || !phi_out || !val_out))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "acq_ascent: null pointer");
    if (!(margin > 0) || !(gamma_t >= 0) || max_iter < 0 || !(tol >= 0))
        SSPBO_FAIL(ctx, SSPBO_EINVAL, "acq_ascent: need margin > 0, gamma >= 0, max_iter >= 0, tol >= 0");
    if (R == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int d = ctx->d;
    const size_t smem = ascent_smem(d);
    int smem_max = 0, coop = 0;
    SSPBO_CUDA(ctx, cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    SSPBO_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "acq_ascent: device lacks cooperative launch");
    if (smem > (size_t)smem_max) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "acq_ascent: ssp_dim %d needs %zu B of shared memory", d, smem);
    SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_acq_ascent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int G = (d + AA_WARPS - 1) / AA_WARPS;
    if (G > ctx->sm_count) G = ctx->sm_count;
    const size_t need = 2 * ((size_t)d * AA_RT + (size_t)G * 3 * AA_RT);
    if (need > ctx->aa_cap) {
        if (ctx->aa_work) { SSPBO_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(ctx->aa_work); ctx->aa_work = nullptr; ctx->aa_cap = 0; }
        SSPBO_CUDA(ctx, cudaMalloc((void**)&ctx->aa_work, need * sizeof(double)));
        ctx->aa_cap = need;
    }
    { const int rc = sspbo_ensure_s(ctx, st); if (rc) return rc; }       // S and m as of the last observation
    for (int r0 = 0; r0 < R; r0 += AA_RT) {
        AscentParams p;
        p.S = ctx->S; p.m = ctx->m; p.phi0 = phi0 + (size_t)r0 * d;
        p.phi_out = phi_out + (size_t)r0 * d; p.val_out = val_out + r0; p.it_out = iters_out ? iters_out + r0 : nullptr;
        p.ybuf = ctx->aa_work; p.part = ctx->aa_work + 2 * (size_t)d * AA_RT;
        p.d = d; p.R = (R - r0 < AA_RT) ? R - r0 : AA_RT; p.max_iter = max_iter;
        p.lam = lam; p.c = gamma_t + 1.0 / ctx->beta; p.sqrt_gamma = sqrt(gamma_t); p.margin = margin; p.tol = tol;
        void* args[] = {(void*)&p};
        SSPBO_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_acq_ascent, dim3(G), dim3(AA_THREADS), args, smem, st));
        ctx->launches++;
    }
    return SSPBO_OK;
}
