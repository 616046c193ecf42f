// K2, CUDA-core engine: fused encode + BLR acquisition + argmax for any d / trajectory length.
// Float64 phase projection and range reduction, float32 contraction against the float32 image
// of the psi-space factor R (R^T R = B^T S B).  Used for small candidate sets, for shapes the
// tcgen05 engine does not cover, and as the in-library cross-check of that engine.
//
// Replaces the composition SSPAgent.eval (agents/ssp_agent.py:125-137) + np.argmax
// (experiments/demo_blr_gif.py:129-136); never materialises phi.
#include "sspbo_internal.cuh"

constexpr int ST_TC = 16;          // candidates per tile
constexpr int ST_THREADS = 256;

template <typename XT>
__global__ void __launch_bounds__(ST_THREADS)
k_score_simt(const XT* __restrict__ x, int64_t N, int64_t index0, const double* __restrict__ A_turns,
             const double* __restrict__ tau_turns, int n, int K, int L, int dp, const float* __restrict__ Rt,
             const float* __restrict__ mp, float dc, int normalize, float inv_beta, float lam, float gamma_t, float* __restrict__ mu_out,
             float* __restrict__ var_out, float* __restrict__ acq_out, unsigned long long* __restrict__ best, int ncls) {
    extern __shared__ __align__(16) float smem[];
    float* psiT = smem;                                   // dp x ST_TC, k-major
    float* red = smem + (size_t)dp * ST_TC;               // ST_TC x 8 (warps)  then  ST_TC x 16 (mu slices)
    float* red_mu = red + ST_TC * 8;
    float* red_n = red_mu + ST_TC * 16;                   // ST_TC x 16 slices of sum_k psi_k^2 (normalised scoring)
    const int d = 2 * K + 1, nl = n * L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned long long my_best = 0ull;

    const int64_t tiles = (N + ST_TC - 1) / ST_TC;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t base = tile * ST_TC;
        __syncthreads();
        // ---- phase 1: psi tile (float64 phase, exact range reduction) ----
        for (int i = tid; i < ST_TC * K; i += ST_THREADS) {
            int c = i % ST_TC, k = i / ST_TC;
            float cs = 0.f, sn = 0.f;
            if (base + c < N) {
                const XT* xr = x + (base + c) * nl;
                for (int j = 0, cl = 0; j < L; ++j) {
                    const double* Ak = A_turns + ((size_t)cl * K + k) * n;      // waypoint j reads frequency table j % ncls
                    if (++cl == ncls) cl = 0;
                    double t = tau_turns ? tau_turns[(size_t)j * K + k] : 0.0;
                    for (int a = 0; a < n; ++a) t = fma(Ak[a], (double)xr[j * n + a], t);
                    float tr = (float)(t - rint(t));
                    float s, c2;
                    sincospif(2.f * tr, &s, &c2);
                    cs += c2; sn += s;
                }
            }
            psiT[(1 + k) * ST_TC + c] = cs;
            psiT[(1 + K + k) * ST_TC + c] = sn;
        }
        for (int i = tid; i < ST_TC * (dp - d + 1); i += ST_THREADS) {
            int c = i % ST_TC, r = i / ST_TC;
            if (r == 0) psiT[c] = dc;                     // DC term: sum of the waypoints' unit phasors at frequency 0
            else psiT[(d + r - 1) * ST_TC + c] = 0.f;     // padding rows
        }
        __syncthreads();
        // ---- phase 2: y = R psi, q = |y|^2 ; mu = m' . psi ----
        float q[ST_TC];
#pragma unroll
        for (int c = 0; c < ST_TC; ++c) q[c] = 0.f;
        for (int j = tid; j < dp; j += ST_THREADS) {
            float acc[ST_TC];
#pragma unroll
            for (int c = 0; c < ST_TC; ++c) acc[c] = 0.f;
            const float* col = Rt + j;
#pragma unroll 4
            for (int k = 0; k < d; ++k) {
                float r = col[(size_t)k * dp];
                const float4* p = reinterpret_cast<const float4*>(psiT + k * ST_TC);
#pragma unroll
                for (int v = 0; v < ST_TC / 4; ++v) {
                    float4 pv = p[v];
                    acc[4 * v + 0] = fmaf(r, pv.x, acc[4 * v + 0]);
                    acc[4 * v + 1] = fmaf(r, pv.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(r, pv.z, acc[4 * v + 2]);
                    acc[4 * v + 3] = fmaf(r, pv.w, acc[4 * v + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < ST_TC; ++c) q[c] = fmaf(acc[c], acc[c], q[c]);
        }
#pragma unroll
        for (int c = 0; c < ST_TC; ++c) {
            float v = q[c];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[c * 8 + warp] = v;
        }
        {
            int c = tid % ST_TC, slice = tid / ST_TC;     // 16 slices of k
            float u = 0.f, ss = 0.f;
            for (int k = slice; k < d; k += ST_THREADS / ST_TC) {
                const float pv = psiT[k * ST_TC + c];
                u = fmaf(mp[k], pv, u);
                ss = fmaf(pv, pv, ss);
            }
            red_mu[c * 16 + slice] = u;
            red_n[c * 16 + slice] = ss;
        }
        __syncthreads();
        // ---- phase 3: acquisition + running argmax ----
        if (tid < ST_TC && base + tid < N) {
            float qq = 0.f, u = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) qq += red[tid * 8 + w];
#pragma unroll
            for (int s = 0; s < 16; ++s) u += red_mu[tid * 16 + s];
            if (normalize) {                                                  // phi / |phi|, |phi|^2 = psi^T D psi
                float ss = 0.f;
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) ss += red_n[tid * 16 + sl];
                const float nrm2 = (2.f * ss - dc * dc) / (float)d;
                u *= rsqrtf(nrm2); qq /= nrm2;
            }
            float var = inv_beta + qq;                                        // blr.py:96
            // ssp_agent.py:136: lam (sqrt(var + gamma) - sqrt(gamma)), written without the float32 cancellation
            float acq = lam * var / (sqrtf(var + gamma_t) + sqrtf(gamma_t));
            float score = u + acq;
            int64_t gi = base + tid;
            if (mu_out) mu_out[gi] = u;
            if (var_out) var_out[gi] = var;
            if (acq_out) acq_out[gi] = acq;
            unsigned long long key = sspbo_pack(score, (unsigned long long)(index0 + gi));
            if (key > my_best) my_best = key;
        }
    }
    if (best) {
        // lanes 0..15 of warp 0 hold candidates
        if (warp == 0) {
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, my_best, o);
                if (other > my_best) my_best = other;
            }
            if (lane == 0 && my_best) atomicMax(best, my_best);
        }
    }
}

int sspbo_score_simt_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, int64_t index0, float lam, float gamma_t,
                            float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st) {
    size_t smem = ((size_t)ctx->dp * ST_TC + ST_TC * 8 + 2 * ST_TC * 16) * sizeof(float);
    int64_t tiles = (N + ST_TC - 1) / ST_TC;
    int blocks = (int)(tiles < (int64_t)ctx->sm_count * 2 ? tiles : (int64_t)ctx->sm_count * 2);
    float inv_beta = (float)(1.0 / ctx->beta);
    if (x_dtype == SSPBO_F32) {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_score_simt<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_score_simt<float><<<blocks, ST_THREADS, smem, st>>>((const float*)x, N, index0, ctx->A_turns, ctx->tau_turns, ctx->n,
                                                              ctx->K, ctx->L, ctx->dp, ctx->Rt_f32, ctx->mp_f32, (float)ctx->dc, ctx->normalize ? 1 : 0, inv_beta, lam,
                                                              gamma_t, mu, var, acq, best, ctx->n_cls);
    } else {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_score_simt<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_score_simt<double><<<blocks, ST_THREADS, smem, st>>>((const double*)x, N, index0, ctx->A_turns, ctx->tau_turns, ctx->n,
                                                               ctx->K, ctx->L, ctx->dp, ctx->Rt_f32, ctx->mp_f32, (float)ctx->dc, ctx->normalize ? 1 : 0, inv_beta, lam,
                                                               gamma_t, mu, var, acq, best, ctx->n_cls);
    }
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}
