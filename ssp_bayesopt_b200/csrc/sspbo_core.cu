// libsspbo.so core: context, SSP space, K1 encode, K3 BLR posterior (float64), K4 from-set
// decode, candidate generators.  Hand-written CUDA for sm_100a; no CPU fallback anywhere.
//
// Reference arithmetic being replaced (ctn-waterloo/ssp-bayesopt, ssp_bayes_opt/):
//   SSPSpace.encode                sspspace.py:254-276
//   BayesianLinearRegression.*     blr.py:27-97
//   SSPSpace.decode('from-set')    sspspace.py:352-356
//   SSPSpace.get_sample_points     sspspace.py:488-495
#include "sspbo_lr_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>

static void timing_release(sspbo_ctx* ctx);

static const double kTwoPi = 6.283185307179586476925286766559;

// =====================================================================================
// small device utilities
// =====================================================================================
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum, result valid in every thread; red must hold >= 32 doubles
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = (l < nw) ? red[l] : 0.0;
    t = warp_sum(t);
    return t;
}

// =====================================================================================
// K1  encode: x -> psi (shared memory) -> phi = B psi   (float64 throughout)
// =====================================================================================
// psi of one candidate row (L waypoints of n coordinates) at frequency k: sum of L unit phasors.
template <typename XT>
__device__ __forceinline__ void psi_entry(const XT* __restrict__ xrow, const double* __restrict__ A_turns,
                                          const double* __restrict__ tau_turns, int n, int K, int L, int k,
                                          double& cs, double& sn, int ncls = 1) {
    // ncls > 1: waypoint j reads the frequency table of class j % ncls (one x space per agent, A_turns = ncls x K x n)
    cs = 0.0; sn = 0.0;
    for (int j = 0, cl = 0; j < L; ++j) {
        const double* Ak = A_turns + ((size_t)cl * K + k) * n;
        if (++cl == ncls) cl = 0;
        double t = tau_turns ? tau_turns[(size_t)j * K + k] : 0.0;
        for (int a = 0; a < n; ++a) t = fma(Ak[a], (double)xrow[j * n + a], t);
        t -= rint(t);                                   // exact range reduction to [-1/2, 1/2] turns
        double s, c;
        sincospi(2.0 * t, &s, &c);
        cs += c; sn += s;
    }
}

// candidates per block: 8 for batches (each twiddle load feeds 16 FMAs), 1 for a handful of points (x_t of a BO step, an
// initial design): the FP64 pipe is narrow, so FMAs spent on absent candidates of a partly filled group are not free
constexpr int ENC_CPB_BATCH = 8, ENC_CPB_SMALL = 1, ENC_SMALL_N = 64;
constexpr int ENC_THREADS = 256;

template <typename XT, int ENC_CPB>
__global__ void __launch_bounds__(ENC_THREADS, 2)
k_encode(const XT* __restrict__ x, int64_t N, const double* __restrict__ A_turns, const double* __restrict__ tau_turns,
         const double2* __restrict__ twiddle, int n, int K, int L, double dc, int normalize, double* __restrict__ phi,
         int deriv_axis, int ncls) {
    // deriv_axis >= 0 (L == 1): d phi / d x_a instead of phi -- the spectrum exp(i theta_k) becomes i w_ka exp(i theta_k),
    // w_ka = A+_ka / l_a, i.e. (cos, sin) -> (-w sin, w cos) and no DC term (sspspace.py:300-305)
    extern __shared__ double2 psi_p[];                  // ENC_CPB x K pairs (cos, sin) of the spectrum: one 16-byte load per frequency
    __shared__ double dc_s[ENC_CPB];                    // its DC term
    __shared__ double scale_s[ENC_CPB];                 // 1 / |phi| per candidate when normalising, else 1
    const int d = 2 * K + 1;
    const int nl = n * L;
    for (int64_t base = (int64_t)blockIdx.x * ENC_CPB; base < N; base += (int64_t)gridDim.x * ENC_CPB) {
        __syncthreads();
        for (int i = threadIdx.x; i < ENC_CPB * K; i += blockDim.x) {
            int c = i / K, k = i - c * K;
            if (base + c < N) {
                double cs, sn;
                psi_entry(x + (base + c) * nl, A_turns, tau_turns, n, K, L, k, cs, sn, ncls);
                if (deriv_axis >= 0) {
                    const double w = 6.283185307179586476925286766559 * A_turns[(size_t)k * n + deriv_axis];
                    const double c0 = cs;
                    cs = -w * sn; sn = w * c0;
                }
                psi_p[c * K + k] = make_double2(cs, sn);
                if (k == 0) dc_s[c] = deriv_axis >= 0 ? 0.0 : dc;
            }
        }
        __syncthreads();
        const double inv_d = 1.0 / (double)d;
        if (normalize) {                                // |phi|^2 = (psi_0^2 + 2 sum_k (C_k^2 + S_k^2)) / d, one warp per candidate
            const int c = threadIdx.x >> 5;
            if (c < ENC_CPB && base + c < N) {
                double ss = 0.0;
                for (int k = (threadIdx.x & 31); k < K; k += 32) {
                    const double2 pv = psi_p[c * K + k];
                    ss = fma(pv.x, pv.x, fma(pv.y, pv.y, ss));
                }
                ss = warp_sum(ss);
                if ((threadIdx.x & 31) == 0) scale_s[c] = rsqrt((dc * dc + 2.0 * ss) * inv_d);
            }
            __syncthreads();
        }
        // Output columns come in pairs: cos(2 pi (d - j) k / d) = cos(2 pi j k / d) and sin(...) changes sign, so with
        // E_j = sum_k C_k cos(2 pi j k / d) and O_j = sum_k S_k sin(2 pi j k / d):  phi_j = (psi_0 + 2 (E_j - O_j)) / d and
        // phi_{d-j} = (psi_0 + 2 (E_j + O_j)) / d -- the same two FMAs per frequency now yield two columns (the FP64 pipe is
        // what bounds this kernel).  Work items are the half-columns h = 0 .. K.
        // small batches: gridDim.y blocks share a group of candidates, each takes a slice of the half-columns
        const int nh = K + 1;
        const int j0 = (int)((long long)nh * blockIdx.y / gridDim.y), j1 = (int)((long long)nh * (blockIdx.y + 1) / gridDim.y);
        if constexpr (ENC_CPB == 1) {
            // one candidate per block: when the slice is narrower than the block, the spare warps split the frequency sum
            // (k-parts) and the partial sums are folded through shared memory in a fixed order
            __shared__ double part_e[ENC_THREADS], part_o[ENC_THREADS];
            const int ncol = j1 - j0;
            const int lanes_per = 32 * ((ncol + 31) / 32);
            const int kparts = lanes_per >= ENC_THREADS ? 1 : ENC_THREADS / lanes_per;
            for (int jb = 0; jb < ncol; jb += ENC_THREADS) {            // one trip when ncol <= ENC_THREADS
                const int part = kparts == 1 ? 0 : threadIdx.x / lanes_per;
                const int jl = jb + (kparts == 1 ? threadIdx.x : threadIdx.x % lanes_per);
                double ae = 0.0, ao = 0.0;
                if (part < kparts && jl < ncol) {
                    const int j = j0 + jl;
                    const int kb = 1 + (int)((long long)K * part / kparts), ke = 1 + (int)((long long)K * (part + 1) / kparts);
                    int idx = (int)(((long long)j * (kb - 1)) % d);
#pragma unroll 4
                    for (int k = kb; k < ke; ++k) {
                        idx += j; if (idx >= d) idx -= d;   // (j*k) mod d
                        const double2 tw = __ldg(twiddle + idx);
                        const double2 pv = psi_p[k - 1];
                        ae = fma(pv.x, tw.x, ae);
                        ao = fma(pv.y, tw.y, ao);
                    }
                }
                if (kparts > 1) {
                    __syncthreads();
                    part_e[threadIdx.x] = ae; part_o[threadIdx.x] = ao;
                    __syncthreads();
                    if (part == 0 && jl < ncol)
                        for (int q = 1; q < kparts; ++q) {
                            ae += part_e[q * lanes_per + (threadIdx.x % lanes_per)];
                            ao += part_o[q * lanes_per + (threadIdx.x % lanes_per)];
                        }
                }
                if (part == 0 && jl < ncol) {
                    const int j = j0 + jl;
                    const double sc = inv_d * (normalize ? scale_s[0] : 1.0);
                    phi[base * d + j] = (dc_s[0] + 2.0 * (ae - ao)) * sc;
                    if (j > 0) phi[base * d + d - j] = (dc_s[0] + 2.0 * (ae + ao)) * sc;
                }
            }
        } else
        // two half-columns per thread (j and j + blockDim.x): every psi value read from shared memory feeds four FMAs
        for (int j = j0 + threadIdx.x; j < j1; j += 2 * blockDim.x) {
            const int jb = j + blockDim.x;
            const bool two = jb < j1;
            double ae[ENC_CPB], ao[ENC_CPB], be[ENC_CPB], bo[ENC_CPB];
#pragma unroll
            for (int c = 0; c < ENC_CPB; ++c) { ae[c] = 0.0; ao[c] = 0.0; be[c] = 0.0; bo[c] = 0.0; }
            int idx = 0, idb = 0;
            const int jbm = two ? jb : j;               // (a lone column computes itself twice; the copy is not stored)
#pragma unroll 2
            for (int k = 1; k <= K; ++k) {
                idx += j; if (idx >= d) idx -= d;       // (j*k) mod d
                idb += jbm; if (idb >= d) idb -= d;
                const double2 tw = __ldg(twiddle + idx);
                const double2 tb = __ldg(twiddle + idb);
#pragma unroll
                for (int c = 0; c < ENC_CPB; ++c) {
                    const double2 pv = psi_p[c * K + k - 1];
                    const double pc = pv.x, ps = pv.y;
                    ae[c] = fma(pc, tw.x, ae[c]);
                    ao[c] = fma(ps, tw.y, ao[c]);
                    be[c] = fma(pc, tb.x, be[c]);
                    bo[c] = fma(ps, tb.y, bo[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < ENC_CPB; ++c)
                if (base + c < N) {
                    const double sc = inv_d * (normalize ? scale_s[c] : 1.0);
                    phi[(base + c) * d + j] = (dc_s[c] + 2.0 * (ae[c] - ao[c])) * sc;
                    if (j > 0) phi[(base + c) * d + d - j] = (dc_s[c] + 2.0 * (ae[c] + ao[c])) * sc;
                    if (two) {
                        phi[(base + c) * d + jb] = (dc_s[c] + 2.0 * (be[c] - bo[c])) * sc;
                        phi[(base + c) * d + d - jb] = (dc_s[c] + 2.0 * (be[c] + bo[c])) * sc;
                    }
                }
        }
    }
}

// Split-fp16 image of the inverse-DFT basis in operand order (B operand of the tensor-core encoder):
//   phi_j = (1/d) psi_0 + (2/d) sum_f [ cos(theta_f) cos(2 pi j f / d) - sin(theta_f) sin(2 pi j f / d) ]
// row j, column c <-> (block c/64, frequency f = 32 (c/64) + c%32, cos for c%64 < 32, sin otherwise); padding stays zero.
__global__ void k_build_basis_operand(const double2* __restrict__ twiddle, int K, int KC_pad, double scale,
                                      __half* __restrict__ Bh, __half* __restrict__ Bl) {
    const int d = 2 * K + 1;
    const int j = blockIdx.x;
    for (int c = threadIdx.x; c < KC_pad; c += blockDim.x) {
        const int b = c >> 6, i = c & 63, f = 32 * b + (i & 31);
        double v = 0.0;
        if (f <= K) {
            if (f == 0) v = (i < 32) ? 1.0 / d : 0.0;
            else {
                const double2 tw = twiddle[(int)(((long long)j * f) % d)];
                v = (i < 32) ? 2.0 * tw.x / d : -2.0 * tw.y / d;
            }
        }
        v *= scale;
        const __half h = __double2half(v);
        Bh[(size_t)j * KC_pad + c] = h;
        Bl[(size_t)j * KC_pad + c] = __double2half(v - (double)__half2float(h));
    }
}

// =====================================================================================
// basis change phi-space <-> psi-space on d-vectors (float64)
//   out[0] = s0 * sum_j in[j];  out[k] = sk * sum_j in[j] cos(w_jk);  out[K+k] = -sk * sum_j in[j] sin(w_jk)
// (s0, sk) = (1, 1) maps phi -> psi ;  (1/d, 2/d) maps m -> m' = B^T m.
// =====================================================================================
__global__ void k_analysis(const double* __restrict__ in, double* __restrict__ out, const double2* __restrict__ twiddle,
                           int K, double s0, double sk, const int* __restrict__ dest) {
    __shared__ double red[64];
    const int d = 2 * K + 1;
    const int k = blockIdx.x;                           // 0..K
    double c = 0.0, s = 0.0;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        int idx = (int)(((long long)j * k) % d);
        double2 tw = twiddle[idx];
        double v = in[j];
        c += v * tw.x; s += v * tw.y;
    }
    c = block_sum(c, red);
    s = block_sum(s, red + 32);
    if (threadIdx.x == 0) {                 // dest (optional) scatters natural psi indices to permuted ranks
        if (k == 0) out[dest ? dest[0] : 0] = s0 * c;
        else { out[dest ? dest[k] : k] = sk * c; out[dest ? dest[K + k] : K + k] = -sk * s; }
    }
}

// =====================================================================================
// float64 GEMV / rank-one kernels for the posterior (d <= a few thousand)
// =====================================================================================
// y = M x   (row-major d x d): one warp per row
__global__ void k_gemv_rows(const double* __restrict__ M, const double* __restrict__ x, double* __restrict__ y, int d) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= d) return;
    const double* r = M + (size_t)row * d;
    double acc = 0.0;
    for (int j = threadIdx.x & 31; j < d; j += 32) acc += r[j] * x[j];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) y[row] = acc;
}
// y = M^T x : block (32 x 8) handles 32 columns, strides over rows
__global__ void k_gemv_cols(const double* __restrict__ M, const double* __restrict__ x, double* __restrict__ y, int d) {
    __shared__ double part[8][33];
    int col = blockIdx.x * 32 + threadIdx.x;
    double acc = 0.0;
    if (col < d)
        for (int i = threadIdx.y; i < d; i += 8) acc += M[(size_t)i * d + col] * x[i];
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && col < d) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += part[r][threadIdx.x];
        y[col] = t;
    }
}
// Sherman-Morrison coefficient of one observation (blr.py:64-67): out[0] = beta / (1 + beta a.b), out[1] = a.b
__global__ void k_update_scalars(const double* __restrict__ a, const double* __restrict__ b, int d, double beta,
                                 double* __restrict__ out) {
    __shared__ double red[32];
    double ab = 0.0;
    for (int j = threadIdx.x; j < d; j += blockDim.x) ab += a[j] * b[j];
    ab = block_sum(ab, red);
    if (threadIdx.x == 0) { out[0] = beta / (1.0 + beta * ab); out[1] = ab; }
}
// M[i][j] += sign * coef * a[i] * b[j]   (coef read from device memory, or the immediate when coef_dev == nullptr)
__global__ void k_rank1(double* __restrict__ M, const double* __restrict__ a, const double* __restrict__ b, int d,
                        const double* __restrict__ coef_dev, double coef_imm) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= d) return;
    double c = coef_dev ? -(*coef_dev) : coef_imm;
    M[(size_t)i * d + j] += c * a[i] * b[j];
}
__global__ void k_axpy(double* __restrict__ y, const double* __restrict__ x, double a, int d) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < d) y[j] += a * x[j];
}
__global__ void k_set_diag(double* __restrict__ M, int d, double v0, double v) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < d) M[(size_t)j * d + j] = (j == 0) ? v0 : v;
}

// =====================================================================================
// Blocked right-looking Cholesky (lower, in place, float64) of the permuted psi-space covariance.
// d^3/3 flops per refresh (3.6e8 at d=1023) in three small kernels per 32-column panel.
// =====================================================================================
constexpr int CH_NB = 32;
// diagonal block: L_kk = chol(A_kk), one CTA of 32x32 threads
__global__ void k_chol_diag(double* __restrict__ A, int d, int k0) {
    __shared__ double a[CH_NB][CH_NB + 1];
    const int nb = min(CH_NB, d - k0);
    const int r = threadIdx.y, c = threadIdx.x;
    a[r][c] = (r < nb && c < nb) ? A[(size_t)(k0 + r) * d + k0 + c] : (r == c ? 1.0 : 0.0);
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        if (r == k && c == k) a[k][k] = sqrt(fmax(a[k][k], 1e-300));
        __syncthreads();
        if (c == k && r > k) a[r][k] /= a[k][k];
        __syncthreads();
        if (r > k && c > k && c <= r) a[r][c] -= a[r][k] * a[c][k];
        __syncthreads();
    }
    if (r < nb && c < nb && c <= r) A[(size_t)(k0 + r) * d + k0 + c] = a[r][c];
}
// panel below the diagonal block: L_ik = A_ik L_kk^-T, one thread per row
__global__ void k_chol_trsm(double* __restrict__ A, int d, int k0) {
    __shared__ double l[CH_NB][CH_NB + 1];
    const int nb = min(CH_NB, d - k0);
    for (int i = threadIdx.x; i < CH_NB * CH_NB; i += blockDim.x) {
        int r = i / CH_NB, c = i % CH_NB;
        l[r][c] = (r < nb && c < nb && c <= r) ? A[(size_t)(k0 + r) * d + k0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    const int row = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= d) return;
    double x[CH_NB];
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) x[c] = (c < nb) ? A[(size_t)row * d + k0 + c] : 0.0;
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) {
        double acc = x[c];
#pragma unroll
        for (int t = 0; t < c; ++t) acc -= x[t] * l[c][t];
        x[c] = acc / l[c][c];
    }
#pragma unroll
    for (int c = 0; c < CH_NB; ++c)
        if (c < nb) A[(size_t)row * d + k0 + c] = x[c];
}
// trailing update (lower tiles only): A_ij -= sum_t L_it L_jt, 32x32 tile per CTA of 32x8 threads
__global__ void k_chol_syrk(double* __restrict__ A, int d, int k0) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;
    __shared__ double li[32][CH_NB + 1], lj[32][CH_NB + 1];
    const int base = k0 + CH_NB;
    const int i0 = base + ti * 32, j0 = base + tj * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int gi = i0 + r, gj = j0 + r;
        li[r][threadIdx.x] = (gi < d) ? A[(size_t)gi * d + k0 + threadIdx.x] : 0.0;
        lj[r][threadIdx.x] = (gj < d) ? A[(size_t)gj * d + k0 + threadIdx.x] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int gi = i0 + r, gj = j0 + threadIdx.x;
        if (gi < d && gj < d && gj <= gi) {
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < CH_NB; ++t) acc += li[r][t] * lj[threadIdx.x][t];
            A[(size_t)gi * d + gj] -= acc;
        }
    }
}

// refresh of the scoring operands from the float64 factor L (lower) and m'
//   SIMT image : Rt_f32[perm[kk] * dp + j] = L[kk][j]   (row = natural psi index, column = output row j)
__global__ void k_refresh_operands(const double* __restrict__ L, const double* __restrict__ mp, const int* __restrict__ perm,
                                   int d, int dp, float* __restrict__ Rt_f32, float* __restrict__ mp_f32) {
    const int kk = blockIdx.y;
    const int row = perm[kk];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < dp; j += gridDim.x * blockDim.x)
        Rt_f32[(size_t)row * dp + j] = (j <= kk) ? (float)L[(size_t)kk * d + j] : 0.f;
    if (kk == 0)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dp; i += gridDim.x * blockDim.x)
            mp_f32[i] = (i < d) ? (float)mp[i] : 0.f;
}
// split-fp16 operands of the tcgen05 scorer in operand order (sspbo_internal.cuh):
//   Rh/Rl[j * KC_pad + col_of_rank[kk]] = split(L[kk][j] * r_scale) for j <= kk; everything else stays zero
__global__ void k_refresh_umma_operands(const double* __restrict__ L, const double* __restrict__ mp,
                                        const int* __restrict__ perm, const int* __restrict__ col_of_rank, int d,
                                        int KC_pad, double r_scale, __half* __restrict__ Rh, __half* __restrict__ Rl,
                                        float* __restrict__ mp_op) {
    __shared__ double tile[32][33];
    const int kk0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (j0 > kk0 + 31) return;                                  // strictly upper tile of L: structurally zero
    for (int r = threadIdx.y; r < 32; r += 8) {
        int kk = kk0 + r, j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (kk < d && j <= kk) ? L[(size_t)kk * d + j] * r_scale : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int j = j0 + r, kk = kk0 + threadIdx.x;
        if (j < d && kk < d) {
            double v = tile[threadIdx.x][r];
            __half h = __double2half(v);
            size_t o = (size_t)j * KC_pad + col_of_rank[kk];
            Rh[o] = h;
            Rl[o] = __double2half(v - (double)__half2float(h));
        }
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) {                  // m' in operand order (padding columns stay zero)
        int kk = kk0 + threadIdx.x;
        if (kk < d) mp_op[col_of_rank[kk]] = (float)mp[perm[kk]];
    }
}

// low-rank scoring operand: row r of V (operand order, split fp16) = y sqrt(coef) r_scale, coef = beta / (1 + beta psi.y)
__global__ void k_lr_append(const double* __restrict__ y, const double* __restrict__ coef, const int* __restrict__ col_of_rank,
                            int d, double r_scale, __half* __restrict__ vh_row, __half* __restrict__ vl_row,
                            double* __restrict__ vd_row) {
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= d) return;
    const double v0 = y[kk] * sqrt(*coef);
    vd_row[kk] = v0;                                    // float64 row in rank order (exact pass over flagged candidates)
    const double v = v0 * r_scale;
    const __half h = __double2half(v);
    const int c = col_of_rank[kk];
    vh_row[c] = h;
    vl_row[c] = __double2half(v - (double)__half2float(h));
}
// rebuilds the split-fp16 operand rows 1 .. rank from the float64 rows of V (after a posterior import)
__global__ void k_lr_rebuild(const double* __restrict__ Vd, const int* __restrict__ col_of_rank, int d, int KC_pad, double r_scale,
                             __half* __restrict__ Vh, __half* __restrict__ Vl) {
    const int j = blockIdx.y;                           // observation
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= d) return;
    const double v = Vd[(size_t)j * d + kk] * r_scale;
    const __half h = __double2half(v);
    const size_t o = (size_t)(j + 1) * KC_pad + col_of_rank[kk];
    Vh[o] = h;
    Vl[o] = __double2half(v - (double)__half2float(h));
}

// m' in operand order (padding columns stay zero), and as row 0 of the low-rank operand images: the tensor core then
// delivers mu = m' . psi as output column 0.  Row 0 carries its own power-of-two scale (|m'| max -> [512, 1024)),
// published in *mscale_out for the epilogue.  One block.
__global__ void k_mp_op(const double* __restrict__ mp, const int* __restrict__ perm, const int* __restrict__ col_of_rank,
                        int d, float* __restrict__ mp_op, __half* __restrict__ vh_row0, __half* __restrict__ vl_row0,
                        float* __restrict__ mscale_out) {
    __shared__ double red[32];
    __shared__ double scale_s;
    double mx = 0.0;
    for (int kk = threadIdx.x; kk < d; kk += blockDim.x) mx = fmax(mx, fabs(mp[kk]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmax(m, red[w]);
        const double sc = (m > 0.0 && isfinite(m)) ? exp2(floor(log2(1024.0 / m))) : 1.0;
        scale_s = sc;
        *mscale_out = (float)(1.0 / sc);
    }
    __syncthreads();
    const double sc = scale_s;
    for (int kk = threadIdx.x; kk < d; kk += blockDim.x) {
        const double m = mp[perm[kk]];
        const int c = col_of_rank[kk];
        mp_op[c] = (float)m;
        if (vh_row0) {
            const double v = m * sc;
            const __half h = __double2half(v);
            vh_row0[c] = h;
            vl_row0[c] = __double2half(v - (double)__half2float(h));
        }
    }
}

// =====================================================================================
// K2 exact engine (float64): var = 1/beta + psi^T Sp psi, mu = m' . psi straight from the float64 psi-space
// covariance.  No factorisation is needed, so it serves (a) the candidates the low-rank tcgen05 pass flags as too
// close to the observations for its cancelling form, (b) tiny candidate sets such as eval(x_t) of every BO step.
//   k_exact_partial : grid (groups, RB); CTA (g, rb) builds psi of EX_G candidates in shared memory and reduces
//                     rows [rb*d/RB, (rb+1)*d/RB) of Sp against them
//   k_exact_finish  : sums the RB partials, applies ssp_agent.py:136, writes outputs, max-combines the key
// =====================================================================================
constexpr int EX_G = 8, EX_THREADS = 256, EX_RB_MAX = 16;
constexpr int EXL_THREADS = 512;    // flagged-candidate pass: few groups in flight, so many warps per group
constexpr int EXL_JB = 4;           // rows of V per warp and pass (register blocking against the shared-memory reads of psi)

template <typename XT>
__global__ void __launch_bounds__(EX_THREADS)
k_exact_partial(const XT* __restrict__ x, int64_t N, const double* __restrict__ A_turns,
                const double* __restrict__ tau_turns, int n, int K, int L, double dc, const int* __restrict__ inv_perm,
                const int* __restrict__ perm, const double* __restrict__ Sp, const double* __restrict__ mp, int RB,
                double* __restrict__ part_q, double* __restrict__ part_mu, int ncls) {
    extern __shared__ double psi_s[];                   // EX_G x d, permuted (rank) order
    __shared__ double red[EX_G][EX_THREADS / 32];
    const int d = 2 * K + 1, nl = n * L;
    const int64_t count = N;
    const int rb = blockIdx.y;
    const int row0 = (int)((long long)d * rb / RB), row1 = (int)((long long)d * (rb + 1) / RB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * EX_G; base < count; base += (int64_t)gridDim.x * EX_G) {
        __syncthreads();
        for (int i = threadIdx.x; i < EX_G * (K + 1); i += blockDim.x) {
            const int c = i / (K + 1), f = i - c * (K + 1);
            double cs = 0.0, sn = 0.0;
            if (base + c < count) {
                const int64_t row = base + c;
                if (f == 0) cs = dc;
                else psi_entry(x + row * nl, A_turns, tau_turns, n, K, L, f - 1, cs, sn, ncls);
            }
            if (f == 0) psi_s[c * d + inv_perm[0]] = cs;
            else { psi_s[c * d + inv_perm[f]] = cs; psi_s[c * d + inv_perm[K + f]] = sn; }
        }
        __syncthreads();
        double q[EX_G];
#pragma unroll
        for (int c = 0; c < EX_G; ++c) q[c] = 0.0;
        for (int i = row0 + warp; i < row1; i += EX_THREADS / 32) {
            const double* srow = Sp + (size_t)i * d;
            double acc[EX_G];
#pragma unroll
            for (int c = 0; c < EX_G; ++c) acc[c] = 0.0;
            for (int j = lane; j < d; j += 32) {
                const double s = srow[j];
#pragma unroll
                for (int c = 0; c < EX_G; ++c) acc[c] = fma(s, psi_s[c * d + j], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < EX_G; ++c) q[c] = fma(psi_s[c * d + i], acc[c], q[c]);   // linear in the lane partials
        }
#pragma unroll
        for (int c = 0; c < EX_G; ++c) {
            const double t = warp_sum(q[c]);
            if (lane == 0) red[c][warp] = t;
        }
        __syncthreads();
        if (threadIdx.x < EX_G && base + threadIdx.x < count) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < EX_THREADS / 32; ++w) t += red[threadIdx.x][w];
            part_q[(base + threadIdx.x) * RB + rb] = t;
        }
        if (rb == 0) {                                   // mu = m' . psi and |phi|^2 = psi^T D psi, one warp per candidate
            if (warp < EX_G && base + warp < count) {
                double u = 0.0, ss = 0.0;
                for (int r = lane; r < d; r += 32) {
                    const double pv = psi_s[warp * d + r];
                    u = fma(mp[perm[r]], pv, u);
                    ss = fma(pv, pv, ss);
                }
                u = warp_sum(u);
                ss = warp_sum(ss);
                if (lane == 0) {
                    const double p0 = psi_s[warp * d + inv_perm[0]];
                    part_mu[base + warp] = u;
                    part_mu[count + base + warp] = (2.0 * ss - p0 * p0) / (double)d;
                }
            }
        }
    }
}

// Flagged candidates of the low-rank pass, re-scored in float64 in the same low-rank form (no cancellation issue at
// 1e-16): var = 1/beta + 1/alpha - sum_j (v_j . psi)^2, mu = m' . psi.  One CTA per group of EX_G candidates
// (persistent over the device-side count), psi in shared memory, one warp per row of V.
// FULL: the same pass with the float64 psi-space covariance itself, var = 1/beta + psi^T Sp psi (d rows of length d instead
// of the r rows of V): for posteriors whose low-rank rows are not available (rank > SSPBO_LR_MAX, imported state), flagged by
// the factor form of the fp16 + e4m3 scorer.  `Vd` then points at Sp and `r` is d.
// Everything one float64 pass needs (one context): k_exact_lowrank takes it by value, k_exact_lowrank_runs reads one per run
struct ExactArgs {
    const void* x; sspbo_lr::CandDesc cand; const unsigned int* list; const unsigned long long* count_ptr; unsigned int cap;
    const double *A_turns, *tau_turns; int n, K, L, ncls; double dc_value; int normalize;
    const int *inv_perm, *perm; const double* Vd; int r; const double* mp;
    double inv_beta, prior_var, lam, gamma_t; int64_t index0;
    float *mu_out, *var_out, *acq_out; unsigned long long* best;
};

// groups cta, cta + ncta, ... of the flag list
template <typename XT, bool FULL>
__device__ __forceinline__ void exact_lowrank_body(const ExactArgs& ea, int cta, int ncta) {
    const XT* __restrict__ x = (const XT*)ea.x;
    const sspbo_lr::CandDesc& cand = ea.cand;
    const unsigned int* __restrict__ list = ea.list;
    const unsigned int cap = ea.cap;
    const double* __restrict__ A_turns = ea.A_turns; const double* __restrict__ tau_turns = ea.tau_turns;
    const int n = ea.n, K = ea.K, L = ea.L, ncls = ea.ncls, normalize = ea.normalize, r = ea.r;
    const double dc_value = ea.dc_value, inv_beta = ea.inv_beta, prior_var = ea.prior_var, lam = ea.lam, gamma_t = ea.gamma_t;
    const int* __restrict__ inv_perm = ea.inv_perm; const int* __restrict__ perm = ea.perm;
    const double* __restrict__ Vd = ea.Vd; const double* __restrict__ mp = ea.mp;
    const int64_t index0 = ea.index0;
    float* __restrict__ mu_out = ea.mu_out; float* __restrict__ var_out = ea.var_out; float* __restrict__ acq_out = ea.acq_out;
    unsigned long long* __restrict__ best = ea.best;
    const unsigned long long* __restrict__ count_ptr = ea.count_ptr;
    extern __shared__ double psi_s[];                   // EX_G x d, permuted (rank) order
    __shared__ double red[EX_G + 1][EXL_THREADS / 32];
    __shared__ double xs_gen[EX_G][sspbo_lr::MAX_N];   // rows of a generated candidate set (cand.mode != 0, L == 1)
    const int d = 2 * K + 1;
    const unsigned long long c0 = *count_ptr;
    if (c0 > cap) return;                               // list overflowed: the host recomputes the call with the dense factor
    const int64_t count = (int64_t)c0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long my_best = 0ull;
    // arg-max only (no per-candidate outputs): a flagged candidate whose score bound cannot beat the best key so far is skipped
    const bool prune = best != nullptr && !mu_out && !var_out && !acq_out;
    // one group: the EX_G (or fewer) flagged rows in grp_rows
    __shared__ unsigned int grp_rows[EX_G];
    __shared__ int grp_n;
    auto score_group = [&]() {
        if (cand.mode != 0) {
            if ((int)threadIdx.x < grp_n)
                sspbo_lr::cand_row_rt(cand, n, index0 + (int64_t)grp_rows[threadIdx.x], xs_gen[threadIdx.x]);
            __syncthreads();
        }
        for (int i = threadIdx.x; i < EX_G * (K + 1); i += blockDim.x) {
            const int c = i / (K + 1), f = i - c * (K + 1);
            double cs = 0.0, sn = 0.0;
            if (c < grp_n) {
                if (f == 0) cs = dc_value;
                else if (cand.mode != 0) psi_entry(xs_gen[c], A_turns, tau_turns, n, K, L, f - 1, cs, sn, ncls);
                else psi_entry(x + (int64_t)grp_rows[c] * n * L, A_turns, tau_turns, n, K, L, f - 1, cs, sn, ncls);
            }
            if (f == 0) psi_s[c * d + inv_perm[0]] = cs;
            else { psi_s[c * d + inv_perm[f]] = cs; psi_s[c * d + inv_perm[K + f]] = sn; }
        }
        __syncthreads();
        double q[EX_G];
#pragma unroll
        for (int c = 0; c < EX_G; ++c) q[c] = 0.0;
        // EXL_JB rows per warp and pass: every psi value read from shared memory feeds EXL_JB FMAs (one row per pass is
        // bound by the shared-memory operand traffic, 8 B per FMA: ~16 FMA/clk/SM)
        for (int j0 = warp * EXL_JB; j0 < r; j0 += (EXL_THREADS / 32) * EXL_JB) {
            const double* vrow[EXL_JB];
#pragma unroll
            for (int jj = 0; jj < EXL_JB; ++jj) vrow[jj] = Vd + (size_t)min(j0 + jj, r - 1) * d;
            double acc[EXL_JB][EX_G];
#pragma unroll
            for (int jj = 0; jj < EXL_JB; ++jj)
#pragma unroll
                for (int c = 0; c < EX_G; ++c) acc[jj][c] = 0.0;
            for (int k = lane; k < d; k += 32) {
                double v[EXL_JB], ps[EX_G];
#pragma unroll
                for (int jj = 0; jj < EXL_JB; ++jj) v[jj] = vrow[jj][k];
#pragma unroll
                for (int c = 0; c < EX_G; ++c) ps[c] = psi_s[c * d + k];
#pragma unroll
                for (int jj = 0; jj < EXL_JB; ++jj)
#pragma unroll
                    for (int c = 0; c < EX_G; ++c) acc[jj][c] = fma(v[jj], ps[c], acc[jj][c]);
            }
#pragma unroll
            for (int jj = 0; jj < EXL_JB; ++jj) {
                if (j0 + jj < r) {                                                      // warp-uniform
#pragma unroll
                    for (int c = 0; c < EX_G; ++c) {
                        const double yj = warp_sum(acc[jj][c]);
                        q[c] = FULL ? fma(psi_s[c * d + j0 + jj], yj, q[c]) : fma(yj, yj, q[c]);   // psi_j (Sp psi)_j  /  (v_j . psi)^2
                    }
                }
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < EX_G; ++c) red[c][warp] = q[c];
        }
        __syncthreads();
        if (warp < grp_n) {       // one warp per candidate: mu, then the final scalars
            double u = 0.0, ss = 0.0;
            for (int rk = lane; rk < d; rk += 32) {
                const double pv = psi_s[warp * d + rk];
                u = fma(mp[perm[rk]], pv, u);
                ss = fma(pv, pv, ss);
            }
            u = warp_sum(u);
            ss = warp_sum(ss);
            if (lane == 0) {
                double qq = 0.0;
#pragma unroll
                for (int w = 0; w < EXL_THREADS / 32; ++w) qq += red[warp][w];
                const double dc = psi_s[warp * d + inv_perm[0]];
                const double nrm2 = (2.0 * ss - dc * dc) / (double)d;                   // |phi|^2 = psi^T D psi (= 1 when L == 1)
                if (normalize) { u *= rsqrt(nrm2); qq /= nrm2; }                        // phi / |phi|
                const double var = FULL ? inv_beta + qq : inv_beta + (prior_var * (normalize ? 1.0 : nrm2) - qq);   // blr.py:96
                const double acq = lam * (sqrt(var + gamma_t) - sqrt(gamma_t));         // ssp_agent.py:136
                const int64_t row = (int64_t)grp_rows[warp];
                if (mu_out) mu_out[row] = (float)u;
                if (var_out) var_out[row] = (float)var;
                if (acq_out) acq_out[row] = (float)acq;
                const unsigned long long key = sspbo_pack((float)(u + acq), (unsigned long long)(index0 + row));
                if (key > my_best) my_best = key;
            }
        }
    };
    if (!prune) {
        for (int64_t base = (int64_t)cta * EX_G; base < count; base += (int64_t)ncta * EX_G) {
            __syncthreads();
            if (threadIdx.x < EX_G && base + threadIdx.x < count) grp_rows[threadIdx.x] = list[base + threadIdx.x];
            if (threadIdx.x == 0) {
                const int64_t nhere = count - base < EX_G ? count - base : EX_G;
                grp_n = (int)nhere;
                atomicAdd(const_cast<unsigned long long*>(count_ptr) - SSPBO_STAT_CUR + SSPBO_STAT_RESCORED, (unsigned long long)nhere);
            }
            __syncthreads();
            score_group();
        }
    } else {
        // the whole block tests one flagged candidate per thread against the best key so far; the survivors are queued in
        // shared memory and scored EX_G at a time (a run of config C4 flags ~10 % of its grid -- everything close to an
        // observation -- and almost none of it can still win)
        __shared__ unsigned int queue[EXL_THREADS + EX_G];
        __shared__ int q_n, warp_cnt[EXL_THREADS / 32];
        if (threadIdx.x == 0) q_n = 0;
        __syncthreads();
        const int T = (int)blockDim.x, nwarps = T >> 5;
        for (int64_t blk = (int64_t)cta * T; blk < count; blk += (int64_t)ncta * T) {
            const int64_t i = blk + threadIdx.x;
            bool keep = false;
            unsigned int row = 0;
            if (i < count) {
                const unsigned int ob = (unsigned int)(*(volatile unsigned long long*)best >> 32);   // orderable bits of the best score
                row = list[i];
                if (ob == 0u) keep = true;                                                       // no key yet
                else if (ob != 0xffffffffu) {                                                    // (else NaN already won)
                    const float bs = __uint_as_float((ob & 0x80000000u) ? (ob & 0x7fffffffu) : ~ob);
                    const float ub = __uint_as_float(list[cap + i]);
                    keep = !(ub + 1e-3f * (fabsf(ub) + fabsf(bs)) < bs);
                }
            }
            const unsigned int bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int off = q_n, tot = 0;
            for (int w = 0; w < nwarps; ++w) { if (w < warp) off += warp_cnt[w]; tot += warp_cnt[w]; }
            if (keep) queue[off + __popc(bal & ((1u << lane) - 1u))] = row;
            __syncthreads();
            if (threadIdx.x == 0) {
                q_n += tot;
                if (tot) atomicAdd(const_cast<unsigned long long*>(count_ptr) - SSPBO_STAT_CUR + SSPBO_STAT_RESCORED, (unsigned long long)tot);
            }
            __syncthreads();
            while (q_n >= EX_G) {                                                                // (block-uniform)
                if (threadIdx.x < EX_G) grp_rows[threadIdx.x] = queue[q_n - EX_G + threadIdx.x];
                if (threadIdx.x == 0) grp_n = EX_G;
                __syncthreads();
                score_group();
                __syncthreads();
                if (threadIdx.x == 0) q_n -= EX_G;
                __syncthreads();
            }
        }
        if (q_n > 0) {
            if ((int)threadIdx.x < q_n) grp_rows[threadIdx.x] = queue[threadIdx.x];
            if (threadIdx.x == 0) grp_n = q_n;
            __syncthreads();
            score_group();
        }
    }
    if (lane == 0 && my_best && best) atomicMax(best, my_best);
}

template <typename XT, bool FULL>
__global__ void __launch_bounds__(EXL_THREADS) k_exact_lowrank(const ExactArgs ea) {
    exact_lowrank_body<XT, FULL>(ea, blockIdx.x, gridDim.x);
}

// Lock-step of many independent runs (BASELINE config C4): the float64 pass of EVERY run in one launch -- CTA b walks the
// flag lists of runs b, b + gridDim.x, ... (a run flags a handful of candidates: its own observations) -- followed, per run,
// by the counter fold of k_flag_fold and the status word of k_batch_status.
// work items (run, part): the groups of a run's flag list are dealt to `parts` CTAs, so a few hundred runs still fill the machine
template <typename XT>
__global__ void __launch_bounds__(EXL_THREADS) k_exact_lowrank_runs(const ExactArgs* __restrict__ runs, int R, int parts) {
    for (int item = blockIdx.x; item < R * parts; item += gridDim.x) {
        __syncthreads();
        exact_lowrank_body<XT, false>(runs[item / parts], item % parts, parts);
    }
}
// ... and then, one thread per run, the status word of k_batch_status and the counters cleared for the next lock-step
__global__ void k_batch_status_runs(const ExactArgs* __restrict__ runs, int R, unsigned long long* __restrict__ status) {
    const int run = blockIdx.x * blockDim.x + threadIdx.x;
    if (run >= R) return;
    unsigned long long* st = const_cast<unsigned long long*>(runs[run].count_ptr) - SSPBO_STAT_CUR;
    status[run] = ((st[SSPBO_STAT_OVERFLOW] || st[SSPBO_STAT_CUR] > runs[run].cap) ? 1ull : 0ull) | (st[SSPBO_STAT_OOR] ? 2ull : 0ull);
    for (int i = 0; i < SSPBO_STAT_COUNT; ++i) st[i] = 0ull;
}
// folds the per-call flag counter into the totals (after the exact pass has consumed it)
__global__ void k_flag_fold(unsigned long long* __restrict__ stats, unsigned int cap, unsigned long long scored) {
    const unsigned long long c = stats[SSPBO_STAT_CUR];
    stats[SSPBO_STAT_FLAGGED] += (c < cap ? c : cap);
    stats[SSPBO_STAT_SCORED] += scored;
    stats[SSPBO_STAT_CUR] = 0ull;
}

__global__ void k_exact_finish(int64_t N, int RB, const double* __restrict__ part_q, const double* __restrict__ part_mu,
                               int normalize, double inv_beta, double lam, double gamma_t, int64_t index0, float* __restrict__ mu_out,
                               float* __restrict__ var_out, float* __restrict__ acq_out, unsigned long long* __restrict__ best,
                               double* __restrict__ out64) {
    const int64_t count = N;
    unsigned long long my_best = 0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double q = 0.0;
        for (int r = 0; r < RB; ++r) q += part_q[i * RB + r];
        double mu = part_mu[i];
        if (normalize) { const double nrm2 = part_mu[count + i]; mu *= rsqrt(nrm2); q /= nrm2; }   // phi / |phi|
        const double var = inv_beta + q;                                        // blr.py:96
        const double acq = lam * (sqrt(var + gamma_t) - sqrt(gamma_t));         // ssp_agent.py:136
        const int64_t row = i;
        if (out64) { out64[row] = mu; out64[64 + row] = var; out64[128 + row] = acq; }   // sspbo_eval_host: float64 as the reference returns
        if (mu_out) mu_out[row] = (float)mu;
        if (var_out) var_out[row] = (float)var;
        if (acq_out) acq_out[row] = (float)acq;
        const unsigned long long key = sspbo_pack((float)(mu + acq), (unsigned long long)(index0 + row));
        if (key > my_best) my_best = key;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, my_best, o);
        if (other > my_best) my_best = other;
    }
    if ((threadIdx.x & 31) == 0 && my_best && best) atomicMax(best, my_best);
}

// =====================================================================================
// BLR.predict on explicit phi (float64): var = 1/beta + phi^T S phi ; mu = m . phi
// =====================================================================================
constexpr int PRED_CPB = 4;
__global__ void __launch_bounds__(256)
k_predict(const double* __restrict__ phi, int64_t N, const double* __restrict__ S, const double* __restrict__ m, int d,
          double inv_beta, double* __restrict__ mu, double* __restrict__ var) {
    extern __shared__ double ph[];                      // PRED_CPB x d
    __shared__ double red[64];
    for (int64_t base = (int64_t)blockIdx.x * PRED_CPB; base < N; base += (int64_t)gridDim.x * PRED_CPB) {
        __syncthreads();
        for (int i = threadIdx.x; i < PRED_CPB * d; i += blockDim.x) {
            int c = i / d;
            ph[i] = (base + c < N) ? phi[(base + c) * d + (i - c * d)] : 0.0;
        }
        __syncthreads();
        double q[PRED_CPB], u[PRED_CPB];
#pragma unroll
        for (int c = 0; c < PRED_CPB; ++c) { q[c] = 0.0; u[c] = 0.0; }
        // blr.py:96: einsum('ij,ij->i', phi, phi @ S.T):  (phi S^T)_j = sum_i phi_i S[j][i]
        for (int j = threadIdx.x; j < d; j += blockDim.x) {
            double t[PRED_CPB];
#pragma unroll
            for (int c = 0; c < PRED_CPB; ++c) t[c] = 0.0;
            for (int i = 0; i < d; ++i) {
                double s = S[(size_t)i * d + j];        // S symmetric to ~1e-13; column read is coalesced
#pragma unroll
                for (int c = 0; c < PRED_CPB; ++c) t[c] += ph[c * d + i] * s;
            }
            double mj = m[j];
#pragma unroll
            for (int c = 0; c < PRED_CPB; ++c) { q[c] += t[c] * ph[c * d + j]; u[c] += mj * ph[c * d + j]; }
        }
#pragma unroll
        for (int c = 0; c < PRED_CPB; ++c) {
            double qq = block_sum(q[c], red);
            double uu = block_sum(u[c], red + 32);
            if (threadIdx.x == 0 && base + c < N) { var[base + c] = inv_beta + qq; mu[base + c] = uu; }
        }
    }
}

// =====================================================================================
// K4 from-set decode: sims = sample_ssps @ unit^T ; first-index argmax per query
// =====================================================================================
__global__ void k_unit_rows(const double* __restrict__ q, double* __restrict__ unit, int d) {
    __shared__ double red[32];
    const double* r = q + (size_t)blockIdx.x * d;
    double s = 0.0;
    for (int j = threadIdx.x; j < d; j += blockDim.x) s += r[j] * r[j];
    s = block_sum(s, red);
    double nrm = fmax(sqrt(s), 1e-6);                   // sspspace.py:352
    for (int j = threadIdx.x; j < d; j += blockDim.x) unit[(size_t)blockIdx.x * d + j] = r[j] / nrm;
}
// one warp per sample row; sims[qi * M + i]
__global__ void k_sims(const double* __restrict__ ssps, int64_t M, const double* __restrict__ unit, int nq, int d,
                       double* __restrict__ sims) {
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const double* r = ssps + row * d;
    for (int qi = 0; qi < nq; ++qi) {
        const double* u = unit + (size_t)qi * d;
        double acc = 0.0;
        for (int j = threadIdx.x & 31; j < d; j += 32) acc += r[j] * u[j];
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) sims[(size_t)qi * M + row] = acc;
    }
}
__device__ __forceinline__ bool better_f64(double v, int64_t i, double bv, int64_t bi) {
    // np.argmax: first maximum; NaN compares as the maximum
    bool vn = isnan(v), bn = isnan(bv);
    if (vn || bn) return vn && (!bn || i < bi);
    return (v > bv) || (v == bv && i < bi);
}
__global__ void k_argmax_rows(const double* __restrict__ sims, int64_t M, int64_t* __restrict__ idx_out) {
    __shared__ double bv_s[32];
    __shared__ long long bi_s[32];
    const double* r = sims + (size_t)blockIdx.x * M;
    double bv = -INFINITY; int64_t bi = INT64_MAX;
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x)
        if (better_f64(r[i], i, bv, bi)) { bv = r[i]; bi = i; }
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        long long oi = __shfl_xor_sync(0xffffffffu, (long long)bi, o);
        if (better_f64(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { bv_s[w] = bv; bi_s[w] = bi; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        bv = (l < nw) ? bv_s[l] : -INFINITY; bi = (l < nw) ? bi_s[l] : INT64_MAX;
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            long long oi = __shfl_xor_sync(0xffffffffu, (long long)bi, o);
            if (better_f64(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (l == 0) idx_out[blockIdx.x] = bi;
    }
}

// =====================================================================================
// K4b direct-optim decode (sspspace.py:358-375): for each query, maximise f(x) = phi(x) . unit(query) over x in the domain
// bounds, starting from x0 (the from-set answer).  The reference runs scipy's L-BFGS-B with numerical gradients on
// -f; here f, its gradient and its Hessian are analytic in the Fourier domain,
//   f(x) = u'_0 + sum_k ( c_k cos theta_k(x) + s_k sin theta_k(x) ),  u' = B^T unit(query),  theta_k = 2 pi A_k . x,
// and a projected, damped Newton iteration (float64, one CTA per query) lands on the same local maximum.
// =====================================================================================
constexpr int DO_THREADS = 128, DO_MAXN = 8, DO_NRED = 1 + DO_MAXN + DO_MAXN * (DO_MAXN + 1) / 2;

__global__ void __launch_bounds__(DO_THREADS)
k_decode_optim(const double* __restrict__ queries, const double* __restrict__ x0, const double* __restrict__ A_turns,
               const double2* __restrict__ twiddle, const double* __restrict__ bounds, int n, int K, int max_iter,
               double* __restrict__ x_out) {
    extern __shared__ double sh[];                      // uc (K), us (K), red (warps x DO_NRED)
    double* uc = sh; double* us = sh + K; double* red = sh + 2 * K;
    __shared__ double xs[DO_MAXN], xtrial[DO_MAXN], fgH[DO_NRED], u0_s, nrm_s;
    __shared__ int done_s;
    const int d = 2 * K + 1, q = blockIdx.x;
    const double* qv = queries + (size_t)q * d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = DO_THREADS / 32;
    // unit(query): norm with the reference's floor (sspspace.py:352)
    {
        double ss = 0.0;
        for (int j = threadIdx.x; j < d; j += DO_THREADS) ss += qv[j] * qv[j];
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < nw; ++w) t += red[w]; nrm_s = fmax(sqrt(t), 1e-6); }
        __syncthreads();
    }
    // u' = B^T unit(query): u'_0 = (1/d) sum_j u_j ; c_k = (2/d) sum_j u_j cos(w_jk) ; s_k = -(2/d) sum_j u_j sin(w_jk)
    for (int k = threadIdx.x; k <= K; k += DO_THREADS) {
        double c = 0.0, sn = 0.0;
        for (int j = 0; j < d; ++j) {
            const double2 tw = twiddle[(int)(((long long)j * k) % d)];
            c = fma(qv[j], tw.x, c); sn = fma(qv[j], tw.y, sn);
        }
        c /= nrm_s; sn /= nrm_s;
        if (k == 0) u0_s = c / d;
        else { uc[k - 1] = 2.0 * c / d; us[k - 1] = -2.0 * sn / d; }
    }
    if (threadIdx.x < n) { xs[threadIdx.x] = x0[(size_t)q * n + threadIdx.x]; xtrial[threadIdx.x] = xs[threadIdx.x]; }
    if (threadIdx.x == 0) done_s = 0;
    __syncthreads();

    const double twopi = 6.283185307179586476925286766559;
    double f_cur = -1e300, step_scale = 1.0;
    double delta[DO_MAXN];                              // thread 0: current Newton direction
    const int nh = n * (n + 1) / 2, nred = 1 + n + nh;
    for (int it = 0; it < max_iter; ++it) {
        // ---- f, gradient, Hessian at xtrial (all threads) ----
        double acc[DO_NRED];
#pragma unroll
        for (int i = 0; i < DO_NRED; ++i) acc[i] = 0.0;
        for (int k = threadIdx.x; k < K; k += DO_THREADS) {
            const double* Ak = A_turns + (size_t)k * n;
            double t = 0.0;
            for (int a = 0; a < n; ++a) t = fma(Ak[a], xtrial[a], t);
            t -= rint(t);
            double sv, cv;
            sincospi(2.0 * t, &sv, &cv);
            const double val = uc[k] * cv + us[k] * sv;          // contribution to f
            const double dth = -uc[k] * sv + us[k] * cv;         // d/dtheta
            acc[0] += val;
            int h = 1 + n;
            for (int a = 0; a < n; ++a) {
                const double wa = twopi * Ak[a];
                acc[1 + a] = fma(dth, wa, acc[1 + a]);
                for (int b = 0; b <= a; ++b) { acc[h] = fma(-val * wa, twopi * Ak[b], acc[h]); ++h; }
            }
        }
        for (int i = 0; i < nred; ++i) {
            const double t = warp_sum(acc[i]);
            if (lane == 0) red[warp * DO_NRED + i] = t;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < nred; ++i) { double t = 0.0; for (int w = 0; w < nw; ++w) t += red[w * DO_NRED + i]; fgH[i] = t; }
            const double f_trial = fgH[0] + u0_s;
            if (f_trial >= f_cur) {                     // accept the trial point, build the next Newton step there
                f_cur = f_trial; step_scale = 1.0;
                for (int a = 0; a < n; ++a) xs[a] = xtrial[a];
                // solve (-H + mu I) delta = g by Cholesky, raising mu until -H + mu I is positive definite
                double M[DO_MAXN][DO_MAXN], g[DO_MAXN];
                double hmax = 0.0;
                for (int i = 0; i < nh; ++i) hmax = fmax(hmax, fabs(fgH[1 + n + i]));
                // active set: a coordinate sitting on a bound with the gradient pushing outwards stays there, the Newton
                // system is solved on the free coordinates only
                bool act[DO_MAXN];
                for (int a = 0; a < n; ++a) {
                    g[a] = fgH[1 + a];
                    act[a] = bounds && ((xs[a] >= bounds[2 * a + 1] && g[a] > 0.0) || (xs[a] <= bounds[2 * a] && g[a] < 0.0));
                    if (act[a]) g[a] = 0.0;
                }
                double mu = 0.0;
                bool ok = false;
                for (int attempt = 0; attempt < 40 && !ok; ++attempt) {
                    int h = 1 + n;
                    for (int a = 0; a < n; ++a)
                        for (int b = 0; b <= a; ++b) {
                            M[a][b] = -fgH[h++];
                            if (a == b) M[a][b] += mu;
                            if (act[a] || act[b]) M[a][b] = (a == b) ? 1.0 : 0.0;
                        }
                    ok = true;
                    for (int a = 0; a < n && ok; ++a) {
                        for (int b = 0; b <= a; ++b) {
                            double sum = M[a][b];
                            for (int c = 0; c < b; ++c) sum -= M[a][c] * M[b][c];
                            if (a == b) { if (sum <= 1e-14 * (hmax + mu)) { ok = false; break; } M[a][a] = sqrt(sum); }
                            else M[a][b] = sum / M[b][b];
                        }
                    }
                    if (!ok) mu = (mu == 0.0) ? 1e-3 * (hmax + 1e-300) : mu * 4.0;
                }
                for (int a = 0; a < n; ++a) {           // forward substitution L y = g
                    double sum = g[a];
                    for (int c = 0; c < a; ++c) sum -= M[a][c] * delta[c];
                    delta[a] = ok ? sum / M[a][a] : 0.0;
                }
                for (int a = n - 1; a >= 0; --a) {      // back substitution L^T delta = y
                    double sum = delta[a];
                    for (int c = a + 1; c < n; ++c) sum -= M[c][a] * delta[c];
                    delta[a] = ok ? sum / M[a][a] : g[a] / (hmax + 1e-300);
                }
            } else {
                step_scale *= 0.5;                      // rejected: shorter step from the last accepted point
            }
            double moved = 0.0, scale = 0.0;
            for (int a = 0; a < n; ++a) {
                double xn = xs[a] + step_scale * delta[a];
                if (bounds) xn = fmin(fmax(xn, bounds[2 * a]), bounds[2 * a + 1]);
                moved = fmax(moved, fabs(xn - xs[a]));
                scale = fmax(scale, fabs(xs[a]));
                xtrial[a] = xn;
            }
            if (moved <= 1e-13 * (1.0 + scale) || step_scale < 1e-12) done_s = 1;
        }
        __syncthreads();
        if (done_s) break;
    }
    if (threadIdx.x < n) x_out[(size_t)q * n + threadIdx.x] = xs[threadIdx.x];
}

// =====================================================================================
// K5 binding: circular convolution out[r][i] = sum_j a[r][j] b[r][(i - j) mod d]   (float64, direct)
// =====================================================================================
__global__ void k_bind(const double* __restrict__ a, int qa, const double* __restrict__ b, int qb, int d,
                       double* __restrict__ out) {
    extern __shared__ double sh[];                      // a row, then b row
    double* as = sh; double* bs = sh + d;
    const int r = blockIdx.y;
    const double* ar = a + (size_t)(qa == 1 ? 0 : r) * d;
    const double* br = b + (size_t)(qb == 1 ? 0 : r) * d;
    for (int j = threadIdx.x; j < d; j += blockDim.x) { as[j] = ar[j]; bs[j] = br[j]; }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
        double acc = 0.0;
        int k = i;                                       // (i - j) mod d, walked downwards
        for (int j = 0; j < d; ++j) {
            acc = fma(as[j], bs[k], acc);
            k = (k == 0) ? d - 1 : k - 1;
        }
        out[(size_t)r * d + i] = acc;
    }
}

// =====================================================================================
// Gram matrix G = Phi Phi^T of n encoded rows (float64): the only O(d) work of the K-fold length-scale search
// (agents/ssp_traj_agent.py:113-143) once every fold's BLR is written in its dual form
// =====================================================================================
__global__ void k_gram(const double* __restrict__ phi, int n, int d, double* __restrict__ G) {
    __shared__ double red[32];
    const int i = blockIdx.y, j = blockIdx.x;
    if (j > i) return;
    const double* a = phi + (size_t)i * d;
    const double* b = phi + (size_t)j * d;
    double acc = 0.0;
    for (int k = threadIdx.x; k < d; k += blockDim.x) acc = fma(a[k], b[k], acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) { G[(size_t)i * n + j] = acc; G[(size_t)j * n + i] = acc; }
}

// =====================================================================================
// candidate generators
// =====================================================================================
__global__ void k_fill_uniform(uint64_t seed, int64_t first_elem, int64_t count, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = (uint64_t)(first_elem + i) + (seed << 40) + 0x9E3779B97F4A7C15ull;   // splitmix64 finaliser
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        out[i] = (float)(z >> 40) * 5.9604644775390625e-8f;                               // 24 bits * 2^-24
    }
}
struct GridDesc { int n; int counts[16]; int offs[16]; };
__global__ void k_grid_points(const double* __restrict__ axes, GridDesc g, int64_t start, int64_t count,
                              double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = start + i;
        // flattened C-order of a meshgrid('xy') array of shape (N1, N0, N2, ..., N_{n-1})
        for (int a = g.n - 1; a >= 2; --a) {
            int ia = (int)(p % g.counts[a]); p /= g.counts[a];
            out[i * g.n + a] = axes[g.offs[a] + ia];
        }
        int i0 = (int)(p % g.counts[0]); p /= g.counts[0];
        out[i * g.n + 0] = axes[g.offs[0] + i0];
        if (g.n >= 2) out[i * g.n + 1] = axes[g.offs[1] + (int)p];
    }
}

// =====================================================================================
// host side
// =====================================================================================
// Context buffers come from the device's stream-ordered pool (kept, not returned to the driver, when freed): a context is ~30
// buffers, and thousands of contexts are created for config C4 -- plain cudaMalloc + synchronous memset cost 20 ms per context.
static void dev_pool_keep(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[device] = true;
}
// (a few buffers are plain cudaMalloc allocations: those go back through cudaFree)
template <typename T>
static void dev_free(T** p) {
    if (!*p) return;
    if (cudaFreeAsync(*p, 0) != cudaSuccess) { cudaGetLastError(); cudaFree(*p); }
    *p = nullptr;
}
template <typename T>
static int dev_alloc(sspbo_ctx* ctx, T** p, size_t count) {
    if (*p) dev_free(p);
    dev_pool_keep(ctx->device);
    SSPBO_CUDA(ctx, cudaMallocAsync((void**)p, count * sizeof(T), 0));
    SSPBO_CUDA(ctx, cudaMemsetAsync(*p, 0, count * sizeof(T), 0));
    SSPBO_CUDA(ctx, cudaStreamSynchronize(0));            // allocated and cleared before any (non-blocking) caller stream can touch it
    return SSPBO_OK;
}


extern "C" int sspbo_abi_version(void) { return 1; }

extern "C" int sspbo_ctx_create(int device, sspbo_ctx** out) {
    if (!out) return SSPBO_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SSPBO_ECUDA;
    sspbo_ctx* ctx = new (std::nothrow) sspbo_ctx();
    if (!ctx) return SSPBO_EINVAL;
    ctx->device = device;
    cudaSetDevice(device);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->cc_major, cudaDevAttrComputeCapabilityMajor, device);
    if (cudaMalloc((void**)&ctx->scal, 16 * sizeof(double)) != cudaSuccess) { delete ctx; return SSPBO_ECUDA; }
    *out = ctx;
    return SSPBO_OK;
}

static void free_blr(sspbo_ctx* ctx) {
    cudaDeviceSynchronize();                              // work on non-blocking streams may still use the buffers (pool frees are stream-ordered on stream 0 only)
    sspbo_f8_release(ctx);
    timing_release(ctx);
    if (ctx->batch_pin) { cudaFreeHost(ctx->batch_pin); cudaFree(ctx->batch_dev); cudaEventDestroy(ctx->batch_event); ctx->batch_pin = nullptr; ctx->batch_dev = nullptr; }
    if (ctx->sbatch_pin) { cudaFreeHost(ctx->sbatch_pin); cudaFree(ctx->sbatch_dev); cudaEventDestroy(ctx->sbatch_event); ctx->sbatch_pin = nullptr; ctx->sbatch_dev = nullptr; }
    dev_free(&ctx->S); dev_free(&ctx->Sinv); dev_free(&ctx->b); dev_free(&ctx->m);
    dev_free(&ctx->R); dev_free(&ctx->Sp); dev_free(&ctx->perm); dev_free(&ctx->inv_perm); dev_free(&ctx->col_of_rank);
    dev_free(&ctx->mp); dev_free(&ctx->Rt_f32); dev_free(&ctx->mp_f32);
    dev_free(&ctx->Rh); dev_free(&ctx->Rl); dev_free(&ctx->mp_op);
    dev_free(&ctx->Vh); dev_free(&ctx->Vl); dev_free(&ctx->Vd); dev_free(&ctx->mscale); ctx->lr_rank = 0; ctx->lr_valid = false;
    dev_free(&ctx->psi_ring); dev_free(&ctx->syn_rows); dev_free(&ctx->bpsi); dev_free(&ctx->lr_gh);
    dev_free(&ctx->ck_vec); dev_free(&ctx->ck_mp_op); dev_free(&ctx->ck_row0); dev_free(&ctx->ck_mscale); ctx->ck_valid = false;
    ctx->sp_applied = 0; ctx->s_applied = 0; ctx->sinv_pending = 0; ctx->bm_stale = false; ctx->bpsi_stale = false;
    dev_free(&ctx->v); dev_free(&ctx->psi); dev_free(&ctx->y); dev_free(&ctx->z); dev_free(&ctx->phi_tmp);
    ctx->blr_ready = false; ctx->has_beta = false;
}

extern "C" void sspbo_ctx_destroy(sspbo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    sspbo_umma_release(ctx);
    sspbo_lr_release(ctx);
    sspbo_dist_release(ctx);
    free_blr(ctx);
    dev_free(&ctx->A_turns); dev_free(&ctx->A_turns_f32); dev_free(&ctx->tau_turns); dev_free(&ctx->tau_turns_f32);
    dev_free(&ctx->twiddle); dev_free(&ctx->scal); dev_free(&ctx->Aq_op); dev_free(&ctx->tauq_op);
    dev_free(&ctx->A32_op); dev_free(&ctx->Af_op); dev_free(&ctx->Af_gen); dev_free(&ctx->Bh_enc); dev_free(&ctx->Bl_enc); dev_free(&ctx->flag_idx); dev_free(&ctx->stats); dev_free(&ctx->ex_part);
    dev_free(&ctx->aa_work);
    if (ctx->gen_rows) cudaFree(ctx->gen_rows);
    if (ctx->eval_dev) cudaFree(ctx->eval_dev);
    if (ctx->obs_pin) { cudaFreeHost(ctx->obs_pin); cudaFree(ctx->obs_x); }
    if (ctx->obs_event) cudaEventDestroy(ctx->obs_event);
    dev_free(&ctx->obs_phi); dev_free(&ctx->obs_out);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    delete ctx;
}

extern "C" const char* sspbo_last_error(const sspbo_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
extern "C" int sspbo_last_engine(const sspbo_ctx* ctx) { return ctx ? ctx->last_engine : 0; }
extern "C" int64_t sspbo_launch_count(const sspbo_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- kernel timing (bench.py's roofline figure: per-launch duration of the dominant kernel, measured on its stream) ----
static const int TIMING_MAX_CALLS = 4096;
static void timing_release(sspbo_ctx* ctx) {
    auto* ev = (std::vector<cudaEvent_t>*)ctx->timing_events;
    if (ev) { for (cudaEvent_t e : *ev) cudaEventDestroy(e); delete ev; }
    ctx->timing_events = nullptr; ctx->timing_used = 0;
}
// event `which` (0 before the scorer kernel, 1 after it, 2 after the flagged pass) of the score call in flight
static void timing_mark(sspbo_ctx* ctx, int which, cudaStream_t st) {
    if (!ctx->timing || ctx->timing_used >= TIMING_MAX_CALLS) return;
    auto* ev = (std::vector<cudaEvent_t>*)ctx->timing_events;
    if (!ev) { ev = new std::vector<cudaEvent_t>(); ctx->timing_events = ev; }
    const size_t idx = (size_t)ctx->timing_used * 3 + which;
    while (ev->size() <= idx) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; ev->push_back(e); }
    cudaEventRecord((*ev)[idx], st);
    if (which == 2) ++ctx->timing_used;
}
extern "C" int sspbo_set_timing(sspbo_ctx* ctx, int enable) {
    if (!ctx) return SSPBO_EINVAL;
    ctx->timing = enable != 0; ctx->timing_used = 0;
    return SSPBO_OK;
}
extern "C" int sspbo_timing_read(sspbo_ctx* ctx, double* scorer_ms, double* flagged_ms, int64_t* calls, int reset, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    double a = 0.0, b = 0.0;
    auto* ev = (std::vector<cudaEvent_t>*)ctx->timing_events;
    SSPBO_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < ctx->timing_used && ev; ++i) {
        float t0 = 0.f, t1 = 0.f;
        SSPBO_CUDA(ctx, cudaEventElapsedTime(&t0, (*ev)[3 * i], (*ev)[3 * i + 1]));
        SSPBO_CUDA(ctx, cudaEventElapsedTime(&t1, (*ev)[3 * i + 1], (*ev)[3 * i + 2]));
        a += t0; b += t1;
    }
    if (scorer_ms) *scorer_ms = a;
    if (flagged_ms) *flagged_ms = b;
    if (calls) *calls = ctx->timing_used;
    if (reset) ctx->timing_used = 0;
    return SSPBO_OK;
}

static int build_p32_tables(sspbo_ctx* ctx);

extern "C" int sspbo_space_set(sspbo_ctx* ctx, const double* A_plus, const double* inv_ls, int K, int n) {
    return sspbo_space_set_classes(ctx, A_plus, inv_ls, K, n, 1);
}

/* n_cls phase matrices of one shape (A_plus: n_cls x K x n, inv_ls: n_cls x n): waypoint j of a candidate row is encoded with
 * table j % n_cls -- SSPMultiAgent with one RandomSSPSpace per agent (ssp_multi_agent.py:58-66, same_agt_x_space=False), rows
 * laid out (traj_len, n_agents, x_dim).  Followed by sspbo_space_set_waypoints with L a multiple of n_cls. */
extern "C" int sspbo_space_set_classes(sspbo_ctx* ctx, const double* A_plus, const double* inv_ls, int K, int n, int n_cls) {
    if (!ctx) return SSPBO_EINVAL;
    if (!A_plus || !inv_ls || K < 1 || n < 1 || n > 16) SSPBO_FAIL(ctx, SSPBO_EINVAL, "space_set: need K>=1, 1<=n<=16");
    if (n_cls < 1 || n_cls > 16) SSPBO_FAIL(ctx, SSPBO_EINVAL, "space_set_classes: 1 <= n_cls <= 16");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->K) SSPBO_CUDA(ctx, cudaDeviceSynchronize());     // re-programming: the old tables may still be read on other streams
    if (K != ctx->K || !ctx->has_factor) { free_blr(ctx); sspbo_umma_release(ctx); }
    if (K != ctx->K) ctx->enc_K = 0;
    const int d = 2 * K + 1;
    std::vector<double> At((size_t)n_cls * K * n);
    std::vector<float> Af(At.size());
    double mx = 0.0;
    for (int k = 0; k < n_cls * K; ++k) {
        const double* il = inv_ls + (size_t)(k / K) * n;
        double rs = 0.0;
        for (int a = 0; a < n; ++a) {
            double t = A_plus[(size_t)k * n + a] * il[a] / kTwoPi;
            At[(size_t)k * n + a] = t; Af[(size_t)k * n + a] = (float)t; rs += fabs(t);
        }
        mx = fmax(mx, rs);
    }
    std::vector<double2> tw(d);
    for (int i = 0; i < d; ++i) {
        // exact-ish twiddles: reduce the angle to the first octant in integer arithmetic via sincospi
        double frac = 2.0 * (double)i / (double)d;
        tw[i].x = cospi(frac); tw[i].y = sinpi(frac);
    }
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->A_turns, At.size()))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A_turns_f32, At.size()))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->twiddle, (size_t)d))) return rc;
    SSPBO_CUDA(ctx, cudaMemcpy(ctx->A_turns, At.data(), At.size() * sizeof(double), cudaMemcpyHostToDevice));
    SSPBO_CUDA(ctx, cudaMemcpy(ctx->A_turns_f32, Af.data(), Af.size() * sizeof(float), cudaMemcpyHostToDevice));
    SSPBO_CUDA(ctx, cudaMemcpy(ctx->twiddle, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));
    dev_free(&ctx->tau_turns); dev_free(&ctx->tau_turns_f32); dev_free(&ctx->tauq_op); dev_free(&ctx->tauf_op);
    {   // frequency table of the tcgen05 scorer: row f = A_turns of frequency f (f = 0: DC, f > K: padding) -> zeros
        int kc, bn, nc, nj;
        sspbo_umma_dims(K, &kc, &bn, &nc, &nj);
        std::vector<long long> op((size_t)n_cls * (kc / 2) * n, 0ll);       // one table per class, back to back
        ctx->umma_range_ok = true;
        for (int c = 0; c < n_cls; ++c)
            for (int f = 1; f <= K; ++f)
                for (int a = 0; a < n; ++a) {
                    double t = At[((size_t)c * K + f - 1) * n + a];
                    if (!(fabs(t) < 2147483648.0)) { ctx->umma_range_ok = false; t = 0.0; }
                    op[((size_t)c * (kc / 2) + f) * n + a] = llrint(t * 4294967296.0);   // Q31.32 turns per unit of x
                }
        if ((rc = dev_alloc(ctx, &ctx->Aq_op, op.size()))) return rc;
        SSPBO_CUDA(ctx, cudaMemcpy(ctx->Aq_op, op.data(), op.size() * sizeof(long long), cudaMemcpyHostToDevice));
    }
    ctx->K = K; ctx->n = n; ctx->d = d; ctx->L = 1; ctx->n_cls = n_cls; ctx->dp = sspbo_pad64(d); ctx->max_turns = mx;
    if (n_cls > 1) { ctx->L = n_cls; ctx->dc = (double)n_cls; ctx->normalize = false; }   // one waypoint per table until set_waypoints
    double amax = 0.0;
    for (size_t i = 0; i < At.size(); ++i) amax = fmax(amax, fabs(At[i]));
    ctx->a_absmax = amax;
    ctx->host_A_turns.assign(At.begin(), At.end());
    ctx->p32_disabled = false;
    return build_p32_tables(ctx);
}

// 32-bit fixed-point phase tables of the fast generator: A in Q.fa, x in Q.fx with fa <= 31 - ceil(log2 |A|max),
// fx <= 31 - ceil(log2 |x|max) and fa + fx = 55; the 64-bit sum of the n products holds the phase in turns at bit 55,
// so bits [32, 55) (the low 23 bits of the high word) are the top 23 bits of the phase modulo one turn.  Enabled when
// the worst-case rounding of the two operands plus the 2^-23 truncation stays below 2^-20 turns (6e-6 rad);
// otherwise the 64-bit path (Q31.32 x Q31.32) is used.
static int build_p32_tables(sspbo_ctx* ctx) {
    ctx->p32_ok = false; ctx->pf32_ok = false;
    if (ctx->K == 0 || !(ctx->x_absmax > 0.0) || !(ctx->a_absmax > 0.0)) return SSPBO_OK;
    {   // float32 path: theta in radians by an FMA chain; partial sums stay below 32 pi (16 turns), so every rounding is
        // below 2^-18 rad and the MUFU range reduction (an RZ multiply by 1/2pi) adds < 1.2e-5 rad
        if (ctx->max_turns * ctx->x_absmax <= 15.5) {           // + half a turn of timestep phase in trajectory mode
            int kc, bn, nc, nj;
            sspbo_umma_dims(ctx->K, &kc, &bn, &nc, &nj);
            const int ncls = ctx->n_cls, nfq = kc / 2;
            std::vector<float> op((size_t)ncls * nfq * ctx->n, 0.f);
            for (int c = 0; c < ncls; ++c)
                for (int f = 1; f <= ctx->K; ++f)
                    for (int a = 0; a < ctx->n; ++a)
                        op[((size_t)c * nfq + f) * ctx->n + a] = (float)(ctx->host_A_turns[((size_t)c * ctx->K + f - 1) * ctx->n + a] * kTwoPi);
            int rc;
            if ((rc = dev_alloc(ctx, &ctx->Af_op, op.size()))) return rc;
            SSPBO_CUDA(ctx, cudaMemcpy(ctx->Af_op, op.data(), op.size() * sizeof(float), cudaMemcpyHostToDevice));
            // the same table in the layout the generators read (axis-major inside every block of 16 frequencies): the grouped
            // scorer of many runs reads it from global memory instead of staging one table per launch in shared memory
            std::vector<float> gen(op.size(), 0.f);
            const int NFq = 16;
            for (int f = 0; f < ncls * nfq; ++f)                              // (nfq is a multiple of 16: blocks never straddle classes)
                for (int a = 0; a < ctx->n; ++a)
                    gen[((size_t)(f / NFq) * ctx->n + a) * NFq + (f % NFq)] = op[(size_t)f * ctx->n + a];
            if ((rc = dev_alloc(ctx, &ctx->Af_gen, gen.size()))) return rc;
            SSPBO_CUDA(ctx, cudaMemcpy(ctx->Af_gen, gen.data(), gen.size() * sizeof(float), cudaMemcpyHostToDevice));
            ctx->pf32_ok = true;
        }
    }
    int ia = 0, ix = 0;
    while (ldexp(1.0, ia) <= ctx->a_absmax && ia < 31) ++ia;     // 2^ia > |A|max
    while (ldexp(1.0, ix) <= ctx->x_absmax && ix < 31) ++ix;     // 2^ix > |x|max
    // the kernel reads the phase at a fixed position: fa + fx = 55 (bits [32, 55) of the 64-bit sum are its top 23 bits)
    const int fa_max = 31 - ia, fx_max = 31 - ix;
    if (fa_max + fx_max < 55) return SSPBO_OK;
    int fa = -1; double err = 1e300;
    for (int cand = 55 - fx_max; cand <= fa_max; ++cand) {
        const double e = ctx->n * (ctx->x_absmax * ldexp(1.0, -(cand + 1)) + ctx->a_absmax * ldexp(1.0, -(55 - cand + 1)));
        if (e < err) { err = e; fa = cand; }
    }
    const int fx = 55 - fa;
    if (fa < 0 || err + ldexp(1.0, -23) > ldexp(1.0, -20)) return SSPBO_OK;
    int kc, bn, nc, nj;
    sspbo_umma_dims(ctx->K, &kc, &bn, &nc, &nj);
    std::vector<int> op((size_t)ctx->n_cls * (kc / 2) * ctx->n, 0);
    for (int c = 0; c < ctx->n_cls; ++c)
        for (int f = 1; f <= ctx->K; ++f)
            for (int a = 0; a < ctx->n; ++a)
            {
                long long q = llrint(ctx->host_A_turns[((size_t)c * ctx->K + f - 1) * ctx->n + a] * ldexp(1.0, fa));
                op[((size_t)c * (kc / 2) + f) * ctx->n + a] = (int)(q > 2147483647ll ? 2147483647ll : (q < -2147483647ll ? -2147483647ll : q));
            }
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->A32_op, op.size()))) return rc;
    SSPBO_CUDA(ctx, cudaMemcpy(ctx->A32_op, op.data(), op.size() * sizeof(int), cudaMemcpyHostToDevice));
    ctx->p32_fa = fa; ctx->p32_fx = fx; ctx->p32_ok = true;
    return SSPBO_OK;
}

extern "C" int sspbo_space_set_xbound(sspbo_ctx* ctx, double x_absmax) {
    if (!ctx) return SSPBO_EINVAL;
    if (!(x_absmax >= 0.0) || !isfinite(x_absmax)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "space_set_xbound: bound must be finite and >= 0");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->x_absmax = x_absmax;
    ctx->p32_disabled = false;
    return build_p32_tables(ctx);
}

extern "C" int sspbo_space_set_traj(sspbo_ctx* ctx, const double* tau_host, int L) {
    return sspbo_space_set_waypoints(ctx, tau_host, nullptr, L, 0);
}

extern "C" int sspbo_space_set_waypoints(sspbo_ctx* ctx, const double* tau_host, const int* dc_sign_host, int L, int normalize) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "space_set_traj before space_set");
    if (L < 1 || L > 64) SSPBO_FAIL(ctx, SSPBO_EINVAL, "traj length must be in [1, 64]");
    if (L % ctx->n_cls) SSPBO_FAIL(ctx, SSPBO_EINVAL, "waypoints per row (%d) must be a multiple of the %d frequency tables", L, ctx->n_cls);
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    dev_free(&ctx->tau_turns); dev_free(&ctx->tau_turns_f32); dev_free(&ctx->tauq_op); dev_free(&ctx->tauf_op);
    ctx->L = L;
    ctx->normalize = normalize != 0;
    ctx->dc = (double)L;
    if (dc_sign_host) {
        ctx->dc = 0.0;
        for (int j = 0; j < L; ++j) {
            if (dc_sign_host[j] != 1 && dc_sign_host[j] != -1) SSPBO_FAIL(ctx, SSPBO_EINVAL, "DC signs must be +1 or -1");
            ctx->dc += dc_sign_host[j];
        }
    }
    if (!tau_host) {
        if (L != 1 || dc_sign_host) SSPBO_FAIL(ctx, SSPBO_EINVAL, "tau is required when L > 1");
        return SSPBO_OK;
    }
    std::vector<double> t((size_t)L * ctx->K);
    std::vector<float> tf((size_t)L * ctx->K);
    for (size_t i = 0; i < t.size(); ++i) {
        double turns = tau_host[i] / kTwoPi;
        turns -= rint(turns);
        t[i] = turns; tf[i] = (float)turns;
    }
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->tau_turns, t.size()))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->tau_turns_f32, t.size()))) return rc;
    SSPBO_CUDA(ctx, cudaMemcpy(ctx->tau_turns, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    SSPBO_CUDA(ctx, cudaMemcpy(ctx->tau_turns_f32, tf.data(), tf.size() * sizeof(float), cudaMemcpyHostToDevice));
    {   // Q0.64 phase offsets in the frequency order of the tcgen05 scorer (f = 0: DC, f > K: padding -> 0)
        int kc, bn, nc, nj;
        sspbo_umma_dims(ctx->K, &kc, &bn, &nc, &nj);
        const int nf = kc / 2;
        std::vector<unsigned long long> tq((size_t)L * nf, 0ull);
        for (int j = 0; j < L; ++j)
            for (int f = 1; f <= ctx->K; ++f) {
                double fr = t[(size_t)j * ctx->K + f - 1];              // in [-1/2, 1/2]
                fr -= floor(fr);                                       // [0, 1)
                long double scaled = (long double)fr * 18446744073709551616.0L;
                tq[(size_t)j * nf + f] = (scaled >= 18446744073709551615.0L) ? 0ull : (unsigned long long)scaled;
            }
        if (dc_sign_host)                                              // DC phasor of waypoint j: +1 (0 turns) or -1 (1/2 turn)
            for (int j = 0; j < L; ++j) if (dc_sign_host[j] < 0) tq[(size_t)j * nf] = 1ull << 63;
        if ((rc = dev_alloc(ctx, &ctx->tauq_op, tq.size()))) return rc;
        SSPBO_CUDA(ctx, cudaMemcpy(ctx->tauq_op, tq.data(), tq.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        std::vector<float> tfl((size_t)L * nf, 0.f);                      // same, float32 radians in [-pi, pi]
        for (int j = 0; j < L; ++j)
            for (int f = 1; f <= ctx->K; ++f) tfl[(size_t)j * nf + f] = (float)(t[(size_t)j * ctx->K + f - 1] * kTwoPi);
        if (dc_sign_host)
            for (int j = 0; j < L; ++j) if (dc_sign_host[j] < 0) tfl[(size_t)j * nf] = (float)(0.5 * kTwoPi);
        if ((rc = dev_alloc(ctx, &ctx->tauf_op, tfl.size()))) return rc;
        SSPBO_CUDA(ctx, cudaMemcpy(ctx->tauf_op, tfl.data(), tfl.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return SSPBO_OK;
}

extern "C" int sspbo_space_dims(const sspbo_ctx* ctx, int* d, int* K, int* n, int* L) {
    if (!ctx) return SSPBO_EINVAL;
    if (d) *d = ctx->d; if (K) *K = ctx->K; if (n) *n = ctx->n; if (L) *L = ctx->L;
    return SSPBO_OK;
}

static int encode_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, double* phi_out, int deriv_axis, void* stream);

extern "C" int sspbo_encode(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, double* phi_out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "encode before space_set");
    if (N < 0 || (N > 0 && (!x || !phi_out))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "encode: null pointer");
    if (N == 0) return SSPBO_OK;
    return encode_launch(ctx, x, x_dtype, N, phi_out, -1, stream);
}

/* SSPSpace.encode_and_deriv (sspspace.py:278-306): phi (N x d) and grad, n planes of N x d with
 * grad[a][s][j] = d phi_j(x_s) / d x_a = Re IDFT_k( i (A_ka / l_a) exp(i A_k . x_s / l) )_j.  Plain spaces only (L == 1). */
extern "C" int sspbo_encode_deriv(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, double* phi_out, double* grad_out,
                                  void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "encode_deriv before space_set");
    if (ctx->L != 1 || ctx->normalize) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "encode_deriv: plain SSP spaces only (no trajectory mode)");
    if (N < 0 || (N > 0 && (!x || !phi_out || !grad_out))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "encode_deriv: null pointer");
    if (N == 0) return SSPBO_OK;
    int rc = encode_launch(ctx, x, x_dtype, N, phi_out, -1, stream);
    for (int a = 0; a < ctx->n && !rc; ++a)
        rc = encode_launch(ctx, x, x_dtype, N, grad_out + (size_t)a * N * ctx->d, a, stream);
    return rc;
}

static int encode_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, double* phi_out, int deriv_axis, void* stream) {
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int cpb = N <= ENC_SMALL_N ? ENC_CPB_SMALL : ENC_CPB_BATCH;
    size_t smem = (size_t)cpb * ctx->K * sizeof(double2);
    int64_t blocks64 = (N + cpb - 1) / cpb;
    int blocks = (int)(blocks64 < (int64_t)ctx->sm_count * 8 ? blocks64 : (int64_t)ctx->sm_count * 8);
    int slices = 1;                                      // a handful of candidates (e.g. x_t of a BO step): spread the columns
    while (slices < 16 && (int64_t)blocks * slices < 2 * ctx->sm_count && (ctx->K + 1) / (slices * 2) >= 32) slices *= 2;
    const dim3 grid(blocks, slices);
#define SSPBO_ENC_LAUNCH(XT, CPB)                                                                                          \
    do {                                                                                                                   \
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_encode<XT, CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        k_encode<XT, CPB><<<grid, ENC_THREADS, smem, st>>>((const XT*)x, N, ctx->A_turns, ctx->tau_turns, ctx->twiddle,     \
                                                           ctx->n, ctx->K, ctx->L, ctx->dc, ctx->normalize ? 1 : 0, phi_out, \
                                                           deriv_axis, ctx->n_cls);                                       \
    } while (0)
    if (x_dtype == SSPBO_F32) {
        if (cpb == ENC_CPB_SMALL) SSPBO_ENC_LAUNCH(float, ENC_CPB_SMALL); else SSPBO_ENC_LAUNCH(float, ENC_CPB_BATCH);
    } else if (x_dtype == SSPBO_F64) {
        if (cpb == ENC_CPB_SMALL) SSPBO_ENC_LAUNCH(double, ENC_CPB_SMALL); else SSPBO_ENC_LAUNCH(double, ENC_CPB_BATCH);
    } else SSPBO_FAIL(ctx, SSPBO_EINVAL, "encode: unknown dtype %d", x_dtype);
#undef SSPBO_ENC_LAUNCH
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

/* Random-Fourier-feature map of agents/rff_agent.py:153-156: phi[s][j] = scale cos(W_j . x_s + B_j), scale = sqrt(2 alpha / m)
 * (float64 throughout; the BLR over these features is a raw-feature context, sspbo_blr_init_dim).  One thread per output. */
template <typename XT>
__global__ void k_rff_encode(const XT* __restrict__ x, int64_t N, int n, const double* __restrict__ W, const double* __restrict__ B,
                             int m, double scale, double* __restrict__ phi) {
    const int64_t total = N * m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / m;
        const int j = (int)(i - row * m);
        double t = B[j];
        for (int a = 0; a < n; ++a) t = fma(W[(size_t)j * n + a], (double)x[row * n + a], t);
        phi[i] = scale * cos(t);
    }
}

extern "C" int sspbo_rff_encode(sspbo_ctx* ctx, const double* W_dev, const double* B_dev, int m, int n, double scale, const void* x,
                                int x_dtype, int64_t N, double* phi_out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (m < 1 || n < 1 || N < 0 || !W_dev || !B_dev || (N > 0 && (!x || !phi_out))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "rff_encode: bad argument");
    if (N == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t total = N * m;
    const int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sm_count * 16 ? (total + 255) / 256 : (int64_t)ctx->sm_count * 16);
    if (x_dtype == SSPBO_F32) k_rff_encode<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, N, n, W_dev, B_dev, m, scale, phi_out);
    else if (x_dtype == SSPBO_F64) k_rff_encode<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)x, N, n, W_dev, B_dev, m, scale, phi_out);
    else SSPBO_FAIL(ctx, SSPBO_EINVAL, "rff_encode: unknown dtype %d", x_dtype);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

/* SSPSpace.encode on the tensor cores (float32 result): phi = psi B^T with the split-fp16 inverse-DFT basis as the B
 * operand of the dense tcgen05 kernel; psi is generated on chip exactly as in the scorer.  Accuracy ~1e-6 of max |phi|
 * (the float64 sspbo_encode gives 1e-13); the kernel is bound by the tensor pipe (3 x 2 x 1024^2 flops per candidate at
 * d = 1023), phi is written once: 4 * pitch bytes per candidate.  phi_out_dev: N x pitch float32, pitch = sspbo_encode_pitch. */
extern "C" int sspbo_encode_pitch(sspbo_ctx* ctx) {
    if (!ctx || ctx->K == 0) return 0;
    int kc, bn, nc, nj;
    sspbo_umma_dims(ctx->K, &kc, &bn, &nc, &nj);
    return nj;
}

extern "C" int sspbo_encode_f32(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, float* phi_out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "encode_f32 before space_set");
    if (N < 0 || (N > 0 && (!x || !phi_out))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "encode_f32: null pointer");
    if (x_dtype != SSPBO_F32 && x_dtype != SSPBO_F64) SSPBO_FAIL(ctx, SSPBO_EINVAL, "encode_f32: unknown dtype %d", x_dtype);
    if (N == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int kc, bn, nc, nj, rc;
    sspbo_umma_dims(ctx->K, &kc, &bn, &nc, &nj);
    if (ctx->n > 8 || !ctx->umma_range_ok || nc > 8 || (ctx->L != 1 && (ctx->n > 3 || !ctx->tauq_op)) || ctx->normalize || ctx->n_cls != 1)
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "encode_f32: shape outside the tcgen05 kernel (n=%d, d=%d, L=%d)", ctx->n, ctx->d, ctx->L);
    if (!ctx->blr_ready || !ctx->has_factor) {           // geometry normally set by blr_init
        ctx->KC_pad = kc; ctx->BN = bn; ctx->n_chunks = nc; ctx->NJ_pad = nj;
        if (!ctx->mp_op && (rc = dev_alloc(ctx, &ctx->mp_op, (size_t)kc))) return rc;
    }
    if (!ctx->Bh_enc || ctx->enc_K != ctx->K) {
        if ((rc = dev_alloc(ctx, &ctx->Bh_enc, (size_t)nj * kc)) || (rc = dev_alloc(ctx, &ctx->Bl_enc, (size_t)nj * kc))) return rc;
        ctx->enc_scale = exp2(floor(log2(1024.0 / (2.0 / ctx->d))));
        k_build_basis_operand<<<ctx->d, 256, 0, st>>>(ctx->twiddle, ctx->K, kc, ctx->enc_scale, ctx->Bh_enc, ctx->Bl_enc);
        SSPBO_LAUNCH_CHECK(ctx);
        ctx->enc_K = ctx->K;
    }
    return sspbo_encode_umma_launch(ctx, x, x_dtype, N, phi_out, nj, st);
}

// ------------------------------------------------------------------------------------ BLR
static int blr_alloc(sspbo_ctx* ctx, int d, double alpha, bool with_factor) {
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    free_blr(ctx);
    const int dp = sspbo_pad64(d);
    const size_t dd = (size_t)d * d, dpp = (size_t)dp * dp;
    int rc;
    if (with_factor) sspbo_umma_dims(ctx->K, &ctx->KC_pad, &ctx->BN, &ctx->n_chunks, &ctx->NJ_pad);
    if ((rc = dev_alloc(ctx, &ctx->S, dd)) || (rc = dev_alloc(ctx, &ctx->Sinv, dd)) ||
        (rc = dev_alloc(ctx, &ctx->b, (size_t)d)) || (rc = dev_alloc(ctx, &ctx->m, (size_t)d)) ||
        (rc = dev_alloc(ctx, &ctx->v, (size_t)d)) || (rc = dev_alloc(ctx, &ctx->z, (size_t)d)))
        return rc;
    if (with_factor) {
        if ((rc = dev_alloc(ctx, &ctx->R, dd)) || (rc = dev_alloc(ctx, &ctx->Sp, dd)) ||
            (rc = dev_alloc(ctx, &ctx->perm, (size_t)d)) || (rc = dev_alloc(ctx, &ctx->inv_perm, (size_t)d)) ||
            (rc = dev_alloc(ctx, &ctx->col_of_rank, (size_t)d)) || (rc = dev_alloc(ctx, &ctx->mp, (size_t)d)) ||
            (rc = dev_alloc(ctx, &ctx->Rt_f32, dpp)) || (rc = dev_alloc(ctx, &ctx->mp_f32, (size_t)dp)) ||
            (rc = dev_alloc(ctx, &ctx->Rh, (size_t)ctx->NJ_pad * ctx->KC_pad)) ||
            (rc = dev_alloc(ctx, &ctx->Rl, (size_t)ctx->NJ_pad * ctx->KC_pad)) ||
            (rc = dev_alloc(ctx, &ctx->mp_op, (size_t)ctx->KC_pad)) ||
            (rc = dev_alloc(ctx, &ctx->Vh, (size_t)(SSPBO_LR_MAX + 1) * ctx->KC_pad)) ||
            (rc = dev_alloc(ctx, &ctx->Vl, (size_t)(SSPBO_LR_MAX + 1) * ctx->KC_pad)) ||
            (rc = dev_alloc(ctx, &ctx->mscale, (size_t)1)) ||
            (rc = dev_alloc(ctx, &ctx->Vd, (size_t)SSPBO_LR_MAX * d)) ||
            (rc = dev_alloc(ctx, &ctx->psi, (size_t)d)) || (rc = dev_alloc(ctx, &ctx->y, (size_t)d)) ||
            (rc = dev_alloc(ctx, &ctx->phi_tmp, (size_t)d)))
            return rc;
    }
    ctx->alpha = alpha; ctx->has_beta = false; ctx->beta = 0.0; ctx->d = d; ctx->dp = dp;
    int blocks = (d + 255) / 256;
    k_set_diag<<<blocks, 256>>>(ctx->S, d, 1.0 / alpha, 1.0 / alpha);                 // blr.py:33 pinv(alpha I)
    SSPBO_LAUNCH_CHECK(ctx);
    k_set_diag<<<blocks, 256>>>(ctx->Sinv, d, alpha, alpha);                           // blr.py:30
    SSPBO_LAUNCH_CHECK(ctx);
    if (with_factor) {
        // permutation of the psi components into operand-column order (cos/sin interleaved per 32-frequency block)
        {
            const int K = ctx->K;
            std::vector<int> perm, col;
            for (int c = 0; c < ctx->KC_pad; ++c) {
                int b = c >> 6, i = c & 63, f = 32 * b + (i & 31);
                int pidx = (f > K) ? -1 : (i < 32 ? f : (f == 0 ? -1 : K + f));
                if (pidx >= 0) { perm.push_back(pidx); col.push_back(c); }
            }
            if ((int)perm.size() != d) SSPBO_FAIL(ctx, SSPBO_EINVAL, "internal: operand permutation has %d of %d entries", (int)perm.size(), d);
            std::vector<int> inv(d);
            for (int r = 0; r < d; ++r) inv[perm[r]] = r;
            SSPBO_CUDA(ctx, cudaMemcpy(ctx->perm, perm.data(), d * sizeof(int), cudaMemcpyHostToDevice));
            SSPBO_CUDA(ctx, cudaMemcpy(ctx->inv_perm, inv.data(), d * sizeof(int), cudaMemcpyHostToDevice));
            SSPBO_CUDA(ctx, cudaMemcpy(ctx->col_of_rank, col.data(), d * sizeof(int), cudaMemcpyHostToDevice));
            ctx->host_col_of_rank = col;
            for (int kb = 0; kb < 32; ++kb) {
                int cnt = 0;
                for (int r = 0; r < d; ++r) cnt += (col[r] < 64 * (kb + 1));
                ctx->nz_rows[kb] = (short)cnt;
            }
            for (int ch = 0; ch < 8; ++ch) {
                int row = ch * ctx->BN;
                ctx->kb_start[ch] = (ch < ctx->n_chunks && row < d) ? col[row] / 64 : 0;
            }
        }
        // Sp0 = P (B^T B / alpha) P^T, B^T B = diag(1/d, 2/d, ..., 2/d); the DC component keeps rank 0
        k_set_diag<<<blocks, 256>>>(ctx->Sp, d, 1.0 / (d * alpha), 2.0 / (d * alpha));
        SSPBO_LAUNCH_CHECK(ctx);
        // Sp only shrinks (in the Loewner order), so |L_ij| <= sqrt(||Sp0||_2) for ever: one fixed power-of-two scale
        double rmax = sqrt(2.0 / (d * alpha));
        ctx->r_scale = exp2(floor(log2(1024.0 / rmax)));
        ctx->factor_dirty = true;
    }
    ctx->has_factor = with_factor;
    ctx->lr_rank = 0; ctx->lr_valid = with_factor; ctx->lr_disabled = false; ctx->lr_flag_frac = 0.0;
    ctx->sp_applied = 0; ctx->s_applied = 0; ctx->sinv_pending = 0; ctx->bm_stale = false; ctx->bpsi_stale = true; ctx->ck_valid = false;
    if (with_factor && !ctx->stats) {
        if ((rc = dev_alloc(ctx, &ctx->stats, (size_t)SSPBO_STAT_COUNT)) ||
            (rc = dev_alloc(ctx, &ctx->flag_idx, (size_t)2 * SSPBO_FLAG_CAP)))
            return rc;
    }
    ctx->blr_ready = true; ctx->operands_dirty = true; ctx->f8_factor_dirty = true;
    SSPBO_CUDA(ctx, cudaDeviceSynchronize());
    return SSPBO_OK;
}

extern "C" int sspbo_blr_init(sspbo_ctx* ctx, double alpha) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_init before space_set");
    if (!(alpha > 0)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "alpha must be positive");
    return blr_alloc(ctx, 2 * ctx->K + 1, alpha, true);
}

extern "C" int sspbo_blr_init_dim(sspbo_ctx* ctx, int d, double alpha) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K != 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_init_dim: this ctx has an SSP space; use blr_init");
    if (d < 1 || d > 16384) SSPBO_FAIL(ctx, SSPBO_EINVAL, "blr_init_dim: d out of range");
    if (!(alpha > 0)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "alpha must be positive");
    return blr_alloc(ctx, d, alpha, false);
}

extern "C" int sspbo_blr_set_beta(sspbo_ctx* ctx, double beta) {
    if (!ctx) return SSPBO_EINVAL;
    if (!(beta > 0) || !isfinite(beta)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "beta must be positive and finite");
    if (ctx->has_beta && beta != ctx->beta && ctx->sinv_pending > 0) {   // pending S_inv terms carry the old precision
        SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
        const int rc = sspbo_ensure_sinv(ctx, 0);
        if (rc) return rc;
        SSPBO_CUDA(ctx, cudaStreamSynchronize(0));
    }
    ctx->beta = beta; ctx->has_beta = true;
    return SSPBO_OK;
}
extern "C" int sspbo_blr_get_beta(const sspbo_ctx* ctx, double* beta, int* is_set) {
    if (!ctx) return SSPBO_EINVAL;
    if (beta) *beta = ctx->beta; if (is_set) *is_set = ctx->has_beta ? 1 : 0;
    return SSPBO_OK;
}

// ---- the state the low-rank updates leave behind (k_blr_update_lr, sspbo_update.cu), brought up to date on demand ----
// M[i][j] -= sum_r rows[r][i] rows[r][j]   (Sp from the rows of V; S from their phi-space images)
__global__ void k_rows_fold(double* __restrict__ M, const double* __restrict__ rows, int r0, int r1, int d) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    double acc = M[(size_t)i * d + j];
    for (int r = r0; r < r1; ++r) acc = fma(-__ldg(rows + (size_t)r * d + i), __ldg(rows + (size_t)r * d + j), acc);
    M[(size_t)i * d + j] = acc;
}
// S_inv[i][j] += beta phi_r[i] phi_r[j] for the observations of the ring, in their order (blr.py:60)
__global__ void k_sinv_flush(double* __restrict__ Sinv, const double* __restrict__ ring, int count, double beta, int d) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    double acc = Sinv[(size_t)i * d + j];
    for (int r = 0; r < count; ++r) acc = fma(beta * __ldg(ring + (size_t)r * d + i), __ldg(ring + (size_t)r * d + j), acc);
    Sinv[(size_t)i * d + j] = acc;
}
// out[row][j] = s0 w_0 + sk sum_k (w_Ck cos(2 pi j k / d) - w_Sk sin(2 pi j k / d)), w = row of `in` read through `src` (rank of
// every natural index; null: natural order).  (s0, sk) = (1/d, 2/d): phi = B psi (sspspace.py:275) and b = B b_psi;
// (1, 1): B D^-1 w, the phi-space image of a row of V and m from m'.
__global__ void k_synth_rows(const double* __restrict__ in, int K, const int* __restrict__ src, const double2* __restrict__ twiddle,
                             double s0, double sk, double* __restrict__ out) {
    extern __shared__ double w_s[];                     // d, natural order
    const int d = 2 * K + 1;
    const double* w = in + (size_t)blockIdx.y * d;
    for (int k = threadIdx.x; k < d; k += blockDim.x) w_s[k] = w[src ? src[k] : k];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    double acc = 0.0;
    int idx = 0;
    for (int k = 1; k <= K; ++k) {
        idx += j; if (idx >= d) idx -= d;
        const double2 tw = __ldg(twiddle + idx);
        acc = fma(w_s[k], tw.x, fma(-w_s[K + k], tw.y, acc));
    }
    out[(size_t)blockIdx.y * d + j] = fma(sk, acc, s0 * w_s[0]);
}

static int synth_rows(sspbo_ctx* ctx, const double* in, int rows, const int* src, double s0, double sk, double* out, cudaStream_t st) {
    const int d = ctx->d;
    k_synth_rows<<<dim3((d + 255) / 256, rows), 256, (size_t)d * sizeof(double), st>>>(in, ctx->K, src, ctx->twiddle, s0, sk, out);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

int sspbo_ensure_sp(sspbo_ctx* ctx, cudaStream_t st) {
    if (!ctx->has_factor || ctx->sp_applied >= ctx->lr_rank) return SSPBO_OK;
    const int d = ctx->d;
    k_rows_fold<<<dim3((d + 255) / 256, d), 256, 0, st>>>(ctx->Sp, ctx->Vd, ctx->sp_applied, ctx->lr_rank, d);
    SSPBO_LAUNCH_CHECK(ctx);
    ctx->sp_applied = ctx->lr_rank;
    return SSPBO_OK;
}

int sspbo_ensure_bm(sspbo_ctx* ctx, cudaStream_t st) {
    if (!ctx->has_factor || !ctx->bm_stale) return SSPBO_OK;
    int rc;
    const double d = (double)ctx->d;
    if ((rc = synth_rows(ctx, ctx->bpsi, 1, ctx->inv_perm, 1.0 / d, 2.0 / d, ctx->b, st))) return rc;     // b = B b_psi
    if ((rc = synth_rows(ctx, ctx->mp, 1, nullptr, 1.0, 1.0, ctx->m, st))) return rc;                      // m = B D^-1 m'
    ctx->bm_stale = false;
    return SSPBO_OK;
}

int sspbo_ensure_s(sspbo_ctx* ctx, cudaStream_t st) {
    int rc;
    if ((rc = sspbo_ensure_bm(ctx, st))) return rc;
    if (!ctx->has_factor || ctx->s_applied >= ctx->lr_rank) return SSPBO_OK;
    const int d = ctx->d;
    while (ctx->s_applied < ctx->lr_rank) {             // S -= u u^T, u = B D^-1 (row of V): a ring's worth of rows per pass over S
        const int cnt = ctx->lr_rank - ctx->s_applied < SSPBO_SINV_RING ? ctx->lr_rank - ctx->s_applied : SSPBO_SINV_RING;
        if ((rc = synth_rows(ctx, ctx->Vd + (size_t)ctx->s_applied * d, cnt, ctx->inv_perm, 1.0, 1.0, ctx->syn_rows, st))) return rc;
        k_rows_fold<<<dim3((d + 255) / 256, d), 256, 0, st>>>(ctx->S, ctx->syn_rows, 0, cnt, d);
        SSPBO_LAUNCH_CHECK(ctx);
        ctx->s_applied += cnt;
    }
    return SSPBO_OK;
}

int sspbo_ensure_sinv(sspbo_ctx* ctx, cudaStream_t st) {
    if (ctx->sinv_pending <= 0) return SSPBO_OK;
    const int d = ctx->d;
    int rc;
    if ((rc = synth_rows(ctx, ctx->psi_ring, ctx->sinv_pending, nullptr, 1.0 / d, 2.0 / d, ctx->syn_rows, st))) return rc;   // phi = B psi
    k_sinv_flush<<<dim3((d + 255) / 256, d), 256, 0, st>>>(ctx->Sinv, ctx->syn_rows, ctx->sinv_pending, ctx->beta, d);
    SSPBO_LAUNCH_CHECK(ctx);
    ctx->sinv_pending = 0;
    return SSPBO_OK;
}

int sspbo_ensure_all(sspbo_ctx* ctx, cudaStream_t st) {
    int rc;
    if ((rc = sspbo_ensure_s(ctx, st)) || (rc = sspbo_ensure_sp(ctx, st)) || (rc = sspbo_ensure_sinv(ctx, st))) return rc;
    return SSPBO_OK;
}

int sspbo_lr_update_prepare(sspbo_ctx* ctx, cudaStream_t st, double** ring_row) {
    int rc;
    const size_t d = ctx->d;
    if (!ctx->psi_ring) {
        if ((rc = dev_alloc(ctx, &ctx->psi_ring, (size_t)SSPBO_SINV_RING * d)) || (rc = dev_alloc(ctx, &ctx->syn_rows, (size_t)SSPBO_SINV_RING * d)) ||
            (rc = dev_alloc(ctx, &ctx->bpsi, d)) || (rc = dev_alloc(ctx, &ctx->lr_gh, (size_t)2 * SSPBO_LR_MAX + 32)))
            return rc;
        ctx->bpsi_stale = true;
    }
    if (ctx->bpsi_stale) {                              // b_psi = D^-1 B^T b after eager updates / an import (b is current then)
        k_analysis<<<ctx->K + 1, 128, 0, st>>>(ctx->b, ctx->bpsi, ctx->twiddle, ctx->K, 1.0, 1.0, ctx->inv_perm);
        SSPBO_LAUNCH_CHECK(ctx);
        ctx->bpsi_stale = false;
    }
    if (ctx->sinv_pending >= SSPBO_SINV_RING && (rc = sspbo_ensure_sinv(ctx, st))) return rc;
    *ring_row = ctx->psi_ring + (size_t)ctx->sinv_pending * d;
    return SSPBO_OK;
}

void sspbo_eager_update_done(sspbo_ctx* ctx) {
    ctx->sp_applied = ctx->lr_rank; ctx->s_applied = ctx->lr_rank;
    ctx->bpsi_stale = true;
}

extern "C" int sspbo_blr_update(sspbo_ctx* ctx, const double* phis, const double* ts_host, int k, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_update before blr_init");
    if (!ctx->has_beta) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_update before blr_set_beta");
    if (k < 0 || (k > 0 && (!phis || !ts_host))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "blr_update: null pointer");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int d = ctx->d, K = ctx->K;
    const double beta = ctx->beta;
    dim3 g1((d + 255) / 256, d);
    bool m_current = false;                              // m, m' already refreshed by the last fused launch
    bool tail_done = false;                              // ... and the operand-order copies of m' too (low-rank kernel)
    for (int i = 0; i < k; ++i) {
        const double* phi = phis + (size_t)i * d;
        if (!ctx->unfused_update && ctx->has_factor && ctx->lr_valid) {      // low-rank form: O(rank d), one cluster launch
            if (ctx->lr_rank >= SSPBO_LR_MAX) ctx->lr_valid = false;
            else {
                sspbo_ctx* one[1] = {ctx};
                const double* ph[1] = {phi};
                const int rc = sspbo_blr_update_lr(one, 1, nullptr, ph, ts_host + i, st, nullptr,
                                                   (ctx->want_obs && k == 1) ? ctx->obs_out : nullptr);
                if (rc == SSPBO_OK) { m_current = true; tail_done = true; continue; }
                if (rc != SSPBO_EUNSUPPORTED) return rc;
            }
        }
        {   // the kernels below keep the d x d state and the phi-space vectors current themselves: fold in what is pending
            const int rc = sspbo_ensure_all(ctx, st);
            if (rc) return rc;
            tail_done = false;
        }
        if (!ctx->unfused_update) {                      // one cooperative launch per observation
            __half *vh = nullptr, *vl = nullptr; double* vd = nullptr;
            bool append = false;
            if (ctx->has_factor && ctx->lr_valid) {
                if (ctx->lr_rank >= SSPBO_LR_MAX) ctx->lr_valid = false;
                else {
                    const size_t off = (size_t)(ctx->lr_rank + 1) * ctx->KC_pad;
                    vh = ctx->Vh + off; vl = ctx->Vl + off; vd = ctx->Vd + (size_t)ctx->lr_rank * d; append = true;
                }
            }
            const int rc = sspbo_blr_update_fused(ctx, phi, ts_host[i], i == k - 1, vh, vl, vd,
                                                  (ctx->want_obs && k == 1) ? ctx->obs_out : nullptr, st);
            if (rc == SSPBO_OK) { if (append) ctx->lr_rank++; sspbo_eager_update_done(ctx); m_current = (i == k - 1); continue; }
            if (rc != SSPBO_EUNSUPPORTED) return rc;
            // nothing was written: fall through to the unfused kernels, for this and every later observation
            ctx->unfused_update = true;
        }
        m_current = false;
        // --- phi-space posterior, blr.py:60-70 ---
        k_gemv_rows<<<(d + 7) / 8, 256, 0, st>>>(ctx->S, phi, ctx->v, d);                       // v = S phi
        SSPBO_LAUNCH_CHECK(ctx);
        k_gemv_cols<<<(d + 31) / 32, dim3(32, 8), 0, st>>>(ctx->S, phi, ctx->z, d);              // z = S^T phi
        SSPBO_LAUNCH_CHECK(ctx);
        k_update_scalars<<<1, 256, 0, st>>>(phi, ctx->v, d, beta, ctx->scal);
        SSPBO_LAUNCH_CHECK(ctx);
        // --- psi-space covariance (permuted): psi = D^-1 B^T phi ; y = Sp psi ; Sp -= beta y y^T / (1 + beta psi.y) ---
        if (ctx->has_factor) {
            k_analysis<<<K + 1, 128, 0, st>>>(phi, ctx->psi, ctx->twiddle, K, 1.0, 1.0, ctx->inv_perm);
            SSPBO_LAUNCH_CHECK(ctx);
            k_gemv_rows<<<(d + 7) / 8, 256, 0, st>>>(ctx->Sp, ctx->psi, ctx->y, d);
            SSPBO_LAUNCH_CHECK(ctx);
            k_update_scalars<<<1, 256, 0, st>>>(ctx->psi, ctx->y, d, beta, ctx->scal + 2);
            SSPBO_LAUNCH_CHECK(ctx);
            k_rank1<<<g1, 256, 0, st>>>(ctx->Sp, ctx->y, ctx->y, d, ctx->scal + 2, 0.0);
            if (ctx->lr_valid) {
                if (ctx->lr_rank >= SSPBO_LR_MAX) ctx->lr_valid = false;        // beyond the low-rank regime
                else {
                    SSPBO_LAUNCH_CHECK(ctx);                                    // (the rank-one launch above)
                    const size_t off = (size_t)(ctx->lr_rank + 1) * ctx->KC_pad;      // row 0 holds m'
                    k_lr_append<<<(d + 255) / 256, 256, 0, st>>>(ctx->y, ctx->scal + 2, ctx->col_of_rank, d, ctx->r_scale,
                                                                 ctx->Vh + off, ctx->Vl + off, ctx->Vd + (size_t)ctx->lr_rank * d);
                    ctx->lr_rank++;
                }
            }
        }
        SSPBO_LAUNCH_CHECK(ctx);
        k_rank1<<<g1, 256, 0, st>>>(ctx->S, ctx->v, ctx->z, d, ctx->scal + 0, 0.0);              // S -= coef v z^T
        SSPBO_LAUNCH_CHECK(ctx);
        k_rank1<<<g1, 256, 0, st>>>(ctx->Sinv, phi, phi, d, nullptr, beta);                      // S_inv += beta phi phi^T
        SSPBO_LAUNCH_CHECK(ctx);
        k_axpy<<<(d + 255) / 256, 256, 0, st>>>(ctx->b, phi, beta * ts_host[i], d);              // b += beta t phi
        SSPBO_LAUNCH_CHECK(ctx);
        sspbo_eager_update_done(ctx);
    }
    if (k > 0) {
        if (!m_current) {
            k_gemv_rows<<<(d + 7) / 8, 256, 0, st>>>(ctx->S, ctx->b, ctx->m, d);                 // m = S b  (blr.py:75)
            SSPBO_LAUNCH_CHECK(ctx);
        }
        if (ctx->has_factor) {
            if (!m_current) {
                k_analysis<<<K + 1, 128, 0, st>>>(ctx->m, ctx->mp, ctx->twiddle, K, 1.0 / d, 2.0 / d, nullptr);   // m' = B^T m
                SSPBO_LAUNCH_CHECK(ctx);
            }
            if (!tail_done) {
                k_mp_op<<<1, 1024, 0, st>>>(ctx->mp, ctx->perm, ctx->col_of_rank, d, ctx->mp_op, ctx->Vh, ctx->Vl, ctx->mscale);
                SSPBO_LAUNCH_CHECK(ctx);
            }
            ctx->v8_row0_dirty = true;
        }
        ctx->operands_dirty = true; ctx->factor_dirty = true; ctx->f8_factor_dirty = true;
    }
    return SSPBO_OK;
}

int sspbo_refresh_factor(sspbo_ctx* ctx, cudaStream_t st) {
    if (!ctx->has_factor) SSPBO_FAIL(ctx, SSPBO_ESTATE, "scoring needs a BLR created over an SSP space (blr_init)");
    const int d = ctx->d;
    if (ctx->factor_dirty) {                                    // L = chol(Sp), blocked, in place on a copy
        const int rc_sp = sspbo_ensure_sp(ctx, st);
        if (rc_sp) return rc_sp;
        SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->R, ctx->Sp, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToDevice, st));
        for (int k0 = 0; k0 < d; k0 += CH_NB) {
            k_chol_diag<<<1, dim3(CH_NB, CH_NB), 0, st>>>(ctx->R, d, k0);
            SSPBO_LAUNCH_CHECK(ctx);
            int rest = d - k0 - CH_NB;
            if (rest > 0) {
                k_chol_trsm<<<(rest + 127) / 128, 128, 0, st>>>(ctx->R, d, k0);
                SSPBO_LAUNCH_CHECK(ctx);
                int tiles = (rest + 31) / 32;
                k_chol_syrk<<<dim3(tiles, tiles), dim3(32, 8), 0, st>>>(ctx->R, d, k0);
                SSPBO_LAUNCH_CHECK(ctx);
            }
        }
        ctx->factor_dirty = false;
    }
    return SSPBO_OK;
}

int sspbo_refresh_operands(sspbo_ctx* ctx, cudaStream_t st) {
    if (!ctx->has_factor) SSPBO_FAIL(ctx, SSPBO_ESTATE, "scoring needs a BLR created over an SSP space (blr_init)");
    if (!ctx->operands_dirty) return SSPBO_OK;
    const int d = ctx->d;
    int rc = sspbo_refresh_factor(ctx, st);
    if (rc) return rc;
    k_refresh_operands<<<dim3((ctx->dp + 255) / 256, d), 256, 0, st>>>(ctx->R, ctx->mp, ctx->perm, d, ctx->dp, ctx->Rt_f32,
                                                                    ctx->mp_f32);
    SSPBO_LAUNCH_CHECK(ctx);
    dim3 g2((d + 31) / 32, (d + 31) / 32);
    k_refresh_umma_operands<<<g2, dim3(32, 8), 0, st>>>(ctx->R, ctx->mp, ctx->perm, ctx->col_of_rank, d, ctx->KC_pad,
                                                        ctx->r_scale, ctx->Rh, ctx->Rl, ctx->mp_op);
    SSPBO_LAUNCH_CHECK(ctx);
    ctx->operands_dirty = false;
    return SSPBO_OK;
}

extern "C" int sspbo_blr_get(sspbo_ctx* ctx, double* m, double* S, double* Sinv, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_get before blr_init");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = ctx->d;
    {
        int rc = 0;
        if (m && !S) rc = sspbo_ensure_bm(ctx, st);
        if (S) rc = sspbo_ensure_s(ctx, st);
        if (!rc && Sinv) rc = sspbo_ensure_sinv(ctx, st);
        if (rc) return rc;
    }
    if (m) SSPBO_CUDA(ctx, cudaMemcpyAsync(m, ctx->m, d * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (S) SSPBO_CUDA(ctx, cudaMemcpyAsync(S, ctx->S, d * d * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (Sinv) SSPBO_CUDA(ctx, cudaMemcpyAsync(Sinv, ctx->Sinv, d * d * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return SSPBO_OK;
}

// packed posterior: S, S_inv, b, m [, Sp, m', rank (one float64, -1 = low-rank rows not valid), rows of V (SSPBO_LR_MAX x d)]
extern "C" int64_t sspbo_blr_export_size(const sspbo_ctx* ctx) {
    if (!ctx || !ctx->blr_ready) return 0;
    int64_t d = ctx->d;
    return ctx->has_factor ? 3 * d * d + 3 * d + 1 + (int64_t)SSPBO_LR_MAX * d : 2 * d * d + 2 * d;
}
extern "C" int sspbo_blr_export(sspbo_ctx* ctx, double* packed, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready || !packed) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_export: not initialised / null");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = ctx->d, dd = d * d;
    {
        const int rc = sspbo_ensure_all(ctx, st);
        if (rc) return rc;
    }
    const double* src[6] = {ctx->S, ctx->Sinv, ctx->b, ctx->m, ctx->Sp, ctx->mp};
    const size_t cnt[6] = {dd, dd, d, d, dd, d};
    size_t off = 0;
    for (int i = 0; i < (ctx->has_factor ? 6 : 4); ++i) {
        SSPBO_CUDA(ctx, cudaMemcpyAsync(packed + off, src[i], cnt[i] * sizeof(double), cudaMemcpyDeviceToDevice, st));
        off += cnt[i];
    }
    if (ctx->has_factor) {
        ctx->export_rank = ctx->lr_valid ? (double)ctx->lr_rank : -1.0;
        SSPBO_CUDA(ctx, cudaMemcpyAsync(packed + off, &ctx->export_rank, sizeof(double), cudaMemcpyHostToDevice, st));
        SSPBO_CUDA(ctx, cudaStreamSynchronize(st));              // export_rank is a host scalar of the ctx
        off += 1;
        SSPBO_CUDA(ctx, cudaMemcpyAsync(packed + off, ctx->Vd, (size_t)SSPBO_LR_MAX * d * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    return SSPBO_OK;
}
static int blr_import_impl(sspbo_ctx* ctx, const double* packed, double beta, bool rank_known, int lr_rank_known, void* stream);
extern "C" int sspbo_blr_import(sspbo_ctx* ctx, const double* packed, double beta, void* stream) {
    return blr_import_impl(ctx, packed, beta, false, -1, stream);
}
extern "C" int sspbo_blr_import_ranked(sspbo_ctx* ctx, const double* packed, double beta, int lr_rank, void* stream) {
    return blr_import_impl(ctx, packed, beta, true, lr_rank, stream);
}
extern "C" int sspbo_blr_rank(const sspbo_ctx* ctx, int* lr_rank) {
    if (!ctx || !lr_rank) return SSPBO_EINVAL;
    *lr_rank = (ctx->blr_ready && ctx->has_factor && ctx->lr_valid) ? ctx->lr_rank : -1;
    return SSPBO_OK;
}
static int blr_import_impl(sspbo_ctx* ctx, const double* packed, double beta, bool rank_known, int lr_rank_known, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready || !packed) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_import: not initialised / null");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = ctx->d, dd = d * d;
    double* dst[6] = {ctx->S, ctx->Sinv, ctx->b, ctx->m, ctx->Sp, ctx->mp};
    const size_t cnt[6] = {dd, dd, d, d, dd, d};
    size_t off = 0;
    for (int i = 0; i < (ctx->has_factor ? 6 : 4); ++i) {
        SSPBO_CUDA(ctx, cudaMemcpyAsync(dst[i], packed + off, cnt[i] * sizeof(double), cudaMemcpyDeviceToDevice, st));
        off += cnt[i];
    }
    if (beta > 0) { ctx->beta = beta; ctx->has_beta = true; }
    ctx->operands_dirty = true; ctx->factor_dirty = true; ctx->f8_factor_dirty = true;
    ctx->sinv_pending = 0; ctx->bm_stale = false; ctx->bpsi_stale = true;   // (S_inv, b, m complete; b_psi follows from b)
    ctx->ck_valid = false;
    if (ctx->has_factor) {
        double rank = (double)lr_rank_known;                     // the rows of V travel with the packed state
        if (!rank_known) {
            SSPBO_CUDA(ctx, cudaMemcpyAsync(&rank, packed + off, sizeof(double), cudaMemcpyDeviceToHost, st));
            SSPBO_CUDA(ctx, cudaStreamSynchronize(st));
        }
        off += 1;
        SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->Vd, packed + off, (size_t)SSPBO_LR_MAX * d * sizeof(double), cudaMemcpyDeviceToDevice, st));
        const size_t img = (size_t)(SSPBO_LR_MAX + 1) * ctx->KC_pad * sizeof(__half);
        SSPBO_CUDA(ctx, cudaMemsetAsync(ctx->Vh, 0, img, st));
        SSPBO_CUDA(ctx, cudaMemsetAsync(ctx->Vl, 0, img, st));
        if (ctx->V8) SSPBO_CUDA(ctx, cudaMemsetAsync(ctx->V8, 0, 2 * img, st));
        ctx->v8_rank = 0; ctx->v8_row0_dirty = true;
        ctx->lr_valid = rank >= 0.0 && rank <= (double)SSPBO_LR_MAX;
        ctx->lr_rank = ctx->lr_valid ? (int)rank : 0;
        ctx->sp_applied = ctx->lr_rank; ctx->s_applied = ctx->lr_rank;     // the imported matrices are complete
        if (ctx->lr_valid && ctx->lr_rank > 0) {
            k_lr_rebuild<<<dim3(((int)d + 255) / 256, ctx->lr_rank), 256, 0, st>>>(ctx->Vd, ctx->col_of_rank, (int)d, ctx->KC_pad,
                                                                                    ctx->r_scale, ctx->Vh, ctx->Vl);
            SSPBO_LAUNCH_CHECK(ctx);
        }
        k_mp_op<<<1, 1024, 0, st>>>(ctx->mp, ctx->perm, ctx->col_of_rank, (int)d, ctx->mp_op, ctx->Vh, ctx->Vl, ctx->mscale);
        SSPBO_LAUNCH_CHECK(ctx);
    }
    return SSPBO_OK;
}

// ---- checkpoint / rollback of the low-rank posterior -----------------------------------------------------------------------
// While updates run in low-rank form an observation only appends a row of V and rewrites b_psi, m' and the scorer's copies of
// m'; undoing the observations since a checkpoint is therefore one small kernel: those vectors back, the appended operand rows
// zeroed, the rank reset.  (Fantasy observations of batch BO -- "kriging believer" / "constant liar" -- and a posterior that a
// benchmark pins at one rank.)  Not possible once the d x d state has absorbed a later observation (something read S, S_inv or
// the psi-space covariance, or the ring of S_inv terms was flushed): rollback then reports SSPBO_EUNSUPPORTED and the caller
// restores an exported state instead.
__global__ void k_ck_copy(const double* __restrict__ vec_src, double* __restrict__ bpsi, double* __restrict__ mp, int d,
                          const float* __restrict__ mpop_src, float* __restrict__ mp_op, const __half* __restrict__ row0_src,
                          __half* __restrict__ vh, __half* __restrict__ vl, int KC, const float* __restrict__ ms_src,
                          float* __restrict__ mscale, int zero_r0, int zero_r1, uint8_t* __restrict__ v8) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int j = gt; j < d; j += nt) { bpsi[j] = vec_src[j]; mp[j] = vec_src[d + j]; }
    for (int c = gt; c < KC; c += nt) { mp_op[c] = mpop_src[c]; vh[c] = row0_src[c]; vl[c] = row0_src[KC + c]; }
    if (gt == 0) *mscale = *ms_src;
    // operand rows of the observations being undone (row r + 1 holds observation r)
    const size_t z0 = (size_t)(zero_r0 + 1) * KC, z1 = (size_t)(zero_r1 + 1) * KC;
    for (size_t i = z0 + gt; i < z1; i += nt) { vh[i] = __float2half(0.f); vl[i] = __float2half(0.f); }
    if (v8) for (size_t i = 2 * z0 + gt; i < 2 * z1; i += nt) v8[i] = 0;
}

extern "C" int sspbo_blr_checkpoint(sspbo_ctx* ctx, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready || !ctx->has_factor || !ctx->lr_valid || ctx->unfused_update)
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "blr_checkpoint: needs a posterior over an SSP space in low-rank form");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    double* ring_row = nullptr;
    if ((rc = sspbo_ensure_bm(ctx, st)) || (rc = sspbo_lr_update_prepare(ctx, st, &ring_row))) return rc;    // b_psi allocated and current
    const int d = ctx->d, KC = ctx->KC_pad;
    if (!ctx->ck_vec) {
        if ((rc = dev_alloc(ctx, &ctx->ck_vec, (size_t)2 * d)) || (rc = dev_alloc(ctx, &ctx->ck_mp_op, (size_t)KC)) ||
            (rc = dev_alloc(ctx, &ctx->ck_row0, (size_t)2 * KC)) || (rc = dev_alloc(ctx, &ctx->ck_mscale, (size_t)1)))
            return rc;
    }
    SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->ck_vec, ctx->bpsi, (size_t)d * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->ck_vec + d, ctx->mp, (size_t)d * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->ck_mp_op, ctx->mp_op, (size_t)KC * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->ck_row0, ctx->Vh, (size_t)KC * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->ck_row0 + KC, ctx->Vl, (size_t)KC * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->ck_mscale, ctx->mscale, sizeof(float), cudaMemcpyDeviceToDevice, st));
    ctx->ck_valid = true; ctx->ck_rank = ctx->lr_rank; ctx->ck_sinv_pending = ctx->sinv_pending;
    return SSPBO_OK;
}

extern "C" int sspbo_blr_rollback(sspbo_ctx* ctx, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->ck_valid) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_rollback without a checkpoint");
    const int undone = ctx->lr_rank - ctx->ck_rank;
    if (!ctx->lr_valid || undone < 0 || ctx->s_applied > ctx->ck_rank || ctx->sp_applied > ctx->ck_rank ||
        ctx->sinv_pending != ctx->ck_sinv_pending + undone || ctx->bpsi_stale)
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "blr_rollback: the dense state has absorbed observations made after the checkpoint");
    if (undone == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int zero_r1 = ctx->lr_rank;
    uint8_t* v8 = (ctx->V8 && ctx->v8_rank > ctx->ck_rank) ? ctx->V8 : nullptr;
    k_ck_copy<<<8, 256, 0, st>>>(ctx->ck_vec, ctx->bpsi, ctx->mp, ctx->d, ctx->ck_mp_op, ctx->mp_op, ctx->ck_row0, ctx->Vh, ctx->Vl,
                                  ctx->KC_pad, ctx->ck_mscale, ctx->mscale, ctx->ck_rank, zero_r1, v8);
    SSPBO_LAUNCH_CHECK(ctx);
    ctx->launches++;
    ctx->lr_rank = ctx->ck_rank; ctx->sinv_pending = ctx->ck_sinv_pending;
    if (ctx->v8_rank > ctx->ck_rank) ctx->v8_rank = ctx->ck_rank;
    ctx->bm_stale = true;                                       // (b and m may have been brought up to date in between)
    ctx->v8_row0_dirty = true; ctx->operands_dirty = true; ctx->factor_dirty = true; ctx->f8_factor_dirty = true;
    return SSPBO_OK;
}

extern "C" int sspbo_blr_predict(sspbo_ctx* ctx, const double* phi, int64_t N, double* mu, double* var, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_predict before blr_init");
    if (!ctx->has_beta) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_predict before beta is set (blr.py:96 divides by beta)");
    if (N < 0 || (N > 0 && (!phi || !mu || !var))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "blr_predict: null pointer");
    if (N == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t smem = (size_t)PRED_CPB * ctx->d * sizeof(double);
    SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks64 = (N + PRED_CPB - 1) / PRED_CPB;
    int blocks = (int)(blocks64 < (int64_t)ctx->sm_count * 4 ? blocks64 : (int64_t)ctx->sm_count * 4);
    { const int rc = sspbo_ensure_s(ctx, (cudaStream_t)stream); if (rc) return rc; }
    k_predict<<<blocks, 256, smem, (cudaStream_t)stream>>>(phi, N, ctx->S, ctx->m, ctx->d, 1.0 / ctx->beta, mu, var);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

// ------------------------------------------------------------------------------------ K2 dispatch
// float64 pass over the candidates a tensor pass flagged (device-side list and count): the low-rank form over the rows of V
// while they are valid, else the quadratic form with the psi-space covariance
static void fill_exact_args(const sspbo_ctx* ctx, const sspbo_lr::CandDesc& cand, int64_t index0, double lam, double gamma_t,
                            float* mu, float* var, float* acq, unsigned long long* best, bool full, ExactArgs* ea) {
    ea->x = cand.x; ea->cand = cand; ea->list = ctx->flag_idx; ea->count_ptr = ctx->stats + SSPBO_STAT_CUR; ea->cap = SSPBO_FLAG_CAP;
    ea->A_turns = ctx->A_turns; ea->tau_turns = ctx->tau_turns; ea->n = ctx->n; ea->K = ctx->K; ea->L = ctx->L; ea->ncls = ctx->n_cls;
    ea->dc_value = ctx->dc; ea->normalize = ctx->normalize ? 1 : 0; ea->inv_perm = ctx->inv_perm; ea->perm = ctx->perm;
    ea->Vd = full ? ctx->Sp : ctx->Vd; ea->r = full ? ctx->d : ctx->lr_rank; ea->mp = ctx->mp;
    ea->inv_beta = 1.0 / ctx->beta; ea->prior_var = 1.0 / ctx->alpha; ea->lam = lam; ea->gamma_t = gamma_t; ea->index0 = index0;
    ea->mu_out = mu; ea->var_out = var; ea->acq_out = acq; ea->best = best;
}

static int score_flagged_launch(sspbo_ctx* ctx, const sspbo_lr::CandDesc& cand, int64_t N, int64_t index0, double lam, double gamma_t,
                                float* mu, float* var, float* acq, unsigned long long* best, cudaStream_t st) {
    const int d = ctx->d;
    const size_t smem = (size_t)EX_G * d * sizeof(double);
    const int gx = 2 * ctx->sm_count;
    const bool full = !ctx->lr_valid;
    if (full) { const int rc_sp = sspbo_ensure_sp(ctx, st); if (rc_sp) return rc_sp; }
    ExactArgs ea;
    fill_exact_args(ctx, cand, index0, lam, gamma_t, mu, var, acq, best, full, &ea);
#define SSPBO_FLAGGED_LAUNCH(XT, FULL)                                                                                             \
    do {                                                                                                                           \
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_exact_lowrank<XT, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        k_exact_lowrank<XT, FULL><<<gx, EXL_THREADS, smem, st>>>(ea);                                                              \
    } while (0)
    if (cand.x_f64) { if (full) SSPBO_FLAGGED_LAUNCH(double, true); else SSPBO_FLAGGED_LAUNCH(double, false); }
    else { if (full) SSPBO_FLAGGED_LAUNCH(float, true); else SSPBO_FLAGGED_LAUNCH(float, false); }
#undef SSPBO_FLAGGED_LAUNCH
    SSPBO_LAUNCH_CHECK(ctx);
    k_flag_fold<<<1, 1, 0, st>>>(ctx->stats, SSPBO_FLAG_CAP, (unsigned long long)N);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

int sspbo_score_exact_launch(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, bool from_flag_list, int64_t index0,
                             double lam, double gamma_t, float* mu, float* var, float* acq, unsigned long long* best,
                             cudaStream_t st) {
    if (!ctx->has_factor) SSPBO_FAIL(ctx, SSPBO_ESTATE, "scoring needs a BLR created over an SSP space (blr_init)");
    const int d = ctx->d;
    if (from_flag_list) {
        sspbo_lr::CandDesc cand;
        memset(&cand, 0, sizeof(cand));
        cand.x = x; cand.x_f64 = (x_dtype == SSPBO_F64);
        return score_flagged_launch(ctx, cand, N, index0, lam, gamma_t, mu, var, acq, best, st);
    }
    // row blocks: a group of EX_G candidates is reduced against Sp by RB CTAs, so that a handful of candidates (eval(x_t))
    // spreads over the machine instead of streaming all of Sp through one SM
    { const int rc_sp = sspbo_ensure_sp(ctx, st); if (rc_sp) return rc_sp; }
    const int64_t groups = (N + EX_G - 1) / EX_G;
    int RB = 1;
    while (RB < EX_RB_MAX && groups * RB < 2 * ctx->sm_count) RB *= 2;
    const size_t need = (size_t)N * (RB + 2);                  // partial quadratic forms, mu, |phi|^2
    if (need > ctx->ex_part_cap) {
        if (ctx->ex_part) { SSPBO_CUDA(ctx, cudaStreamSynchronize(st)); }
        int rc = dev_alloc(ctx, &ctx->ex_part, need);
        if (rc) return rc;
        ctx->ex_part_cap = need;
    }
    double* part_q = ctx->ex_part;
    double* part_mu = ctx->ex_part + (size_t)N * RB;
    const size_t smem = (size_t)EX_G * d * sizeof(double);
    const int gx = (int)(groups < 4 * ctx->sm_count ? groups : 4 * ctx->sm_count);
    if (x_dtype == SSPBO_F32) {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_exact_partial<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_exact_partial<float><<<dim3(gx, RB), EX_THREADS, smem, st>>>((const float*)x, N, ctx->A_turns,
            ctx->tau_turns, ctx->n, ctx->K, ctx->L, ctx->dc, ctx->inv_perm, ctx->perm, ctx->Sp, ctx->mp, RB, part_q, part_mu, ctx->n_cls);
    } else {
        SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_exact_partial<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_exact_partial<double><<<dim3(gx, RB), EX_THREADS, smem, st>>>((const double*)x, N, ctx->A_turns,
            ctx->tau_turns, ctx->n, ctx->K, ctx->L, ctx->dc, ctx->inv_perm, ctx->perm, ctx->Sp, ctx->mp, RB, part_q, part_mu, ctx->n_cls);
    }
    SSPBO_LAUNCH_CHECK(ctx);
    const int fb = (int)((N + 255) / 256 < 1024 ? (N + 255) / 256 : 1024);
    k_exact_finish<<<fb, 256, 0, st>>>(N, RB, part_q, part_mu, ctx->normalize ? 1 : 0, 1.0 / ctx->beta, lam, gamma_t, index0, mu, var, acq, best,
                                         ctx->exact_out64);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

// rows [index0, index0 + N) of a generated candidate set written out for the engines that read explicit rows
static int materialise_rows(sspbo_ctx* ctx, sspbo_lr::CandDesc* cand, int64_t N, int64_t index0, cudaStream_t st) {
    const int n = ctx->n;
    const size_t bytes = (size_t)N * n * (cand->mode == 1 ? sizeof(float) : sizeof(double));
    if (bytes > ctx->gen_rows_bytes) {
        if (ctx->gen_rows) { SSPBO_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(ctx->gen_rows); ctx->gen_rows = nullptr; ctx->gen_rows_bytes = 0; }
        SSPBO_CUDA(ctx, cudaMalloc(&ctx->gen_rows, bytes));
        ctx->gen_rows_bytes = bytes;
    }
    const int64_t total = N * n;
    const int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sm_count * 16 ? (total + 255) / 256 : (int64_t)ctx->sm_count * 16);
    if (cand->mode == 1) {
        k_fill_uniform<<<blocks, 256, 0, st>>>(cand->seed, index0 * n, total, (float*)ctx->gen_rows);
        cand->x_f64 = 0;
    } else {
        GridDesc g; g.n = n;
        for (int a = 0; a < n; ++a) { g.counts[a] = (int)cand->counts[a]; g.offs[a] = cand->offs[a]; }
        k_grid_points<<<blocks, 256, 0, st>>>(cand->axes, g, index0, N, (double*)ctx->gen_rows);
        cand->x_f64 = 1;
    }
    SSPBO_LAUNCH_CHECK(ctx);
    cand->x = ctx->gen_rows; cand->mode = 0;
    return SSPBO_OK;
}

static int score_dispatch(sspbo_ctx* ctx, sspbo_lr::CandDesc cand, int64_t N, int64_t index0, double lam, double gamma_t,
                          float* mu, float* var, float* acq, unsigned long long* best, int engine, cudaStream_t st) {
    if (!ctx->blr_ready || !ctx->has_beta) SSPBO_FAIL(ctx, SSPBO_ESTATE, "score: BLR posterior (and beta) must be set first");
    if (index0 < 0 || index0 + N > 0xffffffffLL) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score: candidate index must stay below 2^32");
    if (!(gamma_t >= 0)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score: gamma_t must be >= 0");
    if (engine < SSPBO_ENGINE_AUTO || engine > SSPBO_ENGINE_UMMA_F8) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score: unknown engine %d", engine);
    if (N == 0) return SSPBO_OK;
    if (!ctx->has_factor) SSPBO_FAIL(ctx, SSPBO_ESTATE, "scoring needs a BLR created over an SSP space (blr_init)");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool umma_ok = sspbo_score_umma_supported(ctx) != 0;
    // low-rank form: split-fp16 operands (one chunk, rank <= 255) and / or fp16 + e4m3 operands (rank <= 511)
    const bool lr16_ok = sspbo_score_lr_supported(ctx) && ctx->operand_format < 1;
    const bool lr8_ok = sspbo_score_f8_supported(ctx, 0) && ctx->operand_format != 0;
    const bool lr_ok = lr16_ok || lr8_ok;
    const bool f8_ok = sspbo_score_f8_supported(ctx, 1) != 0;
    if (engine == SSPBO_ENGINE_UMMA && !umma_ok)
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "score: tcgen05 engine does not cover d=%d L=%d", ctx->d, ctx->L);
    if (engine == SSPBO_ENGINE_UMMA_LR && !lr_ok)
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "score: low-rank tcgen05 engine needs 1 <= rank <= %d rows appended since blr_init "
                   "(rank %d, valid %d, operand format %d)", SSPBO_LR_MAX, ctx->lr_rank, (int)ctx->lr_valid, ctx->operand_format);
    if (engine == SSPBO_ENGINE_UMMA_F8 && !f8_ok)
        SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "score: factor-form tcgen05 engine does not cover d=%d n=%d L=%d", ctx->d, ctx->n, ctx->L);
    if (engine == SSPBO_ENGINE_AUTO) {
        const bool lr_wanted = lr_ok && !ctx->lr_disabled;
        if (N <= 64) engine = SSPBO_ENGINE_EXACT;               // e.g. eval(x_t) of every BO step: no factorisation needed
        else if (N >= 1024 && f8_ok && !ctx->f8_disabled)       // whichever form is estimated to take less time (sspbo_f8_cost)
            engine = (lr_wanted && sspbo_f8_cost(ctx, 0, N) <= sspbo_f8_cost(ctx, 1, N)) ? SSPBO_ENGINE_UMMA_LR : SSPBO_ENGINE_UMMA_F8;
        else if (N >= 1024 && lr_wanted && 2 * (ctx->lr_rank + 1) <= (ctx->d >= 128 ? 3 : 1) * ctx->d)
            engine = SSPBO_ENGINE_UMMA_LR;
        else if (N >= 1024 && umma_ok) engine = SSPBO_ENGINE_UMMA;
        else engine = SSPBO_ENGINE_SIMT;
    }
    ctx->last_engine = engine;
    if (cand.mode != 0 && engine != SSPBO_ENGINE_UMMA_LR && engine != SSPBO_ENGINE_UMMA_F8) {
        int rc = materialise_rows(ctx, &cand, N, index0, st);
        if (rc) return rc;
    }
    const void* x = cand.x;
    const int x_dtype = cand.x_f64 ? SSPBO_F64 : SSPBO_F32;
    if (engine == SSPBO_ENGINE_EXACT)
        return sspbo_score_exact_launch(ctx, x, x_dtype, N, false, index0, lam, gamma_t, mu, var, acq, best, st);
    if (engine == SSPBO_ENGINE_UMMA_LR) {
        // operand format: split fp16 while the kernel is bound by generating psi, fp16 + e4m3 once the tensor pipe binds
        const int bn = (ctx->lr_rank + 1 + 15) / 16 * 16;
        const bool use16 = lr16_ok && (!lr8_ok || ctx->operand_format == 0 || bn <= SSPBO_F8_MIN_BN - 16);
        timing_mark(ctx, 0, st);
        int rc = use16 ? sspbo_score_lr_launch(ctx, cand, N, index0, (float)lam, (float)gamma_t, mu, var, acq, best, st)
                       : sspbo_score_f8_launch(ctx, cand, N, index0, (float)lam, (float)gamma_t, mu, var, acq, best, 0, st);
        if (rc) return rc;
        timing_mark(ctx, 1, st);
        // exact pass over the candidates the low-rank pass flagged (device-side count; usually none)
        if (!(sspbo_debug_env() & 4))                            // (switch for timing experiments, compiled out of product builds)
            rc = score_flagged_launch(ctx, cand, N, index0, lam, gamma_t, mu, var, acq, best, st);
        timing_mark(ctx, 2, st);
        return rc;
    }
    if (engine == SSPBO_ENGINE_UMMA_F8) {
        timing_mark(ctx, 0, st);
        int rc = sspbo_score_f8_launch(ctx, cand, N, index0, (float)lam, (float)gamma_t, mu, var, acq, best, 1, st);
        if (rc) return rc;
        timing_mark(ctx, 1, st);
        // float64 pass over the candidates whose variance is a small fraction of the prior's (SSPBO_F8_FLAG_FRACTION)
        rc = score_flagged_launch(ctx, cand, N, index0, lam, gamma_t, mu, var, acq, best, st);
        timing_mark(ctx, 2, st);
        return rc;
    }
    int rc = sspbo_refresh_operands(ctx, st);
    if (rc) return rc;
    timing_mark(ctx, 0, st);
    if (engine == SSPBO_ENGINE_UMMA)
        rc = sspbo_score_umma_launch(ctx, x, x_dtype, N, index0, (float)lam, (float)gamma_t, mu, var, acq, best, st);
    else
        rc = sspbo_score_simt_launch(ctx, x, x_dtype, N, index0, (float)lam, (float)gamma_t, mu, var, acq, best, st);
    timing_mark(ctx, 1, st);
    timing_mark(ctx, 2, st);
    return rc;
}

extern "C" int sspbo_score(sspbo_ctx* ctx, const void* x, int x_dtype, int64_t N, int64_t index0, double lam, double gamma_t,
                           float* mu, float* var, float* acq, unsigned long long* best, int engine, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (N < 0 || (N > 0 && !x)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score: null candidates");
    if (x_dtype != SSPBO_F32 && x_dtype != SSPBO_F64) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score: unknown dtype %d", x_dtype);
    sspbo_lr::CandDesc cand;
    memset(&cand, 0, sizeof(cand));
    cand.x = x; cand.x_f64 = (x_dtype == SSPBO_F64);
    return score_dispatch(ctx, cand, N, index0, lam, gamma_t, mu, var, acq, best, engine, (cudaStream_t)stream);
}

/* Fused scoring of a candidate set that is GENERATED inside the kernel (no HBM or PCIe byte per candidate): rows
 * [index0, index0 + N) of the counter-based uniform stream of sspbo_fill_uniform (kind SSPBO_CAND_UNIFORM, `seed`), or of
 * the grid of sspbo_grid_points (kind SSPBO_CAND_GRID: axes_dev / counts_host as there; SSPSpace.get_sample_points('grid'),
 * sspspace.py:488-495).  Keys carry the global row index.  Engines that read explicit rows get them written to a scratch
 * buffer first. */
extern "C" int sspbo_score_generated(sspbo_ctx* ctx, int kind, uint64_t seed, const double* axes_dev, const int* counts_host,
                                     int64_t N, int64_t index0, double lam, double gamma_t, float* mu, float* var, float* acq,
                                     unsigned long long* best, int engine, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "score_generated before space_set");
    if (ctx->L != 1) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "score_generated: trajectory contexts take explicit candidate rows");
    if (N < 0 || index0 < 0) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score_generated: negative range");
    sspbo_lr::CandDesc cand;
    memset(&cand, 0, sizeof(cand));
    if (kind == SSPBO_CAND_UNIFORM) {
        cand.mode = 1; cand.seed = seed;
    } else if (kind == SSPBO_CAND_GRID) {
        if (!axes_dev || !counts_host) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score_generated: grid needs axes and counts");
        cand.mode = 2; cand.axes = axes_dev;
        int off = 0; double total = 1.0;
        for (int a = 0; a < ctx->n; ++a) {
            if (counts_host[a] < 1) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score_generated: counts must be >= 1");
            cand.counts[a] = (unsigned int)counts_host[a]; cand.offs[a] = off; off += counts_host[a]; total *= counts_host[a];
        }
        if ((double)(index0 + N) > total) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score_generated: range exceeds the grid");
    } else SSPBO_FAIL(ctx, SSPBO_EINVAL, "score_generated: unknown kind %d", kind);
    return score_dispatch(ctx, cand, N, index0, lam, gamma_t, mu, var, acq, best, engine, (cudaStream_t)stream);
}

extern "C" int sspbo_set_operand_format(sspbo_ctx* ctx, int format) {
    if (!ctx) return SSPBO_EINVAL;
    if (format < -1 || format > 2)
        SSPBO_FAIL(ctx, SSPBO_EINVAL, "operand format: -1 (by rank), 0 (split fp16), 1 (fp16 + e4m3, single CTAs) or 2 (fp16 + e4m3, CTA pairs)");
    ctx->operand_format = format;
    return SSPBO_OK;
}

/* counters of the low-rank pass since the last reset; synchronises `stream`.  Applies the engine policy: when more
 * than 1 % of the scored candidates needed the exact pass (or the flag list overflowed) the AUTO engine switches to the
 * full factor; candidates outside the declared |x| bound switch the generator to the 64-bit phase path.  *rerun is set
 * when the results accumulated since the last reset must be recomputed (overflow or out-of-range candidates). */
// pinned host staging for the small synchronous read-backs / uploads of a BO step (key + counters, eval of x_t)
static int host_stage(sspbo_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pin_bytes) return SSPBO_OK;
    if (ctx->pin) { cudaFreeHost(ctx->pin); ctx->pin = nullptr; ctx->pin_bytes = 0; }
    SSPBO_CUDA(ctx, cudaMallocHost(&ctx->pin, bytes));
    ctx->pin_bytes = bytes;
    return SSPBO_OK;
}

extern "C" int sspbo_score_stats(sspbo_ctx* ctx, int64_t* scored, int64_t* flagged, int64_t* overflow, int64_t* out_of_range,
                                 int* rerun, int reset, void* stream) {
    return sspbo_score_finish(ctx, nullptr, nullptr, scored, flagged, overflow, out_of_range, rerun, reset, stream);
}

extern "C" int sspbo_score_finish(sspbo_ctx* ctx, const unsigned long long* key_dev, unsigned long long* key_host,
                                  int64_t* scored, int64_t* flagged, int64_t* overflow, int64_t* out_of_range, int* rerun,
                                  int reset, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (scored) *scored = 0; if (flagged) *flagged = 0; if (overflow) *overflow = 0; if (out_of_range) *out_of_range = 0;
    if (rerun) *rerun = 0;
    if ((key_dev == nullptr) != (key_host == nullptr)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "score_finish: key_dev and key_host go together");
    if (!ctx->stats && !key_dev) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = host_stage(ctx, 4096);
    if (rc) return rc;
    unsigned long long* h = (unsigned long long*)ctx->pin;          // [0, STAT_COUNT): counters, [STAT_COUNT]: key
    for (int i = 0; i <= SSPBO_STAT_COUNT; ++i) h[i] = 0ull;
    if (ctx->stats)
        SSPBO_CUDA(ctx, cudaMemcpyAsync(h, ctx->stats, SSPBO_STAT_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (key_dev)
        SSPBO_CUDA(ctx, cudaMemcpyAsync(h + SSPBO_STAT_COUNT, key_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SSPBO_CUDA(ctx, cudaStreamSynchronize(st));                     // the one synchronisation of a selection step
    if (key_host) *key_host = h[SSPBO_STAT_COUNT];
    if (!ctx->stats) return SSPBO_OK;
    if (reset) SSPBO_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, SSPBO_STAT_COUNT * sizeof(unsigned long long), st));
    if (scored) *scored = (int64_t)h[SSPBO_STAT_SCORED];
    if (flagged) *flagged = (int64_t)h[SSPBO_STAT_FLAGGED];
    if (overflow) *overflow = (int64_t)h[SSPBO_STAT_OVERFLOW];
    if (out_of_range) *out_of_range = (int64_t)h[SSPBO_STAT_OOR];
    // more flagged candidates than the list holds: the low-rank form gives way to the factor form, and that one (whose own
    // flags then overflowed as well) to the engines that never flag
    if (h[SSPBO_STAT_OVERFLOW] > 0 && (ctx->lr_disabled || !ctx->lr_valid)) ctx->f8_disabled = true;
    if (h[SSPBO_STAT_OVERFLOW] > 0 || (h[SSPBO_STAT_SCORED] >= 4096 && h[SSPBO_STAT_RESCORED] * 100 > h[SSPBO_STAT_SCORED]))
        ctx->lr_disabled = true;
    if (h[SSPBO_STAT_OOR] > 0) ctx->p32_disabled = true;
    if (ctx->last_engine == SSPBO_ENGINE_UMMA_LR && h[SSPBO_STAT_SCORED] >= 4096)       // what the float64 pass really costs
        ctx->lr_flag_frac = (double)h[SSPBO_STAT_RESCORED] / (double)h[SSPBO_STAT_SCORED];
    if (rerun) *rerun = (h[SSPBO_STAT_OVERFLOW] > 0 || h[SSPBO_STAT_OOR] > 0 || (h[SSPBO_STAT_PEER] & 1ull)) ? 1 : 0;
    if (h[SSPBO_STAT_PEER] & 2ull) SSPBO_FAIL(ctx, SSPBO_ECUDA, "key exchange: a peer rank did not arrive within the time limit");
    return SSPBO_OK;
}

/* SSPAgent.eval (agents/ssp_agent.py:125-137) for a handful of HOST points, e.g. x_t of a BO step: one pinned upload, the
 * exact float64 engine, one pinned read-back, one synchronisation. */
extern "C" int sspbo_eval_host(sspbo_ctx* ctx, const double* x_host, int N, double lam, double gamma_t, double* mu_host,
                               double* var_host, double* acq_host, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!ctx->blr_ready || !ctx->has_beta) SSPBO_FAIL(ctx, SSPBO_ESTATE, "eval_host: BLR posterior (and beta) must be set first");
    if (N < 0 || N > 64) SSPBO_FAIL(ctx, SSPBO_EINVAL, "eval_host: 0 <= N <= 64 (larger sets go through sspbo_score)");
    if (N > 0 && (!x_host || !mu_host || !var_host || !acq_host)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "eval_host: null pointer");
    if (!(gamma_t >= 0)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "eval_host: gamma_t must be >= 0");
    if (N == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nl = (size_t)ctx->n * ctx->L;
    const size_t xb = 64 * nl * sizeof(double), ob = 3 * 64 * sizeof(double);
    int rc = host_stage(ctx, xb + ob > 4096 ? xb + ob : 4096);
    if (rc) return rc;
    if (xb + ob > ctx->eval_dev_bytes) {                 // sized by the row width (n * L doubles), 64 rows
        if (ctx->eval_dev) { SSPBO_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(ctx->eval_dev); ctx->eval_dev = nullptr; ctx->eval_dev_bytes = 0; }
        SSPBO_CUDA(ctx, cudaMalloc(&ctx->eval_dev, xb + ob));
        ctx->eval_dev_bytes = xb + ob;
    }
    double* xd = (double*)ctx->eval_dev;
    double* od = (double*)((char*)ctx->eval_dev + xb);
    double* xp = (double*)ctx->pin;
    double* op = (double*)((char*)ctx->pin + xb);
    memcpy(xp, x_host, (size_t)N * nl * sizeof(double));
    SSPBO_CUDA(ctx, cudaMemcpyAsync(xd, xp, (size_t)N * nl * sizeof(double), cudaMemcpyHostToDevice, st));
    ctx->last_engine = SSPBO_ENGINE_EXACT;
    ctx->exact_out64 = od;                               // float64 mu / var / acq straight from the float64 engine
    rc = sspbo_score_exact_launch(ctx, xd, SSPBO_F64, N, false, 0, lam, gamma_t, nullptr, nullptr, nullptr, nullptr, st);
    ctx->exact_out64 = nullptr;
    if (rc) return rc;
    SSPBO_CUDA(ctx, cudaMemcpyAsync(op, od, ob, cudaMemcpyDeviceToHost, st));
    SSPBO_CUDA(ctx, cudaStreamSynchronize(st));
    for (int i = 0; i < N; ++i) { mu_host[i] = op[i]; var_host[i] = op[64 + i]; acq_host[i] = op[128 + i]; }
    return SSPBO_OK;
}

/* SSPAgent.update (agents/ssp_agent.py:187-206) for HOST points: encode x_t on the device and apply the posterior update,
 * in one call and without a synchronisation (x is staged through pinned memory guarded by an event). */
extern "C" int sspbo_blr_update_x(sspbo_ctx* ctx, const double* x_host, const double* ts_host, int k, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_update_x needs an SSP space");
    if (!ctx->blr_ready || !ctx->has_beta) SSPBO_FAIL(ctx, SSPBO_ESTATE, "blr_update_x before blr_init / blr_set_beta");
    if (k < 0 || k > 64 || (k > 0 && (!x_host || !ts_host))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "blr_update_x: 0 <= k <= 64 points");
    if (k == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (k == 1) {                                        // low-rank form: the cluster encodes the point itself, any d
        sspbo_ctx* one[1] = {ctx};
        const int rc = sspbo_blr_update_lr(one, 1, x_host, nullptr, ts_host, st, nullptr, ctx->want_obs ? ctx->obs_out : nullptr);
        if (rc != SSPBO_EUNSUPPORTED) return rc;
    }
    if (k == 1 && ctx->d <= SSPBO_CLUSTER_UPDATE_MAX_D) {
        // small spaces: the observation of a BO step in ONE launch -- a cluster of four CTAs encodes the point, applies the
        // update and refreshes the scoring operands (the cooperative kernel's grid barriers and the two helper launches around
        // it cost more than the arithmetic at d = 151); larger spaces want the whole machine's bandwidth for the d x d state
        sspbo_ctx* one[1] = {ctx};
        const int rc = sspbo_blr_update_runs(one, 1, x_host, ts_host, st, nullptr, ctx->want_obs ? ctx->obs_out : nullptr);
        if (rc != SSPBO_EUNSUPPORTED) return rc;
    }
    const size_t nl = (size_t)ctx->n * ctx->L;
    const size_t stage = 64 * nl * sizeof(double);       // sized by the row width: thousands of contexts may coexist (C4)
    if (ctx->obs_pin) SSPBO_CUDA(ctx, cudaEventSynchronize(ctx->obs_event));       // the previous upload has left the pinned buffer
    if (stage > ctx->obs_bytes) {
        if (ctx->obs_pin) { SSPBO_CUDA(ctx, cudaStreamSynchronize(st)); cudaFreeHost(ctx->obs_pin); cudaFree(ctx->obs_x); ctx->obs_pin = nullptr; ctx->obs_x = nullptr; }
        else SSPBO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->obs_event, cudaEventDisableTiming));
        ctx->obs_bytes = 0;
        SSPBO_CUDA(ctx, cudaMallocHost(&ctx->obs_pin, stage));
        SSPBO_CUDA(ctx, cudaMalloc(&ctx->obs_x, stage));
        ctx->obs_bytes = stage;
    }
    if ((size_t)k * ctx->d > ctx->obs_phi_cap) {
        if (ctx->obs_phi) { SSPBO_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(ctx->obs_phi); ctx->obs_phi = nullptr; ctx->obs_phi_cap = 0; }
        SSPBO_CUDA(ctx, cudaMalloc((void**)&ctx->obs_phi, (size_t)k * ctx->d * sizeof(double)));
        ctx->obs_phi_cap = (size_t)k * ctx->d;
    }
    memcpy(ctx->obs_pin, x_host, (size_t)k * nl * sizeof(double));
    SSPBO_CUDA(ctx, cudaMemcpyAsync(ctx->obs_x, ctx->obs_pin, (size_t)k * nl * sizeof(double), cudaMemcpyHostToDevice, st));
    SSPBO_CUDA(ctx, cudaEventRecord(ctx->obs_event, st));
    int rc = sspbo_encode(ctx, ctx->obs_x, SSPBO_F64, k, ctx->obs_phi, stream);
    if (rc) return rc;
    return sspbo_blr_update(ctx, ctx->obs_phi, ts_host, k, stream);
}

/* One BO observation: SSPAgent.eval(x_t) (agents/ssp_agent.py:125-137) under the current posterior followed by
 * SSPAgent.update(x_t, y_t) (:187-206), for one HOST point.  The predictive mean and variance fall out of the update
 * kernel (mu = m . phi, var = 1/beta + phi^T S phi with v = S phi already formed), so the pair costs one upload, the
 * encode, one cooperative launch and ONE synchronising read-back of two float64 values. */
extern "C" int sspbo_observe(sspbo_ctx* ctx, const double* x_host, double t, double* mu_host, double* var_host, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (!x_host || !mu_host || !var_host) SSPBO_FAIL(ctx, SSPBO_EINVAL, "observe: null pointer");
    if (ctx->K == 0 || !ctx->blr_ready || !ctx->has_beta) SSPBO_FAIL(ctx, SSPBO_ESTATE, "observe needs a space and a posterior with beta");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (ctx->unfused_update || !sspbo_fused_update_available(ctx)) {   // no single-launch update on this device: the two calls it replaces
        double acq;
        if ((rc = sspbo_eval_host(ctx, x_host, 1, 0.0, 0.0, mu_host, var_host, &acq, stream))) return rc;
        return sspbo_blr_update_x(ctx, x_host, &t, 1, stream);
    }
    if (!ctx->obs_out) SSPBO_CUDA(ctx, cudaMalloc((void**)&ctx->obs_out, 2 * sizeof(double)));
    if ((rc = host_stage(ctx, 4096))) return rc;
    ctx->want_obs = true;
    rc = sspbo_blr_update_x(ctx, x_host, &t, 1, stream);
    ctx->want_obs = false;
    if (rc) return rc;
    if (ctx->unfused_update) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "observe: fused update unavailable; state was updated, call again for eval");
    double* h = (double*)ctx->pin + 64;                   // clear of the words sspbo_score_finish uses
    SSPBO_CUDA(ctx, cudaMemcpyAsync(h, ctx->obs_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SSPBO_CUDA(ctx, cudaStreamSynchronize(st));
    *mu_host = h[0]; *var_host = h[1];
    return SSPBO_OK;
}

/* Lock-step of many independent runs (BASELINE config C4): the per-run calls of a selection / an observation issued from
 * one C loop over the runs' contexts, round-robin over the caller's streams.  Returns the first non-zero return code;
 * *failed (nullable) receives the index of that run (its message is in sspbo_last_error of its context). */
// parameter blocks of the grouped launches: RunScore[R] | ExactArgs[R] | status[R], pinned staging + device copy owned by the
// first context of the batch
static int score_batch_grouped(sspbo_ctx* const* ctxs, int R, const void* x_dev, int x_dtype, int64_t N, const double* lam,
                               const double* gamma_t, unsigned long long* keys_dev, cudaStream_t st, int* failed) {
    sspbo_ctx* c0 = ctxs[0];
    if (!c0) return SSPBO_EINVAL;
    const size_t off_ex = (size_t)R * sizeof(sspbo_lr::RunScore);
    const size_t off_st = off_ex + (size_t)R * sizeof(ExactArgs);
    const size_t bytes = off_st + (size_t)R * sizeof(unsigned long long);
    SSPBO_CUDA(c0, cudaSetDevice(c0->device));
    if (c0->sbatch_pin) SSPBO_CUDA(c0, cudaEventSynchronize(c0->sbatch_event));
    if (bytes > c0->sbatch_bytes) {
        if (c0->sbatch_pin) { SSPBO_CUDA(c0, cudaStreamSynchronize(st)); cudaFreeHost(c0->sbatch_pin); cudaFree(c0->sbatch_dev); c0->sbatch_pin = nullptr; c0->sbatch_dev = nullptr; }
        else SSPBO_CUDA(c0, cudaEventCreateWithFlags(&c0->sbatch_event, cudaEventDisableTiming));
        c0->sbatch_bytes = 0;
        SSPBO_CUDA(c0, cudaMallocHost(&c0->sbatch_pin, bytes));
        SSPBO_CUDA(c0, cudaMalloc(&c0->sbatch_dev, bytes));
        c0->sbatch_bytes = bytes;
    }
    sspbo_lr::RunScore* hs = (sspbo_lr::RunScore*)c0->sbatch_pin;
    ExactArgs* he = (ExactArgs*)((char*)c0->sbatch_pin + off_ex);
    int rc = sspbo_score_lr_runs_prepare(ctxs, R, N, lam, gamma_t, keys_dev, hs);
    if (rc) return rc;                                                  // EUNSUPPORTED: nothing enqueued
    sspbo_lr::CandDesc cand;
    memset(&cand, 0, sizeof(cand));
    cand.x = x_dev; cand.x_f64 = (x_dtype == SSPBO_F64);
    for (int r = 0; r < R; ++r)
        fill_exact_args(ctxs[r], cand, 0, lam[r], gamma_t[r], nullptr, nullptr, nullptr, keys_dev + r, false, he + r);
    SSPBO_CUDA(c0, cudaMemcpyAsync(c0->sbatch_dev, c0->sbatch_pin, off_st, cudaMemcpyHostToDevice, st));
    SSPBO_CUDA(c0, cudaEventRecord(c0->sbatch_event, st));
    rc = sspbo_score_lr_runs_launch(ctxs, R, cand, N, (const sspbo_lr::RunScore*)c0->sbatch_dev, st);
    if (rc) { if (failed) *failed = 0; return rc; }
    int d_max = 0;
    for (int r = 0; r < R; ++r) d_max = ctxs[r]->d > d_max ? ctxs[r]->d : d_max;
    const size_t smem = (size_t)EX_G * d_max * sizeof(double);
    int parts = (4 * c0->sm_count) / R;
    parts = parts < 1 ? 1 : (parts > 16 ? 16 : parts);
    const int items = R * parts;
    const int gx = items < 4 * c0->sm_count ? items : 4 * c0->sm_count;
    const ExactArgs* de = (const ExactArgs*)((const char*)c0->sbatch_dev + off_ex);
    unsigned long long* dst = (unsigned long long*)((char*)c0->sbatch_dev + off_st);
    if (cand.x_f64) {
        SSPBO_CUDA(c0, cudaFuncSetAttribute(k_exact_lowrank_runs<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_exact_lowrank_runs<double><<<gx, EXL_THREADS, smem, st>>>(de, R, parts);
    } else {
        SSPBO_CUDA(c0, cudaFuncSetAttribute(k_exact_lowrank_runs<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_exact_lowrank_runs<float><<<gx, EXL_THREADS, smem, st>>>(de, R, parts);
    }
    SSPBO_LAUNCH_CHECK(c0);
    k_batch_status_runs<<<(R + 255) / 256, 256, 0, st>>>(de, R, dst);
    SSPBO_LAUNCH_CHECK(c0);
    c0->batch_status_valid = R;
    return SSPBO_OK;
}

extern "C" int sspbo_score_batch(sspbo_ctx* const* ctxs, int R, const void* x_dev, int x_dtype, int64_t N, const double* lam,
                                 const double* gamma_t, unsigned long long* keys_dev, void* const* streams, int n_streams,
                                 int* failed) {
    if (failed) *failed = -1;
    if (!ctxs || R < 0 || !lam || !gamma_t || !keys_dev || !streams || n_streams < 1) return SSPBO_EINVAL;
    if (R >= 1 && ctxs[0]) ctxs[0]->batch_status_valid = 0;
    if (R >= 2 && x_dev && (x_dtype == SSPBO_F32 || x_dtype == SSPBO_F64)) {
        // Runs that share a layout (lock-step: same rank): ONE launch of the grouped scorer and ONE of the grouped float64 pass,
        // which also leaves every run's status word (read by sspbo_score_batch_status); both on streams[0], which the caller
        // has ordered after the work before the lock-step.
        const int rc = score_batch_grouped(ctxs, R, x_dev, x_dtype, N, lam, gamma_t, keys_dev, (cudaStream_t)streams[0], failed);
        if (rc != SSPBO_EUNSUPPORTED) return rc;
    }
    for (int r = 0; r < R; ++r) {
        const int rc = sspbo_score(ctxs[r], x_dev, x_dtype, N, 0, lam[r], gamma_t[r], nullptr, nullptr, nullptr, keys_dev + r,
                                   SSPBO_ENGINE_AUTO, streams[r % n_streams]);
        if (rc) { if (failed) *failed = r; return rc; }
    }
    return SSPBO_OK;
}

// bit 0: the flag list of the run's low-rank / factor pass overflowed, bit 1: candidates outside the declared |x| bound; the
// counters are cleared for the next lock-step
__global__ void k_batch_status(unsigned long long* __restrict__ stats, unsigned long long* __restrict__ out) {
    *out = (stats[SSPBO_STAT_OVERFLOW] ? 1ull : 0ull) | (stats[SSPBO_STAT_OOR] ? 2ull : 0ull);
    for (int i = 0; i < SSPBO_STAT_COUNT; ++i) stats[i] = 0ull;
}

extern "C" int sspbo_score_batch_status(sspbo_ctx* const* ctxs, int R, unsigned long long* status_dev, void* const* streams,
                                        int n_streams, int* failed) {
    if (failed) *failed = -1;
    if (!ctxs || R < 0 || !status_dev || !streams || n_streams < 1) return SSPBO_EINVAL;
    if (R >= 1 && ctxs[0] && ctxs[0]->batch_status_valid == R) {          // the grouped float64 pass already produced the status words
        sspbo_ctx* c0 = ctxs[0];
        const size_t off_st = (size_t)R * (sizeof(sspbo_lr::RunScore) + sizeof(ExactArgs));
        if (cudaMemcpyAsync(status_dev, (const char*)c0->sbatch_dev + off_st, (size_t)R * sizeof(unsigned long long),
                            cudaMemcpyDeviceToDevice, (cudaStream_t)streams[0]) != cudaSuccess) { if (failed) *failed = 0; return SSPBO_ECUDA; }
        c0->batch_status_valid = 0;
        return SSPBO_OK;
    }
    for (int r = 0; r < R; ++r) {
        sspbo_ctx* ctx = ctxs[r];
        cudaStream_t st = (cudaStream_t)streams[r % n_streams];
        if (!ctx->stats) { if (cudaMemsetAsync(status_dev + r, 0, sizeof(unsigned long long), st) != cudaSuccess) { if (failed) *failed = r; return SSPBO_ECUDA; } continue; }
        k_batch_status<<<1, 1, 0, st>>>(ctx->stats, status_dev + r);
        if (cudaGetLastError() != cudaSuccess) { if (failed) *failed = r; return SSPBO_ECUDA; }
        ++ctx->launches;
    }
    return SSPBO_OK;
}

extern "C" int sspbo_update_batch(sspbo_ctx* const* ctxs, int R, const double* x_host, const double* ts_host, void* const* streams,
                                  int n_streams, int* failed) {
    if (failed) *failed = -1;
    if (!ctxs || R < 0 || !x_host || !ts_host || !streams || n_streams < 1) return SSPBO_EINVAL;
    if (R >= 2) {
        // plain SSP-space runs: every run's encode + update + operand refresh in ONE launch (a cluster per run); the caller has
        // ordered all `streams` after the work that precedes the lock-step, the launch goes to streams[0]
        const int rc = sspbo_blr_update_runs(ctxs, R, x_host, ts_host, (cudaStream_t)streams[0], failed);
        if (rc != SSPBO_EUNSUPPORTED) return rc;
    }
    for (int r = 0; r < R; ++r) {
        const size_t nl = (size_t)ctxs[r]->n * ctxs[r]->L;
        const int rc = sspbo_blr_update_x(ctxs[r], x_host + (size_t)r * nl, ts_host + r, 1, streams[r % n_streams]);
        if (rc) { if (failed) *failed = r; return rc; }
    }
    return SSPBO_OK;
}

extern "C" int sspbo_lowrank_info(const sspbo_ctx* ctx, int* rank, int* valid, int* disabled, int* phase32) {
    if (!ctx) return SSPBO_EINVAL;
    if (rank) *rank = ctx->lr_rank;
    if (valid) *valid = ctx->lr_valid ? 1 : 0;
    if (disabled) *disabled = ctx->lr_disabled ? 1 : 0;
    if (phase32) *phase32 = ((ctx->p32_ok || ctx->pf32_ok) && !ctx->p32_disabled) ? 1 : 0;
    return SSPBO_OK;
}
extern "C" int sspbo_set_policy(sspbo_ctx* ctx, int lowrank_enabled, int phase32_enabled) {
    if (!ctx) return SSPBO_EINVAL;
    if (lowrank_enabled >= 0) { ctx->lr_disabled = !lowrank_enabled; if (lowrank_enabled) { ctx->f8_disabled = false; ctx->lr_flag_frac = 0.0; } }
    if (phase32_enabled >= 0) {                          // 0: Q31.32 only; 1: fastest valid path; 2: fixed point only (no float32 chain)
        ctx->p32_disabled = (phase32_enabled == 0);
        ctx->pf32_disabled = (phase32_enabled == 2);
    }
    return SSPBO_OK;
}

extern "C" int sspbo_set_update_path(sspbo_ctx* ctx, int fused) {
    if (!ctx) return SSPBO_EINVAL;
    ctx->unfused_update = !fused;
    return SSPBO_OK;
}

extern "C" void sspbo_key_unpack(unsigned long long key, float* score, int64_t* index) {
    unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)(key & 0xffffffffu);
    if (index) *index = (int64_t)(0xffffffffu - lo);
    if (score) {
        unsigned int u;
        if (hi == 0xffffffffu) u = 0x7fc00000u;                      // NaN
        else u = (hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi;
        memcpy(score, &u, 4);
    }
}
extern "C" unsigned long long sspbo_key_pack(float score, int64_t index) { return sspbo_pack(score, (unsigned long long)index); }

// ------------------------------------------------------------------------------------ K4
extern "C" int sspbo_from_set_argmax(sspbo_ctx* ctx, const double* ssps, int64_t M, const double* queries, int q,
                                     int64_t* idx_out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (M < 1 || q < 1 || !ssps || !queries || !idx_out) SSPBO_FAIL(ctx, SSPBO_EINVAL, "from_set: empty set / null pointer");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int d = ctx->d;
    if (d < 1) SSPBO_FAIL(ctx, SSPBO_ESTATE, "from_set before space_set");
    double *unit = nullptr, *sims = nullptr;
    SSPBO_CUDA(ctx, cudaMallocAsync((void**)&unit, (size_t)q * d * sizeof(double), st));
    SSPBO_CUDA(ctx, cudaMallocAsync((void**)&sims, (size_t)q * M * sizeof(double), st));
    k_unit_rows<<<q, 256, 0, st>>>(queries, unit, d);
    SSPBO_LAUNCH_CHECK(ctx);
    k_sims<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(ssps, M, unit, q, d, sims);
    SSPBO_LAUNCH_CHECK(ctx);
    k_argmax_rows<<<q, 1024, 0, st>>>(sims, M, idx_out);
    SSPBO_LAUNCH_CHECK(ctx);
    SSPBO_CUDA(ctx, cudaFreeAsync(unit, st));
    SSPBO_CUDA(ctx, cudaFreeAsync(sims, st));
    return SSPBO_OK;
}

extern "C" int sspbo_decode_optim(sspbo_ctx* ctx, const double* queries, int q, const double* x0, const double* bounds_host,
                                  double* x_out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->K == 0) SSPBO_FAIL(ctx, SSPBO_ESTATE, "decode_optim before space_set");
    if (ctx->L != 1) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "decode_optim: trajectory contexts decode per waypoint through the x space");
    if (q < 0 || (q > 0 && (!queries || !x0 || !x_out))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "decode_optim: null pointer");
    if (ctx->n > DO_MAXN) SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "decode_optim: domain dimension %d > %d", ctx->n, DO_MAXN);
    if (q == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    double* bounds_dev = nullptr;
    if (bounds_host) {
        for (int a = 0; a < ctx->n; ++a)
            if (!(bounds_host[2 * a] <= bounds_host[2 * a + 1])) SSPBO_FAIL(ctx, SSPBO_EINVAL, "decode_optim: empty bound on axis %d", a);
        SSPBO_CUDA(ctx, cudaMallocAsync((void**)&bounds_dev, 2 * ctx->n * sizeof(double), st));
        SSPBO_CUDA(ctx, cudaMemcpyAsync(bounds_dev, bounds_host, 2 * ctx->n * sizeof(double), cudaMemcpyHostToDevice, st));
        SSPBO_CUDA(ctx, cudaStreamSynchronize(st));     // bounds_host may be a temporary of the caller
    }
    const size_t smem = (size_t)(2 * ctx->K + (DO_THREADS / 32) * DO_NRED) * sizeof(double);
    SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_decode_optim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_decode_optim<<<q, DO_THREADS, smem, st>>>(queries, x0, ctx->A_turns, ctx->twiddle, bounds_dev, ctx->n, ctx->K, 200, x_out);
    SSPBO_LAUNCH_CHECK(ctx);
    if (bounds_dev) SSPBO_CUDA(ctx, cudaFreeAsync(bounds_dev, st));
    return SSPBO_OK;
}

// ------------------------------------------------------------------------------------ K5
extern "C" int sspbo_bind(sspbo_ctx* ctx, const double* a, int qa, const double* b, int qb, double* out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    const int d = ctx->d;
    if (d < 1) SSPBO_FAIL(ctx, SSPBO_ESTATE, "bind before space_set");
    if (!a || !b || !out || qa < 1 || qb < 1 || (qa != qb && qa != 1 && qb != 1))
        SSPBO_FAIL(ctx, SSPBO_EINVAL, "bind: row counts %d and %d do not broadcast", qa, qb);
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    const int q = qa > qb ? qa : qb;
    size_t smem = 2 * (size_t)d * sizeof(double);
    SSPBO_CUDA(ctx, cudaFuncSetAttribute(k_bind, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bind<<<dim3((d + 255) / 256, q), 256, smem, (cudaStream_t)stream>>>(a, qa, b, qb, d, out);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

extern "C" int sspbo_gram(sspbo_ctx* ctx, const double* phi, int n, double* gram_out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (ctx->d < 1) SSPBO_FAIL(ctx, SSPBO_ESTATE, "gram before space_set / blr_init_dim");
    if (n < 0 || n > 65535 || (n > 0 && (!phi || !gram_out))) SSPBO_FAIL(ctx, SSPBO_EINVAL, "gram: need 0 <= n <= 65535 rows");
    if (n == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    k_gram<<<dim3(n, n), 128, 0, (cudaStream_t)stream>>>(phi, n, ctx->d, gram_out);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

// ------------------------------------------------------------------------------------ generators
extern "C" int sspbo_fill_uniform(sspbo_ctx* ctx, uint64_t seed, int64_t start, int64_t count, int n, float* out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (count < 0 || n < 1 || start < 0 || (count > 0 && !out)) SSPBO_FAIL(ctx, SSPBO_EINVAL, "fill_uniform: bad arguments");
    if (count == 0) return SSPBO_OK;
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t total = count * n;
    int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sm_count * 16 ? (total + 255) / 256 : (int64_t)ctx->sm_count * 16);
    k_fill_uniform<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, start * n, total, out);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}

extern "C" int sspbo_grid_points(sspbo_ctx* ctx, const double* axes, const int* counts, int n, int64_t start, int64_t count,
                                 double* out, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    if (n < 1 || n > 16 || !axes || !counts || count < 0 || start < 0 || (count > 0 && !out))
        SSPBO_FAIL(ctx, SSPBO_EINVAL, "grid_points: bad arguments");
    if (count == 0) return SSPBO_OK;
    GridDesc g; g.n = n;
    int off = 0; double total = 1.0;
    for (int a = 0; a < n; ++a) {
        if (counts[a] < 1) SSPBO_FAIL(ctx, SSPBO_EINVAL, "grid_points: counts must be >= 1");
        g.counts[a] = counts[a]; g.offs[a] = off; off += counts[a]; total *= counts[a];
    }
    if ((double)(start + count) > total) SSPBO_FAIL(ctx, SSPBO_EINVAL, "grid_points: range exceeds the grid");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    int blocks = (int)((count + 255) / 256 < (int64_t)ctx->sm_count * 16 ? (count + 255) / 256 : (int64_t)ctx->sm_count * 16);
    k_grid_points<<<blocks, 256, 0, (cudaStream_t)stream>>>(axes, g, start, count, out);
    SSPBO_LAUNCH_CHECK(ctx);
    return SSPBO_OK;
}
