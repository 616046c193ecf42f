/* Pure-C caller of the drop-in boundary (include/sspbo.h): no Python, no torch -- cudart for the caller-owned device
 * buffers only.  One BO step on a small random SSP space: space, posterior from 16 observations, fused scoring of
 * generated candidates with every engine, key read-back, observation.  Prints one line per engine; the GPU test
 * (tests/test_gpu_parity.py::test_c_abi_from_plain_c) builds it with gcc and checks the engines agree.
 *   gcc -std=c99 -Iinclude tests/c_abi_smoke.c -Lssp_bayesopt_b200/lib -lsspbo -L/usr/local/cuda/lib64 -lcudart -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime_api.h>
#include "sspbo.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? sspbo_last_error(ctx) : ""); return 1; } \
    } while (0)

static double lcg(unsigned long long* s) {                 /* deterministic uniform(0, 1) for the set-up */
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(*s >> 11) / 9007199254740992.0;
}

int main(void) {
    enum { K = 100, NDIM = 3, D = 2 * K + 1, NOBS = 16, NCAND = 20000 };
    sspbo_ctx* ctx = NULL;
    unsigned long long seed = 42;
    double A[K * NDIM], inv_ls[NDIM] = {2.5, 2.5, 2.5}, xs[NOBS * NDIM], ts[NOBS];
    for (int i = 0; i < K * NDIM; ++i) A[i] = (2.0 * lcg(&seed) - 1.0) * 3.14159;
    for (int i = 0; i < NOBS; ++i) {
        double sum = 0.0;
        for (int a = 0; a < NDIM; ++a) { xs[i * NDIM + a] = lcg(&seed); sum += xs[i * NDIM + a]; }
        ts[i] = sin(3.0 * sum);
    }
    if (sspbo_ctx_create(0, &ctx) != 0) { fprintf(stderr, "no device context\n"); return 1; }
    CHECK(sspbo_space_set(ctx, A, inv_ls, K, NDIM));
    CHECK(sspbo_space_set_xbound(ctx, 1.0));
    CHECK(sspbo_blr_init(ctx, 0.01));
    CHECK(sspbo_blr_set_beta(ctx, 4.0));
    CHECK(sspbo_blr_update_x(ctx, xs, ts, NOBS, NULL));          /* encode + 16 posterior updates on the device */

    float* cand = NULL; unsigned long long* key_dev = NULL;
    if (cudaMalloc((void**)&cand, sizeof(float) * NCAND * NDIM) != cudaSuccess) return 1;
    if (cudaMalloc((void**)&key_dev, sizeof(unsigned long long)) != cudaSuccess) return 1;
    CHECK(sspbo_fill_uniform(ctx, 7, 0, NCAND, NDIM, cand, NULL));

    const int engines[] = {SSPBO_ENGINE_AUTO, SSPBO_ENGINE_UMMA_LR, SSPBO_ENGINE_UMMA, SSPBO_ENGINE_SIMT};
    long long first_index = -1;
    for (int e = 0; e < 4; ++e) {
        unsigned long long key = 0;
        int64_t scored, flagged, overflow, oor, index;
        int rerun;
        float score;
        cudaMemset(key_dev, 0, sizeof(key));
        CHECK(sspbo_score(ctx, cand, SSPBO_F32, NCAND, 0, 10.0, 0.0, NULL, NULL, NULL, key_dev, engines[e], NULL));
        CHECK(sspbo_score_finish(ctx, key_dev, &key, &scored, &flagged, &overflow, &oor, &rerun, 1, NULL));
        sspbo_key_unpack(key, &score, &index);
        printf("engine %d (ran as %d): score %.6f index %lld rerun %d flagged %lld\n", engines[e], sspbo_last_engine(ctx),
               score, (long long)index, rerun, (long long)flagged);
        if (rerun) return 2;
        if (first_index < 0) first_index = index;
    }
    /* observe the winner: predictive mean / variance under the current posterior, then the update */
    float row[NDIM];
    double x_t[NDIM], mu, var, mu2, var2, acq2;
    cudaMemcpy(row, cand + first_index * NDIM, sizeof(row), cudaMemcpyDeviceToHost);
    for (int a = 0; a < NDIM; ++a) x_t[a] = row[a];
    CHECK(sspbo_eval_host(ctx, x_t, 1, 10.0, 0.0, &mu2, &var2, &acq2, NULL));
    CHECK(sspbo_observe(ctx, x_t, 0.25, &mu, &var, NULL));
    printf("observe: mu %.6f var %.6f | eval_host: mu %.6f var %.6f acq %.6f\n", mu, var, mu2, var2, acq2);
    if (fabs(mu - mu2) > 1e-4 * (1.0 + fabs(mu2)) || fabs(var - var2) > 1e-4 * var2) return 3;
    cudaFree(cand); cudaFree(key_dev);
    sspbo_ctx_destroy(ctx);
    printf("c abi ok\n");
    return 0;
}
