// K2, tcgen05/TMEM engine (placeholder until the kernel lands; reports "unsupported").
#include "sspbo_internal.cuh"
int sspbo_score_umma_supported(const sspbo_ctx*) { return 0; }
int sspbo_score_umma_launch(sspbo_ctx* ctx, const void*, int, int64_t, int64_t, float, float, float*, float*, float*,
                            unsigned long long*, cudaStream_t) {
    SSPBO_FAIL(ctx, SSPBO_EUNSUPPORTED, "tcgen05 engine not built");
}
void sspbo_umma_release(sspbo_ctx*) {}
