// Multi-GPU winner selection over peer memory (one process per GPU, NVLink / NVSwitch): every rank's packed (score, index) key
// and its "rerun" flag are exchanged by ONE small kernel per step that writes them straight into every peer's slot array and
// waits for the peers' words to arrive -- the all_reduce(MAX) of SURVEY.md section 8e without a collective library in the
// step (settle + xor + NCCL all-reduce + xor + read-back become: this launch + the usual read-back of sspbo_score_finish).
//
// Protocol.  Rank r owns a slot array of 2 x world x 4 words (two buffers used alternately by the parity of the exchange
// number, so a fast rank's next exchange never overwrites words a slow rank has yet to read: a rank cannot finish exchange
// e + 1 before every peer has started it, i.e. finished e).  In exchange e lane p of the one warp writes {key, flag, e} into
// slot r of peer p's array (the epoch word last, after a system-scope fence), then polls slot p of its own array until the
// epoch word equals e.  The warp then reduces max(key) / or(flag); lane 0 stores the global key over the local one and the
// flag into the context's counters, where sspbo_score_finish finds it.  A bounded poll (about ten seconds: ranks may be
// that far apart on their first step) turns a missing peer into an error instead of a hang.
#include "sspbo_internal.cuh"
#include <string.h>

namespace {

constexpr int DIST_MAX_WORLD = 16;
constexpr int SLOT_WORDS = 4;

struct DistState {
    int rank = 0, world = 1;
    unsigned long long* slots = nullptr;                     // 2 x world x SLOT_WORDS, cudaMalloc'ed (IPC-exportable)
    unsigned long long* peer[DIST_MAX_WORLD] = {};           // peer[p]: rank p's slot array mapped into this process
    bool opened[DIST_MAX_WORLD] = {};
    unsigned long long epoch = 0;
    bool connected = false;
};

struct ExchangeArgs {
    unsigned long long* peer[DIST_MAX_WORLD];
    unsigned long long* mine;
    unsigned long long* best;
    unsigned long long* stats;
    unsigned long long epoch;
    int rank, world;
};

__global__ void k_key_exchange(const ExchangeArgs a) {
    const int p = threadIdx.x;
    const unsigned long long key = *a.best;
    const unsigned long long flag = (a.stats && (a.stats[SSPBO_STAT_OVERFLOW] || a.stats[SSPBO_STAT_OOR])) ? 1ull : 0ull;
    const size_t buf = (size_t)(a.epoch & 1ull) * a.world * SLOT_WORDS;
    unsigned long long k = 0ull, f = 0ull;
    if (p < a.world) {
        volatile unsigned long long* dst = a.peer[p] + buf + (size_t)a.rank * SLOT_WORDS;
        dst[0] = key; dst[1] = flag;
        __threadfence_system();
        dst[2] = a.epoch;
        volatile unsigned long long* src = a.mine + buf + (size_t)p * SLOT_WORDS;
        const long long t0 = clock64();
        bool ok = true;
        while (src[2] != a.epoch) {
            if (clock64() - t0 > 20000000000ll) { ok = false; break; }                 // ~10 s: a peer never arrived
            __nanosleep(200);
        }
        __threadfence_system();
        if (ok) { k = src[0]; f = src[1]; } else f = 2ull;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ko = __shfl_xor_sync(0xffffffffu, k, o);
        const unsigned long long fo = __shfl_xor_sync(0xffffffffu, f, o);
        k = ko > k ? ko : k; f |= fo;
    }
    if (p == 0) {
        *a.best = k;
        if (a.stats) a.stats[SSPBO_STAT_PEER] = f;
    }
}

}  // namespace

void sspbo_dist_release(sspbo_ctx* ctx) {
    DistState* ds = (DistState*)ctx->dist_state;
    if (!ds) return;
    for (int p = 0; p < ds->world; ++p)
        if (ds->opened[p]) cudaIpcCloseMemHandle(ds->peer[p]);
    if (ds->slots) cudaFree(ds->slots);
    delete ds;
    ctx->dist_state = nullptr;
}

/* Step 1 of the set-up: this rank's slot array; handle_out (64 bytes) is what the other ranks need to map it. */
extern "C" int sspbo_dist_init(sspbo_ctx* ctx, int rank, int world, unsigned char* handle_out) {
    if (!ctx) return SSPBO_EINVAL;
    if (!handle_out || world < 1 || world > DIST_MAX_WORLD || rank < 0 || rank >= world)
        SSPBO_FAIL(ctx, SSPBO_EINVAL, "dist_init: need 0 <= rank < world <= %d", DIST_MAX_WORLD);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    sspbo_dist_release(ctx);
    DistState* ds = new DistState();
    ctx->dist_state = ds;
    ds->rank = rank; ds->world = world;
    const size_t bytes = (size_t)2 * world * SLOT_WORDS * sizeof(unsigned long long);
    SSPBO_CUDA(ctx, cudaMalloc((void**)&ds->slots, bytes));
    SSPBO_CUDA(ctx, cudaMemset(ds->slots, 0xff, bytes));                 // epoch words start at a value no exchange uses
    SSPBO_CUDA(ctx, cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    SSPBO_CUDA(ctx, cudaIpcGetMemHandle(&h, ds->slots));
    memcpy(handle_out, &h, sizeof(h));
    return SSPBO_OK;
}

/* Step 2: the handles of all ranks (world x 64 bytes, in rank order) -> peer mappings.  Every rank calls it with the same
 * array; a rank that fails here must tell the others (the caller agrees on the outcome before the first exchange). */
extern "C" int sspbo_dist_connect(sspbo_ctx* ctx, const unsigned char* handles) {
    if (!ctx) return SSPBO_EINVAL;
    DistState* ds = (DistState*)ctx->dist_state;
    if (!ds || !handles) SSPBO_FAIL(ctx, SSPBO_ESTATE, "dist_connect before dist_init");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int p = 0; p < ds->world; ++p) {
        if (p == ds->rank) { ds->peer[p] = ds->slots; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)p * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        SSPBO_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ds->peer[p] = (unsigned long long*)ptr; ds->opened[p] = true;
    }
    ds->connected = true;
    return SSPBO_OK;
}

/* The same for ranks that live in ONE process (a thread per GPU, or several contexts in a test): IPC handles cannot be opened
 * by the process that exported them, so such ranks hand each other the slot arrays themselves.  slots_out: this rank's array
 * (after sspbo_dist_init); slot_ptrs: the arrays of all ranks in rank order (peer access between their devices enabled by
 * the caller). */
extern "C" int sspbo_dist_slots(sspbo_ctx* ctx, void** slots_out) {
    if (!ctx || !slots_out) return SSPBO_EINVAL;
    DistState* ds = (DistState*)ctx->dist_state;
    if (!ds) SSPBO_FAIL(ctx, SSPBO_ESTATE, "dist_slots before dist_init");
    *slots_out = ds->slots;
    return SSPBO_OK;
}
extern "C" int sspbo_dist_connect_ptrs(sspbo_ctx* ctx, void* const* slot_ptrs) {
    if (!ctx) return SSPBO_EINVAL;
    DistState* ds = (DistState*)ctx->dist_state;
    if (!ds || !slot_ptrs) SSPBO_FAIL(ctx, SSPBO_ESTATE, "dist_connect_ptrs before dist_init");
    for (int p = 0; p < ds->world; ++p) {
        if (!slot_ptrs[p]) SSPBO_FAIL(ctx, SSPBO_EINVAL, "dist_connect_ptrs: null slot array of rank %d", p);
        ds->peer[p] = (p == ds->rank) ? ds->slots : (unsigned long long*)slot_ptrs[p];
    }
    ds->connected = true;
    return SSPBO_OK;
}

extern "C" int sspbo_dist_ready(const sspbo_ctx* ctx) {
    const DistState* ds = ctx ? (const DistState*)ctx->dist_state : nullptr;
    return (ds && ds->connected) ? 1 : 0;
}

/* One exchange: best_dev (this rank's packed key, device) is replaced by the maximum over all ranks; the counters of the
 * context receive whether any rank asks for a rerun (sspbo_score_finish reports it).  Asynchronous on `stream`; every rank
 * must call it the same number of times. */
extern "C" int sspbo_dist_exchange(sspbo_ctx* ctx, unsigned long long* best_dev, void* stream) {
    if (!ctx) return SSPBO_EINVAL;
    DistState* ds = (DistState*)ctx->dist_state;
    if (!ds || !ds->connected) SSPBO_FAIL(ctx, SSPBO_ESTATE, "dist_exchange before dist_connect");
    if (!best_dev) SSPBO_FAIL(ctx, SSPBO_EINVAL, "dist_exchange: null key");
    SSPBO_CUDA(ctx, cudaSetDevice(ctx->device));
    ExchangeArgs a;
    memset(&a, 0, sizeof(a));
    for (int p = 0; p < ds->world; ++p) a.peer[p] = ds->peer[p];
    a.mine = ds->slots; a.best = best_dev; a.stats = ctx->stats; a.epoch = ds->epoch++; a.rank = ds->rank; a.world = ds->world;
    k_key_exchange<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    SSPBO_LAUNCH_CHECK(ctx);
    ctx->launches++;
    return SSPBO_OK;
}
