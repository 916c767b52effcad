// common.cuh -- context, error plumbing and buffer staging shared by every translation unit
// of librupphash_b200.so.  Nothing here depends on PyTorch; the library is plain CUDA C++.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rupphash_b200.h"

struct rh_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;   // compute stream owned by the ctx
    cudaStream_t stream = nullptr;       // stream in use (own_stream or the caller's)
    cudaStream_t copy_stream = nullptr;  // H2D staging stream of the batching pipeline
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
    std::string err;
    uint64_t launches = 0;
    double last_ms = 0.0, last_units = 0.0;
    int last_hamming_variant = -1;   // rh_hamming_last_variant
    int sm_count = 148;
    // growable device scratch, one buffer per slot
    static constexpr int kSlots = 28;
    void *slot_ptr[kSlots] = {};
    size_t slot_bytes[kSlots] = {};
    // growable pinned host scratch
    static constexpr int kHostSlots = 4;
    void *hslot_ptr[kHostSlots] = {};
    size_t hslot_bytes[kHostSlots] = {};
    // buffers a growing slot left behind: queued work may still use them, so they are freed later, when the
    // ctx is idle (rh::reap_retired) -- growing a slot never waits for the device
    std::vector<void *> retired, retired_host;
    size_t retired_bytes = 0;
    bool dct_ready = false;
    bool timing_pending = false;   // ev_a / ev_b bracket a kernel whose time has not been read yet
    // tuning knobs of rh_ctx_set_option (benchmarks / A-B runs); the defaults are the product path
    int force_prefilter = -1;      // "hamming.prefilter": -1 = sampled selectivity, 0 / 3..7 = pin the variant
    int pdq_force_generic = 0;     // "pdq.force_generic"
    int pdq_prefetch = 2;          // "pdq.prefetch": L2 prefetch mode of the fused kernel's front end
    int pdq_prefetch_rows = 32;    // "pdq.prefetch_rows"
    int pdq_phase_clocks = 0;      // "pdq.phase_clocks": print per-phase cycle shares after each fused launch
    int pdq_variant = 0;           // "pdq.variant": 0 = default front end, other values = experiments
    static constexpr int kTickets = 8;   // completion events of the asynchronous calls (rh_ctx_wait)
    cudaEvent_t ev_ticket[kTickets] = {};
    uint64_t tickets_issued = 0;
    uint64_t stage_seq = 0;        // chunks staged so far (alternates the two H2D staging buffers across calls)
};

namespace rh {

enum Slot {
    S_IN0 = 0, S_IN1, S_IN2, S_IN3, S_IN4,   // staged inputs
    S_OUT0, S_OUT1, S_OUT2, S_OUT3, S_OUT4,  // staged outputs
    S_W0, S_W1, S_W2, S_W3, S_W4, S_W5, S_W6, S_W7, S_W8, S_W9, S_W10, S_W11, S_W12, S_W13,  // work
    S_W14, S_W15, S_DCT  // S_DCT holds the PDQ DCT matrix for the life of the ctx
};

inline int fail(rh_ctx *ctx, int code, const char *what, cudaError_t ce = cudaSuccess) {
    if (ctx) {
        char buf[512];
        if (ce != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
        else
            snprintf(buf, sizeof buf, "%s", what);
        ctx->err = buf;
    }
    return code;
}

#define RH_CUDA(ctx, call)                                                        \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return rh::fail((ctx), RH_ECUDA, #call, _e);       \
    } while (0)

#define RH_TRY(expr)                  \
    do {                              \
        int _s = (expr);              \
        if (_s != RH_OK) return _s;   \
    } while (0)

// Launch check: counts the launch and turns a launch failure into RH_ECUDA.
#define RH_LAUNCHED(ctx, name)                                                    \
    do {                                                                          \
        (ctx)->launches++;                                                        \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess) return rh::fail((ctx), RH_ECUDA, "launch " name, _e); \
    } while (0)

// Frees the buffers that growing slots left behind once nothing queued can still use them: `force` after the
// caller has synchronised both streams, otherwise only if both streams are idle right now (cudaFree would wait
// for the device otherwise).
inline void reap_retired(rh_ctx *ctx, bool force) {
    if (ctx->retired.empty() && ctx->retired_host.empty()) return;
    if (!force) {
        const bool idle = cudaStreamQuery(ctx->stream) == cudaSuccess && cudaStreamQuery(ctx->copy_stream) == cudaSuccess;
        if (!idle) {
            cudaGetLastError();   // cudaErrorNotReady is an answer, not a failure (launch errors are checked at the launch)
            return;
        }
    }
    for (void *p : ctx->retired) cudaFree(p);
    for (void *p : ctx->retired_host) cudaFreeHost(p);
    ctx->retired.clear();
    ctx->retired_host.clear();
    ctx->retired_bytes = 0;
}

inline int scratch(rh_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (ctx->slot_bytes[slot] < bytes) {
        reap_retired(ctx, false);
        if (ctx->slot_ptr[slot]) {
            // the old buffer may still be in use by queued work: it is retired, not freed, so growing a slot does
            // not wait for the device (round 1 synchronised both streams here); above 16 GiB of retired memory
            // the wait is taken after all
            ctx->retired.push_back(ctx->slot_ptr[slot]);
            ctx->retired_bytes += ctx->slot_bytes[slot];
            ctx->slot_ptr[slot] = nullptr;
            ctx->slot_bytes[slot] = 0;
            if (ctx->retired_bytes > (size_t(16) << 30)) {
                RH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                RH_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
                reap_retired(ctx, true);
            }
        }
        size_t want = bytes + bytes / 8;
        want = (want + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&ctx->slot_ptr[slot], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, RH_ENOMEM, "cudaMalloc scratch", e);
        }
        ctx->slot_bytes[slot] = want;
    }
    *out = ctx->slot_ptr[slot];
    return RH_OK;
}

inline int host_scratch(rh_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (ctx->hslot_bytes[slot] < bytes) {
        reap_retired(ctx, false);
        if (ctx->hslot_ptr[slot]) {
            ctx->retired_host.push_back(ctx->hslot_ptr[slot]);
            ctx->hslot_ptr[slot] = nullptr;
            ctx->hslot_bytes[slot] = 0;
        }
        cudaError_t e = cudaMallocHost(&ctx->hslot_ptr[slot], bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, RH_ENOMEM, "cudaMallocHost scratch", e);
        }
        ctx->hslot_bytes[slot] = bytes;
    }
    *out = ctx->hslot_ptr[slot];
    return RH_OK;
}

// true when p is device-accessible memory of some CUDA device (device or managed)
inline bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Input staging: a device pointer is used in place, a host pointer is copied into `slot`.
template <typename T>
inline int stage_in(rh_ctx *ctx, const T *p, size_t count, int slot, const T **out) {
    if (!p) {
        *out = nullptr;
        return RH_OK;
    }
    if (is_device_ptr(p)) {
        *out = p;
        return RH_OK;
    }
    void *d = nullptr;
    RH_TRY(scratch(ctx, slot, count * sizeof(T), &d));
    RH_CUDA(ctx, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = static_cast<const T *>(d);
    return RH_OK;
}

// Output staging: device pointer in place; host pointer gets a scratch buffer and a
// finish() that copies it back.
template <typename T>
struct OutBuf {
    T *dev = nullptr;
    T *user = nullptr;
    size_t count = 0;
    bool staged = false;
    int prepare(rh_ctx *ctx, T *p, size_t n, int slot) {
        user = p;
        count = n;
        if (!p) return RH_OK;
        if (is_device_ptr(p)) {
            dev = p;
            return RH_OK;
        }
        void *d = nullptr;
        RH_TRY(scratch(ctx, slot, n * sizeof(T), &d));
        dev = static_cast<T *>(d);
        staged = true;
        return RH_OK;
    }
    int finish(rh_ctx *ctx) {
        if (staged && count)
            RH_CUDA(ctx, cudaMemcpyAsync(user, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        return RH_OK;
    }
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace rh
