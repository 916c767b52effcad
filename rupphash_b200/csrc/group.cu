// group.cu -- rh_group: the GPUs of one box driven from ONE process through the C ABI (no Python, no
// torch): what a Rust caller of scanner::group_with_pdqhash (scanner.rs:1550-1551, :1827-1832) binds
// to use 2 / 4 / 8 B200s.
//
//   inputs      host buffers are cut into one slice per GPU, every GPU copies its slice over its own
//               PCIe link and an NCCL all-gather over NVLink replicates them; a buffer that already
//               lives on one GPU of the group is broadcast from there
//   search      every GPU runs the persistent tile kernel of hamming.cu on the tiles t = rank (mod n_dev),
//               claimed by its CTAs from a local counter.  RH_GROUP_STEAL_TILES: ONE claim counter in GPU 0's
//               memory instead (system-scope atomics over NVLink, peer access): the GPUs steal tiles from a
//               common pool, which absorbs start-up skew and clock differences between GPUs -- measured on
//               8 equal B200s it is 1 % slower than the static split (7.99 vs 7.91 ms per GPU at 500k
//               hashes), so it is opt-in
//   exchange    ncclAllGather of the n x u32 forests + ncclAllReduce of the edge counts, then the merge
//               (rh_uf_merge's kernels) on GPU 0 and the labels go to the caller
//
// NCCL is loaded with dlopen at rh_group_create, so librupphash_b200.so itself has no NCCL dependency
// and single-GPU users never need it; RH_GROUP_NO_NCCL exchanges with cudaMemcpyPeerAsync instead.
// One worker thread per GPU issues that GPU's work, so the GPUs start within microseconds of each other.
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "hamming_internal.cuh"

#if __has_include(<nccl.h>)
#include <nccl.h>
#else
// the handful of declarations used below (stable since NCCL 2.0)
typedef struct ncclComm *ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
#endif

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string *why) {
        // RH_NCCL_LIB names the copy to use; otherwise a libnccl.so.2 the process already holds (a host that
        // also runs torch has torch's bundled one) is shared, and only then the loader's search path is used --
        // two different NCCL builds behind one soname cannot live in one process
        const char *names[] = {getenv("RH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (names[0] && *names[0] && !handle) handle = dlopen(names[0], RTLD_NOW | RTLD_GLOBAL);
        for (int k = 1; k < 3 && !handle; k++) handle = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
        if (!handle) {
            *why = std::string("dlopen libnccl.so.2: ") + dlerror();
            return false;
        }
#define RH_SYM(field, name)                                              \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));      \
    if (!field) {                                                        \
        *why = std::string("libnccl lacks ") + name;                     \
        return false;                                                    \
    }
        RH_SYM(GetVersion, "ncclGetVersion")
        RH_SYM(CommInitAll, "ncclCommInitAll")
        RH_SYM(CommDestroy, "ncclCommDestroy")
        RH_SYM(AllGather, "ncclAllGather")
        RH_SYM(Broadcast, "ncclBroadcast")
        RH_SYM(AllReduce, "ncclAllReduce")
        RH_SYM(GroupStart, "ncclGroupStart")
        RH_SYM(GroupEnd, "ncclGroupEnd")
        RH_SYM(GetErrorString, "ncclGetErrorString")
#undef RH_SYM
        return true;
    }
};

// one thread per GPU, parked on a condition variable between calls
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    void loop() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> j = std::move(job);
            has_job = false;
            lk.unlock();
            int r = j();
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    }
};

}  // namespace

struct rh_group {
    int n_dev = 0;
    std::vector<int> devices;
    std::vector<rh_ctx *> ctx;
    std::vector<ncclComm_t> comms;
    std::vector<Worker *> workers;
    NcclApi nccl;
    bool use_nccl = false, steal = false;
    int nccl_version = 0;
    unsigned long long *shared_counter = nullptr;   // in devices[0]'s memory
    cudaEvent_t ev_reset = nullptr;
    std::vector<cudaEvent_t> ev_forest;             // forest of device i is ready (peer-copy exchange)
    std::string err;
    double times[12] = {};
    cudaEvent_t ev_t[4] = {};                       // GPU 0 timeline marks of the last group call
    std::mutex err_m;
    void set_err(const std::string &s) {
        std::lock_guard<std::mutex> lk(err_m);
        if (err.empty()) err = s;
    }
    int run_all(const std::function<int(int)> &fn) {
        for (int i = 0; i < n_dev; i++) {
            Worker *w = workers[i];
            std::lock_guard<std::mutex> lk(w->m);
            w->job = [fn, i] { return fn(i); };
            w->has_job = true;
            w->done = false;
            w->cv.notify_all();
        }
        int rc = RH_OK;
        for (int i = 0; i < n_dev; i++) {
            Worker *w = workers[i];
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->done; });
            if (w->rc != RH_OK && rc == RH_OK) rc = w->rc;
        }
        return rc;
    }
};

namespace {

using namespace rh;

#define RH_NCCL(g, ctx, call)                                                                     \
    do {                                                                                          \
        ncclResult_t _r = (call);                                                                 \
        if (_r != ncclSuccess) {                                                                  \
            std::string _m = std::string(#call) + ": " + (g)->nccl.GetErrorString(_r);            \
            (g)->set_err(_m);                                                                     \
            return fail((ctx), RH_ENCCL, _m.c_str());                                             \
        }                                                                                         \
    } while (0)

// which device of the group owns p (-1: host memory, -2: a device outside the group)
int owner_of(const rh_group *g, const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
    for (int i = 0; i < g->n_dev; i++)
        if (g->devices[i] == a.device) return i;
    return -2;
}

struct Input {
    const uint8_t *src = nullptr;
    size_t bytes = 0;
    int slot = 0;
    int owner = -1;
};

// Replicates one input array on device `i`; *out = the device-local pointer the search reads.
int replicate_input(rh_group *g, int i, const Input &in, const uint8_t **out) {
    rh_ctx *ctx = g->ctx[i];
    cudaStream_t st = ctx->stream;
    *out = nullptr;
    if (!in.src) return RH_OK;
    const int world = g->n_dev;
    if (in.owner >= 0) {   // device-resident on one GPU of the group
        if (world == 1 || in.owner == i) {
            *out = in.src;
            if (world == 1) return RH_OK;
        }
        void *buf = nullptr;
        if (in.owner != i) {
            RH_TRY(scratch(ctx, in.slot, in.bytes, &buf));
            *out = (const uint8_t *)buf;
        }
        if (g->use_nccl) {
            void *recv = in.owner == i ? (void *)in.src : buf;   // in place on the root
            RH_NCCL(g, ctx, g->nccl.Broadcast(recv, recv, in.bytes, ncclUint8, in.owner, g->comms[i], st));
        } else if (in.owner != i) {
            RH_CUDA(ctx, cudaMemcpyPeerAsync(buf, g->devices[i], in.src, g->devices[in.owner], in.bytes, st));
        }
        return RH_OK;
    }
    // host memory
    if (world == 1 || !g->use_nccl) {   // every GPU pulls the whole array over its own PCIe link
        void *buf;
        RH_TRY(scratch(ctx, in.slot, in.bytes, &buf));
        RH_CUDA(ctx, cudaMemcpyAsync(buf, in.src, in.bytes, cudaMemcpyHostToDevice, st));
        *out = (const uint8_t *)buf;
        return RH_OK;
    }
    // one slice per GPU over PCIe, then an all-gather over NVLink
    const size_t chunk = ((in.bytes + world - 1) / world + 15) & ~size_t(15);
    void *buf;
    RH_TRY(scratch(ctx, in.slot, chunk * world, &buf));
    const size_t lo = chunk * (size_t)i, hi = std::min(in.bytes, chunk * (size_t)(i + 1));
    if (hi > lo) RH_CUDA(ctx, cudaMemcpyAsync((uint8_t *)buf + lo, in.src + lo, hi - lo, cudaMemcpyHostToDevice, st));
    RH_NCCL(g, ctx, g->nccl.AllGather((const uint8_t *)buf + lo, buf, chunk, ncclUint8, g->comms[i], st));
    *out = (const uint8_t *)buf;
    return RH_OK;
}

}  // namespace

extern "C" {

int rh_group_create(const int *devices, int n_dev, unsigned flags, rh_group **out) {
    if (!out) return RH_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return RH_ECUDA;   // no CPU fallback
    }
    if (n_dev <= 0) n_dev = devices ? 0 : count;
    if (n_dev <= 0 || n_dev > count || n_dev > 64) return RH_EINVAL;
    rh_group *g = new (std::nothrow) rh_group();
    if (!g) return RH_ENOMEM;
    g->n_dev = n_dev;
    for (int i = 0; i < n_dev; i++) {
        const int d = devices ? devices[i] : i;
        for (int k = 0; k < i; k++)
            if (g->devices[k] == d) {
                delete g;
                return RH_EINVAL;
            }
        g->devices.push_back(d);
    }
    int rc = RH_OK;
    for (int i = 0; i < n_dev && rc == RH_OK; i++) {
        rh_ctx *c = nullptr;
        rc = rh_ctx_create(g->devices[i], &c);
        if (rc == RH_OK) g->ctx.push_back(c);
    }
    if (rc != RH_OK) {
        rh_group_destroy(g);
        return rc;
    }
    // peer access: every GPU reaches every other GPU's memory (claim counter, peer copies)
    bool peers = n_dev > 1;
    for (int i = 0; i < n_dev; i++) {
        cudaSetDevice(g->devices[i]);
        for (int k = 0; k < n_dev; k++) {
            if (k == i) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->devices[i], g->devices[k]);
            if (!can) {
                peers = false;
                continue;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[k], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peers = false;
            cudaGetLastError();
        }
    }
    g->steal = peers && (flags & RH_GROUP_STEAL_TILES) && !(flags & RH_GROUP_STATIC_TILES);
    cudaSetDevice(g->devices[0]);
    if (cudaMalloc(&g->shared_counter, 64) != cudaSuccess ||
        cudaEventCreateWithFlags(&g->ev_reset, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        rh_group_destroy(g);
        return RH_ECUDA;
    }
    cudaSetDevice(g->devices[0]);
    for (int k = 0; k < 4; k++)
        if (cudaEventCreate(&g->ev_t[k]) != cudaSuccess) {
            cudaGetLastError();
            rh_group_destroy(g);
            return RH_ECUDA;
        }
    for (int i = 0; i < n_dev; i++) {
        cudaSetDevice(g->devices[i]);
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            rh_group_destroy(g);
            return RH_ECUDA;
        }
        g->ev_forest.push_back(e);
    }
    if (n_dev > 1 && !(flags & RH_GROUP_NO_NCCL)) {
        std::string why;
        if (!g->nccl.load(&why)) {
            fprintf(stderr, "rh_group_create: %s\n", why.c_str());
            rh_group_destroy(g);
            return RH_ENCCL;
        }
        g->nccl.GetVersion(&g->nccl_version);
        g->comms.assign(n_dev, nullptr);
        ncclResult_t r = g->nccl.CommInitAll(g->comms.data(), n_dev, g->devices.data());
        if (r != ncclSuccess) {
            fprintf(stderr, "rh_group_create: ncclCommInitAll: %s\n", g->nccl.GetErrorString(r));
            g->comms.clear();
            rh_group_destroy(g);
            return RH_ENCCL;
        }
        g->use_nccl = true;
    } else if (n_dev > 1 && !peers) {
        fprintf(stderr, "rh_group_create: RH_GROUP_NO_NCCL needs peer access between all GPUs\n");
        rh_group_destroy(g);
        return RH_EUNSUPPORTED;
    }
    for (int i = 0; i < n_dev; i++) {
        Worker *w = new Worker();
        w->th = std::thread([w] { w->loop(); });
        g->workers.push_back(w);
    }
    *out = g;
    return RH_OK;
}

int rh_group_destroy(rh_group *g) {
    if (!g) return RH_OK;
    for (Worker *w : g->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
    }
    for (size_t i = 0; i < g->comms.size(); i++)
        if (g->comms[i]) g->nccl.CommDestroy(g->comms[i]);
    for (size_t i = 0; i < g->ev_forest.size(); i++) {
        cudaSetDevice(g->devices[i]);
        cudaEventDestroy(g->ev_forest[i]);
    }
    if (!g->devices.empty()) cudaSetDevice(g->devices[0]);
    if (g->ev_reset) cudaEventDestroy(g->ev_reset);
    for (int k = 0; k < 4; k++)
        if (g->ev_t[k]) cudaEventDestroy(g->ev_t[k]);
    if (g->shared_counter) cudaFree(g->shared_counter);
    for (rh_ctx *c : g->ctx) rh_ctx_destroy(c);
    delete g;
    return RH_OK;
}

int rh_group_size(const rh_group *g) { return g ? g->n_dev : 0; }

rh_ctx *rh_group_ctx(rh_group *g, int i) { return (g && i >= 0 && i < g->n_dev) ? g->ctx[i] : nullptr; }

const char *rh_group_last_error(const rh_group *g) { return g ? g->err.c_str() : "null group"; }

int rh_group_info(const rh_group *g, int *nccl_version, int *work_stealing) {
    if (!g) return RH_EINVAL;
    if (nccl_version) *nccl_version = g->use_nccl ? g->nccl_version : 0;
    if (work_stealing) *work_stealing = g->steal ? 1 : 0;
    return RH_OK;
}

int rh_group_last_times(const rh_group *g, double *out, int n_out) {
    if (!g || !out) return RH_EINVAL;
    for (int i = 0; i < n_out; i++) out[i] = i < 12 ? g->times[i] : 0.0;
    return RH_OK;
}

int rh_hamming_group_multi(rh_group *g, const uint8_t *hashes, const uint8_t *has_hash, const uint8_t *variants,
                           const uint8_t *n_variants, const uint8_t *low_conf, int64_t n, uint32_t similarity,
                           uint32_t *out_label, uint64_t *out_edge_count) {
    if (!g) return RH_EINVAL;
    g->err.clear();
    const int world = g->n_dev;
    rh_ctx *c0 = g->ctx[0];
    if (n < 0 || n > 0x7FFFFFF0ll || (n > 0 && !hashes) || (n_variants && !variants) || similarity > RH_MAX_SIMILARITY_256) {
        g->set_err("rh_hamming_group_multi: bad arguments (similarity <= 63, scanner.rs:1650-1655)");
        return RH_EINVAL;
    }
    if (out_edge_count) *out_edge_count = 0;
    if (n == 0) return RH_OK;
    const auto t_start = std::chrono::steady_clock::now();

    Input in[5];
    const uint8_t *srcs[5] = {hashes, has_hash, variants, n_variants, low_conf};
    const size_t sizes[5] = {(size_t)n * 32, (size_t)n, (size_t)n * 256, (size_t)n, (size_t)n};
    for (int k = 0; k < 5; k++) {
        in[k].src = srcs[k];
        in[k].bytes = sizes[k];
        in[k].slot = S_IN0 + k;
        in[k].owner = srcs[k] ? owner_of(g, srcs[k]) : -1;
        if (in[k].owner == -2) {
            g->set_err("rh_hamming_group_multi: a buffer lives on a GPU outside the group");
            return RH_EINVAL;
        }
    }
    const int out_owner = out_label ? owner_of(g, out_label) : -1;
    if (out_owner > 0 || out_owner == -2) {
        g->set_err("rh_hamming_group_multi: out_label must be host memory or memory of the group's first GPU");
        return RH_EINVAL;
    }
    // the claim counter is reset before any GPU can reach its tile kernel
    if (cudaSetDevice(g->devices[0]) != cudaSuccess || cudaEventRecord(g->ev_t[0], c0->stream) != cudaSuccess ||
        cudaMemsetAsync(g->shared_counter, 0, 64, c0->stream) != cudaSuccess ||
        cudaEventRecord(g->ev_reset, c0->stream) != cudaSuccess) {
        g->set_err("rh_hamming_group_multi: resetting the claim counter failed");
        cudaGetLastError();
        return RH_ECUDA;
    }

    std::vector<uint32_t *> d_forest(world, nullptr);
    std::vector<const unsigned long long *> d_cnt(world, nullptr);
    std::vector<unsigned long long *> d_sum(world, nullptr);
    std::vector<uint32_t *> d_gather(world, nullptr);

    // phase 1 on every GPU: inputs, search, exchange (NCCL) -- all queued, nothing waits for the host
    int rc = g->run_all([&](int i) -> int {
        rh_ctx *ctx = g->ctx[i];
        RH_CUDA(ctx, cudaSetDevice(g->devices[i]));
        cudaStream_t st = ctx->stream;
        const uint8_t *d_in[5];
        if (g->use_nccl) RH_NCCL(g, ctx, g->nccl.GroupStart());
        for (int k = 0; k < 5; k++) {
            int s = replicate_input(g, i, in[k], &d_in[k]);
            if (s != RH_OK) {
                if (g->use_nccl) g->nccl.GroupEnd();
                g->set_err(ctx->err);
                return s;
            }
        }
        if (g->use_nccl) RH_NCCL(g, ctx, g->nccl.GroupEnd());
        if (i == 0) RH_CUDA(ctx, cudaEventRecord(g->ev_t[1], st));
        void *p;
        RH_TRY(scratch(ctx, S_OUT0, (size_t)n * 4, &p));
        d_forest[i] = (uint32_t *)p;
        RH_TRY(scratch(ctx, S_OUT3, 8 * (size_t)(world + 2), &p));
        d_sum[i] = (unsigned long long *)p;
        if (i == 0 || g->use_nccl) {
            RH_TRY(scratch(ctx, S_OUT1, (size_t)n * 4 * world, &p));
            d_gather[i] = (uint32_t *)p;
        }
        HammingPlan plan;
        plan.rank = i;
        plan.world = world;
        plan.shared_counter = (world > 1 && g->steal) ? g->shared_counter : nullptr;
        plan.before_tiles = g->ev_reset;
        int s = hamming_group_enqueue(ctx, d_in[0], d_in[1], d_in[2], d_in[3], d_in[4], n, similarity, plan, d_forest[i],
                                      &d_cnt[i]);
        if (s != RH_OK) {
            g->set_err(ctx->err);
            return s;
        }
        if (world > 1 && g->use_nccl) {
            RH_NCCL(g, ctx, g->nccl.GroupStart());
            RH_NCCL(g, ctx, g->nccl.AllGather(d_forest[i], d_gather[i], (size_t)n, ncclUint32, g->comms[i], st));
            RH_NCCL(g, ctx, g->nccl.AllReduce(d_cnt[i], d_sum[i], 1, ncclUint64, ncclSum, g->comms[i], st));
            RH_NCCL(g, ctx, g->nccl.GroupEnd());
        } else if (world > 1) {
            RH_CUDA(ctx, cudaEventRecord(g->ev_forest[i], st));
        }
        if (i == 0) RH_CUDA(ctx, cudaEventRecord(g->ev_t[2], st));
        return RH_OK;
    });
    if (rc != RH_OK) {
        for (int i = 0; i < world; i++) {
            cudaSetDevice(g->devices[i]);
            cudaStreamSynchronize(g->ctx[i]->stream);
        }
        cudaGetLastError();
        return rc;
    }

    // phase 2 on GPU 0: merge the forests, labels to the caller
    RH_CUDA(c0, cudaSetDevice(g->devices[0]));
    cudaStream_t st0 = c0->stream;
    std::vector<unsigned long long> peer_counts(world, 0ull);
    OutBuf<uint32_t> lab;
    int s = lab.prepare(c0, out_label, (size_t)n, S_OUT2);
    if (s == RH_OK && world > 1 && !g->use_nccl) {
        // peer-copy exchange: GPU 0 pulls every forest over NVLink once it is ready
        for (int i = 0; i < world && s == RH_OK; i++) {
            if (cudaStreamWaitEvent(st0, g->ev_forest[i], 0) != cudaSuccess ||
                cudaMemcpyPeerAsync(d_gather[0] + (size_t)i * n, g->devices[0], d_forest[i], g->devices[i], (size_t)n * 4,
                                    st0) != cudaSuccess ||
                cudaMemcpyPeerAsync(d_sum[0] + 1 + i, g->devices[0], d_cnt[i], g->devices[i], 8, st0) != cudaSuccess)
                s = fail(c0, RH_ECUDA, "peer copy of a forest", cudaGetLastError());
        }
    }
    if (s == RH_OK && lab.dev) {
        if (world > 1)
            s = uf_merge_enqueue(c0, d_gather[0], world, n, lab.dev);
        else if (cudaMemcpyAsync(lab.dev, d_forest[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, st0) != cudaSuccess)
            s = fail(c0, RH_ECUDA, "copy of the labels", cudaGetLastError());
    }
    if (s == RH_OK) s = lab.finish(c0);
    unsigned long long total = 0;
    if (s == RH_OK) {
        cudaError_t e;
        if (world == 1)
            e = cudaMemcpyAsync(&total, d_cnt[0], 8, cudaMemcpyDeviceToHost, st0);
        else if (g->use_nccl)
            e = cudaMemcpyAsync(&total, d_sum[0], 8, cudaMemcpyDeviceToHost, st0);
        else
            e = cudaMemcpyAsync(peer_counts.data(), d_sum[0] + 1, 8 * (size_t)world, cudaMemcpyDeviceToHost, st0);
        if (e != cudaSuccess) s = fail(c0, RH_ECUDA, "copy of the edge count", e);
    }
    cudaEventRecord(g->ev_t[3], st0);
    // the single synchronisation: GPU 0 last (its stream carries the merge)
    for (int i = world - 1; i >= 0; i--) {
        cudaSetDevice(g->devices[i]);
        cudaError_t e = cudaStreamSynchronize(g->ctx[i]->stream);
        if (e != cudaSuccess && s == RH_OK) s = fail(g->ctx[i], RH_ECUDA, "synchronising a GPU of the group", e);
    }
    if (s != RH_OK) {
        g->set_err(c0->err);
        cudaGetLastError();
        return s;
    }
    if (world > 1 && !g->use_nccl)
        for (int i = 0; i < world; i++) total += peer_counts[i];
    if (out_edge_count) *out_edge_count = total;
    double tmax = 0.0, tmin = 1e30, tsum = 0.0;
    for (int i = 0; i < world; i++) {
        cudaSetDevice(g->devices[i]);
        finish_timing(g->ctx[i]);
        const double ms = g->ctx[i]->last_ms;
        tmax = std::max(tmax, ms);
        tmin = std::min(tmin, ms);
        tsum += ms;
    }
    g->times[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    g->times[1] = tmax;   // tile kernel, slowest GPU
    g->times[2] = tmin;   // tile kernel, fastest GPU
    g->times[3] = tsum;   // GPU-milliseconds spent in the tile kernels
    // GPU 0's timeline: inputs | dense arrays | tiles | exchange | merge + copy-out
    cudaSetDevice(g->devices[0]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g->ev_t[0], g->ev_t[1]) == cudaSuccess) g->times[5] = ms;
    if (cudaEventElapsedTime(&ms, g->ev_t[1], c0->ev_a) == cudaSuccess) g->times[6] = ms;
    if (cudaEventElapsedTime(&ms, c0->ev_b, g->ev_t[2]) == cudaSuccess) g->times[7] = ms;
    if (cudaEventElapsedTime(&ms, g->ev_t[2], g->ev_t[3]) == cudaSuccess) g->times[8] = ms;
    if (cudaEventElapsedTime(&ms, g->ev_t[0], g->ev_t[3]) == cudaSuccess) g->times[9] = ms;
    cudaGetLastError();
    return RH_OK;
}

int rh_pdq_hash_batch_multi(rh_group *g, const uint8_t *pixels, int layout, int64_t n, int w, int h, size_t row_pitch,
                            size_t img_pitch, uint8_t *out_hash, float *out_quality, float *out_coeffs,
                            uint8_t *out_dihedral, uint8_t *out_valid) {
    if (!g) return RH_EINVAL;
    g->err.clear();
    if (n < 0 || w <= 0 || h <= 0 || (n > 0 && !pixels)) {
        g->set_err("rh_pdq_hash_batch_multi: bad arguments");
        return RH_EINVAL;
    }
    const int world = g->n_dev;
    if (world > 1) {
        const void *ptrs[6] = {pixels, out_hash, out_quality, out_coeffs, out_dihedral, out_valid};
        for (const void *p : ptrs)
            if (p && owner_of(g, p) != -1) {
                g->set_err("rh_pdq_hash_batch_multi: buffers must be host memory (each GPU hashes one slice of the batch)");
                return RH_EINVAL;
            }
    }
    const int ch = layout == RH_LAYOUT_RGB8 ? 3 : (layout == RH_LAYOUT_RGBA8 ? 4 : 1);
    const size_t rp = row_pitch ? row_pitch : (size_t)w * ch;
    const size_t ip = img_pitch ? img_pitch : rp * (size_t)h;
    const auto t_start = std::chrono::steady_clock::now();
    int rc = g->run_all([&](int i) -> int {
        const int64_t lo = n * i / world, hi = n * (i + 1) / world;
        if (hi <= lo) return RH_OK;
        int s = rh_pdq_hash_batch(g->ctx[i], pixels + (size_t)lo * ip, layout, hi - lo, w, h, row_pitch, img_pitch,
                                  out_hash ? out_hash + (size_t)lo * 32 : nullptr, out_quality ? out_quality + lo : nullptr,
                                  out_coeffs ? out_coeffs + (size_t)lo * 256 : nullptr,
                                  out_dihedral ? out_dihedral + (size_t)lo * 256 : nullptr, out_valid ? out_valid + lo : nullptr);
        if (s != RH_OK) g->set_err(g->ctx[i]->err);
        return s;
    });
    g->times[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    return rc;
}

}  // extern "C"
