// C-ABI of the engine (include/arvc_icp.h): context, scan store, batch orchestration.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/arvc_icp.h"
#include "engine.cuh"

using namespace arvc;

namespace {

std::string g_create_error;

struct Scan {
    int64_t id = 0;
    int n_raw = 0;
    bool f64 = false;
    void* d_raw = nullptr;
    cudaEvent_t up_ev = nullptr;   // recorded on the copy stream after the upload; the compute stream waits for it once
    bool up_pending = false;
    // persistent slab
    void* slab = nullptr;
    size_t slab_bytes = 0;
    ScanDev dev{};              // host copy (persistent pointers only)
    ScanDev* d_dev = nullptr;   // device copy used by the ICP kernels
    bool preprocessed = false;
    bool has_normals = false;
    bool voxel_on = false;
    arvc_preprocess_params params{};
};

struct PendingBatch {
    int n_pairs = 0;
    void* slab = nullptr;
    size_t slab_bytes = 0;
    PairState* d_states = nullptr;
    PairState* h_states = nullptr;   // pinned, from the context's pool: state heads, only read back for ARVC_DEBUG_STATS
    size_t h_bytes = 0;
    char* h_stage = nullptr;         // pinned: [initial guesses | result records | status word]
    size_t h_stage_bytes = 0;
    arvc_result_record* h_records = nullptr;   // inside h_stage
    int* h_status = nullptr;                   // inside h_stage
    void* d_records = nullptr;       // device copy of the records (what the multi-GPU gather reads), inside the slab
    cudaEvent_t done_ev = nullptr;   // recorded after the batch's last device->host copy
    const IcpGraph* graph = nullptr; // the device-terminated loop this batch was enqueued with (launch accounting)
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct SlabPlanner {   // two-pass bump allocator: plan sizes, then hand out pointers
    size_t off = 0;
    char* base = nullptr;
    template <typename T> T* take(size_t count) {
        const size_t o = off;
        off += align_up(count * sizeof(T));
        return base ? reinterpret_cast<T*>(base + o) : nullptr;
    }
};

}  // namespace

struct arvc_ctx {
    int device = 0;
    Launcher L;
    cudaStream_t copy_stream = nullptr;   // host -> device uploads, so that they overlap the kernels of earlier scans
    // Look-ahead preprocessing (arvc_scan_preprocess_ahead): a second compute stream with a scratch block of its own, so
    // that the next scan of a sequential caller is preprocessed while the registration enqueued before it still runs.
    // `ahead_ev` marks the end of the last look-ahead job; the main stream waits for it at the next entry point.
    cudaStream_t ahead_stream = nullptr;
    cudaEvent_t ahead_ev = nullptr;
    bool ahead_pending = false;
    void* ahead_scratch = nullptr;
    size_t ahead_scratch_bytes = 0;
    int enter() {                         // every entry point but the look-ahead one: select the device, join the side stream
        const cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return (int)e;
        if (ahead_pending) { cudaStreamWaitEvent(L.stream, ahead_ev, 0); ahead_pending = false; }
        return 0;
    }
    std::string error;
    std::unordered_map<int64_t, std::unique_ptr<Scan>> scans;
    std::map<uint64_t, PendingBatch> pending;
    uint64_t next_ticket = 1;
    std::vector<std::pair<size_t, void*>> pinned_free;   // recycled pinned staging buffers
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t event_get() {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        return e;
    }
    IcpGraphCache icp_graphs;
    bool normals_tap = false;           // arvc_ctx_set_option("normals_tap", 1): record the neighbour sets of the normals
    bool icp_use_graph = true;          // ARVC_ICP_LOOP=unrolled / arvc_ctx_set_option("icp_loop_graph", 0) switch it off

    void* pinned_get(size_t bytes, size_t* got) {
        for (size_t i = 0; i < pinned_free.size(); ++i)
            if (pinned_free[i].first >= bytes) {
                void* p = pinned_free[i].second;
                *got = pinned_free[i].first;
                pinned_free.erase(pinned_free.begin() + i);
                return p;
            }
        size_t cap = 1 << 16;
        while (cap < bytes) cap <<= 1;
        void* p = nullptr;
        if (cudaMallocHost(&p, cap) != cudaSuccess) return nullptr;
        *got = cap;
        return p;
    }
    void pinned_put(void* p, size_t bytes) { if (p) pinned_free.emplace_back(bytes, p); }

    // Large transient device blocks (preprocessing scratch, per-batch ICP state) are recycled here instead of going
    // back to the stream-ordered pool: a batch-sized block that the pool has split in the meantime would have to be
    // re-created from the OS (hundreds of ms).  All use is on the context's single stream, so reuse is stream ordered.
    std::vector<std::pair<size_t, void*>> dev_free;
    void* dev_get(size_t bytes, size_t* got) {
        int best = -1;
        for (size_t i = 0; i < dev_free.size(); ++i)
            if (dev_free[i].first >= bytes && (best < 0 || dev_free[i].first < dev_free[best].first)) best = (int)i;
        if (best >= 0 && dev_free[best].first <= 2 * bytes + (1u << 20)) {
            void* p = dev_free[best].second;
            *got = dev_free[best].first;
            dev_free.erase(dev_free.begin() + best);
            return p;
        }
        const size_t cap = bytes + bytes / 8 + 256;      // head-room so that slightly larger batches still fit
        void* p = nullptr;
        if (cudaMallocAsync(&p, cap, L.stream) != cudaSuccess) return nullptr;
        *got = cap;
        return p;
    }
    void dev_put(void* p, size_t bytes) {
        if (!p) return;
        dev_free.emplace_back(bytes, p);
        while (dev_free.size() > 12) {                   // keep the cache small: drop the smallest block
            size_t k = 0;
            for (size_t i = 1; i < dev_free.size(); ++i) if (dev_free[i].first < dev_free[k].first) k = i;
            cudaFreeAsync(dev_free[k].second, L.stream);
            dev_free.erase(dev_free.begin() + k);
        }
    }

    int fail(int code, const std::string& msg) { error = msg; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
        error = buf;
        return ARVC_E_CUDA;
    }
    Scan* find(int64_t id) {
        auto it = scans.find(id);
        return it == scans.end() ? nullptr : it->second.get();
    }
};

#define CK(call)                                                      \
    do {                                                              \
        const cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return ctx->cuda_fail(e__, #call);    \
    } while (0)

namespace {

// the compute stream may touch (or free) the raw cloud only after its upload on the copy stream has finished
void await_upload(arvc_ctx* ctx, Scan* s) {
    if (s->up_pending) cudaStreamWaitEvent(ctx->L.stream, s->up_ev, 0);
    s->up_pending = false;
}

void release_scan(arvc_ctx* ctx, Scan* s) {
    cudaStream_t st = ctx->L.stream;
    await_upload(ctx, s);
    if (s->up_ev) { cudaEventDestroy(s->up_ev); s->up_ev = nullptr; }
    if (s->d_raw) cudaFreeAsync(s->d_raw, st);
    if (s->slab) cudaFreeAsync(s->slab, st);
    if (s->d_dev) cudaFreeAsync(s->d_dev, st);
    s->d_raw = s->slab = nullptr;
    s->d_dev = nullptr;
}

bool same_params(const arvc_preprocess_params& a, const arvc_preprocess_params& b) {
    auto eq = [](double x, double y) { return x == y || (x != x && y != y); };
    const bool va = a.voxel_size > 0, vb = b.voxel_size > 0;
    return eq(a.min_radius2, b.min_radius2) && eq(a.max_radius2, b.max_radius2) && eq(a.min_height, b.min_height) &&
           eq(a.max_height, b.max_height) && va == vb && (!va || eq(a.voxel_size, b.voxel_size)) &&
           eq(a.grid_cell, b.grid_cell) && eq(a.grid_max_dist, b.grid_max_dist);
}

int bits_for(double cells) {
    int b = 1;
    while (b < 31 && (double)(1ll << b) < cells) ++b;
    return b;
}

int upload(arvc_ctx* ctx, int64_t id, const void* xyz, int n, bool f64) {
    if (!ctx) return ARVC_E_ARG;
    if (n < 0 || (n > 0 && !xyz)) return ctx->fail(ARVC_E_ARG, "scan_upload: bad pointer/size");
    CK((cudaError_t)ctx->enter());
    Scan* old = ctx->find(id);
    if (old) { release_scan(ctx, old); ctx->scans.erase(id); }
    auto s = std::make_unique<Scan>();
    s->id = id; s->n_raw = n; s->f64 = f64;
    const size_t bytes = (size_t)n * 3 * (f64 ? sizeof(double) : sizeof(float));
    if (n > 0) {
        CK(cudaMallocAsync(&s->d_raw, bytes, ctx->copy_stream));
        CK(cudaMemcpyAsync(s->d_raw, xyz, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventCreateWithFlags(&s->up_ev, cudaEventDisableTiming));
        CK(cudaEventRecord(s->up_ev, ctx->copy_stream));
        s->up_pending = true;
    }
    ctx->scans[id] = std::move(s);
    return ARVC_OK;
}

// device buffers that outlive preprocessing
void plan_persistent(SlabPlanner& P, Scan& s, bool wide, bool voxel_on, unsigned table_cap, int tap_nn) {
    const size_t cap = (size_t)std::max(s.n_raw, 1);
    ScanDev& d = s.dev;
    d.tap_cnt = tap_nn > 0 ? P.take<int>(cap) : nullptr;
    d.tap_idx = tap_nn > 0 ? P.take<int>(cap * (size_t)tap_nn) : nullptr;
    d.tap_stride = tap_nn;
    d.counts = P.take<int>(CNT_WORDS);
    d.recs = wide ? (void*)P.take<RecD>(cap) : (void*)P.take<RecF>(cap);
    d.normals = P.take<double>(cap * 4);
    d.nn_count = P.take<int>(cap);
    d.table = P.take<HashEntry>(table_cap);
    d.raw_index = P.take<int>(cap);
    d.vox_keys = voxel_on ? P.take<int>(cap * 3) : nullptr;
    d.vox_counts = voxel_on ? P.take<int>(cap) : nullptr;
}

void plan_scratch(SlabPlanner& P, ScanDev& d, int n_raw, bool voxel_on) {
    const size_t cap = (size_t)std::max(n_raw, 1);
    d.fx = P.take<double>(cap); d.fy = P.take<double>(cap); d.fz = P.take<double>(cap);
    if (voxel_on) {
        d.vx = P.take<double>(cap); d.vy = P.take<double>(cap); d.vz = P.take<double>(cap);
        d.key64[0] = P.take<unsigned long long>(cap); d.key64[1] = P.take<unsigned long long>(cap);
    } else {
        d.vx = d.vy = d.vz = nullptr;
        d.key64[0] = d.key64[1] = nullptr;
    }
    d.key32[0] = P.take<unsigned>(cap); d.key32[1] = P.take<unsigned>(cap);
    d.val[0] = P.take<int>(cap); d.val[1] = P.take<int>(cap);
    d.hist = P.take<int>(256 * ((cap + 2047) / 2048));
    d.blk = P.take<int>((cap + 1023) / 1024 + 8);
    d.bbox = P.take<double>(8);
    d.moments = P.take<double>(cap * 10);
    d.redo_list = P.take<int>(cap);
    d.fb_list = P.take<int>(cap);
}

__global__ void k_clear_scan(const ScanDev* __restrict__ scans) {
    const ScanDev& s = scans[blockIdx.y];
    const unsigned cap = s.table_mask + 1u;
    uint4* t = reinterpret_cast<uint4*>(s.table);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) t[i] = make_uint4(0, 0, 0, 0);
    if (blockIdx.x == 0) {
        if (threadIdx.x < CNT_WORDS) s.counts[threadIdx.x] = 0;
        if (threadIdx.x < 8) reinterpret_cast<unsigned long long*>(s.bbox)[threadIdx.x] = 0xffffffffffffffffULL;
    }
}

}  // namespace

extern "C" {

int arvc_version(void) { return 100; }

int arvc_ctx_create(int device, arvc_ctx** out) {
    if (!out) return ARVC_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this engine has no CPU fallback";
        return ARVC_E_CUDA;
    }
    if (device < 0 || device >= count) { g_create_error = "bad device index"; return ARVC_E_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return ARVC_E_CUDA; }
    auto ctx = new arvc_ctx();
    ctx->device = device;
    // the look-ahead stream yields to the context stream: blocks of a registration in flight (short, latency-bound
    // kernels) are scheduled before the blocks of the next scan's preprocessing, which only fills the gaps
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (getenv("ARVC_STREAM_PRIO") && atoi(getenv("ARVC_STREAM_PRIO")) == 0) prio_greatest = prio_least = 0;      // A/B knob
    e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->L.stream, cudaStreamNonBlocking, prio_greatest);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->ahead_stream, cudaStreamNonBlocking, prio_least);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ahead_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return ARVC_E_CUDA; }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    const char* loop = getenv("ARVC_ICP_LOOP");
    if (loop && std::string(loop) == "unrolled") ctx->icp_use_graph = false;
    *out = ctx;
    return ARVC_OK;
}

int arvc_ctx_set_option(arvc_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return ARVC_E_ARG;
    if (std::string(name) == "icp_loop_graph") { ctx->icp_use_graph = value != 0; return ARVC_OK; }
    if (std::string(name) == "normals_tap") { ctx->normals_tap = value != 0; return ARVC_OK; }
    return ctx->fail(ARVC_E_ARG, std::string("ctx_set_option: unknown option ") + name);
}

void arvc_ctx_destroy(arvc_ctx* ctx) {
    if (!ctx) return;
    ctx->enter();
    cudaStreamSynchronize(ctx->ahead_stream);
    cudaStreamSynchronize(ctx->L.stream);
    if (ctx->ahead_scratch) cudaFreeAsync(ctx->ahead_scratch, ctx->L.stream);
    icp_graphs_destroy(ctx->icp_graphs);
    for (auto& df : ctx->dev_free) cudaFreeAsync(df.second, ctx->L.stream);
    for (auto& kv : ctx->pending) {
        if (kv.second.slab) cudaFreeAsync(kv.second.slab, ctx->L.stream);
        if (kv.second.h_states) cudaFreeHost(kv.second.h_states);
    }
    for (auto& kv : ctx->pending) {
        if (kv.second.h_stage) cudaFreeHost(kv.second.h_stage);
        if (kv.second.done_ev) cudaEventDestroy(kv.second.done_ev);
    }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    for (auto& pf : ctx->pinned_free) cudaFreeHost(pf.second);
    for (auto& kv : ctx->scans) release_scan(ctx, kv.second.get());
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->L.stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->ahead_stream);
    cudaStreamDestroy(ctx->L.stream);
    if (ctx->ahead_ev) cudaEventDestroy(ctx->ahead_ev);
    delete ctx;
}

const char* arvc_last_error(const arvc_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int arvc_sync(arvc_ctx* ctx) {
    if (!ctx) return ARVC_E_ARG;
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->ahead_stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    if (ctx->L.err != cudaSuccess) return ctx->cuda_fail(ctx->L.err, "kernel launch");
    return ARVC_OK;
}

void* arvc_stream(arvc_ctx* ctx) { return ctx ? (void*)ctx->L.stream : nullptr; }
int64_t arvc_kernel_launches(const arvc_ctx* ctx) { return ctx ? ctx->L.launches : 0; }

int arvc_profile_enable(arvc_ctx* ctx, int on) {
    if (!ctx) return ARVC_E_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->L.stream);
    for (auto& r : ctx->L.recs) { ctx->L.pool.push_back(r.a); ctx->L.pool.push_back(r.b); }
    ctx->L.recs.clear();
    ctx->L.profile = on != 0;
    return ARVC_OK;
}

int arvc_profile_report(arvc_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return ARVC_E_ARG;
    CK((cudaError_t)ctx->enter());
    CK(cudaStreamSynchronize(ctx->L.stream));
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : ctx->L.recs) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto& a = agg[r.name]; a.first += 1; a.second += ms; }
        ctx->L.pool.push_back(r.a); ctx->L.pool.push_back(r.b);
    }
    ctx->L.recs.clear();
    std::string out;
    char line[160];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s,%lld,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (out.size() + 1 > cap) return ctx->fail(ARVC_E_ARG, "profile_report: buffer too small");
    std::memcpy(buf, out.c_str(), out.size() + 1);
    return ARVC_OK;
}

int arvc_scan_invalidate(arvc_ctx* ctx, int64_t scan_id) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s) return ctx->fail(ARVC_E_STATE, "scan_invalidate: unknown scan id");
    s->preprocessed = false;
    s->has_normals = false;
    return ARVC_OK;
}

int arvc_scan_upload_f32(arvc_ctx* ctx, int64_t scan_id, const float* xyz, int n) { return upload(ctx, scan_id, xyz, n, false); }
int arvc_scan_upload_f64(arvc_ctx* ctx, int64_t scan_id, const double* xyz, int n) { return upload(ctx, scan_id, xyz, n, true); }

int arvc_scan_wait_upload(arvc_ctx* ctx, int64_t scan_id) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s) return ctx->fail(ARVC_E_STATE, "scan_wait_upload: unknown scan id");
    if (s->up_ev) CK(cudaEventSynchronize(s->up_ev));
    return ARVC_OK;
}

int arvc_scan_free(arvc_ctx* ctx, int64_t scan_id) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s) return ARVC_OK;
    ctx->enter();
    release_scan(ctx, s);
    ctx->scans.erase(scan_id);
    return ARVC_OK;
}

namespace {
// arvc_scan_preprocess (ahead = false) and arvc_scan_preprocess_ahead (ahead = true: same work on the look-ahead stream)
int preprocess_impl(arvc_ctx* ctx, int n_scans, const int64_t* scan_ids, const arvc_preprocess_params* p, bool ahead) {
    if (!ctx) return ARVC_E_ARG;
    if (n_scans < 0 || (n_scans > 0 && !scan_ids) || !p) return ctx->fail(ARVC_E_ARG, "scan_preprocess: bad arguments");
    if (ahead) {
        // only scans without device state of their own may go ahead: nothing on the main stream can be using them
        for (int i = 0; i < n_scans; ++i) {
            const Scan* s = ctx->find(scan_ids[i]);
            if (s && (s->slab || s->d_dev)) ahead = false;
        }
    }
    if (ahead) CK(cudaSetDevice(ctx->device)); else CK((cudaError_t)ctx->enter());
    // every launch, copy and allocation below goes through ctx->L.stream: the look-ahead variant swaps the stream in
    struct StreamSwap {
        arvc_ctx* c; cudaStream_t saved; bool on;
        StreamSwap(arvc_ctx* c_, bool on_) : c(c_), saved(c_->L.stream), on(on_) { if (on) c->L.stream = c->ahead_stream; }
        ~StreamSwap() { if (on) c->L.stream = saved; }
    } swap_guard(ctx, ahead);
    const bool voxel_on = p->voxel_size > 0.0;          // NaN and <= 0 mean "voxel_size is None"
    const bool want_normals = p->want_normals != 0;
    if (!(p->max_radius2 > 0) || !std::isfinite(p->max_radius2) || !std::isfinite(p->min_height) || !std::isfinite(p->max_height) ||
        !(p->max_height > p->min_height))
        return ctx->fail(ARVC_E_ARG, "scan_preprocess: the radius/height filter must be bounded (finite max_radius, min/max_height)");
    if (want_normals && (!(p->normal_radius > 0) || p->max_nn < 1)) return ctx->fail(ARVC_E_ARG, "scan_preprocess: bad normal parameters");

    // ---- grid specification from the filter bounds (data independent, so no device->host round trip)
    const double R = std::sqrt(p->max_radius2);
    const double pad = 0.5;
    const double ext = std::max(2 * (R + pad), p->max_height - p->min_height + 2 * pad);
    double c0 = p->grid_cell > 0 ? p->grid_cell : (getenv("ARVC_GRID_CELL") ? atof(getenv("ARVC_GRID_CELL")) : 0.09);
    const double ncell_max = (double)((1 << kMortonBits) - 1);
    if (ext / c0 > ncell_max) c0 = ext / ncell_max;
    const double max_dist = p->grid_max_dist > 0 ? p->grid_max_dist : 10.0;
    GridSpec g{};
    g.ox = -(R + pad); g.oy = -(R + pad); g.oz = p->min_height - pad;
    g.c0 = c0; g.inv_c0 = 1.0 / c0;
    g.top_level = 0;
    while (g.top_level < kMortonBits && c0 * (double)(1 << g.top_level) < max_dist * 1.001) ++g.top_level;
    NormalParams np{};
    np.radius = p->normal_radius; np.max_nn = p->max_nn; np.level = 0;
    if (want_normals)
        while (np.level < kMortonBits && c0 * (double)(1 << np.level) < p->normal_radius * (1.0 - 1e-12)) ++np.level;

    VoxelParams vp{};
    if (voxel_on) {
        vp.voxel = p->voxel_size;
        vp.bx = bits_for(2 * R / p->voxel_size + 3);
        vp.by = vp.bx;
        vp.bz = bits_for((p->max_height - p->min_height) / p->voxel_size + 3);
        if (vp.bx + vp.by + vp.bz > 63) return ctx->fail(ARVC_E_ARG, "scan_preprocess: voxel_size too small for the filter extent");
    }
    FilterParams fp{p->min_radius2, p->max_radius2, p->min_height, p->max_height};

    // ---- collect the scans that actually need work
    std::vector<Scan*> todo, normals_only;
    for (int i = 0; i < n_scans; ++i) {
        Scan* s = ctx->find(scan_ids[i]);
        if (!s) return ctx->fail(ARVC_E_STATE, "scan_preprocess: unknown scan id (upload first)");
        bool dup = false;
        for (Scan* t : todo) dup |= (t == s);
        if (dup) continue;
        if (s->preprocessed && same_params(s->params, *p)) {
            if (!want_normals || (s->has_normals && s->params.normal_radius == p->normal_radius && s->params.max_nn == p->max_nn)) continue;
        }
        todo.push_back(s);
    }
    if (todo.empty()) return ARVC_OK;
    // from here on a failure leaves these scans without usable device state: they count as not preprocessed until the
    // publish loop at the end marks them again (an early return must never leave `preprocessed` set over freed buffers)
    for (Scan* s : todo) { s->preprocessed = false; s->has_normals = false; }
    for (Scan* s : todo) await_upload(ctx, s);

    // ---- allocate: persistent slab per scan, one scratch slab for the batch
    int cap_max = 0;
    bool any_wide = false, any_narrow = false;
    std::vector<ScanDev> h_batch(todo.size());
    SlabPlanner scratch_plan;
    for (Scan* s : todo) {
        const bool wide = s->f64 || voxel_on;
        unsigned tcap = 1024;
        const unsigned tfac = getenv("ARVC_TABLE_FACTOR") ? (unsigned)atoi(getenv("ARVC_TABLE_FACTOR")) : (c0 < 0.12 ? 8u : 4u);
        while (tcap < tfac * (unsigned)std::max(s->n_raw, 1)) tcap <<= 1;      // finer grids occupy more cells
        if (s->slab) { cudaFreeAsync(s->slab, ctx->L.stream); s->slab = nullptr; }
        const int tap_nn = (ctx->normals_tap && want_normals) ? p->max_nn : 0;
        SlabPlanner pp;
        plan_persistent(pp, *s, wide, voxel_on, tcap, tap_nn);
        s->slab_bytes = pp.off;
        CK(cudaMallocAsync(&s->slab, s->slab_bytes, ctx->L.stream));
        SlabPlanner pq;
        pq.base = reinterpret_cast<char*>(s->slab);
        s->dev = ScanDev{};
        plan_persistent(pq, *s, wide, voxel_on, tcap, tap_nn);
        if (tap_nn > 0) CK(cudaMemsetAsync(s->dev.tap_cnt, 0, sizeof(int) * (size_t)std::max(s->n_raw, 1), ctx->L.stream));
        s->dev.raw = s->d_raw; s->dev.n_raw = s->n_raw; s->dev.raw_f64 = s->f64 ? 1 : 0; s->dev.cap = std::max(s->n_raw, 1);
        s->dev.wide = wide ? 1 : 0; s->dev.table_mask = tcap - 1; s->dev.grid = g;
        ScanDev tmp = s->dev;
        plan_scratch(scratch_plan, tmp, s->n_raw, voxel_on);
        cap_max = std::max(cap_max, s->n_raw);
        any_wide |= wide; any_narrow |= !wide;
    }
    const size_t scratch_bytes = scratch_plan.off + align_up(sizeof(ScanDev) * todo.size());
    size_t scratch_cap = 0;
    void* scratch = nullptr;
    if (ahead) {      // the side stream's own block: the recycled blocks of dev_get are ordered by the main stream only
        if (ctx->ahead_scratch_bytes < scratch_bytes) {
            if (ctx->ahead_scratch) cudaFreeAsync(ctx->ahead_scratch, ctx->L.stream);
            ctx->ahead_scratch = nullptr;
            ctx->ahead_scratch_bytes = 0;
            const size_t cap = scratch_bytes + scratch_bytes / 4;
            if (cudaMallocAsync(&ctx->ahead_scratch, cap, ctx->L.stream) == cudaSuccess) ctx->ahead_scratch_bytes = cap;
        }
        scratch = ctx->ahead_scratch;
    } else {
        scratch = ctx->dev_get(scratch_bytes, &scratch_cap);
    }
    if (!scratch) return ctx->fail(ARVC_E_NOMEM, "scan_preprocess: device allocation failed");
    SlabPlanner sp;
    sp.base = reinterpret_cast<char*>(scratch);
    for (size_t i = 0; i < todo.size(); ++i) {
        h_batch[i] = todo[i]->dev;
        plan_scratch(sp, h_batch[i], todo[i]->n_raw, voxel_on);
    }
    ScanDev* d_batch = sp.take<ScanDev>(todo.size());
    CK(cudaMemcpyAsync(d_batch, h_batch.data(), sizeof(ScanDev) * todo.size(), cudaMemcpyHostToDevice, ctx->L.stream));

    // ---- run
    unsigned tmax = 0;
    for (Scan* s : todo) tmax = std::max(tmax, s->dev.table_mask + 1u);
    ctx->L.launch("clear_scan", k_clear_scan, dim3(std::min(256u, (tmax + 255u) / 256u), (unsigned)todo.size()), dim3(256), (const ScanDev*)d_batch);
    run_preprocess(ctx->L, d_batch, (int)todo.size(), cap_max, fp, vp, voxel_on);
    if (want_normals) run_normals(ctx->L, d_batch, (int)todo.size(), cap_max, np, any_wide, any_narrow, ctx->normals_tap);

    // ---- publish the persistent device descriptors used by the ICP kernels
    for (Scan* s : todo) {
        if (!s->d_dev) CK(cudaMallocAsync(&s->d_dev, sizeof(ScanDev), ctx->L.stream));
        CK(cudaMemcpyAsync(s->d_dev, &s->dev, sizeof(ScanDev), cudaMemcpyHostToDevice, ctx->L.stream));
        s->preprocessed = true;
        s->has_normals = want_normals;
        s->voxel_on = voxel_on;
        s->params = *p;
    }
    if (ahead) {
        CK(cudaEventRecord(ctx->ahead_ev, ctx->L.stream));
        ctx->ahead_pending = true;
    } else {
        ctx->dev_put(scratch, scratch_cap);
    }
    if (ctx->L.err != cudaSuccess) return ctx->cuda_fail(ctx->L.err, "kernel launch");
    return ARVC_OK;
}
}  // namespace

int arvc_scan_preprocess(arvc_ctx* ctx, int n_scans, const int64_t* scan_ids, const arvc_preprocess_params* p) {
    return preprocess_impl(ctx, n_scans, scan_ids, p, false);
}

int arvc_scan_preprocess_ahead(arvc_ctx* ctx, int n_scans, const int64_t* scan_ids, const arvc_preprocess_params* p) {
    return preprocess_impl(ctx, n_scans, scan_ids, p, true);
}

static int fetch_counts(arvc_ctx* ctx, Scan* s, int* counts) {
    CK(cudaMemcpyAsync(counts, s->dev.counts, sizeof(int) * CNT_WORDS, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    if (counts[CNT_ERR] & ERR_HASH_FULL) return ctx->fail(ARVC_E_CAPACITY, "hash grid overflow");
    if (counts[CNT_ERR] & ERR_VOXEL_RANGE) return ctx->fail(ARVC_E_CAPACITY, "voxel index outside the range implied by the filter bounds");
    return ARVC_OK;
}

int arvc_scan_info(arvc_ctx* ctx, int64_t scan_id, int* n_raw, int* n_filtered, int* n_points, int* has_normals) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s) return ctx->fail(ARVC_E_STATE, "scan_info: unknown scan id");
    CK((cudaError_t)ctx->enter());
    if (n_raw) *n_raw = s->n_raw;
    if (has_normals) *has_normals = s->has_normals ? 1 : 0;
    if (!s->preprocessed) {
        if (n_filtered) *n_filtered = -1;
        if (n_points) *n_points = -1;
        return ARVC_OK;
    }
    int c[CNT_WORDS];
    const int rc = fetch_counts(ctx, s, c);
    if (rc) return rc;
    if (n_filtered) *n_filtered = c[CNT_NFILT];
    if (n_points) *n_points = c[CNT_NPTS];
    return ARVC_OK;
}

int arvc_scan_get_points(arvc_ctx* ctx, int64_t scan_id, double* xyz, double* normals) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s || !s->preprocessed) return ctx->fail(ARVC_E_STATE, "scan_get_points: scan not preprocessed");
    if (normals && !s->has_normals) return ctx->fail(ARVC_E_STATE, "scan_get_points: no normals (point-to-point preprocessing)");
    CK((cudaError_t)ctx->enter());
    int c[CNT_WORDS];
    const int rc = fetch_counts(ctx, s, c);
    if (rc) return rc;
    const int n = c[CNT_NPTS];
    if (n == 0) return ARVC_OK;
    std::vector<double> hn;
    std::vector<char> hr((size_t)n * (s->dev.wide ? sizeof(RecD) : sizeof(RecF)));
    CK(cudaMemcpyAsync(hr.data(), s->dev.recs, hr.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    if (normals) {
        hn.resize((size_t)n * 4);
        CK(cudaMemcpyAsync(hn.data(), s->dev.normals, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, ctx->L.stream));
    }
    CK(cudaStreamSynchronize(ctx->L.stream));
    for (int p = 0; p < n; ++p) {
        int idx;
        double x, y, z;
        if (s->dev.wide) { const RecD& r = reinterpret_cast<const RecD*>(hr.data())[p]; x = r.x; y = r.y; z = r.z; idx = r.idx; }
        else { const RecF& r = reinterpret_cast<const RecF*>(hr.data())[p]; x = r.x; y = r.y; z = r.z; idx = r.idx; }
        if (idx < 0 || idx >= n) return ctx->fail(ARVC_E_CUDA, "scan_get_points: corrupt permutation");
        if (xyz) { xyz[3 * (size_t)idx] = x; xyz[3 * (size_t)idx + 1] = y; xyz[3 * (size_t)idx + 2] = z; }
        if (normals) for (int d = 0; d < 3; ++d) normals[3 * (size_t)idx + d] = hn[4 * (size_t)p + d];
    }
    return ARVC_OK;
}

int arvc_map_build(arvc_ctx* ctx, int n_scans, const int64_t* scan_ids, const double* T, const arvc_preprocess_params* p,
                   double* xyz_out, int64_t capacity_points, int64_t* offsets_out) {
    if (!ctx) return ARVC_E_ARG;
    if (n_scans < 0 || (n_scans > 0 && (!scan_ids || !T)) || !p || !offsets_out || capacity_points < 0 || (capacity_points > 0 && !xyz_out))
        return ctx->fail(ARVC_E_ARG, "map_build: bad arguments");
    if (n_scans > 32768) return ctx->fail(ARVC_E_ARG, "map_build: at most 32768 keyframes per call (split the batch)");
    offsets_out[0] = 0;
    if (n_scans == 0) return ARVC_OK;
    CK((cudaError_t)ctx->enter());
    std::vector<int64_t> uniq(scan_ids, scan_ids + n_scans);          // a keyframe may appear more than once in a map
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    int rc = arvc_scan_preprocess(ctx, (int)uniq.size(), uniq.data(), p);
    if (rc) return rc;
    int cap_max = 1;
    size_t h_bytes = 0;
    const size_t ptr_bytes = align_up(sizeof(ScanDev*) * (size_t)n_scans), t_bytes = align_up(sizeof(double) * 16 * (size_t)n_scans),
                 off_bytes = align_up(sizeof(long long) * ((size_t)n_scans + 2));   // + total, + OR of the scans' error flags
    char* h = reinterpret_cast<char*>(ctx->pinned_get(ptr_bytes + t_bytes + off_bytes, &h_bytes));
    if (!h) return ctx->fail(ARVC_E_NOMEM, "map_build: pinned host allocation failed");
    const ScanDev** hp = reinterpret_cast<const ScanDev**>(h);
    for (int i = 0; i < n_scans; ++i) {
        Scan* s = ctx->find(scan_ids[i]);
        hp[i] = s->d_dev;
        cap_max = std::max(cap_max, s->dev.cap);
    }
    std::memcpy(h + ptr_bytes, T, sizeof(double) * 16 * (size_t)n_scans);
    size_t head_cap = 0;
    char* d_head = reinterpret_cast<char*>(ctx->dev_get(ptr_bytes + t_bytes + off_bytes, &head_cap));
    if (!d_head) { ctx->pinned_put(h, h_bytes); return ctx->fail(ARVC_E_NOMEM, "map_build: device allocation failed"); }
    const ScanDev* const* d_scans = reinterpret_cast<const ScanDev* const*>(d_head);
    const double* d_T = reinterpret_cast<const double*>(d_head + ptr_bytes);
    long long* d_off = reinterpret_cast<long long*>(d_head + ptr_bytes + t_bytes);
    long long* h_off = reinterpret_cast<long long*>(h + ptr_bytes + t_bytes);
    auto release = [&]() { ctx->dev_put(d_head, head_cap); ctx->pinned_put(h, h_bytes); };
    cudaError_t e = cudaMemcpyAsync(d_head, h, ptr_bytes + t_bytes, cudaMemcpyHostToDevice, ctx->L.stream);
    run_map_build(ctx->L, d_scans, d_T, n_scans, cap_max, d_off, nullptr, 0, true);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_off, d_off, sizeof(long long) * ((size_t)n_scans + 2), cudaMemcpyDeviceToHost, ctx->L.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->L.stream);
    if (e != cudaSuccess || ctx->L.err != cudaSuccess) { release(); return ctx->cuda_fail(e != cudaSuccess ? e : ctx->L.err, "map_build"); }
    for (int i = 0; i <= n_scans; ++i) offsets_out[i] = (int64_t)h_off[i];
    const long long total = h_off[n_scans], flags = h_off[n_scans + 1];
    if (flags & ERR_HASH_FULL) { release(); return ctx->fail(ARVC_E_CAPACITY, "hash grid overflow"); }
    if (flags & ERR_VOXEL_RANGE) { release(); return ctx->fail(ARVC_E_CAPACITY, "voxel index outside the range implied by the filter bounds"); }
    if (total > capacity_points) { release(); return ctx->fail(ARVC_E_CAPACITY, "map_build: output capacity too small (offsets are valid)"); }
    if (total > 0) {
        size_t out_cap = 0;
        double* d_out = reinterpret_cast<double*>(ctx->dev_get(sizeof(double) * 3 * (size_t)total, &out_cap));
        if (!d_out) { release(); return ctx->fail(ARVC_E_NOMEM, "map_build: device allocation failed"); }
        run_map_build(ctx->L, d_scans, d_T, n_scans, cap_max, d_off, d_out, total, false);
        e = cudaMemcpyAsync(xyz_out, d_out, sizeof(double) * 3 * (size_t)total, cudaMemcpyDeviceToHost, ctx->L.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->L.stream);
        ctx->dev_put(d_out, out_cap);
        if (e != cudaSuccess || ctx->L.err != cudaSuccess) { release(); return ctx->cuda_fail(e != cudaSuccess ? e : ctx->L.err, "map_build"); }
    }
    release();
    return ARVC_OK;
}

int arvc_scan_fit_plane(arvc_ctx* ctx, int64_t scan_id, double max_z, double dist_threshold, int iterations, uint64_t seed,
                        double* plane_out, int32_t* n_inliers) {
    if (!ctx) return ARVC_E_ARG;
    if (!plane_out || iterations < 1 || iterations > 65535 || !(dist_threshold > 0)) return ctx->fail(ARVC_E_ARG, "scan_fit_plane: bad arguments");
    Scan* s = ctx->find(scan_id);
    if (!s || !s->preprocessed) return ctx->fail(ARVC_E_STATE, "scan_fit_plane: scan not preprocessed");
    CK((cudaError_t)ctx->enter());
    const size_t cap = (size_t)std::max(s->dev.cap, 1);
    const size_t orig_b = align_up(sizeof(double) * 3 * cap), score_b = align_up(sizeof(int) * (size_t)iterations), res_b = align_up(sizeof(double) * 5);
    size_t got = 0;
    char* d = reinterpret_cast<char*>(ctx->dev_get(orig_b + score_b + res_b, &got));
    if (!d) return ctx->fail(ARVC_E_NOMEM, "scan_fit_plane: device allocation failed");
    double* d_res = reinterpret_cast<double*>(d + orig_b + score_b);
    run_plane_fit(ctx->L, s->d_dev, s->dev.cap, reinterpret_cast<double*>(d), reinterpret_cast<int*>(d + orig_b), d_res, max_z, dist_threshold,
                  iterations, (unsigned long long)seed);
    double h_res[5] = {0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h_res, d_res, sizeof(h_res), cudaMemcpyDeviceToHost, ctx->L.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->L.stream);
    ctx->dev_put(d, got);
    if (e != cudaSuccess || ctx->L.err != cudaSuccess) return ctx->cuda_fail(e != cudaSuccess ? e : ctx->L.err, "scan_fit_plane");
    for (int k = 0; k < 4; ++k) plane_out[k] = h_res[k];
    if (n_inliers) *n_inliers = (int32_t)h_res[4];
    if (h_res[4] < 3) return ctx->fail(ARVC_E_STATE, "scan_fit_plane: fewer than three usable points below max_z");
    return ARVC_OK;
}

int arvc_scan_split_plane(arvc_ctx* ctx, int64_t src_id, const double* plane, double threshold, int64_t near_id, int64_t far_id,
                          int32_t* n_near, int32_t* n_far) {
    if (!ctx) return ARVC_E_ARG;
    if (!plane || near_id == far_id || near_id == src_id || far_id == src_id) return ctx->fail(ARVC_E_ARG, "scan_split_plane: bad arguments");
    Scan* s = ctx->find(src_id);
    if (!s || !s->preprocessed) return ctx->fail(ARVC_E_STATE, "scan_split_plane: scan not preprocessed");
    const double norm = std::sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2]);    // np.sqrt(a*a + b*b + c*c)
    if (!(norm > 0)) return ctx->fail(ARVC_E_ARG, "scan_split_plane: degenerate plane");
    CK((cudaError_t)ctx->enter());
    for (int64_t id : {near_id, far_id}) {
        Scan* old = ctx->find(id);
        if (old) { release_scan(ctx, old); ctx->scans.erase(id); }
    }
    const size_t cap = (size_t)std::max(s->dev.cap, 1);
    const size_t orig_b = align_up(sizeof(double) * 3 * cap), blk_b = align_up(sizeof(int) * ((cap + 1023) / 1024 + 8)), cnt_b = align_up(sizeof(int) * 2);
    size_t got = 0;
    char* d = reinterpret_cast<char*>(ctx->dev_get(orig_b + blk_b + cnt_b, &got));
    if (!d) return ctx->fail(ARVC_E_NOMEM, "scan_split_plane: device allocation failed");
    void *d_near = nullptr, *d_far = nullptr;
    cudaError_t e = cudaMallocAsync(&d_near, sizeof(double) * 3 * cap, ctx->L.stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_far, sizeof(double) * 3 * cap, ctx->L.stream);
    int h_cnt[2] = {0, 0};
    if (e == cudaSuccess) {
        int* d_cnt = reinterpret_cast<int*>(d + orig_b + blk_b);
        run_plane_split(ctx->L, s->d_dev, s->dev.cap, reinterpret_cast<double*>(d), reinterpret_cast<int*>(d + orig_b), reinterpret_cast<double*>(d_near),
                        reinterpret_cast<double*>(d_far), d_cnt, plane, norm, threshold);
        e = cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->L.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->L.stream);
    }
    ctx->dev_put(d, got);
    if (e != cudaSuccess || ctx->L.err != cudaSuccess) {
        if (d_near) cudaFreeAsync(d_near, ctx->L.stream);
        if (d_far) cudaFreeAsync(d_far, ctx->L.stream);
        return ctx->cuda_fail(e != cudaSuccess ? e : ctx->L.err, "scan_split_plane");
    }
    const int64_t ids[2] = {near_id, far_id};
    void* bufs[2] = {d_near, d_far};
    for (int k = 0; k < 2; ++k) {
        auto ns = std::make_unique<Scan>();
        ns->id = ids[k]; ns->n_raw = h_cnt[k]; ns->f64 = true;
        if (h_cnt[k] > 0) ns->d_raw = bufs[k];
        else cudaFreeAsync(bufs[k], ctx->L.stream);
        ctx->scans[ids[k]] = std::move(ns);
    }
    if (n_near) *n_near = h_cnt[0];
    if (n_far) *n_far = h_cnt[1];
    return ARVC_OK;
}

int arvc_scan_get_filter_indices(arvc_ctx* ctx, int64_t scan_id, int32_t* raw_index) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s || !s->preprocessed || !raw_index) return ctx->fail(ARVC_E_STATE, "scan_get_filter_indices: scan not preprocessed");
    CK((cudaError_t)ctx->enter());
    int c[CNT_WORDS];
    const int rc = fetch_counts(ctx, s, c);
    if (rc) return rc;
    if (c[CNT_NFILT] > 0) {
        CK(cudaMemcpyAsync(raw_index, s->dev.raw_index, sizeof(int) * c[CNT_NFILT], cudaMemcpyDeviceToHost, ctx->L.stream));
        CK(cudaStreamSynchronize(ctx->L.stream));
    }
    return ARVC_OK;
}

int arvc_scan_get_voxels(arvc_ctx* ctx, int64_t scan_id, int32_t* keys, int32_t* counts) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s || !s->preprocessed || !s->voxel_on) return ctx->fail(ARVC_E_STATE, "scan_get_voxels: scan not preprocessed with a voxel size");
    CK((cudaError_t)ctx->enter());
    int c[CNT_WORDS];
    const int rc = fetch_counts(ctx, s, c);
    if (rc) return rc;
    const int n = c[CNT_NPTS];
    if (n > 0) {
        if (keys) CK(cudaMemcpyAsync(keys, s->dev.vox_keys, sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, ctx->L.stream));
        if (counts) CK(cudaMemcpyAsync(counts, s->dev.vox_counts, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->L.stream));
        CK(cudaStreamSynchronize(ctx->L.stream));
    }
    return ARVC_OK;
}

int arvc_scan_get_nn_counts(arvc_ctx* ctx, int64_t scan_id, int32_t* nn_count) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s || !s->preprocessed || !s->has_normals || !nn_count) return ctx->fail(ARVC_E_STATE, "scan_get_nn_counts: no normals");
    CK((cudaError_t)ctx->enter());
    int c[CNT_WORDS];
    const int rc = fetch_counts(ctx, s, c);
    if (rc) return rc;
    const int n = c[CNT_NPTS];
    if (n == 0) return ARVC_OK;
    std::vector<int> hc(n);
    std::vector<char> hr((size_t)n * (s->dev.wide ? sizeof(RecD) : sizeof(RecF)));
    CK(cudaMemcpyAsync(hr.data(), s->dev.recs, hr.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(hc.data(), s->dev.nn_count, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    for (int p = 0; p < n; ++p) {
        const int idx = s->dev.wide ? reinterpret_cast<const RecD*>(hr.data())[p].idx : reinterpret_cast<const RecF*>(hr.data())[p].idx;
        nn_count[idx] = hc[p];
    }
    return ARVC_OK;
}

int arvc_scan_get_neighbors(arvc_ctx* ctx, int64_t scan_id, int n_query, const int32_t* point_ids, int32_t* out_idx, int32_t* out_cnt) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s || !s->preprocessed || !s->has_normals) return ctx->fail(ARVC_E_STATE, "scan_get_neighbors: no normals");
    if (!s->dev.tap_idx) return ctx->fail(ARVC_E_STATE, "scan_get_neighbors: preprocess the scan with the context option \"normals_tap\" set");
    if (n_query < 0 || (n_query > 0 && (!point_ids || !out_idx || !out_cnt))) return ctx->fail(ARVC_E_ARG, "scan_get_neighbors: bad arguments");
    CK((cudaError_t)ctx->enter());
    int c[CNT_WORDS];
    const int rc = fetch_counts(ctx, s, c);
    if (rc) return rc;
    const int n = c[CNT_NPTS], K = s->dev.tap_stride;
    // cloud order -> Morton position
    std::vector<char> hr((size_t)n * (s->dev.wide ? sizeof(RecD) : sizeof(RecF)));
    std::vector<int> hcnt(std::max(n, 1));
    if (n > 0) {
        CK(cudaMemcpyAsync(hr.data(), s->dev.recs, hr.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
        CK(cudaMemcpyAsync(hcnt.data(), s->dev.tap_cnt, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->L.stream));
        CK(cudaStreamSynchronize(ctx->L.stream));
    }
    std::vector<int> pos(std::max(n, 1), -1);
    for (int p = 0; p < n; ++p) {
        const int idx = s->dev.wide ? reinterpret_cast<const RecD*>(hr.data())[p].idx : reinterpret_cast<const RecF*>(hr.data())[p].idx;
        if (idx >= 0 && idx < n) pos[idx] = p;
    }
    for (int q = 0; q < n_query; ++q) {
        const int id = point_ids[q];
        if (id < 0 || id >= n || pos[id] < 0) return ctx->fail(ARVC_E_ARG, "scan_get_neighbors: point id out of range");
        const int p = pos[id], cnt = std::min(hcnt[p], K);
        out_cnt[q] = hcnt[p];
        for (int k = 0; k < K; ++k) out_idx[(size_t)q * K + k] = -1;
        if (cnt > 0) CK(cudaMemcpyAsync(out_idx + (size_t)q * K, s->dev.tap_idx + (size_t)p * K, sizeof(int) * cnt, cudaMemcpyDeviceToHost, ctx->L.stream));
    }
    CK(cudaStreamSynchronize(ctx->L.stream));
    return ARVC_OK;
}

int arvc_scan_get_counters(arvc_ctx* ctx, int64_t scan_id, int32_t* counters) {
    if (!ctx) return ARVC_E_ARG;
    Scan* s = ctx->find(scan_id);
    if (!s || !s->preprocessed || !counters) return ctx->fail(ARVC_E_STATE, "scan_get_counters: scan not preprocessed");
    CK((cudaError_t)ctx->enter());
    static_assert(CNT_WORDS == 16, "arvc_scan_get_counters documents 16 words");
    CK(cudaMemcpyAsync(counters, s->dev.counts, sizeof(int) * CNT_WORDS, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return ARVC_OK;
}

// ---- registration --------------------------------------------------------------------------------------
static int icp_enqueue(arvc_ctx* ctx, int n_pairs, const int64_t* tgt_ids, const int64_t* src_ids, const double* init_T,
                       const arvc_icp_params* p, bool trace, PendingBatch& pb, int** d_corr_trace, double** d_state_trace,
                       int* src_cap_out) {
    if (n_pairs < 0 || (n_pairs > 0 && (!tgt_ids || !src_ids || !init_T)) || !p) return ctx->fail(ARVC_E_ARG, "icp: bad arguments");
    if (p->method != ARVC_P2P && p->method != ARVC_P2PLANE) return ctx->fail(ARVC_E_ARG, "icp: unknown method");
    if (p->max_iter < 0) return ctx->fail(ARVC_E_ARG, "icp: max_iter < 0");
    if (n_pairs > 32768) return ctx->fail(ARVC_E_ARG, "icp: at most 32768 pairs per call (split the batch)");
    CK((cudaError_t)ctx->enter());
    pb.n_pairs = n_pairs;
    if (n_pairs == 0) return ARVC_OK;
    int combos = 0, src_cap_max = 1;
    std::vector<Scan*> S(n_pairs), T(n_pairs);
    for (int i = 0; i < n_pairs; ++i) {
        S[i] = ctx->find(src_ids[i]);
        T[i] = ctx->find(tgt_ids[i]);
        if (!S[i] || !T[i] || !S[i]->preprocessed || !T[i]->preprocessed) return ctx->fail(ARVC_E_STATE, "icp: scan not uploaded/preprocessed");
        if (p->method == ARVC_P2PLANE && !T[i]->has_normals) return ctx->fail(ARVC_E_STATE, "icp: point-to-plane needs target normals");
        const double need = p->max_corr_dist * 1.001;
        const GridSpec& g = T[i]->dev.grid;
        if (g.top_level < kMortonBits && g.c0 * (double)(1 << g.top_level) < need)
            return ctx->fail(ARVC_E_STATE, "icp: target grid was built for a smaller max_corr_dist (set grid_max_dist at preprocess)");
        combos |= 1 << (2 * S[i]->dev.wide + T[i]->dev.wide);
        src_cap_max = std::max(src_cap_max, S[i]->dev.cap);
    }
    const int passes = p->max_iter + 1;
    SlabPlanner plan;
    for (int pass = 0; pass < 2; ++pass) {
        SlabPlanner P;
        P.base = pass ? reinterpret_cast<char*>(pb.slab) : nullptr;
        PairDev* d_pairs = P.take<PairDev>(n_pairs);
        PairState* d_states = P.take<PairState>(n_pairs);
        BatchDesc* d_bd = P.take<BatchDesc>(1);
        double* d_init = P.take<double>(16 * (size_t)n_pairs);
        arvc_result_record* d_records = P.take<arvc_result_record>(n_pairs);
        int* d_status = P.take<int>(4);
        std::vector<PairDev> hp(pass ? n_pairs : 0);
        for (int i = 0; i < n_pairs; ++i) {
            const size_t cap = (size_t)S[i]->dev.cap;
            const size_t nrows = (cap + kIcpBlock - 1) / kIcpBlock * (kIcpBlock / 32);
            double* partials = P.take<double>(nrows * kSumStride);
            int* prev = P.take<int>(cap);
            float* lb2 = P.take<float>(cap);
            unsigned char* cpass = P.take<unsigned char>(cap);
            int* list = P.take<int>(cap + (cap + 255) / 256 * 32);   // + padding per block of the select kernel
            int* far_list = P.take<int>(cap);
            int* ct = trace ? P.take<int>(cap * passes) : nullptr;
            double* stt = trace ? P.take<double>((size_t)passes * 18) : nullptr;
            if (pass) {
                hp[i].src = S[i]->d_dev; hp[i].tgt = T[i]->d_dev; hp[i].state = d_states + i;
                hp[i].partials = partials; hp[i].prev = prev; hp[i].lb2 = lb2; hp[i].cert_pass = cpass; hp[i].list = list; hp[i].far_list = far_list; hp[i].corr_trace = ct; hp[i].state_trace = stt;
                if (trace && d_corr_trace) { *d_corr_trace = ct; *d_state_trace = stt; }
            }
        }
        if (!pass) {
            plan = P;
            pb.slab = ctx->dev_get(plan.off, &pb.slab_bytes);
            if (!pb.slab) return ctx->fail(ARVC_E_NOMEM, "icp: device allocation failed");
            if (trace) CK(cudaMemsetAsync(pb.slab, 0xff, plan.off, ctx->L.stream));   // untouched trace entries read as -1 / NaN
        } else {
            pb.d_states = d_states;
            pb.d_records = d_records;
            const bool debug_stats = getenv("ARVC_DEBUG_STATS") != nullptr;
            const size_t init_bytes = sizeof(double) * 16 * (size_t)n_pairs, rec_bytes = sizeof(arvc_result_record) * (size_t)n_pairs;
            pb.h_stage = reinterpret_cast<char*>(ctx->pinned_get(init_bytes + rec_bytes + 16, &pb.h_stage_bytes));
            if (debug_stats) pb.h_states = reinterpret_cast<PairState*>(ctx->pinned_get(sizeof(PairState) * n_pairs, &pb.h_bytes));
            if (!pb.h_stage || (debug_stats && !pb.h_states)) return ctx->fail(ARVC_E_NOMEM, "icp: pinned host allocation failed");
            pb.h_records = reinterpret_cast<arvc_result_record*>(pb.h_stage + init_bytes);
            pb.h_status = reinterpret_cast<int*>(pb.h_stage + init_bytes + rec_bytes);
            std::memcpy(pb.h_stage, init_T, init_bytes);
            CK(cudaMemcpyAsync(d_pairs, hp.data(), sizeof(PairDev) * n_pairs, cudaMemcpyHostToDevice, ctx->L.stream));
            CK(cudaMemcpyAsync(d_init, pb.h_stage, init_bytes, cudaMemcpyHostToDevice, ctx->L.stream));
            launch_icp_init(ctx->L, d_pairs, n_pairs, d_init, d_status);
            BatchDesc bd{};
            bd.pairs = d_pairs; bd.n_pairs = n_pairs;
            IcpParams& ip = bd.ip;
            ip.max_d = p->max_corr_dist;
            ip.max_d2 = p->max_corr_dist > 0 ? p->max_corr_dist * p->max_corr_dist : 0.0;
            ip.rel_fitness = p->rel_fitness; ip.rel_rmse = p->rel_rmse; ip.max_iter = p->max_iter; ip.method = p->method;
            {
                const char* cm = getenv("ARVC_CERT_MARGIN");      // tuning knob; results do not depend on it
                ip.cert_margin = cm ? atof(cm) : 0.0075;
                ip.debug = getenv("ARVC_DEBUG_STATS") ? atoi(getenv("ARVC_DEBUG_STATS")) : 0;
            }
            pb.graph = run_icp(ctx->L, ctx->icp_graphs, bd, d_bd, src_cap_max, combos, ctx->icp_use_graph);
            launch_icp_pack(ctx->L, d_pairs, n_pairs, d_records, d_status);
            CK(cudaMemcpyAsync(pb.h_records, d_records, rec_bytes, cudaMemcpyDeviceToHost, ctx->L.stream));
            CK(cudaMemcpyAsync(pb.h_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
            static_assert(offsetof(PairState, Thist) == kPairStateHead, "PairState head layout");
            if (debug_stats)
                CK(cudaMemcpy2DAsync(pb.h_states, sizeof(PairState), d_states, sizeof(PairState), kPairStateHead, n_pairs,
                                     cudaMemcpyDeviceToHost, ctx->L.stream));
            pb.done_ev = ctx->event_get();
            CK(cudaEventRecord(pb.done_ev, ctx->L.stream));
        }
    }
    if (src_cap_out) *src_cap_out = src_cap_max;
    if (ctx->L.err != cudaSuccess) return ctx->cuda_fail(ctx->L.err, "kernel launch");
    return ARVC_OK;
}

// `abandoned`: the batch failed while it was being enqueued - copies from its staging buffer may still be queued, so the
// stream is drained before the buffers go back to the pools (a later batch must not overwrite them in flight).
static void icp_release(arvc_ctx* ctx, PendingBatch& pb, bool abandoned = false) {
    if (abandoned) cudaStreamSynchronize(ctx->L.stream);
    if (pb.slab) { ctx->dev_put(pb.slab, pb.slab_bytes); pb.slab = nullptr; }
    ctx->pinned_put(pb.h_states, pb.h_bytes); pb.h_states = nullptr;
    ctx->pinned_put(pb.h_stage, pb.h_stage_bytes); pb.h_stage = nullptr;
    if (pb.done_ev) { ctx->event_pool.push_back(pb.done_ev); pb.done_ev = nullptr; }
}

// Waits for THIS batch only (its own event): batches enqueued later keep running, which is what lets a caller overlap the
// delivery of batch k with the kernels of batch k + 1.
static int icp_collect(arvc_ctx* ctx, PendingBatch& pb, arvc_result_record* rec) {
    if (pb.n_pairs == 0) return ARVC_OK;
    const cudaError_t werr = pb.done_ev ? cudaEventSynchronize(pb.done_ev) : cudaStreamSynchronize(ctx->L.stream);
    struct Release { arvc_ctx* c; PendingBatch& b; ~Release() { icp_release(c, b); } } release{ctx, pb};
    if (werr != cudaSuccess) return ctx->cuda_fail(werr, "icp: waiting for the batch");
    if (ctx->L.err != cudaSuccess) return ctx->cuda_fail(ctx->L.err, "kernel launch");
    if (pb.h_states) {      // ARVC_DEBUG_STATS
        unsigned long long tot[8] = {0};
        long long passes = 0;
        for (int i = 0; i < pb.n_pairs; ++i) { for (int k = 0; k < 8; ++k) tot[k] += pb.h_states[i].dbg[k]; passes += pb.h_states[i].passes; }
        if (atoi(getenv("ARVC_DEBUG_STATS")) & 8) {
            unsigned long long mx = 0;
            for (int i = 0; i < pb.n_pairs; ++i) mx = std::max(mx, (unsigned long long)pb.h_states[i].dbg[0]);
            fprintf(stderr, "[arvc warp-time] warm-pass search warps by duration: <10us:%llu <20:%llu <40:%llu <80:%llu <160:%llu <320:%llu >=320:%llu  max=%.0f us\n",
                    tot[1], tot[2], tot[3], tot[4], tot[5], tot[6], tot[7], (double)mx / 1965.0);
        }
        fprintf(stderr, "[arvc stats] pairs=%d passes=%lld queries=%llu searched=%llu (union=%llu fallback=%llu)\n", pb.n_pairs, passes,
                tot[3], tot[0], tot[1], tot[2]);
    }
    if (pb.graph) {      // kernels the device-terminated loop executed: the prefix + one body per pass beyond it
        int max_passes = 0;
        for (int i = 0; i < pb.n_pairs; ++i) max_passes = std::max(max_passes, (int)pb.h_records[i].passes);
        ctx->L.launches += pb.graph->kernels_prefix + (long long)pb.graph->kernels_body * std::max(0, max_passes - kIcpUnrolled);
    }
    const int flags = *pb.h_status;
    if (flags & 0x100) return ctx->fail(ARVC_E_CUDA, "icp: a pair did not terminate (internal error)");
    if (rec) std::memcpy(rec, pb.h_records, sizeof(arvc_result_record) * (size_t)pb.n_pairs);
    if (flags & ERR_HASH_FULL) return ctx->fail(ARVC_E_CAPACITY, "hash grid overflow in a scan of this batch");
    if (flags & ERR_VOXEL_RANGE) return ctx->fail(ARVC_E_CAPACITY, "voxel index outside the range implied by the filter bounds");
    return ARVC_OK;
}

int arvc_icp_batch_async(arvc_ctx* ctx, int n_pairs, const int64_t* tgt_ids, const int64_t* src_ids, const double* init_T,
                         const arvc_icp_params* p, uint64_t* ticket) {
    if (!ctx || !ticket) return ARVC_E_ARG;
    PendingBatch pb;
    const int rc = icp_enqueue(ctx, n_pairs, tgt_ids, src_ids, init_T, p, false, pb, nullptr, nullptr, nullptr);
    if (rc) { icp_release(ctx, pb, true); return rc; }
    *ticket = ctx->next_ticket++;
    ctx->pending[*ticket] = std::move(pb);
    return ARVC_OK;
}

int arvc_icp_batch_device_records(arvc_ctx* ctx, uint64_t ticket, const void** d_records, int* n_pairs) {
    if (!ctx || !d_records) return ARVC_E_ARG;
    auto it = ctx->pending.find(ticket);
    if (it == ctx->pending.end()) return ctx->fail(ARVC_E_ARG, "icp_batch_device_records: unknown ticket");
    *d_records = it->second.d_records;
    if (n_pairs) *n_pairs = it->second.n_pairs;
    return ARVC_OK;
}

int arvc_icp_batch_finish(arvc_ctx* ctx, uint64_t ticket, arvc_result_record* records) {
    if (!ctx) return ARVC_E_ARG;
    auto it = ctx->pending.find(ticket);
    if (it == ctx->pending.end()) return ctx->fail(ARVC_E_ARG, "icp_batch_finish: unknown ticket");
    const int rc = icp_collect(ctx, it->second, records);
    ctx->pending.erase(it);
    return rc;
}

int arvc_icp_batch(arvc_ctx* ctx, int n_pairs, const int64_t* tgt_ids, const int64_t* src_ids, const double* init_T,
                   const arvc_icp_params* p, double* out_T, double* fitness, double* rmse, int32_t* updates, int32_t* n_corr) {
    if (!ctx) return ARVC_E_ARG;
    uint64_t ticket = 0;
    int rc = arvc_icp_batch_async(ctx, n_pairs, tgt_ids, src_ids, init_T, p, &ticket);
    if (rc) return rc;
    std::vector<arvc_result_record> rec(std::max(n_pairs, 1));
    rc = arvc_icp_batch_finish(ctx, ticket, rec.data());
    if (rc) return rc;
    for (int i = 0; i < n_pairs; ++i) {
        if (out_T) std::memcpy(out_T + 16 * (size_t)i, rec[i].T, sizeof(double) * 16);
        if (fitness) fitness[i] = rec[i].fitness;
        if (rmse) rmse[i] = rec[i].rmse;
        if (updates) updates[i] = rec[i].updates;
        if (n_corr) n_corr[i] = rec[i].n_corr;
    }
    return ARVC_OK;
}

int arvc_icp_trace(arvc_ctx* ctx, int64_t tgt_id, int64_t src_id, const double* init_T, const arvc_icp_params* p, int32_t* corr,
                   double* trace_T, double* trace_fitness, double* trace_rmse, int32_t* n_passes, arvc_result_record* result) {
    if (!ctx) return ARVC_E_ARG;
    PendingBatch pb;
    int* d_ct = nullptr;
    double* d_st = nullptr;
    int cap = 0;
    int rc = icp_enqueue(ctx, 1, &tgt_id, &src_id, init_T, p, true, pb, &d_ct, &d_st, &cap);
    if (rc) { icp_release(ctx, pb, true); return rc; }
    const int passes_max = p->max_iter + 1;
    std::vector<int> hct((size_t)cap * passes_max);
    std::vector<double> hst((size_t)passes_max * 18);
    CK(cudaMemcpyAsync(hct.data(), d_ct, sizeof(int) * hct.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(hst.data(), d_st, sizeof(double) * hst.size(), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));      // the trace copies are queued behind the batch's own completion event
    arvc_result_record rec;
    rc = icp_collect(ctx, pb, &rec);
    if (rc) return rc;
    Scan* s = ctx->find(src_id);
    int c[CNT_WORDS];
    rc = fetch_counts(ctx, s, c);
    if (rc) return rc;
    const int ns = c[CNT_NPTS];
    for (int ps = 0; ps < rec.passes; ++ps) {
        if (corr) std::memcpy(corr + (size_t)ps * ns, hct.data() + (size_t)ps * cap, sizeof(int) * ns);
        if (trace_T) std::memcpy(trace_T + 16 * (size_t)ps, hst.data() + 18 * (size_t)ps, sizeof(double) * 16);
        if (trace_fitness) trace_fitness[ps] = hst[18 * (size_t)ps + 16];
        if (trace_rmse) trace_rmse[ps] = hst[18 * (size_t)ps + 17];
    }
    if (n_passes) *n_passes = rec.passes;
    if (result) *result = rec;
    return ARVC_OK;
}

// LZF decompression (liblzf stream format) for DATA binary_compressed PCD files; host-only helper of the load path.
// Returns the number of bytes written, or -1 on a malformed / oversized stream.
long long arvc_lzf_decompress(const unsigned char* in, size_t n_in, unsigned char* out, size_t n_out) {
    if ((!in && n_in) || (!out && n_out)) return -1;
    size_t ip = 0, op = 0;
    while (ip < n_in) {
        const unsigned ctrl = in[ip++];
        if (ctrl < 32) {                       // literal run of ctrl + 1 bytes
            const size_t len = ctrl + 1;
            if (ip + len > n_in || op + len > n_out) return -1;
            std::memcpy(out + op, in + ip, len);
            ip += len; op += len;
        } else {                               // back reference
            size_t len = ctrl >> 5;
            if (len == 7) { if (ip >= n_in) return -1; len += in[ip++]; }
            if (ip >= n_in) return -1;
            const size_t off = ((size_t)(ctrl & 0x1f) << 8) + in[ip++] + 1;
            len += 2;
            if (off > op || op + len > n_out) return -1;
            for (size_t k = 0; k < len; ++k) { out[op] = out[op - off]; ++op; }     // may overlap: byte by byte
        }
    }
    return (long long)op;
}

void* arvc_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void* arvc_ctx_host_alloc(arvc_ctx* ctx, size_t bytes) {
    if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    return arvc_host_alloc(bytes);
}
void arvc_host_free(void* p) { if (p) cudaFreeHost(p); }

int arvc_ctx_reserve(arvc_ctx* ctx, size_t bytes) {
    if (!ctx) return ARVC_E_ARG;
    CK((cudaError_t)ctx->enter());
    if (bytes == 0) return ARVC_OK;
    // one allocation of that size, given straight back: the device's memory pool keeps it (release threshold = max), and
    // later per-scan / per-batch allocations are carved out of it instead of growing the pool while kernels run
    void* p = nullptr;
    const cudaError_t e = cudaMallocAsync(&p, bytes, ctx->L.stream);
    if (e != cudaSuccess) { cudaGetLastError(); return ctx->fail(ARVC_E_NOMEM, "ctx_reserve: device allocation failed"); }
    CK(cudaFreeAsync(p, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return ARVC_OK;
}

}  // extern "C"
