// api.cu -- host side of the engine and the C ABI declared in include/sbr_b200.h.
//
// Mirrors, in C++ (no Rust toolchain in this image), the reference's host logic for the fit()/predict() path:
//   data.rs:213-265,406-432      CompressedInteractions (stable sort, CSR, first-chunk-smallest chunker)
//   sequence_model.rs:70-98      sub-sequence build, master shuffle, partitioning, per-partition rngs
//   lstm.rs:54-202 / ewma.rs     Hyperparameters builder and defaults
// and drives the sm_100a kernels.  There is no CPU compute path: without a usable device every compute entry
// point fails with SBR_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include <dlfcn.h>
#include <type_traits>

#include "../../include/sbr_b200.h"
#include "engine.h"
#include "nccl_dyn.h"

using namespace sbr;

namespace sbr {
// NCCL is resolved on first use (nccl_dyn.h): the copy already mapped in the process, else the system library.
const NcclApi* nccl_api(std::string* why) {
    static NcclApi api;
    static std::string err;
    static std::once_flag once;
    static bool ok = false;
    std::call_once(once, []() {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) { const char* e = dlerror(); err = std::string("cannot load libnccl.so.2: ") + (e ? e : "?"); return; }
        bool all = true;
        auto bind = [&](auto& fn, const char* name) {
            void* s = dlsym(h, name);
            if (!s) { all = false; err = std::string("libnccl.so.2 lacks ") + name; }
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(s);
        };
        bind(api.GetUniqueId, "ncclGetUniqueId"); bind(api.CommInitRank, "ncclCommInitRank");
        bind(api.CommDestroy, "ncclCommDestroy"); bind(api.GetErrorString, "ncclGetErrorString");
        bind(api.GroupStart, "ncclGroupStart"); bind(api.GroupEnd, "ncclGroupEnd");
        bind(api.Send, "ncclSend"); bind(api.Recv, "ncclRecv");
        bind(api.AllReduce, "ncclAllReduce"); bind(api.AllGather, "ncclAllGather");
        bind(api.GetVersion, "ncclGetVersion");
        ok = all;
    });
    if (!ok) { if (why) *why = err; return nullptr; }
    return &api;
}
}  // namespace sbr

namespace {

thread_local std::string g_err;
int g_device = 0;
ncclComm_t g_comm = nullptr;   // process-wide communicator for the synchronous exchange (sbr_dist_init)
int g_rank = 0, g_world = 1;

sbr_status fail(sbr_status s, const std::string& msg) { g_err = msg; return s; }
sbr_status cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return SBR_ERR_CUDA;
}
#define CU(expr)                                                      \
    do {                                                              \
        cudaError_t e__ = (expr);                                     \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr);         \
    } while (0)

struct DeviceInfo { bool ok = false; int sms = 0; int device = -1; std::string why; };

// properties of the CURRENTLY selected device (re-queried whenever sbr_set_device picked another one)
DeviceInfo& device_info() {
    static DeviceInfo info;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (info.device == g_device) return info;
    info = DeviceInfo();
    info.device = g_device;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { info.why = "no CUDA device available (this library has no CPU fallback)"; cudaGetLastError(); return info; }
    if (g_device >= n) { info.why = "selected device index out of range"; return info; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, g_device) != cudaSuccess) { info.why = "cudaGetDeviceProperties failed"; return info; }
    if (p.major != 10) { info.why = "device is not sm_100 (B200): kernels are built for sm_100a only"; return info; }
    info.sms = p.multiProcessorCount;
    info.ok = true;
    return info;
}

sbr_status require_device() {
    DeviceInfo& d = device_info();
    if (!d.ok) return fail(SBR_ERR_CUDA, d.why);
    CU(cudaSetDevice(g_device));
    return SBR_OK;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Transient device buffers of a fit() (id stream mirror, schedule arrays) are recycled: cudaMalloc / cudaFree of a
// 128 MB buffer cost milliseconds each and cudaFree synchronises the device, which shows up directly in the
// end-to-end (host CSR in, loss out) rate.  Freed blocks are kept on a bounded free list and handed out again to
// requests of a similar size; everything here is used on one stream and released only after that stream was synced.
struct DevPool {
    std::mutex mu;
    struct Blk { void* p; size_t bytes; };
    std::vector<Blk> free_;
    size_t held = 0;
    static constexpr size_t kMaxHeld = (size_t)2 << 30; static constexpr size_t kMaxBlocks = 24;
    cudaError_t alloc(void** out, size_t bytes) {
        bytes = std::max<size_t>((bytes + 255) & ~(size_t)255, 256);
        {
            std::lock_guard<std::mutex> lk(mu);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].bytes >= bytes && free_[i].bytes <= 2 * bytes + (1u << 20) && (best == free_.size() || free_[i].bytes < free_[best].bytes)) best = i;
            if (best != free_.size()) { *out = free_[best].p; held -= free_[best].bytes; sizes[*out] = free_[best].bytes; free_.erase(free_.begin() + best); return cudaSuccess; }
        }
        cudaError_t e = cudaMalloc(out, bytes);
        if (e != cudaSuccess) {   // make room and retry once
            trim(0);
            e = cudaMalloc(out, bytes);
        }
        if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(mu); sizes[*out] = bytes; }
        return e;
    }
    void release(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        auto it = sizes.find(p);
        if (it == sizes.end()) { cudaFree(p); return; }
        const size_t bytes = it->second; sizes.erase(it);
        if (bytes > kMaxHeld / 2) { cudaFree(p); return; }
        free_.push_back({p, bytes}); held += bytes;
        while (held > kMaxHeld || free_.size() > kMaxBlocks) { held -= free_.front().bytes; cudaFree(free_.front().p); free_.erase(free_.begin()); }
    }
    void trim(size_t keep) {
        std::lock_guard<std::mutex> lk(mu);
        while (held > keep && !free_.empty()) { held -= free_.front().bytes; cudaFree(free_.front().p); free_.erase(free_.begin()); }
    }
    std::unordered_map<void*, size_t> sizes;
};
DevPool g_pool;
template <typename T> cudaError_t pool_alloc(T** out, size_t bytes) { void* p = nullptr; cudaError_t e = g_pool.alloc(&p, bytes); *out = static_cast<T*>(p); return e; }

// pinned double-buffer used to narrow usize ids to u32 on the way to HBM
struct Staging {
    static constexpr size_t kElems = 8u << 20;  // 32 MiB per buffer
    uint32_t* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2];
    bool ready = false;
    std::mutex mu;
    sbr_status init() {
        if (ready) return SBR_OK;
        for (int i = 0; i < 2; ++i) {
            CU(cudaHostAlloc(&buf[i], kElems * sizeof(uint32_t), cudaHostAllocDefault));
            CU(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        ready = true;
        return SBR_OK;
    }
};
Staging g_staging;

// narrow + upload n ids (validating < bound); dst is device memory.
// Pageable source: up to 12 host threads narrow usize -> u32 into a pinned double buffer while the previous piece is on
// the wire (half the PCIe bytes).  Page-locked source (cudaHostAlloc / cudaHostRegister'ed by the caller): the raw
// 64-bit words are DMA'ed straight from the caller's buffer, piece by piece, and narrowed by a kernel behind each
// piece -- no host thread touches the stream.
sbr_status upload_ids_pinned(const uint64_t* src, size_t n, uint32_t* dst, cudaStream_t st, uint64_t bound, size_t* h2d_bytes) {
    constexpr size_t kPiece = 4u << 20;   // 32 MiB of raw words per piece; copy and narrow kernel alternate on the stream
    uint64_t* raw = nullptr; int* d_bad = nullptr;
    CU(pool_alloc(&raw, std::min(n, kPiece) * sizeof(uint64_t)));
    if (pool_alloc(&d_bad, 256) != cudaSuccess) { g_pool.release(raw); return cuda_fail(cudaGetLastError(), "cudaMalloc"); }
    auto done_ = [&](sbr_status r) { g_pool.release(raw); g_pool.release(d_bad); return r; };
#define CUX(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return done_(cuda_fail(e__, #expr)); } while (0)
    CUX(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    for (size_t done = 0; done < n; done += kPiece) {
        const size_t cnt = std::min(kPiece, n - done);
        CUX(cudaMemcpyAsync(raw, src + done, cnt * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        CUX(launch_narrow_ids(raw, cnt, bound, dst + done, d_bad, st));
    }
    int bad = 0;
    CUX(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUX(cudaStreamSynchronize(st));
#undef CUX
    if (h2d_bytes) *h2d_bytes += n * sizeof(uint64_t);
    if (bad) return done_(fail(SBR_ERR_INVALID_ARGUMENT, "item id out of range"));
    return done_(SBR_OK);
}

sbr_status upload_ids_u32(const uint64_t* src, size_t n, uint32_t* dst, cudaStream_t st, uint64_t bound, size_t* h2d_bytes) {
    if (n >= (1u << 16)) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost) return upload_ids_pinned(src, n, dst, st, bound, h2d_bytes);
        cudaGetLastError();
    }
    std::lock_guard<std::mutex> lk(g_staging.mu);
    sbr_status s = g_staging.init();
    if (s) return s;
    size_t done = 0; int b = 0; bool used[2] = {false, false};
    while (done < n) {
        const size_t cnt = std::min(Staging::kElems, n - done);
        if (used[b]) CU(cudaEventSynchronize(g_staging.ev[b]));
        uint32_t* out = g_staging.buf[b];
        std::atomic<uint64_t> bad{0};
        auto narrow = [&](size_t lo, size_t hi) {
            uint64_t bd = 0;
            for (size_t i = lo; i < hi; ++i) { const uint64_t v = src[done + i]; bd |= (uint64_t)(v >= bound); out[i] = (uint32_t)v; }
            if (bd) bad.store(1);
        };
        const size_t nth = cnt >= (1u << 20) ? std::min<size_t>(12, std::max(1u, std::thread::hardware_concurrency())) : 1;
        if (nth <= 1) narrow(0, cnt);
        else {
            std::vector<std::thread> th;
            const size_t per = (cnt + nth - 1) / nth;
            for (size_t k = 0; k < nth; ++k) th.emplace_back(narrow, std::min(cnt, k * per), std::min(cnt, (k + 1) * per));
            for (auto& t : th) t.join();
        }
        if (bad.load()) return fail(SBR_ERR_INVALID_ARGUMENT, "item id out of range");
        CU(cudaMemcpyAsync(dst + done, out, cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(g_staging.ev[b], st));
        used[b] = true; b ^= 1; done += cnt;
    }
    for (int i = 0; i < 2; ++i) if (used[i]) CU(cudaEventSynchronize(g_staging.ev[i]));
    if (h2d_bytes) *h2d_bytes += n * sizeof(uint32_t);
    return SBR_OK;
}

}  // namespace

// ============================================================================================================
// handles
// ============================================================================================================
struct sbr_compressed {
    size_t num_users = 0, num_items = 0;
    std::vector<uint64_t> user_ptr, item_ids, timestamps;   // owned storage (empty when borrowed)
    const uint64_t* up_ = nullptr; const uint64_t* ii_ = nullptr; const uint64_t* tt_ = nullptr;  // views actually used
    size_t nnz_ = 0;
    void own_views() { up_ = user_ptr.data(); ii_ = item_ids.data(); tt_ = timestamps.data(); nnz_ = item_ids.size(); }
    // lazily created HBM mirror (immutable data => safe to share between const users)
    mutable std::mutex mu;
    mutable uint32_t* d_item_ids = nullptr;
    mutable uint64_t* d_user_ptr = nullptr;
    mutable size_t upload_bytes = 0;  // bytes moved by the most recent upload (0 if it was already resident)
    ~sbr_compressed() {
        g_pool.release(d_item_ids);
        g_pool.release(d_user_ptr);
    }
};

struct sbr_hyperparameters {
    int model = MODEL_LSTM;
    size_t num_items = 0, max_sequence_length = 0, embedding_dim = 16;   // lstm.rs:58-60
    float learning_rate = 0.01f, l2_penalty = 0.0f;                      // lstm.rs:61-62
    int lstm_variant = SBR_LSTM_COUPLED, loss = SBR_LOSS_BPR;            // lstm.rs:63-64
    int optimizer = SBR_OPTIMIZER_ADAM, parallelism = SBR_PARALLELISM_SYNCHRONOUS;  // lstm.rs:65-66
    uint8_t seed[16];
    size_t num_threads = 0;                                              // lstm.rs:68 (0 = fill the device)
    size_t num_epochs = 10;                                              // lstm.rs:69
    int shard_rank = 0, shard_world = 1;   // one process per GPU: this rank owns rows with id % world == rank
    int virtual_shards = 1;                // > 1: shard the table inside this process (exercises the sharded addressing)
    int exact = 0;                         // sbr_hyper_exact_arithmetic
};

struct sbr_model {
    sbr_hyperparameters h;
    ModelDev dev{};
    XorShift rng{};          // Hyperparameters.rng (lstm.rs:49), advanced by every fit (sequence_model.rs:84,97)
    uint64_t num_updates = 0;
    cudaStream_t stream = nullptr;
    mutable std::mutex mu;
    sbr_fit_stats last{};
    // sharding: own[i] != 0 => Es[i]/Bs[i] were cudaMalloc'ed here; otherwise they are CUDA-IPC mappings of a peer
    bool own[8] = {false, false, false, false, false, false, false, false};
    float* own_dense = nullptr;   // this process's dense buffer (dev.dense points at rank 0's after attach)
    bool attached = true;         // false between build() and sbr_model_ipc_attach() when shard_world > 1
    float* scratch_cache = nullptr; size_t scratch_cap = 0; bool scratch_busy = false;  // grow-only activation scratch
    uint8_t* h_stage = nullptr; size_t h_stage_cap = 0;   // grow-only pinned staging area of fit(): shuffled order | partition rngs | keys
    float* replica_snap = nullptr; float* replica_delta = nullptr; size_t replica_n = 0;   // sbr_model_replica_sync: common starting point, delta buffer
    ~sbr_model() {
        for (int i = 0; i < 8; ++i) {
            if (own[i]) { if (dev.Es[i]) cudaFree(dev.Es[i]); }
            else if (dev.Es[i]) cudaIpcCloseMemHandle(dev.Es[i]);
        }
        if (dev.dense && dev.dense != own_dense) cudaIpcCloseMemHandle(dev.dense);
        if (own_dense) cudaFree(own_dense);
        if (scratch_cache) cudaFree(scratch_cache);
        if (h_stage) cudaFreeHost(h_stage);
        if (replica_snap) cudaFree(replica_snap);
        if (replica_delta) cudaFree(replica_delta);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct sbr_fit_plan {
    sbr_model* model = nullptr;
    PlanDev dev{};
    uint64_t* d_seq_start = nullptr; uint32_t* d_seq_len = nullptr;
    void* tmp[2] = {nullptr, nullptr};   // chunker scratch, alive until the plan's staging work has run
    size_t nsub = 0, P = 0, n = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    sbr_fit_stats stats{};
    bool scratch_borrowed = false;
    SyncBuffers* sync = nullptr;
    BatchBuffers* batch = nullptr;
    bool use_batch = false;   // LSTM on the round-synchronous batched tensor-core engine (lstm_batch.cuh)
    bool cold_bounded = false;   // automatic partition count was held down because the model is still cold
    int epochs_override = -1;    // >= 0: run this many epochs instead of hyper.num_epochs
    ~sbr_fit_plan() {
        if (sync) sync_buffers_free(sync);
        if (batch) batch_buffers_free(batch);
        g_pool.release(d_seq_start); g_pool.release(d_seq_len); g_pool.release(tmp[0]); g_pool.release(tmp[1]);
        g_pool.release(dev.order); g_pool.release(dev.rng); g_pool.release(dev.keys);
        g_pool.release(dev.step_ctr); g_pool.release(dev.loss_acc); g_pool.release(dev.examples);
        if (dev.scratch && !scratch_borrowed) cudaFree(dev.scratch);
        if (scratch_borrowed && model) model->scratch_busy = false;
        for (cudaEvent_t e : {ev0, ev1, evk0, evk1}) if (e) cudaEventDestroy(e);
    }
};

namespace {

// the CSR's HBM mirror, created lazily: user_ptr first (the device chunker needs nothing else), then the id stream
sbr_status ensure_user_ptr(const sbr_compressed* c, cudaStream_t st, size_t* bytes) {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->d_user_ptr) return SBR_OK;
    sbr_status s = require_device();
    if (s) return s;
    uint64_t* dp = nullptr;
    CU(pool_alloc(&dp, (c->num_users + 1) * sizeof(uint64_t)));
    cudaError_t e = cudaMemcpyAsync(dp, c->up_, (c->num_users + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { g_pool.release(dp); return cuda_fail(e, "upload user_ptr"); }
    c->d_user_ptr = dp;
    if (bytes) *bytes += (c->num_users + 1) * sizeof(uint64_t);
    return SBR_OK;
}
sbr_status ensure_ids(const sbr_compressed* c, cudaStream_t st, size_t* bytes) {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->d_item_ids || c->nnz_ == 0) return SBR_OK;
    sbr_status s = require_device();
    if (s) return s;
    uint32_t* d = nullptr;
    CU(pool_alloc(&d, std::max<size_t>(c->nnz_, 1) * sizeof(uint32_t)));
    s = upload_ids_u32(c->ii_, c->nnz_, d, st, c->num_items, bytes);
    if (s) { g_pool.release(d); return s; }
    c->d_item_ids = d;
    return SBR_OK;
}
sbr_status ensure_uploaded(const sbr_compressed* c, cudaStream_t st) {
    size_t bytes = 0;
    sbr_status s = ensure_user_ptr(c, st, &bytes);
    if (s == SBR_OK) s = ensure_ids(c, st, &bytes);
    if (s == SBR_OK && cudaStreamSynchronize(st) != cudaSuccess) s = cuda_fail(cudaGetLastError(), "upload");
    c->upload_bytes = bytes;
    return s;
}

// kept chunks of one user (data.rs:406-432 + the len > 2 filter of sequence_model.rs:81)
inline size_t user_kept_chunks(uint64_t len, uint64_t T) {
    if (len == 0) return 0;
    const uint64_t first = ((len | T) >> 32) == 0 ? (uint64_t)((uint32_t)len % (uint32_t)T) : len % T;
    return (T > 2 ? (len - first) / T : 0) + (first > 2 ? 1 : 0);
}
// number of sub-sequences fit() trains on -- all the host needs to know about them: the chunks themselves are built on
// the device (data_prep.cu device_schedule) while the host runs the master shuffle over their indices
size_t host_count_subsequences(const sbr_compressed* c, size_t T) {
    const size_t U = c->num_users;
    const size_t nth = U >= (1u << 18) ? std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
    std::vector<size_t> part(nth, 0);
    auto work = [&](size_t k) {
        const size_t lo = U * k / nth, hi = U * (k + 1) / nth;
        size_t n = 0;
        for (size_t u = lo; u < hi; ++u) n += user_kept_chunks(c->up_[u + 1] - c->up_[u], T);
        part[k] = n;
    };
    if (nth == 1) work(0);
    else { std::vector<std::thread> th; for (size_t k = 0; k < nth; ++k) th.emplace_back(work, k); for (auto& t : th) t.join(); }
    size_t n = 0; for (size_t v : part) n += v;
    return n;
}

// ---- xorshift128 jump-ahead ----
// The master rng's stream is what makes the schedule sequential: 2 (nsub - 1) outputs for the shuffle's swap partners
// (sequence_model.rs:84; ~1.4 x that with gen_range's redraws), then 16 per partition for the thread rngs (:97) -- ~3.5 M
// dependent steps for the bench stream, 11 ms on one core.  xorshift128 is linear over GF(2): the state after k steps is
// J^k s.  The stream is cut into segments of 2^14 steps whose start states come from the cached matrix J^(2^14); host
// threads then produce the segments side by side and the swap loop (inherently sequential, ~3 ns per swap with the
// partners computed 16 positions ahead) follows behind them.  Same stream, same order, same results as the one-thread
// loop (tests/test_abi_and_host.py compares both paths).
struct U128 { uint64_t lo, hi; };   // x = lo[0,32) y = lo[32,64) z = hi[0,32) w = hi[32,64)
inline U128 xs_pack(const XorShift& r) { return {(uint64_t)r.x | ((uint64_t)r.y << 32), (uint64_t)r.z | ((uint64_t)r.w << 32)}; }
inline XorShift xs_unpack(const U128& v) { XorShift r; r.x = (uint32_t)v.lo; r.y = (uint32_t)(v.lo >> 32); r.z = (uint32_t)v.hi; r.w = (uint32_t)(v.hi >> 32); return r; }
struct XsMat {
    U128 col[128];   // col[i] = image of basis vector e_i
    U128 apply(U128 v) const {
        U128 r{0, 0};
        for (uint64_t m = v.lo; m; m &= m - 1) { const U128& c = col[__builtin_ctzll(m)]; r.lo ^= c.lo; r.hi ^= c.hi; }
        for (uint64_t m = v.hi; m; m &= m - 1) { const U128& c = col[64 + __builtin_ctzll(m)]; r.lo ^= c.lo; r.hi ^= c.hi; }
        return r;
    }
};
constexpr int kSegLog2 = 14;
constexpr size_t kSegSteps = (size_t)1 << kSegLog2;
const XsMat& xs_segment_jump() {
    static const XsMat J = [] {
        XsMat m;
        for (int i = 0; i < 128; ++i) {
            U128 e{0, 0};
            if (i < 64) e.lo = 1ull << i; else e.hi = 1ull << (i - 64);
            XorShift r = xs_unpack(e);
            xs_next_u32(r);
            m.col[i] = xs_pack(r);
        }
        for (int k = 0; k < kSegLog2; ++k) {   // square: J <- J . J
            XsMat n;
            for (int i = 0; i < 128; ++i) n.col[i] = m.apply(m.col[i]);
            m = n;
        }
        return m;
    }();
    return J;
}

// One-thread reference of the master rng's work in fit(): Fisher-Yates from the top over the indices 0..nsub
// (parameters.rng().shuffle(&mut subsequences), sequence_model.rs:84), then XorShiftRng::from_seed(parameters.rng().gen())
// per partition (:97; gen::<[u8; 16]>() takes the low byte of 16 successive outputs).
void master_shuffle_serial(XorShift& rng, uint32_t* order, size_t nsub) {
    for (size_t i = 0; i < nsub; ++i) order[i] = (uint32_t)i;
    constexpr size_t K = 16;
    size_t js[K];
    size_t i = nsub, drawn = nsub;
    auto draw = [&]() {
        drawn -= 1;
        const size_t j = (size_t)xs_gen_below(rng, (uint64_t)drawn + 1);
        js[drawn % K] = j;
        __builtin_prefetch(order + j, 1);
    };
    for (size_t k = 0; k < K && drawn >= 2; ++k) draw();
    while (i >= 2) {
        i -= 1;
        const size_t j = js[i % K];
        if (drawn >= 2) draw();
        std::swap(order[i], order[j]);
    }
}
void partition_seeds_serial(XorShift& rng, size_t P, XorShift* rngs, uint64_t* keys) {
    for (size_t p = 0; p < P; ++p) {
        uint8_t seed[16];
        for (int i = 0; i < 16; ++i) seed[i] = (uint8_t)xs_next_u32(rng);
        xs_from_seed(rngs[p], seed);
        uint64_t k = 0; for (int i = 0; i < 8; ++i) k |= (uint64_t)seed[i] << (8 * i);
        keys[p] = k;
    }
}

// the same two steps with the rng's output stream produced by `nth` host threads (jump-ahead); P == 0: shuffle only.
// gen_range redraws while the low product word exceeds its zone (p up to 1/2 per draw), so a swap partner does not sit
// at a fixed stream position.  Three stages, pipelined: nth - 1 producers fill the raw output stream segment by segment;
// one thread turns it into the partner list (widening multiply + zone test, in stream order); the caller swaps.
void master_schedule(XorShift& rng, uint32_t* order, size_t nsub, size_t P, XorShift* rngs, uint64_t* keys, size_t nth) {
    const size_t ndraw = nsub >= 2 ? nsub - 1 : 0;               // positions nsub - 1 .. 1 each pick a partner
    if (nth <= 1 || 2 * ndraw + 16 * P < 16 * kSegSteps) {
        master_shuffle_serial(rng, order, nsub);
        if (P) partition_seeds_serial(rng, P, rngs, keys);
        return;
    }
    // capacity: every draw accepts with p >= 1/2, so 2 x (2 ndraw) outputs is already twice the expected need
    const size_t nseg = (4 * ndraw + 16 * P) / kSegSteps + 2;
    const XorShift rng0 = rng;
    const XsMat& J = xs_segment_jump();
    std::vector<U128> start(nseg);
    start[0] = xs_pack(rng);
    for (size_t sg = 1; sg < nseg; ++sg) start[sg] = J.apply(start[sg - 1]);
    // scratch of the calling thread, grow-only: fresh pages cost more than the arithmetic
    static thread_local std::vector<uint32_t> raw_buf, js_buf;
    if (raw_buf.size() < nseg * kSegSteps) raw_buf.resize(nseg * kSegSteps);
    if (js_buf.size() < ndraw) js_buf.resize(ndraw);
    uint32_t* raw = raw_buf.data();
    uint32_t* js = js_buf.data();
    std::unique_ptr<std::atomic<int>[]> ready(new std::atomic<int>[nseg]);
    for (size_t sg = 0; sg < nseg; ++sg) ready[sg].store(0, std::memory_order_relaxed);
    std::atomic<size_t> next_seg{0};
    std::atomic<int> stop{0};
    auto produce = [&]() {
        for (;;) {
            if (stop.load(std::memory_order_relaxed)) return;
            const size_t sg = next_seg.fetch_add(1);
            if (sg >= nseg) return;
            XorShift r = xs_unpack(start[sg]);
            uint32_t* out = raw + sg * kSegSteps;
            for (size_t g = 0; g < kSegSteps; ++g) out[g] = xs_next_u32(r);
            ready[sg].store(1, std::memory_order_release);
        }
    };
    // stage 2, one thread: gen_range over the produced stream, branch-free (a redraw is taken with p up to 1/2 -- as a branch
    // it mispredicts every third time): the candidate partner is always written, the cursor moves only on acceptance
    std::atomic<size_t> jdone{0};
    std::atomic<int> overflow{0};
    size_t pos_end = 0;
    auto compact = [&]() {
        size_t pos = 0, avail = 0, seg_ok = 0, d = 0;
        uint64_t range = nsub;                       // position i = nsub - 1 - d draws from [0, i] : range i + 1
        while (d < ndraw) {
            if (pos + 2 > avail) {
                if (seg_ok >= nseg) { overflow.store(1); jdone.store(ndraw, std::memory_order_release); return; }
                while (!ready[seg_ok].load(std::memory_order_acquire)) std::this_thread::yield();
                ++seg_ok;
                avail = seg_ok * kSegSteps;
                jdone.store(d, std::memory_order_release);
            }
            const size_t stop_at = avail;
            while (pos + 2 <= stop_at && d < ndraw) {
                const uint64_t v = (uint64_t)raw[pos] | ((uint64_t)raw[pos + 1] << 32);
                pos += 2;
                const uint64_t zone = (range << clz64(range)) - 1;
                uint64_t lo;
                const uint64_t hi = mulhi64(v, range, &lo);
                js[d] = (uint32_t)hi;
                const uint64_t acc = lo <= zone ? 1 : 0;
                d += acc; range -= acc;
            }
        }
        pos_end = pos;
        jdone.store(ndraw, std::memory_order_release);
    };
    std::vector<std::thread> th;
    for (size_t k = 0; k + 1 < nth; ++k) th.emplace_back(produce);
    th.emplace_back(compact);
    struct Stopper { std::atomic<int>& s; std::vector<std::thread>& t; ~Stopper() { s.store(1); for (auto& x : t) x.join(); } } stopper{stop, th};
    // stage 3, this thread: the swaps, partners known ahead
    for (size_t i = 0; i < nsub; ++i) order[i] = (uint32_t)i;
    {
        constexpr size_t K = 16;
        size_t have = 0;
        for (size_t d = 0; d < ndraw; ++d) {
            const size_t pf = std::min(d + K, ndraw - 1);
            while (pf >= have) { have = jdone.load(std::memory_order_acquire); if (pf >= have) std::this_thread::yield(); }
            __builtin_prefetch(order + js[pf], 1);
            std::swap(order[nsub - 1 - d], order[js[d]]);
        }
    }
    th.back().join(); th.pop_back();   // the compactor (pos_end is final)
    size_t pos = pos_end;
    bool bad = overflow.load() != 0;
    if (!bad) {
        const size_t need_seg = (pos + 16 * P + kSegSteps - 1) / kSegSteps;
        if (need_seg > nseg) bad = true;
        else for (size_t sg = 0; sg < need_seg; ++sg) while (!ready[sg].load(std::memory_order_acquire)) std::this_thread::yield();
    }
    if (bad) {   // (the stream ran past twice its expected length: never observed) redo as the plain loop
        rng = rng0;
        master_shuffle_serial(rng, order, nsub);
        if (P) partition_seeds_serial(rng, P, rngs, keys);
        return;
    }
    for (size_t p = 0; p < P; ++p) {
        uint8_t seed[16];
        for (int i = 0; i < 16; ++i) seed[i] = (uint8_t)raw[pos + 16 * p + i];
        xs_from_seed(rngs[p], seed);
        uint64_t k = 0; for (int i = 0; i < 8; ++i) k |= (uint64_t)seed[i] << (8 * i);
        keys[p] = k;
    }
    pos += 16 * P;
    XorShift r = xs_unpack(start[pos / kSegSteps]);   // the rng after everything consumed
    for (size_t g = 0; g < pos % kSegSteps; ++g) xs_next_u32(r);
    rng = r;
}
void master_shuffle(XorShift& rng, uint32_t* order, size_t nsub) { master_schedule(rng, order, nsub, 0, nullptr, nullptr, 1); }

// sequence_model.rs:76-84 entirely on the host (the sbr_host_schedule hook: CPU-only CI pins it on the oracle; fit() builds
// the chunks on the device instead and is tested against this function): chunk every user (data.rs:406-432: the FIRST chunk
// is the short one -- len % T items, if that is not 0 -- every later chunk has exactly T items), keep len > 2 (:81), then
// the master shuffle (:84).
sbr_status host_schedule(const sbr_compressed* c, size_t T, XorShift& rng, std::vector<uint64_t>& starts, std::vector<uint32_t>& lens,
                         std::vector<uint32_t>& order) {
    if (T == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "max_sequence_length must be positive");
    starts.clear(); lens.clear();
    starts.reserve(c->nnz_ / T + c->num_users); lens.reserve(c->nnz_ / T + c->num_users);
    for (size_t u = 0; u < c->num_users; ++u) {
        const size_t b = c->up_[u], len = c->up_[u + 1] - b;
        if (len == 0) continue;
        size_t idx = 0;
        const size_t first = len % T;
        if (first != 0) { if (first > 2) { starts.push_back(b); lens.push_back((uint32_t)first); } idx = first; }
        if (T > 2) for (; idx < len; idx += T) { starts.push_back(b + idx); lens.push_back((uint32_t)T); }
    }
    const size_t nsub = starts.size();
    if (nsub == 0) return fail(SBR_ERR_NO_INTERACTIONS, "No interactions were supplied.");  // :86-88
    if (nsub > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "too many sub-sequences");
    order.resize(nsub);
    master_shuffle(rng, order.data(), nsub);
    return SBR_OK;
}

struct ParamRef { int kind; int slot; size_t off, len; };  // kind 0: E rows, 1: bias, 2: dense

sbr_status resolve_param(const sbr_model* m, const char* name, ParamRef* out) {
    std::string base(name); int slot = 0;
    if (base.size() > 3 && (base.compare(base.size() - 3, 3, ".s1") == 0 || base.compare(base.size() - 3, 3, ".s2") == 0)) {
        slot = base[base.size() - 1] - '0'; base.resize(base.size() - 3);
    }
    const size_t N = m->dev.N, D = m->dev.D;
    if (slot >= m->dev.S && !(slot <= 2 && base != "item_embeddings")) return fail(SBR_ERR_INVALID_ARGUMENT, "optimizer state slot not present for this optimizer");
    if (slot == 2 && m->dev.opt != 1) return fail(SBR_ERR_INVALID_ARGUMENT, "'.s2' exists only for Adam");
    if (base == "item_embeddings") { *out = {0, slot, 0, N * D}; return SBR_OK; }
    if (base == "item_biases") { *out = {1, slot, 0, N}; return SBR_OK; }
    if (m->dev.model == MODEL_LSTM && base == "lstm_weights") { *out = {2, slot, 0, 2 * D * 4 * D}; return SBR_OK; }
    if (m->dev.model == MODEL_LSTM && base == "lstm_biases") { *out = {2, slot, 2 * D * 4 * D, 4 * D}; return SBR_OK; }
    if (m->dev.model == MODEL_EWMA && base == "alpha") { *out = {2, slot, 0, D}; return SBR_OK; }
    return fail(SBR_ERR_INVALID_ARGUMENT, std::string("unknown parameter name: ") + name);
}

}  // namespace

// ============================================================================================================
// C ABI
// ============================================================================================================
extern "C" {

const char* sbr_last_error_string(void) { return g_err.c_str(); }

int sbr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

sbr_status sbr_set_device(int device) {
    if (device < 0) return fail(SBR_ERR_INVALID_ARGUMENT, "negative device index");
    g_device = device;
    return SBR_OK;
}

// ------------------------------------------------------------------------------------------------ data.rs --
sbr_status sbr_compressed_from_triplets(const uint64_t* user_ids, const uint64_t* item_ids, const uint64_t* timestamps,
                                        size_t nnz, size_t num_users, size_t num_items, sbr_compressed** out) try {
    if (!out || (nnz && (!user_ids || !item_ids || !timestamps))) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (num_items > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "num_items must fit in 32 bits");
    for (size_t i = 0; i < nnz; ++i) {
        if (user_ids[i] >= num_users) return fail(SBR_ERR_INVALID_ARGUMENT, "user id >= num_users");
        if (item_ids[i] >= num_items) return fail(SBR_ERR_INVALID_ARGUMENT, "item id >= num_items");
    }
    sbr_compressed* c = new (std::nothrow) sbr_compressed();
    if (!c) return fail(SBR_ERR_INVALID_ARGUMENT, "out of memory");
    c->num_users = num_users; c->num_items = num_items;
    // data.rs:240 `data.sort_by(cmp_timestamp)`: STABLE sort on (user, timestamp); ties keep input order
    std::vector<uint64_t> idx(nnz);
    for (size_t i = 0; i < nnz; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
        if (user_ids[a] != user_ids[b]) return user_ids[a] < user_ids[b];
        return timestamps[a] < timestamps[b];
    });
    c->user_ptr.assign(num_users + 1, 0);
    c->item_ids.resize(nnz); c->timestamps.resize(nnz);
    for (size_t i = 0; i < nnz; ++i) {  // data.rs:246-251
        const uint64_t s = idx[i];
        c->item_ids[i] = item_ids[s]; c->timestamps[i] = timestamps[s];
        c->user_ptr[user_ids[s] + 1] += 1;
    }
    for (size_t u = 1; u <= num_users; ++u) c->user_ptr[u] += c->user_ptr[u - 1];  // data.rs:253-255
    c->own_views();
    *out = c;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_compressed_from_triplets: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_compressed_from_triplets: unknown C++ exception"); }

sbr_status sbr_compressed_from_triplets_device(const uint64_t* user_ids, const uint64_t* item_ids, const uint64_t* timestamps,
                                               size_t nnz, size_t num_users, size_t num_items, sbr_compressed** out) try {
    if (!out || (nnz && (!user_ids || !item_ids || !timestamps))) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (num_items > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "num_items must fit in 32 bits");
    sbr_status s = require_device();
    if (s) return s;
    sbr_compressed* c = new (std::nothrow) sbr_compressed();
    if (!c) return fail(SBR_ERR_INVALID_ARGUMENT, "out of memory");
    c->num_users = num_users; c->num_items = num_items;
    c->user_ptr.assign(num_users + 1, 0);
    c->item_ids.resize(nnz); c->timestamps.resize(nnz);
    uint32_t* d_ids = nullptr; uint64_t* d_ptr = nullptr;
    cudaError_t e = pool_alloc(&d_ids, std::max<size_t>(nnz, 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = pool_alloc(&d_ptr, (num_users + 1) * sizeof(uint64_t));
    if (e != cudaSuccess) { g_pool.release(d_ids); g_pool.release(d_ptr); delete c; return cuda_fail(e, "cudaMalloc CSR mirror"); }
    std::string err;
    const int rc = device_csr_build(user_ids, item_ids, timestamps, nnz, num_users, num_items, c->user_ptr.data(), c->item_ids.data(),
                                    c->timestamps.data(), d_ids, d_ptr, nullptr, &err);
    if (rc) { g_pool.release(d_ids); g_pool.release(d_ptr); delete c; return fail(rc == 2 ? SBR_ERR_INVALID_ARGUMENT : SBR_ERR_CUDA, err); }
    c->own_views();
    c->d_item_ids = d_ids; c->d_user_ptr = d_ptr;   // the CSR is already resident: the first fit() uploads nothing
    *out = c;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_compressed_from_triplets_device: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_compressed_from_triplets_device: unknown C++ exception"); }

sbr_status sbr_compressed_from_csr(const uint64_t* user_pointers, const uint64_t* item_ids, const uint64_t* timestamps,
                                   size_t num_users, size_t num_items, sbr_compressed** out) try {
    if (!out || !user_pointers) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (num_items > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "num_items must fit in 32 bits");
    if (user_pointers[0] != 0) return fail(SBR_ERR_INVALID_ARGUMENT, "user_pointers[0] must be 0");
    for (size_t u = 0; u < num_users; ++u)
        if (user_pointers[u + 1] < user_pointers[u]) return fail(SBR_ERR_INVALID_ARGUMENT, "user_pointers must be non-decreasing");
    const size_t nnz = user_pointers[num_users];
    if (nnz && !item_ids) return fail(SBR_ERR_INVALID_ARGUMENT, "null item_ids");
    sbr_compressed* c = new (std::nothrow) sbr_compressed();
    if (!c) return fail(SBR_ERR_INVALID_ARGUMENT, "out of memory");
    c->num_users = num_users; c->num_items = num_items;
    c->user_ptr.assign(user_pointers, user_pointers + num_users + 1);
    c->item_ids.assign(item_ids, item_ids + nnz);   // range-checked when narrowed for upload
    if (timestamps) c->timestamps.assign(timestamps, timestamps + nnz);
    else c->timestamps.assign(nnz, 0);
    c->own_views();
    *out = c;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_compressed_from_csr: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_compressed_from_csr: unknown C++ exception"); }

sbr_status sbr_compressed_borrow_csr(const uint64_t* user_pointers, const uint64_t* item_ids, const uint64_t* timestamps,
                                     size_t num_users, size_t num_items, sbr_compressed** out) {
    if (!out || !user_pointers) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (num_items > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "num_items must fit in 32 bits");
    if (user_pointers[0] != 0) return fail(SBR_ERR_INVALID_ARGUMENT, "user_pointers[0] must be 0");
    for (size_t u = 0; u < num_users; ++u)
        if (user_pointers[u + 1] < user_pointers[u]) return fail(SBR_ERR_INVALID_ARGUMENT, "user_pointers must be non-decreasing");
    if (user_pointers[num_users] && !item_ids) return fail(SBR_ERR_INVALID_ARGUMENT, "null item_ids");
    sbr_compressed* c = new (std::nothrow) sbr_compressed();
    if (!c) return fail(SBR_ERR_INVALID_ARGUMENT, "out of memory");
    c->num_users = num_users; c->num_items = num_items;
    c->up_ = user_pointers; c->ii_ = item_ids; c->tt_ = timestamps; c->nnz_ = user_pointers[num_users];
    *out = c;
    return SBR_OK;
}

size_t sbr_compressed_num_users(const sbr_compressed* c) { return c ? c->num_users : 0; }
size_t sbr_compressed_num_items(const sbr_compressed* c) { return c ? c->num_items : 0; }
size_t sbr_compressed_len(const sbr_compressed* c) { return c ? c->nnz_ : 0; }

sbr_status sbr_compressed_borrow(const sbr_compressed* c, const uint64_t** user_pointers, const uint64_t** item_ids,
                                 const uint64_t** timestamps) {
    if (!c) return fail(SBR_ERR_INVALID_ARGUMENT, "null handle");
    if (user_pointers) *user_pointers = c->up_;
    if (item_ids) *item_ids = c->ii_;
    if (timestamps) *timestamps = c->tt_;   /* NULL for a borrowed CSR without timestamps */
    return SBR_OK;
}

sbr_status sbr_compressed_user_chunks(const sbr_compressed* c, size_t user_id, size_t chunk_size, uint64_t* starts,
                                      uint64_t* lens, size_t cap, size_t* n) try {
    if (!c || !n) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (user_id >= c->num_users) return fail(SBR_ERR_INVALID_ARGUMENT, "user id out of range");  // data.rs:278-280
    if (chunk_size == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "chunk_size must be > 0");
    const size_t len = c->up_[user_id + 1] - c->up_[user_id];
    size_t idx = 0, k = 0;
    while (idx < len) {  // data.rs:406-432
        const size_t mod = (len - idx) % chunk_size;
        const size_t cs = mod == 0 ? chunk_size : mod;
        if (k < cap) { if (starts) starts[k] = idx; if (lens) lens[k] = cs; }
        idx += cs; ++k;
    }
    *n = k;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_compressed_user_chunks: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_compressed_user_chunks: unknown C++ exception"); }

sbr_status sbr_compressed_upload(sbr_compressed* c) {
    if (!c) return fail(SBR_ERR_INVALID_ARGUMENT, "null handle");
    sbr_status s = require_device();
    if (s) return s;
    return ensure_uploaded(c, nullptr);
}

void sbr_compressed_free(sbr_compressed* c) { delete c; }

// ----------------------------------------------------------------------------------------- hyperparameters --
static sbr_hyperparameters* hyper_new(int model, size_t num_items, size_t max_sequence_length) {
    sbr_hyperparameters* h = new (std::nothrow) sbr_hyperparameters();
    if (!h) return nullptr;
    h->model = model; h->num_items = num_items; h->max_sequence_length = max_sequence_length;
    // lstm.rs:67 seeds from thread_rng(); here: wall clock mixed with the address (call sbr_hyper_from_seed to pin)
    uint64_t s = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count() ^ (uint64_t)(uintptr_t)h;
    for (int i = 0; i < 16; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; h->seed[i] = (uint8_t)(s >> 56); }
    return h;
}
sbr_hyperparameters* sbr_lstm_hyperparameters_new(size_t n, size_t t) { return hyper_new(MODEL_LSTM, n, t); }
sbr_hyperparameters* sbr_ewma_hyperparameters_new(size_t n, size_t t) { return hyper_new(MODEL_EWMA, n, t); }

#define HYPER_SETTER(name, type, field, check)                                                      \
    sbr_status name(sbr_hyperparameters* h, type v) {                                               \
        if (!h) return fail(SBR_ERR_INVALID_ARGUMENT, "null hyperparameters");                      \
        if (!(check)) return fail(SBR_ERR_INVALID_ARGUMENT, #name ": value out of range");          \
        h->field = v;                                                                               \
        return SBR_OK;                                                                              \
    }
HYPER_SETTER(sbr_hyper_learning_rate, float, learning_rate, std::isfinite(v))
HYPER_SETTER(sbr_hyper_l2_penalty, float, l2_penalty, std::isfinite(v))
HYPER_SETTER(sbr_hyper_embedding_dim, size_t, embedding_dim, v > 0)
HYPER_SETTER(sbr_hyper_num_epochs, size_t, num_epochs, true)
HYPER_SETTER(sbr_hyper_loss, sbr_loss, loss, v >= 0 && v <= 2)
HYPER_SETTER(sbr_hyper_num_threads, size_t, num_threads, true)
HYPER_SETTER(sbr_hyper_parallelism, sbr_parallelism, parallelism, v >= 0 && v <= 1)
HYPER_SETTER(sbr_hyper_optimizer, sbr_optimizer, optimizer, v >= 0 && v <= 1)

sbr_status sbr_hyper_lstm_variant(sbr_hyperparameters* h, sbr_lstm_variant v) {
    if (!h) return fail(SBR_ERR_INVALID_ARGUMENT, "null hyperparameters");
    if (h->model != MODEL_LSTM) return fail(SBR_ERR_INVALID_ARGUMENT, "lstm_variant applies to LSTM hyperparameters only");
    if (v < 0 || v > 1) return fail(SBR_ERR_INVALID_ARGUMENT, "bad lstm variant");
    h->lstm_variant = v;
    return SBR_OK;
}
sbr_status sbr_hyper_exact_arithmetic(sbr_hyperparameters* h, int on) {
    if (!h) return fail(SBR_ERR_INVALID_ARGUMENT, "null hyperparameters");
    h->exact = on != 0;
    return SBR_OK;
}
sbr_status sbr_hyper_from_seed(sbr_hyperparameters* h, const uint8_t seed[16]) {
    if (!h || !seed) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    std::memcpy(h->seed, seed, 16);
    return SBR_OK;
}
void sbr_hyper_free(sbr_hyperparameters* h) { delete h; }

sbr_status sbr_hyper_build(sbr_hyperparameters* hp, sbr_model** out) try {
    if (!hp || !out) { delete hp; return fail(SBR_ERR_INVALID_ARGUMENT, "null argument"); }
    sbr_hyperparameters h = *hp;
    delete hp;  // build(self) consumes
    if (h.num_items == 0 || h.num_items > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "num_items must be in [1, 2^32)");
    if (h.max_sequence_length < 1 || h.max_sequence_length > 65535) return fail(SBR_ERR_INVALID_ARGUMENT, "max_sequence_length must be in [1, 65535]");
    sbr_status s = require_device();
    if (s) return s;
    sbr_model* m = new (std::nothrow) sbr_model();
    if (!m) return fail(SBR_ERR_INVALID_ARGUMENT, "out of memory");
    m->h = h;
    ModelDev& d = m->dev;
    d.model = h.model; d.variant = h.lstm_variant; d.loss = h.loss; d.opt = h.optimizer;
    d.N = (uint32_t)h.num_items; d.D = (int)h.embedding_dim; d.T = (int)h.max_sequence_length;
    d.S = h.optimizer == SBR_OPTIMIZER_ADAM ? 3 : 2;
    d.lr = h.learning_rate; d.l2 = h.l2_penalty; d.exact = h.exact;
    d.hbm_resident = (size_t)d.N * d.S * d.D * sizeof(float) > (size_t)96 << 20;
    d.ndense = h.model == MODEL_LSTM ? (size_t)2 * d.D * 4 * d.D + 4 * d.D : (size_t)d.D;
    const char* why = nullptr;
    if (!train_supported(d, &why)) { delete m; return fail(SBR_ERR_UNSUPPORTED, why); }
    xs_from_seed(m->rng, h.seed);
    auto bail = [&](cudaError_t e, const char* what) { sbr_status r = cuda_fail(e, what); delete m; return r; };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    // item table + optimizer state, row-sharded by id % G (G = 1: one shard)
    const int G = h.shard_world > 1 ? h.shard_world : h.virtual_shards;
    d.gmask = (uint32_t)G - 1;
    d.gshift = 0; while ((1 << d.gshift) < G) ++d.gshift;
    d.own_shard = h.shard_world > 1 ? h.shard_rank : -1;
    m->attached = h.shard_world <= 1;
    for (int g = 0; g < G; ++g) {
        if (h.shard_world > 1 && g != h.shard_rank) continue;   // peers' shards are mapped by sbr_model_ipc_attach
        const size_t rows = ((size_t)d.N + G - 1 - g) / G;       // ids g, g+G, g+2G, ...
        if ((e = cudaMalloc(&d.Es[g], std::max<size_t>(rows, 1) * rec_floats(d) * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc item table");
        m->own[g] = true;
    }
    if ((e = cudaMalloc(&d.dense, 3 * d.ndense * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc dense parameters");
    m->own_dense = d.dense;
    // parameter init (lstm.rs:174-186): embeddings N(0, 1/D) in HBM, biases 0, alpha 0, LSTM U(+-1/sqrt(D)) [wyrm-recalled]
    const uint64_t seed64 = xs_next_u64(m->rng);
    if ((e = launch_init_embeddings(d, seed64, m->stream)) != cudaSuccess) return bail(e, "init embeddings");
    std::vector<float> dense(3 * d.ndense, 0.0f);
    if (h.model == MODEL_LSTM) {
        const double a = 1.0 / std::sqrt((double)d.D);
        for (size_t i = 0; i < d.ndense; ++i) {
            const double u = ((double)(xs_next_u64(m->rng) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
            dense[i] = (float)((2.0 * u - 1.0) * a);
        }
    }
    if ((e = cudaMemcpyAsync(d.dense, dense.data(), dense.size() * sizeof(float), cudaMemcpyHostToDevice, m->stream)) != cudaSuccess) return bail(e, "upload dense");
    if ((e = cudaStreamSynchronize(m->stream)) != cudaSuccess) return bail(e, "build sync");
    *out = m;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_hyper_build: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_hyper_build: unknown C++ exception"); }

// --------------------------------------------------------------------------------------------- multi-GPU ----
sbr_status sbr_hyper_shard(sbr_hyperparameters* h, int rank, int world) {
    if (!h) return fail(SBR_ERR_INVALID_ARGUMENT, "null hyperparameters");
    if (!(world == 1 || world == 2 || world == 4 || world == 8) || rank < 0 || rank >= world)
        return fail(SBR_ERR_INVALID_ARGUMENT, "world must be 1, 2, 4 or 8 and 0 <= rank < world");
    h->shard_rank = rank; h->shard_world = world; h->virtual_shards = 1;
    return SBR_OK;
}
sbr_status sbr_hyper_virtual_shards(sbr_hyperparameters* h, int g) {
    if (!h) return fail(SBR_ERR_INVALID_ARGUMENT, "null hyperparameters");
    if (!(g == 1 || g == 2 || g == 4 || g == 8)) return fail(SBR_ERR_INVALID_ARGUMENT, "shard count must be 1, 2, 4 or 8");
    h->virtual_shards = g; h->shard_rank = 0; h->shard_world = 1;
    return SBR_OK;
}

sbr_status sbr_dist_unique_id(uint8_t out[128]) {
    if (!out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::string why;
    const NcclApi* nc = nccl_api(&why);
    if (!nc) return fail(SBR_ERR_NCCL, why);
    ncclUniqueId id;
    ncclResult_t r = nc->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(SBR_ERR_NCCL, std::string("ncclGetUniqueId: ") + nc->GetErrorString(r));
    std::memcpy(out, &id, 128);
    return SBR_OK;
}
sbr_status sbr_dist_init(int rank, int world, const uint8_t id_bytes[128]) {
    if (!id_bytes || world < 1 || rank < 0 || rank >= world) return fail(SBR_ERR_INVALID_ARGUMENT, "bad rank / world / id");
    sbr_status s = require_device();
    if (s) return s;
    std::string why;
    const NcclApi* nc = nccl_api(&why);
    if (!nc) return fail(SBR_ERR_NCCL, why);
    if (g_comm) { nc->CommDestroy(g_comm); g_comm = nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, 128);
    ncclResult_t r = nc->CommInitRank(&g_comm, world, id, rank);
    if (r != ncclSuccess) { g_comm = nullptr; return fail(SBR_ERR_NCCL, std::string("ncclCommInitRank: ") + nc->GetErrorString(r)); }
    g_rank = rank; g_world = world;
    return SBR_OK;
}
void sbr_dist_finalize(void) {
    if (g_comm) { if (const NcclApi* nc = nccl_api(nullptr)) nc->CommDestroy(g_comm); g_comm = nullptr; }
    g_rank = 0; g_world = 1;
}

}  // extern "C"
namespace {
// replicas: what this rank changed since the common starting point / the starting point plus everybody's changes
__global__ void replica_delta_kernel(const float* __restrict__ a, size_t na, const float* __restrict__ b, size_t nb, const float* __restrict__ snap,
                                     float* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < na + nb; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (i < na ? a[i] : b[i - na]) - snap[i];
}
__global__ void replica_apply_kernel(float* __restrict__ a, size_t na, float* __restrict__ b, size_t nb, float* __restrict__ snap,
                                     const float* __restrict__ sum) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < na + nb; i += (size_t)gridDim.x * blockDim.x) {
        const float v = snap[i] + sum[i];
        snap[i] = v;
        if (i < na) a[i] = v; else b[i - na] = v;
    }
}
}  // namespace
extern "C" {

// Data-parallel replicas of a small (L2-resident) model: every rank trains its own users on a full replica; this call makes
// the replicas identical again by adding up what each of them changed -- the deltas of every parameter AND of the optimizer
// state (item records, dense weights) since the last call are summed over the ranks with one ncclAllReduce on device buffers
// and applied to the common starting point.  The first call only records the starting point (the replicas must be identical
// then: same seed / same checkpoint).  Needs sbr_dist_init; the model must not be row-sharded.
sbr_status sbr_model_replica_sync(sbr_model* m, size_t* bytes_reduced) try {
    if (!m) return fail(SBR_ERR_INVALID_ARGUMENT, "null model");
    sbr_status s = require_device();
    if (s) return s;
    if (m->h.shard_world > 1 || m->dev.gmask != 0) return fail(SBR_ERR_INVALID_ARGUMENT, "replica sync needs an unsharded model");
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    const size_t na = (size_t)m->dev.N * rec_floats(m->dev), nb = 3 * m->dev.ndense, n = na + nb;
    if (bytes_reduced) *bytes_reduced = 0;
    if (!m->replica_snap || m->replica_n != n) {
        if (m->replica_snap) { cudaFree(m->replica_snap); cudaFree(m->replica_delta); m->replica_snap = m->replica_delta = nullptr; }
        CU(cudaMalloc(&m->replica_snap, n * sizeof(float)));
        CU(cudaMalloc(&m->replica_delta, n * sizeof(float)));
        m->replica_n = n;
        CU(cudaMemcpyAsync(m->replica_snap, m->dev.Es[0], na * sizeof(float), cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(m->replica_snap + na, m->dev.dense, nb * sizeof(float), cudaMemcpyDeviceToDevice, st));
        CU(cudaStreamSynchronize(st));
        return SBR_OK;
    }
    if (g_world <= 1) return SBR_OK;
    if (!g_comm) return fail(SBR_ERR_NCCL, "sbr_model_replica_sync needs sbr_dist_init");
    std::string why;
    const NcclApi* nc = nccl_api(&why);
    if (!nc) return fail(SBR_ERR_NCCL, why);
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    replica_delta_kernel<<<blocks, 256, 0, st>>>(m->dev.Es[0], na, m->dev.dense, nb, m->replica_snap, m->replica_delta);
    ncclResult_t r = nc->AllReduce(m->replica_delta, m->replica_delta, n, ncclFloat, ncclSum, g_comm, st);
    if (r != ncclSuccess) return fail(SBR_ERR_NCCL, std::string("ncclAllReduce: ") + nc->GetErrorString(r));
    replica_apply_kernel<<<blocks, 256, 0, st>>>(m->dev.Es[0], na, m->dev.dense, nb, m->replica_snap, m->replica_delta);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    if (bytes_reduced) *bytes_reduced = n * sizeof(float);
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_replica_sync: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_replica_sync: unknown C++ exception"); }

size_t sbr_model_ipc_handle_size(void) { return 2 * sizeof(cudaIpcMemHandle_t); }

sbr_status sbr_model_ipc_export(const sbr_model* m, void* out) {
    if (!m || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    sbr_status s = require_device();
    if (s) return s;
    const int r = m->h.shard_rank;
    if (m->h.shard_world <= 1) return fail(SBR_ERR_INVALID_ARGUMENT, "model was not built with sbr_hyper_shard");
    cudaIpcMemHandle_t* hs = reinterpret_cast<cudaIpcMemHandle_t*>(out);
    CU(cudaIpcGetMemHandle(&hs[0], m->dev.Es[r]));
    CU(cudaIpcGetMemHandle(&hs[1], m->own_dense));
    return SBR_OK;
}

sbr_status sbr_model_ipc_attach(sbr_model* m, const void* all_handles) {
    if (!m || !all_handles) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    sbr_status s = require_device();
    if (s) return s;
    const int W = m->h.shard_world, r = m->h.shard_rank;
    if (W <= 1) return fail(SBR_ERR_INVALID_ARGUMENT, "model was not built with sbr_hyper_shard");
    if (m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "peers already attached");
    std::lock_guard<std::mutex> lk(m->mu);
    const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int g = 0; g < W; ++g) {
        if (g == r) continue;
        void* pe = nullptr;
        CU(cudaIpcOpenMemHandle(&pe, hs[2 * g + 0], cudaIpcMemLazyEnablePeerAccess));
        m->dev.Es[g] = static_cast<float*>(pe);
    }
    // Asynchronous: the dense parameters (LSTM weights / alpha) live on rank 0 and everyone updates them Hogwild.
    // Synchronous: every rank keeps its own replica (identical on all ranks: same all-reduced gradient every round).
    if (r != 0 && m->h.parallelism != SBR_PARALLELISM_SYNCHRONOUS) {
        void* pd = nullptr;
        CU(cudaIpcOpenMemHandle(&pd, hs[1], cudaIpcMemLazyEnablePeerAccess));
        m->dev.dense = static_cast<float*>(pd);
    }
    m->attached = true;
    return SBR_OK;
}

size_t sbr_model_embedding_dim(const sbr_model* m) { return m ? (size_t)m->dev.D : 0; }
size_t sbr_model_num_items(const sbr_model* m) { return m ? (size_t)m->dev.N : 0; }
void sbr_model_free(sbr_model* m) { if (m) { cudaSetDevice(g_device); delete m; } }

sbr_status sbr_model_parameter_len(const sbr_model* m, const char* name, size_t* len) {
    if (!m || !name || !len) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    ParamRef r; sbr_status s = resolve_param(m, name, &r);
    if (s) return s;
    *len = r.len;
    return SBR_OK;
}

static sbr_status param_io(sbr_model* m, const char* name, float* host, size_t len, bool set) {
    if (!m || !name || !host) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    sbr_status s = require_device();
    if (s) return s;
    ParamRef r; s = resolve_param(m, name, &r);
    if (s) return s;
    if (len != r.len) return fail(SBR_ERR_INVALID_ARGUMENT, "parameter length mismatch");
    if (!set && r.kind != 2 && !m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "sharded model: call sbr_model_ipc_attach before reading the table");
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    if (r.kind == 2) {
        float* p = m->dev.dense + (size_t)r.slot * m->dev.ndense + r.off;
        if (set) CU(cudaMemcpyAsync(p, host, len * sizeof(float), cudaMemcpyHostToDevice, st));
        else CU(cudaMemcpyAsync(host, p, len * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return SBR_OK;
    }
    float* tmp = nullptr;
    CU(cudaMalloc(&tmp, len * sizeof(float)));
    cudaError_t e = cudaSuccess;
    if (set) {
        e = cudaMemcpyAsync(tmp, host, len * sizeof(float), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = r.kind == 0 ? launch_pack_rows(m->dev, r.slot, tmp, st) : launch_pack_bias(m->dev, r.slot, tmp, st);
    } else {
        e = r.kind == 0 ? launch_unpack_rows(m->dev, r.slot, tmp, st) : launch_unpack_bias(m->dev, r.slot, tmp, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(host, tmp, len * sizeof(float), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) return cuda_fail(e, "parameter transfer");
    return SBR_OK;
}

sbr_status sbr_model_get_parameter(const sbr_model* m, const char* name, float* out, size_t len) {
    return param_io(const_cast<sbr_model*>(m), name, out, len, false);
}
sbr_status sbr_model_set_parameter(sbr_model* m, const char* name, const float* data, size_t len) {
    return param_io(m, name, const_cast<float*>(data), len, true);
}
sbr_status sbr_model_get_num_updates(const sbr_model* m, uint64_t* out) {
    if (!m || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    *out = m->num_updates;
    return SBR_OK;
}
sbr_status sbr_model_set_num_updates(sbr_model* m, uint64_t v) {
    if (!m) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    m->num_updates = v;
    return SBR_OK;
}

sbr_status sbr_model_get_rng_state(const sbr_model* m, uint32_t out[4]) {
    if (!m || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    out[0] = m->rng.x; out[1] = m->rng.y; out[2] = m->rng.z; out[3] = m->rng.w;
    return SBR_OK;
}
sbr_status sbr_model_set_rng_state(sbr_model* m, const uint32_t state[4]) {
    if (!m || !state) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if ((state[0] | state[1] | state[2] | state[3]) == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "xorshift state must not be all zero");
    m->rng.x = state[0]; m->rng.y = state[1]; m->rng.z = state[2]; m->rng.w = state[3];
    return SBR_OK;
}

// ------------------------------------------------------------------------- hyperparameter values / random ----
static void fill_hyper_values(const sbr_hyperparameters& h, sbr_hyper_values* v) {
    std::memset(v, 0, sizeof(*v));
    v->model = h.model; v->lstm_variant = h.lstm_variant; v->loss = h.loss; v->optimizer = h.optimizer;
    v->parallelism = h.parallelism; v->exact_arithmetic = h.exact;
    v->num_items = h.num_items; v->max_sequence_length = h.max_sequence_length; v->embedding_dim = h.embedding_dim;
    v->num_threads = h.num_threads; v->num_epochs = h.num_epochs;
    v->learning_rate = h.learning_rate; v->l2_penalty = h.l2_penalty;
    std::memcpy(v->seed, h.seed, 16);
}
sbr_status sbr_hyper_get_values(const sbr_hyperparameters* h, sbr_hyper_values* out) {
    if (!h || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    fill_hyper_values(*h, out);
    return SBR_OK;
}
sbr_status sbr_model_get_hyper_values(const sbr_model* m, sbr_hyper_values* out) {
    if (!m || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    fill_hyper_values(m->h, out);
    return SBR_OK;
}

// rand 0.5 Uniform<f32>::sample [rand-recalled]: value in [1, 2) from the top 23 bits of one u32, then * scale + offset
static float xs_uniform_f32(XorShift& r, float low, float high) {
    const float scale = high - low, offset = low - scale;
    const uint32_t bits = (xs_next_u32(r) >> 9) | 0x3f800000u;
    float v12; std::memcpy(&v12, &bits, 4);
    return v12 * scale + offset;
}
static sbr_hyperparameters* hyper_random(int model, size_t num_items, uint32_t st[4]) {
    if (!st || (st[0] | st[1] | st[2] | st[3]) == 0) { g_err = "xorshift state must not be null / all zero"; return nullptr; }
    sbr_hyperparameters* h = hyper_new(model, num_items, 0);
    if (!h) return nullptr;
    XorShift r; r.x = st[0]; r.y = st[1]; r.z = st[2]; r.w = st[3];
    auto upow2 = [&](uint64_t lo, uint64_t hi) { return (size_t)1 << (lo + xs_gen_below(r, hi - lo)); };   // 2_usize.pow(Uniform::new(lo, hi))
    auto coin = [&]() { return xs_uniform_f32(r, 0.0f, 1.0f) < 0.5f; };
    h->max_sequence_length = upow2(4, 8);                                   // lstm.rs:144
    h->embedding_dim = upow2(4, 8);                                         // :145
    h->learning_rate = std::pow(10.0f, xs_uniform_f32(r, -3.0f, 0.5f));     // :146
    h->l2_penalty = std::pow(10.0f, xs_uniform_f32(r, -7.0f, -3.0f));       // :147
    h->loss = coin() ? SBR_LOSS_BPR : SBR_LOSS_HINGE;                       // :148-152
    h->optimizer = coin() ? SBR_OPTIMIZER_ADAM : SBR_OPTIMIZER_ADAGRAD;     // :153-157
    if (model == MODEL_LSTM) h->lstm_variant = coin() ? SBR_LSTM_NORMAL : SBR_LSTM_COUPLED;   // :158-162 (absent in ewma.rs)
    h->parallelism = coin() ? SBR_PARALLELISM_ASYNCHRONOUS : SBR_PARALLELISM_SYNCHRONOUS;     // :163-167
    // :168 rng: from thread_rng() -- hyper_new seeded it from the clock
    const uint64_t host_threads = std::max(1u, std::thread::hardware_concurrency());
    h->num_threads = 1 + (size_t)xs_gen_below(r, host_threads);            // :169 Uniform::new(1, rayon::current_num_threads() + 1)
    h->num_epochs = upow2(3, 7);                                            // :170
    st[0] = r.x; st[1] = r.y; st[2] = r.z; st[3] = r.w;
    return h;
}
sbr_hyperparameters* sbr_lstm_hyperparameters_random(size_t num_items, uint32_t rng_state[4]) { return hyper_random(MODEL_LSTM, num_items, rng_state); }
sbr_hyperparameters* sbr_ewma_hyperparameters_random(size_t num_items, uint32_t rng_state[4]) { return hyper_random(MODEL_EWMA, num_items, rng_state); }

// ------------------------------------------------------------------------------------------- checkpoint ----
namespace {
constexpr char kCkptMagic[8] = {'S', 'B', 'R', 'B', '2', '0', '0', '\0'};
struct CkptBlob { char name[32]; uint64_t len, off; };
static_assert(sizeof(sbr_hyper_values) == 88, "checkpoint layout documents an 88-byte sbr_hyper_values");
static_assert(sizeof(CkptBlob) == 48, "checkpoint blob directory entry is 48 bytes");

std::vector<std::string> ckpt_blob_names(const sbr_hyperparameters& h) {
    std::vector<std::string> base = {"item_embeddings", "item_biases"};
    if (h.model == MODEL_LSTM) { base.push_back("lstm_weights"); base.push_back("lstm_biases"); } else base.push_back("alpha");
    std::vector<std::string> out;
    for (const std::string& b : base) {
        out.push_back(b); out.push_back(b + ".s1");
        if (h.optimizer == SBR_OPTIMIZER_ADAM) out.push_back(b + ".s2");
    }
    return out;
}
struct FileCloser { FILE* f; ~FileCloser() { if (f) fclose(f); } };
}  // namespace

sbr_status sbr_model_save(const sbr_model* m, const char* path) try {
    if (!m || !path) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    sbr_status s = require_device();
    if (s) return s;
    const std::vector<std::string> names = ckpt_blob_names(m->h);
    std::vector<CkptBlob> dir(names.size());
    const uint64_t header = (128 + 8 + names.size() * sizeof(CkptBlob) + 63) / 64 * 64;
    uint64_t off = header;
    for (size_t i = 0; i < names.size(); ++i) {
        std::memset(&dir[i], 0, sizeof(CkptBlob));
        std::snprintf(dir[i].name, sizeof(dir[i].name), "%s", names[i].c_str());
        size_t len = 0;
        if ((s = sbr_model_parameter_len(m, names[i].c_str(), &len))) return s;
        dir[i].len = len; dir[i].off = off; off += (uint64_t)len * 4;
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SBR_ERR_INVALID_ARGUMENT, std::string("cannot open for writing: ") + path);
    FileCloser fc{f};
    std::vector<uint8_t> head(header, 0);
    std::memcpy(head.data(), kCkptMagic, 8);
    const uint32_t version = 1, hb = (uint32_t)header, nb = (uint32_t)names.size();
    std::memcpy(head.data() + 8, &version, 4); std::memcpy(head.data() + 12, &hb, 4);
    sbr_hyper_values hv; fill_hyper_values(m->h, &hv);
    std::memcpy(head.data() + 16, &hv, sizeof(hv));
    const uint32_t rs[4] = {m->rng.x, m->rng.y, m->rng.z, m->rng.w};
    std::memcpy(head.data() + 104, rs, 16);
    std::memcpy(head.data() + 120, &m->num_updates, 8);
    std::memcpy(head.data() + 128, &nb, 4);
    std::memcpy(head.data() + 136, dir.data(), dir.size() * sizeof(CkptBlob));
    if (fwrite(head.data(), 1, head.size(), f) != head.size()) return fail(SBR_ERR_INVALID_ARGUMENT, "short write (header)");
    std::vector<float> buf;
    for (size_t i = 0; i < names.size(); ++i) {
        buf.resize(dir[i].len);
        if ((s = sbr_model_get_parameter(m, names[i].c_str(), buf.data(), buf.size()))) return s;
        if (fwrite(buf.data(), 4, buf.size(), f) != buf.size()) return fail(SBR_ERR_INVALID_ARGUMENT, "short write (blob)");
    }
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_save: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_save: unknown C++ exception"); }

static sbr_status ckpt_read(const char* path, sbr_hyper_values* hv, uint32_t rs[4], uint64_t* num_updates, std::vector<CkptBlob>* dir, FILE** fout) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(SBR_ERR_INVALID_ARGUMENT, std::string("cannot open: ") + path);
    uint8_t fixed[136];
    uint32_t version = 0, hb = 0, nb = 0;
    bool ok = fread(fixed, 1, sizeof(fixed), f) == sizeof(fixed) && std::memcmp(fixed, kCkptMagic, 8) == 0;
    if (ok) { std::memcpy(&version, fixed + 8, 4); std::memcpy(&hb, fixed + 12, 4); std::memcpy(&nb, fixed + 128, 4); ok = version == 1 && nb <= 16 && hb >= 136 + nb * sizeof(CkptBlob); }
    if (ok) { dir->resize(nb); ok = fread(dir->data(), sizeof(CkptBlob), nb, f) == nb; }
    if (!ok) { fclose(f); return fail(SBR_ERR_INVALID_ARGUMENT, "not a version-1 sbr_b200 checkpoint"); }
    std::memcpy(hv, fixed + 16, sizeof(*hv)); std::memcpy(rs, fixed + 104, 16); std::memcpy(num_updates, fixed + 120, 8);
    for (CkptBlob& b : *dir) b.name[31] = 0;
    *fout = f;
    return SBR_OK;
}

static sbr_status ckpt_apply(sbr_model* m, const sbr_hyper_values& hv, const uint32_t rs[4], uint64_t num_updates, const std::vector<CkptBlob>& dir, FILE* f) {
    if (hv.model != m->h.model || hv.num_items != m->h.num_items || hv.embedding_dim != m->h.embedding_dim || hv.optimizer != m->h.optimizer)
        return fail(SBR_ERR_INVALID_ARGUMENT, "checkpoint does not match the model (model kind / num_items / embedding_dim / optimizer)");
    std::vector<float> buf;
    for (const CkptBlob& b : dir) {
        size_t len = 0;
        sbr_status s = sbr_model_parameter_len(m, b.name, &len);
        if (s) return s;
        if (len != b.len) return fail(SBR_ERR_INVALID_ARGUMENT, std::string("checkpoint blob has the wrong length: ") + b.name);
        buf.resize(len);
        if (fseek(f, (long)b.off, SEEK_SET) != 0 || fread(buf.data(), 4, len, f) != len) return fail(SBR_ERR_INVALID_ARGUMENT, std::string("truncated checkpoint blob: ") + b.name);
        if ((s = sbr_model_set_parameter(m, b.name, buf.data(), len))) return s;
    }
    if ((rs[0] | rs[1] | rs[2] | rs[3]) == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "checkpoint holds an all-zero rng state");
    m->rng.x = rs[0]; m->rng.y = rs[1]; m->rng.z = rs[2]; m->rng.w = rs[3];
    m->num_updates = num_updates;
    return SBR_OK;
}

sbr_status sbr_model_restore(sbr_model* m, const char* path) try {
    if (!m || !path) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    sbr_status s = require_device();
    if (s) return s;
    sbr_hyper_values hv; uint32_t rs[4]; uint64_t nu = 0; std::vector<CkptBlob> dir; FILE* f = nullptr;
    if ((s = ckpt_read(path, &hv, rs, &nu, &dir, &f))) return s;
    FileCloser fc{f};
    return ckpt_apply(m, hv, rs, nu, dir, f);
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_restore: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_restore: unknown C++ exception"); }

sbr_status sbr_model_load(const char* path, sbr_model** out) try {
    if (!path || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    sbr_status s = require_device();
    if (s) return s;
    sbr_hyper_values hv; uint32_t rs[4]; uint64_t nu = 0; std::vector<CkptBlob> dir; FILE* f = nullptr;
    if ((s = ckpt_read(path, &hv, rs, &nu, &dir, &f))) return s;
    FileCloser fc{f};
    if (hv.model != MODEL_LSTM && hv.model != MODEL_EWMA) return fail(SBR_ERR_INVALID_ARGUMENT, "checkpoint: unknown model kind");
    sbr_hyperparameters* h = hyper_new(hv.model, hv.num_items, hv.max_sequence_length);
    if (!h) return fail(SBR_ERR_INVALID_ARGUMENT, "out of memory");
    h->embedding_dim = hv.embedding_dim; h->learning_rate = hv.learning_rate; h->l2_penalty = hv.l2_penalty;
    h->lstm_variant = hv.lstm_variant; h->loss = hv.loss; h->optimizer = hv.optimizer; h->parallelism = hv.parallelism;
    h->num_threads = hv.num_threads; h->num_epochs = hv.num_epochs; h->exact = hv.exact_arithmetic;
    std::memcpy(h->seed, hv.seed, 16);
    sbr_model* m = nullptr;
    if ((s = sbr_hyper_build(h, &m))) return s;
    if ((s = ckpt_apply(m, hv, rs, nu, dir, f))) { const std::string keep = g_err; sbr_model_free(m); g_err = keep; return s; }
    *out = m;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_load: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_load: unknown C++ exception"); }

// ---------------------------------------------------------------------------------- data.rs:54-88 splits ----
// SipHash-2-4 (Aumasson & Bernstein) of one 8-byte little-endian word: what siphasher's `write_usize` + `finish` compute
static inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
static uint64_t siphash24_u64(uint64_t k0, uint64_t k1, uint64_t m) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL, v1 = k1 ^ 0x646f72616e646f6dULL, v2 = k0 ^ 0x6c7967656e657261ULL, v3 = k1 ^ 0x7465646279746573ULL;
    auto round = [&]() {
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
    };
    v3 ^= m; round(); round(); v0 ^= m;
    const uint64_t b = (uint64_t)8 << 56;          // total length 8 bytes, no tail bytes
    v3 ^= b; round(); round(); v0 ^= b;
    v2 ^= 0xff; round(); round(); round(); round();
    return v0 ^ v1 ^ v2 ^ v3;
}
// data.rs:69-88 user_based_split: is_train(x) = siphash(key_0, key_1, user) % 100_000 > (test_fraction * 100_000) as u64,
// the keys being two Uniform<u64>[0, MAX) draws from the caller's rng.  Host-only (data preparation of the tests / examples).
sbr_status sbr_user_based_split(const uint64_t* user_ids, size_t nnz, uint32_t rng_state[4], float test_fraction, uint8_t* out_is_train) try {
    if ((nnz && (!user_ids || !out_is_train)) || !rng_state) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if ((rng_state[0] | rng_state[1] | rng_state[2] | rng_state[3]) == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "xorshift state must not be all zero");
    XorShift rng; rng.x = rng_state[0]; rng.y = rng_state[1]; rng.z = rng_state[2]; rng.w = rng_state[3];
    const uint64_t denominator = 100000;
    const uint64_t cutoff = (uint64_t)(test_fraction * (float)denominator);
    const uint64_t k0 = xs_gen_below(rng, UINT64_MAX), k1 = xs_gen_below(rng, UINT64_MAX);
    for (size_t i = 0; i < nnz; ++i) out_is_train[i] = (siphash24_u64(k0, k1, user_ids[i]) % denominator) > cutoff;
    rng_state[0] = rng.x; rng_state[1] = rng.y; rng_state[2] = rng.z; rng_state[3] = rng.w;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_user_based_split: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_user_based_split: unknown C++ exception"); }
// data.rs:54-64 train_test_split: interactions.shuffle(rng) (Fisher-Yates from the top), then the FIRST
// (test_fraction * len) as usize shuffled interactions are the test set.  perm[k] = original index of the k-th
// shuffled interaction; perm[0 .. *num_test) is test, the rest is train.
sbr_status sbr_train_test_split(size_t nnz, uint32_t rng_state[4], float test_fraction, uint64_t* perm, size_t* num_test) try {
    if ((nnz && !perm) || !rng_state || !num_test) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if ((rng_state[0] | rng_state[1] | rng_state[2] | rng_state[3]) == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "xorshift state must not be all zero");
    XorShift rng; rng.x = rng_state[0]; rng.y = rng_state[1]; rng.z = rng_state[2]; rng.w = rng_state[3];
    for (size_t i = 0; i < nnz; ++i) perm[i] = i;
    for (size_t i = nnz; i >= 2;) { i -= 1; const size_t j = (size_t)xs_gen_below(rng, (uint64_t)i + 1); std::swap(perm[i], perm[j]); }
    *num_test = (size_t)(test_fraction * (float)nnz);
    rng_state[0] = rng.x; rng_state[1] = rng.y; rng_state[2] = rng.z; rng_state[3] = rng.w;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_train_test_split: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_train_test_split: unknown C++ exception"); }

// host-only view of the schedule that fit() builds (no device needed): lets CPU-only CI pin the chunker / filter /
// master shuffle against the oracle
sbr_status sbr_host_schedule(const sbr_compressed* c, size_t max_sequence_length, uint32_t rng_state[4], uint64_t* starts, uint32_t* lens,
                             uint32_t* order, size_t cap, size_t* nsub) try {
    if (!c || !rng_state || !nsub) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if ((rng_state[0] | rng_state[1] | rng_state[2] | rng_state[3]) == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "xorshift state must not be all zero");
    XorShift rng; rng.x = rng_state[0]; rng.y = rng_state[1]; rng.z = rng_state[2]; rng.w = rng_state[3];
    std::vector<uint64_t> st; std::vector<uint32_t> ln, od;
    sbr_status s = host_schedule(c, max_sequence_length, rng, st, ln, od);
    if (s) return s;
    // fit() itself only counts the sub-sequences on the host (the chunks are built on the device): the count it would use
    // must be the number of chunks this one-thread reference pass produced
    if (host_count_subsequences(c, max_sequence_length) != st.size()) return fail(SBR_ERR_INVALID_ARGUMENT, "internal: threaded sub-sequence count disagrees with the chunker");
    *nsub = st.size();
    if (starts && lens && order && cap >= st.size()) {
        std::memcpy(starts, st.data(), st.size() * sizeof(uint64_t));
        std::memcpy(lens, ln.data(), ln.size() * sizeof(uint32_t));
        std::memcpy(order, od.data(), od.size() * sizeof(uint32_t));
        rng_state[0] = rng.x; rng_state[1] = rng.y; rng_state[2] = rng.z; rng_state[3] = rng.w;
    }
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_host_schedule: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_host_schedule: unknown C++ exception"); }

sbr_status sbr_host_master_schedule(uint32_t rng_state[4], size_t nsub, size_t partitions, size_t threads, uint32_t* order, uint64_t* keys) try {
    if (!rng_state || !order || (partitions && !keys)) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if ((rng_state[0] | rng_state[1] | rng_state[2] | rng_state[3]) == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "xorshift state must not be all zero");
    if (nsub > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "too many sub-sequences");
    XorShift rng; rng.x = rng_state[0]; rng.y = rng_state[1]; rng.z = rng_state[2]; rng.w = rng_state[3];
    std::vector<XorShift> rngs(partitions);
    master_schedule(rng, order, nsub, partitions, rngs.data(), keys, threads);
    rng_state[0] = rng.x; rng_state[1] = rng.y; rng_state[2] = rng.z; rng_state[3] = rng.w;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_host_master_schedule: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_host_master_schedule: unknown C++ exception"); }

// --------------------------------------------------------------------------------------------------- fit ----
// Order of business (everything the device can do is on the model's stream, the host never waits before the end):
//   main thread : user_ptr -> HBM, device chunker (data_prep.cu) enqueued          | count sub-sequences (threads), master shuffle
//   upload thread:                                  id stream -> HBM (narrowed)     |   of their indices into a pinned buffer,
//   then: order / rngs / keys -> HBM from the pinned buffer; the plan is ready when the stream is.                partition rngs
sbr_status sbr_fit_plan_create(sbr_model* m, const sbr_compressed* c, sbr_fit_plan** out) try {
    if (!m || !c || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    sbr_status s = require_device();
    if (s) return s;
    if (c->num_items > m->dev.N) return fail(SBR_ERR_INVALID_ARGUMENT, "interactions.num_items exceeds the model's num_items");
    const double t0 = now_ms();
    const size_t T = (size_t)m->dev.T;
    if (T == 0) return fail(SBR_ERR_INVALID_ARGUMENT, "max_sequence_length must be positive");
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    // sequence_model.rs:76-81: how many sub-sequences survive the len > 2 filter (the chunks themselves: device, below)
    const size_t nsub = host_count_subsequences(c, T);
    if (nsub == 0) return fail(SBR_ERR_NO_INTERACTIONS, "No interactions were supplied.");  // :86-88
    if (nsub > 0xffffffffull) return fail(SBR_ERR_INVALID_ARGUMENT, "too many sub-sequences");
    // :90-98 partitions
    size_t P = m->h.num_threads;
    bool cold_bounded = false;
    if (P == 0) {
        const size_t autoP = (size_t)train_auto_partitions(m->dev, device_info().sms);
        P = std::min(autoP, std::max<size_t>(1, nsub / 16));
        // the D = 32 tile kernels take whole tiles of 128 partitions (2 tiles per CTA): round down so that real data
        // (nsub / 16 is almost never a multiple of 128) still runs on them
        if (m->dev.D == 32 && !m->dev.exact) { if (P >= 256) P -= P % 256; else if (P >= 128) P = 128; }
        if (m->dev.model == MODEL_LSTM && m->dev.D > 32 && !m->dev.exact && P >= 128) P -= P % 128;
        // Cold-start bound.  An LSTM that starts from random parameters cannot absorb thousands of concurrent
        // sequences per item row: before the first feedback every row takes hundreds of coherent Adagrad steps (scale-invariant:
        // lr-sized whatever the gradient's size), embeddings and gate weights blow up together and the cell saturates for
        // good (measured: 1,683 items, 37,888 partitions: MRR stays at the untrained level for 64 epochs; <= 4,096 partitions
        // learn in one epoch; a warm model trains fine at 37,888).  Until the model has seen ~100 steps per item the automatic
        // choice therefore stays below 2.5 partitions per item; afterwards it fills the device.
        const size_t bound = std::max<size_t>(256, (size_t)(2.5 * (double)m->dev.N) / 256 * 256);
        if (m->dev.model == MODEL_LSTM && m->num_updates < 100ull * m->dev.N && P > bound) { P = bound; cold_bounded = true; }
    }
    if (P > nsub) return fail(SBR_ERR_INVALID_ARGUMENT, "num_threads exceeds the number of sub-sequences (the reference panics in chunks_mut(0))");
    const size_t n = nsub / P;  // :91, remainder dropped by the zip at :94-96

    sbr_fit_plan* pl = new (std::nothrow) sbr_fit_plan();
    if (!pl) return fail(SBR_ERR_INVALID_ARGUMENT, "out of memory");
    pl->model = m; pl->nsub = nsub; pl->P = P; pl->n = n; pl->cold_bounded = cold_bounded;
    {   // which engine: LSTM under Parallelism::Synchronous (any width), and the wide LSTMs with many partitions, run in rounds
        const char* why = nullptr;
        const bool can = m->dev.model == MODEL_LSTM && m->h.shard_world <= 1 && batch_lstm_supported(m->dev, (uint32_t)P, &why);
        const bool sync_lstm = m->dev.model == MODEL_LSTM && m->h.parallelism == SBR_PARALLELISM_SYNCHRONOUS && P > 1;
        if (sync_lstm && !can) { delete pl; return fail(SBR_ERR_INVALID_ARGUMENT, std::string("Parallelism::Synchronous LSTM fit: ") + (why ? why : "one process / unsharded table only")); }
        pl->use_batch = can && (sync_lstm || (m->dev.D > 32 && !m->dev.exact && P >= 128));
    }
    std::string up_err;
    std::future<sbr_status> up;
    struct Joiner { std::future<sbr_status>& f; ~Joiner() { if (f.valid()) f.wait(); } } joiner{up};   // never outlive the upload thread
    auto bail = [&](sbr_status r) { if (up.valid()) up.wait(); delete pl; return r; };
#define CUP(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return bail(cuda_fail(e__, #expr)); } while (0)
    for (cudaEvent_t* e : {&pl->ev0, &pl->ev1, &pl->evk0, &pl->evk1}) CUP(cudaEventCreate(e));
    CUP(cudaEventRecord(pl->ev0, st));
    size_t h2d = 0;
    PlanDev& d = pl->dev;
    // ---- device: user_ptr, then the chunker ----
    s = ensure_user_ptr(c, st, &h2d);
    if (s) return bail(s);
    CUP(pool_alloc(&pl->d_seq_start, nsub * sizeof(uint64_t)));
    CUP(pool_alloc(&pl->d_seq_len, nsub * sizeof(uint32_t)));
    {
        uint32_t* d_counts = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0;
        CUP(device_schedule(nullptr, c->num_users, T, nullptr, nullptr, &tmp_bytes, nsub, nullptr, nullptr, st));
        CUP(pool_alloc(&d_counts, (c->num_users + 1) * sizeof(uint32_t)));
        pl->tmp[0] = d_counts;
        CUP(pool_alloc(&d_tmp, std::max<size_t>(tmp_bytes, 256)));
        pl->tmp[1] = d_tmp;
        CUP(device_schedule(c->d_user_ptr, c->num_users, T, d_counts, d_tmp, &tmp_bytes, nsub, pl->d_seq_start, pl->d_seq_len, st));
    }
    // ---- upload thread: the id stream (a no-op when the CSR is already resident) ----
    const double tu0 = now_ms();
    double upload_ms = 0.0;
    size_t id_bytes = 0;
    up = std::async(std::launch::async, [&]() {
        cudaSetDevice(g_device);
        sbr_status r = ensure_ids(c, st, &id_bytes);
        if (r) up_err = g_err;
        upload_ms = now_ms() - tu0;
        return r;
    });
    // ---- host: master shuffle of the indices (:84) + per-partition rngs (:97) into the model's pinned staging area ----
    const size_t stage_bytes = nsub * sizeof(uint32_t) + P * (sizeof(XorShift) + sizeof(uint64_t)) + 64;
    if (m->h_stage_cap < stage_bytes) {
        if (m->h_stage) { cudaFreeHost(m->h_stage); m->h_stage = nullptr; m->h_stage_cap = 0; }
        CUP(cudaHostAlloc(&m->h_stage, stage_bytes + (stage_bytes >> 2), cudaHostAllocDefault));
        m->h_stage_cap = stage_bytes + (stage_bytes >> 2);
    }
    uint32_t* h_order = reinterpret_cast<uint32_t*>(m->h_stage);
    XorShift* h_rngs = reinterpret_cast<XorShift*>(m->h_stage + ((nsub * sizeof(uint32_t) + 15) & ~(size_t)15));
    uint64_t* h_keys = reinterpret_cast<uint64_t*>(h_rngs + P);
    master_schedule(m->rng, h_order, nsub, P, h_rngs, h_keys, std::min<size_t>(4, std::max(1u, std::thread::hardware_concurrency())));
    pl->stats.host_prepare_ms = now_ms() - t0;

    CUP(pool_alloc(&d.order, P * n * sizeof(uint32_t)));
    CUP(pool_alloc(&d.rng, P * sizeof(XorShift)));
    CUP(pool_alloc(&d.keys, P * sizeof(uint64_t)));
    CUP(pool_alloc(&d.step_ctr, P * sizeof(uint64_t)));
    CUP(pool_alloc(&d.loss_acc, P * sizeof(float)));
    CUP(pool_alloc(&d.examples, P * sizeof(unsigned long long)));
    d.scratch_stride = pl->use_batch ? 32 : train_scratch_floats_per_warp(m->dev);
    {
        const size_t need = P * d.scratch_stride * sizeof(float);
        if (!m->scratch_busy) {   // the model keeps one grow-only scratch buffer across fit() calls
            if (m->scratch_cap < need) {
                if (m->scratch_cache) { cudaFree(m->scratch_cache); m->scratch_cache = nullptr; m->scratch_cap = 0; }
                CUP(cudaMalloc(&m->scratch_cache, need));
                m->scratch_cap = need;
            }
            d.scratch = m->scratch_cache; pl->scratch_borrowed = true; m->scratch_busy = true;
        } else CUP(cudaMalloc(&d.scratch, need));
    }
    s = up.get();   // the id copies are enqueued (and done); what follows queues behind them
    if (s) { g_err = up_err; return bail(s); }
    h2d += id_bytes;
    c->upload_bytes = id_bytes;
    pl->stats.upload_ms = upload_ms;
    d.item_ids = c->d_item_ids;
    CUP(cudaMemcpyAsync(d.order, h_order, P * n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUP(cudaMemcpyAsync(d.rng, h_rngs, P * sizeof(XorShift), cudaMemcpyHostToDevice, st));
    CUP(cudaMemcpyAsync(d.keys, h_keys, P * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUP(cudaMemsetAsync(d.step_ctr, 0, P * sizeof(uint64_t), st));
    CUP(cudaMemsetAsync(d.loss_acc, 0, P * sizeof(float), st));
    CUP(cudaMemsetAsync(d.examples, 0, P * sizeof(unsigned long long), st));
    CUP(cudaStreamSynchronize(st));   // the model's pinned staging area is free again; chunker scratch can go back to the pool
    g_pool.release(pl->tmp[0]); g_pool.release(pl->tmp[1]); pl->tmp[0] = pl->tmp[1] = nullptr;
    h2d += P * n * 4 + P * (sizeof(XorShift) + 8);
    d.seq_start = pl->d_seq_start; d.seq_len = pl->d_seq_len;
    d.n = (uint32_t)n; d.P = (uint32_t)P;
    d.neg_range = (uint32_t)c->num_items;  // :74 Uniform::new(0, interactions.num_items())
    pl->stats.h2d_bytes = h2d;
    pl->stats.partitions = P;
#undef CUP
    *out = pl;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_fit_plan_create: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_fit_plan_create: unknown C++ exception"); }

sbr_status sbr_fit_plan_run(sbr_fit_plan* pl, float* loss_out) try {
    if (!pl) return fail(SBR_ERR_INVALID_ARGUMENT, "null plan");
    sbr_status s = require_device();
    if (s) return s;
    sbr_model* m = pl->model;
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    const size_t P = pl->P;
    pl->dev.epochs = pl->epochs_override >= 0 ? pl->epochs_override : (int)m->h.num_epochs;
    pl->dev.adam_t0 = m->num_updates;
    CU(cudaMemsetAsync(pl->dev.loss_acc, 0, P * sizeof(float), st));
    CU(cudaMemsetAsync(pl->dev.examples, 0, P * sizeof(unsigned long long), st));
    int launches = 0;
    CU(cudaEventRecord(pl->evk0, st));
    const char* why = nullptr;
    uint64_t sync_rounds = 0;
    const bool sync_mode = m->h.parallelism == SBR_PARALLELISM_SYNCHRONOUS && (P > 1 || g_world > 1) && sync_supported(m->dev, &why) &&
                           (int)(m->dev.gmask + 1) == (m->h.shard_world > 1 ? m->h.shard_world : 1);
    if (m->h.parallelism == SBR_PARALLELISM_SYNCHRONOUS && (P > 1 || g_world > 1) && !sync_mode && !pl->use_batch)
        // never fall back to the Hogwild schedule behind the caller's back (the reference default is Synchronous, lstm.rs:66)
        return fail(SBR_ERR_INVALID_ARGUMENT, std::string("Parallelism::Synchronous is not available for this configuration: ") +
                                                   (why ? why : "the item table's sharding does not match the process group") +
                                                   " (use Parallelism::Asynchronous or num_threads = 1)");
    if (pl->dev.epochs > 0 && pl->use_batch) {
        if (!pl->batch) pl->batch = batch_buffers_new();
        std::string err;
        const int rc = run_batch_lstm(m->dev, pl->dev, *pl->batch, m->num_updates, device_info().sms, st, &launches, &sync_rounds, &err);
        if (rc) return fail(SBR_ERR_CUDA, err);
        std::snprintf(pl->stats.kernel, sizeof(pl->stats.kernel), "bl_fwd/bl_dz/bl_dw_kernel D=%d (batched tcgen05 LSTM engine)", m->dev.D);
    } else if (pl->dev.epochs > 0 && sync_mode) {
        // Parallelism::Synchronous: round-synchronous schedule with an explicit row exchange (sync_engine.cu)
        const int world = m->h.shard_world > 1 ? m->h.shard_world : 1;
        if (world > 1 && (!g_comm || g_world != world || g_rank != m->h.shard_rank))
            return fail(SBR_ERR_NCCL, "synchronous multi-GPU fit needs sbr_dist_init(rank, world, id) matching sbr_hyper_shard");
        if (!pl->sync) pl->sync = sync_buffers_new();
        std::string err;
        // every rank steps its OWN replica of the dense parameters (after sbr_model_ipc_attach dev.dense of rank != 0 points
        // at rank 0's buffer, which all ranks would otherwise step once each per round)
        ModelDev md = m->dev; md.dense = m->own_dense;
        const int rc = run_sync_ewma(md, pl->dev, *pl->sync, g_comm, m->h.shard_rank, world, m->num_updates, st, &launches, &sync_rounds, &err);
        if (rc) return fail(rc == 2 ? SBR_ERR_NCCL : rc == 3 ? SBR_ERR_INVALID_ARGUMENT : SBR_ERR_CUDA, err);
        std::snprintf(pl->stats.kernel, sizeof(pl->stats.kernel), "sync_ewma_compute_kernel<%d> (round-synchronous%s)", m->dev.D,
                      world == 1 ? "" : sync_buffers_copy_engine(pl->sync) ? ", p2p copy engines" : sync_buffers_p2p(pl->sync) ? ", p2p stores" : ", nccl");
    } else if (pl->dev.epochs > 0) {
        if (!m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "sharded model: call sbr_model_ipc_attach before an asynchronous fit");
        cudaError_t e;
        const char* kname = "";
        launches = launch_train(m->dev, pl->dev, device_info().sms, st, &e, &kname);
        std::snprintf(pl->stats.kernel, sizeof(pl->stats.kernel), "%s", kname);
        if (e != cudaSuccess) return cuda_fail(e, "launch_train");
    }
    CU(cudaEventRecord(pl->evk1, st));
    std::vector<float> loss(P); std::vector<unsigned long long> ex(P);
    CU(cudaMemcpyAsync(loss.data(), pl->dev.loss_acc, P * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ex.data(), pl->dev.examples, P * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(pl->ev1, st));
    CU(cudaStreamSynchronize(st));
    float total = 0.0f; uint64_t tsteps = 0;
    for (size_t p = 0; p < P; ++p) { total += loss[p] / (1.0f + (float)ex[p]); tsteps += ex[p]; }  // :173-175
    const uint64_t steps = ((sync_mode || pl->use_batch) && pl->dev.epochs > 0) ? (uint64_t)P * sync_rounds : (uint64_t)P * pl->n * (uint64_t)pl->dev.epochs;
    m->num_updates += steps;
    float kms = 0.0f, tms = 0.0f;
    CU(cudaEventElapsedTime(&kms, pl->evk0, pl->evk1));
    CU(cudaEventElapsedTime(&tms, pl->ev0, pl->ev1));
    pl->stats.steps = steps; pl->stats.timesteps = tsteps;
    pl->stats.kernel_launches = (uint64_t)launches;
    pl->stats.d2h_bytes = P * 12;
    pl->stats.train_kernel_ms = kms; pl->stats.total_device_ms = tms;
    if (loss_out) *loss_out = total;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_fit_plan_run: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_fit_plan_run: unknown C++ exception"); }

sbr_status sbr_fit_plan_stats(const sbr_fit_plan* p, sbr_fit_stats* out) {
    if (!p || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    *out = p->stats;
    return SBR_OK;
}
sbr_status sbr_fit_plan_read_schedule(const sbr_fit_plan* p, uint64_t* starts, uint32_t* lens, uint32_t* order, size_t cap, size_t* nsub,
                                      size_t* norder) try {
    if (!p || !nsub || !norder) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    *nsub = p->nsub; *norder = p->P * p->n;
    if (!starts || !lens || !order || cap < p->nsub) return SBR_OK;
    sbr_status s = require_device();
    if (s) return s;
    std::lock_guard<std::mutex> lk(p->model->mu);
    cudaStream_t st = p->model->stream;
    CU(cudaMemcpyAsync(starts, p->d_seq_start, p->nsub * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(lens, p->d_seq_len, p->nsub * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(order, p->dev.order, p->P * p->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_fit_plan_read_schedule: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_fit_plan_read_schedule: unknown C++ exception"); }
void sbr_fit_plan_free(sbr_fit_plan* p) { if (p) { cudaSetDevice(g_device); delete p; } }

sbr_status sbr_model_fit(sbr_model* m, const sbr_compressed* c, float* loss_out) {
    sbr_fit_plan* pl = nullptr;
    sbr_status s = sbr_fit_plan_create(m, c, &pl);
    if (s) return s;
    // automatic partitions on a cold model: the first epoch runs at the bounded count, the remaining epochs
    // on a second schedule at whatever the (now warmer) model allows; the returned loss is the second schedule's
    const bool split = pl->cold_bounded && m->h.num_epochs > 1;
    if (split) pl->epochs_override = 1;
    s = sbr_fit_plan_run(pl, loss_out);
    if (s == SBR_OK) m->last = pl->stats;
    sbr_fit_plan_free(pl);
    if (s == SBR_OK && split) {
        pl = nullptr;
        s = sbr_fit_plan_create(m, c, &pl);
        if (s) return s;
        pl->epochs_override = (int)m->h.num_epochs - 1;
        s = sbr_fit_plan_run(pl, loss_out);
        if (s == SBR_OK) m->last = pl->stats;
        sbr_fit_plan_free(pl);
    }
    return s;
}

sbr_status sbr_model_set_num_threads(sbr_model* m, size_t num_threads) {
    if (!m) return fail(SBR_ERR_INVALID_ARGUMENT, "null model");
    std::lock_guard<std::mutex> lk(m->mu);
    m->h.num_threads = num_threads;
    return SBR_OK;
}

sbr_status sbr_model_last_fit_stats(const sbr_model* m, sbr_fit_stats* out) {
    if (!m || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    *out = m->last;
    return SBR_OK;
}

// --------------------------------------------------------------------------------------------- inference ----
sbr_status sbr_model_user_representations(const sbr_model* m, const uint64_t* ptr, const uint64_t* item_ids, size_t num_users,
                                          float* out) try {
    if (!m || !ptr || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "sharded model: call sbr_model_ipc_attach before using it");
    sbr_status s = require_device();
    if (s) return s;
    if (num_users == 0) return SBR_OK;
    const size_t nnz = ptr[num_users];
    if (nnz && !item_ids) return fail(SBR_ERR_INVALID_ARGUMENT, "null item_ids");
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    uint32_t* d_ids = nullptr; uint64_t* d_ptr = nullptr; float* d_out = nullptr;
    const size_t D = m->dev.D;
    CU(cudaMalloc(&d_ids, std::max<size_t>(nnz, 1) * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_ptr, (num_users + 1) * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, num_users * D * sizeof(float));
    if (e == cudaSuccess) {
        s = upload_ids_u32(item_ids, nnz, d_ids, st, m->dev.N, nullptr);
        if (s == SBR_OK) {
            e = cudaMemcpyAsync(d_ptr, ptr, (num_users + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = launch_user_representations(m->dev, d_ptr, d_ids, num_users, d_out, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, num_users * D * sizeof(float), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
    }
    cudaFree(d_ids); cudaFree(d_ptr); cudaFree(d_out);
    if (s) return s;
    if (e != cudaSuccess) return cuda_fail(e, "user_representations");
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_user_representations: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_user_representations: unknown C++ exception"); }

sbr_status sbr_model_user_representation(const sbr_model* m, const uint64_t* item_ids, size_t n, float* out) {
    const uint64_t ptr[2] = {0, (uint64_t)n};
    return sbr_model_user_representations(m, ptr, item_ids, 1, out);
}

sbr_status sbr_model_predict(const sbr_model* m, const float* user, const uint64_t* item_ids, size_t k, float* out) try {
    if (!m || !user || (k && (!item_ids || !out))) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "sharded model: call sbr_model_ipc_attach before using it");
    sbr_status s = require_device();
    if (s) return s;
    if (k == 0) return SBR_OK;
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    const size_t D = m->dev.D;
    uint32_t* d_ids = nullptr; float* d_user = nullptr; float* d_out = nullptr; int* d_flag = nullptr;
    CU(cudaMalloc(&d_ids, k * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_user, D * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, k * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_flag, sizeof(int));
    int flag = 0;
    if (e == cudaSuccess) {
        s = upload_ids_u32(item_ids, k, d_ids, st, m->dev.N, nullptr);
        if (s == SBR_OK) {
            e = cudaMemcpyAsync(d_user, user, D * sizeof(float), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, sizeof(int), st);
            if (e == cudaSuccess) e = launch_predict(m->dev, d_user, d_ids, k, d_out, d_flag, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, k * sizeof(float), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
    }
    cudaFree(d_ids); cudaFree(d_user); cudaFree(d_out); cudaFree(d_flag);
    if (s) return s;
    if (e != cudaSuccess) return cuda_fail(e, "predict");
    if (flag) return fail(SBR_ERR_INVALID_PREDICTION, "Invalid prediction value: non-finite or not a number.");
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_predict: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_predict: unknown C++ exception"); }

sbr_status sbr_model_mrr_score(const sbr_model* m, const sbr_compressed* test, float* out) try {
    if (!m || !test || !out) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "sharded model: call sbr_model_ipc_attach before using it");
    sbr_status s = require_device();
    if (s) return s;
    if (test->num_items > m->dev.N) return fail(SBR_ERR_INVALID_ARGUMENT, "test.num_items exceeds the model's num_items");
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    s = ensure_uploaded(test, st);
    if (s) return s;
    const size_t U = test->num_users;
    std::vector<float> rr(U, 0.0f);
    int flag = 0;
    if (U && test->d_item_ids) {
        float* d_rr = nullptr; int* d_flag = nullptr;
        CU(cudaMalloc(&d_rr, U * sizeof(float)));
        cudaError_t e = cudaMalloc(&d_flag, sizeof(int));
        if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, sizeof(int), st);
        if (e == cudaSuccess) e = launch_mrr(m->dev, (uint32_t)test->num_items, test->d_user_ptr, test->d_item_ids, U, d_rr, d_flag, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(rr.data(), d_rr, U * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_rr); cudaFree(d_flag);
        if (e != cudaSuccess) return cuda_fail(e, "mrr_score");
    }
    if (flag) return fail(SBR_ERR_INVALID_PREDICTION, "Invalid prediction value: non-finite or not a number.");
    float sum = 0.0f; size_t cnt = 0;  // evaluation.rs:47 mrrs.iter().sum::<f32>() / mrrs.len() as f32, user order
    for (size_t u = 0; u < U; ++u)
        if (test->up_[u + 1] - test->up_[u] >= 2) { sum += rr[u]; ++cnt; }
    *out = cnt ? sum / (float)cnt : NAN;
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_mrr_score: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_mrr_score: unknown C++ exception"); }

// Stand-alone embedding gather timed on the device (BASELINE.json metric "embed-gather HBM GB/s"): the ids are uploaded
// once, gather_rows_kernel runs 1 + iters times on the model's stream between CUDA events, *kernel_ms is the mean of the
// timed launches.  Algorithmic bytes per row: 4 D read + 4 D written + 4 (u32 id).  `out` may be NULL (nothing copied back).
sbr_status sbr_model_gather_rows_timed(const sbr_model* m, const uint64_t* item_ids, size_t n, float* out, int iters, double* kernel_ms) try {
    if (!m || !kernel_ms || iters < 1 || (n && !item_ids)) return fail(SBR_ERR_INVALID_ARGUMENT, "bad argument");
    if (!m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "sharded model: call sbr_model_ipc_attach before using it");
    sbr_status s = require_device();
    if (s) return s;
    *kernel_ms = 0.0;
    if (n == 0) return SBR_OK;
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    const size_t D = m->dev.D;
    uint32_t* d_ids = nullptr; float* d_out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CU(cudaMalloc(&d_ids, n * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_out, n * D * sizeof(float));
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) {
        s = upload_ids_u32(item_ids, n, d_ids, st, m->dev.N, nullptr);
        if (s == SBR_OK) {
            e = launch_gather_rows(m->dev, d_ids, n, d_out, st);   // warm-up
            if (e == cudaSuccess) e = cudaEventRecord(e0, st);
            for (int i = 0; i < iters && e == cudaSuccess; ++i) e = launch_gather_rows(m->dev, d_ids, n, d_out, st);
            if (e == cudaSuccess) e = cudaEventRecord(e1, st);
            if (e == cudaSuccess && out) e = cudaMemcpyAsync(out, d_out, n * D * sizeof(float), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            float ms = 0.0f;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
            *kernel_ms = (double)ms / iters;
        }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_ids); cudaFree(d_out);
    if (s) return s;
    if (e != cudaSuccess) return cuda_fail(e, "gather_rows_timed");
    return SBR_OK;
} catch (const std::exception& e) { return fail(SBR_ERR_INVALID_ARGUMENT, std::string("sbr_model_gather_rows_timed: ") + e.what()); } catch (...) { return fail(SBR_ERR_INVALID_ARGUMENT, "sbr_model_gather_rows_timed: unknown C++ exception"); }

sbr_status sbr_model_gather_rows(const sbr_model* m, const uint64_t* item_ids, size_t n, float* out) {
    if (!m || (n && (!item_ids || !out))) return fail(SBR_ERR_INVALID_ARGUMENT, "null argument");
    if (!m->attached) return fail(SBR_ERR_INVALID_ARGUMENT, "sharded model: call sbr_model_ipc_attach before using it");
    sbr_status s = require_device();
    if (s) return s;
    if (n == 0) return SBR_OK;
    std::lock_guard<std::mutex> lk(m->mu);
    cudaStream_t st = m->stream;
    const size_t D = m->dev.D;
    uint32_t* d_ids = nullptr; float* d_out = nullptr;
    CU(cudaMalloc(&d_ids, n * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_out, n * D * sizeof(float));
    if (e == cudaSuccess) {
        s = upload_ids_u32(item_ids, n, d_ids, st, m->dev.N, nullptr);
        if (s == SBR_OK) {
            e = launch_gather_rows(m->dev, d_ids, n, d_out, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * D * sizeof(float), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
    }
    cudaFree(d_ids); cudaFree(d_out);
    if (s) return s;
    if (e != cudaSuccess) return cuda_fail(e, "gather_rows");
    return SBR_OK;
}

}  // extern "C"
