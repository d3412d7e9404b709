// sync_engine.cu -- round-synchronous training with an explicit row exchange (Parallelism::Synchronous, mod.rs:39-40;
// barrier-coupled optimizer at sequence_model.rs:92,163-166) for the EWMA model, on 1..8 GPUs.
//
// A round = every partition ("thread") takes its next sub-sequence.  All gradients of a round are computed from the
// round-start parameters, then every recorded (row, gradient) entry is applied once, un-merged -- the reference's
// semantics when its threads meet at the optimizer barrier.  Because nothing reads the table between "gather" and
// "apply", the table traffic of a round can be made explicit, and that is what makes a catalogue that is row-sharded
// over several GPUs practical (BASELINE config C4: 50M items x 128):
//     requests (item ids)  --all-to-all-->  owners gather rows from their own HBM (local, coalesced)
//     rows                 --all-to-all-->  requesters run the fused forward/backward on the received rows
//     gradient rows        --all-to-all-->  owners apply the sparse Adagrad/Adam visits to their own shard
// Random access stays inside each GPU's HBM; NVLink only carries contiguous buffers (NCCL grouped send/recv).  Direct
// peer loads (the Hogwild path, kernels_train.cu) collapse once the imported range exceeds a few GB (measured: 8M
// items/2 GPUs 2.4M steps/s, 50M items/8 GPUs 0.65M steps/s -- fabric-side address translation thrashes).
// With one GPU the same kernels run without NCCL (the exchange is the identity).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <string>
#include <vector>

#include <cuda_bf16.h>

#include <algorithm>

#include "engine.h"
#include "nccl_dyn.h"
#include "tc_tile.cuh"

namespace sbr {

namespace {

constexpr uint32_t kInvalid = 0xffffffffu;

// thread per partition: epoch shuffle (thread_rng.shuffle(partition), sequence_model.rs:109)
__global__ void sync_shuffle_kernel(PlanDev pl) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pl.P) return;
    XorShift rng = pl.rng[p];
    uint32_t* ord = pl.order + (size_t)p * pl.n;
    uint32_t i = pl.n;
    while (i >= 2) {
        i -= 1;
        const uint32_t j = (uint32_t)xs_gen_below(rng, (uint64_t)i + 1);
        const uint32_t a = ord[i], b = ord[j];
        ord[i] = b; ord[j] = a;
    }
    pl.rng[p] = rng;
}

// Requests of one round.  slot = (p*(T-1) + t)*3 + k, k = 0 input, 1 target, 2 negative (uniform draw; WARP's
// data-dependent resampling is not supported in this mode).
// req_ord = position of the entry in the reference's application order: partition (thread) major, then t descending, then
// E[neg], E[out], E[in] (the oracle's order); owners sort by (row, ord) so that entries naming the same row are applied
// sequentially in exactly that order -- consecutive timesteps always share a row (out_{t-1} == in_t).
__global__ void sync_request_kernel(ModelDev m, PlanDev pl, uint32_t it, uint64_t step_base, uint32_t part_base,
                                    uint32_t* __restrict__ req_id, uint32_t* __restrict__ req_ord) {
    const int Tm1 = m.T - 1;
    const size_t total = (size_t)pl.P * Tm1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t p = (uint32_t)(i / Tm1); const int t = (int)(i - (size_t)p * Tm1);
        const uint32_t sq = pl.order[(size_t)p * pl.n + it];
        const int Tn = (int)pl.seq_len[sq] - 1;
        uint32_t a = kInvalid, b = kInvalid, c = kInvalid;
        if (t < Tn) {
            const uint32_t* ids = pl.item_ids + pl.seq_start[sq];
            a = __ldg(ids + t); b = __ldg(ids + t + 1);
            c = draw_item(pl.keys[p], step_base + pl.step_ctr[p], (uint32_t)t, 0u, pl.neg_range);
        }
        req_id[3 * i] = a; req_id[3 * i + 1] = b; req_id[3 * i + 2] = c;
        const uint32_t base = ((part_base + p) << 17) | ((uint32_t)(Tm1 - 1 - t) << 2);
        req_ord[3 * i] = base | 2u; req_ord[3 * i + 1] = base | 1u; req_ord[3 * i + 2] = base | 0u;
    }
}

// bucket the requests by owner GPU (warp-aggregated): pass 0 counts, pass 1 scatters
__global__ void sync_bucket_kernel(ModelDev m, const uint32_t* __restrict__ req_id, size_t nslots, int pass,
                                   unsigned int* counts /*[G]*/, unsigned int* cursor /*[G]*/, uint32_t* __restrict__ send_row,
                                   uint32_t* __restrict__ pos_of_slot, const uint32_t* __restrict__ req_ord, uint32_t* __restrict__ send_ord) {
    const int G = (int)m.gmask + 1, lane = threadIdx.x & 31;
    __shared__ unsigned int off[8];
    if (pass == 1) {
        if (threadIdx.x < 8) { unsigned int o = 0; for (int g = 0; g < (int)threadIdx.x && g < G; ++g) o += counts[g]; off[threadIdx.x] = o; }
        __syncthreads();
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nround = (nslots + stride - 1) / stride * stride;  // keep warps converged for the ballots
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        const uint32_t id = i < nslots ? req_id[i] : kInvalid;
        const int owner = id == kInvalid ? -1 : (int)(id & m.gmask);
        for (int g = 0; g < G; ++g) {
            const unsigned mask = __ballot_sync(kFull, owner == g);
            if (!mask) continue;
            const int leader = __ffs(mask) - 1;
            unsigned int base = 0;
            if (lane == leader) base = atomicAdd(pass == 0 ? &counts[g] : &cursor[g], (unsigned int)__popc(mask));
            base = __shfl_sync(kFull, base, leader);
            if (pass == 1 && owner == g) {
                const unsigned int pos = off[g] + base + (unsigned int)__popc(mask & ((1u << lane) - 1));
                send_row[pos] = id >> m.gshift;
                send_ord[pos] = req_ord[i];
                pos_of_slot[i] = pos;
            }
        }
        if (pass == 1 && i < nslots && owner < 0) pos_of_slot[i] = kInvalid;
    }
}

// owner side: copy the requested rows (weights only) and biases out of the local shard, warp per request
template <int D>
__global__ void __launch_bounds__(256) sync_gather_kernel(ModelDev m, int self, const uint32_t* __restrict__ rows, size_t n,
                                                          float* __restrict__ out_rows, float* __restrict__ out_bias) {
    constexpr int V = VecOf<D>::V;
    const int lane = threadIdx.x & 31;
    for (size_t j = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n; j += (size_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t r = rows[j];
        float w[V];
        row_load_cg<D>(shard_item_rec(m, self, r), lane, w);
        vec_store<D>(out_rows + j * D, lane, w);
        if (lane == 0) out_bias[j] = __ldcg(reinterpret_cast<const float*>(shard_bias_rec(m, self, r)));
    }
}

// requester side: fused EWMA forward/backward of one sub-sequence per warp on the received rows (ewma.rs:266-352)
template <int D>
__global__ void __launch_bounds__(256) sync_ewma_compute_kernel(ModelDev m, PlanDev pl, uint32_t it, float* __restrict__ rows,
                                                                float* __restrict__ biases, const uint32_t* __restrict__ pos_of_slot,
                                                                float* __restrict__ grads, float* __restrict__ bgrads,
                                                                float* __restrict__ dalpha_sum, uint32_t* __restrict__ own_rows) {
    constexpr int V = VecOf<D>::V;
    const int lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= pl.P) return;
    const int Tm1 = m.T - 1;
    const uint32_t sq = pl.order[(size_t)p * pl.n + it];
    const int Tn = (int)pl.seq_len[sq] - 1;
    float* S_ = pl.scratch + (size_t)p * pl.scratch_stride;          // [T][D] states
    float* G_ = S_ + (size_t)m.T * D;                                // [T] loss gradients
    const uint32_t* pos = pos_of_slot + (size_t)p * Tm1 * 3;
    float al[V], a[V], s[V];
    row_load_cg<D>(m.dense, lane, al);
#pragma unroll
    for (int v = 0; v < V; ++v) { a[v] = sigmoidf_(al[v]); s[v] = 0.0f; }
    float loss_seq = 0.0f;
    for (int t = 0; t < Tn; ++t) {
        const uint32_t px = pos[3 * t], pp = pos[3 * t + 1], pq = pos[3 * t + 2];
        float x[V], pv[V], qv[V];
        vec_load<D>(rows + (size_t)px * D, lane, x); vec_load<D>(rows + (size_t)pp * D, lane, pv); vec_load<D>(rows + (size_t)pq * D, lane, qv);
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] = t == 0 ? x[v] : a[v] * s[v] + (1.0f - a[v]) * x[v];
        vec_store<D>(S_ + (size_t)t * D, lane, s);
        const float posv = warp_dot<D>(s, pv) + biases[pp];
        float ngs = warp_dot<D>(s, qv) + biases[pq];
        if (m.loss == 2 && own_rows) {
            // WARP (sequence_model.rs:47-68) on one GPU: the requested row is candidate 0; further candidates are read from the
            // table itself (nothing writes it before the round's apply stage) and the accepted one replaces the request
            const uint64_t key = pl.keys[p], step = pl.step_ctr[p];
            for (int j = 1; j < 5 && !(1.0f - posv + ngs > 0.0f); ++j) {
                const uint32_t cand = draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range);
                row_load_cg<D>(item_rec(m, cand), lane, qv);
                const float bq = __ldcg(reinterpret_cast<const float*>(bias_rec(m, cand)));
                ngs = warp_dot<D>(s, qv) + bq;
                vec_store<D>(rows + (size_t)pq * D, lane, qv);
                if (lane == 0) { biases[pq] = bq; own_rows[pq] = cand; }
                __syncwarp();
            }
        }
        float l, g;
        if (m.loss == 0) { const float sg = sigmoidf_(ngs - posv); l = sg; g = sg * (1.0f - sg); }
        else { const float vv = 1.0f + ngs - posv; l = vv > 0.0f ? vv : 0.0f; g = vv > 0.0f ? 1.0f : 0.0f; }
        loss_seq += l;
        if (lane == 0) G_[t] = g;
    }
    __syncwarp();
    float ds[V], da[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { ds[v] = 0.0f; da[v] = 0.0f; }
    for (int t = Tn - 1; t >= 0; --t) {
        const uint32_t px = pos[3 * t], pp = pos[3 * t + 1], pq = pos[3 * t + 2];
        const float g = G_[t];
        float st[V], pv[V], qv[V], dh[V], dx[V], gn[V], gp[V];
        vec_load<D>(S_ + (size_t)t * D, lane, st);
        vec_load<D>(rows + (size_t)pp * D, lane, pv); vec_load<D>(rows + (size_t)pq * D, lane, qv);
#pragma unroll
        for (int v = 0; v < V; ++v) dh[v] = ds[v] + g * (qv[v] - pv[v]);
        if (t == 0) {
#pragma unroll
            for (int v = 0; v < V; ++v) { dx[v] = dh[v]; ds[v] = 0.0f; }
        } else {
            float sp[V], x[V];
            vec_load<D>(S_ + (size_t)(t - 1) * D, lane, sp);
            vec_load<D>(rows + (size_t)px * D, lane, x);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                dx[v] = (1.0f - a[v]) * dh[v];
                da[v] += dh[v] * (sp[v] - x[v]);
                ds[v] = a[v] * dh[v];
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) { gn[v] = g * st[v]; gp[v] = -g * st[v]; }
        vec_store<D>(grads + (size_t)pq * D, lane, gn);
        vec_store<D>(grads + (size_t)pp * D, lane, gp);
        vec_store<D>(grads + (size_t)px * D, lane, dx);
        if (lane == 0) { bgrads[pq] = g; bgrads[pp] = -g; bgrads[px] = __int_as_float(0x7fc00000); }  // NaN: no bias entry for inputs
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float dal = da[v] * a[v] * (1.0f - a[v]);
        if (D >= 32 || lane < D) atomicAdd(dalpha_sum + (D < 32 ? lane : lane * V + v), dal);
    }
    if (lane == 0) { pl.loss_acc[p] += loss_seq; pl.examples[p] += (unsigned long long)Tn; pl.step_ctr[p] += 1; }
}

__global__ void sync_keys_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ ords, size_t n,
                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        keys[j] = ((unsigned long long)rows[j] << 32) | ords[j];
        vals[j] = (uint32_t)j;
    }
}

// owner side: entries sorted by (row, application order); the warp that sees the first entry of a row applies that row's
// entries one after the other, un-merged (wyrm's sparse optimizer loop), different rows in parallel
template <int D>
__global__ void __launch_bounds__(256) sync_apply_kernel(ModelDev m, int self, const unsigned long long* __restrict__ keys,
                                                         const uint32_t* __restrict__ vals, const float* __restrict__ grads,
                                                         const float* __restrict__ bgrads, size_t n, OptCfg o) {
    constexpr int V = VecOf<D>::V;
    const int lane = threadIdx.x & 31;
    const size_t stride = (size_t)gridDim.x * (blockDim.x >> 5);
    const int rec_lines = (int)((rec_floats(m) * 4 + 127) / 128), grad_lines = (D * 4 + 127) / 128;
    for (size_t j = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n; j += stride) {
        // the warp's next entry: its record and gradient row travel towards L2 while this one is applied (one line per lane)
        if (j + stride < n) {
            const unsigned long long kn = keys[j + stride];
            const uint32_t rn = (uint32_t)(kn >> 32);
            if (rn != kInvalid) {
                if (lane < rec_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(shard_bias_rec(m, self, rn)) + lane * 128));
                else if (lane < rec_lines + grad_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(grads + (size_t)vals[j + stride] * D) + (lane - rec_lines) * 128));
            }
        }
        const uint32_t r = (uint32_t)(keys[j] >> 32);
        if (r == kInvalid) continue;                                 // unused slot (sorted to the end)
        if (j > 0 && (uint32_t)(keys[j - 1] >> 32) == r) continue;   // not the first entry of its row
        // the row's record is read once, takes the entries of its run one after the other in registers (same arithmetic
        // and order as one read-modify-write per entry), and is written back once
        float* rec = shard_item_rec(m, self, r);
        float w[V], s1[V], s2[V];
        row_load_cg<D>(rec, lane, w);
        row_load_cg<D>(rec + D, lane, s1);
        if (o.adam) row_load_cg<D>(rec + 2 * D, lane, s2);
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        bool bias_dirty = false;
        if (lane == 0) bq = __ldcg(shard_bias_rec(m, self, r));
        for (size_t e = j; e < n && (uint32_t)(keys[e] >> 32) == r; ++e) {
            const uint32_t src = vals[e];
            float g[V];
            vec_load<D>(grads + (size_t)src * D, lane, g);
            if (!o.adam) {
#pragma unroll
                for (int v = 0; v < V; ++v) adagrad_elem(w[v], s1[v], g[v], o.lr, o.l2);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) adam_elem(w[v], s1[v], s2[v], g[v], o);
            }
            if (lane == 0) {
                const float bg = bgrads[src];
                if (bg == bg) {
                    if (!o.adam) adagrad_elem(bq.x, bq.y, bg, o.lr, o.l2); else adam_elem(bq.x, bq.y, bq.z, bg, o);
                    bias_dirty = true;
                }
            }
        }
        row_store_cg<D>(rec, lane, w);
        row_store_cg<D>(rec + D, lane, s1);
        if (o.adam) row_store_cg<D>(rec + 2 * D, lane, s2);
        if (lane == 0 && bias_dirty) __stcg(shard_bias_rec(m, self, r), bq);
    }
}

__global__ void sync_dense_kernel(ModelDev m, float* gsum, OptCfg o) {
    const size_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.ndense) return;
    float w = m.dense[i], s1 = m.dense[m.ndense + i];
    if (o.adam) { float s2 = m.dense[2 * m.ndense + i]; adam_elem(w, s1, s2, gsum[i], o); m.dense[2 * m.ndense + i] = s2; }
    else adagrad_elem(w, s1, gsum[i], o.lr, o.l2);
    m.dense[i] = w; m.dense[m.ndense + i] = s1;
    gsum[i] = 0.0f;
}

#define SYNC_DISPATCH_D(D_, ...)                                 \
    switch (D_) {                                                \
        case 16: { constexpr int kD = 16; __VA_ARGS__; } break;  \
        case 32: { constexpr int kD = 32; __VA_ARGS__; } break;  \
        case 64: { constexpr int kD = 64; __VA_ARGS__; } break;  \
        case 128: { constexpr int kD = 128; __VA_ARGS__; } break;\
        case 256: { constexpr int kD = 256; __VA_ARGS__; } break;\
        default: break;                                          \
    }

struct Buf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    ~Buf() { if (p) cudaFree(p); }
};

}  // namespace

struct SyncBuffers {
    Buf req_id, req_ord, send_row, send_ord, pos_of_slot, counts, allcounts, recv_row, recv_ord, keys_in, keys_out, vals_in, vals_out, cub_tmp, rows_req, bias_req, rows_own, bias_own, grads_req, bgrads_req,
        grads_own, bgrads_own, dalpha;
    unsigned int* h_counts = nullptr;  // pinned [G*G + G]
    ~SyncBuffers() { if (h_counts) cudaFreeHost(h_counts); }
};

SyncBuffers* sync_buffers_new() { return new SyncBuffers(); }
void sync_buffers_free(SyncBuffers* b) { delete b; }

#define SCU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { *err = std::string(#expr) + ": " + cudaGetErrorString(e__); return 1; } } while (0)
#define SNC(expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) { *err = std::string(#expr) + ": " + NC->GetErrorString(r__); return 2; } } while (0)

#include "lstm_batch.cuh"

BatchBuffers* batch_buffers_new() { return new BatchBuffers(); }
void batch_buffers_free(BatchBuffers* b) { delete b; }

bool sync_supported(const ModelDev& m, const char** why) {
    if (m.model != MODEL_EWMA) { *why = "Parallelism::Synchronous with num_threads > 1 is implemented for the EWMA model only"; return false; }
    if (m.loss == 2 && m.gmask != 0) { *why = "Parallelism::Synchronous with WARP needs the whole item table on one GPU (candidates are resampled against it)"; return false; }
    return true;
}

size_t sync_scratch_floats_per_partition(const ModelDev& m) { return (((size_t)m.T * m.D + m.T) + 31) / 32 * 32; }

// returns 0 ok, 1 cuda error, 2 nccl error, 3 capacity error.  `comm` may be null when world == 1.
int run_sync_ewma(const ModelDev& m, PlanDev& pl, SyncBuffers& B, void* comm_v, int rank, int world, uint64_t num_updates,
                  cudaStream_t st, int* launches, uint64_t* rounds_out, std::string* err) {
    ncclComm_t comm = static_cast<ncclComm_t>(comm_v);
    const NcclApi* NC = nullptr;
    if (world > 1) { NC = nccl_api(err); if (!NC) return 2; }
    const int G = world, D = m.D, Tm1 = m.T - 1;
    const size_t nslots = (size_t)pl.P * Tm1 * 3;
    // Rows other ranks may request from this shard per round.  The usual load is ~nslots; a skewed id distribution (popular
    // ids in one residue class mod world) can send up to world * nslots rows to one owner, so the owner-side buffers GROW
    // when a round needs more (the stream is idle at that point: the counts were just read back).
    size_t cap_own = 0;
    auto ensure_own = [&](size_t need) -> int {
        if (need <= cap_own) return 0;
        cap_own = need;
        SCU(B.keys_in.ensure(cap_own * 8)); SCU(B.keys_out.ensure(cap_own * 8)); SCU(B.vals_in.ensure(cap_own * 4)); SCU(B.vals_out.ensure(cap_own * 4));
        size_t cub_bytes = 0;
        SCU(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<unsigned long long*>(nullptr), static_cast<unsigned long long*>(nullptr),
                                            static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), (int)cap_own, 0, 64, st));
        SCU(B.cub_tmp.ensure(cub_bytes));
        if (world > 1) {
            SCU(B.recv_row.ensure(cap_own * 4)); SCU(B.recv_ord.ensure(cap_own * 4)); SCU(B.rows_own.ensure(cap_own * D * 4)); SCU(B.bias_own.ensure(cap_own * 4));
            SCU(B.grads_own.ensure(cap_own * D * 4)); SCU(B.bgrads_own.ensure(cap_own * 4));
        }
        return 0;
    };
    SCU(B.req_id.ensure(nslots * 4)); SCU(B.send_row.ensure(nslots * 4)); SCU(B.pos_of_slot.ensure(nslots * 4));
    SCU(B.req_ord.ensure(nslots * 4)); SCU(B.send_ord.ensure(nslots * 4));
    if (int rc = ensure_own(world == 1 ? nslots : nslots * 3 / 2 + 4096)) return rc;
    SCU(B.counts.ensure(64)); SCU(B.allcounts.ensure(8 * 8 * 4));
    SCU(B.rows_req.ensure(nslots * D * 4)); SCU(B.bias_req.ensure(nslots * 4));
    SCU(B.grads_req.ensure(nslots * D * 4)); SCU(B.bgrads_req.ensure(nslots * 4));
    SCU(B.dalpha.ensure(m.ndense * 4));
    if (!B.h_counts) SCU(cudaHostAlloc(&B.h_counts, (8 * 8 + 8) * sizeof(unsigned int), cudaHostAllocDefault));
    SCU(cudaMemsetAsync(B.dalpha.p, 0, m.ndense * 4, st));
    unsigned int* counts = static_cast<unsigned int*>(B.counts.p);       // [0..8) counts, [8..16) cursors
    uint32_t* req_id = static_cast<uint32_t*>(B.req_id.p);
    uint32_t* send_row = static_cast<uint32_t*>(B.send_row.p);
    uint32_t* pos_of_slot = static_cast<uint32_t*>(B.pos_of_slot.p);
    uint32_t* req_ord = static_cast<uint32_t*>(B.req_ord.p);
    uint32_t* send_ord = static_cast<uint32_t*>(B.send_ord.p);
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;
    const int grid_p = (int)((pl.P + 7) / 8);
    uint64_t rounds_done = 0;

    auto exchange = [&](const void* sendbuf, const size_t* scnt, const size_t* soff, void* recvbuf, const size_t* rcnt, const size_t* roff,
                        size_t elem) -> ncclResult_t {
        ncclResult_t r = NC->GroupStart();
        if (r != ncclSuccess) return r;
        for (int g = 0; g < G && r == ncclSuccess; ++g) {
            if (scnt[g]) r = NC->Send(static_cast<const char*>(sendbuf) + soff[g] * elem, scnt[g] * elem, ncclChar, g, comm, st);
            if (r == ncclSuccess && rcnt[g]) r = NC->Recv(static_cast<char*>(recvbuf) + roff[g] * elem, rcnt[g] * elem, ncclChar, g, comm, st);
        }
        const ncclResult_t re = NC->GroupEnd();   // the group is always closed, also after a failed Send / Recv
        return r != ncclSuccess ? r : re;
    };

    // every rank must run the same number of rounds (collectives inside): the ranks agree on the smallest partition length
    uint32_t n_rounds = pl.n;
    if (world > 1) {
        B.h_counts[64] = pl.n;
        SCU(cudaMemcpyAsync(counts, &B.h_counts[64], 4, cudaMemcpyHostToDevice, st));
        SNC(NC->AllReduce(counts, counts, 1, ncclUint32, ncclMin, comm, st));
        SCU(cudaMemcpyAsync(&B.h_counts[64], counts, 4, cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        n_rounds = B.h_counts[64];
    }
    *rounds_out = (uint64_t)n_rounds * (uint64_t)pl.epochs;
    for (int ep = 0; ep < pl.epochs; ++ep) {
        sync_shuffle_kernel<<<(pl.P + 127) / 128, 128, 0, st>>>(pl);
        ++*launches;
        for (uint32_t it = 0; it < n_rounds; ++it, ++rounds_done) {
            // 1. requests + bucketing by owner
            SCU(cudaMemsetAsync(counts, 0, 64, st));
            sync_request_kernel<<<148 * 8, 256, 0, st>>>(m, pl, it, 0, (uint32_t)rank * pl.P, req_id, req_ord);
            sync_bucket_kernel<<<148 * 4, 256, 0, st>>>(m, req_id, nslots, 0, counts, counts + 8, send_row, pos_of_slot, req_ord, send_ord);
            sync_bucket_kernel<<<148 * 4, 256, 0, st>>>(m, req_id, nslots, 1, counts, counts + 8, send_row, pos_of_slot, req_ord, send_ord);
            *launches += 3;
            size_t scnt[8] = {0}, soff[8] = {0}, rcnt[8] = {0}, roff[8] = {0}, nown = 0;
            const uint32_t* own_rows = send_row; const uint32_t* own_ords = send_ord; float* rows_for_compute = nullptr; float* bias_for_compute = nullptr;
            if (world > 1) {
                SNC(NC->AllGather(counts, B.allcounts.p, 8, ncclUint32, comm, st));
                SCU(cudaMemcpyAsync(B.h_counts, B.allcounts.p, (size_t)G * 8 * 4, cudaMemcpyDeviceToHost, st));
                SCU(cudaStreamSynchronize(st));
                size_t so = 0, ro = 0;
                for (int g = 0; g < G; ++g) {
                    scnt[g] = B.h_counts[rank * 8 + g]; soff[g] = so; so += scnt[g];
                    rcnt[g] = B.h_counts[g * 8 + rank]; roff[g] = ro; ro += rcnt[g];
                }
                nown = ro;
                if (int rc = ensure_own(nown)) return rc;
                // 2. ids to owners
                SNC(exchange(send_row, scnt, soff, B.recv_row.p, rcnt, roff, 4));
                SNC(exchange(send_ord, scnt, soff, B.recv_ord.p, rcnt, roff, 4));
                own_rows = static_cast<const uint32_t*>(B.recv_row.p); own_ords = static_cast<const uint32_t*>(B.recv_ord.p);
            } else {
                SCU(cudaMemcpyAsync(B.h_counts, counts, 4, cudaMemcpyDeviceToHost, st));
                SCU(cudaStreamSynchronize(st));
                nown = B.h_counts[0];
            }
            // 3. owners gather, rows travel back
            float* own_rows_buf = static_cast<float*>(world > 1 ? B.rows_own.p : B.rows_req.p);
            float* own_bias_buf = static_cast<float*>(world > 1 ? B.bias_own.p : B.bias_req.p);
            if (nown) { SYNC_DISPATCH_D(D, sync_gather_kernel<kD><<<148 * 8, 256, 0, st>>>(m, rank, own_rows, nown, own_rows_buf, own_bias_buf)); ++*launches; }
            if (world > 1) {
                SNC(exchange(own_rows_buf, rcnt, roff, B.rows_req.p, scnt, soff, (size_t)D * 4));
                SNC(exchange(own_bias_buf, rcnt, roff, B.bias_req.p, scnt, soff, 4));
            }
            rows_for_compute = static_cast<float*>(B.rows_req.p); bias_for_compute = static_cast<float*>(B.bias_req.p);
            // 4. fused forward/backward on the received rows
            SYNC_DISPATCH_D(D, sync_ewma_compute_kernel<kD><<<grid_p, 256, 0, st>>>(m, pl, it, rows_for_compute, bias_for_compute, pos_of_slot,
                                                                                     static_cast<float*>(B.grads_req.p), static_cast<float*>(B.bgrads_req.p),
                                                                                     static_cast<float*>(B.dalpha.p), world == 1 ? send_row : nullptr));
            ++*launches;
            // 5. gradients to owners, sparse visits on the owner's shard
            const float* g_own = static_cast<const float*>(B.grads_req.p); const float* bg_own = static_cast<const float*>(B.bgrads_req.p);
            if (world > 1) {
                SNC(exchange(B.grads_req.p, scnt, soff, B.grads_own.p, rcnt, roff, (size_t)D * 4));
                SNC(exchange(B.bgrads_req.p, scnt, soff, B.bgrads_own.p, rcnt, roff, 4));
                g_own = static_cast<const float*>(B.grads_own.p); bg_own = static_cast<const float*>(B.bgrads_own.p);
            }
            const uint64_t t_adam = num_updates + (rounds_done + 1) * (uint64_t)pl.P * world;
            if (o.adam) { o.c1 = 1.0f - powf(0.9f, (float)t_adam); o.c2 = 1.0f - powf(0.999f, (float)t_adam); }
            if (nown) {
                unsigned long long* k_in = static_cast<unsigned long long*>(B.keys_in.p); unsigned long long* k_out = static_cast<unsigned long long*>(B.keys_out.p);
                uint32_t* v_in = static_cast<uint32_t*>(B.vals_in.p); uint32_t* v_out = static_cast<uint32_t*>(B.vals_out.p);
                sync_keys_kernel<<<148 * 4, 256, 0, st>>>(own_rows, own_ords, nown, k_in, v_in);
                size_t tmp = B.cub_tmp.cap;
                SCU(cub::DeviceRadixSort::SortPairs(B.cub_tmp.p, tmp, k_in, k_out, v_in, v_out, (int)nown, 0, 64, st));
                SYNC_DISPATCH_D(D, sync_apply_kernel<kD><<<148 * 8, 256, 0, st>>>(m, rank, k_out, v_out, g_own, bg_own, nown, o));
                *launches += 3;
            }
            // 6. dense parameters: gradient summed over every partition of every rank, one step on each replica
            if (world > 1) SNC(NC->AllReduce(B.dalpha.p, B.dalpha.p, m.ndense, ncclFloat, ncclSum, comm, st));
            sync_dense_kernel<<<(unsigned)((m.ndense + 127) / 128), 128, 0, st>>>(m, static_cast<float*>(B.dalpha.p), o);
            ++*launches;
        }
    }
    SCU(cudaGetLastError());
    return 0;
}

}  // namespace sbr
