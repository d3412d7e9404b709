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
#include <cstdlib>
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
    // step_base = rounds of this run already scheduled: the step counters in HBM move only at the end of a run, so that the
    // requests of the next round can be built while the current one is still computing
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

// bucket the requests by (half-round group, owner GPU), warp-aggregated: pass 0 counts, pass 1 scatters.  Group = which half
// of the partitions the slot belongs to (the two halves of a round are exchanged and computed as two overlapping pipelines);
// a group's entries for owner g sit at  goff[grp] + sum of the group's counts below g.  pos_of_slot is group-local.
__global__ void sync_bucket_kernel(ModelDev m, const uint32_t* __restrict__ req_id, size_t nslots, size_t slots_grp0, size_t goff1, int ngrp, int pass,
                                   unsigned int* counts /*[2G]*/, unsigned int* cursor /*[2G]*/, uint2* __restrict__ send_pair,
                                   uint32_t* __restrict__ pos_of_slot, const uint32_t* __restrict__ req_ord) {
    const int G = (int)m.gmask + 1, lane = threadIdx.x & 31;
    __shared__ unsigned int off[16];
    if (pass == 1) {
        if (threadIdx.x < 16) {
            const int grp = threadIdx.x / 8, g = threadIdx.x % 8;
            unsigned int o = 0;
            for (int q = 0; q < g && q < G; ++q) o += counts[grp * G + q];
            off[threadIdx.x] = o;
        }
        __syncthreads();
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nround = (nslots + stride - 1) / stride * stride;  // keep warps converged for the ballots
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        const uint32_t id = i < nslots ? req_id[i] : kInvalid;
        const int grp = (ngrp > 1 && i >= slots_grp0) ? 1 : 0;
        const int bucket = id == kInvalid ? -1 : grp * G + (int)(id & m.gmask);
        for (int b = 0; b < ngrp * G; ++b) {
            const unsigned mask = __ballot_sync(kFull, bucket == b);
            if (!mask) continue;
            const int leader = __ffs(mask) - 1;
            unsigned int base = 0;
            if (lane == leader) base = atomicAdd(pass == 0 ? &counts[b] : &cursor[b], (unsigned int)__popc(mask));
            base = __shfl_sync(kFull, base, leader);
            if (pass == 1 && bucket == b) {
                const unsigned int pos = off[(b / G) * 8 + b % G] + base + (unsigned int)__popc(mask & ((1u << lane) - 1));
                send_pair[(b >= G ? goff1 : 0) + pos] = make_uint2(id >> m.gshift, req_ord[i]);
                pos_of_slot[i] = pos;
            }
        }
        if (pass == 1 && i < nslots && bucket < 0) pos_of_slot[i] = kInvalid;
    }
}

// owner side: copy the requested rows (weights only) and biases out of the local shard, warp per request
// Where entry j of a list that is segmented by peer GPU goes: segment g = [lo[g], lo[g + 1]) lands in rows[g] / bias[g] (peer
// g's buffer, mapped through CUDA IPC -- a plain store over NVLink -- or a local one) at element base[g] + (j - lo[g]).
struct Route {
    int G;
    size_t lo[9], base[8];
    float* rows[8]; float* bias[8];
    __device__ __forceinline__ void find(size_t j, int& g, size_t& idx) const {
        g = 0;
#pragma unroll
        for (int q = 1; q < 8; ++q) if (q < G && j >= lo[q]) g = q;
        idx = base[g] + (j - lo[g]);
    }
};
template <int D>
__global__ void __launch_bounds__(256) sync_gather_kernel(ModelDev m, int self, const uint2* __restrict__ pairs, size_t n, Route rt, size_t rot) {
    constexpr int V = VecOf<D>::V;
    const int lane = threadIdx.x & 31;
    for (size_t i = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += (size_t)gridDim.x * (blockDim.x >> 5)) {
        // the list is segmented by requester; every owner starts with the requester AFTER itself and wraps around, so that at
        // any moment each requester's NVLink ingress is fed by one owner (walking the list from 0, all owners store into
        // requester 0 first, then all into requester 1, ...: measured 3.7 ms instead of ~0.5 ms per half-round on 8 GPUs)
        size_t j = i + rot;
        if (j >= n) j -= n;
        const uint32_t r = pairs[j].x;
        float w[V];
        row_load_cg<D>(shard_item_rec(m, self, r), lane, w);
        int g; size_t idx;
        rt.find(j, g, idx);
        vec_store<D>(rt.rows[g] + idx * D, lane, w);
        if (lane == 0) rt.bias[g][idx] = __ldcg(reinterpret_cast<const float*>(shard_bias_rec(m, self, r)));
    }
}

// requester side: fused EWMA forward/backward of one sub-sequence per warp on the received rows (ewma.rs:266-352)
template <int D>
__global__ void __launch_bounds__(256) sync_ewma_compute_kernel(ModelDev m, PlanDev pl, uint32_t it, float* __restrict__ rows,
                                                                float* __restrict__ biases, const uint32_t* __restrict__ pos_of_slot,
                                                                Route gr, float* __restrict__ dalpha_sum, uint2* __restrict__ own_pairs,
                                                                uint32_t p_lo, uint32_t p_hi, uint64_t step_base) {
    constexpr int V = VecOf<D>::V;
    const int lane = threadIdx.x & 31;
    const uint32_t p = p_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= p_hi) return;
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
        if (m.loss == 2 && own_pairs) {
            // WARP (sequence_model.rs:47-68) on one GPU: the requested row is candidate 0; further candidates are read from the
            // table itself (nothing writes it before the round's apply stage) and the accepted one replaces the request
            const uint64_t key = pl.keys[p], step = step_base + pl.step_ctr[p];
            for (int j = 1; j < 5 && !(1.0f - posv + ngs > 0.0f); ++j) {
                const uint32_t cand = draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range);
                row_load_cg<D>(item_rec(m, cand), lane, qv);
                const float bq = __ldcg(reinterpret_cast<const float*>(bias_rec(m, cand)));
                ngs = warp_dot<D>(s, qv) + bq;
                vec_store<D>(rows + (size_t)pq * D, lane, qv);
                if (lane == 0) { biases[pq] = bq; own_pairs[pq].x = cand; }
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
        // the three entries go where their rows live: the owner's gradient buffer (peer memory over NVLink, or local)
        int gq, gp_, gx; size_t iq, ip, ix;
        gr.find(pq, gq, iq); gr.find(pp, gp_, ip); gr.find(px, gx, ix);
        vec_store<D>(gr.rows[gq] + iq * D, lane, gn);
        vec_store<D>(gr.rows[gp_] + ip * D, lane, gp);
        vec_store<D>(gr.rows[gx] + ix * D, lane, dx);
        if (lane == 0) { gr.bias[gq][iq] = g; gr.bias[gp_][ip] = -g; gr.bias[gx][ix] = __int_as_float(0x7fc00000); }  // NaN: no bias entry for inputs
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const float dal = da[v] * a[v] * (1.0f - a[v]);
        if (D >= 32 || lane < D) atomicAdd(dalpha_sum + (D < 32 ? lane : lane * V + v), dal);
    }
    if (lane == 0) { pl.loss_acc[p] += loss_seq; pl.examples[p] += (unsigned long long)Tn; }
}

__global__ void sync_advance_steps_kernel(PlanDev pl, uint64_t rounds) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < pl.P) pl.step_ctr[p] += rounds;
}

__global__ void sync_keys_kernel(const uint2* __restrict__ pairs, size_t n, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        keys[j] = ((unsigned long long)pairs[j].x << 32) | pairs[j].y;
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
                else if (lane < rec_lines + grad_lines)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(grads + (size_t)vals[j + stride] * D) + (lane - rec_lines) * 128));
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
        const float rc1 = 1.0f / o.c1, rc2 = 1.0f / o.c2;
        if (lane == 0) bq = __ldcg(shard_bias_rec(m, self, r));
        for (size_t e = j; e < n && (uint32_t)(keys[e] >> 32) == r; ++e) {
            const size_t src = vals[e];
            float g[V];
            vec_load<D>(grads + src * D, lane, g);
            if (!o.adam) {
#pragma unroll
                for (int v = 0; v < V; ++v) adagrad_elem(w[v], s1[v], g[v], o.lr, o.l2);
            } else {
                // adam_elem with the two bias corrections as reciprocals (one division per entry instead of three per element;
                // profiles/r2_c3_sync_apply_ncu_full.txt: this kernel was issue-bound on IEEE divisions, not memory-bound)
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float gg = g[v] + w[v] * o.l2;
                    s1[v] = 0.9f * s1[v] + (1.0f - 0.9f) * gg;
                    s2[v] = 0.999f * s2[v] + (1.0f - 0.999f) * gg * gg;
                    w[v] -= __fdividef(o.lr * (s1[v] * rc1), sqrtf(s2[v] * rc2) + 1e-8f);
                }
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

// request side of a round (double-buffered: the next round's requests are built and counted while this one computes)
struct ReqSet {
    Buf req_id, req_ord, send_pair, pos_of_slot, counts, allcounts;
    unsigned int* h_counts = nullptr;   // pinned [G][16]: every rank's (group, owner) counts
    cudaEvent_t ready = nullptr;
    ~ReqSet() { if (h_counts) cudaFreeHost(h_counts); if (ready) cudaEventDestroy(ready); }
};
// data side of one half-round group
struct GrpBufs {
    Buf recv_pair, rows_req, bias_req, rows_own, bias_own, grads_req, bgrads_req, grads_own, bgrads_own, keys_in, keys_out, vals_in, vals_out, cub_tmp;
    size_t cap_own = 0;
    cudaEvent_t ev_gather = nullptr, ev_compute = nullptr, ev_apply = nullptr;
    ~GrpBufs() { for (cudaEvent_t e : {ev_gather, ev_compute, ev_apply}) if (e) cudaEventDestroy(e); }
};
struct SyncBuffers {
    ReqSet rs[2];
    GrpBufs grp[2];
    Buf dalpha, scal;
    unsigned int* h_scal = nullptr;     // pinned scalar for the round-count agreement
    cudaStream_t s_b = nullptr, s_req = nullptr;
    cudaEvent_t ev_round = nullptr, ev_consumed = nullptr;
    // peer mappings (CUDA IPC) of every rank's receive buffers: [group][0 rows_req, 1 bias_req, 2 grads_own, 3 bgrads_own][rank]
    bool p2p = false, p2p_tried = false;
    bool ce = false;                    // p2p transport: copy engines (cudaMemcpyAsync from a local staging buffer, one stream per peer) instead of SM stores
    cudaStream_t s_copy[8] = {};
    cudaEvent_t ev_src[2] = {}, ev_copy[2][8] = {};
    float* peer[2][4][8] = {};
    std::vector<void*> ipc_opened;
    Buf ipc_dev;
    ~SyncBuffers() {
        for (void* q : ipc_opened) cudaIpcCloseMemHandle(q);
        for (cudaStream_t c : s_copy) if (c) cudaStreamDestroy(c);
        for (cudaEvent_t e : ev_src) if (e) cudaEventDestroy(e);
        for (auto& row : ev_copy) for (cudaEvent_t e : row) if (e) cudaEventDestroy(e);
        if (h_scal) cudaFreeHost(h_scal);
        if (s_b) cudaStreamDestroy(s_b);
        if (s_req) cudaStreamDestroy(s_req);
        for (cudaEvent_t e : {ev_round, ev_consumed}) if (e) cudaEventDestroy(e);
    }
};

SyncBuffers* sync_buffers_new() { return new SyncBuffers(); }
bool sync_buffers_p2p(const SyncBuffers* b) { return b && b->p2p; }
bool sync_buffers_copy_engine(const SyncBuffers* b) { return b && b->ce; }
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

// returns 0 ok, 1 cuda error, 2 nccl error.  `comm` may be null when world == 1.
//
// One round on G GPUs (G == 1: the exchanges are the identity and there is one group):
//   request stream : ids + uniform negatives of the NEXT round, bucketed by (group, owner); counts all-gathered and copied to
//                    the host while the current round runs -- the host never waits on the critical path
//   group A / B    : the two halves of the partitions, each a pipeline on its own stream:
//                      (row, order) pairs --all-to-all--> owners gather rows + biases from their shard
//                      rows + biases      --all-to-all--> fused EWMA forward/backward on the received rows
//                      gradient rows      --all-to-all--> owners sort by (row, reference order) and apply, A before B
//                    NCCL serialises the transfers of one communicator, so A's transfers overlap B's kernels and vice versa.
//   dense          : all-reduce of the round-summed alpha gradient, one step on every replica.
int run_sync_ewma(const ModelDev& m, PlanDev& pl, SyncBuffers& B, void* comm_v, int rank, int world, uint64_t num_updates,
                  cudaStream_t st, int* launches, uint64_t* rounds_out, std::string* err) {
    ncclComm_t comm = static_cast<ncclComm_t>(comm_v);
    const NcclApi* NC = nullptr;
    if (world > 1) { NC = nccl_api(err); if (!NC) return 2; }
    const int G = world, D = m.D, Tm1 = m.T - 1;
    const int NGRP = (world > 1 && pl.P >= 2) ? 2 : 1;
    const uint32_t p_split = NGRP == 2 ? pl.P / 2 : pl.P;                  // group 0 = partitions [0, p_split)
    const size_t nslots = (size_t)pl.P * Tm1 * 3;
    const size_t slots_grp[2] = {(size_t)p_split * Tm1 * 3, nslots - (size_t)p_split * Tm1 * 3};
    if (!B.s_b) { SCU(cudaStreamCreateWithFlags(&B.s_b, cudaStreamNonBlocking)); SCU(cudaStreamCreateWithFlags(&B.s_req, cudaStreamNonBlocking)); }
    if (!B.ev_round) { SCU(cudaEventCreateWithFlags(&B.ev_round, cudaEventDisableTiming)); SCU(cudaEventCreateWithFlags(&B.ev_consumed, cudaEventDisableTiming)); }
    if (!B.h_scal) SCU(cudaHostAlloc(&B.h_scal, 64, cudaHostAllocDefault));
    SCU(B.scal.ensure(64)); SCU(B.dalpha.ensure(m.ndense * 4));
    for (ReqSet& r : B.rs) {
        SCU(r.req_id.ensure(nslots * 4)); SCU(r.req_ord.ensure(nslots * 4)); SCU(r.send_pair.ensure(nslots * 8)); SCU(r.pos_of_slot.ensure(nslots * 4));
        SCU(r.counts.ensure(128)); SCU(r.allcounts.ensure(8 * 16 * 4));
        if (!r.h_counts) SCU(cudaHostAlloc(&r.h_counts, 8 * 16 * sizeof(unsigned int), cudaHostAllocDefault));
        if (!r.ready) SCU(cudaEventCreateWithFlags(&r.ready, cudaEventDisableTiming));
    }
    // Rows other ranks may request from this shard per round and group.  The usual load is ~ the group's own slot count; a
    // skewed id distribution (popular ids in one residue class mod world) can send up to world x that to one owner, so the
    // owner-side buffers GROW when a round needs more (both pipelines are drained first) -- except when peers hold mappings
    // of them (p2p mode), where they are sized at twice the expected load once and a round that needs more is an error.
    auto ensure_own = [&](GrpBufs& g, size_t need) -> int {
        if (need <= g.cap_own) return 0;
        if (B.p2p) { *err = "synchronous exchange: one owner was asked for more rows than its peer-mapped buffers hold (id distribution too skewed for id % world sharding)"; return 3; }
        SCU(cudaDeviceSynchronize());
        g.cap_own = need;
        SCU(g.keys_in.ensure(need * 8)); SCU(g.keys_out.ensure(need * 8)); SCU(g.vals_in.ensure(need * 4)); SCU(g.vals_out.ensure(need * 4));
        size_t cub_bytes = 0;
        SCU(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<unsigned long long*>(nullptr), static_cast<unsigned long long*>(nullptr),
                                            static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), (int)need, 0, 64, st));
        SCU(g.cub_tmp.ensure(cub_bytes));
        if (world > 1) { SCU(g.recv_pair.ensure(need * 8)); SCU(g.grads_own.ensure(need * D * 4)); SCU(g.bgrads_own.ensure(need * 4)); }
        return 0;
    };
    for (int q = 0; q < NGRP; ++q) {
        GrpBufs& g = B.grp[q];
        SCU(g.rows_req.ensure(slots_grp[q] * D * 4)); SCU(g.bias_req.ensure(slots_grp[q] * 4));
        if (world == 1) { SCU(g.grads_req.ensure(slots_grp[q] * D * 4)); SCU(g.bgrads_req.ensure(slots_grp[q] * 4)); }
        if (int rc = ensure_own(g, world == 1 ? slots_grp[q] : slots_grp[q] * 2 + 4096)) return rc;
        for (cudaEvent_t* e : {&g.ev_gather, &g.ev_compute, &g.ev_apply}) if (!*e) SCU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    // ---- peer mappings of the receive buffers (once per plan): rows and gradient rows then cross NVLink as plain stores of the
    // gather / compute kernels, no staging copy and no collective on the data path.  Falls back to NCCL send/recv when any
    // rank cannot map a peer (SBR_SYNC_NO_P2P=1 forces the fallback).
    if (world > 1 && !B.p2p_tried) {
        B.p2p_tried = true;
        struct Pack { cudaIpcMemHandle_t h[2][4]; };
        static_assert(sizeof(Pack) == 512, "8 IPC handles");
        Pack mine{};
        int ok = getenv("SBR_SYNC_NO_P2P") ? 0 : 1;
        for (int q = 0; q < NGRP && ok; ++q) {
            void* bufs[4] = {B.grp[q].rows_req.p, B.grp[q].bias_req.p, B.grp[q].grads_own.p, B.grp[q].bgrads_own.p};
            for (int k = 0; k < 4 && ok; ++k) if (cudaIpcGetMemHandle(&mine.h[q][k], bufs[k]) != cudaSuccess) { ok = 0; cudaGetLastError(); }
        }
        SCU(B.ipc_dev.ensure((size_t)world * sizeof(Pack) + 64));
        std::vector<Pack> all(world);
        SCU(cudaMemcpyAsync(static_cast<char*>(B.ipc_dev.p) + (size_t)rank * sizeof(Pack), &mine, sizeof(Pack), cudaMemcpyHostToDevice, st));
        SNC(NC->AllGather(static_cast<char*>(B.ipc_dev.p) + (size_t)rank * sizeof(Pack), B.ipc_dev.p, sizeof(Pack), ncclChar, comm, st));
        SCU(cudaMemcpyAsync(all.data(), B.ipc_dev.p, (size_t)world * sizeof(Pack), cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        for (int q = 0; q < NGRP && ok; ++q) {
            void* bufs[4] = {B.grp[q].rows_req.p, B.grp[q].bias_req.p, B.grp[q].grads_own.p, B.grp[q].bgrads_own.p};
            for (int k = 0; k < 4 && ok; ++k)
                for (int g = 0; g < world && ok; ++g) {
                    if (g == rank) { B.peer[q][k][g] = static_cast<float*>(bufs[k]); continue; }
                    void* ptr = nullptr;
                    if (cudaIpcOpenMemHandle(&ptr, all[g].h[q][k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
                    B.ipc_opened.push_back(ptr);
                    B.peer[q][k][g] = static_cast<float*>(ptr);
                }
        }
        // everybody or nobody
        int* d_ok = reinterpret_cast<int*>(static_cast<char*>(B.ipc_dev.p) + (size_t)world * sizeof(Pack));
        SCU(cudaMemcpyAsync(d_ok, &ok, 4, cudaMemcpyHostToDevice, st));
        SNC(NC->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, st));
        SCU(cudaMemcpyAsync(&ok, d_ok, 4, cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        B.p2p = ok != 0;
        // Transport of the peer-mapped mode.  Default: the gather / compute kernels store straight into the peer buffers (6.5 M
        // steps/s on 8 GPUs, 3.05 M on 2).  SBR_SYNC_P2P=copy: kernels write a local staging buffer and one cudaMemcpyAsync per
        // peer (own stream each) pushes it over NVLink with the copy engines -- the SMs stay free for the other half-round's
        // kernels (3.18 M on 2 GPUs), but seven concurrent peer copies per GPU run at ~230 GB/s on 8 GPUs (5.5 M).
        const char* tr_env = getenv("SBR_SYNC_P2P");
        B.ce = B.p2p && tr_env && std::string(tr_env) == "copy";
        if (B.ce) {
            for (int g = 0; g < world; ++g) if (g != rank && !B.s_copy[g]) SCU(cudaStreamCreateWithFlags(&B.s_copy[g], cudaStreamNonBlocking));
            for (int q = 0; q < 2; ++q) {
                if (!B.ev_src[q]) SCU(cudaEventCreateWithFlags(&B.ev_src[q], cudaEventDisableTiming));
                for (int g = 0; g < world; ++g) if (!B.ev_copy[q][g]) SCU(cudaEventCreateWithFlags(&B.ev_copy[q][g], cudaEventDisableTiming));
            }
        }
    }
    if (world > 1 && (!B.p2p || B.ce))   // NCCL fallback / copy-engine transport: staging buffers on both sides
        for (int q = 0; q < NGRP; ++q) {
            GrpBufs& g = B.grp[q];
            SCU(g.rows_own.ensure(g.cap_own * D * 4)); SCU(g.bias_own.ensure(g.cap_own * 4));
            SCU(g.grads_req.ensure(slots_grp[q] * D * 4)); SCU(g.bgrads_req.ensure(slots_grp[q] * 4));
        }
    SCU(cudaMemsetAsync(B.dalpha.p, 0, m.ndense * 4, st));
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;
    cudaStream_t gst[2] = {st, B.s_b};
    uint64_t rounds_done = 0;

    // every rank must run the same number of rounds (collectives inside): the ranks agree on the smallest partition length
    uint32_t n_rounds = pl.n;
    if (world > 1) {
        unsigned int* d_scal = static_cast<unsigned int*>(B.scal.p);
        B.h_scal[0] = pl.n;
        SCU(cudaMemcpyAsync(d_scal, B.h_scal, 4, cudaMemcpyHostToDevice, st));
        SNC(NC->AllReduce(d_scal, d_scal, 1, ncclUint32, ncclMin, comm, st));
        SCU(cudaMemcpyAsync(B.h_scal, d_scal, 4, cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        n_rounds = B.h_scal[0];
    }
    *rounds_out = (uint64_t)n_rounds * (uint64_t)pl.epochs;

    // requests of round `it` into set `r` on the request stream (after `after` has happened: the epoch shuffle / the host
    // having consumed the set's previous counts is implied by program order)
    auto stage_requests = [&](ReqSet& r, uint32_t it, uint64_t step_base, cudaEvent_t after) -> int {
        cudaStream_t sr = B.s_req;
        if (after) SCU(cudaStreamWaitEvent(sr, after, 0));
        unsigned int* counts = static_cast<unsigned int*>(r.counts.p);       // [0..16) counts, [16..32) cursors
        SCU(cudaMemsetAsync(counts, 0, 128, sr));
        uint32_t* req_id = static_cast<uint32_t*>(r.req_id.p); uint32_t* req_ord = static_cast<uint32_t*>(r.req_ord.p);
        sync_request_kernel<<<148 * 8, 256, 0, sr>>>(m, pl, it, step_base, (uint32_t)rank * pl.P, req_id, req_ord);
        for (int pass = 0; pass < 2; ++pass)
            sync_bucket_kernel<<<148 * 4, 256, 0, sr>>>(m, req_id, nslots, slots_grp[0], slots_grp[0], NGRP, pass, counts, counts + 16,
                                                        static_cast<uint2*>(r.send_pair.p), static_cast<uint32_t*>(r.pos_of_slot.p), req_ord);
        *launches += 3;
        if (world > 1) {
            SNC(NC->AllGather(counts, r.allcounts.p, 16, ncclUint32, comm, sr));
            SCU(cudaMemcpyAsync(r.h_counts, r.allcounts.p, (size_t)G * 16 * 4, cudaMemcpyDeviceToHost, sr));
        } else SCU(cudaMemcpyAsync(r.h_counts, counts, 16 * 4, cudaMemcpyDeviceToHost, sr));
        SCU(cudaEventRecord(r.ready, sr));
        return 0;
    };
    auto exchange = [&](cudaStream_t s, const void* sendbuf, const size_t* scnt, const size_t* soff, void* recvbuf, const size_t* rcnt, const size_t* roff, size_t elem,
                        const void* sendbuf2, void* recvbuf2, size_t elem2, bool skip_self) -> ncclResult_t {
        ncclResult_t r = NC->GroupStart();
        if (r != ncclSuccess) return r;
        for (int g = 0; g < G && r == ncclSuccess; ++g) {
            if (g == rank && skip_self) continue;
            if (scnt[g]) r = NC->Send(static_cast<const char*>(sendbuf) + soff[g] * elem, scnt[g] * elem, ncclChar, g, comm, s);
            if (r == ncclSuccess && rcnt[g]) r = NC->Recv(static_cast<char*>(recvbuf) + roff[g] * elem, rcnt[g] * elem, ncclChar, g, comm, s);
            if (sendbuf2 && r == ncclSuccess && scnt[g]) r = NC->Send(static_cast<const char*>(sendbuf2) + soff[g] * elem2, scnt[g] * elem2, ncclChar, g, comm, s);
            if (sendbuf2 && r == ncclSuccess && rcnt[g]) r = NC->Recv(static_cast<char*>(recvbuf2) + roff[g] * elem2, rcnt[g] * elem2, ncclChar, g, comm, s);
        }
        const ncclResult_t re = NC->GroupEnd();   // the group is always closed, also after a failed Send / Recv
        return r != ncclSuccess ? r : re;
    };

    // SBR_SYNC_TRACE=1: CUDA-event timeline of round 2 on stderr (which stage of which pipeline ends when)
    const bool trace = getenv("SBR_SYNC_TRACE") != nullptr;
    std::vector<std::pair<std::string, cudaEvent_t>> tr;
    auto mark = [&](const char* name, int q, cudaStream_t s, bool on) {
        if (!on) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s);
        tr.emplace_back(std::string(name) + (q == 0 ? " A" : q == 1 ? " B" : ""), e);
    };
    int cur = 0;
    for (int ep = 0; ep < pl.epochs; ++ep) {
        sync_shuffle_kernel<<<(pl.P + 127) / 128, 128, 0, st>>>(pl);
        ++*launches;
        SCU(cudaEventRecord(B.ev_consumed, st));
        if (n_rounds) if (int rc = stage_requests(B.rs[cur], 0, rounds_done, B.ev_consumed)) return rc;
        for (uint32_t it = 0; it < n_rounds; ++it, ++rounds_done) {
            ReqSet& R = B.rs[cur];
            SCU(cudaEventSynchronize(R.ready));     // counts of this round on the host (enqueued one round ago)
            size_t scnt[2][8] = {}, soff[2][8] = {}, rcnt[2][8] = {}, roff[2][8] = {}, nown[2] = {0, 0};
            for (int q = 0; q < NGRP; ++q) {
                size_t so = 0, ro = 0;
                for (int g = 0; g < G; ++g) {
                    scnt[q][g] = R.h_counts[(world > 1 ? rank * 16 : 0) + q * G + g]; soff[q][g] = so; so += scnt[q][g];
                    rcnt[q][g] = world > 1 ? R.h_counts[g * 16 + q * G + rank] : scnt[q][g]; roff[q][g] = ro; ro += rcnt[q][g];
                }
                nown[q] = ro;
                if (B.p2p)   // peer-mapped buffers cannot grow: every rank sees every rank's load, so all ranks give up together
                    for (int g = 0; g < G; ++g) {
                        size_t load = 0;
                        for (int src = 0; src < G; ++src) load += R.h_counts[src * 16 + q * G + g];
                        if (load > B.grp[q].cap_own) { *err = "synchronous exchange: one owner was asked for more rows than its peer-mapped buffers hold (id distribution too skewed for id % world sharding)"; return 3; }
                    }
                if (int rc = ensure_own(B.grp[q], nown[q])) return rc;
            }
            // both pipelines start once the previous round's table is final and the epoch shuffle (stream `st`) has run
            if (NGRP == 2) { SCU(cudaEventRecord(B.ev_consumed, st)); SCU(cudaStreamWaitEvent(gst[1], B.ev_consumed, 0)); }
            if (rounds_done > 0 && NGRP == 2) SCU(cudaStreamWaitEvent(gst[0], B.ev_round, 0));
            const bool tron = trace && rounds_done == 2;
            mark("start", 0, gst[0], tron); mark("start", 1, gst[NGRP - 1], tron && NGRP == 2);
            const uint2* send_pair = static_cast<const uint2*>(R.send_pair.p);
            const uint32_t* pos_of_slot = static_cast<const uint32_t*>(R.pos_of_slot.p);
            const uint2* own_pairs[2];
            // 1. (row, order) pairs to the owners; owners gather
            for (int q = 0; q < NGRP; ++q) {
                GrpBufs& g = B.grp[q];
                const uint2* sp = send_pair + (q ? slots_grp[0] : 0);
                own_pairs[q] = sp;
                if (world > 1) {
                    SNC(exchange(gst[q], sp, scnt[q], soff[q], g.recv_pair.p, rcnt[q], roff[q], 8, nullptr, nullptr, 0, false));
                    own_pairs[q] = static_cast<const uint2*>(g.recv_pair.p);
                    mark("pairs exchanged", q, gst[q], tron);
                }
                // where the gathered rows go: segment g of the owner-side list came from rank g and belongs at that rank's
                // requester-side position soff_g[me] + i -- in its rows_req buffer (p2p: a store over NVLink; own segment: local;
                // NCCL fallback: the local staging buffer, identity position)
                Route rt; rt.G = G;
                for (int g2 = 0; g2 < G; ++g2) {
                    rt.lo[g2] = roff[q][g2];
                    size_t soff_at = 0;   // position of my segment in rank g2's send list of this group
                    for (int o2 = 0; o2 < rank; ++o2) soff_at += world > 1 ? R.h_counts[g2 * 16 + q * G + o2] : 0;
                    const bool direct = (B.p2p && !B.ce) || g2 == rank || world == 1;
                    rt.base[g2] = direct ? soff_at : roff[q][g2];
                    rt.rows[g2] = direct ? (world > 1 && B.p2p ? B.peer[q][0][g2] : static_cast<float*>(g.rows_req.p)) : static_cast<float*>(g.rows_own.p);
                    rt.bias[g2] = direct ? (world > 1 && B.p2p ? B.peer[q][1][g2] : static_cast<float*>(g.bias_req.p)) : static_cast<float*>(g.bias_own.p);
                }
                for (int g2 = G; g2 <= 8; ++g2) rt.lo[g2] = nown[q];
                const size_t rot = (world > 1 && rank + 1 < G) ? roff[q][rank + 1] : 0;   // first entry of requester (rank + 1) % G
                if (nown[q]) { SYNC_DISPATCH_D(D, sync_gather_kernel<kD><<<148 * 8, 256, 0, gst[q]>>>(m, rank, own_pairs[q], nown[q], rt, rot)); ++*launches; }
                if (B.ce) {   // the staged rows of every other rank go out through the copy engines, one stream per peer
                    SCU(cudaEventRecord(B.ev_src[q], gst[q]));
                    for (int g2 = 0; g2 < G; ++g2) {
                        if (g2 == rank || !rcnt[q][g2]) continue;
                        size_t soff_at = 0;
                        for (int o2 = 0; o2 < rank; ++o2) soff_at += R.h_counts[g2 * 16 + q * G + o2];
                        SCU(cudaStreamWaitEvent(B.s_copy[g2], B.ev_src[q], 0));
                        SCU(cudaMemcpyAsync(B.peer[q][0][g2] + soff_at * D, static_cast<float*>(g.rows_own.p) + roff[q][g2] * D, rcnt[q][g2] * (size_t)D * 4, cudaMemcpyDeviceToDevice, B.s_copy[g2]));
                        SCU(cudaMemcpyAsync(B.peer[q][1][g2] + soff_at, static_cast<float*>(g.bias_own.p) + roff[q][g2], rcnt[q][g2] * 4, cudaMemcpyDeviceToDevice, B.s_copy[g2]));
                        SCU(cudaEventRecord(B.ev_copy[q][g2], B.s_copy[g2]));
                    }
                }
                SCU(cudaEventRecord(g.ev_gather, gst[q]));
                mark("gathered", q, gst[q], tron);
            }
            // the NEXT round's requests go to the other set, whose last readers were the kernels of the previous round
            if (it + 1 < n_rounds) if (int rc = stage_requests(B.rs[cur ^ 1], it + 1, rounds_done + 1, it > 0 ? B.ev_round : nullptr)) return rc;
            // 2. rows + biases back to the requesters; fused forward / backward
            for (int q = 0; q < NGRP; ++q) {
                GrpBufs& g = B.grp[q];
                if (world > 1 && !B.p2p) SNC(exchange(gst[q], g.rows_own.p, rcnt[q], roff[q], g.rows_req.p, scnt[q], soff[q], (size_t)D * 4, g.bias_own.p, g.bias_req.p, 4, true));
                if (B.ce) for (int g2 = 0; g2 < G; ++g2) if (g2 != rank && rcnt[q][g2]) SCU(cudaStreamWaitEvent(gst[q], B.ev_copy[q][g2], 0));
                if (B.p2p) SNC(NC->AllReduce(static_cast<int*>(B.scal.p) + 4 + q, static_cast<int*>(B.scal.p) + 4 + q, 1, ncclInt, ncclSum, comm, gst[q]));   // every rank's rows have landed
                mark("rows exchanged", q, gst[q], tron && world > 1);
                const uint32_t p_lo = q ? p_split : 0, p_hi = (q || NGRP == 1) ? pl.P : p_split;
                const int grid_p = (int)((p_hi - p_lo + 7) / 8);
                uint2* warp_pairs = (world == 1 && m.loss == 2) ? static_cast<uint2*>(R.send_pair.p) : nullptr;
                // where the gradient entries go: requester-side position pos in segment g belongs at owner g's list position
                // roff_g[me] + i -- in its grads_own buffer (p2p / own segment) or in the local staging buffer (NCCL fallback)
                Route gr; gr.G = G;
                size_t sent = 0;
                for (int g2 = 0; g2 < G; ++g2) {
                    gr.lo[g2] = soff[q][g2]; sent = soff[q][g2] + scnt[q][g2];
                    size_t roff_at = 0;   // where my entries start in rank g2's owner-side list of this group
                    for (int s2 = 0; s2 < rank; ++s2) roff_at += world > 1 ? R.h_counts[s2 * 16 + q * G + g2] : 0;
                    const bool direct = (B.p2p && !B.ce) || (g2 == rank && world > 1);
                    gr.base[g2] = direct ? roff_at : soff[q][g2];
                    gr.rows[g2] = direct ? (B.p2p ? B.peer[q][2][g2] : static_cast<float*>(g.grads_own.p)) : static_cast<float*>(g.grads_req.p);
                    gr.bias[g2] = direct ? (B.p2p ? B.peer[q][3][g2] : static_cast<float*>(g.bgrads_own.p)) : static_cast<float*>(g.bgrads_req.p);
                }
                for (int g2 = G; g2 <= 8; ++g2) gr.lo[g2] = sent;
                if (grid_p) {
                    SYNC_DISPATCH_D(D, sync_ewma_compute_kernel<kD><<<grid_p, 256, 0, gst[q]>>>(m, pl, it, static_cast<float*>(g.rows_req.p), static_cast<float*>(g.bias_req.p),
                                                                                             pos_of_slot, gr, static_cast<float*>(B.dalpha.p), warp_pairs, p_lo, p_hi, rounds_done));
                    ++*launches;
                }
                SCU(cudaEventRecord(g.ev_compute, gst[q]));
                mark("computed", q, gst[q], tron);
                if (B.ce) {   // the staged entries go to their owners through the copy engines
                    for (int g2 = 0; g2 < G; ++g2) {
                        if (g2 == rank || !scnt[q][g2]) continue;
                        size_t roff_at = 0;
                        for (int s2 = 0; s2 < rank; ++s2) roff_at += R.h_counts[s2 * 16 + q * G + g2];
                        SCU(cudaStreamWaitEvent(B.s_copy[g2], g.ev_compute, 0));
                        SCU(cudaMemcpyAsync(B.peer[q][2][g2] + roff_at * D, static_cast<float*>(g.grads_req.p) + soff[q][g2] * D, scnt[q][g2] * (size_t)D * 4, cudaMemcpyDeviceToDevice, B.s_copy[g2]));
                        SCU(cudaMemcpyAsync(B.peer[q][3][g2] + roff_at, static_cast<float*>(g.bgrads_req.p) + soff[q][g2], scnt[q][g2] * 4, cudaMemcpyDeviceToDevice, B.s_copy[g2]));
                        SCU(cudaEventRecord(B.ev_copy[q][g2], B.s_copy[g2]));
                    }
                }
            }
            // 3. gradient rows to the owners; sparse visits on the owner's shard, group A's entries before group B's
            const uint64_t t_adam = num_updates + (rounds_done + 1) * (uint64_t)pl.P * world;
            if (o.adam) { o.c1 = 1.0f - powf(0.9f, (float)t_adam); o.c2 = 1.0f - powf(0.999f, (float)t_adam); }
            for (int q = 0; q < NGRP; ++q) {
                GrpBufs& g = B.grp[q];
                const float* g_own = static_cast<const float*>(world > 1 ? g.grads_own.p : g.grads_req.p);
                const float* bg_own = static_cast<const float*>(world > 1 ? g.bgrads_own.p : g.bgrads_req.p);
                if (world > 1) {
                    if (!B.p2p) SNC(exchange(gst[q], g.grads_req.p, scnt[q], soff[q], g.grads_own.p, rcnt[q], roff[q], (size_t)D * 4, g.bgrads_req.p, g.bgrads_own.p, 4, true));
                    else {
                        if (B.ce) for (int g2 = 0; g2 < G; ++g2) if (g2 != rank && scnt[q][g2]) SCU(cudaStreamWaitEvent(gst[q], B.ev_copy[q][g2], 0));
                        SNC(NC->AllReduce(static_cast<int*>(B.scal.p) + 6 + q, static_cast<int*>(B.scal.p) + 6 + q, 1, ncclInt, ncclSum, comm, gst[q]));   // every rank's entries have landed
                    }
                    mark("grads exchanged", q, gst[q], tron);
                }
                if (q == 0 && NGRP == 2) SCU(cudaStreamWaitEvent(gst[0], B.grp[1].ev_gather, 0));   // nobody still reads the table
                if (q == 1) SCU(cudaStreamWaitEvent(gst[1], B.grp[0].ev_apply, 0));
                if (nown[q]) {
                    unsigned long long* k_in = static_cast<unsigned long long*>(g.keys_in.p); unsigned long long* k_out = static_cast<unsigned long long*>(g.keys_out.p);
                    uint32_t* v_in = static_cast<uint32_t*>(g.vals_in.p); uint32_t* v_out = static_cast<uint32_t*>(g.vals_out.p);
                    sync_keys_kernel<<<148 * 4, 256, 0, gst[q]>>>(own_pairs[q], nown[q], k_in, v_in);
                    size_t tmp = g.cub_tmp.cap;
                    SCU(cub::DeviceRadixSort::SortPairs(g.cub_tmp.p, tmp, k_in, k_out, v_in, v_out, (int)nown[q], 0, 64, gst[q]));
                    SYNC_DISPATCH_D(D, sync_apply_kernel<kD><<<148 * 8, 256, 0, gst[q]>>>(m, rank, k_out, v_out, g_own, bg_own, nown[q], o));
                    *launches += 3;
                }
                SCU(cudaEventRecord(g.ev_apply, gst[q]));
                mark("applied", q, gst[q], tron);
            }
            // 4. dense parameters: gradient summed over every partition of every rank, one step on each replica
            cudaStream_t sl = gst[NGRP - 1];
            if (NGRP == 2) SCU(cudaStreamWaitEvent(sl, B.grp[0].ev_compute, 0));
            if (world > 1) SNC(NC->AllReduce(B.dalpha.p, B.dalpha.p, m.ndense, ncclFloat, ncclSum, comm, sl));
            sync_dense_kernel<<<(unsigned)((m.ndense + 127) / 128), 128, 0, sl>>>(m, static_cast<float*>(B.dalpha.p), o);
            ++*launches;
            SCU(cudaEventRecord(B.ev_round, sl));
            mark("round done", 2, sl, tron);
            cur ^= 1;
        }
        if (NGRP == 2) SCU(cudaStreamWaitEvent(st, B.ev_round, 0));   // the next epoch's shuffle / the caller see a finished epoch
    }
    sync_advance_steps_kernel<<<(pl.P + 127) / 128, 128, 0, st>>>(pl, rounds_done);
    ++*launches;
    if (!tr.empty()) {
        cudaDeviceSynchronize();
        for (auto& pr : tr) { float ms = 0.f; cudaEventElapsedTime(&ms, tr[0].second, pr.second); fprintf(stderr, "[sync trace rank %d] %-18s %8.3f ms\n", rank, pr.first.c_str(), ms); }
        for (auto& pr : tr) cudaEventDestroy(pr.second);
    }
    SCU(cudaGetLastError());
    return 0;
}

}  // namespace sbr
