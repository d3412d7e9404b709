// kernels_train.cu -- Hogwild training kernels (one warp == one reference "thread" / partition).
//
// Replaces the rayon body of fit_sequence_model (sequence_model.rs:100-175) and the wyrm graph it drives
// (lstm.rs:258-337, ewma.rs:266-352): for every sub-sequence of its partition a warp does gather -> recurrent
// forward -> negative sampling / scoring -> backward -> sparse optimizer visits, with the parameters shared
// lock-free in HBM/L2 exactly like Arc<HogwildParameter> (lstm.rs:175-181,259-260).  One launch runs all epochs.
#include <cuda_runtime.h>


#include "engine.h"

namespace sbr {

namespace {

// In-place Fisher-Yates of one partition by lane 0 (thread_rng.shuffle(partition), sequence_model.rs:109)
__device__ __forceinline__ void shuffle_partition(uint32_t* ord, uint32_t n, XorShift& rng) {
    uint32_t i = n;
    while (i >= 2) {
        i -= 1;
        uint32_t j = (uint32_t)xs_gen_below(rng, (uint64_t)i + 1);
        uint32_t a = ord[i], b = ord[j];
        ord[i] = b; ord[j] = a;
    }
}

struct StepOut { float loss; float g; };

__device__ __forceinline__ StepOut pair_loss(int loss_kind, float pos, float neg) {
    StepOut r;
    if (loss_kind == 0) {  // BPR: sigmoid(neg - pos)   lstm.rs:317
        float s = sigmoidf_(neg - pos);
        r.loss = s; r.g = s * (1.0f - s);
    } else {               // Hinge / WARP: relu(1 + neg - pos)   lstm.rs:318
        float v = 1.0f + neg - pos;
        r.loss = v > 0.0f ? v : 0.0f; r.g = v > 0.0f ? 1.0f : 0.0f;
    }
    return r;
}

// Scores h against the target and draws the negative (uniform, or WARP rejection sampling,
// sequence_model.rs:47-68): returns pos / neg scores, leaves the rows in p / q.
template <int D>
__device__ __forceinline__ void score_and_sample(const ModelDev& m, int lane, const float (&h)[VecOf<D>::V], uint32_t out,
                                                 uint64_t key, uint64_t step, uint32_t t, uint32_t range, float (&p)[VecOf<D>::V],
                                                 float (&q)[VecOf<D>::V], uint32_t& neg, float& pos, float& ngs) {
    row_load_cg<D>(item_rec(m, out), lane, p);
    pos = warp_dot<D>(h, p) + __ldcg(reinterpret_cast<const float*>(bias_rec(m, out)));
    const int tries = m.loss == 2 ? 5 : 1;
    for (int j = 0; j < tries; ++j) {
        neg = draw_item(key, step, t, (uint32_t)j, range);
        row_load_cg<D>(item_rec(m, neg), lane, q);
        ngs = warp_dot<D>(h, q) + __ldcg(reinterpret_cast<const float*>(bias_rec(m, neg)));
        if (1.0f - pos + ngs > 0.0f) break;  // warp-uniform
    }
}

// Pull an item's row record (and its bias record) towards L2 a couple of timesteps ahead of use: on HBM-resident
// tables every row visit is a DRAM round trip that nothing else in the warp can hide.
template <int D>
__device__ __forceinline__ void prefetch_item(const ModelDev& m, uint32_t id, int lane, int vectors) {
    // local shards only: prefetch.global.L2 on a peer (NVLink) address stalls for tens of microseconds (measured:
    // 25x slowdown of the whole kernel), and the peer's L2 is not ours to fill anyway
    if (m.own_shard >= 0 && (int)(id & m.gmask) != m.own_shard) return;
    const char* rec = reinterpret_cast<const char*>(item_rec(m, id));
    const int lines = (vectors * D * 4 + 127) / 128;
    if (lane < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + lane * 128));
    else if (lane == 31) asm volatile("prefetch.global.L2 [%0];" ::"l"(bias_rec(m, id)));
}

// dense (non-embedding) parameter visit by one warp: element i of [w | s1 | s2] arrays of length nd
template <int D>
__device__ __forceinline__ void update_dense_vec(float* dense, size_t nd, size_t off, int lane, const float (&g)[VecOf<D>::V],
                                                 const OptCfg& o) {
    constexpr int V = VecOf<D>::V;
    float w[V], s1[V], s2[V];
    row_load_cg<D>(dense + off, lane, w);
    row_load_cg<D>(dense + nd + off, lane, s1);
    if (o.adam) row_load_cg<D>(dense + 2 * nd + off, lane, s2);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (o.adam) adam_elem(w[v], s1[v], s2[v], g[v], o);
        else adagrad_elem(w[v], s1[v], g[v], o.lr, o.l2);
    }
    row_store_cg<D>(dense + off, lane, w);
    row_store_cg<D>(dense + nd + off, lane, s1);
    if (o.adam) row_store_cg<D>(dense + 2 * nd + off, lane, s2);
}

// =====================================================================================================
// EWMA (ewma.rs:266-352): s_0 = x_0, s_t = a*s_{t-1} + (1-a)*x_t, a = sigmoid(alpha)
// scratch per warp: S[T][D] X[T][D] DQ[T][D] G[T] NEG[T]
// =====================================================================================================
template <int D>
__global__ void __launch_bounds__(256) ewma_train_kernel(ModelDev m, PlanDev pl) {
    constexpr int V = VecOf<D>::V;
    constexpr int kPF = 2;  // prefetch distance in timesteps
    const int lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= pl.P) return;
    const int T = m.T;
    float* ws = pl.scratch + (size_t)p * pl.scratch_stride;
    float* S_ = ws; float* X_ = ws + (size_t)T * D; float* DQ = ws + 2 * (size_t)T * D;
    float* G_ = ws + 3 * (size_t)T * D; uint32_t* NEG = reinterpret_cast<uint32_t*>(G_ + T);

    XorShift rng = pl.rng[p];
    const uint64_t key = pl.keys[p];
    uint64_t step = pl.step_ctr[p];
    uint32_t* ord = pl.order + (size_t)p * pl.n;
    float loss_acc = 0.0f; unsigned long long ex = 0;
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;
    // Hogwild with more than one partition: Adagrad visits go through L2 atomics (common.cuh); one partition keeps
    // the plain load / store visit, which is the reference's single-thread arithmetic to the last bit
    const bool atomics = !o.adam && pl.P > 1 && !m.hbm_resident;   // (HBM-resident tables: no hot rows, plain visits are 8 % faster)

    for (int ep = 0; ep < pl.epochs; ++ep) {
        if (lane == 0) shuffle_partition(ord, pl.n, rng);
        __syncwarp();
        for (uint32_t i = 0; i < pl.n; ++i, ++step) {
            const uint32_t sq = ord[i];
            const uint32_t* ids = pl.item_ids + pl.seq_start[sq];
            const int Tn = (int)pl.seq_len[sq] - 1;
            adam_corrections(o, pl.adam_t0 + step * pl.P + p + 1);

            float al[V], a[V], s[V];
            row_load_cg<D>(m.dense, lane, al);
#pragma unroll
            for (int v = 0; v < V; ++v) { a[v] = sigmoidf_(al[v]); s[v] = 0.0f; }

            float loss_seq = 0.0f;
            // ---- forward: the three rows of timestep t+1 (input, target, first negative candidate -- all known in
            // advance) are loaded while timestep t is being computed, so a row visit costs one exposed round trip at
            // most, local HBM or a peer GPU over NVLink alike ----
            float xn[V], pn[V], qn[V], bpn = 0.0f, bqn = 0.0f; uint32_t candn = 0;
            auto issue_rows = [&](int t) {
                const uint32_t in = __ldg(ids + t), out = __ldg(ids + t + 1);
                candn = draw_item(key, step, (uint32_t)t, 0u, pl.neg_range);
                row_load_cg<D>(item_rec(m, in), lane, xn);     // item_embeddings.index(input)
                row_load_cg<D>(item_rec(m, out), lane, pn);
                row_load_cg<D>(item_rec(m, candn), lane, qn);
                bpn = __ldcg(reinterpret_cast<const float*>(bias_rec(m, out)));
                bqn = __ldcg(reinterpret_cast<const float*>(bias_rec(m, candn)));
            };
            if (Tn > 0) issue_rows(0);
            for (int t = 0; t < Tn; ++t) {
                float x[V], pv[V], qv[V];
#pragma unroll
                for (int v = 0; v < V; ++v) { x[v] = xn[v]; pv[v] = pn[v]; qv[v] = qn[v]; }
                const float bp = bpn; float bq = bqn; uint32_t neg = candn;
                if (m.hbm_resident && t + kPF < Tn) {
                    prefetch_item<D>(m, __ldg(ids + t + kPF + 1), lane, 1);
                    prefetch_item<D>(m, draw_item(key, step, (uint32_t)(t + kPF), 0u, pl.neg_range), lane, 1);
                }
                if (t + 1 < Tn) issue_rows(t + 1);
#pragma unroll
                for (int v = 0; v < V; ++v) s[v] = t == 0 ? x[v] : a[v] * s[v] + (1.0f - a[v]) * x[v];
                vec_store<D>(S_ + (size_t)t * D, lane, s);
                vec_store<D>(X_ + (size_t)t * D, lane, x);
                const float pos = warp_dot<D>(s, pv) + bp;
                float ngs = warp_dot<D>(s, qv) + bq;
                if (m.loss == 2) {  // WARP: further candidates only while the current one is not a violator (sequence_model.rs:58-65)
                    for (int j = 1; j < 5 && !(1.0f - pos + ngs > 0.0f); ++j) {
                        neg = draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range);
                        row_load_cg<D>(item_rec(m, neg), lane, qv);
                        ngs = warp_dot<D>(s, qv) + __ldcg(reinterpret_cast<const float*>(bias_rec(m, neg)));
                    }
                }
                StepOut lo = pair_loss(m.loss, pos, ngs);
                loss_seq += lo.loss;
                float dq[V];
#pragma unroll
                for (int v = 0; v < V; ++v) dq[v] = lo.g * (qv[v] - pv[v]);
                vec_store<D>(DQ + (size_t)t * D, lane, dq);
                if (lane == 0) { G_[t] = lo.g; NEG[t] = neg; }
            }
            __syncwarp();
            // ---- backward + sparse optimizer visits (t descending: E[neg], E[out], E[in]) ----
            float ds[V], da[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { ds[v] = 0.0f; da[v] = 0.0f; }
            for (int t = Tn - 1; t >= 0; --t) {
                const uint32_t in = __ldg(ids + t), out = __ldg(ids + t + 1);
                const uint32_t neg = NEG[t]; const float g = G_[t];
                if (m.hbm_resident && t >= kPF) {  // full records (weights + optimizer state) of timestep t - kPF
                    prefetch_item<D>(m, NEG[t - kPF], lane, m.S);
                    prefetch_item<D>(m, __ldg(ids + t - kPF + 1), lane, m.S);
                    prefetch_item<D>(m, __ldg(ids + t - kPF), lane, m.S);
                }
                // when the three rows are distinct their records are loaded together (one round trip instead of three)
                const bool distinct = neg != out && neg != in && out != in;
                float* rn = item_rec(m, neg); float* ro = item_rec(m, out); float* ri = item_rec(m, in);
                float wn[V], gnn[V], wo[V], goo[V], wi[V], gii[V], vn[V], vo[V], vi[V];
                if (distinct && !atomics) {
                    row_load_cg<D>(rn, lane, wn); row_load_cg<D>(rn + D, lane, gnn);
                    row_load_cg<D>(ro, lane, wo); row_load_cg<D>(ro + D, lane, goo);
                    row_load_cg<D>(ri, lane, wi); row_load_cg<D>(ri + D, lane, gii);
                    if (o.adam) { row_load_cg<D>(rn + 2 * D, lane, vn); row_load_cg<D>(ro + 2 * D, lane, vo); row_load_cg<D>(ri + 2 * D, lane, vi); }
                }
                float st[V], dq[V], dh[V], dx[V], gn[V], gp[V];
                vec_load<D>(S_ + (size_t)t * D, lane, st);
                vec_load<D>(DQ + (size_t)t * D, lane, dq);
#pragma unroll
                for (int v = 0; v < V; ++v) dh[v] = ds[v] + dq[v];
                if (t == 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) { dx[v] = dh[v]; ds[v] = 0.0f; }
                } else {
                    float sp[V], x[V];
                    vec_load<D>(S_ + (size_t)(t - 1) * D, lane, sp);
                    vec_load<D>(X_ + (size_t)t * D, lane, x);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        dx[v] = (1.0f - a[v]) * dh[v];
                        da[v] += dh[v] * (sp[v] - x[v]);
                        ds[v] = a[v] * dh[v];
                    }
                }
#pragma unroll
                for (int v = 0; v < V; ++v) { gn[v] = g * st[v]; gp[v] = -g * st[v]; }
                if (atomics) {
                    if (distinct) {   // the three visits share one round trip: {ld w, atom G += g^2} x 3, then the reductions
                        float qn[V], qo[V], qi[V];
#pragma unroll
                        for (int v = 0; v < V; ++v) { qn[v] = gn[v] * gn[v]; qo[v] = qn[v]; qi[v] = dx[v] * dx[v]; }
                        row_load_cg<D>(rn, lane, wn); row_atom_add<D>(rn + D, lane, qn, gnn);
                        row_load_cg<D>(ro, lane, wo); row_atom_add<D>(ro + D, lane, qo, goo);
                        row_load_cg<D>(ri, lane, wi); row_atom_add<D>(ri + D, lane, qi, gii);
                        float dwn[V], dGn[V], dwo[V], dGo[V], dwi[V], dGi[V];
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            adagrad_atomic_elem(wn[v], gnn[v], gn[v], qn[v], o.lr, o.l2, dwn[v], dGn[v]);
                            adagrad_atomic_elem(wo[v], goo[v], gp[v], qo[v], o.lr, o.l2, dwo[v], dGo[v]);
                            adagrad_atomic_elem(wi[v], gii[v], dx[v], qi[v], o.lr, o.l2, dwi[v], dGi[v]);
                        }
                        row_red_add<D>(rn, lane, dwn); row_red_add<D>(ro, lane, dwo); row_red_add<D>(ri, lane, dwi);
                        if (o.l2 != 0.0f) { row_red_add<D>(rn + D, lane, dGn); row_red_add<D>(ro + D, lane, dGo); row_red_add<D>(ri + D, lane, dGi); }
                    } else {
                        update_row_atomic<D>(rn, lane, gn, o);
                        update_row_atomic<D>(ro, lane, gp, o);
                        update_row_atomic<D>(ri, lane, dx, o);
                    }
                    if (lane == 0) {
                        update_bias_atomic(bias_rec(m, neg), g, o);
                        update_bias_atomic(bias_rec(m, out), -g, o);
                    }
                } else if (distinct) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (!o.adam) {
                            adagrad_elem(wn[v], gnn[v], gn[v], o.lr, o.l2); adagrad_elem(wo[v], goo[v], gp[v], o.lr, o.l2);
                            adagrad_elem(wi[v], gii[v], dx[v], o.lr, o.l2);
                        } else {
                            adam_elem(wn[v], gnn[v], vn[v], gn[v], o); adam_elem(wo[v], goo[v], vo[v], gp[v], o);
                            adam_elem(wi[v], gii[v], vi[v], dx[v], o);
                        }
                    }
                    row_store_cg<D>(rn, lane, wn); row_store_cg<D>(rn + D, lane, gnn);
                    row_store_cg<D>(ro, lane, wo); row_store_cg<D>(ro + D, lane, goo);
                    row_store_cg<D>(ri, lane, wi); row_store_cg<D>(ri + D, lane, gii);
                    if (o.adam) { row_store_cg<D>(rn + 2 * D, lane, vn); row_store_cg<D>(ro + 2 * D, lane, vo); row_store_cg<D>(ri + 2 * D, lane, vi); }
                } else {  // repeated item inside the timestep: strictly sequential visits
                    update_row<D>(rn, lane, gn, o);
                    update_row<D>(ro, lane, gp, o);
                    update_row<D>(ri, lane, dx, o);
                }
                if (lane == 0 && !atomics) {
                    update_bias(bias_rec(m, neg), g, o);
                    update_bias(bias_rec(m, out), -g, o);
                }
                __syncwarp();
            }
            // ---- dense: alpha ----
            float dal[V];
#pragma unroll
            for (int v = 0; v < V; ++v) dal[v] = da[v] * a[v] * (1.0f - a[v]);
            update_dense_vec<D>(m.dense, m.ndense, 0, lane, dal, o);
            loss_acc += loss_seq; ex += (unsigned long long)Tn;
        }
    }
    if (lane == 0) {
        pl.rng[p] = rng; pl.step_ctr[p] = step;
        pl.loss_acc[p] += loss_acc; pl.examples[p] += ex;
    }
}

// =====================================================================================================
// LSTM, FFMA path, D in {16, 32} (lstm.rs:258-337 + wyrm::nn::lstm).  Weights live in shared memory as
// Ws[k][d] = float4{f,i,g,o}; the CTA's warps step in lock-step rounds: after each round the CTA-summed dense
// gradient is applied to the global weights (Hogwild across CTAs) and the shared copy is refreshed.
// With num_threads == 1 this is exactly the reference's one-dense-step-per-sequence order.
// scratch per warp: [T][8][D] = x,h,c,f,i,g,o,dq ; then G[T], NEG[T]
// =====================================================================================================
template <int NV>
__device__ __forceinline__ void warp_multi_reduce(float (&v)[NV], int lane) {
    // Reduces NV independent sums across the 32 lanes with NV-NV/32 shuffles; afterwards lane L holds, in
    // v[0..NV/32), the complete sums of the slots that were stored at positions (NV/32)*L + i.
    int n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const int half = n >> 1;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
            if (i < half) {
                float keep = up ? v[i + half] : v[i];
                float send = up ? v[i] : v[i + half];
                v[i] = keep + __shfl_xor_sync(kFull, send, off);
            }
        }
        n = half;
    }
}

template <int D, int WPC>
__global__ void __launch_bounds__(WPC * 32) lstm_train_kernel(ModelDev m, PlanDev pl) {
    static_assert(D == 16 || D == 32, "FFMA LSTM path supports D in {16, 32}");
    constexpr int NK = 2 * D;       // rows of W: [h ; x]
    constexpr int R = NK / 32;      // sums per lane after the multi-reduce
    extern __shared__ float4 smem4[];
    float4* Ws = smem4;                       // [NK][D]
    float4* dWs = Ws + NK * D;                // [NK][D]
    float4* Bs = dWs + NK * D;                // [D]
    float4* dBs = Bs + D;                     // [D]
    float* zbuf = reinterpret_cast<float*>(dBs + D);  // [WPC][NK]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool act = lane < D;
    const int ld = act ? lane : 0;
    const uint32_t p = blockIdx.x * WPC + warp;
    const bool live = p < pl.P;
    const int T = m.T;
    const size_t nd = m.ndense;
    const bool coupled = m.variant == 1;
    float* myz = zbuf + warp * NK;
    float* ws = pl.scratch + (size_t)(live ? p : 0) * pl.scratch_stride;
    float* G_ = ws + (size_t)T * 8 * D; uint32_t* NEG = reinterpret_cast<uint32_t*>(G_ + T);
    auto slot = [&](int t, int which) { return ws + ((size_t)t * 8 + which) * D; };
    enum { SX = 0, SH = 1, SC = 2, SF = 3, SI = 4, SG = 5, SO = 6, SDQ = 7 };

    // stage weights: global canonical W[k][q][d], B[q][d]
    for (int idx = threadIdx.x; idx < NK * D; idx += WPC * 32) {
        const int k = idx / D, d = idx % D;
        const float* src = m.dense + (size_t)k * 4 * D + d;
        Ws[idx] = make_float4(__ldcg(src), __ldcg(src + D), __ldcg(src + 2 * D), __ldcg(src + 3 * D));
        dWs[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int d = threadIdx.x; d < D; d += WPC * 32) {
        const float* src = m.dense + (size_t)NK * 4 * D + d;
        Bs[d] = make_float4(__ldcg(src), __ldcg(src + D), __ldcg(src + 2 * D), __ldcg(src + 3 * D));
        dBs[d] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    XorShift rng; uint64_t key = 0; uint32_t* ord = nullptr;
    uint64_t step = pl.step_ctr[live ? p : 0];  // all partitions have done the same number of steps
    if (live) { rng = pl.rng[p]; key = pl.keys[p]; ord = pl.order + (size_t)p * pl.n; }
    float loss_acc = 0.0f; unsigned long long ex = 0;
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;

    for (int ep = 0; ep < pl.epochs; ++ep) {
        if (live && lane == 0) shuffle_partition(ord, pl.n, rng);
        __syncwarp();
        for (uint32_t i = 0; i < pl.n; ++i, ++step) {
            adam_corrections(o, pl.adam_t0 + step * pl.P + (live ? p : blockIdx.x * WPC) + 1);
            if (live) {
                const uint32_t sq = ord[i];
                const uint32_t* ids = pl.item_ids + pl.seq_start[sq];
                const int Tn = (int)pl.seq_len[sq] - 1;
                float h = 0.0f, c = 0.0f, loss_seq = 0.0f;
                // ---------------- forward ----------------
                for (int t = 0; t < Tn; ++t) {
                    const uint32_t in = __ldg(ids + t), out = __ldg(ids + t + 1);
                    float x[1], hv[1], pv[1], qv[1];
                    row_load_cg<D>(item_rec(m, in), lane, x);
                    if (act) { myz[lane] = h; myz[D + lane] = x[0]; }
                    __syncwarp();
                    float4 pre = Bs[ld];
#pragma unroll 4
                    for (int k4 = 0; k4 < NK / 4; ++k4) {
                        const float4 z4 = reinterpret_cast<const float4*>(myz)[k4];
                        const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const float4 w4 = Ws[(k4 * 4 + kk) * D + ld];
                            pre.x = fmaf(zz[kk], w4.x, pre.x); pre.y = fmaf(zz[kk], w4.y, pre.y);
                            pre.z = fmaf(zz[kk], w4.z, pre.z); pre.w = fmaf(zz[kk], w4.w, pre.w);
                        }
                    }
                    __syncwarp();
                    const float f = sigmoidf_(pre.x);
                    const float ig = coupled ? 1.0f - f : sigmoidf_(pre.y);
                    const float gg = tanhf(pre.z);
                    const float og = sigmoidf_(pre.w);
                    const float cn = f * c + ig * gg;
                    const float tc = tanhf(cn);
                    c = act ? cn : 0.0f;
                    h = act ? og * tc : 0.0f;
                    if (act) {
                        slot(t, SX)[lane] = x[0]; slot(t, SH)[lane] = h; slot(t, SC)[lane] = c;
                        slot(t, SF)[lane] = f; slot(t, SI)[lane] = ig; slot(t, SG)[lane] = gg; slot(t, SO)[lane] = og;
                    }
                    hv[0] = h;
                    uint32_t neg; float pos, ngs;
                    score_and_sample<D>(m, lane, hv, out, key, step, (uint32_t)t, pl.neg_range, pv, qv, neg, pos, ngs);
                    StepOut lo = pair_loss(m.loss, pos, ngs);
                    loss_seq += lo.loss;
                    if (act) slot(t, SDQ)[lane] = lo.g * (qv[0] - pv[0]);
                    if (lane == 0) { G_[t] = lo.g; NEG[t] = neg; }
                }
                __syncwarp();
                // ---------------- backward ----------------
                float dh_rec = 0.0f, dc_rec = 0.0f;
                float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int t = Tn - 1; t >= 0; --t) {
                    const uint32_t in = __ldg(ids + t), out = __ldg(ids + t + 1);
                    const uint32_t neg = NEG[t]; const float g = G_[t];
                    float ht = 0.f, f = 0.f, ig = 0.f, gg = 0.f, og = 0.f, ct = 0.f, cp = 0.f, dq = 0.f;
                    if (act) {
                        ht = slot(t, SH)[lane]; ct = slot(t, SC)[lane]; f = slot(t, SF)[lane]; ig = slot(t, SI)[lane];
                        gg = slot(t, SG)[lane]; og = slot(t, SO)[lane]; dq = slot(t, SDQ)[lane];
                        cp = t ? slot(t - 1, SC)[lane] : 0.0f;
                    }
                    const float tc = tanhf(ct);
                    const float dh = dh_rec + dq;
                    const float d_o = dh * tc;
                    const float dc = dc_rec + dh * og * (1.0f - tc * tc);
                    float d_f = dc * cp, d_i = dc * gg;
                    const float d_g = dc * ig;
                    dc_rec = dc * f;
                    if (coupled) { d_f -= d_i; d_i = 0.0f; }
                    float4 del;
                    del.x = d_f * f * (1.0f - f);
                    del.y = coupled ? 0.0f : d_i * ig * (1.0f - ig);
                    del.z = d_g * (1.0f - gg * gg);
                    del.w = d_o * og * (1.0f - og);
                    if (!act) del = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (act) {  // deltas replace the gates (consumed by the dW pass below)
                        slot(t, SF)[lane] = del.x; slot(t, SI)[lane] = del.y; slot(t, SG)[lane] = del.z; slot(t, SO)[lane] = del.w;
                    }
                    db.x += del.x; db.y += del.y; db.z += del.z; db.w += del.w;
                    // dz[k] = sum_{q,d} del[q][d] W[k][q][d]: lane-partials then a 32-lane multi-reduce
                    float part[NK];
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        const float4 w4 = Ws[k * D + ld];
                        float a0 = del.x * w4.x;
                        a0 = fmaf(del.y, w4.y, a0); a0 = fmaf(del.z, w4.z, a0); a0 = fmaf(del.w, w4.w, a0);
                        part[R * (k % 32) + k / 32] = a0;
                    }
                    warp_multi_reduce<NK>(part, lane);
                    float dx;
                    if constexpr (D == 32) { dh_rec = part[0]; dx = part[1]; }
                    else { dx = __shfl_sync(kFull, part[0], (lane + 16) & 31); dh_rec = act ? part[0] : 0.0f; }
                    float gn[1] = {g * ht}, gp[1] = {-g * ht}, gx[1] = {dx};
                    update_row<D>(item_rec(m, neg), lane, gn, o);
                    update_row<D>(item_rec(m, out), lane, gp, o);
                    update_row<D>(item_rec(m, in), lane, gx, o);
                    if (lane == 0) {
                        update_bias(bias_rec(m, neg), g, o);
                        update_bias(bias_rec(m, out), -g, o);
                    }
                    __syncwarp();
                }
                // ---------------- dW = sum_t z_t^T delta_t, 8 rows of W at a time ----------------
                if (act) {
                    atomicAdd(&dBs[lane].x, db.x); atomicAdd(&dBs[lane].y, db.y);
                    atomicAdd(&dBs[lane].z, db.z); atomicAdd(&dBs[lane].w, db.w);
                }
                for (int kc = 0; kc < NK / 8; ++kc) {
                    float acc[8][4];
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) { acc[kk][0] = acc[kk][1] = acc[kk][2] = acc[kk][3] = 0.0f; }
                    const bool hpart = kc * 8 < D;
                    const int off = hpart ? kc * 8 : kc * 8 - D;
                    for (int t = Tn - 1; t >= 0; --t) {
                        float4 z0, z1;
                        if (hpart && t == 0) { z0 = make_float4(0.f, 0.f, 0.f, 0.f); z1 = z0; }
                        else {
                            const float* zs = (hpart ? slot(t - 1, SH) : slot(t, SX)) + off;
                            z0 = *reinterpret_cast<const float4*>(zs); z1 = *reinterpret_cast<const float4*>(zs + 4);
                        }
                        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                        float dl[4] = {0.f, 0.f, 0.f, 0.f};
                        if (act) { dl[0] = slot(t, SF)[lane]; dl[1] = slot(t, SI)[lane]; dl[2] = slot(t, SG)[lane]; dl[3] = slot(t, SO)[lane]; }
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                            for (int q = 0; q < 4; ++q) acc[kk][q] = fmaf(zz[kk], dl[q], acc[kk][q]);
                    }
                    if (act) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            float4* dst = &dWs[(kc * 8 + kk) * D + lane];
                            atomicAdd(&dst->x, acc[kk][0]); atomicAdd(&dst->y, acc[kk][1]);
                            atomicAdd(&dst->z, acc[kk][2]); atomicAdd(&dst->w, acc[kk][3]);
                        }
                    }
                }
                loss_acc += loss_seq; ex += (unsigned long long)Tn;
            }
            // ---------------- CTA round: dense optimizer step on the CTA-summed gradient ----------------
            __syncthreads();
            for (int idx = threadIdx.x; idx < (int)nd; idx += WPC * 32) {
                float gsum; float* sw; float* sdw;
                if (idx < NK * 4 * D) {
                    const int k = idx / (4 * D), q = (idx / D) % 4, d = idx % D;
                    sw = reinterpret_cast<float*>(&Ws[k * D + d]) + q; sdw = reinterpret_cast<float*>(&dWs[k * D + d]) + q;
                } else {
                    const int j = idx - NK * 4 * D, q = j / D, d = j % D;
                    sw = reinterpret_cast<float*>(&Bs[d]) + q; sdw = reinterpret_cast<float*>(&dBs[d]) + q;
                }
                gsum = *sdw; *sdw = 0.0f;
                float w = __ldcg(m.dense + idx), s1 = __ldcg(m.dense + nd + idx);
                if (o.adam) {
                    float s2 = __ldcg(m.dense + 2 * nd + idx);
                    adam_elem(w, s1, s2, gsum, o);
                    __stcg(m.dense + 2 * nd + idx, s2);
                } else adagrad_elem(w, s1, gsum, o.lr, o.l2);
                __stcg(m.dense + idx, w); __stcg(m.dense + nd + idx, s1);
                *sw = w;
            }
            __syncthreads();
        }
    }
    if (live && lane == 0) {
        pl.rng[p] = rng; pl.step_ctr[p] = step;
        pl.loss_acc[p] += loss_acc; pl.examples[p] += ex;
    }
}


// =====================================================================================================
// LSTM, generic-width FFMA path, D in {64, 128, 256} (BASELINE configs C3, C5).  One warp == one partition; lane l owns
// units l*V .. l*V+V-1 (V = D/32).  The weights (8*D*D floats: 131 KB .. 2 MB) do not fit shared memory, so they are read
// from L2 (ld.cg: they are Hogwild-shared) as coalesced V-float vectors, three passes per timestep (gates, dz, dW).
// The dense optimizer step is applied per sequence by the warp itself, rows of W at a time, exactly like the
// reference (one optimizer.step per sub-sequence, sequence_model.rs:163-169).  Correct and parity-tested; the
// tensor-core tile kernel (kernels_lstm_tile.cu) is the fast path and currently covers D = 32 only.
// scratch per warp: [T][8][D] = x,h,c,f,i,g,o,dq ; then G[T], NEG[T]
// =====================================================================================================
template <int D, int WPC>
__global__ void __launch_bounds__(WPC * 32) lstm_wide_train_kernel(ModelDev m, PlanDev pl) {
    constexpr int V = D / 32, NK = 2 * D;
    constexpr int KC = 64 / (4 * V) > 0 ? 64 / (4 * V) : 1;   // rows of W per dW chunk: 4*V*KC = 64 accumulators
    extern __shared__ float smem_w[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * WPC + warp;
    if (p >= pl.P) return;
    float* zbuf = smem_w + warp * (NK + NK);       // [NK] z = [h, x]
    float* dzbuf = zbuf + NK;                      // [NK] dz
    const int T = m.T;
    const size_t nd = m.ndense;
    const bool coupled = m.variant == 1;
    float* ws = pl.scratch + (size_t)p * pl.scratch_stride;
    float* G_ = ws + (size_t)T * 8 * D; uint32_t* NEG = reinterpret_cast<uint32_t*>(G_ + T);
    auto slot = [&](int t, int which) { return ws + ((size_t)t * 8 + which) * D; };
    enum { SX = 0, SH = 1, SC = 2, SF = 3, SI = 4, SG = 5, SO = 6, SDQ = 7 };
    const float* W = m.dense; const float* Bv = m.dense + (size_t)NK * 4 * D;

    XorShift rng = pl.rng[p];
    const uint64_t key = pl.keys[p];
    uint64_t step = pl.step_ctr[p];
    uint32_t* ord = pl.order + (size_t)p * pl.n;
    float loss_acc = 0.0f; unsigned long long ex = 0;
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;

    for (int ep = 0; ep < pl.epochs; ++ep) {
        if (lane == 0) shuffle_partition(ord, pl.n, rng);
        __syncwarp();
        for (uint32_t it = 0; it < pl.n; ++it, ++step) {
            const uint32_t sq = ord[it];
            const uint32_t* ids = pl.item_ids + pl.seq_start[sq];
            const int Tn = (int)pl.seq_len[sq] - 1;
            adam_corrections(o, pl.adam_t0 + step * pl.P + p + 1);
            float h[V], c[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { h[v] = 0.0f; c[v] = 0.0f; }
            float loss_seq = 0.0f;
            // ---------------- forward ----------------
            for (int t = 0; t < Tn; ++t) {
                const uint32_t in = __ldg(ids + t), out = __ldg(ids + t + 1);
                float x[V], pre[4][V], pv[V], qv[V];
                row_load_cg<D>(item_rec(m, in), lane, x);
                vec_store<D>(zbuf, lane, h); vec_store<D>(zbuf + D, lane, x);
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; ++q) row_load_cg<D>(Bv + q * D, lane, pre[q]);
                for (int k = 0; k < NK; ++k) {
                    const float zk = zbuf[k];
                    float w[V];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        row_load_cg<D>(W + ((size_t)k * 4 + q) * D, lane, w);
#pragma unroll
                        for (int v = 0; v < V; ++v) pre[q][v] = fmaf(zk, w[v], pre[q][v]);
                    }
                }
                __syncwarp();
                float f[V], ig[V], gg[V], og[V];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    f[v] = sigmoidf_(pre[0][v]);
                    ig[v] = coupled ? 1.0f - f[v] : sigmoidf_(pre[1][v]);
                    gg[v] = tanhf(pre[2][v]);
                    og[v] = sigmoidf_(pre[3][v]);
                    c[v] = f[v] * c[v] + ig[v] * gg[v];
                    h[v] = og[v] * tanhf(c[v]);
                }
                vec_store<D>(slot(t, SX), lane, x); vec_store<D>(slot(t, SH), lane, h); vec_store<D>(slot(t, SC), lane, c);
                vec_store<D>(slot(t, SF), lane, f); vec_store<D>(slot(t, SI), lane, ig); vec_store<D>(slot(t, SG), lane, gg);
                vec_store<D>(slot(t, SO), lane, og);
                uint32_t neg; float pos, ngs;
                score_and_sample<D>(m, lane, h, out, key, step, (uint32_t)t, pl.neg_range, pv, qv, neg, pos, ngs);
                StepOut lo = pair_loss(m.loss, pos, ngs);
                loss_seq += lo.loss;
                float dq[V];
#pragma unroll
                for (int v = 0; v < V; ++v) dq[v] = lo.g * (qv[v] - pv[v]);
                vec_store<D>(slot(t, SDQ), lane, dq);
                if (lane == 0) { G_[t] = lo.g; NEG[t] = neg; }
            }
            __syncwarp();
            // ---------------- backward ----------------
            float dh_rec[V], dc_rec[V], db[4][V];
#pragma unroll
            for (int v = 0; v < V; ++v) { dh_rec[v] = 0.0f; dc_rec[v] = 0.0f; db[0][v] = db[1][v] = db[2][v] = db[3][v] = 0.0f; }
            for (int t = Tn - 1; t >= 0; --t) {
                const uint32_t in = __ldg(ids + t), out = __ldg(ids + t + 1);
                const uint32_t neg = NEG[t]; const float g = G_[t];
                float ht[V], ct[V], cp[V], f[V], ig[V], gg[V], og[V], dq[V], del[4][V];
                vec_load<D>(slot(t, SH), lane, ht); vec_load<D>(slot(t, SC), lane, ct); vec_load<D>(slot(t, SF), lane, f);
                vec_load<D>(slot(t, SI), lane, ig); vec_load<D>(slot(t, SG), lane, gg); vec_load<D>(slot(t, SO), lane, og);
                vec_load<D>(slot(t, SDQ), lane, dq);
                if (t > 0) vec_load<D>(slot(t - 1, SC), lane, cp);
                else {
#pragma unroll
                    for (int v = 0; v < V; ++v) cp[v] = 0.0f;
                }
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float tc = tanhf(ct[v]);
                    const float dh = dh_rec[v] + dq[v];
                    const float d_o = dh * tc;
                    const float dc = dc_rec[v] + dh * og[v] * (1.0f - tc * tc);
                    float d_f = dc * cp[v], d_i = dc * gg[v];
                    const float d_g = dc * ig[v];
                    dc_rec[v] = dc * f[v];
                    if (coupled) { d_f -= d_i; d_i = 0.0f; }
                    del[0][v] = d_f * f[v] * (1.0f - f[v]);
                    del[1][v] = coupled ? 0.0f : d_i * ig[v] * (1.0f - ig[v]);
                    del[2][v] = d_g * (1.0f - gg[v] * gg[v]);
                    del[3][v] = d_o * og[v] * (1.0f - og[v]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) db[q][v] += del[q][v];
                }
                // deltas replace the gates in scratch (consumed by the dW pass)
                vec_store<D>(slot(t, SF), lane, del[0]); vec_store<D>(slot(t, SI), lane, del[1]);
                vec_store<D>(slot(t, SG), lane, del[2]); vec_store<D>(slot(t, SO), lane, del[3]);
                // dz[k] = sum_{q,d} del[q][d] W[k][q][d]: 64 rows of W per multi-reduce
                for (int kb = 0; kb < NK; kb += 64) {
                    float part[64];
#pragma unroll
                    for (int kk = 0; kk < 64; ++kk) {
                        float a0 = 0.0f, w[V];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            row_load_cg<D>(W + ((size_t)(kb + kk) * 4 + q) * D, lane, w);
#pragma unroll
                            for (int v = 0; v < V; ++v) a0 = fmaf(del[q][v], w[v], a0);
                        }
                        part[2 * (kk % 32) + kk / 32] = a0;
                    }
                    warp_multi_reduce<64>(part, lane);
                    dzbuf[kb + lane] = part[0]; dzbuf[kb + 32 + lane] = part[1];
                }
                __syncwarp();
                float dx[V];
                vec_load<D>(dzbuf, lane, dh_rec); vec_load<D>(dzbuf + D, lane, dx);
                __syncwarp();
                float gn[V], gp[V];
#pragma unroll
                for (int v = 0; v < V; ++v) { gn[v] = g * ht[v]; gp[v] = -g * ht[v]; }
                update_row<D>(item_rec(m, neg), lane, gn, o);
                update_row<D>(item_rec(m, out), lane, gp, o);
                update_row<D>(item_rec(m, in), lane, dx, o);
                if (lane == 0) {
                    update_bias(bias_rec(m, neg), g, o);
                    update_bias(bias_rec(m, out), -g, o);
                }
                __syncwarp();
            }
            // ---------------- dense step: dW = sum_t z_t^T delta_t, KC rows of W at a time, applied in place ----------------
            for (int kb = 0; kb < NK; kb += KC) {
                float acc[KC][4][V];
#pragma unroll
                for (int kk = 0; kk < KC; ++kk)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[kk][q][v] = 0.0f;
                const bool hpart = kb < D;
                const int off = hpart ? kb : kb - D;
                for (int t = Tn - 1; t >= 0; --t) {
                    if (hpart && t == 0) continue;   // h_{-1} = 0
                    const float* zs = (hpart ? slot(t - 1, SH) : slot(t, SX)) + off;
                    float dl[4][V];
                    vec_load<D>(slot(t, SF), lane, dl[0]); vec_load<D>(slot(t, SI), lane, dl[1]);
                    vec_load<D>(slot(t, SG), lane, dl[2]); vec_load<D>(slot(t, SO), lane, dl[3]);
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) {
                        const float zk = zs[kk];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int v = 0; v < V; ++v) acc[kk][q][v] = fmaf(zk, dl[q][v], acc[kk][q][v]);
                    }
                }
#pragma unroll
                for (int kk = 0; kk < KC; ++kk)
#pragma unroll
                    for (int q = 0; q < 4; ++q) update_dense_vec<D>(m.dense, nd, ((size_t)(kb + kk) * 4 + q) * D, lane, acc[kk][q], o);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) update_dense_vec<D>(m.dense, nd, (size_t)NK * 4 * D + q * D, lane, db[q], o);
            loss_acc += loss_seq; ex += (unsigned long long)Tn;
        }
    }
    if (lane == 0) {
        pl.rng[p] = rng; pl.step_ctr[p] = step;
        pl.loss_acc[p] += loss_acc; pl.examples[p] += ex;
    }
}

template <int D, int WPC>
constexpr size_t lstm_smem_bytes() { return sizeof(float4) * (2 * (2 * D) * D + 2 * D) + sizeof(float) * WPC * 2 * D; }

constexpr int kLstmWPC = 8;
constexpr int kEwmaWPC = 8;

}  // namespace

bool train_supported(const ModelDev& m, const char** why) {
    if (m.model == MODEL_EWMA) {
        if (m.D == 16 || m.D == 32 || m.D == 64 || m.D == 128 || m.D == 256) return true;
        *why = "EWMA embedding_dim must be one of 16, 32, 64, 128, 256";
        return false;
    }
    if (m.D == 16 || m.D == 32 || m.D == 64 || m.D == 128 || m.D == 256) return true;
    *why = "LSTM embedding_dim must be one of 16, 32, 64, 128, 256";
    return false;
}

size_t train_scratch_floats_per_warp(const ModelDev& m) {
    size_t T = (size_t)m.T, D = (size_t)m.D;
    size_t n = (m.model == MODEL_EWMA ? 3 * T * D : 8 * T * D) + 2 * T;
    return (n + 31) / 32 * 32;  // keep every warp's block 128-byte aligned
}

int train_auto_partitions(const ModelDev& m, int num_sms) {
    // EWMA / FFMA LSTM: resident warps per SM (registers / shared memory); tensor-core LSTM: 2 tiles of 128 per SM
    if (m.D == 32 && !m.exact) return num_sms * (m.opt == 1 ? 128 : 256);   // tile kernels: 2 x 128 partitions per SM (Adam records: 1 x 128)
    if (m.model == MODEL_LSTM && m.D > 32) return m.exact ? num_sms * 8 : (m.D == 64 ? 32768 : 16384);   // wide LSTM: rounds of the batched tensor-core engine
    int per_sm = m.model == MODEL_EWMA ? (m.D <= 64 ? 32 : 16) : 16;
    return num_sms * per_sm;
}

// which LSTM kernel a plan runs on: 0 = FFMA (exact fp32, warp per partition), 1 = the tensor-core tile kernel
// (kernels_lstm_tile.cu), which takes whole tiles of 128 partitions; sbr_hyper_exact_arithmetic(h, 1) keeps the exact path.
int lstm_kernel_choice(const ModelDev& m, uint32_t P) {
    return (m.exact || m.model != MODEL_LSTM || m.D != 32 || P < 128 || P % 128 != 0) ? 0 : 1;
}

int launch_train(const ModelDev& m, const PlanDev& p, int num_sms, cudaStream_t st, cudaError_t* err, const char** kernel_name) {
    (void)num_sms;
    *err = cudaSuccess;
    const char* dummy; const char*& kn = kernel_name ? *kernel_name : dummy;
    kn = "";
    if (m.model == MODEL_EWMA && !m.exact && m.D == 32 && p.P >= 128 && p.P % 128 == 0) {
        // many partitions: thread-per-half-sequence kernel with bulk-copied records (kernels_ewma_tile.cu)
        *err = launch_ewma_tile(m, p, st);
        kn = m.opt == 1 ? "ewma_tile_train_kernel<1,3>" : ewma_tile_tiles_per_cta(m, p.P) == 2 ? "ewma_tile_train_kernel<2,2>" : "ewma_tile_train_kernel<1,2>";
        return 1;
    } else if (m.model == MODEL_EWMA) {
        kn = m.D == 16 ? "ewma_train_kernel<16>" : m.D == 32 ? "ewma_train_kernel<32>" : m.D == 64 ? "ewma_train_kernel<64>"
             : m.D == 128 ? "ewma_train_kernel<128>" : "ewma_train_kernel<256>";
        dim3 block(kEwmaWPC * 32), grid((p.P + kEwmaWPC - 1) / kEwmaWPC);
        switch (m.D) {
            case 16: ewma_train_kernel<16><<<grid, block, 0, st>>>(m, p); break;
            case 32: ewma_train_kernel<32><<<grid, block, 0, st>>>(m, p); break;
            case 64: ewma_train_kernel<64><<<grid, block, 0, st>>>(m, p); break;
            case 128: ewma_train_kernel<128><<<grid, block, 0, st>>>(m, p); break;
            case 256: ewma_train_kernel<256><<<grid, block, 0, st>>>(m, p); break;
            default: *err = cudaErrorInvalidValue; return 0;
        }
    } else if (lstm_kernel_choice(m, p.P)) {
        *err = launch_lstm_tile(m, p, st);
        kn = m.opt == 1 ? "lstm_tile_train_kernel<1,3>" : lstm_tile_tiles_per_cta(m, p.P) == 2 ? "lstm_tile_train_kernel<2,2>" : "lstm_tile_train_kernel<1,2>";
        return 1;
    } else {
        dim3 block(kLstmWPC * 32), grid((p.P + kLstmWPC - 1) / kLstmWPC);
        if (m.D == 32) {
            constexpr size_t smem = lstm_smem_bytes<32, kLstmWPC>();
            *err = cudaFuncSetAttribute(lstm_train_kernel<32, kLstmWPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (*err != cudaSuccess) return 0;
            lstm_train_kernel<32, kLstmWPC><<<grid, block, smem, st>>>(m, p);
            kn = "lstm_train_kernel<32,8>";
        } else if (m.D == 16) {
            constexpr size_t smem = lstm_smem_bytes<16, kLstmWPC>();
            *err = cudaFuncSetAttribute(lstm_train_kernel<16, kLstmWPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (*err != cudaSuccess) return 0;
            lstm_train_kernel<16, kLstmWPC><<<grid, block, smem, st>>>(m, p);
            kn = "lstm_train_kernel<16,8>";
        } else if (m.D == 64 || m.D == 128 || m.D == 256) {
            constexpr int WPCW = 4;
            dim3 blockw(WPCW * 32), gridw((p.P + WPCW - 1) / WPCW);
            const size_t smem = sizeof(float) * WPCW * 4 * m.D;
            if (m.D == 64) lstm_wide_train_kernel<64, WPCW><<<gridw, blockw, smem, st>>>(m, p);
            else if (m.D == 128) lstm_wide_train_kernel<128, WPCW><<<gridw, blockw, smem, st>>>(m, p);
            else lstm_wide_train_kernel<256, WPCW><<<gridw, blockw, smem, st>>>(m, p);
            kn = m.D == 64 ? "lstm_wide_train_kernel<64,4>" : m.D == 128 ? "lstm_wide_train_kernel<128,4>" : "lstm_wide_train_kernel<256,4>";
        } else { *err = cudaErrorInvalidValue; return 0; }
    }
    *err = cudaGetLastError();
    return 1;
}

}  // namespace sbr
