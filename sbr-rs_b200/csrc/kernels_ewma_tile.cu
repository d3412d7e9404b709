// kernels_ewma_tile.cu -- EWMA training kernel for many concurrent Hogwild partitions (D = 32): the throughput path of
// ImplicitEWMAModel::fit (ewma.rs:266-352 driven by sequence_model.rs:100-175).
//
// Same skeleton as the LSTM tile kernel (kernels_lstm_tile.cu) without its tensor-core part: a sequence is owned by two
// threads in two warps (owner `part` holds units [16 part, 16 part + 16) of every vector), item records {bias quad | w |
// G} arrive by TMA bulk copies into the sequence's shared-memory slots and optimizer steps leave as bulk reduce-adds.
// The EWMA recurrence itself is a handful of packed FMAs per timestep, so nothing hides a record's latency inside a
// step: the slots are DOUBLE-buffered and the records of timestep t+1 (t-1 in the backward pass) are requested while
// timestep t is being computed.  There is no block-wide barrier in the time loop -- the two owners of a sequence meet at
// a 64-thread named barrier to exchange their partial dot products, nothing else.
//
// Per timestep (ewma.rs:302-343):  s_0 = x_0, s_t = a s_{t-1} + (1 - a) x_t with a = sigmoid(alpha);
// pos = s_t . E[out_t] + b[out_t], neg = s_t . E[neg_t] + b[neg_t]; BPR sigmoid(neg - pos) or hinge max(0, 1 + neg - pos).
// Order of the sparse visits of one sub-sequence (as in the LSTM tile kernel):
//   forward,  t ascending : E[neg_t] += step(+g_t s_t), b[neg_t] += step(+g_t)
//   backward, t descending: E[ids[t+1]] += step(dx_{t+1}) then step(-g_t s_t), b[ids[t+1]] += step(-g_t)
//   then the dense step on alpha (Adagrad: reductions of the step and of the accumulator's increment; Adam: plain
//   read-modify-write).
#include <cuda_runtime.h>

#include "engine.h"
#include "tc_tile.cuh"

namespace sbr {

namespace {

using namespace tc;

constexpr uint32_t OFF_MISC = 0;                 // mbarriers, shard pointers
constexpr uint32_t OFF_TILES = 1024;
constexpr uint32_t PSLOT = 144;                  // bias quad + w
constexpr uint32_t XS_BYTES_PER_QUAD = 2 * 2 * 32 * 4;

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigm(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float2 fma2(const float2& a, const float2& b, const float2& c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 mul2(const float2& a, const float2& b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 add2(const float2& a, const float2& b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 bf2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { uint32_t o; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(hi), "f"(lo)); return o; }
__device__ __forceinline__ uint32_t word_of(const uint4& u, int pr) { return pr == 0 ? u.x : pr == 1 ? u.y : pr == 2 ? u.z : u.w; }

__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar_smem) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar_smem) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar_smem, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar_smem, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tEW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra ED_%=;\n\tbra EW_%=;\n\tED_%=:\n\t}\n" ::"r"(bar_smem),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void red_add4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct OptC { float lr, l2, c1, c2; float chat; };
// Adagrad on a pair of elements, as deltas: g' = g + l2 w ; dG = g'^2 ; dw = -lr g' / sqrt(G + dG)   (w, G are updated too)
__device__ __forceinline__ void adagrad2(float2& w, float2& G, const float2& g, const OptC& o, float2& dw, float2& dG) {
    const float2 gg = fma2(w, splat2(o.l2), g);
    dG = mul2(gg, gg);
    G = add2(G, dG);
    // cold start under massive concurrency (o.chat = expected concurrent visitors of a row, 0 when <= 1): a visitor that finds
    // the accumulator (nearly) empty is one of ~chat visitors that all see it empty; each steps as if its peers' g^2 were in
    const float2 rs = make_float2(rsqrt_approx(fmaxf(G.x, fmaxf(o.chat * dG.x, 1e-20f))), rsqrt_approx(fmaxf(G.y, fmaxf(o.chat * dG.y, 1e-20f))));
    dw = mul2(mul2(gg, splat2(-o.lr)), rs);
    w = add2(w, dw);
}
__device__ __forceinline__ void adagrad1(float& w, float& G, float g, const OptC& o) {
    g = fmaf(w, o.l2, g);
    G = fmaf(g, g, G);
    w = fmaf(-o.lr * g, rsqrt_approx(fmaxf(G, fmaxf(o.chat * g * g, 1e-20f))), w);
}
__device__ __forceinline__ void adam1(float& w, float& m, float& v, float g, const OptC& o) {
    g = fmaf(w, o.l2, g);
    m = 0.9f * m + 0.1f * g;
    v = 0.999f * v + 0.001f * g * g;
    const float mhat = __fdividef(m, o.c1), vhat = __fdividef(v, o.c2);
    w -= __fdividef(o.lr * mhat, sqrtf(vhat) + 1e-8f);
}
// one optimizer application on a 4-element chunk held as {w, s1, s2}.  Adagrad: the deltas are accumulated into dw, d1
// (they leave as reduce-adds: the accumulator is additive, no update is lost between partitions).  Adam: dw, d1, d2 end up
// holding the NEW values -- Adam's moments are decaying averages, m += 0.1 (g - m_seen) summed over K concurrent visitors
// of a hot row multiplies m by (1 - 0.1 K) and diverges for K > 20, so Adam records are written back with a plain bulk
// store (Hogwild in the reference's sense: concurrent visitors may overwrite each other, nothing can blow up).
template <bool ADAM>
__device__ __forceinline__ void apply_chunk(float4& w, float4& s1, float4& s2, const float4& g, const OptC& o, float4& dw, float4& d1, float4& d2) {
    if (!ADAM) {
        float2 wa = make_float2(w.x, w.y), wb = make_float2(w.z, w.w), Ga = make_float2(s1.x, s1.y), Gb = make_float2(s1.z, s1.w), ta, tb, ua, ub;
        adagrad2(wa, Ga, make_float2(g.x, g.y), o, ta, ua);
        adagrad2(wb, Gb, make_float2(g.z, g.w), o, tb, ub);
        w = make_float4(wa.x, wa.y, wb.x, wb.y); s1 = make_float4(Ga.x, Ga.y, Gb.x, Gb.y);
        dw.x += ta.x; dw.y += ta.y; dw.z += tb.x; dw.w += tb.y;
        d1.x += ua.x; d1.y += ua.y; d1.z += ub.x; d1.w += ub.y;
    } else {
        adam1(w.x, s1.x, s2.x, g.x, o); adam1(w.y, s1.y, s2.y, g.y, o); adam1(w.z, s1.z, s2.z, g.z, o); adam1(w.w, s1.w, s2.w, g.w, o);
        dw = w; d1 = s1; d2 = s2;
    }
}

struct Table { float* e0; float* const* es; uint32_t stride, gmask; int gshift; };
template <bool FLAT>
__device__ __forceinline__ float* trec(const Table& tb, uint32_t id) {   // base of the record (the bias quad; w follows at + 4 floats)
    if (FLAT) return tb.e0 + (size_t)id * tb.stride;
    return tb.es[id & tb.gmask] + (size_t)(id >> tb.gshift) * tb.stride;
}

template <int NT, int S, bool FLAT>
__global__ void __launch_bounds__(256 * NT, 1) ewma_tile_train_kernel(ModelDev m, PlanDev pl) {
    constexpr int DPT = 16, NCH = 4, TT = 256;
    constexpr bool ADAM = S == 3;
    constexpr uint32_t REC = 16 + S * 128;                 // bytes of an item record / record slot
    // P[2][128] target-row slots (forward), R[2][128] record slots; the backward pass rotates THREE record buffers: the third
    // one aliases the idle P slots when a record fits (Adagrad), else it follows R
    constexpr uint32_t TILE_P = 0, TILE_R = 2 * 128 * PSLOT;
    constexpr bool R3_ALIAS = 32 * REC <= 64 * PSLOT;
    constexpr uint32_t TILE_R3 = R3_ALIAS ? TILE_P : TILE_R + 2 * 128 * REC, TILE_BYTES = TILE_R + (R3_ALIAS ? 2 : 3) * 128 * REC;

    extern __shared__ __align__(1024) uint8_t smem[];
    float** es_s = reinterpret_cast<float**>(smem + OFF_MISC);               // [8] shard base pointers
    uint64_t* ldbar = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 64);     // [NT][4][3] record loads: per lane quarter, per buffer

    const int tid = threadIdx.x, tile = tid / TT, tt = tid % TT;
    const int wi = tt >> 5, q = wi & 3, part = wi >> 2, lane = tid & 31;
    const int r = q * 32 + lane;                     // sequence row of the tile
    uint8_t* Tb = smem + OFF_TILES + tile * TILE_BYTES;
    float* xs = reinterpret_cast<float*>(smem + OFF_TILES + NT * TILE_BYTES + (tile * 4 + q) * XS_BYTES_PER_QUAD);
    // P slots are grouped per lane quarter ([4][2][32] x 144 bytes): a quarter's 9216 bytes are its own third record buffer
    // in the backward pass (quarters are not synchronised with each other)
    auto pslot_g = [&](int b) -> const uint8_t* { return Tb + TILE_P + (size_t)((q * 2 + b) * 32 + lane) * PSLOT; };
    auto rslot_g = [&](int b) -> uint8_t* { return Tb + TILE_R + (size_t)(b * 128 + r) * REC; };
    auto rslot3_g = [&](int k) -> uint8_t* {
        if (k < 2) return Tb + TILE_R + (size_t)(k * 128 + r) * REC;
        return R3_ALIAS ? Tb + TILE_P + (size_t)q * (64 * PSLOT) + (size_t)lane * REC : Tb + TILE_R3 + (size_t)r * REC;
    };
    auto qbar = [&](int b) -> uint32_t { return smem_u32(ldbar + (tile * 4 + q) * 3 + b); };
    const uint32_t p = (blockIdx.x * NT + tile) * 128u + r;
    const bool live = p < pl.P;
    const bool lead = part == 0;
    const int T = m.T;
    if (tid < 8) es_s[tid] = m.Es[tid];
    Table tb; tb.e0 = m.Es[0]; tb.es = es_s; tb.stride = (uint32_t)rec_floats(m); tb.gmask = m.gmask; tb.gshift = m.gshift;

    auto quad_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + tile * 4 + q), "n"(64) : "memory"); };
    int xslot = 0;
    auto xsum = [&](float partial) -> float {   // sum of the two owners' partial values, identical (same order) in both
        xs[(xslot * 2 + part) * 32 + lane] = partial;
        quad_bar();
        const float s = xs[(xslot * 2) * 32 + lane] + xs[(xslot * 2 + 1) * 32 + lane];
        xslot ^= 1;
        return s;
    };
    uint32_t ldph = 0u;   // bit b = phase parity of buffer b's barrier
    auto rec_wait = [&](int b) { mbar_wait_u32(qbar(b), (ldph >> b) & 1u); ldph ^= 1u << b; };

    // partition scratch, timestep-major: [T][tiles][S X DQ : 3 x 4 units-of-8][128 seq] 16-byte pieces (bf16), then G [T][tiles][128]
    const size_t ntiles = (size_t)gridDim.x * NT;
    const uint32_t tile_gid = blockIdx.x * NT + tile;
    constexpr int kStepU4 = 12 * 128;
    uint4* sbase = reinterpret_cast<uint4*>(pl.scratch);
    float* G_ = reinterpret_cast<float*>(sbase + (size_t)T * ntiles * kStepU4) + (size_t)tile_gid * 128 + r;
    const size_t gstride = ntiles * 128;
    enum { AS = 0, AX = 1, ADQ = 2 };
    auto sb8 = [&](int t, int which, int c8) -> uint4* { return sbase + ((size_t)t * ntiles + tile_gid) * kStepU4 + (size_t)(which * 4 + part * 2 + c8) * 128 + r; };

    if (tid == 0) { for (int i = 0; i < NT * 12; ++i) mbar_init(ldbar + i, 1); fence_mbar_init(); }
    __syncthreads();

    uint64_t key = 0; uint32_t* ord = nullptr;
    uint64_t step = pl.step_ctr[live ? p : 0];
    if (live) { key = pl.keys[p]; ord = pl.order + (size_t)p * pl.n; }
    OptC o; o.lr = m.lr; o.l2 = m.l2; o.c1 = 1.0f; o.c2 = 1.0f;
    {   // expected concurrent visitors of an item row (see kernels_lstm_tile.cu)
        const float c = 3.0f * (float)pl.P / (float)m.N;
        o.chat = c > 1.0f ? c : 0.0f;
    }
    const int tries = m.loss == 2 ? 5 : 1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float* alpha = m.dense + part * DPT;
    const size_t nd = m.ndense;

    for (int ep = 0; ep < pl.epochs; ++ep) {
        if (live && lead) {  // thread_rng.shuffle(partition)  sequence_model.rs:109
            XorShift rng = pl.rng[p];
            uint32_t i = pl.n;
            while (i >= 2) {
                i -= 1;
                const uint32_t j = (uint32_t)xs_gen_below(rng, (uint64_t)i + 1);
                const uint32_t a = ord[i], b = ord[j];
                ord[i] = b; ord[j] = a;
            }
            pl.rng[p] = rng;
        }
        quad_bar();   // the other owner reads the shuffled order
        for (uint32_t it = 0; it < pl.n; ++it, ++step) {
            if (ADAM) {
                const float tt_ = (float)(pl.adam_t0 + step * pl.P + (live ? p : 0) + 1);
                o.c1 = 1.0f - powf(0.9f, tt_); o.c2 = 1.0f - powf(0.999f, tt_);
            }
            const uint32_t* ids = pl.item_ids;
            int Tn = 0;
            if (live) { const uint32_t sq = __ldcg(ord + it); ids = pl.item_ids + pl.seq_start[sq]; Tn = (int)pl.seq_len[sq] - 1; }
            // every lane of the two warps walks max(Tn) timesteps (the exchange barriers are warp-uniform)
            int Tmax = Tn;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) Tmax = max(Tmax, __shfl_xor_sync(kFull, Tmax, off));
            // a = sigmoid(alpha), read live at the start of the step (ewma.rs:302)
            float a[DPT];
#pragma unroll
            for (int cc = 0; cc < NCH; ++cc) {
                const float4 al = __ldcg(reinterpret_cast<const float4*>(alpha) + cc);
                a[4 * cc] = sigm(al.x); a[4 * cc + 1] = sigm(al.y); a[4 * cc + 2] = sigm(al.z); a[4 * cc + 3] = sigm(al.w);
            }
            auto idat = [&](int k) -> uint32_t { return (k <= Tn && Tn > 0) ? __ldg(ids + k) : 0u; };   // ids[k] or the dummy row 0

            // =========================== forward ===========================
            // Loads of timestep t are issued one timestep ahead: the target row {bias, w} -> P[t & 1] by the lead owner, the first
            // candidate's full record -> R[t & 1] by the second owner; the negative's reduce-add comes from the lead owner (an
            // even split of the per-lane issue loops: bulk copies take uniform-register operands).  x_0 goes through P[1].
            float s[DPT], x[DPT];
            float loss_seq = 0.0f;
            uint32_t idn = 0;   // ids[t + 2] of the coming iteration, loaded one iteration ahead (lead owner)
            if (Tmax > 0) {
                if (lead) {
                    bulk_load(smem_u32(pslot_g(1)), trec<FLAT>(tb, idat(0)), PSLOT, qbar(1));
                    bulk_load(smem_u32(pslot_g(0)), trec<FLAT>(tb, idat(1)), PSLOT, qbar(0));
                    idn = idat(2);
                } else {
                    bulk_load(smem_u32(rslot_g(0)), trec<FLAT>(tb, draw_item(key, step, 0u, 0u, pl.neg_range)), REC, qbar(0));
                }
                if (lead && lane == 0) { mbar_arrive_expect_tx(qbar(1), 32 * PSLOT); mbar_arrive_expect_tx(qbar(0), 32 * (PSLOT + REC)); }
                rec_wait(1);
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) {
                    const float4 v = *reinterpret_cast<const float4*>(pslot_g(1) + 16 + (part * NCH + cc) * 16);
                    x[4 * cc] = v.x; x[4 * cc + 1] = v.y; x[4 * cc + 2] = v.z; x[4 * cc + 3] = v.w;
                }
            }
            for (int t = 0; t < Tmax; ++t) {
                const bool act = t < Tn;
                const int b = t & 1;
                if (lead) bulk_wait_read();   // the previous step's reduce-add has read R[b ^ 1]: it may be loaded again (below, by the other owner)
                // s_t
#pragma unroll
                for (int d = 0; d < DPT; d += 2) {
                    const float2 xv = make_float2(x[d], x[d + 1]);
                    const float2 sv = t == 0 ? xv : fma2(make_float2(a[d], a[d + 1]), add2(make_float2(s[d], s[d + 1]), mul2(xv, splat2(-1.0f))), xv);
                    s[d] = act ? sv.x : 0.0f; s[d + 1] = act ? sv.y : 0.0f;
                }
                rec_wait(b);
                float4 pv[NCH], qv[NCH];
                float pos;
                {
                    float acc = lead ? *reinterpret_cast<const float*>(pslot_g(b)) : 0.0f;   // b[out]
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        pv[cc] = *reinterpret_cast<const float4*>(pslot_g(b) + 16 + (part * NCH + cc) * 16);
                        acc = fmaf(s[4 * cc], pv[cc].x, acc); acc = fmaf(s[4 * cc + 1], pv[cc].y, acc);
                        acc = fmaf(s[4 * cc + 2], pv[cc].z, acc); acc = fmaf(s[4 * cc + 3], pv[cc].w, acc);
                    }
                    pos = xsum(acc);
                }
                // the lead owner has passed its wait above before this exchange: the other buffers are free
                if (t + 1 < Tmax) {
                    if (lead) {
                        bulk_load(smem_u32(pslot_g(b ^ 1)), trec<FLAT>(tb, idn), PSLOT, qbar(b ^ 1));
                        if (lane == 0) mbar_arrive_expect_tx(qbar(b ^ 1), 32 * (PSLOT + REC));
                        idn = idat(t + 3);
                    } else bulk_load(smem_u32(rslot_g(b ^ 1)), trec<FLAT>(tb, draw_item(key, step, (uint32_t)(t + 1), 0u, pl.neg_range)), REC, qbar(b ^ 1));
                }
                bool done = !act; uint32_t neg = draw_item(key, step, (uint32_t)t, 0u, pl.neg_range); float ngs = 0.0f;
                auto score = [&]() {
                    float acc = lead ? *reinterpret_cast<const float*>(rslot_g(b)) : 0.0f;   // b[candidate]
                    float4 qt[NCH];
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        qt[cc] = *reinterpret_cast<const float4*>(rslot_g(b) + 16 + (part * NCH + cc) * 16);
                        acc = fmaf(s[4 * cc], qt[cc].x, acc); acc = fmaf(s[4 * cc + 1], qt[cc].y, acc);
                        acc = fmaf(s[4 * cc + 2], qt[cc].z, acc); acc = fmaf(s[4 * cc + 3], qt[cc].w, acc);
                    }
                    const float tot = xsum(acc);
                    if (!done) {
                        ngs = tot;
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) qv[cc] = qt[cc];
                        if (1.0f - pos + tot > 0.0f) done = true;
                    }
                };
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) qv[cc] = zero4;
                score();
                for (int j = 1; j < tries; ++j) {   // WARP: further candidates (sequence_model.rs:58-65); decisions are identical in both owners
                    if (__all_sync(kFull, done)) break;
                    const uint32_t cj = draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range);
                    if (!done) { neg = cj; if (!lead) bulk_load(smem_u32(rslot_g(b)), trec<FLAT>(tb, cj), REC, qbar(b)); }
                    const uint32_t n = __popc(__ballot_sync(kFull, !done));
                    if (lead && lane == 0) mbar_arrive_expect_tx(qbar(b), n * REC);
                    rec_wait(b);
                    score();
                }
                float g = 0.0f;
                if (act) {
                    float l;
                    if (m.loss == 0) { const float sg = sigm(ngs - pos); l = sg; g = sg * (1.0f - sg); }
                    else { const float v = 1.0f + ngs - pos; l = v > 0.0f ? v : 0.0f; g = v > 0.0f ? 1.0f : 0.0f; }
                    loss_seq += l;
                    if (lead) G_[(size_t)t * gstride] = g;
                    // activation copies for the backward pass (bf16): s_t, x_t, g (q - p)
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        *sb8(t, AS, c8) = make_uint4(pack2(s[8 * c8], s[8 * c8 + 1]), pack2(s[8 * c8 + 2], s[8 * c8 + 3]), pack2(s[8 * c8 + 4], s[8 * c8 + 5]), pack2(s[8 * c8 + 6], s[8 * c8 + 7]));
                        *sb8(t, AX, c8) = make_uint4(pack2(x[8 * c8], x[8 * c8 + 1]), pack2(x[8 * c8 + 2], x[8 * c8 + 3]), pack2(x[8 * c8 + 4], x[8 * c8 + 5]), pack2(x[8 * c8 + 6], x[8 * c8 + 7]));
                        const float4 qa = qv[2 * c8], qb = qv[2 * c8 + 1], pa = pv[2 * c8], pb = pv[2 * c8 + 1];
                        *sb8(t, ADQ, c8) = make_uint4(pack2(g * (qa.x - pa.x), g * (qa.y - pa.y)), pack2(g * (qa.z - pa.z), g * (qa.w - pa.w)),
                                                       pack2(g * (qb.x - pb.x), g * (qb.y - pb.y)), pack2(g * (qb.z - pb.z), g * (qb.w - pb.w)));
                    }
                    // the negative's visit, in its slot: E[neg_t] += step(+g s_t), b[neg_t] += step(+g)
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        float* sl = reinterpret_cast<float*>(rslot_g(b) + 16 + (part * NCH + cc) * 16);
                        float4 w = qv[cc], s1 = *reinterpret_cast<const float4*>(sl + 32), s2 = zero4, dw = zero4, d1 = zero4, d2 = zero4;
                        if (ADAM) s2 = *reinterpret_cast<const float4*>(sl + 64);
                        apply_chunk<ADAM>(w, s1, s2, make_float4(g * s[4 * cc], g * s[4 * cc + 1], g * s[4 * cc + 2], g * s[4 * cc + 3]), o, dw, d1, d2);
                        *reinterpret_cast<float4*>(sl) = dw; *reinterpret_cast<float4*>(sl + 32) = d1;
                        if (ADAM) *reinterpret_cast<float4*>(sl + 64) = d2;
                    }
                    if (lead) {
                        float4 bq = *reinterpret_cast<const float4*>(rslot_g(b));
                        const float4 b0 = bq;
                        if (!ADAM) adagrad1(bq.x, bq.y, g, o); else adam1(bq.x, bq.y, bq.z, g, o);
                        *reinterpret_cast<float4*>(rslot_g(b)) = ADAM ? bq : make_float4(bq.x - b0.x, bq.y - b0.y, 0.0f, 0.0f);
                    }
                }
                fence_async_smem();
                quad_bar();              // both owners have written their halves / are done reading the slots of buffer b
                if (lead && act) {
                    if (ADAM) bulk_store(trec<FLAT>(tb, neg), smem_u32(rslot_g(b)), REC); else bulk_reduce_add(trec<FLAT>(tb, neg), smem_u32(rslot_g(b)), REC);
                    bulk_commit();
                }
                // x_{t+1} = E[out_t]: the copy fetched for the score
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) { x[4 * cc] = pv[cc].x; x[4 * cc + 1] = pv[cc].y; x[4 * cc + 2] = pv[cc].z; x[4 * cc + 3] = pv[cc].w; }
            }
            if (lead) bulk_wait_read();
            quad_bar();                  // every slot is free

            // =========================== backward ===========================
            // visit of backward timestep t: row ids[t+1] (entries dx_{t+1}, then -g_t s_t) from record buffer (t + 3) % 3.  The
            // second owner requests the records TWO timesteps ahead (three rotating buffers), the lead owner issues the
            // reduce-adds -- again an even split of the issue loops.  t = -1: only E[ids[0]] += step(dx_0).
            float ds[DPT], da[DPT], dxp[DPT];
#pragma unroll
            for (int d = 0; d < DPT; ++d) { ds[d] = 0.0f; da[d] = 0.0f; dxp[d] = 0.0f; }
            auto kb = [&](int t) -> int { return (t + 3) % 3; };
            auto pf = [&](const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); };
            if (Tmax > 0 && !lead) {
                const int t0 = Tmax - 1;
                bulk_load(smem_u32(rslot3_g(kb(t0))), trec<FLAT>(tb, idat(t0 + 1)), REC, qbar(kb(t0)));
                bulk_load(smem_u32(rslot3_g(kb(t0 - 1))), trec<FLAT>(tb, idat(t0)), REC, qbar(kb(t0 - 1)));
                if (lane == 0) { mbar_arrive_expect_tx(qbar(kb(t0)), 32 * REC); mbar_arrive_expect_tx(qbar(kb(t0 - 1)), 32 * REC); }
            }
            for (int t = Tmax - 1; t >= (Tmax > 0 ? -1 : 0); --t) {
                const bool act = t >= 0 && t < Tn;
                const bool has_dx = t + 1 < Tn;
                const int b = kb(t);
                // ids for the end of this iteration: the lead's reduce-add goes to ids[t + 1], the other owner's load (iteration
                // t - 2) to ids[t - 1]
                const uint32_t id_end = lead ? idat(t + 1) : (t >= 1 ? idat(t - 1) : 0u);
                if ((lane & 7) == 0 && t >= 1 && t - 1 < Tn) {   // next timestep's activation copies towards L2 (one 128-byte line per 8 lanes)
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) { pf(sb8(t - 1, AX, c8)); pf(sb8(t - 1, ADQ, c8)); if (t >= 2) pf(sb8(t - 2, AS, c8)); }
                    if (lead) pf(G_ + (size_t)(t - 1) * gstride);
                }
                float g = 0.0f;
                float4 ghn[NCH];         // -g_t s_t
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) ghn[cc] = zero4;
                float dx[DPT];
#pragma unroll
                for (int d = 0; d < DPT; ++d) dx[d] = 0.0f;
                if (act) {
                    g = __ldcg(G_ + (size_t)t * gstride);
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        const uint4 us = __ldcg(sb8(t, AS, c8)), uq = __ldcg(sb8(t, ADQ, c8));
                        uint4 ux = make_uint4(0u, 0u, 0u, 0u), up = ux;
                        if (t > 0) { ux = __ldcg(sb8(t, AX, c8)); up = __ldcg(sb8(t - 1, AS, c8)); }
#pragma unroll
                        for (int pr = 0; pr < 4; ++pr) {
                            const int d = 8 * c8 + 2 * pr;
                            const float2 sv = bf2(word_of(us, pr)), dq = bf2(word_of(uq, pr));
                            const float2 dh = add2(make_float2(ds[d], ds[d + 1]), dq);
                            const float2 gn = mul2(splat2(-g), sv);
                            if (pr < 2) { if (pr == 0) { ghn[2 * c8].x = gn.x; ghn[2 * c8].y = gn.y; } else { ghn[2 * c8].z = gn.x; ghn[2 * c8].w = gn.y; } }
                            else { if (pr == 2) { ghn[2 * c8 + 1].x = gn.x; ghn[2 * c8 + 1].y = gn.y; } else { ghn[2 * c8 + 1].z = gn.x; ghn[2 * c8 + 1].w = gn.y; } }
                            if (t == 0) { dx[d] = dh.x; dx[d + 1] = dh.y; ds[d] = 0.0f; ds[d + 1] = 0.0f; }
                            else {
                                const float2 av = make_float2(a[d], a[d + 1]);
                                const float2 adh = mul2(av, dh);
                                const float2 dxv = add2(dh, mul2(adh, splat2(-1.0f)));            // (1 - a) dh
                                const float2 dav = fma2(dh, add2(bf2(word_of(up, pr)), mul2(bf2(word_of(ux, pr)), splat2(-1.0f))), make_float2(da[d], da[d + 1]));
                                dx[d] = dxv.x; dx[d + 1] = dxv.y; ds[d] = adh.x; ds[d + 1] = adh.y; da[d] = dav.x; da[d + 1] = dav.y;
                            }
                        }
                    }
                }
                rec_wait(b);
                const bool visit = act || has_dx;
                uint8_t* const rsl = rslot3_g(b);
                if (visit) {
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        float* sl = reinterpret_cast<float*>(rsl + 16 + (part * NCH + cc) * 16);
                        float4 w = *reinterpret_cast<const float4*>(sl), s1 = *reinterpret_cast<const float4*>(sl + 32), s2 = zero4, dw = zero4, d1 = zero4, d2 = zero4;
                        if (ADAM) { s2 = *reinterpret_cast<const float4*>(sl + 64); dw = w; d1 = s1; d2 = s2; }
                        if (has_dx) apply_chunk<ADAM>(w, s1, s2, make_float4(dxp[4 * cc], dxp[4 * cc + 1], dxp[4 * cc + 2], dxp[4 * cc + 3]), o, dw, d1, d2);
                        if (act) apply_chunk<ADAM>(w, s1, s2, ghn[cc], o, dw, d1, d2);
                        *reinterpret_cast<float4*>(sl) = dw; *reinterpret_cast<float4*>(sl + 32) = d1;
                        if (ADAM) *reinterpret_cast<float4*>(sl + 64) = d2;
                    }
                    if (lead) {
                        float4 bq = *reinterpret_cast<const float4*>(rsl);
                        const float4 b0 = bq;
                        if (act) { if (!ADAM) adagrad1(bq.x, bq.y, -g, o); else adam1(bq.x, bq.y, bq.z, -g, o); }
                        *reinterpret_cast<float4*>(rsl) = ADAM ? bq : make_float4(bq.x - b0.x, bq.y - b0.y, 0.0f, 0.0f);
                    }
                }
                fence_async_smem();
                if (lead) bulk_wait_read();   // the previous iteration's reduce-add has read its buffer: the other owner may refill it
                quad_bar();
                if (lead) {
                    if (visit) {
                        if (ADAM) bulk_store(trec<FLAT>(tb, id_end), smem_u32(rsl), REC); else bulk_reduce_add(trec<FLAT>(tb, id_end), smem_u32(rsl), REC);
                        bulk_commit();
                    }
                } else if (t >= 1) {          // record of iteration t - 2 (row ids[t - 1]) into the buffer iteration t + 1 used
                    bulk_load(smem_u32(rslot3_g(kb(t - 2))), trec<FLAT>(tb, id_end), REC, qbar(kb(t - 2)));
                    if (lane == 0) mbar_arrive_expect_tx(qbar(kb(t - 2)), 32 * REC);
                }
#pragma unroll
                for (int d = 0; d < DPT; ++d) dxp[d] = dx[d];
            }
            if (lead) bulk_wait_read();
            quad_bar();                  // every slot is free for the next sub-sequence's forward loads
            if (live && lead) { pl.loss_acc[p] += loss_seq; pl.examples[p] += (unsigned long long)Tn; }

            // =========================== dense step: alpha (ewma.rs:302; one optimizer step per sub-sequence) ===========================
            if (live && Tn > 0) {
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) {
                    float4 w = __ldcg(reinterpret_cast<const float4*>(alpha) + cc), s1 = __ldcg(reinterpret_cast<const float4*>(alpha + nd) + cc), s2 = zero4;
                    if (ADAM) s2 = __ldcg(reinterpret_cast<const float4*>(alpha + 2 * nd) + cc);
                    float4 dw = zero4, d1 = zero4, d2 = zero4;
                    const float4 gal = make_float4(da[4 * cc] * a[4 * cc] * (1.0f - a[4 * cc]), da[4 * cc + 1] * a[4 * cc + 1] * (1.0f - a[4 * cc + 1]),
                                                   da[4 * cc + 2] * a[4 * cc + 2] * (1.0f - a[4 * cc + 2]), da[4 * cc + 3] * a[4 * cc + 3] * (1.0f - a[4 * cc + 3]));
                    apply_chunk<ADAM>(w, s1, s2, gal, o, dw, d1, d2);
                    if (!ADAM) { red_add4(alpha + 4 * cc, dw); red_add4(alpha + nd + 4 * cc, d1); }
                    else {   // plain read-modify-write (see apply_chunk): Hogwild as in the reference
                        __stcg(reinterpret_cast<float4*>(alpha) + cc, dw); __stcg(reinterpret_cast<float4*>(alpha + nd) + cc, d1);
                        __stcg(reinterpret_cast<float4*>(alpha + 2 * nd) + cc, d2);
                    }
                }
            }
        }
    }
    if (live && lead) pl.step_ctr[p] = step;
}

template <int NT, int S, bool FLAT>
cudaError_t launch_one(const ModelDev& m, const PlanDev& p, cudaStream_t st) {
    const size_t rec = 16 + S * 128;
    const size_t smem = OFF_TILES + (size_t)NT * (2 * 128 * PSLOT + (rec <= 2 * PSLOT ? 2 : 3) * 128 * rec) + (size_t)NT * 4 * XS_BYTES_PER_QUAD;
    const int seq_per_cta = 128 * NT;
    dim3 grid((p.P + seq_per_cta - 1) / seq_per_cta);
    cudaError_t e = cudaFuncSetAttribute(ewma_tile_train_kernel<NT, S, FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ewma_tile_train_kernel<NT, S, FLAT><<<grid, seq_per_cta * 2, smem, st>>>(m, p);
    return cudaGetLastError();
}

}  // namespace

// Adagrad: two groups of 128 partitions per CTA; Adam (400-byte records): one
int ewma_tile_tiles_per_cta(const ModelDev& m, uint32_t P) { return (m.opt == 1 || P % 256 != 0) ? 1 : 2; }

cudaError_t launch_ewma_tile(const ModelDev& m, const PlanDev& p, cudaStream_t st) {
    const bool flat = m.gmask == 0;
    const int nt = ewma_tile_tiles_per_cta(m, p.P);
    if (m.opt == 1) return flat ? launch_one<1, 3, true>(m, p, st) : launch_one<1, 3, false>(m, p, st);
    if (nt == 2) return flat ? launch_one<2, 2, true>(m, p, st) : launch_one<2, 2, false>(m, p, st);
    return flat ? launch_one<1, 2, true>(m, p, st) : launch_one<1, 2, false>(m, p, st);
}

}  // namespace sbr
