// kernels_lstm_tile.cu -- tensor-core LSTM training kernel (D = 32): the throughput path of fit() for LSTM models.
//
// Replaces, for many concurrent Hogwild partitions, the rayon body of fit_sequence_model (sequence_model.rs:100-175) and
// the wyrm graph it drives (lstm.rs:258-337): gather -> LSTM forward -> WARP / uniform negative + scores -> loss ->
// backward through time -> sparse optimizer visits -> dense step.  One tile = 128 partitions advancing in lock-step, two
// tiles per CTA, one CTA per SM.  A sequence is owned by DS = 2 threads in two warps of the same TMEM lane quarter
// (TMEM lane == sequence): owner `part` holds hidden units [16 part, 16 part + 16) of everything.
//
// What moves the bytes (this is what the kernel is built around):
//   * item records travel by TMA bulk copies, one instruction per record: `cp.async.bulk.shared.global` brings the
//     272-byte record {bias quad | w[32] | G[32]} of an item into the sequence's own shared-memory slot (mbarrier
//     completion), and `cp.reduce.async.bulk.global.shared.add.f32` adds the optimizer step {db, dGb | dw[32] | dG[32]}
//     back into the live table at L2 -- no per-16-byte loads, atomics, shuffles or pointer arithmetic in the SM, no lost
//     updates between the thousands of partitions that hit the same hot rows (Adagrad's accumulator is additive).  Adam
//     records go back with a plain bulk STORE of the new {b, m, v | w | m | v}: its moments are decaying averages, and
//     m += 0.1 (g - m_seen) summed over K concurrent visitors of a hot row diverges for K > 20 -- Hogwild overwrites as in
//     the reference cannot.  Slot strides (144 / 272 / 400 bytes) are = 16 mod 128: lane-per-row 16-byte
//     accesses are bank-conflict free.
//   * the three contractions of a timestep run on the tensor cores from shared-memory operand tiles with TMEM
//     accumulators (tc_tile.cuh: tcgen05.mma kind::f16 on bf16 operands, fp32 accumulation):
//       gates = [h_{t-1}, x_t, 1] . [W ; b]     128 x 128 x 80   A = Z tile K-major, B = W tile MN-major (bias = row 64)
//       dz    = delta_t . W^T                    128 x  64 x 128  A = delta tile K-major, B = the same W tile K-major
//       dW^T += delta_t^T . [h_{t-1}, x_t, 1]    128 x  80 x 128  A = delta tile MN-major, B = Z tile MN-major
//     (no-swizzle core-matrix layout: the same bytes serve as K-major along one axis and MN-major along the other).
//
// Order of the sparse visits of one sub-sequence (the reference records (row, gradient) entries during backward and
// applies them un-merged; wyrm's order inside a step is not known):
//   forward,  t ascending : E[neg_t] += step(+g_t h_t), b[neg_t] += step(+g_t)      -- as soon as the loss of timestep t
//                           is known; the record is already in shared memory for the score
//   backward, t descending: E[ids[t+1]] += step(dx_{t+1}) then step(-g_t h_t), b[ids[t+1]] += step(-g_t)
//                           -- in_{t+1} and out_t are always the same row: one record load, two sequential applications,
//                           one reduce-add
// x_{t+1} is the copy of E[out_t] that was fetched for the score of timestep t (fetched before any update of this
// sub-sequence touches the row, like the reference reads it).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "engine.h"
#include "tc_tile.cuh"

namespace sbr {

namespace {

using namespace tc;

constexpr int kD = 32, kNG = 128, kKP = 80;      // K of the forward / dW products: h (32), x (32), 1, 15 zeros
constexpr uint32_t OFF_WB = 0;                   // bf16 [80 feat][128 gd]
constexpr uint32_t OFF_MISC = 20480;             // mbarriers, tmem base, tile maxima, shard pointers
constexpr uint32_t OFF_TILES = 21504;
constexpr uint32_t TILE_DB = 0;                  // bf16 [128 seq][128 gd]   (forward: slot P, 128 x 144 bytes)
constexpr uint32_t TILE_ZB = 32768;              // bf16 [128 seq][80 feat]
constexpr uint32_t TILE_ST = 53248;              // record slots: 128 x REC bytes (forward: candidate / negative, backward: chain row)
constexpr uint32_t PSLOT = 144;                  // bias quad + w
constexpr uint32_t XS_BYTES_PER_QUAD = 2 * 2 * 32 * 4;   // score exchange: 2 slots x 2 parts x 32 lanes floats

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / (1 + 2^(-x log2 e)); saturates cleanly: ex2 -> +inf gives rcp -> 0, ex2 -> 0 gives 1
__device__ __forceinline__ float sigm(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tnh(float x) { return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f); }

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of a vector) ----
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
__device__ __forceinline__ float2 sigm2(const float2& x) {
    const float2 t = mul2(x, splat2(-1.4426950408889634f));
    const float2 e = add2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), splat2(1.0f));
    return make_float2(rcp_approx(e.x), rcp_approx(e.y));
}
__device__ __forceinline__ float2 tnh2(const float2& x) {
    const float2 t = mul2(x, splat2(-2.8853900817779268f));
    const float2 e = add2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), splat2(1.0f));
    return fma2(make_float2(rcp_approx(e.x), rcp_approx(e.y)), splat2(2.0f), splat2(-1.0f));
}
__device__ __forceinline__ float2 bf2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

__device__ __forceinline__ uint32_t word_of(const uint4& u, int pr) { return pr == 0 ? u.x : pr == 1 ? u.y : pr == 2 ? u.z : u.w; }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { uint32_t o; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(hi), "f"(lo)); return o; }
__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) { return make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w)); }

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA bulk copies of whole item records ----
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

// one element of a sparse Adagrad / Adam visit on register copies; returns the step in w (the additive deltas of the
// state are what the caller accumulates: G' - G resp. m' - m, v' - v)
struct OptC { float lr, l2, c1, c2; int adam; float chat; };
__device__ __forceinline__ void adagrad1(float& w, float& G, float g, const OptC& o) {
    g = fmaf(w, o.l2, g);
    G = fmaf(g, g, G);
    w = fmaf(-o.lr * g, rsqrt_approx(fmaxf(G, fmaxf(o.chat * g * g, 1e-20f))), w);   // lr / (1e-10 + sqrt(G)) * g
}
__device__ __forceinline__ void adam1(float& w, float& m, float& v, float g, const OptC& o) {
    g = fmaf(w, o.l2, g);
    m = 0.9f * m + 0.1f * g;
    v = 0.999f * v + 0.001f * g * g;
    const float mhat = __fdividef(m, o.c1), vhat = __fdividef(v, o.c2);
    w -= __fdividef(o.lr * mhat, sqrtf(vhat) + 1e-8f);
}
__device__ __forceinline__ void apply4(float4& w, float4& s, float4& v, const float4& g, float sign, const OptC& o) {
    if (!o.adam) {
        adagrad1(w.x, s.x, sign * g.x, o); adagrad1(w.y, s.y, sign * g.y, o); adagrad1(w.z, s.z, sign * g.z, o); adagrad1(w.w, s.w, sign * g.w, o);
    } else {
        adam1(w.x, s.x, v.x, sign * g.x, o); adam1(w.y, s.y, v.y, sign * g.y, o); adam1(w.z, s.z, v.z, sign * g.z, o); adam1(w.w, s.w, v.w, sign * g.w, o);
    }
}
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
__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

struct Table { float* e0; float* const* es; uint32_t stride, gmask; int gshift; };
// base of the record of item `id` (the bias quad; w follows at + 4 floats)
template <bool FLAT>
__device__ __forceinline__ float* trec(const Table& tb, uint32_t id) {
    if (FLAT) return tb.e0 + (size_t)id * tb.stride;
    return tb.es[id & tb.gmask] + (size_t)(id >> tb.gshift) * tb.stride;
}

// one 8-unit block of a timestep's saved activations (backward operand), bf16 x 8 each
struct ActB { uint4 f, i, g, o, q, cp, tc, h; };

template <int NT, int S, bool FLAT>
__global__ void __launch_bounds__(256 * NT, 1) lstm_tile_train_kernel(ModelDev m, PlanDev pl) {
    constexpr int DS = 2, DPT = 16, NCH = 4, NB8 = 2, TT = 256;
    constexpr uint32_t REC = 16 + S * 128;                 // bytes of an item record / record slot
    constexpr uint32_t TILE_BYTES = TILE_ST + 128 * REC;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Wb = smem + OFF_WB;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_MISC);           // [NT] MMA completion
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 32);
    int* tmax_s = reinterpret_cast<int*>(smem + OFF_MISC + 40);              // [NT]
    float** es_s = reinterpret_cast<float**>(smem + OFF_MISC + 64);          // [8] shard base pointers
    uint64_t* ldbar = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 128);    // [NT][4] record loads, one per lane quarter

    const int tid = threadIdx.x, tile = tid / TT, tt = tid % TT;
    const int wi = tt >> 5, q = wi & 3, part = wi >> 2, lane = tid & 31;
    const int r = q * 32 + lane;                     // sequence row of the tile == TMEM lane
    const int gb0 = part * NB8;                      // first 8-unit block owned by this thread
    uint8_t* Db = smem + OFF_TILES + tile * TILE_BYTES + TILE_DB;
    uint8_t* Zb = smem + OFF_TILES + tile * TILE_BYTES + TILE_ZB;
    uint8_t* St = smem + OFF_TILES + tile * TILE_BYTES + TILE_ST;
    float* xs = reinterpret_cast<float*>(smem + OFF_TILES + NT * TILE_BYTES + (tile * 4 + q) * XS_BYTES_PER_QUAD);
    const uint32_t pslot = smem_u32(Db) + (uint32_t)r * PSLOT;    // forward: {bias quad | w} of the target row
    const uint32_t rslot = smem_u32(St) + (uint32_t)r * REC;      // the full record of the candidate / chain row
    const uint8_t* pslot_g = Db + (size_t)r * PSLOT;
    uint8_t* rslot_g = St + (size_t)r * REC;
    const uint32_t qbar = smem_u32(ldbar + tile * 4 + q);
    const uint32_t tile_gid = blockIdx.x * NT + tile;
    const uint32_t p = tile_gid * 128u + r;
    const bool live = p < pl.P;
    const bool lead = part == 0;                     // the owner that issues the record traffic and does per-sequence scalar work
    const size_t nd = m.ndense;
    const bool coupled = m.variant == 1;
    const int T = m.T;
    if (tid < 8) es_s[tid] = m.Es[tid];
    Table tb; tb.e0 = m.Es[0]; tb.es = es_s; tb.stride = (uint32_t)rec_floats(m); tb.gmask = m.gmask; tb.gshift = m.gshift;

    auto tile_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(tile + 1), "n"(TT) : "memory"); };
    auto quad_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(3 + tile * 4 + q), "n"(64) : "memory"); };
    int xslot = 0;
    // sum of the two owners' partial values, identical (same order) in both; warp-uniform call sites only
    auto xsum = [&](float partial) -> float {
        xs[(xslot * 2 + part) * 32 + lane] = partial;
        quad_bar();
        const float s = xs[(xslot * 2) * 32 + lane] + xs[(xslot * 2 + 1) * 32 + lane];
        xslot ^= 1;
        return s;
    };
    uint32_t ldph = 0;
    auto rec_wait = [&]() {
        asm volatile(
            "{\n\t.reg .pred p;\n\tLW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra LD_%=;\n\tbra LW_%=;\n\tLD_%=:\n\t}\n" ::"r"(qbar),
            "r"(ldph)
            : "memory");
        ldph ^= 1;
    };

    // tile scratch, timestep-major across the grid's tiles: [T][tiles][9 x 4 units][128 seq] 16-byte pieces (bf16 F I G O X DQ C
    // TANH(C) H: 72 KB per tile-timestep), then G [T][tiles][128]
    const size_t ntiles = (size_t)gridDim.x * NT;
    constexpr int kStepU4 = 36 * 128;
    uint4* sbase = reinterpret_cast<uint4*>(pl.scratch);
    float* G_ = reinterpret_cast<float*>(sbase + (size_t)T * ntiles * kStepU4) + (size_t)tile_gid * 128 + r;   // + t * gstride
    const size_t gstride = ntiles * 128;
    enum { AF = 0, AI = 1, AG = 2, AO = 3, AX = 4, ADQ = 5, AC = 6, ATC = 7, AHB = 8 };
    auto step_base = [&](int t) -> uint4* { return sbase + ((size_t)t * ntiles + tile_gid) * kStepU4; };
    auto sb8 = [&](int t, int which, int c8) -> uint4* { return step_base(t) + (size_t)(which * 4 + c8) * 128 + r; };
    auto prefetch_step = [&](int t) {   // 72 KB = 576 lines towards L2, spread over the tile's threads
        constexpr int LPT = (576 + TT - 1) / TT;
        const char* base = reinterpret_cast<const char*>(step_base(t));
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            const int line = tt * LPT + i;
            if (line < 576) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)line * 128));
        }
    };
    auto load_act = [&](ActB& a, int t, int gb, bool on) {
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        a.f = a.i = a.g = a.o = a.q = a.cp = a.tc = a.h = z4;
        if (on) {
            a.f = __ldcg(sb8(t, AF, gb)); a.i = __ldcg(sb8(t, AI, gb)); a.g = __ldcg(sb8(t, AG, gb)); a.o = __ldcg(sb8(t, AO, gb));
            a.q = __ldcg(sb8(t, ADQ, gb)); a.tc = __ldcg(sb8(t, ATC, gb)); a.h = __ldcg(sb8(t, AHB, gb));
            if (t > 0) a.cp = __ldcg(sb8(t - 1, AC, gb));
        }
    };
    // Z_t = [h_{t-1}, x_t] (bf16) straight from the scratch into this thread's row of the Z tile; rows of finished
    // sequences are zero-filled (src-size 0): their deltas are 0, but 0 x stale bits must not become NaN in dW
    auto stage_z_async = [&](int t, bool on) {
        const uint32_t nb = on ? 16u : 0u;
#pragma unroll
        for (int b = 0; b < NB8; ++b) {
            const int gb = gb0 + b;
            cp_async16_zfill(smem_u32(Zb + tile_chunk_off(r, 4 + gb, 10)), sb8(t, AX, gb), nb);
            if (t > 0) cp_async16_zfill(smem_u32(Zb + tile_chunk_off(r, gb, 10)), sb8(t - 1, AHB, gb), nb);
            else *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, gb, 10)) = make_uint4(0u, 0u, 0u, 0u);
        }
        cp_commit();
    };
    // this thread's 16-byte chunk cc (4 units) of vector `vec` (0 = w, 1 = s1, 2 = s2) of the record in its slot
    auto rs_ld = [&](int vec, int cc) -> float4 { return *reinterpret_cast<const float4*>(rslot_g + 16 + vec * 128 + (part * NCH + cc) * 16); };
    auto rs_st = [&](int vec, int cc, const float4& v) { *reinterpret_cast<float4*>(rslot_g + 16 + vec * 128 + (part * NCH + cc) * 16) = v; };

    // ---- one-time setup ----
    if (tid < 32) tmem_alloc<(NT == 1 ? 256 : 512)>(tmem_ptr);
    if (tid == 0) {
        for (int i = 0; i < NT; ++i) mbar_init(mbar + i, 1);
        for (int i = 0; i < NT * 4; ++i) mbar_init(ldbar + i, 1);
        fence_mbar_init();
    }
    if (lead) {   // constant columns of the Z tile: col 64 = 1 (bias input / bias gradient), 65..79 = 0
        const float one8[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 8, 10)) = pack_bf16x8(one8);
        *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 9, 10)) = pack_bf16x8(zero8);
    }
    if (tile == 0) {  // weights: thread (gd, part) stages its share of column gd of [W ; b ; 0]
        const int gd = r;
        __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(Wb);
        for (int k = part * 40; k < (part + 1) * 40; ++k) {
            const float v = k <= 64 ? __ldcg(m.dense + (size_t)k * kNG + gd) : 0.0f;   // k == 64: bias[gd]
            wb[(tile_chunk_off(k, gd >> 3, 16) >> 1) + (gd & 7)] = __float2bfloat16_rn(v);
        }
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_ptr + (uint32_t)tile * 256u + ((uint32_t)(q * 32) << 16);   // this thread's TMEM lane, tile's columns
    const uint32_t tcol0 = *tmem_ptr + (uint32_t)tile * 256u;                                 // for the MMA issuer
    const uint32_t tcs = tbase + 208u + (uint32_t)(part * DPT);                               // this thread's cell-state columns
    const uint32_t db_a = smem_u32(Db), zb_a = smem_u32(Zb), wb_a = smem_u32(Wb);
    constexpr uint32_t IDESC_F = make_idesc_bf16(128, 128, 0, 1);
    constexpr uint32_t IDESC_G2 = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t IDESC_G3 = make_idesc_bf16(128, 80, 1, 1);
    const bool issuer = tt == 0;
    uint32_t phase = 0;

    uint64_t key = 0; uint32_t* ord = nullptr;
    uint64_t step = pl.step_ctr[live ? p : 0];
    if (live) { key = pl.keys[p]; ord = pl.order + (size_t)p * pl.n; }
    constexpr bool ADAM = S == 3;   // 400-byte records <=> Adam
    OptC o; o.lr = m.lr; o.l2 = m.l2; o.adam = ADAM ? 1 : 0; o.c1 = 1.0f; o.c2 = 1.0f;
    {   // expected concurrent visitors of an item row: 3 visits per partition-timestep over m.N rows (DESIGN 3.4 "cold start")
        const float c = 3.0f * (float)pl.P / (float)m.N;
        o.chat = c > 1.0f ? c : 0.0f;
    }
    const int tries = m.loss == 2 ? 5 : 1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

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
        quad_bar();   // the other owner reads the shuffled order (same SM: visible after the barrier)
        for (uint32_t it = 0; it < pl.n; ++it, ++step) {
            if (ADAM) {
                const float tt_ = (float)(pl.adam_t0 + step * pl.P + (live ? p : 0) + 1);
                o.c1 = 1.0f - powf(0.9f, tt_); o.c2 = 1.0f - powf(0.999f, tt_);
            }
            const uint32_t* ids = pl.item_ids;
            int Tn = 0;
            if (live) { const uint32_t sq = __ldcg(ord + it); ids = pl.item_ids + pl.seq_start[sq]; Tn = (int)pl.seq_len[sq] - 1; }
            if (tt == 0) tmax_s[tile] = 0;   // tile-wide number of lock-step timesteps
            tile_bar();
            if (lead) atomicMax(&tmax_s[tile], Tn);
            tile_bar();
            const int Tmax = tmax_s[tile];

            // =========================== forward ===========================
            float h[DPT];
#pragma unroll
            for (int d = 0; d < DPT; ++d) h[d] = 0.0f;
            {
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int b = 0; b < NB8; ++b) tmem_st8(tcs + b * 8, z8);   // c_{-1} = 0
                tmem_st_wait();
            }
            float loss_seq = 0.0f;
            uint32_t idB = 0, idC = 0;   // ids[t+1], ids[t+2] travel in registers of the lead owner (ids has Tn + 1 entries)
            if (Tmax > 0) {
                // x_0 = E[ids[0]]: one record copy {bias quad | w} into the P slot
                uint32_t idA = 0;
                if (Tn > 0) { idA = __ldg(ids); idB = __ldg(ids + 1); }
                if (Tn > 1) idC = __ldg(ids + 2);
                if (!lead) {
                    bulk_load(pslot, trec<FLAT>(tb, idA), PSLOT, qbar);
                    if (lane == 0) mbar_arrive_expect_tx(qbar, 32 * PSLOT);
                }
                rec_wait();
                const bool a0 = Tn > 0;
#pragma unroll
                for (int b = 0; b < NB8; ++b) {
                    const float4 xa = *reinterpret_cast<const float4*>(pslot_g + 16 + (part * NCH + 2 * b) * 16);
                    const float4 xb = *reinterpret_cast<const float4*>(pslot_g + 16 + (part * NCH + 2 * b + 1) * 16);
                    const uint4 xp = pack8(xa, xb);
                    *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 4 + gb0 + b, 10)) = xp;
                    *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, gb0 + b, 10)) = make_uint4(0u, 0u, 0u, 0u);
                    if (a0) *sb8(0, AX, gb0 + b) = xp;
                }
            }
            for (int t = 0; t < Tmax; ++t) {
                const bool act = t < Tn;
                // Bulk copies take uniform-register operands: per-lane addresses run as a 32-iteration loop in the issuing warp.
                // The two loads of a forward timestep are issued by the SECOND owner's warp, the reduce-adds (and the chain
                // record's load behind its reduce-add, same slot) by the lead owner's -- an even split of those loops.
                uint32_t out = 0, c0 = 0, idD = 0;
                out = act ? idB : 0u;
                if (t + 3 <= Tn) idD = __ldg(ids + t + 3);
                c0 = draw_item(key, step, (uint32_t)t, 0u, pl.neg_range);
                fence_async_smem();          // Z_t (written by every owner during the previous step) -> tensor-core proxy
                tc_fence_before_sync();
                tile_bar();
                if (issuer) {
                    tc_fence_after_sync();
#pragma unroll
                    for (int k = 0; k < kKP / 16; ++k)
                        mma_bf16(tcol0, make_smem_desc(zb_a + k * 256, 128, 1280), make_smem_desc(wb_a + k * 4096, 2048, 128), IDESC_F, k > 0);
                    mma_commit(mbar + tile);
                }
                // the records this timestep scores against: the target row {bias, w} and the first candidate (full record: it
                // becomes the negative's optimizer visit); the slots were released by the previous step
                if (!lead) {                 // (the lead owner waited for its reduce-add to have read the slot before the tile barrier)
                    bulk_load(pslot, trec<FLAT>(tb, out), PSLOT, qbar);
                    bulk_load(rslot, trec<FLAT>(tb, c0), REC, qbar);
                    if (lane == 0) mbar_arrive_expect_tx(qbar, 32 * (PSLOT + REC));
                }
                mbar_wait(mbar + tile, phase); phase ^= 1;
                tc_fence_after_sync();
#pragma unroll
                for (int b = 0; b < NB8; ++b) {
                    const int gb = gb0 + b;
                    float pf[8], pi[8], pg[8], po[8], pc[8], ptc[8];
                    tmem_ld8(tcs + b * 8, pc);   // c_{t-1}
                    tmem_ld8x4(tbase + gb * 8, tbase + 32 + gb * 8, tbase + 64 + gb * 8, tbase + 96 + gb * 8, pf, pi, pg, po);
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {   // two hidden units per packed instruction
                        const float2 f = sigm2(make_float2(pf[j], pf[j + 1]));
                        const float2 ig = coupled ? fma2(f, splat2(-1.0f), splat2(1.0f)) : sigm2(make_float2(pi[j], pi[j + 1]));
                        const float2 gg = tnh2(make_float2(pg[j], pg[j + 1]));
                        const float2 og = sigm2(make_float2(po[j], po[j + 1]));
                        const float2 cn = fma2(f, make_float2(pc[j], pc[j + 1]), mul2(ig, gg));
                        const float2 tcn = tnh2(cn);
                        const float2 hn = mul2(og, tcn);
                        h[b * 8 + j] = act ? hn.x : 0.0f; h[b * 8 + j + 1] = act ? hn.y : 0.0f;
                        pf[j] = f.x; pf[j + 1] = f.y; pi[j] = ig.x; pi[j + 1] = ig.y; pg[j] = gg.x; pg[j + 1] = gg.y;
                        po[j] = og.x; po[j + 1] = og.y; pc[j] = cn.x; pc[j + 1] = cn.y; ptc[j] = tcn.x; ptc[j + 1] = tcn.y;
                    }
                    tmem_st8(tcs + b * 8, pc);   // (finished sequences carry garbage from here on: never read again as a live value)
                    if (act) {
                        *sb8(t, AF, gb) = pack_bf16x8(pf); *sb8(t, AI, gb) = pack_bf16x8(pi);
                        *sb8(t, AG, gb) = pack_bf16x8(pg); *sb8(t, AO, gb) = pack_bf16x8(po);
                        *sb8(t, AC, gb) = pack_bf16x8(pc); *sb8(t, ATC, gb) = pack_bf16x8(ptc);
                    }
                }
                tmem_st_wait();
                tc_fence_before_sync();  // TMEM reads ordered before the next MMA (issued after the next tile barrier)
                // ---- scoring + negative sampling (sequence_model.rs:47-68, lstm.rs:300-320) ----
                rec_wait();
                float4 pv[NCH], qv[NCH];
                float pos;
                {
                    float a = lead ? *reinterpret_cast<const float*>(pslot_g) : 0.0f;   // b[out]
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        pv[cc] = *reinterpret_cast<const float4*>(pslot_g + 16 + (part * NCH + cc) * 16);
                        a = fmaf(h[4 * cc], pv[cc].x, a); a = fmaf(h[4 * cc + 1], pv[cc].y, a);
                        a = fmaf(h[4 * cc + 2], pv[cc].z, a); a = fmaf(h[4 * cc + 3], pv[cc].w, a);
                    }
                    pos = xsum(a);
                }
                bool done = !act; uint32_t neg = c0; float ngs = 0.0f;
                auto score = [&]() {
                    float a = lead ? *reinterpret_cast<const float*>(rslot_g) : 0.0f;   // b[candidate]
                    float4 qt[NCH];
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        qt[cc] = rs_ld(0, cc);
                        a = fmaf(h[4 * cc], qt[cc].x, a); a = fmaf(h[4 * cc + 1], qt[cc].y, a);
                        a = fmaf(h[4 * cc + 2], qt[cc].z, a); a = fmaf(h[4 * cc + 3], qt[cc].w, a);
                    }
                    const float tot = xsum(a);
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
                // further candidates only while some sequence of the quad still has none that violates the margin; the
                // decisions are bit-identical in both owners, so both warps of a quad take the same path
                for (int j = 1; j < tries; ++j) {
                    if (__all_sync(kFull, done)) break;
                    {
                        const uint32_t cj = draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range);
                        if (!done) { neg = cj; if (!lead) bulk_load(rslot, trec<FLAT>(tb, cj), REC, qbar); }
                        const uint32_t n = __popc(__ballot_sync(kFull, !done));
                        if (!lead && lane == 0) mbar_arrive_expect_tx(qbar, n * REC);
                    }
                    rec_wait();
                    score();
                }
                float g = 0.0f;
                if (act) {
                    float l;
                    if (m.loss == 0) { const float s = sigm(ngs - pos); l = s; g = s * (1.0f - s); }
                    else { const float v = 1.0f + ngs - pos; l = v > 0.0f ? v : 0.0f; g = v > 0.0f ? 1.0f : 0.0f; }
                    loss_seq += l;
                    if (lead) G_[(size_t)t * gstride] = g;
                }
                // ---- Z_{t+1} = [h_t, x_{t+1} = E[out_t]] into the Z tile (its MMA has completed); activation copies ----
#pragma unroll
                for (int b = 0; b < NB8; ++b) {
                    const int gb = gb0 + b;
                    const float4 qa = qv[2 * b], qb = qv[2 * b + 1], pa = pv[2 * b], pb = pv[2 * b + 1];
                    const uint4 hp = make_uint4(pack2(h[8 * b], h[8 * b + 1]), pack2(h[8 * b + 2], h[8 * b + 3]), pack2(h[8 * b + 4], h[8 * b + 5]), pack2(h[8 * b + 6], h[8 * b + 7]));
                    const uint4 xp = pack8(pa, pb);
                    *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, gb, 10)) = hp;
                    *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 4 + gb, 10)) = xp;
                    if (act) {
                        *sb8(t, AHB, gb) = hp;
                        if (t + 1 < Tn) *sb8(t + 1, AX, gb) = xp;
                        const float2 g2 = splat2(g), ng2 = splat2(-g);   // g (q - p) = g q - g p
                        const float2 d0 = fma2(g2, make_float2(qa.x, qa.y), mul2(ng2, make_float2(pa.x, pa.y)));
                        const float2 d1 = fma2(g2, make_float2(qa.z, qa.w), mul2(ng2, make_float2(pa.z, pa.w)));
                        const float2 d2 = fma2(g2, make_float2(qb.x, qb.y), mul2(ng2, make_float2(pb.x, pb.y)));
                        const float2 d3 = fma2(g2, make_float2(qb.z, qb.w), mul2(ng2, make_float2(pb.z, pb.w)));
                        *sb8(t, ADQ, gb) = make_uint4(pack2(d0.x, d0.y), pack2(d1.x, d1.y), pack2(d2.x, d2.y), pack2(d3.x, d3.y));
                    }
                }
                // ---- the negative's visit, in its slot: E[neg_t] += step(+g h_t), b[neg_t] += step(+g) ----
                if (act) {
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        if (!ADAM) {
                            const float4 G4 = rs_ld(1, cc);
                            float2 wa = make_float2(qv[cc].x, qv[cc].y), wb_ = make_float2(qv[cc].z, qv[cc].w);
                            float2 Ga = make_float2(G4.x, G4.y), Gb = make_float2(G4.z, G4.w), dwa, dwb, dGa, dGb;
                            adagrad2(wa, Ga, mul2(splat2(g), make_float2(h[4 * cc], h[4 * cc + 1])), o, dwa, dGa);
                            adagrad2(wb_, Gb, mul2(splat2(g), make_float2(h[4 * cc + 2], h[4 * cc + 3])), o, dwb, dGb);
                            rs_st(0, cc, make_float4(dwa.x, dwa.y, dwb.x, dwb.y)); rs_st(1, cc, make_float4(dGa.x, dGa.y, dGb.x, dGb.y));
                        } else {
                            float4 w = qv[cc], s1 = rs_ld(1, cc), s2 = rs_ld(2, cc);
                            const float4 gh = make_float4(g * h[4 * cc], g * h[4 * cc + 1], g * h[4 * cc + 2], g * h[4 * cc + 3]);
                            apply4(w, s1, s2, gh, 1.0f, o);
                            rs_st(0, cc, w); rs_st(1, cc, s1); rs_st(2, cc, s2);   // Adam: new values, plain store (below)
                        }
                    }
                    if (lead) {
                        float4 bq = *reinterpret_cast<const float4*>(rslot_g);
                        const float4 b0 = bq;
                        if (!o.adam) adagrad1(bq.x, bq.y, g, o); else adam1(bq.x, bq.y, bq.z, g, o);
                        *reinterpret_cast<float4*>(rslot_g) = ADAM ? bq : make_float4(bq.x - b0.x, bq.y - b0.y, 0.0f, 0.0f);
                    }
                }
                fence_async_smem();      // the deltas in the slot (and Z_{t+1}) -> async proxy
                quad_bar();              // both owners have written their halves / are done reading the slots
                if (lead) {
                    if (act) {
                        if (ADAM) bulk_store(trec<FLAT>(tb, neg), rslot, REC); else bulk_reduce_add(trec<FLAT>(tb, neg), rslot, REC);
                        bulk_commit();
                    }
                    bulk_wait_read();    // the slot is free again before this owner reaches the next tile barrier
                }
                idB = idC; idC = idD;
            }

            // =========================== backward ===========================
            // dz of timestep t+1 (dh_t in TMEM columns 0..31, dx_{t+1} in 32..63) is consumed straight from TMEM inside
            // timestep t's delta loop -- nothing but the cell-gradient recurrence lives across timesteps (in TMEM).
            {
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int b = 0; b < NB8; ++b) tmem_st8(tcs + b * 8, z8);   // dc_t recurrence starts at 0
                tmem_st_wait();
            }
            // chain rows: the visit of backward timestep t touches row ids[t+1] (entries dx_{t+1} and -g_t h_t); its record is
            // requested one timestep earlier.  t = -1 is a pseudo-timestep: only E[ids[0]] += step(dx_0).
            auto chain_id = [&](int t) -> uint32_t { return (t + 1 <= Tn && Tn > 0) ? __ldg(ids + t + 1) : 0u; };   // lead only
            uint32_t row_c = 0;
            if (Tmax > 0 && lead) {
                row_c = chain_id(Tmax - 1);
                bulk_load(rslot, trec<FLAT>(tb, row_c), REC, qbar);
                if (lane == 0) mbar_arrive_expect_tx(qbar, 32 * REC);
            }
            float g_c = 0.0f;
            ActB cur;
            {
                const int t = Tmax - 1;
                const bool a0 = t >= 0 && t < Tn;
                if (a0) g_c = __ldcg(G_ + (size_t)t * gstride);
                load_act(cur, t > 0 ? t : 0, gb0, a0);
            }
            bool prev_valid = false, prev_act = false;   // a dz of the previous (later) timestep is pending in TMEM
            for (int t = Tmax - 1; t >= (Tmax > 0 ? -1 : 0); --t) {
                const bool act = t >= 0 && t < Tn;
                const float g = g_c;
                const bool actn = t >= 1 && (t - 1) < Tn;   // the next (earlier) timestep
                float g_n = 0.0f;
                if (actn) g_n = __ldcg(G_ + (size_t)(t - 1) * gstride);
                if (t >= 1) prefetch_step(t - 1);
                const bool has_dx = t + 1 < Tn;   // a deferred E[in_{t+1}] entry exists (t + 1 >= 0 always)
                float4 dxv[NCH], ghv[NCH];
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) { dxv[cc] = zero4; ghv[cc] = zero4; }
                if (prev_valid) { mbar_wait(mbar + tile, phase); phase ^= 1; tc_fence_after_sync(); }
                if (t >= 0) {
                    stage_z_async(t, act);   // the previous MMA is done with the Z tile
#pragma unroll
                    for (int b = 0; b < NB8; ++b) {
                        const int gb = gb0 + b;
                        if (b > 0) load_act(cur, t, gb, act);
                        float dhv[8], dcv[8];
                        tmem_ld8(tcs + b * 8, dcv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) dhv[e] = 0.0f;
                        if (prev_valid) {   // dh_t and dx_{t+1}
                            uint32_t ra[8], rb[8];
                            tmem_ld8_issue(tbase + gb * 8, ra); tmem_ld8_issue(tbase + 32 + gb * 8, rb);
                            tmem_wait8(ra); tmem_wait8(rb);
#pragma unroll
                            for (int e = 0; e < 8; ++e) dhv[e] = prev_act ? __uint_as_float(ra[e]) : 0.0f;
                            dxv[2 * b] = make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3]));
                            dxv[2 * b + 1] = make_float4(__uint_as_float(rb[4]), __uint_as_float(rb[5]), __uint_as_float(rb[6]), __uint_as_float(rb[7]));
                        }
                        // gradient of the target row: g h_t (h_t from its bf16 copy)
                        {   // (stored negated: the entry of the target row is -g h_t)
                            const float2 ng2 = splat2(-g);
                            const float2 a0_ = mul2(ng2, bf2(cur.h.x)), a1_ = mul2(ng2, bf2(cur.h.y)), a2_ = mul2(ng2, bf2(cur.h.z)), a3_ = mul2(ng2, bf2(cur.h.w));
                            ghv[2 * b] = make_float4(a0_.x, a0_.y, a1_.x, a1_.y);
                            ghv[2 * b + 1] = make_float4(a2_.x, a2_.y, a3_.x, a3_.y);
                        }
                        uint32_t wdf[4], wdi[4], wdg[4], wdo[4];
#pragma unroll
                        for (int pr = 0; pr < 4; ++pr) {   // two hidden units at a time, straight from / to packed bf16 words
                            const uint32_t uf = word_of(cur.f, pr), ui = word_of(cur.i, pr), ug = word_of(cur.g, pr), uo = word_of(cur.o, pr);
                            const uint32_t uq = word_of(cur.q, pr), ucp = word_of(cur.cp, pr), utc = word_of(cur.tc, pr);
                            const float2 f_ = bf2(uf), i_ = bf2(ui), g_ = bf2(ug), o_ = bf2(uo), q_ = bf2(uq), cp_ = bf2(ucp), tcv = bf2(utc);
                            const int e = 2 * pr;
                            const float2 one = splat2(1.0f);
                            const float2 dh = add2(make_float2(dhv[e], dhv[e + 1]), q_);
                            const float2 d_o = mul2(dh, tcv);
                            const float2 dc = fma2(mul2(dh, o_), fma2(mul2(tcv, splat2(-1.0f)), tcv, one), make_float2(dcv[e], dcv[e + 1]));
                            float2 d_f = mul2(dc, cp_), d_i = mul2(dc, g_);
                            const float2 d_g = mul2(dc, i_);
                            const float2 dcn = mul2(dc, f_);
                            dcv[e] = act ? dcn.x : 0.0f; dcv[e + 1] = act ? dcn.y : 0.0f;
                            if (coupled) { d_f = fma2(d_i, splat2(-1.0f), d_f); d_i = splat2(0.0f); }
                            const float2 nf = mul2(f_, splat2(-1.0f)), ni = mul2(i_, splat2(-1.0f)), ng = mul2(g_, splat2(-1.0f)), no = mul2(o_, splat2(-1.0f));
                            const float2 rdf = mul2(d_f, fma2(nf, f_, f_));                 // d_f f (1 - f)
                            const float2 rdi = coupled ? splat2(0.0f) : mul2(d_i, fma2(ni, i_, i_));
                            const float2 rdg = mul2(d_g, fma2(ng, g_, one));                // d_g (1 - g^2)
                            const float2 rdo = mul2(d_o, fma2(no, o_, o_));
                            wdf[pr] = pack2(rdf.x, rdf.y); wdi[pr] = pack2(rdi.x, rdi.y);
                            wdg[pr] = pack2(rdg.x, rdg.y); wdo[pr] = pack2(rdo.x, rdo.y);
                        }
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 0 + gb, 16)) = make_uint4(wdf[0], wdf[1], wdf[2], wdf[3]);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 4 + gb, 16)) = make_uint4(wdi[0], wdi[1], wdi[2], wdi[3]);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 8 + gb, 16)) = make_uint4(wdg[0], wdg[1], wdg[2], wdg[3]);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 12 + gb, 16)) = make_uint4(wdo[0], wdo[1], wdo[2], wdo[3]);
                        tmem_st8(tcs + b * 8, dcv);
                    }
                    tmem_st_wait();
                    cp_wait<0>();             // Z_t rows have landed
                    fence_async_smem();
                    tc_fence_before_sync();   // also orders this thread's TMEM reads of dz_{t+1} before the MMA that overwrites them
                    tile_bar();
                    if (issuer) {
                        tc_fence_after_sync();
#pragma unroll
                        for (int k = 0; k < 8; ++k)  // dz = delta . W^T
                            mma_bf16(tcol0, make_smem_desc(db_a + k * 256, 128, 2048), make_smem_desc(wb_a + k * 256, 128, 2048), IDESC_G2, k > 0);
#pragma unroll
                        for (int k = 0; k < 8; ++k)  // dW^T += delta^T . [Z, 1]
                            mma_bf16(tcol0 + 128, make_smem_desc(db_a + k * 4096, 2048, 128), make_smem_desc(zb_a + k * 2560, 1280, 128), IDESC_G3,
                                     (k > 0 || t < Tmax - 1) ? 1u : 0u);
                        mma_commit(mbar + tile);
                    }
                    // first activation block of the next timestep: requested before this timestep's visit
                    if (t >= 1) load_act(cur, t - 1, gb0, actn);
                } else {   // t == -1: dx_0 out of TMEM, nothing else
#pragma unroll
                    for (int b = 0; b < NB8; ++b) {
                        uint32_t rb[8];
                        tmem_ld8_issue(tbase + 32 + (gb0 + b) * 8, rb);
                        tmem_wait8(rb);
                        dxv[2 * b] = make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3]));
                        dxv[2 * b + 1] = make_float4(__uint_as_float(rb[4]), __uint_as_float(rb[5]), __uint_as_float(rb[6]), __uint_as_float(rb[7]));
                    }
                    tc_fence_before_sync();
                }
                // ---- the chain row's visit (overlaps the MMAs): E[ids[t+1]] += step(dx_{t+1}), then step(-g h_t); bias -g ----
                rec_wait();
                const bool visit = act || has_dx;
                if (visit) {
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        if (!ADAM) {
                            const float4 w4 = rs_ld(0, cc), G4 = rs_ld(1, cc);
                            float2 wa = make_float2(w4.x, w4.y), wb_ = make_float2(w4.z, w4.w), Ga = make_float2(G4.x, G4.y), Gb = make_float2(G4.z, G4.w);
                            float2 dwa = splat2(0.0f), dwb = dwa, dGa = dwa, dGb = dwa, t0, t1;
                            if (has_dx) {
                                adagrad2(wa, Ga, make_float2(dxv[cc].x, dxv[cc].y), o, dwa, dGa);
                                adagrad2(wb_, Gb, make_float2(dxv[cc].z, dxv[cc].w), o, dwb, dGb);
                            }
                            if (act) {   // ghv holds -g h_t
                                adagrad2(wa, Ga, make_float2(ghv[cc].x, ghv[cc].y), o, t0, t1); dwa = add2(dwa, t0); dGa = add2(dGa, t1);
                                adagrad2(wb_, Gb, make_float2(ghv[cc].z, ghv[cc].w), o, t0, t1); dwb = add2(dwb, t0); dGb = add2(dGb, t1);
                            }
                            rs_st(0, cc, make_float4(dwa.x, dwa.y, dwb.x, dwb.y)); rs_st(1, cc, make_float4(dGa.x, dGa.y, dGb.x, dGb.y));
                        } else {
                            float4 w = rs_ld(0, cc), s1 = rs_ld(1, cc), s2 = rs_ld(2, cc);
                            if (has_dx) apply4(w, s1, s2, dxv[cc], 1.0f, o);
                            if (act) apply4(w, s1, s2, ghv[cc], 1.0f, o);   // ghv holds -g h_t
                            rs_st(0, cc, w); rs_st(1, cc, s1); rs_st(2, cc, s2);
                        }
                    }
                    if (lead) {
                        float4 bq = *reinterpret_cast<const float4*>(rslot_g);
                        const float4 b0 = bq;
                        if (act) { if (!o.adam) adagrad1(bq.x, bq.y, -g, o); else adam1(bq.x, bq.y, bq.z, -g, o); }
                        *reinterpret_cast<float4*>(rslot_g) = ADAM ? bq : make_float4(bq.x - b0.x, bq.y - b0.y, 0.0f, 0.0f);
                    }
                }
                fence_async_smem();
                quad_bar();
                if (lead) {
                    if (visit) {
                        if (ADAM) bulk_store(trec<FLAT>(tb, row_c), rslot, REC); else bulk_reduce_add(trec<FLAT>(tb, row_c), rslot, REC);
                        bulk_commit();
                    }
                    if (t >= 0) {   // the next (earlier) timestep's record, as soon as the reduce-add has read the slot
                        row_c = chain_id(t - 1);
                        bulk_wait_read();
                        bulk_load(rslot, trec<FLAT>(tb, row_c), REC, qbar);
                        if (lane == 0) mbar_arrive_expect_tx(qbar, 32 * REC);
                    }
                }
                prev_valid = true; prev_act = act;
                g_c = g_n;
            }
            if (lead) bulk_wait_read();
            if (live && lead) { pl.loss_acc[p] += loss_seq; pl.examples[p] += (unsigned long long)Tn; }

            // =========================== dense step on the CTA-summed gradient ===========================
            // TMEM lane gd = r of a tile holds row gd of its dW^T: columns 0..63 = dW[k][gd], column 64 = dbias[gd];
            // owner `part` takes columns [32 part, 32 part + 32), the second owner also column 64
            __syncthreads();
            constexpr int CW = 32;
            float* xch = reinterpret_cast<float*>(smem + OFF_TILES + TILE_BYTES + TILE_DB);  // tile 1's delta area: [65][128]
            float dwr[CW + 1];
            {
                const bool have = Tmax > 0;  // a tile whose partitions are all dead issued no MMA this round
#pragma unroll
                for (int cb = 0; cb < CW / 8; ++cb) {
                    float v8[8];
                    tmem_ld8(tbase + 128 + part * CW + cb * 8, v8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dwr[cb * 8 + j] = have ? v8[j] : 0.0f;
                }
                float v8[8];
                tmem_ld8(tbase + 128 + 64, v8);
                dwr[CW] = have ? v8[0] : 0.0f;
            }
            tc_fence_before_sync();
            if (NT == 2) {
                if (tile == 1) {
#pragma unroll
                    for (int k = 0; k < CW; ++k) xch[(part * CW + k) * 128 + r] = dwr[k];
                    if (part == DS - 1) xch[64 * 128 + r] = dwr[CW];
                }
                __syncthreads();
                if (tile == 0) {
#pragma unroll
                    for (int k = 0; k < CW; ++k) dwr[k] += xch[(part * CW + k) * 128 + r];
                    dwr[CW] += xch[64 * 128 + r];
                }
            }
            if (tile == 0) {
                OptC od = o;
                od.chat = 0.0f;   // dense weights: one read-modify-write per CTA round, last writer wins
                if (od.adam) {
                    const float tt_ = (float)(pl.adam_t0 + step * pl.P + (uint64_t)blockIdx.x * NT * 128 + 1);
                    od.c1 = 1.0f - powf(0.9f, tt_); od.c2 = 1.0f - powf(0.999f, tt_);
                }
                const int gd = r;
                __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(Wb);
#pragma unroll
                for (int kk = 0; kk <= CW; ++kk) {
                    if (kk == CW && part != DS - 1) continue;
                    const int k = kk == CW ? 64 : part * CW + kk;
                    const size_t idx = (size_t)k * kNG + gd;  // k == 64: bias[gd]
                    float w = __ldcg(m.dense + idx), s1 = __ldcg(m.dense + nd + idx);
                    if (od.adam) {
                        float s2 = __ldcg(m.dense + 2 * nd + idx);
                        adam1(w, s1, s2, dwr[kk], od);
                        __stcg(m.dense + 2 * nd + idx, s2);
                    } else adagrad1(w, s1, dwr[kk], od);
                    __stcg(m.dense + idx, w); __stcg(m.dense + nd + idx, s1);
                    wb[(tile_chunk_off(k, gd >> 3, 16) >> 1) + (gd & 7)] = __float2bfloat16_rn(w);
                }
            }
            fence_async_smem();
            __syncthreads();
        }
    }
    if (live && lead) pl.step_ctr[p] = step;
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc<(NT == 1 ? 256 : 512)>(*tmem_ptr);
}

template <int NT, int S, bool FLAT>
cudaError_t launch_one(const ModelDev& m, const PlanDev& p, cudaStream_t st) {
    const size_t smem = OFF_TILES + (size_t)NT * (TILE_ST + 128 * (16 + S * 128)) + (size_t)NT * 4 * XS_BYTES_PER_QUAD;
    const int seq_per_cta = 128 * NT;
    dim3 grid((p.P + seq_per_cta - 1) / seq_per_cta);
    cudaError_t e = cudaFuncSetAttribute(lstm_tile_train_kernel<NT, S, FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lstm_tile_train_kernel<NT, S, FLAT><<<grid, seq_per_cta * 2, smem, st>>>(m, p);
    return cudaGetLastError();
}

}  // namespace

// floats of activation scratch per PARTITION for T timesteps (plan scratch is sized per partition)
size_t lstm_tile_scratch_floats_per_partition(int T) { return (size_t)T * (36 * 4 + 1); }

// Adagrad: two tiles of 128 partitions per CTA; Adam (400-byte records): one tile per CTA
int lstm_tile_tiles_per_cta(const ModelDev& m, uint32_t P) { return (m.opt == 1 || P % 256 != 0) ? 1 : 2; }

cudaError_t launch_lstm_tile(const ModelDev& m, const PlanDev& p, cudaStream_t st) {
    const bool flat = m.gmask == 0;
    const int nt = lstm_tile_tiles_per_cta(m, p.P);
    if (m.opt == 1) return flat ? launch_one<1, 3, true>(m, p, st) : launch_one<1, 3, false>(m, p, st);
    if (nt == 2) return flat ? launch_one<2, 2, true>(m, p, st) : launch_one<2, 2, false>(m, p, st);
    return flat ? launch_one<1, 2, true>(m, p, st) : launch_one<1, 2, false>(m, p, st);
}

}  // namespace sbr
