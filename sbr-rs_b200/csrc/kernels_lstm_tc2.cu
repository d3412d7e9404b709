// kernels_lstm_tc2.cu -- tensor-core LSTM training kernel (D = 32), second generation: same tcgen05 / TMEM tile
// structure and numerics as kernels_lstm_tc.cu, but every row that a timestep needs is on its way to shared memory
// before the warp asks for it.
//
// Why (profiles/r1_v5_lstm_tc_stalls_by_line.txt): the first tile kernel issues 14 % of the time; 46 % of all warp
// samples wait on the scoreboard of a global load -- the three sparse optimizer visits (12.5 %), the x / target /
// candidate gathers (11 %), the activation reloads of the backward pass (10 %) -- and 11 % at tile barriers behind
// the slowest warp.  Each of those is a dependent L2 round trip, ~25 per timestep, with 2 warps per scheduler.
//
// What changes:
//  * forward: x_{t+1}, the target row and all WARP candidates of a timestep (ids come from the counter-based
//    sampler, so they are known up front) are fetched with cp.async (LDGSTS, .cg = L2-coherent) into per-warp
//    staging slices that alias whatever operand tile is idle -- no registers are held while they fly; the ids and
//    bias scalars are loaded one step early.
//  * backward: the visit of E[in_{t+1}] is merged with the visit of E[out_t] -- they are ALWAYS the same row
//    (out_t == ids[t+1] == in_{t+1}) -- one load, two sequential optimizer applications, one store, in the reference
//    order (t descending; E[neg], E[out], E[in]).  The E[neg_t] visit shares the load batch (distinct row, or folded
//    in as a third application when neg_t == out_t).  Six dependent round trips per timestep become two.
//  * backward activations are double-buffered in registers one block ahead, the first block of timestep t-1 is
//    requested before the visits of timestep t.
//  * FASTM: gates with MUFU.TANH (tanh.approx.f32, sigmoid(x) = 0.5 tanh(x/2) + 0.5) and Adagrad with
//    rsqrt.approx.ftz -- a third of the epilogue instructions.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "engine.h"
#include "tc_tile.cuh"

namespace sbr {

namespace {

using namespace tc;

constexpr int kD = 32, kNK = 64, kNG = 128;
constexpr uint32_t OFF_WT = 0;              // tf32 [128 gd][64 feat]
constexpr uint32_t OFF_WB = 32768;          // bf16 [64 feat][128 gd]
constexpr uint32_t OFF_BIAS = 49152;        // float[128]
constexpr uint32_t OFF_MISC = 49664;        // mbarriers, tmem base, tile maxima
constexpr uint32_t OFF_TILES = 50176;
constexpr uint32_t TILE_ZT = 0;             // tf32 [128 seq][64 feat]
constexpr uint32_t TILE_DB = 32768;         // bf16 [128 seq][128 gd]
constexpr uint32_t TILE_ZB = 65536;         // bf16 [128 seq][80 feat]
constexpr uint32_t TILE_BYTES = 86016;

// ---------------------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <bool F> __device__ __forceinline__ float sigm(float x) {
    if (F) return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f);
    return __fdividef(1.0f, 1.0f + __expf(-x));
}
template <bool F> __device__ __forceinline__ float tnh(float x) {
    if (F) return tanh_approx(x);
    return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f;
}
__device__ __forceinline__ void adagrad1(float& w, float& G, float g, float lr, float l2) {
    g = fmaf(w, l2, g);
    G = fmaf(g, g, G);
    w = fmaf(-lr * g, rsqrt_approx(fmaxf(G, 1e-20f)), w);  // lr / (1e-10 + sqrt(G)) * g
}
__device__ __forceinline__ void adam1(float& w, float& m, float& v, float g, const OptCfg& o) {
    g = fmaf(w, o.l2, g);
    m = 0.9f * m + 0.1f * g;
    v = 0.999f * v + 0.001f * g * g;
    const float mhat = __fdividef(m, o.c1), vhat = __fdividef(v, o.c2);
    w -= __fdividef(o.lr * mhat, sqrtf(vhat) + 1e-8f);
}
// one optimizer application on a 16-byte piece of a row record: w, s (Adagrad G / Adam m), v (Adam v), gradient sign*g
__device__ __forceinline__ void apply4(float4& w, float4& s, float4& v, const float4& g, float sign, const OptCfg& o) {
    if (!o.adam) {
        adagrad1(w.x, s.x, sign * g.x, o.lr, o.l2); adagrad1(w.y, s.y, sign * g.y, o.lr, o.l2);
        adagrad1(w.z, s.z, sign * g.z, o.lr, o.l2); adagrad1(w.w, s.w, sign * g.w, o.lr, o.l2);
    } else {
        adam1(w.x, s.x, v.x, sign * g.x, o); adam1(w.y, s.y, v.y, sign * g.y, o);
        adam1(w.z, s.z, v.z, sign * g.z, o); adam1(w.w, s.w, v.w, sign * g.w, o);
    }
}

// L2 atomics for the Hogwild Adagrad visits.  A visit as load -> compute -> store makes every hot row a read-modify-write
// hazard between thousands of concurrent sequences (1,683 rows, 37,888 partitions): measured on the ML-100K-shaped
// stream the stores alone cost 7 ms and the 16-byte bias records 11 ms of a 44 ms epoch (ablation in DESIGN.md 3.4).
// The accumulator G is additive, so the visit becomes: {atom.add G += sum g^2 (returns the old G), ld w} in one round
// trip, the sequential Adagrad applications in registers from the returned G, then two fire-and-forget reductions
// (w += sum of steps; G += the l2 cross terms).  No update of G is ever lost, and the L2 atomic unit serialises hot
// addresses at about a cycle per operation instead of a memory round trip.
__device__ __forceinline__ float4 atom_add4(float* p, const float4& v) {
    float4 o;
    asm volatile("atom.global.add.v4.f32 {%0, %1, %2, %3}, [%4], {%5, %6, %7, %8};"
                 : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    return o;
}
__device__ __forceinline__ void red_add4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 sq4(const float4& a) { return make_float4(a.x * a.x, a.y * a.y, a.z * a.z, a.w * a.w); }
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// ---------------------------------------------------------------------------------------------------------
// Staging slices.  A slice holds 32 rows x 128 B in core-matrix order: row r, 16-byte chunk c at
//   (r >> 3) * gs + c * 128 + (r & 7) * 16          (gs = 1024 compact; 1280 inside the bf16 Z tile)
// Lane-per-row 16-byte accesses (8 consecutive rows = 128 contiguous bytes) and the cooperative mapping below
// (8 rows x 4 chunks per instruction: for a fixed chunk 8 rows are again 128 contiguous bytes) are both free of
// bank conflicts.  Every slice lives in the part of an operand tile that holds the warp's OWN 32 rows, so staging
// never needs a barrier wider than the warp.
// ---------------------------------------------------------------------------------------------------------
struct Slice { uint8_t* p; uint32_t gs; };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Table { float* e0; uint32_t stride; bool flat; };   // item_rec() with the unsharded case resolved once
__device__ __forceinline__ float* trec(const ModelDev& m, const Table& tb, uint32_t id) {
    return tb.flat ? tb.e0 + (size_t)id * tb.stride : item_rec(m, id);
}

// rows named by the lanes' ids -> slice, asynchronously (lane l moves chunk (l >> 3) and (l >> 3) + 4 of rows 8g + (l & 7))
__device__ __forceinline__ void gather_async(const ModelDev& m, const Table& tb, uint32_t my_id, int lane, const Slice& s) {
    const int rl = lane & 7, ch = lane >> 3;
    const uint32_t base = smem_u32(s.p) + ch * 128 + rl * 16;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t id = __shfl_sync(kFull, my_id, 8 * g + rl);
        const float* src = trec(m, tb, id) + ch * 4;
        cp_async16(base + g * s.gs, src);
        cp_async16(base + g * s.gs + 512, src + 16);
    }
}
__device__ __forceinline__ float4 slice_ld(const Slice& s, int lane, int c) {
    return *reinterpret_cast<const float4*>(s.p + (lane >> 3) * s.gs + c * 128 + (lane & 7) * 16);
}
__device__ __forceinline__ void slice_st(const Slice& s, int lane, int c, const float4& v) {
    *reinterpret_cast<float4*>(s.p + (lane >> 3) * s.gs + c * 128 + (lane & 7) * 16) = v;
}
__device__ __forceinline__ void slice_read_row(const Slice& s, int lane, float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 t = slice_ld(s, lane, c);
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
}

// ---------------------------------------------------------------------------------------------------------
// The sparse visits of one backward timestep (sequence_model.rs:163-169 -> wyrm sparse optimizer, un-merged entries):
//   V_neg: E[neg_t]  += step(+g h_t)                                   flag bit 0
//   V_out: E[out_t]  += step(dx_{t+1}) [bit 2: a deferred E[in_{t+1}] entry exists], then
//                       step(+g h_t)   [bit 3: neg_t == out_t, the E[neg_t] entry lands on this row], then
//                       step(-g h_t)   [bit 4: timestep t is active]    flag bit 1 = any of the three
// (the pseudo-timestep t = -1 carries only the last deferred entry, E[in_0] += step(dx_0): out_{-1} == ids[0])
// Loads of both visits for four (row-group, half-row) items are issued before the first use.  Two different
// sequences of the warp naming the same row is the usual Hogwild race (last store wins).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void coop_visits(const ModelDev& m, const Table& tb, uint32_t neg, uint32_t out, uint32_t fl, int lane,
                                            const Slice& gh, const Slice& dx, const OptCfg& o, const bool noatom = false) {
    const int rl = lane & 7, ch = lane >> 3;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        float* rn[2]; float* ro[2]; uint32_t f[2];
        float4 wn[4], sn[4], wo[4], so[4];
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
            const int row = 8 * (pass * 2 + gg) + rl;
            const uint32_t idn = __shfl_sync(kFull, neg, row), ido = __shfl_sync(kFull, out, row);
            f[gg] = __shfl_sync(kFull, fl, row);
            rn[gg] = trec(m, tb, idn) + ch * 4; ro[gg] = trec(m, tb, ido) + ch * 4;
        }
        const bool atomics = !o.adam && !noatom;
        if (atomics) {
            // ---- Adagrad through L2 atomics: {ld w, atom G += sum g^2} -> sequential applications -> red w, red G ----
            float4 gn[4], go[4];   // sum of squared raw gradients of the visit (what the atom adds up front)
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int gg = it >> 1, hf = it & 1;
                const uint32_t off = (uint32_t)(pass * 2 + gg) * 1024u + (uint32_t)(ch + 4 * hf) * 128u + (uint32_t)rl * 16u;
                const float4 g4 = *reinterpret_cast<const float4*>(gh.p + off);
                const float4 s4 = sq4(g4);
                if (f[gg] & 1u) {
                    gn[it] = s4;
                    wn[it] = __ldcg(reinterpret_cast<const float4*>(rn[gg] + hf * 16));
                    sn[it] = atom_add4(rn[gg] + hf * 16 + kD, s4);
                }
                if (f[gg] & 2u) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (f[gg] & 4u) t = sq4(*reinterpret_cast<const float4*>(dx.p + off));
                    if (f[gg] & 8u) t = add4(t, s4);
                    if (f[gg] & 16u) t = add4(t, s4);
                    go[it] = t;
                    wo[it] = __ldcg(reinterpret_cast<const float4*>(ro[gg] + hf * 16));
                    so[it] = atom_add4(ro[gg] + hf * 16 + kD, t);
                }
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int gg = it >> 1, hf = it & 1;
                const uint32_t off = (uint32_t)(pass * 2 + gg) * 1024u + (uint32_t)(ch + 4 * hf) * 128u + (uint32_t)rl * 16u;
                const float4 g4 = *reinterpret_cast<const float4*>(gh.p + off);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f[gg] & 1u) {
                    const float4 w0 = wn[it], G0 = sn[it];
                    apply4(wn[it], sn[it], v, g4, 1.0f, o);
                    red_add4(rn[gg] + hf * 16, sub4(wn[it], w0));
                    if (o.l2 != 0.0f) red_add4(rn[gg] + hf * 16 + kD, sub4(sub4(sn[it], G0), gn[it]));
                }
                if (f[gg] & 2u) {
                    const float4 w0 = wo[it], G0 = so[it];
                    if (f[gg] & 4u) { const float4 d4 = *reinterpret_cast<const float4*>(dx.p + off); apply4(wo[it], so[it], v, d4, 1.0f, o); }
                    if (f[gg] & 8u) apply4(wo[it], so[it], v, g4, 1.0f, o);
                    if (f[gg] & 16u) apply4(wo[it], so[it], v, g4, -1.0f, o);
                    red_add4(ro[gg] + hf * 16, sub4(wo[it], w0));
                    if (o.l2 != 0.0f) red_add4(ro[gg] + hf * 16 + kD, sub4(sub4(so[it], G0), go[it]));
                }
            }
            continue;
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int gg = it >> 1, hf = it & 1;
            if (f[gg] & 1u) {
                wn[it] = __ldcg(reinterpret_cast<const float4*>(rn[gg] + hf * 16));
                sn[it] = __ldcg(reinterpret_cast<const float4*>(rn[gg] + hf * 16 + kD));
            }
            if (f[gg] & 2u) {
                wo[it] = __ldcg(reinterpret_cast<const float4*>(ro[gg] + hf * 16));
                so[it] = __ldcg(reinterpret_cast<const float4*>(ro[gg] + hf * 16 + kD));
            }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int gg = it >> 1, hf = it & 1;
            const uint32_t off = (uint32_t)(pass * 2 + gg) * 1024u + (uint32_t)(ch + 4 * hf) * 128u + (uint32_t)rl * 16u;
            const float4 g4 = *reinterpret_cast<const float4*>(gh.p + off);
            if (f[gg] & 1u) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o.adam) v = __ldcg(reinterpret_cast<const float4*>(rn[gg] + hf * 16 + 2 * kD));
                apply4(wn[it], sn[it], v, g4, 1.0f, o);
                __stcg(reinterpret_cast<float4*>(rn[gg] + hf * 16), wn[it]);
                __stcg(reinterpret_cast<float4*>(rn[gg] + hf * 16 + kD), sn[it]);
                if (o.adam) __stcg(reinterpret_cast<float4*>(rn[gg] + hf * 16 + 2 * kD), v);
            }
            if (f[gg] & 2u) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o.adam) v = __ldcg(reinterpret_cast<const float4*>(ro[gg] + hf * 16 + 2 * kD));
                if (f[gg] & 4u) {
                    const float4 d4 = *reinterpret_cast<const float4*>(dx.p + off);
                    apply4(wo[it], so[it], v, d4, 1.0f, o);
                }
                if (f[gg] & 8u) apply4(wo[it], so[it], v, g4, 1.0f, o);
                if (f[gg] & 16u) apply4(wo[it], so[it], v, g4, -1.0f, o);
                __stcg(reinterpret_cast<float4*>(ro[gg] + hf * 16), wo[it]);
                __stcg(reinterpret_cast<float4*>(ro[gg] + hf * 16 + kD), so[it]);
                if (o.adam) __stcg(reinterpret_cast<float4*>(ro[gg] + hf * 16 + 2 * kD), v);
            }
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void tile_bar(int tile) { asm volatile("bar.sync %0, 128;" ::"r"(tile + 1) : "memory"); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}

// one 8-d block of a timestep's saved activations (backward operand)
struct ActB {
    uint4 f, i, g, o, q, cp, tc;   // gates, g (q - p), c_{t-1}, tanh(c_t): bf16 x 8
    float4 h0, h1;                  // h_t fp32
};

// TIMING: a debug instantiation in which thread 0 of CTA 0 prints clock64() deltas between the phases of one timestep
template <int NT, bool FASTM, bool TIMING>
__global__ void __launch_bounds__(128 * NT, 1) lstm_tc2_train_kernel(ModelDev m, PlanDev pl) {
    long long tk[24];
    auto tick = [&](int i) { if (TIMING) tk[i] = clock64(); };
#pragma unroll
    for (int i = 0; i < 24; ++i) tk[i] = 0;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Wt = smem + OFF_WT;
    uint8_t* Wb = smem + OFF_WB;
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_MISC);           // [NT]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 32);
    int* tmax_s = reinterpret_cast<int*>(smem + OFF_MISC + 40);              // [NT]

    const int tid = threadIdx.x, tile = tid >> 7, r = tid & 127, wq = (tid >> 5) & 3, lane = tid & 31;
    uint8_t* Zt = smem + OFF_TILES + tile * TILE_BYTES + TILE_ZT;
    uint8_t* Db = smem + OFF_TILES + tile * TILE_BYTES + TILE_DB;
    uint8_t* Zb = smem + OFF_TILES + tile * TILE_BYTES + TILE_ZB;
    // staging slices in the warp's own rows of the operand tiles (see Slice)
    const Slice SD0{Db + wq * 8192, 1024u}, SD1{Db + wq * 8192 + 4096, 1024u};   // forward: delta tile is idle
    const Slice SB{Zb + wq * 5120, 1280u};                                       // forward: bf16 Z tile is idle (chunks 0..7)
    const Slice SZ0{Zt + wq * 8192, 1024u}, SZ1{Zt + wq * 8192 + 4096, 1024u};   // tf32 Z tile: idle between MMA and next step / all of backward
    const uint32_t tile_gid = blockIdx.x * NT + tile;
    const uint32_t p = tile_gid * 128u + r;
    const bool live = p < pl.P;
    const size_t nd = m.ndense;
    const bool coupled = m.variant == 1;
    const int T = m.T;
    Table tb; tb.e0 = m.Es[0]; tb.stride = (uint32_t)(m.S * m.D); tb.flat = m.gmask == 0;
    // tile scratch, 16-byte structure-of-arrays so that lane == sequence accesses are coalesced (512 B per warp
    // instruction).  Per timestep: H fp32 ([8 chunks][128 seq] float4) then nine bf16 arrays F, I, G, O, X, DQ, C, TANH(C), H
    // ([9][4 chunks][128 seq] uint4) = 88 KB per tile-timestep; then G, NEG [T][128].  The backward products are bf16
    // anyway (delta tile, Z tile), so everything that only feeds a delta is kept in bf16; h_t stays fp32 for the
    // embedding-row gradients g h_t.  X and the bf16 H are copied global -> Z tile by cp.async without touching registers.
    // Layout is TIMESTEP-major across the grid's tiles -- [T][tiles][88 KB] -- because all tiles walk their sequences
    // at about the same t: the chip's hot scratch set is then ~tiles x 88 KB = 26 MB of contiguous address space
    // (a dozen 2 MB pages) instead of one 88 KB window in each of 296 4-MB-apart tile regions (~600 pages, far beyond
    // the TLB reach).
    const size_t ntiles = (size_t)gridDim.x * NT;
    constexpr int kStepU4 = (8 + 36) * 128;   // uint4 units per tile-timestep
    uint4* sbase = reinterpret_cast<uint4*>(pl.scratch);
    float* G_ = reinterpret_cast<float*>(sbase + (size_t)T * ntiles * kStepU4) + (size_t)tile_gid * 128 + r;   // + t * gstride
    uint32_t* NEG = reinterpret_cast<uint32_t*>(G_ + (size_t)T * ntiles * 128);
    const size_t gstride = ntiles * 128;
    enum { AF = 0, AI = 1, AG = 2, AO = 3, AX = 4, ADQ = 5, AC = 6, ATC = 7, AHB = 8 };
    auto step_base = [&](int t) -> uint4* { return sbase + ((size_t)t * ntiles + tile_gid) * kStepU4; };
    auto sf4 = [&](int t, int c4) -> float4* { return reinterpret_cast<float4*>(step_base(t) + (size_t)c4 * 128 + r); };
    auto sb8 = [&](int t, int which, int c8) -> uint4* { return step_base(t) + (size_t)(8 + which * 4 + c8) * 128 + r; };
    auto prefetch_step = [&](int t) {   // one timestep of the tile's scratch (88 KB = 704 lines) towards L2: up to 6 lines per thread
        const char* base = reinterpret_cast<const char*>(step_base(t));
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int line = r * 6 + i;
            if (line < 704) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)line * 128));
        }
    };
    auto load_act = [&](ActB& a, int t, int db, bool on) {
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u); const float4 zf = make_float4(0.f, 0.f, 0.f, 0.f);
        a.f = a.i = a.g = a.o = a.q = a.cp = a.tc = z4; a.h0 = a.h1 = zf;
        if (on) {
            a.f = __ldcg(sb8(t, AF, db)); a.i = __ldcg(sb8(t, AI, db)); a.g = __ldcg(sb8(t, AG, db)); a.o = __ldcg(sb8(t, AO, db));
            a.q = __ldcg(sb8(t, ADQ, db)); a.tc = __ldcg(sb8(t, ATC, db));
            a.h0 = __ldcg(sf4(t, 2 * db)); a.h1 = __ldcg(sf4(t, 2 * db + 1));
            if (t > 0) a.cp = __ldcg(sb8(t - 1, AC, db));
        }
    };
    // Z_t = [h_{t-1}, x_t] (bf16) straight from the scratch into this thread's row of the bf16 Z tile; rows of finished
    // sequences are zero-filled (src-size 0): their deltas are 0, but 0 x stale bits must not become NaN in dW
    auto stage_z_async = [&](int t, bool on) {
        const uint32_t nb = on ? 16u : 0u;
#pragma unroll
        for (int db = 0; db < 4; ++db) {
            const uint32_t dx_ = smem_u32(Zb + tile_chunk_off(r, 4 + db, 10)), dh_ = smem_u32(Zb + tile_chunk_off(r, db, 10));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dx_), "l"(sb8(t, AX, db)), "r"(nb) : "memory");
            if (t > 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dh_), "l"(sb8(t - 1, AHB, db)), "r"(nb) : "memory");
            else *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, db, 10)) = make_uint4(0u, 0u, 0u, 0u);
        }
        cp_commit();
    };

    // ---- one-time setup ----
    if (tid < 32) tmem_alloc<(NT == 1 ? 256 : 512)>(tmem_ptr);
    if (tid == 0) { for (int i = 0; i < NT; ++i) mbar_init(mbar + i, 1); fence_mbar_init(); }
    {   // constant columns of the bf16 Z tile: col 64 = 1 (bias gradient), 65..79 = 0
        const float one8[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 8, 10)) = pack_bf16x8(one8);
        *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 9, 10)) = pack_bf16x8(zero8);
    }
    if (tile == 0) {  // weights: thread gd stages column gd of W (both operand tiles) and bias[gd]
        const int gd = r;
        __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(Wb);
        for (int c4 = 0; c4 < 16; ++c4) {
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                wv[j] = __ldcg(m.dense + (size_t)(4 * c4 + j) * kNG + gd);
                wb[(tile_chunk_off(4 * c4 + j, gd >> 3, 16) >> 1) + (gd & 7)] = __float2bfloat16_rn(wv[j]);
            }
            *reinterpret_cast<float4*>(Wt + tile_chunk_off(gd, c4, 16)) =
                make_float4(to_tf32(wv[0]), to_tf32(wv[1]), to_tf32(wv[2]), to_tf32(wv[3]));
        }
        bias_s[gd] = __ldcg(m.dense + (size_t)kNK * kNG + gd);
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_ptr + (uint32_t)tile * 256u + ((uint32_t)(wq * 32) << 16);  // this thread's lane, tile's columns
    const uint32_t tcol0 = *tmem_ptr + (uint32_t)tile * 256u;                                // for the MMA issuer
    const uint32_t zt_a = smem_u32(Zt), db_a = smem_u32(Db), zb_a = smem_u32(Zb), wt_a = smem_u32(Wt), wb_a = smem_u32(Wb);
    constexpr uint32_t IDESC_G1 = make_idesc_tf32(128, 128, 0, 0);
    constexpr uint32_t IDESC_G2 = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t IDESC_G3 = make_idesc_bf16(128, 80, 1, 1);
    uint32_t phase = 0;

    XorShift rng; rng.x = rng.y = rng.z = rng.w = 1; uint64_t key = 0; uint32_t* ord = nullptr;
    uint64_t step = pl.step_ctr[live ? p : 0];
    if (live) { rng = pl.rng[p]; key = pl.keys[p]; ord = pl.order + (size_t)p * pl.n; }
    float loss_acc = 0.0f; unsigned long long ex = 0;
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;
    const int tries = m.loss == 2 ? 5 : 1;

    for (int ep = 0; ep < pl.epochs; ++ep) {
        if (live) {  // thread_rng.shuffle(partition)  sequence_model.rs:109
            uint32_t i = pl.n;
            while (i >= 2) {
                i -= 1;
                const uint32_t j = (uint32_t)xs_gen_below(rng, (uint64_t)i + 1);
                const uint32_t a = ord[i], b = ord[j];
                ord[i] = b; ord[j] = a;
            }
        }
        for (uint32_t it = 0; it < pl.n; ++it, ++step) {
            adam_corrections(o, pl.adam_t0 + step * pl.P + (live ? p : 0) + 1);
            const uint32_t* ids = pl.item_ids;
            int Tn = 0;
            if (live) { const uint32_t sq = ord[it]; ids = pl.item_ids + pl.seq_start[sq]; Tn = (int)pl.seq_len[sq] - 1; }
            if (r == 0) tmax_s[tile] = 0;   // tile-wide number of lock-step timesteps
            tile_bar(tile);
            atomicMax(&tmax_s[tile], Tn);
            tile_bar(tile);
            const int Tmax = tmax_s[tile];

            // =========================== forward ===========================
            float h[32], c[32];
#pragma unroll
            for (int d = 0; d < 32; ++d) { h[d] = 0.0f; c[d] = 0.0f; }
            float loss_seq = 0.0f;
            // ids[t], ids[t+1], ids[t+2] travel in registers, loaded one step ahead of their use (ids has Tn + 1 entries)
            uint32_t idA = 0, idB = 0, idC = 0;
            if (Tn > 0) { idA = __ldg(ids); idB = __ldg(ids + 1); }
            if (Tn > 1) idC = __ldg(ids + 2);
            if (Tmax > 0) { gather_async(m, tb, idA, lane, SD1); }   // x_0
            cp_commit();
            for (int t = 0; t < Tmax; ++t) {
                const bool act = t < Tn;
                const uint32_t out = act ? idB : 0u;
                uint32_t idD = 0;
                if (t + 3 <= Tn) idD = __ldg(ids + t + 3);
                uint32_t cand[5]; float bc[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) cand[j] = j < tries ? draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range) : 0u;
                // ---- x_t has landed in SD1 (issued a step ago) ----
                tick(0);
                cp_wait<0>();
                __syncwarp();
                tick(1);
                {
                    float x[32];
                    slice_read_row(SD1, lane, x);
                    __syncwarp();
                    // target row and the first candidates of this step: G1
                    gather_async(m, tb, out, lane, SD0);
                    gather_async(m, tb, cand[0], lane, SD1);
                    if (tries > 1) gather_async(m, tb, cand[1], lane, SB);
                    cp_commit();
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        *reinterpret_cast<float4*>(Zt + tile_chunk_off(r, c4, 16)) =
                            make_float4(to_tf32(h[4 * c4]), to_tf32(h[4 * c4 + 1]), to_tf32(h[4 * c4 + 2]), to_tf32(h[4 * c4 + 3]));
                        *reinterpret_cast<float4*>(Zt + tile_chunk_off(r, 8 + c4, 16)) =
                            make_float4(to_tf32(x[4 * c4]), to_tf32(x[4 * c4 + 1]), to_tf32(x[4 * c4 + 2]), to_tf32(x[4 * c4 + 3]));
                    }
                    if (act) {
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            const float x8[8] = {x[8 * c8], x[8 * c8 + 1], x[8 * c8 + 2], x[8 * c8 + 3], x[8 * c8 + 4], x[8 * c8 + 5], x[8 * c8 + 6], x[8 * c8 + 7]};
                            *sb8(t, AX, c8) = pack_bf16x8(x8);
                        }
                    }
                }
                fence_async_smem();
                tc_fence_before_sync();
                tick(2);
                tile_bar(tile);
                tick(3);
                if (r == 0) {
                    tc_fence_after_sync();
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_tf32(tcol0, make_smem_desc(zt_a + k * 256, 128, 2048), make_smem_desc(wt_a + k * 256, 128, 2048), IDESC_G1, k > 0);
                    mma_commit(mbar + tile);
                }
                // bias scalars of the target and of every candidate, long before they are needed
                const float bp = act ? __ldcg(reinterpret_cast<const float*>(bias_rec(m, out))) : 0.0f;
#pragma unroll
                for (int j = 0; j < 5; ++j) bc[j] = (j < tries && act) ? __ldcg(reinterpret_cast<const float*>(bias_rec(m, cand[j]))) : 0.0f;
                tick(4);
                mbar_wait(mbar + tile, phase); phase ^= 1;
                tc_fence_after_sync();
                tick(5);
                // the tf32 Z tile is idle until the next step: candidates 2 and 3 go there (G2)
                if (tries > 2) { gather_async(m, tb, cand[2], lane, SZ0); gather_async(m, tb, cand[3], lane, SZ1); }
                cp_commit();
#pragma unroll
                for (int db = 0; db < 4; ++db) {
                    float pf[8], pi[8], pg[8], po[8], pc[8], ptc[8];
                    tmem_ld8x4(tbase + db * 8, tbase + 32 + db * 8, tbase + 64 + db * 8, tbase + 96 + db * 8, pf, pi, pg, po);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int d = db * 8 + j;
                        const float f = sigm<FASTM>(pf[j] + bias_s[d]);
                        const float ig = coupled ? 1.0f - f : sigm<FASTM>(pi[j] + bias_s[32 + d]);
                        const float gg = tnh<FASTM>(pg[j] + bias_s[64 + d]);
                        const float og = sigm<FASTM>(po[j] + bias_s[96 + d]);
                        const float cn = f * c[d] + ig * gg;
                        const float tcn = tnh<FASTM>(cn);
                        const float hn = og * tcn;
                        if (act) { c[d] = cn; h[d] = hn; } else h[d] = 0.0f;
                        pf[j] = f; pi[j] = ig; pg[j] = gg; po[j] = og; pc[j] = cn; ptc[j] = tcn;
                    }
                    if (act) {
                        *sb8(t, AF, db) = pack_bf16x8(pf); *sb8(t, AI, db) = pack_bf16x8(pi);
                        *sb8(t, AG, db) = pack_bf16x8(pg); *sb8(t, AO, db) = pack_bf16x8(po);
                        *sb8(t, AC, db) = pack_bf16x8(pc); *sb8(t, ATC, db) = pack_bf16x8(ptc);
                        const float h8[8] = {h[8 * db], h[8 * db + 1], h[8 * db + 2], h[8 * db + 3], h[8 * db + 4], h[8 * db + 5], h[8 * db + 6], h[8 * db + 7]};
                        *sb8(t, AHB, db) = pack_bf16x8(h8);
                        *sf4(t, 2 * db) = make_float4(h8[0], h8[1], h8[2], h8[3]);
                        *sf4(t, 2 * db + 1) = make_float4(h8[4], h8[5], h8[6], h8[7]);
                    }
                }
                tc_fence_before_sync();  // TMEM reads ordered before the next MMA (issued after the next tile barrier)
                tick(6);
                // scoring + negative sampling (sequence_model.rs:47-68, lstm.rs:300-320)
                cp_wait<1>();            // G1 (target, candidates 0 and 1) has landed; G2 may still fly
                __syncwarp();
                tick(7);
                float pos = bp;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 pv = slice_ld(SD0, lane, c4);
                    pos = fmaf(h[4 * c4], pv.x, pos); pos = fmaf(h[4 * c4 + 1], pv.y, pos);
                    pos = fmaf(h[4 * c4 + 2], pv.z, pos); pos = fmaf(h[4 * c4 + 3], pv.w, pos);
                }
                bool done = !act; uint32_t neg = 0; float ngs = 0.0f;
                float qv[32];
#pragma unroll
                for (int d = 0; d < 32; ++d) qv[d] = 0.0f;
                auto score = [&](const Slice& s, uint32_t cd, float bcj) {
                    if (!done) {
                        neg = cd;
                        slice_read_row(s, lane, qv);
                        float a = bcj;
#pragma unroll
                        for (int d = 0; d < 32; ++d) a = fmaf(h[d], qv[d], a);
                        ngs = a;
                        if (1.0f - pos + ngs > 0.0f) done = true;
                    }
                };
                tick(8);
                bool alld = __all_sync(kFull, done);
                if (!alld) { score(SD1, cand[0], bc[0]); __syncwarp(); }
                // x_{t+1} = E[ids[t+1]] takes the slice candidate 0 just left (G3)
                if (t + 1 < Tmax) gather_async(m, tb, (t + 1 < Tn) ? idB : 0u, lane, SD1);
                cp_commit();
                if (tries > 1) {
                    alld = __all_sync(kFull, done);
                    if (!alld) {
                        score(SB, cand[1], bc[1]);
                        __syncwarp();
                        gather_async(m, tb, cand[4], lane, SB);   // candidate 4 takes candidate 1's slice (G4)
                        cp_commit();
                        alld = __all_sync(kFull, done);
                    }
                    if (!alld) {
                        cp_wait<2>();   // G2 (candidates 2, 3) landed; G3, G4 may fly
                        __syncwarp();
                        score(SZ0, cand[2], bc[2]);
                        alld = __all_sync(kFull, done);
                    }
                    if (!alld) { score(SZ1, cand[3], bc[3]); alld = __all_sync(kFull, done); }
                    if (!alld) {
                        cp_wait<0>();
                        __syncwarp();
                        score(SB, cand[4], bc[4]);
                    }
                }
                tick(9);
                if (act) {
                    float l, g;
                    if (m.loss == 0) { const float s = sigm<FASTM>(ngs - pos); l = s; g = s * (1.0f - s); }
                    else { const float v = 1.0f + ngs - pos; l = v > 0.0f ? v : 0.0f; g = v > 0.0f ? 1.0f : 0.0f; }
                    loss_seq += l;
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const float4 pa = slice_ld(SD0, lane, 2 * c8), pb = slice_ld(SD0, lane, 2 * c8 + 1);
                        const float d8[8] = {g * (qv[8 * c8] - pa.x), g * (qv[8 * c8 + 1] - pa.y), g * (qv[8 * c8 + 2] - pa.z), g * (qv[8 * c8 + 3] - pa.w),
                                             g * (qv[8 * c8 + 4] - pb.x), g * (qv[8 * c8 + 5] - pb.y), g * (qv[8 * c8 + 6] - pb.z), g * (qv[8 * c8 + 7] - pb.w)};
                        *sb8(t, ADQ, c8) = pack_bf16x8(d8);
                    }
                    G_[(size_t)t * gstride] = g; NEG[(size_t)t * gstride] = neg;
                }
                __syncwarp();            // every lane is done with SD0 before the next step's target row overwrites it
                tick(10);
                if (TIMING && blockIdx.x == 0 && tid == 0 && it == 1 && (t == 10 || t == 11))
                    printf("FWD t=%d: wait_x %lld | Zt+scratch %lld | tile_bar %lld | mma_issue+bias_ld %lld | mma_wait %lld | gates %lld | wait_p %lld | pos %lld | tries %lld | loss+dq %lld | total %lld\n",
                           t, tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], tk[7] - tk[6], tk[8] - tk[7], tk[9] - tk[8], tk[10] - tk[9], tk[10] - tk[0]);
                idA = idB; idB = idC; idC = idD;
            }
            cp_wait<0>();
            __syncwarp();

            // =========================== backward ===========================
            // dz of timestep t+1 (dh_t in TMEM columns 0..31, dx_{t+1} in 32..63) is consumed straight from TMEM inside
            // timestep t's delta loop -- nothing but the cell-gradient recurrence lives in registers across timesteps.
            float dc_rec[32];
#pragma unroll
            for (int d = 0; d < 32; ++d) dc_rec[d] = 0.0f;
            float g_c = 0.0f; uint32_t neg_c = 0, out_c = 0;
            ActB cur;
            {
                const int t = Tmax - 1;
                const bool a0 = t >= 0 && t < Tn;
                if (a0) { g_c = G_[(size_t)t * gstride]; neg_c = NEG[(size_t)t * gstride]; out_c = __ldg(ids + t + 1); }
                load_act(cur, t > 0 ? t : 0, 0, a0);
            }
            bool prev_valid = false, prev_act = false;   // a dz of the previous (later) timestep is pending in TMEM
            // t = -1 is a pseudo-timestep: no deltas, no MMA, only the visit of the last deferred entry E[in_0] += step(dx_0)
            for (int t = Tmax - 1; t >= (Tmax > 0 ? -1 : 0); --t) {
                const bool act = t >= 0 && t < Tn;
                const float g = g_c; const uint32_t neg = neg_c, out = out_c;
                const bool actn = t >= 1 && (t - 1) < Tn;   // the next (earlier) timestep
                float g_n = 0.0f; uint32_t neg_n = 0, out_n = 0;
                if (actn) { g_n = G_[(size_t)(t - 1) * gstride]; neg_n = NEG[(size_t)(t - 1) * gstride]; out_n = __ldg(ids + t); }
                else if (t == 0 && Tn > 0) out_n = __ldg(ids);   // out_{-1} = ids[0] = in_0
                tick(12);
                if (t >= 1) prefetch_step(t - 1);
                if (prev_valid) { mbar_wait(mbar + tile, phase); phase ^= 1; tc_fence_after_sync(); }
                tick(13);
                if (t >= 0) {
                    stage_z_async(t, act);   // the previous MMA is done with the Z tile
#pragma unroll
                    for (int db = 0; db < 4; ++db) {
                        ActB nxt;
                        if (db < 3) load_act(nxt, t, db + 1, act);
                        float dhv[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) dhv[e] = 0.0f;
                        if (prev_valid) {   // dh_t and dx_{t+1}
                            uint32_t ra[8], rb[8];
                            tmem_ld8_issue(tbase + db * 8, ra); tmem_ld8_issue(tbase + 32 + db * 8, rb);
                            tmem_wait8(ra); tmem_wait8(rb);
#pragma unroll
                            for (int e = 0; e < 8; ++e) dhv[e] = prev_act ? __uint_as_float(ra[e]) : 0.0f;
                            slice_st(SZ1, lane, 2 * db, make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3])));
                            slice_st(SZ1, lane, 2 * db + 1, make_float4(__uint_as_float(rb[4]), __uint_as_float(rb[5]), __uint_as_float(rb[6]), __uint_as_float(rb[7])));
                        }
                        float df[8], di[8], dg[8], dO[8];
                        float f8[8], i8[8], g8[8], o8[8], q8[8], cp8[8], tc8[8];
                        unpack8(cur.f, f8); unpack8(cur.i, i8); unpack8(cur.g, g8); unpack8(cur.o, o8); unpack8(cur.q, q8);
                        unpack8(cur.cp, cp8); unpack8(cur.tc, tc8);
                        // gradient of the two rows that only need h_t goes straight to the staging slice
                        slice_st(SZ0, lane, 2 * db, make_float4(g * cur.h0.x, g * cur.h0.y, g * cur.h0.z, g * cur.h0.w));
                        slice_st(SZ0, lane, 2 * db + 1, make_float4(g * cur.h1.x, g * cur.h1.y, g * cur.h1.z, g * cur.h1.w));
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int d = db * 8 + e;
                            const float tcv = tc8[e];
                            const float dh = dhv[e] + q8[e];
                            const float d_o = dh * tcv;
                            const float dc = dc_rec[d] + dh * o8[e] * (1.0f - tcv * tcv);
                            float d_f = dc * cp8[e], d_i = dc * g8[e];
                            const float d_g = dc * i8[e];
                            dc_rec[d] = act ? dc * f8[e] : 0.0f;
                            if (coupled) { d_f -= d_i; d_i = 0.0f; }
                            df[e] = d_f * f8[e] * (1.0f - f8[e]);
                            di[e] = coupled ? 0.0f : d_i * i8[e] * (1.0f - i8[e]);
                            dg[e] = d_g * (1.0f - g8[e] * g8[e]);
                            dO[e] = d_o * o8[e] * (1.0f - o8[e]);
                        }
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 0 + db, 16)) = pack_bf16x8(df);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 4 + db, 16)) = pack_bf16x8(di);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 8 + db, 16)) = pack_bf16x8(dg);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 12 + db, 16)) = pack_bf16x8(dO);
                        if (db < 3) cur = nxt;
                    }
                    tick(14);
                    cp_wait<0>();             // Z_t rows have landed
                    fence_async_smem();
                    tc_fence_before_sync();   // also orders this thread's TMEM reads of dz_{t+1} before the MMA that overwrites them
                    tick(15);
                    tile_bar(tile);
                    tick(16);
                    if (r == 0) {
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
                    // first activation block of the next timestep: requested before this timestep's visits
                    if (t >= 1) load_act(cur, t - 1, 0, actn);
                } else {   // t == -1: dx_0 out of TMEM, nothing else
#pragma unroll
                    for (int db = 0; db < 4; ++db) {
                        uint32_t rb[8];
                        tmem_ld8_issue(tbase + 32 + db * 8, rb);
                        tmem_wait8(rb);
                        slice_st(SZ1, lane, 2 * db, make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3])));
                        slice_st(SZ1, lane, 2 * db + 1, make_float4(__uint_as_float(rb[4]), __uint_as_float(rb[5]), __uint_as_float(rb[6]), __uint_as_float(rb[7])));
                    }
                    tc_fence_before_sync();
                }
                __syncwarp();            // the g h_t and dx_{t+1} slices are complete
                tick(17);
                // ---- sparse visits of this timestep (overlap the MMAs): E[neg_t]; E[out_t] with the deferred E[in_{t+1}] ----
                {
                    const bool triple = act && neg == out;
                    const bool has_dx = t + 1 < Tn;   // a deferred E[in_{t+1}] entry exists (t + 1 >= 0 always)
                    const uint32_t fl = (act && !triple ? 1u : 0u) | ((act || has_dx) ? 2u : 0u) | (has_dx ? 4u : 0u) | (triple ? 8u : 0u) | (act ? 16u : 0u);
                    float4* rn = bias_rec(m, neg); float4* ro = bias_rec(m, out);
                    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
                    const bool bv = act;
                    const bool batom = !o.adam && !(pl.dbg_flags & 8) && !m.hbm_resident;
                    float* fn_ = reinterpret_cast<float*>(rn); float* fo_ = reinterpret_cast<float*>(ro);
                    if (bv) {
                        if (batom) {   // {ld b, atom G_b += g^2}: b[neg] takes +g; b[out] takes -g (two entries on one record when neg == out)
                            ba.x = __ldcg(fn_); ba.y = atomicAdd(fn_ + 1, neg != out ? g * g : 2.0f * g * g);
                            if (neg != out) { bb.x = __ldcg(fo_); bb.y = atomicAdd(fo_ + 1, g * g); }
                        } else { ba = __ldcg(rn); if (neg != out) bb = __ldcg(ro); }
                    }
                    coop_visits(m, tb, neg, out, fl, lane, SZ0, SZ1, o, (pl.dbg_flags & 8) != 0 || m.hbm_resident != 0);
                    if (bv) {   // b[neg] += step(+g), b[out] += step(-g)
                        const float4 a0 = ba, b0 = bb;
                        if (neg != out) {
                            if (!o.adam) { adagrad1(ba.x, ba.y, g, o.lr, o.l2); adagrad1(bb.x, bb.y, -g, o.lr, o.l2); }
                            else { adam1(ba.x, ba.y, ba.z, g, o); adam1(bb.x, bb.y, bb.z, -g, o); }
                            if (batom) {
                                atomicAdd(fn_, ba.x - a0.x); atomicAdd(fo_, bb.x - b0.x);
                                if (o.l2 != 0.0f) { atomicAdd(fn_ + 1, ba.y - a0.y - g * g); atomicAdd(fo_ + 1, bb.y - b0.y - g * g); }
                            } else { __stcg(rn, ba); __stcg(ro, bb); }
                        } else {
                            if (!o.adam) { adagrad1(ba.x, ba.y, g, o.lr, o.l2); adagrad1(ba.x, ba.y, -g, o.lr, o.l2); }
                            else { adam1(ba.x, ba.y, ba.z, g, o); adam1(ba.x, ba.y, ba.z, -g, o); }
                            if (batom) {
                                atomicAdd(fn_, ba.x - a0.x);
                                if (o.l2 != 0.0f) atomicAdd(fn_ + 1, ba.y - a0.y - 2.0f * g * g);
                            } else __stcg(rn, ba);
                        }
                    }
                }
                tick(18);
                if (TIMING && blockIdx.x == 0 && tid == 0 && it == 1 && (t == 10 || t == 11))
                    printf("BWD t=%d: scalars+prefetch+mma_wait %lld | deltas(4 blocks) %lld | wait_z %lld | tile_bar %lld | mma_issue+preload %lld | visits %lld | total %lld\n",
                           t, tk[13] - tk[12], tk[14] - tk[13], tk[15] - tk[14], tk[16] - tk[15], tk[17] - tk[16], tk[18] - tk[17], tk[18] - tk[12]);
                prev_valid = true; prev_act = act;
                g_c = g_n; neg_c = neg_n; out_c = out_n;
            }
            if (live) { loss_acc += loss_seq; ex += (unsigned long long)Tn; }

            // =========================== dense step on the CTA-summed gradient ===========================
            // thread (tile, r) holds row gd = r of its tile's dW^T: columns 0..63 = dW[k][gd], column 64 = dbias[gd]
            __syncthreads();
            float* xch = reinterpret_cast<float*>(smem + OFF_TILES + TILE_BYTES + TILE_DB);  // tile 1's delta area: [65][128]
            float dwr[65];
            {
                const bool have = Tmax > 0;  // a tile whose partitions are all dead issued no MMA this round
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) {
                    float v8[8];
                    tmem_ld8(tbase + 128 + cb * 8, v8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dwr[cb * 8 + j] = have ? v8[j] : 0.0f;
                }
                float v8[8];
                tmem_ld8(tbase + 128 + 64, v8);
                dwr[64] = have ? v8[0] : 0.0f;
            }
            tc_fence_before_sync();
            if (NT == 2) {
                if (tile == 1) {
#pragma unroll
                    for (int k = 0; k < 65; ++k) xch[k * 128 + r] = dwr[k];
                }
                __syncthreads();
                if (tile == 0) {
#pragma unroll
                    for (int k = 0; k < 65; ++k) dwr[k] += xch[k * 128 + r];
                }
            }
            if (tile == 0) {
                OptCfg od = o;
                adam_corrections(od, pl.adam_t0 + step * pl.P + (uint64_t)blockIdx.x * NT * 128 + 1);
                const int gd = r;
                __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(Wb);
#pragma unroll
                for (int k = 0; k < 65; ++k) {
                    const size_t idx = (size_t)k * kNG + gd;  // k == 64: bias[gd]
                    float w = __ldcg(m.dense + idx), s1 = __ldcg(m.dense + nd + idx);
                    if (od.adam) {
                        float s2 = __ldcg(m.dense + 2 * nd + idx);
                        adam1(w, s1, s2, dwr[k], od);
                        __stcg(m.dense + 2 * nd + idx, s2);
                    } else adagrad1(w, s1, dwr[k], od.lr, od.l2);
                    __stcg(m.dense + idx, w); __stcg(m.dense + nd + idx, s1);
                    if (k < 64) {
                        reinterpret_cast<float*>(Wt + tile_chunk_off(gd, k >> 2, 16))[k & 3] = to_tf32(w);
                        wb[(tile_chunk_off(k, gd >> 3, 16) >> 1) + (gd & 7)] = __float2bfloat16_rn(w);
                    } else bias_s[gd] = w;
                }
            }
            fence_async_smem();
            __syncthreads();
        }
    }
    if (live) {
        pl.rng[p] = rng; pl.step_ctr[p] = step;
        pl.loss_acc[p] += loss_acc; pl.examples[p] += ex;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc<(NT == 1 ? 256 : 512)>(*tmem_ptr);
}

template <int NT, bool FASTM, bool TIMING = false>
cudaError_t launch_one(const ModelDev& m, const PlanDev& p, cudaStream_t st) {
    const size_t smem = OFF_TILES + (size_t)NT * TILE_BYTES;
    const int per_cta = 128 * NT;
    dim3 grid((p.P + per_cta - 1) / per_cta);
    cudaError_t e = cudaFuncSetAttribute(lstm_tc2_train_kernel<NT, FASTM, TIMING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lstm_tc2_train_kernel<NT, FASTM, TIMING><<<grid, per_cta, smem, st>>>(m, p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_lstm_tc2(const ModelDev& m, const PlanDev& p, int nt, bool fast_math, cudaStream_t st) {
    if (nt == 2 && (p.dbg_flags & 16)) return launch_one<2, false, true>(m, p, st);
    if (nt == 2) return fast_math ? launch_one<2, true>(m, p, st) : launch_one<2, false>(m, p, st);
    return fast_math ? launch_one<1, true>(m, p, st) : launch_one<1, false>(m, p, st);
}

}  // namespace sbr
