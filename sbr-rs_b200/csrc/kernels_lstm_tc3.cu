// kernels_lstm_tc3.cu -- tensor-core LSTM training kernel (D = 32), third generation: DS threads per sequence.
//
// The first two tile kernels give one THREAD a whole sequence (32 hidden units): 255 registers, 8 warps per SM,
// ~6,300 straight-line instructions per warp and timestep.  Their profile (profiles/r1_v5_*, r1_v6_*) shows 11-14 %
// issue utilisation no matter how well the memory round trips are hidden: with two warps per scheduler the kernel is
// bound by the per-warp dependent-issue latency (fixed-latency waits, MIO / MUFU scoreboards, instruction-cache
// misses of the huge unrolled body), not by HBM or L2.
//
// Here a sequence is owned by DS threads in DS different warps of the same lane quarter (TMEM lane == sequence, and
// warps w, w+4, w+8, .. may all read lanes 32 (w % 4) ..): thread `part` owns hidden units [part * 32/DS, ..) of
// everything -- h, c, gate columns, deltas, the 16-byte chunks of every gathered / updated item row.  16 (DS = 2) or
// 32 (DS = 4) warps per SM, a quarter of the registers and of the unrolled code per thread.  The only data that
// must cross between the DS threads of a sequence are the partial dot products of the scores (h . p, h . q): they
// go through a 2-slot shared-memory exchange and a named barrier of the DS warps; all threads add the partials
// in the same order, so accept / reject decisions are bit-identical in every owner of a sequence.
//
// Everything else is the second generation's pipeline (kernels_lstm_tc2.cu): tcgen05 MMAs from shared-memory tiles
// with TMEM accumulators (tf32 forward, bf16 backward), cp.async row prefetch into staging slices that alias idle
// operand tiles, merged sparse visits in the reference order, bf16 activation copies, timestep-major scratch.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

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
constexpr uint32_t XS_BYTES_PER_QUAD = 2 * 4 * 32 * 4;   // score exchange: 2 slots x (up to 4 parts) x 32 lanes floats

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / (1 + 2^(-x log2 e)); saturates cleanly: ex2 -> +inf gives rcp -> 0, ex2 -> 0 gives 1
__device__ __forceinline__ float sigm(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tnh(float x) { return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f); }

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
__device__ __forceinline__ void apply4(float4& w, float4& s, float4& v, const float4& g, float sign, const OptCfg& o) {
    if (!o.adam) {
        adagrad1(w.x, s.x, sign * g.x, o.lr, o.l2); adagrad1(w.y, s.y, sign * g.y, o.lr, o.l2);
        adagrad1(w.z, s.z, sign * g.z, o.lr, o.l2); adagrad1(w.w, s.w, sign * g.w, o.lr, o.l2);
    } else {
        adam1(w.x, s.x, v.x, sign * g.x, o); adam1(w.y, s.y, v.y, sign * g.y, o);
        adam1(w.z, s.z, v.z, sign * g.z, o); adam1(w.w, s.w, v.w, sign * g.w, o);
    }
}

// L2 atomics for the Hogwild Adagrad visits (see kernels_lstm_tc2.cu): {ld w, atom G += sum g^2} in one round trip, the
// sequential applications in registers from the returned G, then fire-and-forget reductions on w and G.
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
// word `pr` (bf16 elements 2 pr, 2 pr + 1) of a packed block; the two halves as floats
__device__ __forceinline__ uint32_t word_of(const uint4& u, int pr) { return pr == 0 ? u.x : pr == 1 ? u.y : pr == 2 ? u.z : u.w; }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { uint32_t o; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(hi), "f"(lo)); return o; }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A staging slice holds this thread-group's part of 32 rows: row r, 16-byte chunk c at (r >> 3) * gs + c * 128 + (r & 7) * 16
// (core-matrix order: lane-per-row accesses and the cooperative 8-rows-per-chunk accesses are both conflict-free).
struct Slice { uint8_t* p; uint32_t gs; };

// item_rec(): one multiply-add on an unsharded table (FLAT); shard base pointers from shared memory otherwise
struct Table { float* e0; float* const* es; uint32_t stride, gmask; int gshift; };
template <bool FLAT>
__device__ __forceinline__ float* trec(const Table& tb, uint32_t id) {
    if (FLAT) return tb.e0 + (size_t)id * tb.stride + 4;
    return tb.es[id & tb.gmask] + (size_t)(id >> tb.gshift) * tb.stride + 4;
}

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

template <int NT, int DS, bool FLAT>
__global__ void __launch_bounds__(128 * NT * DS, 1) lstm_tc3_train_kernel(ModelDev m, PlanDev pl) {
    constexpr int DPT = 32 / DS;       // hidden units per thread
    constexpr int NCH = DPT / 4;       // 16-byte fp32 chunks of a row per thread
    constexpr int NB8 = DPT / 8;       // 8-unit blocks per thread
    constexpr int RPI = 32 / NCH;      // rows covered by one cooperative warp instruction (lane -> row l % RPI, chunk l / RPI)
    constexpr int NGI = 32 / RPI;      // cooperative instructions per 32 rows
    constexpr int TT = 128 * DS;       // threads per tile
    constexpr uint32_t SL = 4096 / DS; // bytes of a compact slice
    constexpr uint32_t GS = NCH * 128; // its row-group stride

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Wt = smem + OFF_WT;
    uint8_t* Wb = smem + OFF_WB;
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_MISC);           // [NT]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 32);
    int* tmax_s = reinterpret_cast<int*>(smem + OFF_MISC + 40);              // [NT]

    const int tid = threadIdx.x, tile = tid / TT, tt = tid % TT;
    const int wi = tt >> 5, q = wi & 3, part = wi >> 2, lane = tid & 31;
    const int r = q * 32 + lane;                     // sequence row of the tile == TMEM lane
    const int gb0 = part * NB8;                      // first 8-unit block owned by this thread
    uint8_t* Zt = smem + OFF_TILES + tile * TILE_BYTES + TILE_ZT;
    uint8_t* Db = smem + OFF_TILES + tile * TILE_BYTES + TILE_DB;
    uint8_t* Zb = smem + OFF_TILES + tile * TILE_BYTES + TILE_ZB;
    float* xs = reinterpret_cast<float*>(smem + OFF_TILES + NT * TILE_BYTES + (tile * 4 + q) * XS_BYTES_PER_QUAD);
    // staging slices: this warp's share of the rows its quad owns in the operand tiles
    const Slice SD0{Db + q * 8192 + part * 2 * SL, GS}, SD1{Db + q * 8192 + part * 2 * SL + SL, GS};   // forward: delta tile idle
    const Slice SB{Zb + q * 5120 + part * NCH * 128, 1280u};                                            // forward: bf16 Z tile idle
    const Slice SZ0{Zt + q * 8192 + part * 2 * SL, GS}, SZ1{Zt + q * 8192 + part * 2 * SL + SL, GS};   // tf32 Z tile: after its MMA / all of backward
    const uint32_t tile_gid = blockIdx.x * NT + tile;
    const uint32_t p = tile_gid * 128u + r;
    const bool live = p < pl.P;
    const bool lead = part == 0;                     // the owner that does per-sequence scalar work
    const size_t nd = m.ndense;
    const bool coupled = m.variant == 1;
    const int T = m.T;
    float** es_s = reinterpret_cast<float**>(smem + OFF_MISC + 64);         // [8] shard base pointers of the item table
    if (tid < 8) es_s[tid] = m.Es[tid];
    Table tb; tb.e0 = m.Es[0]; tb.es = es_s; tb.stride = (uint32_t)rec_floats(m); tb.gmask = m.gmask; tb.gshift = m.gshift;

    auto tile_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(tile + 1), "n"(TT) : "memory"); };
    auto quad_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(3 + tile * 4 + q), "n"(32 * DS) : "memory"); };
    int xslot = 0;
    // sum of the DS owners' partial values, identical (same order) in every owner; warp-uniform call sites only
    auto xsum = [&](float partial) -> float {
        xs[(xslot * 4 + part) * 32 + lane] = partial;
        quad_bar();
        float s = xs[(xslot * 4) * 32 + lane];
#pragma unroll
        for (int k = 1; k < DS; ++k) s += xs[(xslot * 4 + k) * 32 + lane];
        xslot ^= 1;
        return s;
    };
    auto slice_ld = [&](const Slice& s, int c) -> float4 {
        return *reinterpret_cast<const float4*>(s.p + (lane >> 3) * s.gs + c * 128 + (lane & 7) * 16);
    };
    auto slice_st = [&](const Slice& s, int c, const float4& v) {
        *reinterpret_cast<float4*>(s.p + (lane >> 3) * s.gs + c * 128 + (lane & 7) * 16) = v;
    };
    // this warp's part of the rows named by the lanes' ids -> slice, asynchronously
    auto gather_async = [&](uint32_t my_id, const Slice& s) {
        const int rl = lane % RPI, ch = lane / RPI;
        const uint32_t base = smem_u32(s.p) + ch * 128;
        const float* src[NGI];
#pragma unroll
        for (int i = 0; i < NGI; ++i)   // all addresses first: the copies below are compiler barriers
            src[i] = trec<FLAT>(tb, __shfl_sync(kFull, my_id, i * RPI + rl)) + part * DPT + ch * 4;
#pragma unroll
        for (int i = 0; i < NGI; ++i) {
            const int row = i * RPI + rl;
            cp_async16(base + (row >> 3) * s.gs + (row & 7) * 16, src[i]);
        }
    };

    // tile scratch, timestep-major across the grid's tiles: [T][tiles][H fp32 8 units | F I G O X DQ C TANH(C) H bf16 9 x 4 units]
    // of [128 seq] 16-byte pieces (88 KB per tile-timestep), then G, NEG [T][tiles][128]  (see kernels_lstm_tc2.cu)
    const size_t ntiles = (size_t)gridDim.x * NT;
    constexpr int kStepU4 = (8 + 36) * 128;
    uint4* sbase = reinterpret_cast<uint4*>(pl.scratch);
    float* G_ = reinterpret_cast<float*>(sbase + (size_t)T * ntiles * kStepU4) + (size_t)tile_gid * 128 + r;   // + t * gstride
    const size_t gstride = ntiles * 128;
#define NEG (reinterpret_cast<uint32_t*>(G_ + (size_t)T * gstride))
    enum { AF = 0, AI = 1, AG = 2, AO = 3, AX = 4, ADQ = 5, AC = 6, ATC = 7, AHB = 8 };
    auto step_base = [&](int t) -> uint4* { return sbase + ((size_t)t * ntiles + tile_gid) * kStepU4; };
    auto sf4 = [&](int t, int c4) -> float4* { return reinterpret_cast<float4*>(step_base(t) + (size_t)c4 * 128 + r); };
    auto sb8 = [&](int t, int which, int c8) -> uint4* { return step_base(t) + (size_t)(8 + which * 4 + c8) * 128 + r; };
    auto prefetch_step = [&](int t) {   // 88 KB = 704 lines towards L2, spread over the tile's threads
        constexpr int LPT = (704 + TT - 1) / TT;
        const char* base = reinterpret_cast<const char*>(step_base(t));
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            const int line = tt * LPT + i;
            if (line < 704) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)line * 128));
        }
    };
    auto load_act = [&](ActB& a, int t, int gb, bool on) {
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u); const float4 zf = make_float4(0.f, 0.f, 0.f, 0.f);
        a.f = a.i = a.g = a.o = a.q = a.cp = a.tc = z4; a.h0 = a.h1 = zf;
        if (on) {
            a.f = __ldcg(sb8(t, AF, gb)); a.i = __ldcg(sb8(t, AI, gb)); a.g = __ldcg(sb8(t, AG, gb)); a.o = __ldcg(sb8(t, AO, gb));
            a.q = __ldcg(sb8(t, ADQ, gb)); a.tc = __ldcg(sb8(t, ATC, gb));
            a.h0 = __ldcg(sf4(t, 2 * gb)); a.h1 = __ldcg(sf4(t, 2 * gb + 1));
            if (t > 0) a.cp = __ldcg(sb8(t - 1, AC, gb));
        }
    };
    // Z_t = [h_{t-1}, x_t] (bf16) straight from the scratch into this thread's row of the bf16 Z tile; rows of finished
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
    // The sparse visits of one backward timestep, this warp's chunks of the 32 rows (flags as in kernels_lstm_tc2.cu):
    //   bit 0: E[neg] += step(+g h)      bit 1: visit E[out]: [bit 2: step(dx_{t+1})] [bit 3: step(+g h)] [bit 4: step(-g h)]
    const bool noatom = m.hbm_resident != 0;   // hot rows only exist on L2-resident tables
    // A pass covers GPP row groups (NGI / GPP passes per timestep).  It is split in two so that the round trip of the
    // first pass -- {ld w, atom G} or {ld w, ld G} -- flies while the tile meets at its barrier and the MMAs are issued.
    constexpr int GPP = NGI >= 2 ? 2 : 1;   // row groups per pass: 4 row visits (w, s each) in flight
    struct VisitState { float* rn[GPP]; float* ro[GPP]; uint32_t f[GPP]; uint32_t off[GPP]; float4 wn[GPP], sn[GPP], wo[GPP], so[GPP]; };
    auto atom_amount_out = [&](uint32_t f, const float4& s4, const Slice& dx, uint32_t off) -> float4 {   // what the atom adds to G[out]
        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f & 4u) t4 = sq4(*reinterpret_cast<const float4*>(dx.p + off));
        if (f & 8u) t4 = add4(t4, s4);
        if (f & 16u) t4 = add4(t4, s4);
        return t4;
    };
    auto visit_issue = [&](int pass, VisitState& V, uint32_t neg, uint32_t out, uint32_t fl, const Slice& gh, const Slice& dx, const OptCfg& o) {
        const int rl = lane % RPI, ch = lane / RPI;
        const bool atomics = !o.adam && !noatom;
#pragma unroll
        for (int gg = 0; gg < GPP; ++gg) {
            const int row = (pass * GPP + gg) * RPI + rl;
            const uint32_t idn = __shfl_sync(kFull, neg, row), ido = __shfl_sync(kFull, out, row);
            V.f[gg] = __shfl_sync(kFull, fl, row);
            V.rn[gg] = trec<FLAT>(tb, idn) + part * DPT + ch * 4; V.ro[gg] = trec<FLAT>(tb, ido) + part * DPT + ch * 4;
            V.off[gg] = (uint32_t)(row >> 3) * GS + (uint32_t)ch * 128u + (uint32_t)(row & 7) * 16u;
        }
#pragma unroll
        for (int gg = 0; gg < GPP; ++gg) {
            if (atomics) {
                const float4 s4 = sq4(*reinterpret_cast<const float4*>(gh.p + V.off[gg]));
                if (V.f[gg] & 1u) { V.wn[gg] = __ldcg(reinterpret_cast<const float4*>(V.rn[gg])); V.sn[gg] = atom_add4(V.rn[gg] + kD, s4); }
                if (V.f[gg] & 2u) { V.wo[gg] = __ldcg(reinterpret_cast<const float4*>(V.ro[gg])); V.so[gg] = atom_add4(V.ro[gg] + kD, atom_amount_out(V.f[gg], s4, dx, V.off[gg])); }
            } else {
                if (V.f[gg] & 1u) { V.wn[gg] = __ldcg(reinterpret_cast<const float4*>(V.rn[gg])); V.sn[gg] = __ldcg(reinterpret_cast<const float4*>(V.rn[gg] + kD)); }
                if (V.f[gg] & 2u) { V.wo[gg] = __ldcg(reinterpret_cast<const float4*>(V.ro[gg])); V.so[gg] = __ldcg(reinterpret_cast<const float4*>(V.ro[gg] + kD)); }
            }
        }
    };
    // flags: bit 0: E[neg] += step(+g h)      bit 1: visit E[out]: [bit 2: step(dx_{t+1})] [bit 3: step(+g h)] [bit 4: step(-g h)]
    auto visit_finish = [&](VisitState& V, const Slice& gh, const Slice& dx, const OptCfg& o) {
        const bool atomics = !o.adam && !noatom;
#pragma unroll
        for (int gg = 0; gg < GPP; ++gg) {
            const float4 g4 = *reinterpret_cast<const float4*>(gh.p + V.off[gg]);
            const float4 d4 = *reinterpret_cast<const float4*>(dx.p + V.off[gg]);
            if (V.f[gg] & 1u) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o.adam) v = __ldcg(reinterpret_cast<const float4*>(V.rn[gg] + 2 * kD));
                const float4 w0 = V.wn[gg], G0 = V.sn[gg];
                apply4(V.wn[gg], V.sn[gg], v, g4, 1.0f, o);
                if (atomics) {
                    red_add4(V.rn[gg], sub4(V.wn[gg], w0));
                    if (o.l2 != 0.0f) red_add4(V.rn[gg] + kD, sub4(sub4(V.sn[gg], G0), sq4(g4)));
                } else {
                    __stcg(reinterpret_cast<float4*>(V.rn[gg]), V.wn[gg]); __stcg(reinterpret_cast<float4*>(V.rn[gg] + kD), V.sn[gg]);
                    if (o.adam) __stcg(reinterpret_cast<float4*>(V.rn[gg] + 2 * kD), v);
                }
            }
            if (V.f[gg] & 2u) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o.adam) v = __ldcg(reinterpret_cast<const float4*>(V.ro[gg] + 2 * kD));
                const float4 w0 = V.wo[gg], G0 = V.so[gg];
                if (V.f[gg] & 4u) apply4(V.wo[gg], V.so[gg], v, d4, 1.0f, o);
                if (V.f[gg] & 8u) apply4(V.wo[gg], V.so[gg], v, g4, 1.0f, o);
                if (V.f[gg] & 16u) apply4(V.wo[gg], V.so[gg], v, g4, -1.0f, o);
                if (atomics) {
                    red_add4(V.ro[gg], sub4(V.wo[gg], w0));
                    if (o.l2 != 0.0f) {
                        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 q4 = sq4(g4);
                        if (V.f[gg] & 4u) t4 = sq4(d4);
                        if (V.f[gg] & 8u) t4 = add4(t4, q4);
                        if (V.f[gg] & 16u) t4 = add4(t4, q4);
                        red_add4(V.ro[gg] + kD, sub4(sub4(V.so[gg], G0), t4));
                    }
                } else {
                    __stcg(reinterpret_cast<float4*>(V.ro[gg]), V.wo[gg]); __stcg(reinterpret_cast<float4*>(V.ro[gg] + kD), V.so[gg]);
                    if (o.adam) __stcg(reinterpret_cast<float4*>(V.ro[gg] + 2 * kD), v);
                }
            }
        }
    };

    // ---- one-time setup ----
    if (tid < 32) tmem_alloc<(NT == 1 ? 256 : 512)>(tmem_ptr);
    if (tid == 0) { for (int i = 0; i < NT; ++i) mbar_init(mbar + i, 1); fence_mbar_init(); }
    if (lead) {   // constant columns of the bf16 Z tile: col 64 = 1 (bias gradient), 65..79 = 0
        const float one8[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 8, 10)) = pack_bf16x8(one8);
        *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 9, 10)) = pack_bf16x8(zero8);
    }
    if (tile == 0) {  // weights: thread (gd, part) stages its share of column gd of W (both operand tiles); bias[gd]
        const int gd = r;
        __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(Wb);
        for (int c4 = part * (16 / DS); c4 < (part + 1) * (16 / DS); ++c4) {
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                wv[j] = __ldcg(m.dense + (size_t)(4 * c4 + j) * kNG + gd);
                wb[(tile_chunk_off(4 * c4 + j, gd >> 3, 16) >> 1) + (gd & 7)] = __float2bfloat16_rn(wv[j]);
            }
            *reinterpret_cast<float4*>(Wt + tile_chunk_off(gd, c4, 16)) =
                make_float4(to_tf32(wv[0]), to_tf32(wv[1]), to_tf32(wv[2]), to_tf32(wv[3]));
        }
        if (lead) bias_s[gd] = __ldcg(m.dense + (size_t)kNK * kNG + gd);
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_ptr + (uint32_t)tile * 256u + ((uint32_t)(q * 32) << 16);   // this thread's TMEM lane, tile's columns
    const uint32_t tcol0 = *tmem_ptr + (uint32_t)tile * 256u;                                 // for the MMA issuer
    const uint32_t tcs = tbase + 208u + (uint32_t)(part * DPT);                               // this thread's cell-state columns
    const uint32_t zt_a = smem_u32(Zt), db_a = smem_u32(Db), zb_a = smem_u32(Zb), wt_a = smem_u32(Wt), wb_a = smem_u32(Wb);
    constexpr uint32_t IDESC_G1 = make_idesc_tf32(128, 128, 0, 0);
    constexpr uint32_t IDESC_G2 = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t IDESC_G3 = make_idesc_bf16(128, 80, 1, 1);
    const bool issuer = tt == 0;
    uint32_t phase = 0;

    uint64_t key = 0; uint32_t* ord = nullptr;
    uint64_t step = pl.step_ctr[live ? p : 0];
    if (live) { key = pl.keys[p]; ord = pl.order + (size_t)p * pl.n; }
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;
    const int tries = m.loss == 2 ? 5 : 1;

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
        quad_bar();   // the other owners read the shuffled order (same SM: visible after the barrier)
        for (uint32_t it = 0; it < pl.n; ++it, ++step) {
            adam_corrections(o, pl.adam_t0 + step * pl.P + (live ? p : 0) + 1);
            const uint32_t* ids = pl.item_ids;
            int Tn = 0;
            if (live) { const uint32_t sq = __ldcg(ord + it); ids = pl.item_ids + pl.seq_start[sq]; Tn = (int)pl.seq_len[sq] - 1; }
            if (tt == 0) tmax_s[tile] = 0;   // tile-wide number of lock-step timesteps
            tile_bar();
            if (lead) atomicMax(&tmax_s[tile], Tn);
            tile_bar();
            const int Tmax = tmax_s[tile];

            // =========================== forward ===========================
            // The cell state (forward) and the cell-gradient recurrence (backward) are touched once per timestep, inside
            // the gate / delta loops: they live in 32 spare TMEM columns of the tile (208..239), not in registers.
            float h[DPT];
#pragma unroll
            for (int d = 0; d < DPT; ++d) h[d] = 0.0f;
            {
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int b = 0; b < NB8; ++b) tmem_st8(tcs + b * 8, z8);
                tmem_st_wait();
            }
            float loss_seq = 0.0f;
            // ids[t+1], ids[t+2] travel in registers, loaded one step ahead of their use (ids has Tn + 1 entries)
            uint32_t idB = 0, idC = 0;
            {
                uint32_t idA = 0;
                if (Tn > 0) { idA = __ldg(ids); idB = __ldg(ids + 1); }
                if (Tn > 1) idC = __ldg(ids + 2);
                if (Tmax > 0) gather_async(idA, SD1);   // x_0
                cp_commit();
            }
            for (int t = 0; t < Tmax; ++t) {
                const bool act = t < Tn;
                const uint32_t out = act ? idB : 0u;
                uint32_t idD = 0;
                if (t + 3 <= Tn) idD = __ldg(ids + t + 3);
                float bc[5];
                auto cand = [&](int j) -> uint32_t { return draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range); };
                // ---- x_t has landed in SD1 (issued a step ago) ----
                cp_wait<0>();
                __syncwarp();
                // The staging slices inside the tf32 Z tile cover cells that the OTHER owners of the quad are about to
                // write as operands: nobody writes Z_t before every owner's copies into that region have landed.
                quad_bar();
                {
                    float4 xv[NCH];
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) xv[cc] = slice_ld(SD1, cc);
                    __syncwarp();
                    // target row and the first candidates of this step: G1
                    gather_async(out, SD0);
                    gather_async(cand(0), SD1);
                    if (tries > 1) gather_async(cand(1), SB);
                    cp_commit();
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        *reinterpret_cast<float4*>(Zt + tile_chunk_off(r, part * NCH + cc, 16)) =
                            make_float4(to_tf32(h[4 * cc]), to_tf32(h[4 * cc + 1]), to_tf32(h[4 * cc + 2]), to_tf32(h[4 * cc + 3]));
                        *reinterpret_cast<float4*>(Zt + tile_chunk_off(r, 8 + part * NCH + cc, 16)) =
                            make_float4(to_tf32(xv[cc].x), to_tf32(xv[cc].y), to_tf32(xv[cc].z), to_tf32(xv[cc].w));
                    }
                    if (act) {
#pragma unroll
                        for (int b = 0; b < NB8; ++b) {
                            const float x8[8] = {xv[2 * b].x, xv[2 * b].y, xv[2 * b].z, xv[2 * b].w, xv[2 * b + 1].x, xv[2 * b + 1].y, xv[2 * b + 1].z, xv[2 * b + 1].w};
                            *sb8(t, AX, gb0 + b) = pack_bf16x8(x8);
                        }
                    }
                }
                fence_async_smem();
                tc_fence_before_sync();
                tile_bar();
                if (issuer) {
                    tc_fence_after_sync();
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_tf32(tcol0, make_smem_desc(zt_a + k * 256, 128, 2048), make_smem_desc(wt_a + k * 256, 128, 2048), IDESC_G1, k > 0);
                    mma_commit(mbar + tile);
                }
                // bias scalars of the target and of every candidate (the lead owner folds them into its partial sums)
                float bp = 0.0f;
#pragma unroll
                for (int j = 0; j < 5; ++j) bc[j] = 0.0f;
                if (lead && act) {
                    bp = __ldcg(reinterpret_cast<const float*>(bias_rec(m, out)));
#pragma unroll
                    for (int j = 0; j < 5; ++j) if (j < tries) bc[j] = __ldcg(reinterpret_cast<const float*>(bias_rec(m, cand(j))));
                }
                mbar_wait(mbar + tile, phase); phase ^= 1;
                tc_fence_after_sync();
                // the tf32 Z tile is idle until the next step: candidates 2 and 3 go there (G2)
                if (tries > 2) { gather_async(cand(2), SZ0); gather_async(cand(3), SZ1); }
                cp_commit();
#pragma unroll
                for (int b = 0; b < NB8; ++b) {
                    const int gb = gb0 + b;
                    float pf[8], pi[8], pg[8], po[8], pc[8], ptc[8];
                    tmem_ld8(tcs + b * 8, pc);   // c_{t-1}
                    tmem_ld8x4(tbase + gb * 8, tbase + 32 + gb * 8, tbase + 64 + gb * 8, tbase + 96 + gb * 8, pf, pi, pg, po);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int d = b * 8 + j, dg_ = gb * 8 + j;
                        const float f = sigm(pf[j] + bias_s[dg_]);
                        const float ig = coupled ? 1.0f - f : sigm(pi[j] + bias_s[32 + dg_]);
                        const float gg = tnh(pg[j] + bias_s[64 + dg_]);
                        const float og = sigm(po[j] + bias_s[96 + dg_]);
                        const float cn = f * pc[j] + ig * gg;
                        const float tcn = tnh(cn);
                        const float hn = og * tcn;
                        h[d] = act ? hn : 0.0f;
                        pf[j] = f; pi[j] = ig; pg[j] = gg; po[j] = og; pc[j] = cn; ptc[j] = tcn;
                    }
                    tmem_st8(tcs + b * 8, pc);   // (finished sequences carry garbage from here on: never read again as a live value)
                    if (act) {
                        *sb8(t, AF, gb) = pack_bf16x8(pf); *sb8(t, AI, gb) = pack_bf16x8(pi);
                        *sb8(t, AG, gb) = pack_bf16x8(pg); *sb8(t, AO, gb) = pack_bf16x8(po);
                        *sb8(t, AC, gb) = pack_bf16x8(pc); *sb8(t, ATC, gb) = pack_bf16x8(ptc);
                        const float h8[8] = {h[8 * b], h[8 * b + 1], h[8 * b + 2], h[8 * b + 3], h[8 * b + 4], h[8 * b + 5], h[8 * b + 6], h[8 * b + 7]};
                        *sb8(t, AHB, gb) = pack_bf16x8(h8);
                        *sf4(t, 2 * gb) = make_float4(h8[0], h8[1], h8[2], h8[3]);
                        *sf4(t, 2 * gb + 1) = make_float4(h8[4], h8[5], h8[6], h8[7]);
                    }
                }
                tmem_st_wait();
                tc_fence_before_sync();  // TMEM reads ordered before the next MMA (issued after the next tile barrier)
                // scoring + negative sampling (sequence_model.rs:47-68, lstm.rs:300-320)
                cp_wait<1>();            // G1 (target, candidates 0 and 1) has landed; G2 may still fly
                __syncwarp();
                float pos;
                {
                    float a = bp;
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        const float4 pv = slice_ld(SD0, cc);
                        a = fmaf(h[4 * cc], pv.x, a); a = fmaf(h[4 * cc + 1], pv.y, a);
                        a = fmaf(h[4 * cc + 2], pv.z, a); a = fmaf(h[4 * cc + 3], pv.w, a);
                    }
                    pos = xsum(a);
                }
                bool done = !act; uint32_t neg = 0; float ngs = 0.0f;
                float4 qv[NCH];
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) qv[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
                // every lane of every owner takes part in the exchange; only lanes still sampling keep the result
                auto score = [&](const Slice& s, uint32_t cd, float bcj) {
                    float4 qt[NCH];
                    float a = bcj;
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        qt[cc] = slice_ld(s, cc);
                        a = fmaf(h[4 * cc], qt[cc].x, a); a = fmaf(h[4 * cc + 1], qt[cc].y, a);
                        a = fmaf(h[4 * cc + 2], qt[cc].z, a); a = fmaf(h[4 * cc + 3], qt[cc].w, a);
                    }
                    const float tot = xsum(a);
                    if (!done) {
                        neg = cd; ngs = tot;
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) qv[cc] = qt[cc];
                        if (1.0f - pos + tot > 0.0f) done = true;
                    }
                };
                bool alld = __all_sync(kFull, done);   // identical in the DS warps of a quad: same lanes, same decisions
                if (!alld) { score(SD1, cand(0), bc[0]); __syncwarp(); }
                // x_{t+1} = E[ids[t+1]] takes the slice candidate 0 just left (G3)
                if (t + 1 < Tmax) gather_async((t + 1 < Tn) ? idB : 0u, SD1);
                cp_commit();
                if (tries > 1) {
                    alld = __all_sync(kFull, done);
                    if (!alld) {
                        score(SB, cand(1), bc[1]);
                        __syncwarp();
                        gather_async(cand(4), SB);   // candidate 4 takes candidate 1's slice (G4)
                        cp_commit();
                        alld = __all_sync(kFull, done);
                    }
                    if (!alld) {
                        cp_wait<2>();   // G2 (candidates 2, 3) landed; G3, G4 may fly
                        __syncwarp();
                        score(SZ0, cand(2), bc[2]);
                        alld = __all_sync(kFull, done);
                    }
                    if (!alld) { score(SZ1, cand(3), bc[3]); alld = __all_sync(kFull, done); }
                    if (!alld) {
                        cp_wait<0>();
                        __syncwarp();
                        score(SB, cand(4), bc[4]);
                    }
                }
                if (act) {
                    float l, g;
                    if (m.loss == 0) { const float s = sigm(ngs - pos); l = s; g = s * (1.0f - s); }
                    else { const float v = 1.0f + ngs - pos; l = v > 0.0f ? v : 0.0f; g = v > 0.0f ? 1.0f : 0.0f; }
                    loss_seq += l;
#pragma unroll
                    for (int b = 0; b < NB8; ++b) {
                        const float4 pa = slice_ld(SD0, 2 * b), pb = slice_ld(SD0, 2 * b + 1);
                        const float4 qa = qv[2 * b], qb = qv[2 * b + 1];
                        const float d8[8] = {g * (qa.x - pa.x), g * (qa.y - pa.y), g * (qa.z - pa.z), g * (qa.w - pa.w),
                                             g * (qb.x - pb.x), g * (qb.y - pb.y), g * (qb.z - pb.z), g * (qb.w - pb.w)};
                        *sb8(t, ADQ, gb0 + b) = pack_bf16x8(d8);
                    }
                    if (lead) { G_[(size_t)t * gstride] = g; NEG[(size_t)t * gstride] = neg; }
                }
                __syncwarp();            // every lane is done with SD0 before the next step's target row overwrites it
                idB = idC; idC = idD;
            }
            cp_wait<0>();
            __syncwarp();
            quad_bar();   // G_ / NEG written by the lead owner are read by all owners below

            // =========================== backward ===========================
            // dz of timestep t+1 (dh_t in TMEM columns 0..31, dx_{t+1} in 32..63) is consumed straight from TMEM inside
            // timestep t's delta loop -- nothing but the cell-gradient recurrence lives in registers across timesteps.
            {
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int b = 0; b < NB8; ++b) tmem_st8(tcs + b * 8, z8);   // dc_t recurrence starts at 0
                tmem_st_wait();
            }
            float g_c = 0.0f; uint32_t neg_c = 0, out_c = 0;
            ActB cur;
            {
                const int t = Tmax - 1;
                const bool a0 = t >= 0 && t < Tn;
                if (a0) { g_c = __ldcg(G_ + (size_t)t * gstride); neg_c = __ldcg(NEG + (size_t)t * gstride); out_c = __ldg(ids + t + 1); }
                load_act(cur, t > 0 ? t : 0, gb0, a0);
            }
            bool prev_valid = false, prev_act = false;   // a dz of the previous (later) timestep is pending in TMEM
            // t = -1 is a pseudo-timestep: no deltas, no MMA, only the visit of the last deferred entry E[in_0] += step(dx_0)
            for (int t = Tmax - 1; t >= (Tmax > 0 ? -1 : 0); --t) {
                const bool act = t >= 0 && t < Tn;
                const float g = g_c; const uint32_t neg = neg_c, out = out_c;
                const bool actn = t >= 1 && (t - 1) < Tn;   // the next (earlier) timestep
                float g_n = 0.0f; uint32_t neg_n = 0, out_n = 0;
                if (actn) { g_n = __ldcg(G_ + (size_t)(t - 1) * gstride); neg_n = __ldcg(NEG + (size_t)(t - 1) * gstride); out_n = __ldg(ids + t); }
                else if (t == 0 && Tn > 0) out_n = __ldg(ids);   // out_{-1} = ids[0] = in_0
                if (t >= 1) prefetch_step(t - 1);
                const bool triple = act && neg == out;
                const bool has_dx = t + 1 < Tn;   // a deferred E[in_{t+1}] entry exists (t + 1 >= 0 always)
                const uint32_t fl = (act && !triple ? 1u : 0u) | ((act || has_dx) ? 2u : 0u) | (has_dx ? 4u : 0u) | (triple ? 8u : 0u) | (act ? 16u : 0u);
                VisitState VS;
                if (prev_valid) { mbar_wait(mbar + tile, phase); phase ^= 1; tc_fence_after_sync(); }
                if (t >= 0) {
                    stage_z_async(t, act);   // the previous MMA is done with the Z tile
#pragma unroll
                    for (int b = 0; b < NB8; ++b) {
                        const int gb = gb0 + b;
                        if (b > 0) load_act(cur, t, gb, act);   // (no double buffering: 128 registers per thread)
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
                            slice_st(SZ1, 2 * b, make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3])));
                            slice_st(SZ1, 2 * b + 1, make_float4(__uint_as_float(rb[4]), __uint_as_float(rb[5]), __uint_as_float(rb[6]), __uint_as_float(rb[7])));
                        }
                        // gradient of the two rows that only need h_t goes straight to the staging slice
                        slice_st(SZ0, 2 * b, make_float4(g * cur.h0.x, g * cur.h0.y, g * cur.h0.z, g * cur.h0.w));
                        slice_st(SZ0, 2 * b + 1, make_float4(g * cur.h1.x, g * cur.h1.y, g * cur.h1.z, g * cur.h1.w));
                        uint32_t wdf[4], wdi[4], wdg[4], wdo[4];
#pragma unroll
                        for (int pr = 0; pr < 4; ++pr) {   // two hidden units at a time, straight from / to packed bf16 words
                            const uint32_t uf = word_of(cur.f, pr), ui = word_of(cur.i, pr), ug = word_of(cur.g, pr), uo = word_of(cur.o, pr);
                            const uint32_t uq = word_of(cur.q, pr), ucp = word_of(cur.cp, pr), utc = word_of(cur.tc, pr);
                            float rdf[2], rdi[2], rdg[2], rdo[2];
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const int e = 2 * pr + k;
                                const float f_ = k ? bf_hi(uf) : bf_lo(uf), i_ = k ? bf_hi(ui) : bf_lo(ui), g_ = k ? bf_hi(ug) : bf_lo(ug);
                                const float o_ = k ? bf_hi(uo) : bf_lo(uo), q_ = k ? bf_hi(uq) : bf_lo(uq);
                                const float cp_ = k ? bf_hi(ucp) : bf_lo(ucp), tcv = k ? bf_hi(utc) : bf_lo(utc);
                                const float dh = dhv[e] + q_;
                                const float d_o = dh * tcv;
                                const float dc = dcv[e] + dh * o_ * (1.0f - tcv * tcv);
                                float d_f = dc * cp_, d_i = dc * g_;
                                const float d_g = dc * i_;
                                dcv[e] = act ? dc * f_ : 0.0f;
                                if (coupled) { d_f -= d_i; d_i = 0.0f; }
                                rdf[k] = d_f * f_ * (1.0f - f_);
                                rdi[k] = coupled ? 0.0f : d_i * i_ * (1.0f - i_);
                                rdg[k] = d_g * (1.0f - g_ * g_);
                                rdo[k] = d_o * o_ * (1.0f - o_);
                            }
                            wdf[pr] = pack2(rdf[0], rdf[1]); wdi[pr] = pack2(rdi[0], rdi[1]);
                            wdg[pr] = pack2(rdg[0], rdg[1]); wdo[pr] = pack2(rdo[0], rdo[1]);
                        }
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 0 + gb, 16)) = make_uint4(wdf[0], wdf[1], wdf[2], wdf[3]);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 4 + gb, 16)) = make_uint4(wdi[0], wdi[1], wdi[2], wdi[3]);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 8 + gb, 16)) = make_uint4(wdg[0], wdg[1], wdg[2], wdg[3]);
                        *reinterpret_cast<uint4*>(Db + tile_chunk_off(r, 12 + gb, 16)) = make_uint4(wdo[0], wdo[1], wdo[2], wdo[3]);
                        tmem_st8(tcs + b * 8, dcv);
                    }
                    tmem_st_wait();
                    __syncwarp();             // the g h_t and dx_{t+1} slices are complete
                    visit_issue(0, VS, neg, out, fl, SZ0, SZ1, o);   // first pass of the visits flies across the barrier and the MMA issue
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
                    // first activation block of the next timestep: requested before this timestep's visits
                    if (t >= 1) load_act(cur, t - 1, gb0, actn);
                } else {   // t == -1: dx_0 out of TMEM, nothing else
#pragma unroll
                    for (int b = 0; b < NB8; ++b) {
                        uint32_t rb[8];
                        tmem_ld8_issue(tbase + 32 + (gb0 + b) * 8, rb);
                        tmem_wait8(rb);
                        slice_st(SZ1, 2 * b, make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3])));
                        slice_st(SZ1, 2 * b + 1, make_float4(__uint_as_float(rb[4]), __uint_as_float(rb[5]), __uint_as_float(rb[6]), __uint_as_float(rb[7])));
                    }
                    tc_fence_before_sync();
                    __syncwarp();            // the dx_0 slice is complete
                    visit_issue(0, VS, neg, out, fl, SZ0, SZ1, o);
                }
                // ---- sparse visits of this timestep (overlap the MMAs): E[neg_t]; E[out_t] with the deferred E[in_{t+1}] ----
                {
                    float4* rn = bias_rec(m, neg); float4* ro = bias_rec(m, out);
                    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
                    const bool bv = act && lead;
                    const bool batom = !o.adam && !noatom;
                    float* fn_ = reinterpret_cast<float*>(rn); float* fo_ = reinterpret_cast<float*>(ro);
                    if (bv) {
                        if (batom) {   // {ld b, atom G_b += g^2}: b[neg] takes +g; b[out] takes -g (two entries on one record when neg == out)
                            ba.x = __ldcg(fn_); ba.y = atomicAdd(fn_ + 1, neg != out ? g * g : 2.0f * g * g);
                            if (neg != out) { bb.x = __ldcg(fo_); bb.y = atomicAdd(fo_ + 1, g * g); }
                        } else { ba = __ldcg(rn); if (neg != out) bb = __ldcg(ro); }
                    }
                    visit_finish(VS, SZ0, SZ1, o);
#pragma unroll
                    for (int pass = 1; pass < NGI / GPP; ++pass) { visit_issue(pass, VS, neg, out, fl, SZ0, SZ1, o); visit_finish(VS, SZ0, SZ1, o); }
                    __syncwarp();
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
                prev_valid = true; prev_act = act;
                g_c = g_n; neg_c = neg_n; out_c = out_n;
            }
            if (live && lead) { pl.loss_acc[p] += loss_seq; pl.examples[p] += (unsigned long long)Tn; }

            // =========================== dense step on the CTA-summed gradient ===========================
            // TMEM lane gd = r of a tile holds row gd of its dW^T: columns 0..63 = dW[k][gd], column 64 = dbias[gd];
            // owner `part` takes columns [part * 64/DS, (part+1) * 64/DS), the last owner also column 64
            __syncthreads();
            constexpr int CW = 64 / DS;
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
                OptCfg od = o;
                adam_corrections(od, pl.adam_t0 + step * pl.P + (uint64_t)blockIdx.x * NT * 128 + 1);
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
                    } else adagrad1(w, s1, dwr[kk], od.lr, od.l2);
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
    if (live && lead) pl.step_ctr[p] = step;
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc<(NT == 1 ? 256 : 512)>(*tmem_ptr);
}

template <int NT, int DS, bool FLAT>
cudaError_t launch_one(const ModelDev& m, const PlanDev& p, cudaStream_t st) {
    const size_t smem = OFF_TILES + (size_t)NT * TILE_BYTES + (size_t)NT * 4 * XS_BYTES_PER_QUAD;
    const int seq_per_cta = 128 * NT;
    dim3 grid((p.P + seq_per_cta - 1) / seq_per_cta);
    cudaError_t e = cudaFuncSetAttribute(lstm_tc3_train_kernel<NT, DS, FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lstm_tc3_train_kernel<NT, DS, FLAT><<<grid, seq_per_cta * DS, smem, st>>>(m, p);
    return cudaGetLastError();
}

#undef NEG

}  // namespace

cudaError_t launch_lstm_tc3(const ModelDev& m, const PlanDev& p, int nt, cudaStream_t st) {
    const bool flat = m.gmask == 0;
    if (nt == 2) return flat ? launch_one<2, 2, true>(m, p, st) : launch_one<2, 2, false>(m, p, st);
    return flat ? launch_one<1, 2, true>(m, p, st) : launch_one<1, 2, false>(m, p, st);
}

}  // namespace sbr
