// common.cuh -- device helpers shared by the sm_100a kernels of the sbr hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbr {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// RNG.  xorshift128 mirrors rand 0.5 XorShiftRng (per-partition epoch shuffles, sequence_model.rs:97,109);
// draw_item is the counter-based negative sampler shared bit-for-bit with oracle/sbr_oracle.c:sbo_draw_item.
// ---------------------------------------------------------------------------------------------
struct XorShift { uint32_t x, y, z, w; };

__host__ __device__ inline uint32_t xs_next_u32(XorShift& r) {
    uint32_t t = r.x ^ (r.x << 11);
    r.x = r.y; r.y = r.z; r.z = r.w;
    r.w = r.w ^ (r.w >> 19) ^ (t ^ (t >> 8));
    return r.w;
}
__host__ __device__ inline uint64_t xs_next_u64(XorShift& r) {
    uint64_t lo = xs_next_u32(r);
    uint64_t hi = xs_next_u32(r);
    return (hi << 32) | lo;
}
__host__ __device__ inline uint64_t mulhi64(uint64_t a, uint64_t b, uint64_t* lo) {
#ifdef __CUDA_ARCH__
    *lo = a * b;
    return __umul64hi(a, b);
#else
    unsigned __int128 m = (unsigned __int128)a * b;
    *lo = (uint64_t)m;
    return (uint64_t)(m >> 64);
#endif
}
__host__ __device__ inline int clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
// Rng::gen_range(0, high) for usize (widening multiply + rejection zone)
__host__ __device__ inline uint64_t xs_gen_below(XorShift& r, uint64_t range) {
    if (range == 0) return 0;
    uint64_t zone = (range << clz64(range)) - 1;
    for (;;) {
        uint64_t v = xs_next_u64(r), lo;
        uint64_t hi = mulhi64(v, range, &lo);
        if (lo <= zone) return hi;
    }
}
__host__ __device__ inline void xs_from_seed(XorShift& r, const uint8_t seed[16]) {
    uint32_t s[4];
    for (int i = 0; i < 4; ++i)
        s[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
               ((uint32_t)seed[4 * i + 3] << 24);
    if ((s[0] | s[1] | s[2] | s[3]) == 0) { s[0] = 0x193a6754u; s[1] = 0xa8a7d469u; s[2] = 0x97830e05u; s[3] = 0x113ba7bbu; }
    r.x = s[0]; r.y = s[1]; r.z = s[2]; r.w = s[3];
}

__host__ __device__ inline uint32_t draw_item(uint64_t key, uint64_t step, uint32_t t, uint32_t j, uint32_t num_items) {
    uint64_t v = key + step * 0x9E3779B97F4A7C15ULL + ((uint64_t)t * 8u + j) * 0xD1B54A32D192ED03ULL;
    v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ULL;
    v ^= v >> 27; v *= 0x94D049BB133111EBULL;
    v ^= v >> 31;
    uint32_t r = (uint32_t)(v >> 32);
    return (uint32_t)(((uint64_t)r * (uint64_t)num_items) >> 32);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Parameter access.  All live (Hogwild-shared) parameters are read/written at L2 (.cg): L1 is not coherent
// across SMs and a row may be rewritten by any other warp at any time.
// Row layout in HBM ("row record"): [w[D] | s1[D] | (s2[D])] contiguous, so one optimizer visit touches one
// contiguous 8D- (Adagrad) or 12D-byte (Adam) region.  Lane ownership of a D-vector:
//   D == 16 : lane l < 16 owns element l            (V = 1)
//   D == 32*V, V in {1,2,4}: lane l owns l*V..l*V+V-1 (one 4V-byte vector access)
//   D == 256 (V = 8): lane l owns j*128 + l*4 + i    (two float4 accesses)
// ---------------------------------------------------------------------------------------------
template <int D> struct VecOf { static constexpr int V = (D + 31) / 32; };

template <int D>
__device__ __forceinline__ void row_load_cg(const float* __restrict__ p, int lane, float (&r)[VecOf<D>::V]) {
    constexpr int V = VecOf<D>::V;
    if constexpr (D < 32) {
        r[0] = lane < D ? __ldcg(p + lane) : 0.0f;
    } else if constexpr (V == 1) {
        r[0] = __ldcg(p + lane);
    } else if constexpr (V == 2) {
        float2 t = __ldcg(reinterpret_cast<const float2*>(p) + lane);
        r[0] = t.x; r[1] = t.y;
    } else {
#pragma unroll
        for (int j = 0; j < V / 4; ++j) {
            float4 t = __ldcg(reinterpret_cast<const float4*>(p) + j * 32 + lane);
            r[4 * j] = t.x; r[4 * j + 1] = t.y; r[4 * j + 2] = t.z; r[4 * j + 3] = t.w;
        }
    }
}

template <int D>
__device__ __forceinline__ void row_store_cg(float* __restrict__ p, int lane, const float (&r)[VecOf<D>::V]) {
    constexpr int V = VecOf<D>::V;
    if constexpr (D < 32) {
        if (lane < D) __stcg(p + lane, r[0]);
    } else if constexpr (V == 1) {
        __stcg(p + lane, r[0]);
    } else if constexpr (V == 2) {
        __stcg(reinterpret_cast<float2*>(p) + lane, make_float2(r[0], r[1]));
    } else {
#pragma unroll
        for (int j = 0; j < V / 4; ++j)
            __stcg(reinterpret_cast<float4*>(p) + j * 32 + lane, make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
    }
}

// warp-private scratch (global or shared; generic addressing, default caching)
template <int D>
__device__ __forceinline__ void vec_load(const float* p, int lane, float (&r)[VecOf<D>::V]) {
    constexpr int V = VecOf<D>::V;
    if constexpr (D < 32) {
        r[0] = lane < D ? p[lane] : 0.0f;
    } else if constexpr (V == 1) {
        r[0] = p[lane];
    } else if constexpr (V == 2) {
        float2 t = reinterpret_cast<const float2*>(p)[lane];
        r[0] = t.x; r[1] = t.y;
    } else {
#pragma unroll
        for (int j = 0; j < V / 4; ++j) {
            float4 t = reinterpret_cast<const float4*>(p)[j * 32 + lane];
            r[4 * j] = t.x; r[4 * j + 1] = t.y; r[4 * j + 2] = t.z; r[4 * j + 3] = t.w;
        }
    }
}
template <int D>
__device__ __forceinline__ void vec_store(float* p, int lane, const float (&r)[VecOf<D>::V]) {
    constexpr int V = VecOf<D>::V;
    if constexpr (D < 32) {
        if (lane < D) p[lane] = r[0];
    } else if constexpr (V == 1) {
        p[lane] = r[0];
    } else if constexpr (V == 2) {
        reinterpret_cast<float2*>(p)[lane] = make_float2(r[0], r[1]);
    } else {
#pragma unroll
        for (int j = 0; j < V / 4; ++j)
            reinterpret_cast<float4*>(p)[j * 32 + lane] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
}

template <int D>
__device__ __forceinline__ float warp_dot(const float (&a)[VecOf<D>::V], const float (&b)[VecOf<D>::V]) {
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < VecOf<D>::V; ++v) s = fmaf(a[v], b[v], s);
    return warp_sum(s);
}

// ---------------------------------------------------------------------------------------------
// optimizers (wyrm::optim::{Adagrad, Adam} semantics as restated in oracle/sbr_oracle.c)
// ---------------------------------------------------------------------------------------------
struct OptCfg {
    float lr, l2;
    int adam;       // 0 Adagrad, 1 Adam
    float c1, c2;   // Adam bias corrections 1-b1^t, 1-b2^t for the current step
};

__device__ __forceinline__ void adagrad_elem(float& w, float& G, float g, float lr, float l2) {
    g = g + w * l2;
    G += g * g;
    w -= lr / (1e-10f + sqrtf(G)) * g;
}
__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, const OptCfg& o) {
    g = g + w * o.l2;
    m = 0.9f * m + (1.0f - 0.9f) * g;
    v = 0.999f * v + (1.0f - 0.999f) * g * g;
    float mhat = m / o.c1, vhat = v / o.c2;
    w -= o.lr / (sqrtf(vhat) + 1e-8f) * mhat;
}

// One sparse optimizer visit of an item row record by a whole warp (Hogwild: plain L2 read-modify-write).
template <int D>
__device__ __forceinline__ void update_row(float* __restrict__ rec, int lane, const float (&g)[VecOf<D>::V], const OptCfg& o) {
    constexpr int V = VecOf<D>::V;
    float w[V], s1[V];
    row_load_cg<D>(rec, lane, w);
    row_load_cg<D>(rec + D, lane, s1);
    if (!o.adam) {
#pragma unroll
        for (int v = 0; v < V; ++v) adagrad_elem(w[v], s1[v], g[v], o.lr, o.l2);
        row_store_cg<D>(rec, lane, w);
        row_store_cg<D>(rec + D, lane, s1);
    } else {
        float s2[V];
        row_load_cg<D>(rec + 2 * D, lane, s2);
#pragma unroll
        for (int v = 0; v < V; ++v) adam_elem(w[v], s1[v], s2[v], g[v], o);
        row_store_cg<D>(rec, lane, w);
        row_store_cg<D>(rec + D, lane, s1);
        row_store_cg<D>(rec + 2 * D, lane, s2);
    }
}

// ---- Adagrad visits through L2 atomics (Hogwild with more than one partition) ----
// A visit as load -> compute -> store turns every popular row into a read-modify-write hazard between thousands of
// concurrent partitions (lost updates, and on B200 a measured 2-3x slow-down on the 1,683-row ML-100K table).  The
// accumulator is additive, so: {ld w, atom.add G += g^2 (returns the old G)} in one round trip, the Adagrad step in
// registers from the returned value, then fire-and-forget reductions w += dw and G += (l2 cross terms).
template <int D>
__device__ __forceinline__ void row_atom_add(float* __restrict__ p, int lane, const float (&v)[VecOf<D>::V], float (&old)[VecOf<D>::V]) {
    constexpr int V = VecOf<D>::V;
    if constexpr (D < 32) {
        old[0] = lane < D ? atomicAdd(p + lane, v[0]) : 0.0f;
    } else if constexpr (V == 1) {
        old[0] = atomicAdd(p + lane, v[0]);
    } else if constexpr (V == 2) {
        asm volatile("atom.global.add.v2.f32 {%0, %1}, [%2], {%3, %4};" : "=f"(old[0]), "=f"(old[1]) : "l"(p + 2 * lane), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
#pragma unroll
        for (int j = 0; j < V / 4; ++j)
            asm volatile("atom.global.add.v4.f32 {%0, %1, %2, %3}, [%4], {%5, %6, %7, %8};"
                         : "=f"(old[4 * j]), "=f"(old[4 * j + 1]), "=f"(old[4 * j + 2]), "=f"(old[4 * j + 3])
                         : "l"(p + j * 128 + lane * 4), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
    }
}
template <int D>
__device__ __forceinline__ void row_red_add(float* __restrict__ p, int lane, const float (&v)[VecOf<D>::V]) {
    constexpr int V = VecOf<D>::V;
    if constexpr (D < 32) {
        if (lane < D) atomicAdd(p + lane, v[0]);
    } else if constexpr (V == 1) {
        atomicAdd(p + lane, v[0]);
    } else if constexpr (V == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p + 2 * lane), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
#pragma unroll
        for (int j = 0; j < V / 4; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + j * 128 + lane * 4), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                         "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
    }
}
// the register part of an atomic visit: from w and the G returned by the atom (which already added q = g^2) to the deltas
__device__ __forceinline__ void adagrad_atomic_elem(float w, float Gold, float g, float q, float lr, float l2, float& dw, float& dG) {
    const float gg = g + w * l2;
    const float Gn = Gold + gg * gg;
    dw = -(lr / (1e-10f + sqrtf(Gn)) * gg);
    dG = Gn - Gold - q;
}
template <int D>
__device__ __forceinline__ void update_row_atomic(float* __restrict__ rec, int lane, const float (&g)[VecOf<D>::V], const OptCfg& o) {
    constexpr int V = VecOf<D>::V;
    float w[V], q[V], Gold[V], dw[V], dG[V];
    row_load_cg<D>(rec, lane, w);
#pragma unroll
    for (int v = 0; v < V; ++v) q[v] = g[v] * g[v];
    row_atom_add<D>(rec + D, lane, q, Gold);
#pragma unroll
    for (int v = 0; v < V; ++v) adagrad_atomic_elem(w[v], Gold[v], g[v], q[v], o.lr, o.l2, dw[v], dG[v]);
    row_red_add<D>(rec, lane, dw);
    if (o.l2 != 0.0f) row_red_add<D>(rec + D, lane, dG);
}
__device__ __forceinline__ void update_bias_atomic(float4* __restrict__ rec, float g, const OptCfg& o) {
    float* f = reinterpret_cast<float*>(rec);
    const float q = g * g;
    const float b = __ldcg(f);
    const float Gold = atomicAdd(f + 1, q);
    float db, dG;
    adagrad_atomic_elem(b, Gold, g, q, o.lr, o.l2, db, dG);
    atomicAdd(f, db);
    if (o.l2 != 0.0f) atomicAdd(f + 1, dG);
}

// bias record: [b, s1, s2, pad] (float4) -- one lane does the visit
__device__ __forceinline__ void update_bias(float4* __restrict__ rec, float g, const OptCfg& o) {
    float4 r = __ldcg(rec);
    if (!o.adam) adagrad_elem(r.x, r.y, g, o.lr, o.l2);
    else adam_elem(r.x, r.y, r.z, g, o);
    __stcg(rec, r);
}

__device__ __forceinline__ void adam_corrections(OptCfg& o, uint64_t t) {
    if (o.adam) {
        o.c1 = 1.0f - powf(0.9f, (float)t);
        o.c2 = 1.0f - powf(0.999f, (float)t);
    }
}
#endif  // __CUDACC__

}  // namespace sbr
