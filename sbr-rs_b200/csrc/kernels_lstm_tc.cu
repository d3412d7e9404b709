// kernels_lstm_tc.cu -- tensor-core LSTM training kernel (D = 32), tcgen05 + TMEM, thread-per-sequence epilogue.
//
// One tile = 128 partitions (reference "threads") advancing in lock-step through their current sub-sequence; one
// thread owns one partition.  Per timestep the three contractions of the LSTM (lstm.rs:293-298 -> wyrm::nn::lstm) run
// on the 5th-gen tensor cores from shared-memory tiles, accumulators in TMEM:
//   forward   gates[128 x 128] = Z_t[128 x 64] . W                (kind::tf32; Z_t = [h_{t-1} , x_t])
//   backward  dz   [128 x 64]  = delta_t[128 x 128] . W^T         (kind::f16 / bf16)
//             dW^T [128 x 80] += delta_t^T . [Z_t , 1 , 0..]      (bf16; column 64 accumulates the bias gradient)
// and everything else (gather, gate non-linearities, scoring, WARP sampling, loss, deltas, sparse optimizer visits)
// is per-thread work on that thread's own sequence: no shuffles, no replication.  The delta tile is read twice from
// the same bytes (K-major for dz, MN-major for dW^T), see tc_tile.cuh.
// After every round (each partition finished one sub-sequence) the CTA applies ONE dense optimizer step with the
// gradient summed over the CTA's 128*NT sequences, Hogwild across CTAs, and restages the weight tiles.
// Numerics: "throughput mode" -- tf32 forward products, bf16 backward products, fp32 accumulation, MUFU-based
// exp / rcp / rsqrt.  The FFMA kernel in kernels_train.cu is the exact-fp32 path (used when num_threads < 128).
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

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.0f * fast_sigmoid(2.0f * x) - 1.0f; }

__device__ __forceinline__ void fast_adagrad(float& w, float& G, float g, float lr, float l2) {
    g = fmaf(w, l2, g);
    G = fmaf(g, g, G);
    w -= lr * g * rsqrtf(fmaxf(G, 1e-20f));  // lr / (1e-10 + sqrt(G)) * g
}
__device__ __forceinline__ void fast_adam(float& w, float& m, float& v, float g, const OptCfg& o) {
    g = fmaf(w, o.l2, g);
    m = 0.9f * m + 0.1f * g;
    v = 0.999f * v + 0.001f * g * g;
    const float mhat = __fdividef(m, o.c1), vhat = __fdividef(v, o.c2);
    w -= __fdividef(o.lr * mhat, sqrtf(vhat) + 1e-8f);
}

// ---------------------------------------------------------------------------------------------------------
// Warp-cooperative row traffic.  The epilogue owns one sequence per THREAD, but item rows are 128-byte lines in
// HBM/L2: a per-thread row access would touch 32 different lines per instruction (32 LSU wavefronts).  Instead a
// warp moves its 32 rows 4 at a time (8 lanes x 16 B per row = one full line per quarter-warp) through a staging
// buffer [32 rows][36 floats] in shared memory; 144-byte rows make both the quarter-warp-per-row 16-byte accesses
// and the lane-per-row 16-byte accesses bank-conflict free.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSS = 36;

__device__ __forceinline__ void coop_gather(const ModelDev& m, uint32_t my_id, int lane, float* stage) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + (lane >> 3);
        const uint32_t id = __shfl_sync(kFull, my_id, row);
        const float4 v = __ldcg(reinterpret_cast<const float4*>(item_rec(m, id)) + (lane & 7));
        *reinterpret_cast<float4*>(stage + row * kSS + (lane & 7) * 4) = v;
    }
    __syncwarp();
}
__device__ __forceinline__ void stage_read_row(const float* stage, int lane, float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(stage + lane * kSS + c * 4);
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
}
__device__ __forceinline__ void stage_write_row(float* stage, int lane, const float (&v)[32], float scale) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(stage + lane * kSS + c * 4) =
            make_float4(scale * v[4 * c], scale * v[4 * c + 1], scale * v[4 * c + 2], scale * v[4 * c + 3]);
    __syncwarp();
}
// sparse optimizer visit of the 32 rows named by the lanes' ids with the staged gradients (times `sign`).
// All 16 row loads are issued before the first use: the 8 row groups belong to 8 x 4 different sequences, and two
// sequences naming the same row is the same benign Hogwild race as between warps (last store wins).
__device__ __forceinline__ void coop_update(const ModelDev& m, uint32_t my_id, bool my_act, int lane,
                                            const float* stage, float sign, const OptCfg& o) {
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {   // two batches of four row groups: 8 loads in flight, 32 registers
        float* rec[4]; bool a[4]; float4 w[4], s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = 4 * (hb * 4 + i) + (lane >> 3);
            const uint32_t id = __shfl_sync(kFull, my_id, row);
            a[i] = __shfl_sync(kFull, (int)my_act, row) != 0;
            rec[i] = item_rec(m, id) + (lane & 7) * 4;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (a[i]) { w[i] = __ldcg(reinterpret_cast<const float4*>(rec[i])); s[i] = __ldcg(reinterpret_cast<const float4*>(rec[i] + kD)); }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (a[i]) {
                const int row = 4 * (hb * 4 + i) + (lane >> 3);
                const float4 g = *reinterpret_cast<const float4*>(stage + row * kSS + (lane & 7) * 4);
                if (!o.adam) {
                    fast_adagrad(w[i].x, s[i].x, sign * g.x, o.lr, o.l2); fast_adagrad(w[i].y, s[i].y, sign * g.y, o.lr, o.l2);
                    fast_adagrad(w[i].z, s[i].z, sign * g.z, o.lr, o.l2); fast_adagrad(w[i].w, s[i].w, sign * g.w, o.lr, o.l2);
                } else {
                    float4 v = __ldcg(reinterpret_cast<const float4*>(rec[i] + 2 * kD));
                    fast_adam(w[i].x, s[i].x, v.x, sign * g.x, o); fast_adam(w[i].y, s[i].y, v.y, sign * g.y, o);
                    fast_adam(w[i].z, s[i].z, v.z, sign * g.z, o); fast_adam(w[i].w, s[i].w, v.w, sign * g.w, o);
                    __stcg(reinterpret_cast<float4*>(rec[i] + 2 * kD), v);
                }
                __stcg(reinterpret_cast<float4*>(rec[i]), w[i]); __stcg(reinterpret_cast<float4*>(rec[i] + kD), s[i]);
            }
        }
    }
    __syncwarp();
}
// the two bias visits of one timestep (b[neg] += g-step, b[out] -= g-step); loads first unless they alias
__device__ __forceinline__ void thread_update_biases(const ModelDev& m, uint32_t neg, uint32_t out, float g, const OptCfg& o) {
    float4* rn = bias_rec(m, neg); float4* ro = bias_rec(m, out);
    float4 a = __ldcg(rn);
    if (neg != out) {
        float4 b = __ldcg(ro);
        if (!o.adam) { fast_adagrad(a.x, a.y, g, o.lr, o.l2); fast_adagrad(b.x, b.y, -g, o.lr, o.l2); }
        else { fast_adam(a.x, a.y, a.z, g, o); fast_adam(b.x, b.y, b.z, -g, o); }
        __stcg(rn, a); __stcg(ro, b);
    } else {
        if (!o.adam) { fast_adagrad(a.x, a.y, g, o.lr, o.l2); fast_adagrad(a.x, a.y, -g, o.lr, o.l2); }
        else { fast_adam(a.x, a.y, a.z, g, o); fast_adam(a.x, a.y, a.z, -g, o); }
        __stcg(rn, a);
    }
}
__device__ __forceinline__ void tile_bar(int tile) { asm volatile("bar.sync %0, 128;" ::"r"(tile + 1) : "memory"); }

enum { SX = 0, SH = 1, SC = 2, SF = 3, SI = 4, SG = 5, SO = 6, SDQ = 7 };

template <int NT>
__global__ void __launch_bounds__(128 * NT, 1) lstm_tc_train_kernel(ModelDev m, PlanDev pl) {
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
    // staging buffers live in whichever tile is idle in the current phase
    float* stage_f = reinterpret_cast<float*>(Db) + wq * (32 * kSS);   // forward: delta tile is unused
    float* stage_b = reinterpret_cast<float*>(Zt) + wq * (32 * kSS);   // backward: tf32 Z tile is unused
    const uint32_t tile_gid = blockIdx.x * NT + tile;
    const uint32_t p = tile_gid * 128u + r;
    const bool live = p < pl.P;
    const size_t nd = m.ndense;
    const bool coupled = m.variant == 1;
    const int T = m.T;
    // tile scratch, 16-byte structure-of-arrays so that lane == sequence accesses are coalesced (512 B per warp
    // instruction): [T][8 arrays][8 chunks of 4 d][128 seq] float4, then G, NEG [T][128]
    float* sbase = pl.scratch + (size_t)tile_gid * 128 * pl.scratch_stride;
    float* G_ = sbase + (size_t)T * 8 * 32 * 128; uint32_t* NEG = reinterpret_cast<uint32_t*>(G_ + (size_t)T * 128);
    // per timestep: fp32 arrays C, H ([2][8 chunks][128] float4) then bf16 arrays F, I, G, O, X, DQ ([6][4 chunks][128] uint4):
    // 80 KB per tile-timestep.  The backward products are bf16 anyway (delta tile, Z tile), so gates / x / dq are kept in
    // bf16; the cell and hidden states stay fp32.
    constexpr int kStepU4 = (16 + 24) * 128;   // uint4 units per timestep
    auto sf4 = [&](int t, int which /*0 C, 1 H*/, int c4) -> float4* {
        return reinterpret_cast<float4*>(sbase) + (size_t)t * kStepU4 + (size_t)(which * 8 + c4) * 128 + r;
    };
    auto sb8 = [&](int t, int which /*0 F,1 I,2 G,3 O,4 X,5 DQ*/, int c8) -> uint4* {
        return reinterpret_cast<uint4*>(sbase) + (size_t)t * kStepU4 + (size_t)(16 + which * 4 + c8) * 128 + r;
    };
    // pull one timestep of the tile's scratch (80 KB = 640 lines) towards L2: 5 lines per thread
    auto prefetch_step = [&](int t) {
        const char* base = reinterpret_cast<const char*>(sbase) + (size_t)t * kStepU4 * 16 + (size_t)r * 5 * 128;
#pragma unroll
        for (int i = 0; i < 5; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + i * 128));
    };
    auto unpack8 = [](const uint4& u, float (&v)[8]) {
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
        v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
        v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
        v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
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
            for (int t = 0; t < Tmax; ++t) {
                const bool act = t < Tn;
                uint32_t in = 0, out = 0;
                if (act) { in = __ldg(ids + t); out = __ldg(ids + t + 1); }
                {
                    float x[32];
                    coop_gather(m, in, lane, stage_f);       // item_embeddings.index(input), one line per row
                    stage_read_row(stage_f, lane, x);
                    __syncwarp();
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
                            *sb8(t, 4, c8) = pack_bf16x8(x8);
                        }
                    }
                }
                fence_async_smem();
                tc_fence_before_sync();
                tile_bar(tile);
                if (r == 0) {
                    tc_fence_after_sync();
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_tf32(tcol0, make_smem_desc(zt_a + k * 256, 128, 2048), make_smem_desc(wt_a + k * 256, 128, 2048), IDESC_G1, k > 0);
                    mma_commit(mbar + tile);
                }
                // overlap with the MMA: target rows
                float pv[32], qv[32];
                coop_gather(m, out, lane, stage_f);
                stage_read_row(stage_f, lane, pv);
                __syncwarp();
                const float bp = act ? __ldcg(reinterpret_cast<const float*>(bias_rec(m, out))) : 0.0f;
                mbar_wait(mbar + tile, phase); phase ^= 1;
                tc_fence_after_sync();
#pragma unroll
                for (int db = 0; db < 4; ++db) {
                    float pf[8], pi[8], pg[8], po[8];
                    tmem_ld8x4(tbase + db * 8, tbase + 32 + db * 8, tbase + 64 + db * 8, tbase + 96 + db * 8, pf, pi, pg, po);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int d = db * 8 + j;
                        const float f = fast_sigmoid(pf[j] + bias_s[d]);
                        const float ig = coupled ? 1.0f - f : fast_sigmoid(pi[j] + bias_s[32 + d]);
                        const float gg = fast_tanh(pg[j] + bias_s[64 + d]);
                        const float og = fast_sigmoid(po[j] + bias_s[96 + d]);
                        const float cn = f * c[d] + ig * gg;
                        const float hn = og * fast_tanh(cn);
                        if (act) { c[d] = cn; h[d] = hn; } else h[d] = 0.0f;
                        pf[j] = f; pi[j] = ig; pg[j] = gg; po[j] = og;
                    }
                    if (act) {
                        *sb8(t, 0, db) = pack_bf16x8(pf); *sb8(t, 1, db) = pack_bf16x8(pi);
                        *sb8(t, 2, db) = pack_bf16x8(pg); *sb8(t, 3, db) = pack_bf16x8(po);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int c4 = db * 2 + hf;
                            *sf4(t, 0, c4) = make_float4(c[4 * c4], c[4 * c4 + 1], c[4 * c4 + 2], c[4 * c4 + 3]);
                            *sf4(t, 1, c4) = make_float4(h[4 * c4], h[4 * c4 + 1], h[4 * c4 + 2], h[4 * c4 + 3]);
                        }
                    }
                }
                tc_fence_before_sync();  // TMEM reads ordered before the next MMA (issued after the next tile barrier)
                // scoring + negative sampling (sequence_model.rs:47-68, lstm.rs:300-320)
                float pos = 0.0f;
#pragma unroll
                for (int d = 0; d < 32; ++d) pos = fmaf(h[d], pv[d], pos);
                pos += bp;
                bool done = !act; uint32_t neg = 0; float ngs = 0.0f;
#pragma unroll
                for (int d = 0; d < 32; ++d) qv[d] = 0.0f;
                for (int j = 0; j < tries; ++j) {
                    if (__all_sync(kFull, done)) break;
                    const uint32_t cand = done ? neg : draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range);
                    coop_gather(m, cand, lane, stage_f);
                    if (!done) {
                        neg = cand;
                        stage_read_row(stage_f, lane, qv);
                        ngs = 0.0f;
#pragma unroll
                        for (int d = 0; d < 32; ++d) ngs = fmaf(h[d], qv[d], ngs);
                        ngs += __ldcg(reinterpret_cast<const float*>(bias_rec(m, neg)));
                        if (1.0f - pos + ngs > 0.0f) done = true;
                    }
                    __syncwarp();
                }
                if (act) {
                    float l, g;
                    if (m.loss == 0) { const float s = fast_sigmoid(ngs - pos); l = s; g = s * (1.0f - s); }
                    else { const float v = 1.0f + ngs - pos; l = v > 0.0f ? v : 0.0f; g = v > 0.0f ? 1.0f : 0.0f; }
                    loss_seq += l;
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float d8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) d8[e] = g * (qv[8 * c8 + e] - pv[8 * c8 + e]);
                        *sb8(t, 5, c8) = pack_bf16x8(d8);
                    }
                    G_[(size_t)t * 128 + r] = g; NEG[(size_t)t * 128 + r] = neg;
                }
            }

            // The forward staging buffers alias the delta tile that backward is about to fill: no warp may start
            // writing deltas while another warp of the tile is still sampling negatives through its staging slice.
            tile_bar(tile);

            // =========================== backward ===========================
            float dh_rec[32], dc_rec[32];
#pragma unroll
            for (int d = 0; d < 32; ++d) { dh_rec[d] = 0.0f; dc_rec[d] = 0.0f; }
            for (int t = Tmax - 1; t >= 0; --t) {
                const bool act = t < Tn;
                float g = 0.0f; uint32_t neg = 0, in = 0, out = 0;
                if (act) { g = G_[(size_t)t * 128 + r]; neg = NEG[(size_t)t * 128 + r]; in = __ldg(ids + t); out = __ldg(ids + t + 1); }
#pragma unroll
                for (int db = 0; db < 4; ++db) {
                    float df[8], di[8], dg[8], dO[8], hp8[8];
                    float f8[8], i8[8], g8[8], o8[8], q8[8], c8v[8], cp8[8], h8[8];
                    uint4 xraw = make_uint4(0u, 0u, 0u, 0u);
                    if (act) {
                        const uint4 uf = *sb8(t, 0, db), ui = *sb8(t, 1, db), ug = *sb8(t, 2, db), uo = *sb8(t, 3, db), uq = *sb8(t, 5, db);
                        xraw = *sb8(t, 4, db);
                        const float4 ca = *sf4(t, 0, 2 * db), cb = *sf4(t, 0, 2 * db + 1), ha = *sf4(t, 1, 2 * db), hb = *sf4(t, 1, 2 * db + 1);
                        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa, hpa = pa, hpb = pa;
                        if (t > 0) { pa = *sf4(t - 1, 0, 2 * db); pb = *sf4(t - 1, 0, 2 * db + 1); hpa = *sf4(t - 1, 1, 2 * db); hpb = *sf4(t - 1, 1, 2 * db + 1); }
                        unpack8(uf, f8); unpack8(ui, i8); unpack8(ug, g8); unpack8(uo, o8); unpack8(uq, q8);
                        c8v[0] = ca.x; c8v[1] = ca.y; c8v[2] = ca.z; c8v[3] = ca.w; c8v[4] = cb.x; c8v[5] = cb.y; c8v[6] = cb.z; c8v[7] = cb.w;
                        cp8[0] = pa.x; cp8[1] = pa.y; cp8[2] = pa.z; cp8[3] = pa.w; cp8[4] = pb.x; cp8[5] = pb.y; cp8[6] = pb.z; cp8[7] = pb.w;
                        h8[0] = ha.x; h8[1] = ha.y; h8[2] = ha.z; h8[3] = ha.w; h8[4] = hb.x; h8[5] = hb.y; h8[6] = hb.z; h8[7] = hb.w;
                        hp8[0] = hpa.x; hp8[1] = hpa.y; hp8[2] = hpa.z; hp8[3] = hpa.w; hp8[4] = hpb.x; hp8[5] = hpb.y; hp8[6] = hpb.z; hp8[7] = hpb.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) { f8[e] = i8[e] = g8[e] = o8[e] = q8[e] = c8v[e] = cp8[e] = h8[e] = hp8[e] = 0.0f; }
                    }
                    // gradient of the two rows that only need h_t goes straight to the staging slice
                    *reinterpret_cast<float4*>(stage_b + lane * kSS + (2 * db) * 4) = make_float4(g * h8[0], g * h8[1], g * h8[2], g * h8[3]);
                    *reinterpret_cast<float4*>(stage_b + lane * kSS + (2 * db + 1) * 4) = make_float4(g * h8[4], g * h8[5], g * h8[6], g * h8[7]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int d = db * 8 + e;
                        const float tcv = fast_tanh(c8v[e]);
                        const float dh = dh_rec[d] + q8[e];
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
                    *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, db, 10)) = pack_bf16x8(hp8);      // Z_t = [h_{t-1}, x_t]
                    *reinterpret_cast<uint4*>(Zb + tile_chunk_off(r, 4 + db, 10)) = xraw;               // x_t was saved as bf16
                }
                fence_async_smem();
                tc_fence_before_sync();
                tile_bar(tile);
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
                // Next timestep's activations: the forward stream evicted them to HBM (296 tiles x 4 MB); fetch them into L2
                // now, one MMA + three row visits ahead of their use -- early enough for DRAM latency, late enough that
                // the other tiles' traffic does not evict them again (the whole L2 turns over in about two timesteps).
                if (t >= 1) prefetch_step(t - 1);
                // overlap with the MMAs: the two visits that need only h_t  (t descending: E[neg], E[out], ..)
                __syncwarp();
                coop_update(m, neg, act, lane, stage_b, 1.0f, o);
                coop_update(m, out, act, lane, stage_b, -1.0f, o);
                mbar_wait(mbar + tile, phase); phase ^= 1;
                tc_fence_after_sync();
                float dx[32];
#pragma unroll
                for (int db = 0; db < 4; ++db) {
                    uint32_t ra[8], rb[8];
                    tmem_ld8_issue(tbase + db * 8, ra); tmem_ld8_issue(tbase + 32 + db * 8, rb);
                    tmem_wait8(ra); tmem_wait8(rb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { dh_rec[db * 8 + j] = act ? __uint_as_float(ra[j]) : 0.0f; dx[db * 8 + j] = __uint_as_float(rb[j]); }
                }
                tc_fence_before_sync();
                stage_write_row(stage_b, lane, dx, 1.0f);
                coop_update(m, in, act, lane, stage_b, 1.0f, o);   // .. E[in], b[neg], b[out]
                if (act) thread_update_biases(m, neg, out, g, o);
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
                        fast_adam(w, s1, s2, dwr[k], od);
                        __stcg(m.dense + 2 * nd + idx, s2);
                    } else fast_adagrad(w, s1, dwr[k], od.lr, od.l2);
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

}  // namespace

size_t lstm_tc_smem_bytes(int nt) { return OFF_TILES + (size_t)nt * TILE_BYTES; }

cudaError_t launch_lstm_tc(const ModelDev& m, const PlanDev& p, int nt, cudaStream_t st) {
    const size_t smem = lstm_tc_smem_bytes(nt);
    const int per_cta = 128 * nt;
    dim3 grid((p.P + per_cta - 1) / per_cta);
    cudaError_t e;
    if (nt == 2) {
        e = cudaFuncSetAttribute(lstm_tc_train_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        lstm_tc_train_kernel<2><<<grid, 256, smem, st>>>(m, p);
    } else {
        e = cudaFuncSetAttribute(lstm_tc_train_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        lstm_tc_train_kernel<1><<<grid, 128, smem, st>>>(m, p);
    }
    return cudaGetLastError();
}

}  // namespace sbr
