// lstm_batch.cuh -- round-synchronous LSTM training on batched tensor-core GEMMs (included by sync_engine.cu).
//
// What it is for
//   * Parallelism::Synchronous (mod.rs:39-40, the reference default lstm.rs:66) for the LSTM model, every loss incl. WARP:
//     a round = every partition ("thread", sequence_model.rs:90-98) takes its next sub-sequence; all gradients of the round
//     come from the round-start parameters; the recorded (row, gradient) entries are then applied once each, un-merged, in
//     the reference order; the dense LSTM weights take ONE step on the gradient summed over the round (oracle: run_partition
//     barrier mode).
//   * the throughput path of the wide LSTMs (embedding_dim 64 / 128 / 256: BASELINE configs C3, C5), where the 4-gate
//     product is a real contraction: per timestep the whole round is one [P x 2D] . [2D x 4D] GEMM on the tcgen05 tensor
//     cores (lstm.rs:293-298 -> wyrm Layer::forward), two more in the backward pass.
//
// Layout.  Every GEMM operand lives in HBM as 128 x 128 bf16 TILES of 32 KB in the canonical no-swizzle core-matrix order
// (tc_tile.cuh: chunk (r, c8) at ((r / 8) * 16 + c8) * 128 + (r % 8) * 16), so that one `cp.async.bulk` brings a tile into
// shared memory and the same bytes serve as a K-major operand along the columns and as an MN-major operand along the
// rows -- the three products of a timestep read the tiles the previous kernels wrote, without any transposition:
//     gates_t = Z_t . W            A = Z tile  (seq x feat)  K-major     B = W tile (feat x gate) MN-major
//     dz_t    = delta_t . W^T      A = delta tile (seq x gate) K-major   B = W tile               K-major
//     dW^T   += delta_t^T . Z_t    A = delta tile MN-major               B = Z tile               MN-major   (K = sequences)
// rows = partitions of the round (padded to 128), feat = [h_{t-1} | x_t] (2D, padded to 128), gate columns in blocks of 128:
// block j = hidden units [32 j, 32 j + 32) x {f, i, g, o}, so that a CTA that owns a block owns whole LSTM cells.
//
// Kernels of one round (stream order):
//   bl_begin (x_0, state reset) -> T-1 x { bl_fwd (GEMM + LSTM cell epilogue) ; bl_score (scores, WARP / uniform negative,
//   loss, entries of the target and negative rows, x_{t+1}) } -> T-1 x { bl_delta (gate deltas, cell-gradient recurrence) ;
//   bl_dz (GEMM; epilogue: dh_{t-1}, entry of the input row) } -> bl_dw (split-K GEMM over all timesteps, bias gradient through
//   a ones column) -> keys / radix sort / sync_apply_kernel (sparse visits in the reference order) -> sync_dense_kernel.
#pragma once

namespace {

using namespace tc;

constexpr uint32_t kTileBytes = 32768;

__device__ __forceinline__ void bl_bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar_smem) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar_smem) : "memory");
}
__device__ __forceinline__ void bl_expect_tx(uint32_t bar_smem, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_smem), "r"(bytes) : "memory");
}

struct BatchDims {
    int D, UB, NBLK, KB;        // units per gate block (32, or 16 when D = 16), gate blocks, feature blocks (2D padded to 128, / 128)
    uint32_t P, Ppad, PB;       // partitions, padded to tiles of 128, row blocks
    int Tm1;                    // timesteps per full sub-sequence
};
__host__ __device__ inline BatchDims batch_dims(int D, uint32_t P, int T) {
    BatchDims d;
    d.D = D; d.UB = D < 32 ? D : 32; d.NBLK = (D + 31) / 32; d.KB = (2 * D + 127) / 128;
    d.P = P; d.PB = (P + 127) / 128; d.Ppad = d.PB * 128; d.Tm1 = T - 1;
    return d;
}

struct BatchPtrs {
    uint8_t* Wt;        // [KB][NBLK] tiles
    uint8_t* Z;         // [T][PB][KB] tiles               Z_t = [h_{t-1} | x_t]
    uint8_t* Dl;        // [T-1][PB][NBLK] tiles           gate deltas
    uint4* act;         // [T-1][6][D/8][Ppad] 8 x bf16:   f i g o c tanh(c)
    float* dQ;          // [T-1][Ppad][D]                  g (q - p)
    float* Hrow;        // [Ppad][D]                       h_t of the current timestep
    float4* C;          // [D/4][Ppad]                     cell state (forward) / cell gradient (backward)
    float4* dzh;        // [D/4][Ppad]                     dh_{t-1} out of dz_t
    float* grads;       // [P (T-1) 3][D]                  sparse entries: slot = (p (T-1) + t) 3 + {0 in, 1 out, 2 neg}
    float* bgrads;      // [P (T-1) 3]
    uint32_t* neg_id;   // [P (T-1)]
    int* Tn;            // [Ppad]                          timesteps of this round's sub-sequence (0 for padding rows)
    uint64_t* base;     // [Ppad]                          offset of the sub-sequence in item_ids
    float* dWsum;       // [ndense]                        round-summed dense gradient (canonical layout)
    uint8_t* ones;      // 128 x 16 bf16 ones (bias gradient column)
};

// ------------------------------------------------------------------------------------------------ small kernels
// fp32 canonical weights W[k][gate][u] -> bf16 tiles with the block-of-cells column order
__global__ void bl_stage_w_kernel(ModelDev m, BatchDims bd, BatchPtrs bp) {
    const int D = bd.D, NG = bd.NBLK * 128, KZ = bd.KB * 128;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < KZ * NG; i += gridDim.x * blockDim.x) {
        const int k = i / NG, n = i % NG;
        const int j = n >> 7, nl = n & 127, g = nl / bd.UB, u = j * 32 + nl % bd.UB;
        float v = 0.0f;
        if (k < 2 * D && g < 4 && u < D) v = m.dense[(size_t)k * 4 * D + g * D + u];
        uint8_t* tile = bp.Wt + ((size_t)(k >> 7) * bd.NBLK + j) * kTileBytes;
        *reinterpret_cast<__nv_bfloat16*>(tile + tile_chunk_off(k & 127, nl >> 3, 16) + (nl & 7) * 2) = __float2bfloat16_rn(v);
    }
}

// warp per partition: this round's sub-sequence, x_0 into Z_0, zero state
template <int D>
__global__ void __launch_bounds__(256) bl_begin_kernel(ModelDev m, PlanDev pl, BatchDims bd, BatchPtrs bp, uint32_t it) {
    constexpr int V = VecOf<D>::V;
    const int lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= bd.Ppad) return;
    int Tn = 0; uint64_t base = 0;
    if (p < bd.P) {
        const uint32_t sq = pl.order[(size_t)p * pl.n + it];
        Tn = (int)pl.seq_len[sq] - 1; base = pl.seq_start[sq];
    }
    if (lane == 0) { bp.Tn[p] = Tn; bp.base[p] = base; }
    float x[V];
#pragma unroll
    for (int v = 0; v < V; ++v) x[v] = 0.0f;
    if (Tn > 0) row_load_cg<D>(item_rec(m, pl.item_ids[base]), lane, x);
    uint8_t* zrow = bp.Z + ((size_t)(p >> 7) * bd.KB) * kTileBytes;   // Z_0, row block of p
    const int r = p & 127;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int u = D < 32 ? lane : (V <= 4 ? lane * V + v : (v / 4) * 128 + lane * 4 + (v % 4));
        if (D >= 32 || lane < D) {
            const int kx = D + u;
            *reinterpret_cast<__nv_bfloat16*>(zrow + (size_t)(kx >> 7) * kTileBytes + tile_chunk_off(r, (kx & 127) >> 3, 16) + (kx & 7) * 2) = __float2bfloat16_rn(x[v]);
            *reinterpret_cast<__nv_bfloat16*>(zrow + (size_t)(u >> 7) * kTileBytes + tile_chunk_off(r, (u & 127) >> 3, 16) + (u & 7) * 2) = __float2bfloat16_rn(0.0f);
        }
    }
    for (int c = lane; c < D / 4; c += 32) bp.C[(size_t)c * bd.Ppad + p] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// warp per partition, timestep t: pos / neg scores on h_t, WARP or uniform negative (sequence_model.rs:47-68,115-142), loss
// (lstm.rs:300-328), the entries of the target and the negative row, g (q - p) for the backward pass, x_{t+1} = E[out_t]
template <int D>
__global__ void __launch_bounds__(256) bl_score_kernel(ModelDev m, PlanDev pl, BatchDims bd, BatchPtrs bp, int t) {
    constexpr int V = VecOf<D>::V;
    const int lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= bd.P) return;
    const int Tn = bp.Tn[p];
    const bool act = t < Tn;
    const int r = p & 127;
    uint8_t* znext = bp.Z + (((size_t)(t + 1) * bd.PB + (p >> 7)) * bd.KB) * kTileBytes;
    float pv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) pv[v] = 0.0f;
    if (act) {
        const uint32_t* ids = pl.item_ids + bp.base[p];
        const uint32_t out = __ldg(ids + t + 1);
        float h[V], qv[V];
        vec_load<D>(bp.Hrow + (size_t)p * D, lane, h);
        row_load_cg<D>(item_rec(m, out), lane, pv);
        const float pos = warp_dot<D>(h, pv) + __ldcg(reinterpret_cast<const float*>(bias_rec(m, out)));
        const uint64_t key = pl.keys[p], step = pl.step_ctr[p];
        const int tries = m.loss == 2 ? 5 : 1;
        uint32_t neg = 0; float ngs = 0.0f;
        for (int j = 0; j < tries; ++j) {
            neg = draw_item(key, step, (uint32_t)t, (uint32_t)j, pl.neg_range);
            row_load_cg<D>(item_rec(m, neg), lane, qv);
            ngs = warp_dot<D>(h, qv) + __ldcg(reinterpret_cast<const float*>(bias_rec(m, neg)));
            if (1.0f - pos + ngs > 0.0f) break;
        }
        float l, g;
        if (m.loss == 0) { const float s = sigmoidf_(ngs - pos); l = s; g = s * (1.0f - s); }
        else { const float vv = 1.0f + ngs - pos; l = vv > 0.0f ? vv : 0.0f; g = vv > 0.0f ? 1.0f : 0.0f; }
        const size_t slot = ((size_t)p * bd.Tm1 + t) * 3;
        float gn[V], gp[V], dq[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { gn[v] = g * h[v]; gp[v] = -g * h[v]; dq[v] = g * (qv[v] - pv[v]); }
        vec_store<D>(bp.grads + (slot + 2) * D, lane, gn);
        vec_store<D>(bp.grads + (slot + 1) * D, lane, gp);
        vec_store<D>(bp.dQ + ((size_t)t * bd.Ppad + p) * D, lane, dq);
        if (lane == 0) {
            bp.bgrads[slot + 2] = g; bp.bgrads[slot + 1] = -g; bp.bgrads[slot] = __int_as_float(0x7fc00000);   // NaN: inputs have no bias entry
            bp.neg_id[(size_t)p * bd.Tm1 + t] = neg;
            pl.loss_acc[p] += l;
        }
    }
    // x_{t+1} (zero once the sub-sequence has ended: rows of finished sequences stay finite and their deltas are zero)
    if (t + 1 < bd.Tm1 + 1) {
        const bool nx = t + 1 < Tn;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int u = D < 32 ? lane : (V <= 4 ? lane * V + v : (v / 4) * 128 + lane * 4 + (v % 4));
            if (D >= 32 || lane < D) {
                const int kx = D + u;
                *reinterpret_cast<__nv_bfloat16*>(znext + (size_t)(kx >> 7) * kTileBytes + tile_chunk_off(r, (kx & 127) >> 3, 16) + (kx & 7) * 2) =
                    __float2bfloat16_rn(nx ? pv[v] : 0.0f);
            }
        }
    }
}

__device__ __forceinline__ float bl_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bl_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void bl_unpack8(const uint4& u, float (&v)[8]) {
    v[0] = bl_lo(u.x); v[1] = bl_hi(u.x); v[2] = bl_lo(u.y); v[3] = bl_hi(u.y); v[4] = bl_lo(u.z); v[5] = bl_hi(u.z); v[6] = bl_lo(u.w); v[7] = bl_hi(u.w);
}

// thread per (row, 8-unit chunk), timestep t: gate deltas -> delta tile, cell-gradient recurrence in place
__global__ void __launch_bounds__(256) bl_delta_kernel(BatchDims bd, BatchPtrs bp, int t, int coupled, int first) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nch = bd.D / 8;
    if (idx >= (size_t)bd.Ppad * nch) return;
    const uint32_t r = (uint32_t)(idx % bd.Ppad); const int c = (int)(idx / bd.Ppad);
    const bool act = t < bp.Tn[r];
    float df[8], di[8], dg[8], d_o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { df[e] = 0.f; di[e] = 0.f; dg[e] = 0.f; d_o[e] = 0.f; }
    float4 dca = bp.C[(size_t)(2 * c) * bd.Ppad + r], dcb = bp.C[(size_t)(2 * c + 1) * bd.Ppad + r];
    if (first) { dca = make_float4(0.f, 0.f, 0.f, 0.f); dcb = dca; }
    float dcv[8] = {dca.x, dca.y, dca.z, dca.w, dcb.x, dcb.y, dcb.z, dcb.w};
    if (act) {
        const uint4* a = bp.act + ((size_t)t * 6 * nch + c) * bd.Ppad + r;
        const size_t gs = (size_t)nch * bd.Ppad;
        float f[8], ig[8], gg[8], og[8], tc[8], cp[8];
        bl_unpack8(a[0], f); bl_unpack8(a[gs], ig); bl_unpack8(a[2 * gs], gg); bl_unpack8(a[3 * gs], og); bl_unpack8(a[5 * gs], tc);
        if (t > 0) bl_unpack8((a - (size_t)6 * nch * bd.Ppad)[4 * gs], cp);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) cp[e] = 0.0f;
        }
        const float* dq = bp.dQ + ((size_t)t * bd.Ppad + r) * bd.D + 8 * c;
        const float4 q0 = *reinterpret_cast<const float4*>(dq), q1 = *reinterpret_cast<const float4*>(dq + 4);
        float dh[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        if (!first) {   // + dh out of dz_{t+1} (zero rows for sequences that had already ended)
            const float4 h0 = bp.dzh[(size_t)(2 * c) * bd.Ppad + r], h1 = bp.dzh[(size_t)(2 * c + 1) * bd.Ppad + r];
            dh[0] += h0.x; dh[1] += h0.y; dh[2] += h0.z; dh[3] += h0.w; dh[4] += h1.x; dh[5] += h1.y; dh[6] += h1.z; dh[7] += h1.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float dho = dh[e] * tc[e];
            const float dc = dh[e] * og[e] * (1.0f - tc[e] * tc[e]) + dcv[e];
            float dff = dc * cp[e], dii = dc * gg[e];
            const float dgg = dc * ig[e];
            dcv[e] = dc * f[e];
            if (coupled) { dff -= dii; dii = 0.0f; }
            df[e] = dff * f[e] * (1.0f - f[e]);
            di[e] = coupled ? 0.0f : dii * ig[e] * (1.0f - ig[e]);
            dg[e] = dgg * (1.0f - gg[e] * gg[e]);
            d_o[e] = dho * og[e] * (1.0f - og[e]);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) dcv[e] = 0.0f;
    }
    bp.C[(size_t)(2 * c) * bd.Ppad + r] = make_float4(dcv[0], dcv[1], dcv[2], dcv[3]);
    bp.C[(size_t)(2 * c + 1) * bd.Ppad + r] = make_float4(dcv[4], dcv[5], dcv[6], dcv[7]);
    const int u0 = 8 * c, j = u0 >> 5, cu = (u0 & 31) >> 3, cpg = bd.UB / 8;
    uint8_t* tile = bp.Dl + (((size_t)t * bd.PB + (r >> 7)) * bd.NBLK + j) * kTileBytes;
    const int rr = r & 127;
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(rr, 0 * cpg + cu, 16)) = pack_bf16x8(df);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(rr, 1 * cpg + cu, 16)) = pack_bf16x8(di);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(rr, 2 * cpg + cu, 16)) = pack_bf16x8(dg);
    *reinterpret_cast<uint4*>(tile + tile_chunk_off(rr, 3 * cpg + cu, 16)) = pack_bf16x8(d_o);
}

// sort keys of the round's entries: (row, position in the reference's application order); unused slots sort to the end
__global__ void bl_keys_kernel(PlanDev pl, BatchDims bd, BatchPtrs bp, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const size_t total = (size_t)bd.P * bd.Tm1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t p = (uint32_t)(i / bd.Tm1); const int t = (int)(i - (size_t)p * bd.Tm1);
        unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull;
        if (t < bp.Tn[p]) {
            const uint32_t* ids = pl.item_ids + bp.base[p];
            const uint32_t ord = (p << 14) | ((uint32_t)(bd.Tm1 - 1 - t) << 2);   // partition (18 bits) | t descending (12 bits) | neg, out, in
            k0 = ((unsigned long long)__ldg(ids + t) << 32) | (ord | 2u);
            k1 = ((unsigned long long)__ldg(ids + t + 1) << 32) | (ord | 1u);
            k2 = ((unsigned long long)bp.neg_id[i] << 32) | (ord | 0u);
        }
        keys[3 * i] = k0; keys[3 * i + 1] = k1; keys[3 * i + 2] = k2;
        vals[3 * i] = (uint32_t)(3 * i); vals[3 * i + 1] = (uint32_t)(3 * i + 1); vals[3 * i + 2] = (uint32_t)(3 * i + 2);
    }
}

__global__ void bl_finish_kernel(PlanDev pl, BatchDims bd, BatchPtrs bp) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= bd.P) return;
    pl.examples[p] += (unsigned long long)bp.Tn[p];
    pl.step_ctr[p] += 1;
}

// ------------------------------------------------------------------------------------------------ GEMM kernels
// One CTA = 256 threads: thread 0 streams tiles (cp.async.bulk -> 2-stage ring, mbarrier transaction counts), thread 32 issues
// the tcgen05 MMAs (a stage is released by tcgen05.commit when the MMAs that read it have finished), then all eight warps run
// the epilogue out of TMEM: warp w reads lanes 32 (w % 4) .., warps w and w + 4 split the columns.
// With a single k-iteration (embedding_dim <= 64 forward) one stage is enough and two CTAs share an SM: the epilogue of one
// overlaps the loads of the other.
constexpr int kGemmThreads = 256;
constexpr uint32_t kStageBytes = 3 * kTileBytes;            // A tile + two B tiles
__host__ __device__ constexpr uint32_t gemm_smem_bytes(int nstage) { return (uint32_t)nstage * kStageBytes + 4096 + 1024; }   // + ones tile + barriers

struct GemmSmem {
    uint8_t* stage[2];
    uint8_t* ones;
    uint64_t* full; uint64_t* empty; uint64_t* done; uint32_t* tmem_ptr;
    int nstage;
};
__device__ __forceinline__ GemmSmem gemm_smem_carve(uint8_t* smem, int nstage) {
    GemmSmem s;
    s.nstage = nstage;
    s.stage[0] = smem; s.stage[1] = smem + (nstage - 1) * kStageBytes; s.ones = smem + nstage * kStageBytes;
    uint8_t* misc = smem + nstage * kStageBytes + 4096;
    s.full = reinterpret_cast<uint64_t*>(misc); s.empty = s.full + 2; s.done = s.full + 4;
    s.tmem_ptr = reinterpret_cast<uint32_t*>(misc + 64);
    return s;
}
template <int COLS>
__device__ __forceinline__ void gemm_setup(const GemmSmem& s) {
    if (threadIdx.x < 32) tmem_alloc<COLS>(s.tmem_ptr);
    if (threadIdx.x == 0) {
        mbar_init(s.full, 1); mbar_init(s.full + 1, 1); mbar_init(s.empty, 1); mbar_init(s.empty + 1, 1); mbar_init(s.done, 1);
        fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
}
template <int COLS>
__device__ __forceinline__ void gemm_teardown(const GemmSmem& s) {
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<COLS>(*s.tmem_ptr);
}

// MODE 0: A K-major, B MN-major (forward)   1: A K-major, B K-major (dz)   2: A MN-major, B MN-major (dW)
// tile_of(iter, &A, &B0, &B1): global tile addresses of k-iteration `iter` (B1 == nullptr: one B tile)
template <int MODE, class TileFn>
__device__ __forceinline__ void gemm_mainloop(const GemmSmem& s, int niter, TileFn tile_of, bool with_ones) {
    const int tid = threadIdx.x;
    const int sh = s.nstage - 1;   // 1 or 2 stages: stage = it & sh, use count = it >> sh
    if (tid == 0) {
        for (int it = 0; it < niter; ++it) {
            const int st = it & sh;
            if (it >= s.nstage) mbar_wait(s.empty + st, ((it >> sh) - 1) & 1);
            const uint8_t *A, *B0, *B1;
            tile_of(it, A, B0, B1);
            const uint32_t bar = smem_u32(s.full + st);
            bl_expect_tx(bar, (B1 ? 3u : 2u) * kTileBytes);
            bl_bulk_load(smem_u32(s.stage[st]), A, kTileBytes, bar);
            bl_bulk_load(smem_u32(s.stage[st]) + kTileBytes, B0, kTileBytes, bar);
            if (B1) bl_bulk_load(smem_u32(s.stage[st]) + 2 * kTileBytes, B1, kTileBytes, bar);
        }
    } else if (tid == 32) {
        const uint32_t tm = *s.tmem_ptr;
        constexpr uint32_t idesc = make_idesc_bf16(128, 128, MODE == 2 ? 1 : 0, MODE == 1 ? 0 : 1);
        constexpr uint32_t idesc1 = make_idesc_bf16(128, 16, 1, 1);
        for (int it = 0; it < niter; ++it) {
            const int st = it & sh;
            mbar_wait(s.full + st, (it >> sh) & 1);
            tc_fence_after_sync();
            const uint8_t *A, *B0, *B1;
            tile_of(it, A, B0, B1);
            const uint32_t a = smem_u32(s.stage[st]), b0 = a + kTileBytes, b1 = a + 2 * kTileBytes;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint64_t ad = MODE == 2 ? make_smem_desc(a + k * 4096, 2048, 128) : make_smem_desc(a + k * 256, 128, 2048);
                const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
                mma_bf16(tm, ad, MODE == 1 ? make_smem_desc(b0 + k * 256, 128, 2048) : make_smem_desc(b0 + k * 4096, 2048, 128), idesc, acc);
                if (B1) mma_bf16(tm + 128, ad, MODE == 1 ? make_smem_desc(b1 + k * 256, 128, 2048) : make_smem_desc(b1 + k * 4096, 2048, 128), idesc, acc);
                if (MODE == 2 && with_ones) mma_bf16(tm + 256, ad, make_smem_desc(smem_u32(s.ones) + k * 512, 256, 128), idesc1, acc);
            }
            mma_commit(s.empty + st);
        }
        mma_commit(s.done);
    }
    __syncwarp();
    mbar_wait(s.done, 0);
    tc_fence_after_sync();
}

// gates_t = Z_t . W (+ bias), LSTM cell (lstm.rs:293-298): CTA = row block x two gate blocks (64 cells)
__device__ __forceinline__ float bl_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float bl_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / (1 + 2^(-x log2 e)) and 2 / (1 + 2^(-2 x log2 e)) - 1 on the MUFU units (saturate cleanly at both ends)
__device__ __forceinline__ float bl_sigm(float x) { return bl_rcp(1.0f + bl_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float bl_tanh(float x) { return fmaf(2.0f, bl_rcp(1.0f + bl_ex2(-2.8853900817779268f * x)), -1.0f); }

__global__ void __launch_bounds__(kGemmThreads, 2) bl_fwd_kernel(ModelDev m, BatchDims bd, BatchPtrs bp, int t) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GemmSmem s = gemm_smem_carve(smem, bd.KB > 1 ? 2 : 1);
    gemm_setup<256>(s);
    const int pb = blockIdx.x, j0 = blockIdx.y * 2, nj = min(2, bd.NBLK - j0);
    const uint8_t* zt = bp.Z + ((size_t)t * bd.PB + pb) * bd.KB * kTileBytes;
    gemm_mainloop<0>(s, bd.KB, [&](int it, const uint8_t*& A, const uint8_t*& B0, const uint8_t*& B1) {
        A = zt + (size_t)it * kTileBytes;
        B0 = bp.Wt + ((size_t)it * bd.NBLK + j0) * kTileBytes;
        B1 = nj > 1 ? B0 + kTileBytes : nullptr;
    }, false);
    // ---- epilogue: thread = (row, every second 8-cell chunk of the CTA's blocks) ----
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, part = warp >> 2;
    const int rr = (warp & 3) * 32 + lane;
    const uint32_t r = (uint32_t)pb * 128 + rr;
    const uint32_t tl = *s.tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
    const bool act = t < bp.Tn[r];
    const bool coupled = m.variant == 1;
    const int D = bd.D, cpg = bd.UB / 8, nch = D / 8;
    const float* bias = m.dense + (size_t)2 * D * 4 * D;
    uint8_t* znext = bp.Z + ((size_t)(t + 1) * bd.PB + pb) * bd.KB * kTileBytes;
    for (int c = part; c < nj * cpg; c += 2) {
        const int jl = c / cpg, cu = c % cpg;
        const int u0 = (j0 + jl) * 32 + cu * 8;
        const uint32_t cb = tl + jl * 128 + cu * 8;
        float pf[8], pi[8], pg[8], po[8], pc[8], ptc[8], hn[8];
        const float4 c0 = bp.C[(size_t)(u0 / 4) * bd.Ppad + r], c1 = bp.C[(size_t)(u0 / 4 + 1) * bd.Ppad + r];
        tmem_ld8x4(cb, cb + bd.UB, cb + 2 * bd.UB, cb + 3 * bd.UB, pf, pi, pg, po);
        const float cprev[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        float bf[8], bi[8], bg[8], bo[8];
#pragma unroll
        for (int h4 = 0; h4 < 2; ++h4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(bias + u0) + h4), b = __ldg(reinterpret_cast<const float4*>(bias + D + u0) + h4);
            const float4 cc = __ldg(reinterpret_cast<const float4*>(bias + 2 * D + u0) + h4), d = __ldg(reinterpret_cast<const float4*>(bias + 3 * D + u0) + h4);
            bf[4 * h4] = a.x; bf[4 * h4 + 1] = a.y; bf[4 * h4 + 2] = a.z; bf[4 * h4 + 3] = a.w;
            bi[4 * h4] = b.x; bi[4 * h4 + 1] = b.y; bi[4 * h4 + 2] = b.z; bi[4 * h4 + 3] = b.w;
            bg[4 * h4] = cc.x; bg[4 * h4 + 1] = cc.y; bg[4 * h4 + 2] = cc.z; bg[4 * h4 + 3] = cc.w;
            bo[4 * h4] = d.x; bo[4 * h4 + 1] = d.y; bo[4 * h4 + 2] = d.z; bo[4 * h4 + 3] = d.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float f = bl_sigm(pf[e] + bf[e]);
            const float ig = coupled ? 1.0f - f : bl_sigm(pi[e] + bi[e]);
            const float gg = bl_tanh(pg[e] + bg[e]);
            const float og = bl_sigm(po[e] + bo[e]);
            const float cn = f * cprev[e] + ig * gg;
            const float tcn = bl_tanh(cn);
            pf[e] = f; pi[e] = ig; pg[e] = gg; po[e] = og; pc[e] = act ? cn : 0.0f; ptc[e] = tcn; hn[e] = act ? og * tcn : 0.0f;
        }
        bp.C[(size_t)(u0 / 4) * bd.Ppad + r] = make_float4(pc[0], pc[1], pc[2], pc[3]);
        bp.C[(size_t)(u0 / 4 + 1) * bd.Ppad + r] = make_float4(pc[4], pc[5], pc[6], pc[7]);
        *reinterpret_cast<uint4*>(znext + (size_t)(u0 >> 7) * kTileBytes + tile_chunk_off(rr, (u0 & 127) >> 3, 16)) = pack_bf16x8(hn);
        float* hrow = bp.Hrow + (size_t)r * D + u0;
        *reinterpret_cast<float4*>(hrow) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(hrow + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
        if (act) {
            uint4* a = bp.act + ((size_t)t * 6 * nch + (u0 >> 3)) * bd.Ppad + r;
            const size_t gs = (size_t)nch * bd.Ppad;
            a[0] = pack_bf16x8(pf); a[gs] = pack_bf16x8(pi); a[2 * gs] = pack_bf16x8(pg); a[3 * gs] = pack_bf16x8(po);
            a[4 * gs] = pack_bf16x8(pc); a[5 * gs] = pack_bf16x8(ptc);
        }
    }
    gemm_teardown<256>(s);
}

// dz_t = delta_t . W^T: CTA = row block x two feature blocks; columns [0, D) -> dh_{t-1}, [D, 2D) -> entry of the input row
__global__ void __launch_bounds__(kGemmThreads, 2) bl_dz_kernel(BatchDims bd, BatchPtrs bp, int t) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GemmSmem s = gemm_smem_carve(smem, bd.NBLK > 1 ? 2 : 1);
    gemm_setup<256>(s);
    const int pb = blockIdx.x, k0 = blockIdx.y * 2, nk = min(2, bd.KB - k0);
    const uint8_t* dt = bp.Dl + ((size_t)t * bd.PB + pb) * bd.NBLK * kTileBytes;
    gemm_mainloop<1>(s, bd.NBLK, [&](int it, const uint8_t*& A, const uint8_t*& B0, const uint8_t*& B1) {
        A = dt + (size_t)it * kTileBytes;
        B0 = bp.Wt + ((size_t)k0 * bd.NBLK + it) * kTileBytes;
        B1 = nk > 1 ? B0 + (size_t)bd.NBLK * kTileBytes : nullptr;
    }, false);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, part = warp >> 2;
    const int rr = (warp & 3) * 32 + lane;
    const uint32_t r = (uint32_t)pb * 128 + rr;
    const uint32_t tl = *s.tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
    const int D = bd.D;
    const bool act = r < bd.P && t < bp.Tn[r];
    float* gin = bp.grads + (((size_t)r * bd.Tm1 + t) * 3) * D;
    for (int c = part; c < nk * 16; c += 2) {   // 8 feature columns per step
        const int k = k0 * 128 + c * 8;
        if (k >= 2 * D) break;
        float v[8];
        tmem_ld8(tl + c * 8, v);
        if (k < D) {
            bp.dzh[(size_t)(k / 4) * bd.Ppad + r] = make_float4(v[0], v[1], v[2], v[3]);
            bp.dzh[(size_t)(k / 4 + 1) * bd.Ppad + r] = make_float4(v[4], v[5], v[6], v[7]);
        } else if (act) {
            *reinterpret_cast<float4*>(gin + (k - D)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(gin + (k - D) + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    gemm_teardown<256>(s);
}

// dW^T[gate block][features] += sum over (timestep, row block) delta^T . Z, bias gradient through a column of ones;
// CTA = gate block x two feature blocks x a slice of the (t, row block) pairs; reduce-adds into the canonical dense layout
__global__ void __launch_bounds__(kGemmThreads, 1) bl_dw_kernel(BatchDims bd, BatchPtrs bp, int nsplit) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GemmSmem s = gemm_smem_carve(smem, 2);
    for (int i = threadIdx.x; i < 4096 / 16; i += blockDim.x) reinterpret_cast<uint4*>(s.ones)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_async_smem();
    gemm_setup<512>(s);
    const int j = blockIdx.x, k0 = blockIdx.y * 2, nk = min(2, bd.KB - k0), sp = blockIdx.z;
    const int total = bd.Tm1 * (int)bd.PB;
    const int lo = (int)((long long)total * sp / nsplit), hi = (int)((long long)total * (sp + 1) / nsplit);
    const bool ones = blockIdx.y == 0;
    if (hi > lo) {
        gemm_mainloop<2>(s, hi - lo, [&](int it, const uint8_t*& A, const uint8_t*& B0, const uint8_t*& B1) {
            const int q = lo + it;            // (t, pb) pair: tiles of Z and delta are stored [t][pb][..]
            A = bp.Dl + ((size_t)q * bd.NBLK + j) * kTileBytes;
            B0 = bp.Z + ((size_t)q * bd.KB + k0) * kTileBytes;
            B1 = nk > 1 ? B0 + kTileBytes : nullptr;
        }, ones);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, part = warp >> 2;
        const int nl = (warp & 3) * 32 + lane;   // gate column of the block = TMEM lane
        const uint32_t tl = *s.tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
        const int D = bd.D, g = nl / bd.UB, u = j * 32 + nl % bd.UB;
        const bool valid = g < 4 && u < D;
        for (int c = part; c < nk * 16; c += 2) {
            const int k = k0 * 128 + c * 8;
            if (k >= 2 * D) break;
            float v[8];
            tmem_ld8(tl + c * 8, v);
            if (valid) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(bp.dWsum + (size_t)(k + e) * 4 * D + g * D + u, v[e]);
            }
        }
        if (ones && part == 0) {
            float v[8];
            tmem_ld8(tl + 256, v);
            if (valid) atomicAdd(bp.dWsum + (size_t)2 * D * 4 * D + g * D + u, v[0]);
        }
    }
    gemm_teardown<512>(s);
}

}  // namespace

struct BatchBuffers {
    Buf Wt, Z, Dl, act, dQ, Hrow, C, dzh, grads, bgrads, neg_id, Tn, base, dWsum, ones, keys_in, keys_out, vals_in, vals_out, cub_tmp;
    bool primed = false;
};

bool batch_lstm_supported(const ModelDev& m, uint32_t P, const char** why) {
    if (m.model != MODEL_LSTM) { *why = "not an LSTM model"; return false; }
    if (m.D != 16 && m.D != 32 && m.D != 64 && m.D != 128 && m.D != 256) { *why = "embedding_dim must be 16, 32, 64, 128 or 256"; return false; }
    if (m.gmask != 0) { *why = "the batched LSTM engine needs an unsharded item table"; return false; }
    if (P >= (1u << 18) || m.T > 4096) { *why = "the batched LSTM engine takes fewer than 262144 partitions and sequences of at most 4096 items"; return false; }
    return true;
}

#define BL_DISPATCH_D(D_, ...)                                   \
    switch (D_) {                                                \
        case 16: { constexpr int kD = 16; __VA_ARGS__; } break;  \
        case 32: { constexpr int kD = 32; __VA_ARGS__; } break;  \
        case 64: { constexpr int kD = 64; __VA_ARGS__; } break;  \
        case 128: { constexpr int kD = 128; __VA_ARGS__; } break;\
        case 256: { constexpr int kD = 256; __VA_ARGS__; } break;\
        default: break;                                          \
    }

// returns 0 ok, 1 cuda error
int run_batch_lstm(const ModelDev& m, PlanDev& pl, BatchBuffers& B, uint64_t num_updates, int num_sms, cudaStream_t st, int* launches,
                   uint64_t* rounds_out, std::string* err) {
    const int D = m.D, T = m.T;
    const BatchDims bd = batch_dims(D, pl.P, T);
    const size_t nslots = (size_t)bd.P * bd.Tm1 * 3;
    const size_t zbytes = (size_t)T * bd.PB * bd.KB * kTileBytes, dbytes = (size_t)bd.Tm1 * bd.PB * bd.NBLK * kTileBytes;
    SCU(B.Wt.ensure((size_t)bd.KB * bd.NBLK * kTileBytes));
    const bool fresh = B.Z.cap < zbytes || B.Dl.cap < dbytes || !B.primed;
    SCU(B.Z.ensure(zbytes)); SCU(B.Dl.ensure(dbytes));
    SCU(B.act.ensure((size_t)bd.Tm1 * 6 * (D / 8) * bd.Ppad * 16));
    SCU(B.dQ.ensure((size_t)bd.Tm1 * bd.Ppad * D * 4));
    SCU(B.Hrow.ensure((size_t)bd.Ppad * D * 4)); SCU(B.C.ensure((size_t)bd.Ppad * D * 4)); SCU(B.dzh.ensure((size_t)bd.Ppad * D * 4));
    SCU(B.grads.ensure(nslots * D * 4)); SCU(B.bgrads.ensure(nslots * 4)); SCU(B.neg_id.ensure((size_t)bd.P * bd.Tm1 * 4));
    SCU(B.Tn.ensure((size_t)bd.Ppad * 4)); SCU(B.base.ensure((size_t)bd.Ppad * 8)); SCU(B.dWsum.ensure(m.ndense * 4));
    SCU(B.keys_in.ensure(nslots * 8)); SCU(B.keys_out.ensure(nslots * 8)); SCU(B.vals_in.ensure(nslots * 4)); SCU(B.vals_out.ensure(nslots * 4));
    size_t cub_bytes = 0;
    SCU(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<unsigned long long*>(nullptr), static_cast<unsigned long long*>(nullptr),
                                        static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), (int)nslots, 0, 64, st));
    SCU(B.cub_tmp.ensure(cub_bytes));
    if (fresh) {   // padding columns / rows of the operand tiles are written once and never again
        SCU(cudaMemsetAsync(B.Z.p, 0, zbytes, st)); SCU(cudaMemsetAsync(B.Dl.p, 0, dbytes, st));
        B.primed = true;
    }
    SCU(cudaMemsetAsync(B.dWsum.p, 0, m.ndense * 4, st));
    BatchPtrs bp;
    bp.Wt = static_cast<uint8_t*>(B.Wt.p); bp.Z = static_cast<uint8_t*>(B.Z.p); bp.Dl = static_cast<uint8_t*>(B.Dl.p);
    bp.act = static_cast<uint4*>(B.act.p); bp.dQ = static_cast<float*>(B.dQ.p); bp.Hrow = static_cast<float*>(B.Hrow.p);
    bp.C = static_cast<float4*>(B.C.p); bp.dzh = static_cast<float4*>(B.dzh.p); bp.grads = static_cast<float*>(B.grads.p);
    bp.bgrads = static_cast<float*>(B.bgrads.p); bp.neg_id = static_cast<uint32_t*>(B.neg_id.p); bp.Tn = static_cast<int*>(B.Tn.p);
    bp.base = static_cast<uint64_t*>(B.base.p); bp.dWsum = static_cast<float*>(B.dWsum.p); bp.ones = nullptr;
    const uint32_t smem_fwd = gemm_smem_bytes(bd.KB > 1 ? 2 : 1), smem_dz = gemm_smem_bytes(bd.NBLK > 1 ? 2 : 1), smem_dw = gemm_smem_bytes(2);
    SCU(cudaFuncSetAttribute(bl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fwd));
    SCU(cudaFuncSetAttribute(bl_dz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dz));
    SCU(cudaFuncSetAttribute(bl_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dw));
    OptCfg o; o.lr = m.lr; o.l2 = m.l2; o.adam = m.opt == 1; o.c1 = 1.0f; o.c2 = 1.0f;
    const int warps_grid = (int)((bd.Ppad + 7) / 8);
    const dim3 g_fwd(bd.PB, (bd.NBLK + 1) / 2), g_dz(bd.PB, (bd.KB + 1) / 2);
    const int dw_tiles = bd.NBLK * ((bd.KB + 1) / 2);
    const int total_pairs = bd.Tm1 * (int)bd.PB;
    const int nsplit = std::max(1, std::min(total_pairs, (2 * num_sms + dw_tiles - 1) / dw_tiles));
    const dim3 g_dw(bd.NBLK, (bd.KB + 1) / 2, nsplit);
    const unsigned delta_blocks = (unsigned)(((size_t)bd.Ppad * (D / 8) + 255) / 256);
    unsigned long long* k_in = static_cast<unsigned long long*>(B.keys_in.p); unsigned long long* k_out = static_cast<unsigned long long*>(B.keys_out.p);
    uint32_t* v_in = static_cast<uint32_t*>(B.vals_in.p); uint32_t* v_out = static_cast<uint32_t*>(B.vals_out.p);
    uint64_t rounds_done = 0;
    *rounds_out = (uint64_t)pl.n * (uint64_t)pl.epochs;
    int key_bits = 33;   // radix passes only over the bits in use: 32 order bits + the row id (unused slots are all ones: still last)
    while (key_bits < 64 && (m.N >> (key_bits - 32)) != 0) ++key_bits;
    for (int ep = 0; ep < pl.epochs; ++ep) {
        sync_shuffle_kernel<<<(pl.P + 127) / 128, 128, 0, st>>>(pl);
        ++*launches;
        for (uint32_t it = 0; it < pl.n; ++it, ++rounds_done) {
            bl_stage_w_kernel<<<148, 256, 0, st>>>(m, bd, bp);
            BL_DISPATCH_D(D, bl_begin_kernel<kD><<<warps_grid, 256, 0, st>>>(m, pl, bd, bp, it));
            *launches += 2;
            for (int t = 0; t < bd.Tm1; ++t) {
                bl_fwd_kernel<<<g_fwd, kGemmThreads, smem_fwd, st>>>(m, bd, bp, t);
                BL_DISPATCH_D(D, bl_score_kernel<kD><<<warps_grid, 256, 0, st>>>(m, pl, bd, bp, t));
                *launches += 2;
            }
            for (int t = bd.Tm1 - 1; t >= 0; --t) {
                bl_delta_kernel<<<delta_blocks, 256, 0, st>>>(bd, bp, t, m.variant == 1 ? 1 : 0, t == bd.Tm1 - 1 ? 1 : 0);
                bl_dz_kernel<<<g_dz, kGemmThreads, smem_dz, st>>>(bd, bp, t);
                *launches += 2;
            }
            bl_dw_kernel<<<g_dw, kGemmThreads, smem_dw, st>>>(bd, bp, nsplit);
            bl_keys_kernel<<<148 * 4, 256, 0, st>>>(pl, bd, bp, k_in, v_in);
            size_t tmp = B.cub_tmp.cap;
            SCU(cub::DeviceRadixSort::SortPairs(B.cub_tmp.p, tmp, k_in, k_out, v_in, v_out, (int)nslots, 0, key_bits, st));
            const uint64_t t_adam = num_updates + (rounds_done + 1) * (uint64_t)pl.P;
            if (o.adam) { o.c1 = 1.0f - powf(0.9f, (float)t_adam); o.c2 = 1.0f - powf(0.999f, (float)t_adam); }
            SYNC_DISPATCH_D(D, sync_apply_kernel<kD><<<148 * 8, 256, 0, st>>>(m, 0, k_out, v_out, bp.grads, bp.bgrads, nslots, o));
            sync_dense_kernel<<<(unsigned)((m.ndense + 127) / 128), 128, 0, st>>>(m, bp.dWsum, o);
            bl_finish_kernel<<<(bd.P + 127) / 128, 128, 0, st>>>(pl, bd, bp);
            *launches += 6;
        }
    }
    SCU(cudaGetLastError());
    return 0;
}
