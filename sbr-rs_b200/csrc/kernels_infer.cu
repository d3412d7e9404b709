// kernels_infer.cu -- gather / inference / evaluation / layout kernels.
//   gather_rows            ParameterNode::index                         lstm.rs:272-283
//   user_representations   OnlineRankingModel::user_representation      sequence_model.rs:182-211
//   predict                OnlineRankingModel::predict / predict_single sequence_model.rs:213-232, lstm.rs:338-350
//   mrr                    evaluation::mrr_score                        evaluation.rs:12-48
#include <cuda_runtime.h>
#include <float.h>

#include "engine.h"

namespace sbr {

namespace {

// ---- K1: bit-exact row gather.  One thread moves one 16-byte piece of one row: a warp covers 512 contiguous
// output bytes and 512/(4D) whole table rows; loads bypass L1 (random rows, no reuse), stores stream.
__global__ void __launch_bounds__(256) gather_rows_kernel(ModelDev m, const uint32_t* __restrict__ ids, size_t n,
                                                          float* __restrict__ out) {
    const int q = m.D >> 2;  // float4 per row
    const size_t total = n * (size_t)q;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / q; const int c = (int)(i - r * q);
        const uint32_t id = __ldg(ids + r);
        const float4 v = __ldcg(reinterpret_cast<const float4*>(item_rec(m, id)) + c);
        __stcs(reinterpret_cast<float4*>(out) + i, v);
    }
}

// ---- recurrent state after consuming a history (one warp per user) ----
template <int D>
__device__ __forceinline__ void ewma_represent(const ModelDev& m, const uint32_t* ids, int n, int lane, float* out) {
    constexpr int V = VecOf<D>::V;
    float al[V], a[V], s[V], x[V];
    row_load_cg<D>(m.dense, lane, al);
#pragma unroll
    for (int v = 0; v < V; ++v) { a[v] = sigmoidf_(al[v]); s[v] = 0.0f; }
    for (int t = 0; t < n; ++t) {
        const uint32_t in = ids ? __ldg(ids + t) : 0u;
        row_load_cg<D>(item_rec(m, in), lane, x);
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] = t == 0 ? x[v] : a[v] * s[v] + (1.0f - a[v]) * x[v];
    }
    vec_store<D>(out, lane, s);
}

// LSTM state after a history; lane l owns units l*V..l*V+V-1 (D < 32: lane l < D owns unit l).  zbuf: 2*D floats of
// shared memory private to the calling warp.
template <int D>
__device__ __forceinline__ void lstm_represent(const ModelDev& m, const uint32_t* ids, int n, int lane, float* out, float* zbuf) {
    constexpr int V = VecOf<D>::V;
    const float* W = m.dense; const float* B = m.dense + (size_t)2 * D * 4 * D;
    const bool coupled = m.variant == 1;
    float h[V], c[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { h[v] = 0.0f; c[v] = 0.0f; }
    for (int t = 0; t < n; ++t) {
        const uint32_t in = ids ? __ldg(ids + t) : 0u;
        float x[V], pre[4][V];
        row_load_cg<D>(item_rec(m, in), lane, x);
        vec_store<D>(zbuf, lane, h); vec_store<D>(zbuf + D, lane, x);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) row_load_cg<D>(B + q * D, lane, pre[q]);
        for (int k = 0; k < 2 * D; ++k) {
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
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float f = sigmoidf_(pre[0][v]);
            const float ig = coupled ? 1.0f - f : sigmoidf_(pre[1][v]);
            const float gg = tanhf(pre[2][v]);
            const float og = sigmoidf_(pre[3][v]);
            c[v] = f * c[v] + ig * gg;
            h[v] = og * tanhf(c[v]);
        }
    }
    vec_store<D>(out, lane, h);
}

template <int D>
__device__ __forceinline__ void represent(const ModelDev& m, const uint32_t* ids, int n, int lane, float* out, float* zbuf) {
    if (n == 0) { ids = nullptr; n = 1; }  // sequence_model.rs:197-200: hidden_states[0] with the default index 0
    if (m.model == MODEL_EWMA) ewma_represent<D>(m, ids, n, lane, out);
    else lstm_represent<D>(m, ids, n, lane, out, zbuf);
}

template <int D>
__global__ void __launch_bounds__(128) user_rep_kernel(ModelDev m, const uint64_t* __restrict__ ptr,
                                                       const uint32_t* __restrict__ ids, size_t num_users, float* out) {
    __shared__ float zb[4][2 * D];
    const int lane = threadIdx.x & 31;
    const size_t u = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= num_users) return;
    uint64_t b = ptr[u], e = ptr[u + 1];
    uint64_t n = e - b;
    if (n > (uint64_t)m.T) { b = e - m.T; n = m.T; }  // sequence_model.rs:188
    represent<D>(m, ids + b, (int)n, lane, out + u * D, zb[threadIdx.x >> 5]);
}

// predict_single: bias + dot(user, row).  One thread per item, 16-byte loads.
template <int D>
__device__ __forceinline__ float score_item(const ModelDev& m, const float* __restrict__ user, uint32_t id) {
    const float4* row = reinterpret_cast<const float4*>(item_rec(m, id));
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
        const float4 r = __ldcg(row + c);
        acc = fmaf(user[4 * c], r.x, acc); acc = fmaf(user[4 * c + 1], r.y, acc);
        acc = fmaf(user[4 * c + 2], r.z, acc); acc = fmaf(user[4 * c + 3], r.w, acc);
    }
    return __ldcg(reinterpret_cast<const float*>(bias_rec(m, id))) + acc;
}

template <int D>
__global__ void __launch_bounds__(256) predict_kernel(ModelDev m, const float* __restrict__ user_g,
                                                      const uint32_t* __restrict__ ids, size_t k, float* out, int* nonfinite) {
    __shared__ float user[D];
    for (int d = threadIdx.x; d < D; d += blockDim.x) user[d] = user_g[d];
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const float v = score_item<D>(m, user, ids[i]);
    if (!isfinite(v)) atomicOr(nonfinite, 1);  // sequence_model.rs:225-229
    out[i] = v;
}

// mrr_score: one CTA per user (grid-strided).  pred is a [gridDim][N] scratch.
template <int D>
__global__ void __launch_bounds__(256) mrr_kernel(ModelDev m, uint32_t num_items, const uint64_t* __restrict__ ptr,
                                                  const uint32_t* __restrict__ ids, size_t num_users, float* pred_all,
                                                  float* rr, int* nonfinite) {
    __shared__ float user[D];
    __shared__ float zb[2 * D];
    __shared__ unsigned int cnt;
    float* pred = pred_all + (size_t)blockIdx.x * num_items;   // evaluation.rs:16: 0..test.num_items()
    for (size_t u = blockIdx.x; u < num_users; u += gridDim.x) {
        const uint64_t b = ptr[u], e = ptr[u + 1];
        const uint64_t len = e - b;
        if (len < 2) { if (threadIdx.x == 0) rr[u] = 0.0f; continue; }           // evaluation.rs:20
        const uint32_t test_item = ids[e - 1];                                   // :25
        uint64_t hb = b, hn = len - 1;                                           // :24 all but the last
        if (hn > (uint64_t)m.T) { hb = (e - 1) - m.T; hn = m.T; }
        if (threadIdx.x < 32) represent<D>(m, ids + hb, (int)hn, threadIdx.x, user, zb);  // :27
        if (threadIdx.x == 0) cnt = 0;
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < num_items; j += blockDim.x) {          // :16,28 all items of the TEST set
            const float v = score_item<D>(m, user, j);
            if (!isfinite(v)) atomicOr(nonfinite, 1);
            pred[j] = v;
        }
        __syncthreads();
        for (uint64_t j = b + threadIdx.x; j < e - 1; j += blockDim.x) pred[ids[j]] = -FLT_MAX;  // :30-32 f32::MIN
        __syncthreads();
        const float ts = pred[test_item];                                        // :34
        unsigned int local = 0;
        for (uint32_t j = threadIdx.x; j < num_items; j += blockDim.x) local += pred[j] >= ts ? 1u : 0u;  // :37-41
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(kFull, local, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, local);
        __syncthreads();
        if (threadIdx.x == 0) rr[u] = 1.0f / (float)cnt;                         // :43
        __syncthreads();
    }
}

// embedding_init: N(0, (1/D)^2) (lstm.rs:22-25), counter-based so tables of any size initialise in HBM
__global__ void __launch_bounds__(256) init_embeddings_kernel(ModelDev m, uint64_t seed) {
    const size_t total = (size_t)m.N * m.D;
    const float sd = 1.0f / (float)m.D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = seed + i * 0x9E3779B97F4A7C15ULL;
        v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ULL; v ^= v >> 27; v *= 0x94D049BB133111EBULL; v ^= v >> 31;
        const float u1 = ((float)(uint32_t)(v >> 40) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(uint32_t)(v & 0xffffffu) + 0.5f) * (1.0f / 16777216.0f);
        const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        const size_t r = i / m.D; const int d = (int)(i - r * m.D);
        if (m.own_shard >= 0 && (int)(r & m.gmask) != m.own_shard) continue;
        float* rec = item_rec(m, (uint32_t)r);
        rec[d] = z * sd;
        for (int s = 1; s < m.S; ++s) rec[s * m.D + d] = 0.0f;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m.N; i += (size_t)gridDim.x * blockDim.x)
        if (m.own_shard < 0 || (int)(i & m.gmask) == m.own_shard) *bias_rec(m, (uint32_t)i) = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) pack_rows_kernel(ModelDev m, int slot, const float* __restrict__ packed, float* out, int dir) {
    const size_t total = (size_t)m.N * m.D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / m.D; const int d = (int)(i - r * m.D);
        if (dir == 0 && m.own_shard >= 0 && (int)(r & m.gmask) != m.own_shard) continue;
        float* cell = item_rec(m, (uint32_t)r) + (size_t)slot * m.D + d;
        if (dir == 0) *cell = packed[i]; else out[i] = *cell;
    }
}
__global__ void __launch_bounds__(256) pack_bias_kernel(ModelDev m, int slot, const float* __restrict__ packed, float* out, int dir) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m.N; i += (size_t)gridDim.x * blockDim.x) {
        if (dir == 0 && m.own_shard >= 0 && (int)(i & m.gmask) != m.own_shard) continue;   // peers' shards are theirs to write
        float* cell = reinterpret_cast<float*>(bias_rec(m, i)) + slot;
        if (dir == 0) *cell = packed[i]; else out[i] = *cell;
    }
}

inline int grid_for(size_t n, int block, int cap = 148 * 16) {
    size_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (size_t)cap) g = cap;
    return (int)g;
}

#define SBR_DISPATCH_D(D_, ...)                                   \
    switch (D_) {                                                 \
        case 16: { constexpr int kD = 16; __VA_ARGS__; } break;   \
        case 32: { constexpr int kD = 32; __VA_ARGS__; } break;   \
        case 64: { constexpr int kD = 64; __VA_ARGS__; } break;   \
        case 128: { constexpr int kD = 128; __VA_ARGS__; } break; \
        case 256: { constexpr int kD = 256; __VA_ARGS__; } break; \
        default: return cudaErrorInvalidValue;                    \
    }

}  // namespace

cudaError_t launch_gather_rows(const ModelDev& m, const uint32_t* ids, size_t n, float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t total = n * (size_t)(m.D / 4);
    gather_rows_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(m, ids, n, out);
    return cudaGetLastError();
}

cudaError_t launch_user_representations(const ModelDev& m, const uint64_t* ptr, const uint32_t* ids, size_t num_users,
                                        float* out, cudaStream_t st) {
    if (num_users == 0) return cudaSuccess;
    const int grid = (int)((num_users + 3) / 4);
    SBR_DISPATCH_D(m.D, user_rep_kernel<kD><<<grid, 128, 0, st>>>(m, ptr, ids, num_users, out));
    return cudaGetLastError();
}

cudaError_t launch_predict(const ModelDev& m, const float* user, const uint32_t* ids, size_t k, float* out, int* nonfinite,
                           cudaStream_t st) {
    if (k == 0) return cudaSuccess;
    const int grid = (int)((k + 255) / 256);
    SBR_DISPATCH_D(m.D, predict_kernel<kD><<<grid, 256, 0, st>>>(m, user, ids, k, out, nonfinite));
    return cudaGetLastError();
}

cudaError_t launch_mrr(const ModelDev& m, uint32_t num_items, const uint64_t* ptr, const uint32_t* ids, size_t num_users, float* rr,
                       int* nonfinite, cudaStream_t st) {
    if (num_users == 0) return cudaSuccess;
    int grid = (int)(num_users < 148 * 2 ? num_users : 148 * 2);
    float* pred = nullptr;
    cudaError_t e = cudaMallocAsync(&pred, sizeof(float) * (size_t)grid * num_items, st);
    if (e != cudaSuccess) return e;
    SBR_DISPATCH_D(m.D, mrr_kernel<kD><<<grid, 256, 0, st>>>(m, num_items, ptr, ids, num_users, pred, rr, nonfinite));
    e = cudaGetLastError();
    cudaFreeAsync(pred, st);
    return e;
}

cudaError_t launch_init_embeddings(const ModelDev& m, uint64_t seed, cudaStream_t st) {
    init_embeddings_kernel<<<grid_for((size_t)m.N * m.D, 256), 256, 0, st>>>(m, seed);
    return cudaGetLastError();
}
cudaError_t launch_pack_rows(const ModelDev& m, int slot, const float* packed, cudaStream_t st) {
    pack_rows_kernel<<<grid_for((size_t)m.N * m.D, 256), 256, 0, st>>>(m, slot, packed, nullptr, 0);
    return cudaGetLastError();
}
cudaError_t launch_unpack_rows(const ModelDev& m, int slot, float* packed, cudaStream_t st) {
    pack_rows_kernel<<<grid_for((size_t)m.N * m.D, 256), 256, 0, st>>>(m, slot, nullptr, packed, 1);
    return cudaGetLastError();
}
cudaError_t launch_pack_bias(const ModelDev& m, int slot, const float* packed, cudaStream_t st) {
    pack_bias_kernel<<<grid_for(m.N, 256), 256, 0, st>>>(m, slot, packed, nullptr, 0);
    return cudaGetLastError();
}
cudaError_t launch_unpack_bias(const ModelDev& m, int slot, float* packed, cudaStream_t st) {
    pack_bias_kernel<<<grid_for(m.N, 256), 256, 0, st>>>(m, slot, nullptr, packed, 1);
    return cudaGetLastError();
}

}  // namespace sbr
