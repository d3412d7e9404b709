// engine.h -- plain structs shared by the host driver (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "common.cuh"

namespace sbr {

enum { MODEL_LSTM = 0, MODEL_EWMA = 1 };

// Device-resident model (HBM layout, see DESIGN.md "Data layout")
struct ModelDev {
    int model, variant, loss, opt;
    uint32_t N;
    int D, T;
    int S;           // float-vectors per item record: 2 (w,G) Adagrad, 3 (w,m,v) Adam
    // Item table, row-sharded by item id: shard = id & gmask, local row = id >> gshift (G = gmask+1 is 1, 2, 4 or 8).
    // With G == 1 everything is in Es[0].  With G > 1 a shard is either local memory (virtual shards, tests)
    // or a peer GPU's memory mapped through CUDA IPC (one process per GPU; loads/stores travel over NVLink).
    // One RECORD per item: {b, b.s1, b.s2, pad} (16 bytes) | w[D] | s1[D] | (s2[D])  --  4 + S*D floats, contiguous, so
    // that ONE bulk copy / bulk reduce-add moves everything a sparse optimizer visit of the item touches.
    float* Es[8];    // [rows of shard][4 + S*D]
    uint32_t gmask;
    int gshift;
    int hbm_resident; // table + state larger than L2: kernels prefetch rows a few timesteps ahead
    int own_shard;   // >= 0: only this shard is written by init / set_parameter (multi-process); -1: all shards are ours
    float* dense;    // [3][ndense]  w | s1 | s2  (LSTM: W[2D][4][D] then bias[4][D]; EWMA: alpha[D])
    size_t ndense;
    float lr, l2;
    int exact;       // sbr_hyper_exact_arithmetic: never use the tf32 / bf16 tile kernel
};

// Device-resident schedule of one fit(): sequence_model.rs:74-98 materialised in HBM
struct PlanDev {
    const uint32_t* item_ids;   // [nnz] narrowed item-id stream
    const uint64_t* seq_start;  // [nsub] offset of each sub-sequence in item_ids
    const uint32_t* seq_len;    // [nsub]
    uint32_t* order;            // [P*n] partition-major indices into seq_*; shuffled in place every epoch
    uint32_t n;                 // sub-sequences per partition (= nsub / P, remainder dropped)
    uint32_t P;                 // partitions ("num_threads")
    uint32_t neg_range;         // negatives ~ U[0, interactions.num_items)  (sequence_model.rs:74)
    XorShift* rng;              // [P] per-partition shuffle rng (persists across runs)
    uint64_t* keys;             // [P] negative-sampler keys
    uint64_t* step_ctr;         // [P] optimizer steps done so far by each partition
    float* loss_acc;            // [P] sum of sequence losses
    unsigned long long* examples;  // [P] sum of timesteps
    float* scratch;             // warp-private activations for backward
    size_t scratch_stride;      // floats per warp
    int epochs;
    unsigned long long adam_t0; // model num_updates before this run
};

#ifdef __CUDACC__
__host__ __device__ __forceinline__ size_t rec_floats(const ModelDev& m) { return 4 + (size_t)m.S * m.D; }
// the record's vectors: item_rec() points at w (s1 at + D, s2 at + 2 D); bias_rec() at the {b, s1, s2, pad} quad in front of it
__device__ __forceinline__ float* item_rec(const ModelDev& m, uint32_t id) {
    return m.Es[id & m.gmask] + (size_t)(id >> m.gshift) * rec_floats(m) + 4;
}
__device__ __forceinline__ float4* bias_rec(const ModelDev& m, uint32_t id) {
    return reinterpret_cast<float4*>(m.Es[id & m.gmask] + (size_t)(id >> m.gshift) * rec_floats(m));
}
// local row r of this process's own shard
__device__ __forceinline__ float* shard_item_rec(const ModelDev& m, int shard, uint32_t r) { return m.Es[shard] + (size_t)r * rec_floats(m) + 4; }
__device__ __forceinline__ float4* shard_bias_rec(const ModelDev& m, int shard, uint32_t r) { return reinterpret_cast<float4*>(m.Es[shard] + (size_t)r * rec_floats(m)); }
#endif

// launchers (kernels_train.cu / kernels_infer.cu); all enqueue on `st` and return the launch count
int launch_train(const ModelDev& m, const PlanDev& p, int num_sms, cudaStream_t st, cudaError_t* err, const char** kernel_name);
size_t train_scratch_floats_per_warp(const ModelDev& m);
int train_auto_partitions(const ModelDev& m, int num_sms);
bool train_supported(const ModelDev& m, const char** why);
int lstm_kernel_choice(const ModelDev& m, uint32_t P);
cudaError_t launch_lstm_tile(const ModelDev& m, const PlanDev& p, cudaStream_t st);
cudaError_t launch_ewma_tile(const ModelDev& m, const PlanDev& p, cudaStream_t st);
int ewma_tile_tiles_per_cta(const ModelDev& m, uint32_t P);
int lstm_tile_tiles_per_cta(const ModelDev& m, uint32_t P);

// round-synchronous engine (sync_engine.cu)
struct SyncBuffers;
SyncBuffers* sync_buffers_new();
void sync_buffers_free(SyncBuffers* b);
bool sync_buffers_copy_engine(const SyncBuffers* b);   // ... pushed by the copy engines from a staging buffer rather than stored by the kernels
bool sync_buffers_p2p(const SyncBuffers* b);   // rows / gradient rows travel as peer stores (CUDA IPC mappings) instead of NCCL send/recv
bool sync_supported(const ModelDev& m, const char** why);
}  // namespace sbr
#include <string>
namespace sbr {
int run_sync_ewma(const ModelDev& m, PlanDev& pl, SyncBuffers& B, void* comm, int rank, int world, uint64_t num_updates,
                  cudaStream_t st, int* launches, uint64_t* rounds_out, std::string* err);

// round-synchronous batched LSTM engine on tcgen05 GEMMs (lstm_batch.cuh): Parallelism::Synchronous for LSTM models and the
// throughput path of embedding_dim 64 / 128 / 256
struct BatchBuffers;
BatchBuffers* batch_buffers_new();
void batch_buffers_free(BatchBuffers* b);
bool batch_lstm_supported(const ModelDev& m, uint32_t P, const char** why);
int run_batch_lstm(const ModelDev& m, PlanDev& pl, BatchBuffers& B, uint64_t num_updates, int num_sms, cudaStream_t st, int* launches,
                   uint64_t* rounds_out, std::string* err);

// Interactions::to_compressed on the device (data_prep.cu): 0 ok, 1 CUDA error, 2 invalid argument
int device_csr_build(const uint64_t* h_user, const uint64_t* h_item, const uint64_t* h_ts, size_t nnz, size_t num_users, size_t num_items,
                     uint64_t* h_user_ptr, uint64_t* h_item_out, uint64_t* h_ts_out, uint32_t* d_item_u32, uint64_t* d_user_ptr,
                     cudaStream_t st, std::string* err);

// sequence_model.rs:76-81 on the device (data_prep.cu): chunks of every user, len > 2 filter, in user order
cudaError_t device_schedule(const uint64_t* d_user_ptr, size_t num_users, size_t T, uint32_t* d_counts, void* d_tmp, size_t* tmp_bytes, size_t nsub,
                            uint64_t* d_seq_start, uint32_t* d_seq_len, cudaStream_t st);
cudaError_t launch_narrow_ids(const uint64_t* d_src, size_t n, uint64_t bound, uint32_t* d_dst, int* d_bad, cudaStream_t st);

cudaError_t launch_gather_rows(const ModelDev& m, const uint32_t* ids_dev, size_t n, float* out_dev, cudaStream_t st);
cudaError_t launch_user_representations(const ModelDev& m, const uint64_t* ptr_dev, const uint32_t* ids_dev, size_t num_users,
                                        float* out_dev, cudaStream_t st);
cudaError_t launch_predict(const ModelDev& m, const float* user_dev, const uint32_t* ids_dev, size_t k, float* out_dev,
                           int* nonfinite_dev, cudaStream_t st);
// reciprocal rank per user with >= 2 interactions (0 for the others); evaluation.rs:12-48
// items 0..num_items (the TEST set's num_items, evaluation.rs:16) are scored and ranked
cudaError_t launch_mrr(const ModelDev& m, uint32_t num_items, const uint64_t* ptr_dev, const uint32_t* ids_dev, size_t num_users, float* rr_dev,
                       int* nonfinite_dev, cudaStream_t st);
cudaError_t launch_init_embeddings(const ModelDev& m, uint64_t seed, cudaStream_t st);
// strided copy between packed host-order blobs and the record layouts
cudaError_t launch_pack_rows(const ModelDev& m, int slot, const float* packed_dev, cudaStream_t st);    // packed[N*D] -> E slot
cudaError_t launch_unpack_rows(const ModelDev& m, int slot, float* packed_dev, cudaStream_t st);        // E slot -> packed
cudaError_t launch_pack_bias(const ModelDev& m, int slot, const float* packed_dev, cudaStream_t st);
cudaError_t launch_unpack_bias(const ModelDev& m, int slot, float* packed_dev, cudaStream_t st);

}  // namespace sbr
