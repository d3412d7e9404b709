/*
 * sbr_b200.h -- C ABI of the B200-native sequence-recommender training engine.
 *
 * This is the drop-in boundary for the `fit()` / `predict()` hot path of maciejkula/sbr-rs
 * (reference @ f01d4a7, citations are into /root/reference/src/).  The reference has no FFI of its own;
 * every entry point below states the Rust item it replaces, and INTEGRATION.md shows the `extern "C"`
 * block and the safe wrappers a maintainer of the Rust crate would add to bind it.
 *
 * Conventions
 *  - `usize` ids (lib.rs:77-81 UserId/ItemId/Timestamp) cross the ABI as uint64_t.
 *  - Every constructor returns an owned opaque handle, released by the matching *_free (Rust Drop).
 *  - No exceptions / panics cross the ABI: all failures are sbr_status codes; Rust index panics
 *    (out-of-range ids) surface as SBR_ERR_INVALID_ARGUMENT.  sbr_last_error_string() gives detail.
 *  - All pointers are HOST pointers unless a name says `_device`.  Parameters and optimizer state live in
 *    HBM for the lifetime of the model and persist across fit() calls (warm restart, benches/benchmark.rs:40-42).
 *  - There is NO CPU fallback: every compute entry point returns SBR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SBR_B200_H
#define SBR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SBR_OK = 0,
    SBR_ERR_NO_INTERACTIONS = 1,    /* FittingError::NoInteractions        lib.rs:93-97, sequence_model.rs:86-88 */
    SBR_ERR_INVALID_PREDICTION = 2, /* PredictionError::InvalidPredictionValue lib.rs:85-89, sequence_model.rs:225-229 */
    SBR_ERR_INVALID_ARGUMENT = 3,   /* replaces Rust bounds / chunks_mut(0) panics */
    SBR_ERR_CUDA = 4,
    SBR_ERR_NCCL = 5,
    SBR_ERR_UNSUPPORTED = 6
} sbr_status;

/* models/mod.rs:16-23 */
typedef enum { SBR_LOSS_BPR = 0, SBR_LOSS_HINGE = 1, SBR_LOSS_WARP = 2 } sbr_loss;
/* models/mod.rs:27-32 */
typedef enum { SBR_OPTIMIZER_ADAGRAD = 0, SBR_OPTIMIZER_ADAM = 1 } sbr_optimizer;
/* models/mod.rs:36-41 */
typedef enum { SBR_PARALLELISM_ASYNCHRONOUS = 0, SBR_PARALLELISM_SYNCHRONOUS = 1 } sbr_parallelism;
/* models/lstm.rs:29-35 */
typedef enum { SBR_LSTM_NORMAL = 0, SBR_LSTM_COUPLED = 1 } sbr_lstm_variant;

typedef struct sbr_compressed sbr_compressed;           /* data.rs:228-234 CompressedInteractions */
typedef struct sbr_hyperparameters sbr_hyperparameters; /* lstm.rs:39-52 / ewma.rs:45-57 Hyperparameters */
typedef struct sbr_model sbr_model;                     /* lstm.rs:387-389 ImplicitLSTMModel / ewma.rs:402-404 */
typedef struct sbr_fit_plan sbr_fit_plan;               /* device-resident schedule of one fit() (no Rust analogue) */

const char* sbr_last_error_string(void);
/* number of usable sm_100 devices; 0 when none (compute entry points then fail with SBR_ERR_CUDA) */
int sbr_device_count(void);
/* one process per GPU: select the device this process (all later handles) uses. Default 0. */
sbr_status sbr_set_device(int device);

/* ------------------------------------------------------------------ data.rs ------------------------------- */
/* Interactions::to_compressed (data.rs:180-182, 236-265): stable sort by (user, timestamp), histogram, prefix sum. */
sbr_status sbr_compressed_from_triplets(const uint64_t* user_ids, const uint64_t* item_ids, const uint64_t* timestamps,
                                        size_t nnz, size_t num_users, size_t num_items, sbr_compressed** out);
/* The same on the device: stable LSD radix sort by timestamp then by user, gather, histogram, scan.
 * Result identical to sbr_compressed_from_triplets including the order of (user, timestamp) ties (data.rs:240 sorts
 * stably); the narrowed item-id stream stays resident in HBM, so the first fit() uploads nothing.  Needs the device. */
sbr_status sbr_compressed_from_triplets_device(const uint64_t* user_ids, const uint64_t* item_ids, const uint64_t* timestamps,
                                               size_t nnz, size_t num_users, size_t num_items, sbr_compressed** out);
/* Adopt an existing CSR (the three Vec fields at data.rs:231-233). timestamps may be NULL. Copies. */
sbr_status sbr_compressed_from_csr(const uint64_t* user_pointers, const uint64_t* item_ids, const uint64_t* timestamps,
                                   size_t num_users, size_t num_items, sbr_compressed** out);
/* Zero-copy variant: the handle only borrows the three arrays (they must outlive it), like the `&CompressedInteractions`
 * that fit() takes in the reference (lstm.rs:395). */
sbr_status sbr_compressed_borrow_csr(const uint64_t* user_pointers, const uint64_t* item_ids, const uint64_t* timestamps,
                                     size_t num_users, size_t num_items, sbr_compressed** out);
size_t sbr_compressed_num_users(const sbr_compressed* c); /* data.rs:293-295 */
size_t sbr_compressed_num_items(const sbr_compressed* c); /* data.rs:298-300 */
size_t sbr_compressed_len(const sbr_compressed* c);
/* borrow the CSR arrays (valid until free): to_interactions (data.rs:308-328) / iter_users (data.rs:269-274) */
sbr_status sbr_compressed_borrow(const sbr_compressed* c, const uint64_t** user_pointers, const uint64_t** item_ids,
                                 const uint64_t** timestamps);
/* CompressedInteractionsUser::chunks (data.rs:363-371, 406-432): first chunk smallest.  Writes up to `cap`
 * (start, len) pairs relative to the user's slice and the total count to *n. */
sbr_status sbr_compressed_user_chunks(const sbr_compressed* c, size_t user_id, size_t chunk_size, uint64_t* starts,
                                      uint64_t* lens, size_t cap, size_t* n);
/* Pre-stage the item-id stream in HBM (otherwise done lazily by the first fit/mrr call that needs it). */
sbr_status sbr_compressed_upload(sbr_compressed* c);
void sbr_compressed_free(sbr_compressed* c);

/* data.rs:69-88 user_based_split (no user in both sets): out_is_train[i] = SipHash-2-4(key_0, key_1, user_ids[i]) % 100000 >
 * (test_fraction * 100000) as u64, key_0 / key_1 = two Uniform<u64> draws from the xorshift128 state `rng_state` (in / out).
 * Host-only: data preparation of the reference's tests and examples (lstm.rs:428-430). */
sbr_status sbr_user_based_split(const uint64_t* user_ids, size_t nnz, uint32_t rng_state[4], float test_fraction,
                                uint8_t* out_is_train);
/* data.rs:54-64 train_test_split: shuffle, then the first (test_fraction * nnz) as usize shuffled interactions are the
 * test set.  perm[k] = original index of the k-th shuffled interaction; perm[0 .. *num_test) is test, the rest train. */
sbr_status sbr_train_test_split(size_t nnz, uint32_t rng_state[4], float test_fraction, uint64_t* perm, size_t* num_test);

/* ------------------------------------------------- lstm.rs:54-202 / ewma.rs:59-206 ------------------------ */
sbr_hyperparameters* sbr_lstm_hyperparameters_new(size_t num_items, size_t max_sequence_length); /* lstm.rs:56-71 */
sbr_hyperparameters* sbr_ewma_hyperparameters_new(size_t num_items, size_t max_sequence_length); /* ewma.rs:61-76 */
sbr_status sbr_hyper_learning_rate(sbr_hyperparameters* h, float v);      /* lstm.rs:74-77 */
sbr_status sbr_hyper_l2_penalty(sbr_hyperparameters* h, float v);         /* lstm.rs:80-83 */
sbr_status sbr_hyper_embedding_dim(sbr_hyperparameters* h, size_t v);     /* lstm.rs:86-89 */
sbr_status sbr_hyper_num_epochs(sbr_hyperparameters* h, size_t v);        /* lstm.rs:92-95 */
sbr_status sbr_hyper_loss(sbr_hyperparameters* h, sbr_loss v);            /* lstm.rs:98-101 */
sbr_status sbr_hyper_lstm_variant(sbr_hyperparameters* h, sbr_lstm_variant v); /* lstm.rs:104-107 (LSTM only) */
/* lstm.rs:110-113.  On the GPU `num_threads` is the number of partitions ("threads" of sequence_model.rs:90-98) that train
 * concurrently.  1 reproduces the reference's single-thread update order exactly.  0 = automatic: fill the device, but never
 * more than the data allows (sub-sequences / 16) and -- for an LSTM that has seen fewer than ~100 steps per item -- never more
 * than 2.5 partitions per item: a cold LSTM does not learn with thousands of concurrent sequences per item row;
 * a multi-epoch fit of a cold model runs its first epoch at the bounded count and the rest device-filling. */
sbr_status sbr_hyper_num_threads(sbr_hyperparameters* h, size_t v);
/* lstm.rs:116-119.  Asynchronous = Hogwild.  Synchronous (the reference default) with num_threads > 1 runs in rounds: gradients
 * from the round-start parameters, entries applied un-merged in thread order, one dense step per round (EWMA: 1..8 GPUs, WARP on
 * one GPU; LSTM: one GPU, every loss).  Configurations the round engines cannot run return SBR_ERR_INVALID_ARGUMENT. */
sbr_status sbr_hyper_parallelism(sbr_hyperparameters* h, sbr_parallelism v);
sbr_status sbr_hyper_from_seed(sbr_hyperparameters* h, const uint8_t seed[16]); /* lstm.rs:129-132 */
sbr_status sbr_hyper_optimizer(sbr_hyperparameters* h, sbr_optimizer v);  /* lstm.rs:135-138 */
/* Hyperparameters::random(num_items, rng) (lstm.rs:141-172 / ewma.rs:139-170), "useful for hyperparameter search":
 * draws, in the reference's field order, max_sequence_length = 2^U[4,8), embedding_dim = 2^U[4,8), learning_rate =
 * 10^U[-3,0.5), l2_penalty = 10^U[-7,-3), loss in {BPR, Hinge}, optimizer in {Adam, Adagrad}, (LSTM: variant in {Normal,
 * Coupled},) parallelism, num_threads = U[1, host threads + 1), num_epochs = 2^U[3,7) from the xorshift128 state
 * `rng_state` (in / out; rand 0.5 Uniform sampling [rand-recalled]); the model seed comes from the clock like
 * thread_rng() at lstm.rs:167.  Returns NULL on a bad argument. */
sbr_hyperparameters* sbr_lstm_hyperparameters_random(size_t num_items, uint32_t rng_state[4]);
sbr_hyperparameters* sbr_ewma_hyperparameters_random(size_t num_items, uint32_t rng_state[4]);
/* Engine knob without a reference analogue: 1 = never use the tf32 / bf16 tensor-core tile kernel, i.e. exact fp32
 * arithmetic on every path (the tile kernel is the default for LSTM dim 32 with >= 128 partitions). */
sbr_status sbr_hyper_exact_arithmetic(sbr_hyperparameters* h, int on);
/* The public fields of Hyperparameters (lstm.rs:39-52) as plain data: what a serde round trip / a Rust shim reads. */
typedef struct {
    int32_t model;            /* 0 LSTM, 1 EWMA */
    int32_t lstm_variant, loss, optimizer, parallelism;
    int32_t exact_arithmetic;
    uint64_t num_items, max_sequence_length, embedding_dim, num_threads, num_epochs;
    float learning_rate, l2_penalty;
    uint8_t seed[16];
} sbr_hyper_values;
sbr_status sbr_hyper_get_values(const sbr_hyperparameters* h, sbr_hyper_values* out);
sbr_status sbr_model_get_hyper_values(const sbr_model* m, sbr_hyper_values* out);
void sbr_hyper_free(sbr_hyperparameters* h);
/* Hyperparameters::build(self) (lstm.rs:197-201 / ewma.rs:201-205): consumes `h` (also on failure). */
sbr_status sbr_hyper_build(sbr_hyperparameters* h, sbr_model** out);

/* ---------------------------------------------------------- models ---------------------------------------- */
/* ImplicitLSTMModel::fit / ImplicitEWMAModel::fit (lstm.rs:395-397, ewma.rs:408-410 -> sequence_model.rs:70-178). */
sbr_status sbr_model_fit(sbr_model* m, const sbr_compressed* interactions, float* loss_out);
/* OnlineRankingModel::user_representation (sequence_model.rs:182-211): state after the last
 * max_sequence_length ids.  out has embedding_dim floats. */
sbr_status sbr_model_user_representation(const sbr_model* m, const uint64_t* item_ids, size_t n, float* out);
/* batched variant: users given as CSR slices ptr[u]..ptr[u+1] of item_ids; out is [num_users, embedding_dim] */
sbr_status sbr_model_user_representations(const sbr_model* m, const uint64_t* ptr, const uint64_t* item_ids,
                                          size_t num_users, float* out);
/* OnlineRankingModel::predict (sequence_model.rs:213-232) */
sbr_status sbr_model_predict(const sbr_model* m, const float* user, const uint64_t* item_ids, size_t k, float* out);
/* evaluation::mrr_score (evaluation.rs:12-48) */
sbr_status sbr_model_mrr_score(const sbr_model* m, const sbr_compressed* test, float* out);
/* ParameterNode::index (lstm.rs:272-283): bit-exact row gather, out is [n, embedding_dim] */
sbr_status sbr_model_gather_rows(const sbr_model* m, const uint64_t* item_ids, size_t n, float* out);
/* the same gather timed on the device (CUDA events around 1 warm-up + `iters` launches; *kernel_ms = mean per launch):
 * the "embed-gather HBM GB/s" of BASELINE.json; algorithmic bytes per row 4 D + 4 D + 4.  `out` may be NULL. */
sbr_status sbr_model_gather_rows_timed(const sbr_model* m, const uint64_t* item_ids, size_t n, float* out, int iters,
                                       double* kernel_ms);

size_t sbr_model_embedding_dim(const sbr_model* m);
size_t sbr_model_num_items(const sbr_model* m);
/* Parameter exchange (the serde surface, lstm.rs:204-210).  Names: "item_embeddings" [N*D], "item_biases" [N],
 * "lstm_weights" [2D*4*D] (row k of [h,x], gate f/i/g/o, unit d), "lstm_biases" [4*D], "alpha" [D];
 * optimizer state with suffix ".s1" (Adagrad sum of squares / Adam m) and ".s2" (Adam v). */
sbr_status sbr_model_parameter_len(const sbr_model* m, const char* name, size_t* len);
sbr_status sbr_model_get_parameter(const sbr_model* m, const char* name, float* out, size_t len);
sbr_status sbr_model_set_parameter(sbr_model* m, const char* name, const float* data, size_t len);
sbr_status sbr_model_get_num_updates(const sbr_model* m, uint64_t* out); /* Adam step counter */
sbr_status sbr_model_set_num_updates(sbr_model* m, uint64_t v);
/* Hyperparameters.rng (lstm.rs:49, serialised with the model): xorshift128 state x,y,z,w */
sbr_status sbr_model_get_rng_state(const sbr_model* m, uint32_t out[4]);
sbr_status sbr_model_set_rng_state(sbr_model* m, const uint32_t state[4]);
/* Checkpoint (the reference derives Serialize / Deserialize for the whole model incl. hyperparameters, rng and the
 * optimizer state inside HogwildParameter: lstm.rs:38-52,204-210,386-389).  ONE flat little-endian file, layout:
 *   0   char[8]  "SBRB200\0"          8   u32 version (1)        12  u32 header_bytes (offset of the first blob)
 *   16  sbr_hyper_values (88 bytes, the struct above)
 *   104 u32[4] xorshift128 state of Hyperparameters.rng           120 u64 num_updates (Adam step counter)
 *   128 u32 n_blobs, u32 pad, then n_blobs x { char name[32]; u64 len_floats; u64 file_offset }
 *   header_bytes.. the blobs as raw f32: "item_embeddings" [N*D], "item_biases" [N], "lstm_weights" [2D*4D] +
 *   "lstm_biases" [4D] (LSTM) or "alpha" [D] (EWMA), each followed by its ".s1" (and ".s2" for Adam) state blob.
 * save reads the device tables (a row-sharded model must be attached); load builds a new single-GPU model on the
 * selected device; restore overwrites the parameters of an existing model of the same shape (every rank of a sharded
 * model restores its own rows).  A restored model continues bit-for-bit (tests/test_gpu_checkpoint.py). */
sbr_status sbr_model_save(const sbr_model* m, const char* path);
sbr_status sbr_model_load(const char* path, sbr_model** out);
sbr_status sbr_model_restore(sbr_model* m, const char* path);
void sbr_model_free(sbr_model* m);

/* ------------------------------------------------ multi-GPU (one process per GPU) ------------------------- */
/* Three ways to use several GPUs.  (i) Small catalogues: build an ordinary model per rank, fit each
 * rank's users, and between fits sum the deltas of every parameter / optimizer-state blob over the ranks
 * (sbr_model_get_parameter -> all-reduce -> sbr_model_set_parameter; bench.py's ReplicaSync) -- needs nothing below.
 * (ii) One shared model over NVLink peer mappings (sbr_hyper_shard + sbr_model_ipc_*), described next.
 * (iii) Catalogues too large for one GPU: row-sharded table with an explicit NCCL exchange (sbr_dist_*).
 *
 * (ii) The reference shares ONE parameter set between its worker threads (Arc<HogwildParameter>, lstm.rs:175-181).  Across
 * GPUs the same is done over NVLink: the item table, biases and their optimizer state are row-sharded (item_id %
 * world) and every rank's training kernel gathers / updates remote rows directly in the owner's HBM through CUDA-IPC
 * peer mappings (no staging copies, no collective on the data path); the small dense parameters live on rank 0.
 * Protocol: every rank calls sbr_hyper_shard(rank, world) before build, exports sbr_model_ipc_handle_size() bytes
 * with sbr_model_ipc_export, the host program all-gathers them in rank order (any transport), and every rank calls
 * sbr_model_ipc_attach.  Ranks then call sbr_model_fit concurrently, each on its own users' interactions. */
sbr_status sbr_hyper_shard(sbr_hyperparameters* h, int rank, int world);   /* world in {1,2,4,8} */
/* Parallelism::Synchronous across GPUs (EWMA, BPR/hinge): round-synchronous schedule; per round the ranks exchange
 * requested item ids, the rows, and the gradient rows with NCCL grouped send/recv (all-to-all over NVLink) and the
 * owners apply the optimizer to their own shard, so random table access never leaves a GPU's HBM -- the mode for
 * catalogues too large for one GPU.  Rank 0 creates the 128-byte NCCL id, the host program broadcasts it, every rank
 * calls sbr_dist_init once per process; models are built with sbr_hyper_shard(rank, world) (no IPC attach needed). */
sbr_status sbr_dist_unique_id(uint8_t out[128]);
sbr_status sbr_dist_init(int rank, int world, const uint8_t id[128]);
void sbr_dist_finalize(void);
/* Data-parallel replicas of a small model (one full replica per GPU, every rank trains its own users): adds up what every
 * rank changed since the previous call -- parameters and optimizer state of the item records and the dense weights -- with
 * one ncclAllReduce on device buffers and applies the sum to the common starting point, so all replicas are identical again
 * (the sum of all partitions' updates, as in one Hogwild run over all ranks' partitions, with a staleness of one call).
 * The first call records the starting point.  Needs sbr_dist_init.  *bytes_reduced (optional): size of the all-reduce. */
sbr_status sbr_model_replica_sync(sbr_model* m, size_t* bytes_reduced);
/* Same sharded addressing with all shards in this process / on this device (used by single-GPU tests). */
sbr_status sbr_hyper_virtual_shards(sbr_hyperparameters* h, int shards);
size_t sbr_model_ipc_handle_size(void);
sbr_status sbr_model_ipc_export(const sbr_model* m, void* out);
sbr_status sbr_model_ipc_attach(sbr_model* m, const void* all_handles /* world * handle_size bytes, rank order */);

/* ------------------------------------------- device-resident fit schedule --------------------------------- */
typedef struct {
    uint64_t steps;          /* optimizer steps = sub-sequences processed (sequence_model.rs:111) */
    uint64_t timesteps;      /* the reference's `examples` counter (sequence_model.rs:158) */
    uint64_t partitions;     /* concurrent Hogwild partitions actually used */
    uint64_t kernel_launches;
    uint64_t h2d_bytes, d2h_bytes;
    double train_kernel_ms;  /* CUDA-event time of the training kernel(s) on the engine's stream */
    double total_device_ms;  /* CUDA-event time from first H2D to last D2H */
    double host_prepare_ms;  /* host side of the schedule: counting sub-sequences, master shuffle, partition rngs */
    double upload_ms;        /* wall time of the id-stream upload (narrow + H2D), which runs beside host_prepare; 0 if resident */
    char kernel[64];         /* name of the training kernel the plan ran on (e.g. "lstm_tile_train_kernel<2,2>") */
} sbr_fit_stats;

/* Host-only test hook: the schedule fit() builds from a CSR -- sequence_model.rs:76-84: chunks of every user
 * (data.rs:406-432), the len > 2 filter (:81) and the master-rng shuffle (:84).  Needs no device.  `rng_state` is the
 * xorshift128 state (in: before the shuffle, out: after it); with cap < *nsub only the count is returned. */
sbr_status sbr_host_schedule(const sbr_compressed* c, size_t max_sequence_length, uint32_t rng_state[4], uint64_t* starts,
                             uint32_t* lens, uint32_t* order, size_t cap, size_t* nsub);

/* Host-only test hook: what fit() does with the master rng (sequence_model.rs:84,97) -- Fisher-Yates over the indices
 * 0..nsub, then one [u8; 16] seed per partition (the negative-sampler key of a partition is the first 8 bytes of its seed) --
 * with the rng stream produced by `threads` host threads (xorshift128 jump-ahead; threads <= 1: the plain loop).
 * order: nsub entries; keys: partitions entries (may be NULL when partitions == 0); rng_state in/out. */
sbr_status sbr_host_master_schedule(uint32_t rng_state[4], size_t nsub, size_t partitions, size_t threads, uint32_t* order, uint64_t* keys);

/* Split of sbr_model_fit into "stage once" + "run": create does sequence_model.rs:74-98 (sub-sequences, shuffle,
 * partitions, per-partition rngs) and puts everything in HBM; run does :100-175 for num_epochs epochs. */
sbr_status sbr_fit_plan_create(sbr_model* m, const sbr_compressed* interactions, sbr_fit_plan** out);
sbr_status sbr_fit_plan_run(sbr_fit_plan* p, float* loss_out);
sbr_status sbr_fit_plan_stats(const sbr_fit_plan* p, sbr_fit_stats* out);
/* Test hook: the schedule the plan staged in HBM -- sub-sequences built by the DEVICE chunker in user order (starts, lens:
 * nsub entries) and the master-shuffled partition-major order (partitions * per-partition entries; epochs permute it in
 * place).  With cap < *nsub only the counts are returned.  tests/test_gpu_data_prep.py compares it with sbr_host_schedule. */
sbr_status sbr_fit_plan_read_schedule(const sbr_fit_plan* p, uint64_t* starts, uint32_t* lens, uint32_t* order, size_t cap, size_t* nsub,
                                      size_t* norder);
void sbr_fit_plan_free(sbr_fit_plan* p);
/* Changes Hyperparameters::num_threads of a built model for its following fits (0 = automatic).  Useful on small catalogues:
 * an LSTM trains its first epoch at a bounded partition count and continues with a device-filling one (the
 * automatic setting does exactly that). */
sbr_status sbr_model_set_num_threads(sbr_model* m, size_t num_threads);
/* stats of the most recent sbr_model_fit on this model */
sbr_status sbr_model_last_fit_stats(const sbr_model* m, sbr_fit_stats* out);

#ifdef __cplusplus
}
#endif
#endif
