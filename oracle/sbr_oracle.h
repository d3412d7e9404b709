/*
 * sbr_oracle.h -- CPU ORACLE (test infrastructure only; NOT a product path).
 *
 * Plain-C restatement of the training / inference hot path of maciejkula/sbr-rs
 * (reference @ f01d4a7).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * library (libsbr_b200.so) never links or calls it and has no CPU fallback.
 *
 * PARITY STATUS: "parity unpinned" at the arithmetic (wyrm) boundary.
 *   The reference is Rust and cannot be compiled in this image (no cargo/rustc,
 *   no network); its numerics live in the un-vendored crate wyrm 0.9.x
 *   (Cargo.toml:29, features=["fast-math"]), ndarray 0.11, rand 0.5.  The
 *   reference's tests contain no golden tensors for forward/backward/optimizer.
 *   What IS pinned against the reference's own tests (tests/test_oracle_*.py):
 *     - data.rs:630-662 chunk golden vector ([0,1],[2,3,4])        -- exact
 *     - data.rs:588-627 compress/decompress round-trip property     -- exact
 *     - lstm.rs:522-530 empty interactions => NoInteractions        -- exact
 *     - MRR floors lstm.rs:450-520 / ewma.rs:463-507 on ML-100K     -- statistical
 *   Everything marked [wyrm-recalled] / [rand-recalled] restates the published
 *   algorithm of the pinned dependency from memory.
 *   Gradient correctness is established independently by finite differences.
 */
#ifndef SBR_ORACLE_H
#define SBR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enums mirror src/models/mod.rs:16-41 and src/models/lstm.rs:29-35 */
enum { SBO_LOSS_BPR = 0, SBO_LOSS_HINGE = 1, SBO_LOSS_WARP = 2 };
enum { SBO_OPT_ADAGRAD = 0, SBO_OPT_ADAM = 1 };
enum { SBO_PAR_ASYNC = 0, SBO_PAR_SYNC = 1 };
enum { SBO_LSTM_NORMAL = 0, SBO_LSTM_COUPLED = 1 };
enum { SBO_MODEL_LSTM = 0, SBO_MODEL_EWMA = 1 };

enum { SBO_OK = 0, SBO_ERR_NO_INTERACTIONS = 1, SBO_ERR_INVALID_PREDICTION = 2, SBO_ERR_INVALID_ARGUMENT = 3 };

/* ---- rand 0.5 XorShiftRng [rand-recalled] ---- */
typedef struct { uint32_t x, y, z, w; } sbo_rng;
void     sbo_rng_from_seed(sbo_rng* r, const uint8_t seed[16]);
uint32_t sbo_rng_next_u32(sbo_rng* r);
uint64_t sbo_rng_next_u64(sbo_rng* r);
uint64_t sbo_rng_gen_range(sbo_rng* r, uint64_t low, uint64_t high);     /* Rng::gen_range, half-open */
void     sbo_rng_gen_seed(sbo_rng* r, uint8_t out[16]);                  /* rng.gen::<[u8;16]>() */
void     sbo_shuffle_u32(sbo_rng* r, uint32_t* v, size_t n);             /* Rng::shuffle */

/* counter-based negative draw shared bit-for-bit with the CUDA engine */
uint32_t sbo_draw_item(uint64_t key, uint64_t step, uint32_t t, uint32_t j, uint32_t num_items);
/* experiment switch: 1 = sum the entries of a row recorded several times in one step, one optimizer visit per row */
void sbo_set_merge_sparse(int on);

/* ---- data.rs ---- */
/* data.rs:236-265: stable sort by (user, timestamp), histogram, prefix sum. Outputs caller-allocated. */
int sbo_compress(const uint64_t* users, const uint64_t* items, const uint64_t* ts, size_t nnz,
                 size_t num_users, uint64_t* user_ptr /*num_users+1*/, uint64_t* item_ids, uint64_t* timestamps);
/* data.rs:406-432: first-chunk-smallest chunking of one user's history of length len.
   Writes (start,len) pairs; returns number of chunks. */
size_t sbo_chunks(size_t len, size_t chunk_size, uint64_t* starts, uint64_t* lens);
/* sequence_model.rs:76-83: flat list of sub-sequences with len > 2 in user order.
   starts/lens may be NULL to count.  Offsets index item_ids. */
size_t sbo_subsequences(const uint64_t* user_ptr, size_t num_users, size_t max_len,
                        uint64_t* starts, uint32_t* lens);
/* data.rs:69-88 user_based_split: out_is_train[i] per interaction. [rand-recalled, siphasher SipHash-2-4] */
void sbo_user_based_split(const uint64_t* users, size_t nnz, sbo_rng* rng, float test_fraction, uint8_t* out_is_train);
uint64_t sbo_siphash24_u64(uint64_t k0, uint64_t k1, uint64_t value);

/* ---- model ---- */
typedef struct sbo_model sbo_model;

typedef struct {
    int model;            /* SBO_MODEL_* */
    size_t num_items, max_sequence_length, embedding_dim;
    float learning_rate, l2_penalty;
    int lstm_variant, loss, optimizer, parallelism;
    int num_threads, num_epochs;
    uint8_t seed[16];
} sbo_hyper;

/* defaults of lstm.rs:56-71 / ewma.rs:61-76 (num_threads := 1, seed := 42s) */
void sbo_hyper_default(sbo_hyper* h, int model, size_t num_items, size_t max_sequence_length);

sbo_model* sbo_model_new(const sbo_hyper* h);     /* lstm.rs:174-201 / ewma.rs:167-205 */
void       sbo_model_free(sbo_model* m);

/* parameter blobs: "item_embeddings"[N*D] "item_biases"[N] "lstm_weights"[2D*4*D] "lstm_biases"[4*D] "alpha"[D];
   optimizer state: append ".s1" (Adagrad G / Adam m) or ".s2" (Adam v).  Returns pointer+len, NULL if absent. */
float* sbo_model_param(sbo_model* m, const char* name, size_t* len);
uint64_t* sbo_model_num_updates(sbo_model* m);
sbo_rng* sbo_model_rng(sbo_model* m);   /* Hyperparameters.rng (lstm.rs:49) */

/* sequence_model.rs:70-178 */
int sbo_fit(sbo_model* m, const uint64_t* user_ptr, const uint64_t* item_ids, size_t num_users, float* loss_out);
/* one optimizer step on one sub-sequence (the body of the loop at sequence_model.rs:111-169).
   key/step feed the negative sampler.  If negatives_out != NULL the drawn negatives are written (len-1).
   If apply_update == 0 only gradients are computed (used by the finite-difference tests):
   dense gradient is written to dense_grad_out (may be NULL). Returns the summed loss. */
float sbo_step(sbo_model* m, const uint64_t* ids, size_t len, uint64_t key, uint64_t step,
               uint32_t* negatives_out, int apply_update, float* dense_grad_out,
               const uint32_t* forced_negatives);
/* loss only (no grad), with forced negatives: for finite differences */
double sbo_loss_only(sbo_model* m, const uint64_t* ids, size_t len, const uint32_t* negatives);
/* sparse grads from the last sbo_step(apply_update=0): entries in application order */
size_t sbo_last_sparse_grads(sbo_model* m, const uint32_t** rows, const float** grads /*[n*D]*/,
                             const uint32_t** brows, const float** bgrads, size_t* nb);

/* sequence_model.rs:182-232 */
int sbo_user_representation(sbo_model* m, const uint64_t* ids, size_t n, float* out_D);
int sbo_predict(sbo_model* m, const float* user_D, const uint64_t* ids, size_t k, float* out);
/* evaluation.rs:12-48 */
int sbo_mrr_score(sbo_model* m, const uint64_t* user_ptr, const uint64_t* item_ids, size_t num_users, float* out);

#ifdef __cplusplus
}
#endif
#endif
