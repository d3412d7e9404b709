/*
 * sbr_oracle.c -- CPU ORACLE (test infrastructure only; see sbr_oracle.h header note).
 * "parity unpinned" at the wyrm arithmetic boundary; pinned on the reference's data-path
 * golden vectors and MRR floors.  Citations are into /root/reference/src/.
 */
#define _GNU_SOURCE
#include "sbr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * rand 0.5 XorShiftRng  [rand-recalled]: xorshift128, seed = 4 little-endian u32
 * ---------------------------------------------------------------------------------------- */
void sbo_rng_from_seed(sbo_rng* r, const uint8_t seed[16]) {
    uint32_t s[4];
    for (int i = 0; i < 4; ++i)
        s[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
               ((uint32_t)seed[4 * i + 3] << 24);
    if ((s[0] | s[1] | s[2] | s[3]) == 0) { /* all-zero seed is replaced by fixed constants */
        s[0] = 0x193a6754u; s[1] = 0xa8a7d469u; s[2] = 0x97830e05u; s[3] = 0x113ba7bbu;
    }
    r->x = s[0]; r->y = s[1]; r->z = s[2]; r->w = s[3];
}

uint32_t sbo_rng_next_u32(sbo_rng* r) {
    uint32_t t = r->x ^ (r->x << 11);
    r->x = r->y; r->y = r->z; r->z = r->w;
    r->w = r->w ^ (r->w >> 19) ^ (t ^ (t >> 8));
    return r->w;
}

uint64_t sbo_rng_next_u64(sbo_rng* r) { /* next_u64_via_u32: low word first */
    uint64_t lo = sbo_rng_next_u32(r);
    uint64_t hi = sbo_rng_next_u32(r);
    return (hi << 32) | lo;
}

/* Rng::gen_range(low, high) for usize: widening multiply with a conservative rejection zone */
uint64_t sbo_rng_gen_range(sbo_rng* r, uint64_t low, uint64_t high) {
    uint64_t range = high - low;
    if (range == 0) return low;
    uint64_t zone = (range << __builtin_clzll(range)) - 1;
    for (;;) {
        uint64_t v = sbo_rng_next_u64(r);
        unsigned __int128 m = (unsigned __int128)v * range;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo <= zone) return low + hi;
    }
}

void sbo_rng_gen_seed(sbo_rng* r, uint8_t out[16]) { /* gen::<[u8;16]>(): one next_u32 per byte */
    for (int i = 0; i < 16; ++i) out[i] = (uint8_t)sbo_rng_next_u32(r);
}

void sbo_shuffle_u32(sbo_rng* r, uint32_t* v, size_t n) { /* Fisher-Yates from the top (rand 0.5 Rng::shuffle) */
    size_t i = n;
    while (i >= 2) {
        i -= 1;
        size_t j = (size_t)sbo_rng_gen_range(r, 0, i + 1);
        uint32_t tmp = v[i]; v[i] = v[j]; v[j] = tmp;
    }
}

/* Counter-based negative draw.  Replaces `negative_item_range.sample(thread_rng)`
 * (sequence_model.rs:59,137): uniform over [0, num_items), no filtering of the positive / history.
 * Keyed by (partition key, optimizer-step index, timestep, draw index) so that a one-ulp difference in a
 * WARP accept/reject decision cannot desynchronise the rest of the stream between CPU and GPU. */
uint32_t sbo_draw_item(uint64_t key, uint64_t step, uint32_t t, uint32_t j, uint32_t num_items) {
    uint64_t v = key + step * 0x9E3779B97F4A7C15ULL + ((uint64_t)t * 8u + j) * 0xD1B54A32D192ED03ULL;
    v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ULL;
    v ^= v >> 27; v *= 0x94D049BB133111EBULL;
    v ^= v >> 31;
    uint32_t r = (uint32_t)(v >> 32);
    return (uint32_t)(((uint64_t)r * (uint64_t)num_items) >> 32);
}

/* ------------------------------------------------------------------------------------------
 * data.rs
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t user, ts, item; size_t pos; } sbo_triplet;

static int cmp_triplet(const void* a, const void* b) { /* data.rs:213-221 cmp_timestamp + stability */
    const sbo_triplet* x = (const sbo_triplet*)a; const sbo_triplet* y = (const sbo_triplet*)b;
    if (x->user != y->user) return x->user < y->user ? -1 : 1;
    if (x->ts != y->ts) return x->ts < y->ts ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1; /* slice::sort_by is stable (data.rs:240) */
    return 0;
}

int sbo_compress(const uint64_t* users, const uint64_t* items, const uint64_t* ts, size_t nnz, size_t num_users,
                 uint64_t* user_ptr, uint64_t* item_ids, uint64_t* timestamps) {
    sbo_triplet* d = (sbo_triplet*)malloc(sizeof(sbo_triplet) * (nnz ? nnz : 1));
    for (size_t i = 0; i < nnz; ++i) {
        if (users[i] >= num_users) { free(d); return SBO_ERR_INVALID_ARGUMENT; } /* Rust: index panic */
        d[i].user = users[i]; d[i].ts = ts[i]; d[i].item = items[i]; d[i].pos = i;
    }
    qsort(d, nnz, sizeof(sbo_triplet), cmp_triplet);
    memset(user_ptr, 0, sizeof(uint64_t) * (num_users + 1));
    for (size_t i = 0; i < nnz; ++i) { /* data.rs:246-251 */
        item_ids[i] = d[i].item; timestamps[i] = d[i].ts;
        user_ptr[d[i].user + 1] += 1;
    }
    for (size_t i = 1; i <= num_users; ++i) user_ptr[i] += user_ptr[i - 1]; /* data.rs:253-255 */
    free(d);
    return SBO_OK;
}

size_t sbo_chunks(size_t len, size_t chunk_size, uint64_t* starts, uint64_t* lens) { /* data.rs:406-432 */
    size_t idx = 0, n = 0;
    while (idx < len) {
        size_t mod = (len - idx) % chunk_size;
        size_t cs = mod == 0 ? chunk_size : mod;
        if (starts) starts[n] = idx;
        if (lens) lens[n] = cs;
        idx += cs; n++;
    }
    return n;
}

size_t sbo_subsequences(const uint64_t* user_ptr, size_t num_users, size_t max_len, uint64_t* starts, uint32_t* lens) {
    size_t n = 0;
    for (size_t u = 0; u < num_users; ++u) { /* sequence_model.rs:76-83 */
        size_t b = user_ptr[u], len = user_ptr[u + 1] - b, idx = 0;
        while (idx < len) {
            size_t mod = (len - idx) % max_len;
            size_t cs = mod == 0 ? max_len : mod;
            if (cs > 2) { /* filter(|item_ids| item_ids.len() > 2) */
                if (starts) starts[n] = b + idx;
                if (lens) lens[n] = (uint32_t)cs;
                n++;
            }
            idx += cs;
        }
    }
    return n;
}

/* SipHash-2-4 (siphasher 0.2) of one usize written with write_usize (8 native-endian bytes) */
#define ROTL64(x, b) (((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND do { v0 += v1; v1 = ROTL64(v1, 13); v1 ^= v0; v0 = ROTL64(v0, 32); v2 += v3; v3 = ROTL64(v3, 16); \
    v3 ^= v2; v0 += v3; v3 = ROTL64(v3, 21); v3 ^= v0; v2 += v1; v1 = ROTL64(v1, 17); v1 ^= v2; v2 = ROTL64(v2, 32); } while (0)
uint64_t sbo_siphash24_u64(uint64_t k0, uint64_t k1, uint64_t m) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL, v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL, v3 = k1 ^ 0x7465646279746573ULL;
    v3 ^= m; SIPROUND; SIPROUND; v0 ^= m;
    uint64_t b = (uint64_t)8 << 56; /* length byte, no tail bytes */
    v3 ^= b; SIPROUND; SIPROUND; v0 ^= b;
    v2 ^= 0xff; SIPROUND; SIPROUND; SIPROUND; SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

void sbo_user_based_split(const uint64_t* users, size_t nnz, sbo_rng* rng, float test_fraction, uint8_t* out) {
    const uint64_t denominator = 100000;                                   /* data.rs:74 */
    uint64_t cutoff = (uint64_t)(test_fraction * (float)denominator);      /* data.rs:75 */
    uint64_t k0 = sbo_rng_gen_range(rng, 0, UINT64_MAX);                   /* data.rs:77-78 */
    uint64_t k1 = sbo_rng_gen_range(rng, 0, UINT64_MAX);
    for (size_t i = 0; i < nnz; ++i)
        out[i] = (sbo_siphash24_u64(k0, k1, users[i]) % denominator) > cutoff; /* data.rs:80-85 */
}

/* ------------------------------------------------------------------------------------------
 * model
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    size_t cap_t, D;
    /* forward saves, [T][D] each */
    float *x, *h, *c, *tc, *gf, *gi, *gg, *go, *dq;
    float* gscal;      /* [T] loss gradient g_t */
    uint32_t* neg;     /* [T] negatives */
    /* gradients in application order */
    uint32_t* rows; float* grads;      /* [3T], [3T*D] */
    uint32_t* brows; float* bgrads;    /* [2T] */
    size_t nrows, nbrows;
    float* dense_grad;                 /* [ndense] */
    float loss;
} sbo_ws;

struct sbo_model {
    sbo_hyper h;
    size_t N, D, T, ndense;
    float *E, *E_s1, *E_s2;
    float *b, *b_s1, *b_s2;
    float *dense, *dense_s1, *dense_s2; /* LSTM: W[2D][4][D] then B[4][D]; EWMA: alpha[D] */
    uint64_t num_updates;               /* Adam bias-correction counter [wyrm-recalled] */
    sbo_rng rng;                        /* master rng (Hyperparameters.rng) */
    sbo_ws ws;
};

void sbo_hyper_default(sbo_hyper* h, int model, size_t num_items, size_t max_sequence_length) {
    memset(h, 0, sizeof(*h));
    h->model = model; h->num_items = num_items; h->max_sequence_length = max_sequence_length;
    h->embedding_dim = 16; h->learning_rate = 0.01f; h->l2_penalty = 0.0f;        /* lstm.rs:60-62 */
    h->lstm_variant = SBO_LSTM_COUPLED; h->loss = SBO_LOSS_BPR; h->optimizer = SBO_OPT_ADAM; /* :63-65 */
    h->parallelism = SBO_PAR_SYNC; h->num_threads = 1; h->num_epochs = 10;        /* :66-69 */
    memset(h->seed, 42, 16);
}

static void ws_init(sbo_ws* w, size_t T, size_t D, size_t ndense) {
    memset(w, 0, sizeof(*w));
    w->cap_t = T; w->D = D;
    size_t n = T * D;
    float** arrs[] = {&w->x, &w->h, &w->c, &w->tc, &w->gf, &w->gi, &w->gg, &w->go, &w->dq};
    for (size_t i = 0; i < sizeof(arrs) / sizeof(arrs[0]); ++i) *arrs[i] = (float*)calloc(n ? n : 1, sizeof(float));
    w->gscal = (float*)calloc(T ? T : 1, sizeof(float));
    w->neg = (uint32_t*)calloc(T ? T : 1, sizeof(uint32_t));
    w->rows = (uint32_t*)calloc(3 * T + 1, sizeof(uint32_t));
    w->grads = (float*)calloc(3 * n + 1, sizeof(float));
    w->brows = (uint32_t*)calloc(2 * T + 1, sizeof(uint32_t));
    w->bgrads = (float*)calloc(2 * T + 1, sizeof(float));
    w->dense_grad = (float*)calloc(ndense ? ndense : 1, sizeof(float));
}

static void ws_free(sbo_ws* w) {
    free(w->x); free(w->h); free(w->c); free(w->tc); free(w->gf); free(w->gi); free(w->gg); free(w->go); free(w->dq);
    free(w->gscal); free(w->neg); free(w->rows); free(w->grads); free(w->brows); free(w->bgrads); free(w->dense_grad);
}

/* standard normal from the xorshift stream (Box-Muller); the reference uses rand's ziggurat, whose
 * stream is not reproducible here -- parity tests inject parameters instead. */
static double rng_uniform01(sbo_rng* r) { return ((double)(sbo_rng_next_u64(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double rng_normal(sbo_rng* r) {
    double u1 = rng_uniform01(r), u2 = rng_uniform01(r);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

sbo_model* sbo_model_new(const sbo_hyper* h) {
    sbo_model* m = (sbo_model*)calloc(1, sizeof(sbo_model));
    m->h = *h; m->N = h->num_items; m->D = h->embedding_dim; m->T = h->max_sequence_length;
    size_t N = m->N, D = m->D;
    m->ndense = h->model == SBO_MODEL_LSTM ? (2 * D * 4 * D + 4 * D) : D;
    sbo_rng_from_seed(&m->rng, h->seed);
    m->E = (float*)malloc(sizeof(float) * N * D);
    m->E_s1 = (float*)calloc(N * D, sizeof(float)); m->E_s2 = (float*)calloc(N * D, sizeof(float));
    m->b = (float*)calloc(N, sizeof(float)); m->b_s1 = (float*)calloc(N, sizeof(float)); m->b_s2 = (float*)calloc(N, sizeof(float));
    m->dense = (float*)calloc(m->ndense, sizeof(float));
    m->dense_s1 = (float*)calloc(m->ndense, sizeof(float)); m->dense_s2 = (float*)calloc(m->ndense, sizeof(float));
    /* embedding_init: Normal(0, 1/D)  (lstm.rs:22-25, ewma.rs:33-36: std = 1/cols) */
    for (size_t i = 0; i < N * D; ++i) m->E[i] = (float)(rng_normal(&m->rng) / (double)D);
    if (h->model == SBO_MODEL_LSTM) { /* wyrm nn::lstm::Parameters::new [wyrm-recalled]: U(-1/sqrt(D), 1/sqrt(D)) */
        double a = 1.0 / sqrt((double)D);
        for (size_t i = 0; i < m->ndense; ++i) m->dense[i] = (float)((2.0 * rng_uniform01(&m->rng) - 1.0) * a);
    } /* EWMA alpha = 0 (ewma.rs:175-178); biases = 0 (lstm.rs:181) */
    ws_init(&m->ws, m->T, D, m->ndense);
    return m;
}

void sbo_model_free(sbo_model* m) {
    if (!m) return;
    free(m->E); free(m->E_s1); free(m->E_s2); free(m->b); free(m->b_s1); free(m->b_s2);
    free(m->dense); free(m->dense_s1); free(m->dense_s2); ws_free(&m->ws); free(m);
}

float* sbo_model_param(sbo_model* m, const char* name, size_t* len) {
    size_t N = m->N, D = m->D; float* p = NULL; size_t n = 0;
    char base[64]; int which = 0;
    strncpy(base, name, 63); base[63] = 0;
    char* dot = strrchr(base, '.');
    if (dot && (!strcmp(dot, ".s1") || !strcmp(dot, ".s2"))) { which = dot[2] - '0'; *dot = 0; }
    if (!strcmp(base, "item_embeddings")) { p = which == 0 ? m->E : which == 1 ? m->E_s1 : m->E_s2; n = N * D; }
    else if (!strcmp(base, "item_biases")) { p = which == 0 ? m->b : which == 1 ? m->b_s1 : m->b_s2; n = N; }
    else if (m->h.model == SBO_MODEL_LSTM && !strcmp(base, "lstm_weights")) {
        p = which == 0 ? m->dense : which == 1 ? m->dense_s1 : m->dense_s2; n = 2 * D * 4 * D; }
    else if (m->h.model == SBO_MODEL_LSTM && !strcmp(base, "lstm_biases")) {
        p = (which == 0 ? m->dense : which == 1 ? m->dense_s1 : m->dense_s2) + 2 * D * 4 * D; n = 4 * D; }
    else if (m->h.model == SBO_MODEL_EWMA && !strcmp(base, "alpha")) {
        p = which == 0 ? m->dense : which == 1 ? m->dense_s1 : m->dense_s2; n = D; }
    if (len) *len = n;
    return p;
}

uint64_t* sbo_model_num_updates(sbo_model* m) { return &m->num_updates; }
sbo_rng* sbo_model_rng(sbo_model* m) { return &m->rng; }

/* ndarray unrolled_dot / wyrm simd_dot [wyrm-recalled]: 8 partial sums, pairwise combine, scalar tail */
static float dot8(const float* x, const float* y, size_t n) {
    float p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0, sum = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        p0 += x[i] * y[i]; p1 += x[i + 1] * y[i + 1]; p2 += x[i + 2] * y[i + 2]; p3 += x[i + 3] * y[i + 3];
        p4 += x[i + 4] * y[i + 4]; p5 += x[i + 5] * y[i + 5]; p6 += x[i + 6] * y[i + 6]; p7 += x[i + 7] * y[i + 7];
    }
    sum += p0 + p4; sum += p1 + p5; sum += p2 + p6; sum += p3 + p7;
    for (; i < n; ++i) sum += x[i] * y[i];
    return sum;
}

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); } /* exact libm; reference uses fastexp */

/* lstm.rs:338-350 / ewma.rs:353-365 */
static inline float predict_single(const sbo_model* m, const float* user, size_t item) {
    return m->b[item] + dot8(user, m->E + item * m->D, m->D);
}

/* ---- optimizers [wyrm-recalled] ---- */
static inline void adagrad_elem(float* w, float* G, float g, float lr, float l2) {
    g = g + *w * l2;                       /* gradient + value * l2 */
    *G += g * g;
    *w -= lr / (1e-10f + sqrtf(*G)) * g;   /* eps = 1e-10 */
}
static inline void adam_elem(float* w, float* mm, float* vv, float g, float lr, float l2, float c1, float c2) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    g = g + *w * l2;
    *mm = b1 * *mm + (1.0f - b1) * g;
    *vv = b2 * *vv + (1.0f - b2) * g * g;
    float mhat = *mm / c1, vhat = *vv / c2; /* c1 = 1 - b1^t, c2 = 1 - b2^t */
    *w -= lr / (sqrtf(vhat) + eps) * mhat;
}

static void adam_coeffs(const sbo_model* m, uint64_t t, float* c1, float* c2) {
    *c1 = 1.0f; *c2 = 1.0f;
    if (m->h.optimizer == SBO_OPT_ADAM) { *c1 = 1.0f - powf(0.9f, (float)t); *c2 = 1.0f - powf(0.999f, (float)t); }
}

/* Experiment switch (profiles/tools/oracle_mrr_seeds.py --merge): what if wyrm's gradient accumulator SUMS the entries
 * of a row that is recorded several times in one step (a dense-shaped gradient buffer plus a set of touched rows) and
 * the optimizer visits every touched row once, in ascending row order?  The source is not on this box; the default (0)
 * is the un-merged list the survey recalled.  DESIGN.md 5 reports what the MRR floors say about the two readings. */
static int g_merge_sparse = 0;
void sbo_set_merge_sparse(int on) { g_merge_sparse = on; }

static int cmp_u32_idx(const void* a, const void* b) {
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}
static void apply_sparse_merged(sbo_model* m, const sbo_ws* w, float c1, float c2) {
    const size_t D = m->D; const float lr = m->h.learning_rate, l2 = m->h.l2_penalty;
    const int adam = m->h.optimizer == SBO_OPT_ADAM;
    uint64_t* key = (uint64_t*)malloc(sizeof(uint64_t) * (w->nrows + w->nbrows + 1));
    float* acc = (float*)malloc(sizeof(float) * D);
    for (size_t e = 0; e < w->nrows; ++e) key[e] = ((uint64_t)w->rows[e] << 32) | e;   /* (row, recording order) */
    qsort(key, w->nrows, sizeof(uint64_t), cmp_u32_idx);
    for (size_t a = 0; a < w->nrows;) {
        const size_t r = (size_t)(key[a] >> 32);
        memset(acc, 0, sizeof(float) * D);
        size_t b = a;
        for (; b < w->nrows && (size_t)(key[b] >> 32) == r; ++b) {
            const float* g = w->grads + (size_t)(key[b] & 0xffffffffu) * D;
            for (size_t d = 0; d < D; ++d) acc[d] += g[d];
        }
        float* wv = m->E + r * D; float* s1 = m->E_s1 + r * D; float* s2 = m->E_s2 + r * D;
        for (size_t d = 0; d < D; ++d) {
            if (adam) adam_elem(wv + d, s1 + d, s2 + d, acc[d], lr, l2, c1, c2);
            else adagrad_elem(wv + d, s1 + d, acc[d], lr, l2);
        }
        a = b;
    }
    for (size_t e = 0; e < w->nbrows; ++e) key[e] = ((uint64_t)w->brows[e] << 32) | e;
    qsort(key, w->nbrows, sizeof(uint64_t), cmp_u32_idx);
    for (size_t a = 0; a < w->nbrows;) {
        const size_t r = (size_t)(key[a] >> 32);
        float g = 0.0f; size_t b = a;
        for (; b < w->nbrows && (size_t)(key[b] >> 32) == r; ++b) g += w->bgrads[key[b] & 0xffffffffu];
        if (adam) adam_elem(m->b + r, m->b_s1 + r, m->b_s2 + r, g, lr, l2, c1, c2);
        else adagrad_elem(m->b + r, m->b_s1 + r, g, lr, l2);
        a = b;
    }
    free(key); free(acc);
}

/* sparse rows: one update per recorded (row, grad) entry, in order, duplicates NOT merged [wyrm-recalled] */
static void apply_sparse(sbo_model* m, const sbo_ws* w, float c1, float c2) {
    if (g_merge_sparse) { apply_sparse_merged(m, w, c1, c2); return; }
    const size_t D = m->D; const float lr = m->h.learning_rate, l2 = m->h.l2_penalty;
    const int adam = m->h.optimizer == SBO_OPT_ADAM;
    for (size_t e = 0; e < w->nrows; ++e) {
        size_t r = w->rows[e]; const float* g = w->grads + e * D;
        float* wv = m->E + r * D; float* s1 = m->E_s1 + r * D; float* s2 = m->E_s2 + r * D;
        for (size_t d = 0; d < D; ++d) {
            if (adam) adam_elem(wv + d, s1 + d, s2 + d, g[d], lr, l2, c1, c2);
            else adagrad_elem(wv + d, s1 + d, g[d], lr, l2);
        }
    }
    for (size_t e = 0; e < w->nbrows; ++e) {
        size_t r = w->brows[e];
        if (adam) adam_elem(m->b + r, m->b_s1 + r, m->b_s2 + r, w->bgrads[e], lr, l2, c1, c2);
        else adagrad_elem(m->b + r, m->b_s1 + r, w->bgrads[e], lr, l2);
    }
}

/* dense parameters (LSTM weights+biases / alpha): every element, every step */
static void apply_dense(sbo_model* m, const float* grad, float c1, float c2) {
    const float lr = m->h.learning_rate, l2 = m->h.l2_penalty;
    const int adam = m->h.optimizer == SBO_OPT_ADAM;
    for (size_t i = 0; i < m->ndense; ++i) {
        if (adam) adam_elem(m->dense + i, m->dense_s1 + i, m->dense_s2 + i, grad[i], lr, l2, c1, c2);
        else adagrad_elem(m->dense + i, m->dense_s1 + i, grad[i], lr, l2);
    }
}

static void apply_update(sbo_model* m, sbo_ws* w) {
    float c1, c2;
    uint64_t t = ++m->num_updates; /* per-parameter num_updates, one per step() [wyrm-recalled] */
    adam_coeffs(m, t, &c1, &c2);
    apply_sparse(m, w, c1, c2);
    apply_dense(m, w->dense_grad, c1, c2);
}

/* ---- recurrent cells ---- */
/* One LSTM cell step (wyrm::nn::lstm [wyrm-recalled]; z = [h_prev, x], gate order f,i,g,o). */
static void lstm_cell(const sbo_model* m, const float* hprev, const float* cprev, const float* x,
                      float* f, float* i_, float* g, float* o, float* c, float* tc, float* h) {
    const size_t D = m->D; const float* W = m->dense; const float* B = m->dense + 2 * D * 4 * D;
    float pre[4][512];
    for (int q = 0; q < 4; ++q) for (size_t d = 0; d < D; ++d) pre[q][d] = B[q * D + d];
    for (size_t k = 0; k < 2 * D; ++k) {
        float zk = k < D ? hprev[k] : x[k - D];
        const float* Wk = W + k * 4 * D;
        for (int q = 0; q < 4; ++q) for (size_t d = 0; d < D; ++d) pre[q][d] = fmaf(zk, Wk[q * D + d], pre[q][d]);
    }
    const int coupled = m->h.lstm_variant == SBO_LSTM_COUPLED;
    for (size_t d = 0; d < D; ++d) {
        f[d] = sigmoidf_(pre[0][d]);
        i_[d] = coupled ? 1.0f - f[d] : sigmoidf_(pre[1][d]);
        g[d] = tanhf(pre[2][d]);
        o[d] = sigmoidf_(pre[3][d]);
        c[d] = f[d] * cprev[d] + i_[d] * g[d];
        tc[d] = tanhf(c[d]);
        h[d] = o[d] * tc[d];
    }
}

static inline void ewma_cell(const float* a, size_t D, int first, const float* sprev, const float* x, float* s) {
    for (size_t d = 0; d < D; ++d) s[d] = first ? x[d] : a[d] * sprev[d] + (1.0f - a[d]) * x[d]; /* ewma.rs:302-313 */
}

/* The body of the hot loop, sequence_model.rs:111-169, for one sub-sequence. */
static float step_ws(sbo_model* m, sbo_ws* w, const uint64_t* ids, size_t len, uint64_t key, uint64_t step,
                     const uint32_t* forced_neg, int compute_grad) {
    const size_t D = m->D, N = m->N;
    const size_t Tn = len - 1; /* izip stops at item_ids.skip(1): len-1 timesteps */
    const int lstm = m->h.model == SBO_MODEL_LSTM, coupled = m->h.lstm_variant == SBO_LSTM_COUPLED;
    const int loss_kind = m->h.loss;
    float zero[512]; memset(zero, 0, sizeof(zero));
    float a[512];
    if (!lstm) for (size_t d = 0; d < D; ++d) a[d] = sigmoidf_(m->dense[d]); /* ewma.rs:302 */

    float total = 0.0f;
    for (size_t t = 0; t < Tn; ++t) {
        size_t in = ids[t], out = ids[t + 1];
        float* x = w->x + t * D; float* h = w->h + t * D;
        memcpy(x, m->E + in * D, sizeof(float) * D); /* item_embeddings.index(input) -- exact copy */
        const float* hprev = t ? w->h + (t - 1) * D : zero;
        if (lstm) {
            const float* cprev = t ? w->c + (t - 1) * D : zero;
            lstm_cell(m, hprev, cprev, x, w->gf + t * D, w->gi + t * D, w->gg + t * D, w->go + t * D, w->c + t * D,
                      w->tc + t * D, h);
        } else {
            ewma_cell(a, D, t == 0, hprev, x, h);
        }
        /* negative: uniform, or WARP rejection sampling (sequence_model.rs:47-68,125-138) */
        uint32_t neg;
        if (forced_neg) neg = forced_neg[t];
        else if (loss_kind == SBO_LOSS_WARP) {
            float pos_pred = predict_single(m, h, out);
            neg = 0;
            for (uint32_t j = 0; j < 5; ++j) {
                neg = sbo_draw_item(key, step, (uint32_t)t, j, (uint32_t)N);
                float neg_pred = predict_single(m, h, neg);
                if (1.0f - pos_pred + neg_pred > 0.0f) break;
            }
        } else neg = sbo_draw_item(key, step, (uint32_t)t, 0, (uint32_t)N);
        w->neg[t] = neg;
        const float* p = m->E + out * D; const float* q = m->E + (size_t)neg * D;
        float pos = dot8(h, p, D) + m->b[out];           /* lstm.rs:300-305 */
        float ng = dot8(h, q, D) + m->b[neg];            /* lstm.rs:306-311 */
        float l, g;
        if (loss_kind == SBO_LOSS_BPR) { float s = sigmoidf_(ng - pos); l = s; g = s * (1.0f - s); } /* lstm.rs:317 */
        else { float v = 1.0f + ng - pos; l = v > 0.0f ? v : 0.0f; g = v > 0.0f ? 1.0f : 0.0f; }      /* lstm.rs:318 */
        total += l;                                      /* summed_losses, lstm.rs:322-328 */
        w->gscal[t] = g;
        float* dq = w->dq + t * D;
        for (size_t d = 0; d < D; ++d) dq[d] = g * (q[d] - p[d]);
    }
    w->loss = total;
    if (!compute_grad) return total;

    /* backward (seed 1.0, sequence_model.rs:161); entries recorded t descending: E[neg], E[out], E[in] */
    memset(w->dense_grad, 0, sizeof(float) * m->ndense);
    w->nrows = 0; w->nbrows = 0;
    float dh_rec[512], dc_rec[512], dh[512], dx[512], del[4][512], da[512];
    memset(dh_rec, 0, sizeof(dh_rec)); memset(dc_rec, 0, sizeof(dc_rec)); memset(da, 0, sizeof(da));
    const float* W = m->dense; float* dW = w->dense_grad; float* dB = w->dense_grad + 2 * D * 4 * D;
    for (size_t tt = Tn; tt-- > 0;) {
        const size_t t = tt;
        size_t in = ids[t], out = ids[t + 1]; uint32_t neg = w->neg[t];
        const float* h = w->h + t * D; const float* x = w->x + t * D; const float g = w->gscal[t];
        const float* hprev = t ? w->h + (t - 1) * D : zero;
        for (size_t d = 0; d < D; ++d) dh[d] = dh_rec[d] + w->dq[t * D + d];
        if (lstm) {
            const float* cprev = t ? w->c + (t - 1) * D : zero;
            const float *f = w->gf + t * D, *ii = w->gi + t * D, *gg = w->gg + t * D, *o = w->go + t * D, *tc = w->tc + t * D;
            for (size_t d = 0; d < D; ++d) {
                float d_o = dh[d] * tc[d];
                float dc = dc_rec[d] + dh[d] * o[d] * (1.0f - tc[d] * tc[d]);
                float d_f = dc * cprev[d], d_i = dc * gg[d], d_g = dc * ii[d];
                dc_rec[d] = dc * f[d];
                if (coupled) { d_f -= d_i; d_i = 0.0f; }          /* i = 1 - f */
                del[0][d] = d_f * f[d] * (1.0f - f[d]);
                del[1][d] = coupled ? 0.0f : d_i * ii[d] * (1.0f - ii[d]);
                del[2][d] = d_g * (1.0f - gg[d] * gg[d]);
                del[3][d] = d_o * o[d] * (1.0f - o[d]);
            }
            for (size_t k = 0; k < 2 * D; ++k) {
                float zk = k < D ? hprev[k] : x[k - D];
                const float* Wk = W + k * 4 * D; float* dWk = dW + k * 4 * D;
                float acc = 0.0f;
                for (int q = 0; q < 4; ++q) for (size_t d = 0; d < D; ++d) {
                    dWk[q * D + d] = fmaf(zk, del[q][d], dWk[q * D + d]);
                    acc = fmaf(del[q][d], Wk[q * D + d], acc);
                }
                if (k < D) dh_rec[k] = acc; else dx[k - D] = acc;
            }
            for (int q = 0; q < 4; ++q) for (size_t d = 0; d < D; ++d) dB[q * D + d] += del[q][d];
        } else {
            for (size_t d = 0; d < D; ++d) {
                if (t == 0) { dx[d] = dh[d]; dh_rec[d] = 0.0f; }
                else { dx[d] = (1.0f - a[d]) * dh[d]; da[d] += dh[d] * (hprev[d] - x[d]); dh_rec[d] = a[d] * dh[d]; }
            }
        }
        size_t e = w->nrows;
        w->rows[e] = neg; for (size_t d = 0; d < D; ++d) w->grads[e * D + d] = g * h[d];
        w->rows[e + 1] = (uint32_t)out; for (size_t d = 0; d < D; ++d) w->grads[(e + 1) * D + d] = -g * h[d];
        w->rows[e + 2] = (uint32_t)in; memcpy(w->grads + (e + 2) * D, dx, sizeof(float) * D);
        w->nrows += 3;
        size_t be = w->nbrows;
        w->brows[be] = neg; w->bgrads[be] = g; w->brows[be + 1] = (uint32_t)out; w->bgrads[be + 1] = -g;
        w->nbrows += 2;
    }
    if (!lstm) for (size_t d = 0; d < D; ++d) w->dense_grad[d] = da[d] * a[d] * (1.0f - a[d]);
    return total;
}

float sbo_step(sbo_model* m, const uint64_t* ids, size_t len, uint64_t key, uint64_t step, uint32_t* negatives_out,
               int apply, float* dense_grad_out, const uint32_t* forced_negatives) {
    if (len < 2 || len > m->T) return NAN;
    float l = step_ws(m, &m->ws, ids, len, key, step, forced_negatives, 1);
    if (negatives_out) memcpy(negatives_out, m->ws.neg, sizeof(uint32_t) * (len - 1));
    if (dense_grad_out) memcpy(dense_grad_out, m->ws.dense_grad, sizeof(float) * m->ndense);
    if (apply) apply_update(m, &m->ws);
    return l;
}

double sbo_loss_only(sbo_model* m, const uint64_t* ids, size_t len, const uint32_t* negatives) {
    return (double)step_ws(m, &m->ws, ids, len, 0, 0, negatives, 0);
}

size_t sbo_last_sparse_grads(sbo_model* m, const uint32_t** rows, const float** grads, const uint32_t** brows,
                             const float** bgrads, size_t* nb) {
    *rows = m->ws.rows; *grads = m->ws.grads; *brows = m->ws.brows; *bgrads = m->ws.bgrads; *nb = m->ws.nbrows;
    return m->ws.nrows;
}

/* ------------------------------------------------------------------------------------------
 * fit: sequence_model.rs:70-178
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    sbo_model* m; sbo_ws* ws; int p, P;
    const uint64_t* item_ids; const uint64_t* starts; const uint32_t* lens;
    uint32_t* order; size_t n; uint8_t seed[16];
    pthread_barrier_t* barrier; void* all;
    float loss_value; size_t examples;
} sbo_part_t;

static void* run_partition(void* arg) {
    sbo_part_t* pt = (sbo_part_t*)arg; sbo_model* m = pt->m;
    sbo_rng rng; sbo_rng_from_seed(&rng, pt->seed);                 /* :97 */
    uint64_t key = 0; for (int i = 0; i < 8; ++i) key |= (uint64_t)pt->seed[i] << (8 * i);
    const int sync = pt->P > 1 && m->h.parallelism == SBO_PAR_SYNC;  /* :163-165 */
    sbo_part_t* all = (sbo_part_t*)pt->all;
    uint64_t step = 0; float loss_value = 0.0f; size_t examples = 0;
    for (int epoch = 0; epoch < m->h.num_epochs; ++epoch) {         /* :108 */
        sbo_shuffle_u32(&rng, pt->order, pt->n);                    /* :109 */
        for (size_t i = 0; i < pt->n; ++i) {                        /* :111 */
            uint32_t s = pt->order[i];
            float l = step_ws(m, pt->ws, pt->item_ids + pt->starts[s], pt->lens[s], key, step++, NULL, 1);
            loss_value += l; examples += pt->lens[s] - 1;           /* :157-158 (fresh, not stale, value) */
            if (sync) { /* barrier-coupled optimizer [wyrm-recalled]: every thread's gradients come from the round-start
                           parameters; the sparse entries are applied one thread at a time, in order, un-merged; the
                           dense parameters take ONE step on the gradient summed over the round (the engine's documented
                           synchronous semantics) */
                pthread_barrier_wait(pt->barrier);
                if (pt->p == 0) {
                    float c1, c2;
                    m->num_updates += (uint64_t)pt->P;
                    adam_coeffs(m, m->num_updates, &c1, &c2);
                    float* sum = all[0].ws->dense_grad;
                    for (int q = 1; q < pt->P; ++q) for (size_t k = 0; k < m->ndense; ++k) sum[k] += all[q].ws->dense_grad[k];
                    for (int q = 0; q < pt->P; ++q) apply_sparse(m, all[q].ws, c1, c2);
                    apply_dense(m, sum, c1, c2);
                }
                pthread_barrier_wait(pt->barrier);
            } else apply_update(m, pt->ws);                         /* :168 (Hogwild when P > 1) */
        }
    }
    pt->loss_value = loss_value; pt->examples = examples;
    return NULL;
}

int sbo_fit(sbo_model* m, const uint64_t* user_ptr, const uint64_t* item_ids, size_t num_users, float* loss_out) {
    size_t n = sbo_subsequences(user_ptr, num_users, m->T, NULL, NULL);
    if (n == 0) return SBO_ERR_NO_INTERACTIONS;                     /* :86-88 */
    uint64_t* starts = (uint64_t*)malloc(sizeof(uint64_t) * n);
    uint32_t* lens = (uint32_t*)malloc(sizeof(uint32_t) * n);
    sbo_subsequences(user_ptr, num_users, m->T, starts, lens);
    uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    sbo_shuffle_u32(&m->rng, order, n);                             /* :84 */
    int P = m->h.num_threads;
    if (P < 1 || (size_t)P > n) { free(starts); free(lens); free(order); return SBO_ERR_INVALID_ARGUMENT; } /* :95 panics */
    size_t chunk = n / (size_t)P;                                   /* :91; remainder dropped by zip (:94-96) */
    sbo_part_t* parts = (sbo_part_t*)calloc(P, sizeof(sbo_part_t));
    sbo_ws* wss = (sbo_ws*)calloc(P, sizeof(sbo_ws));
    pthread_barrier_t barrier; pthread_barrier_init(&barrier, NULL, P);
    for (int p = 0; p < P; ++p) {
        ws_init(&wss[p], m->T, m->D, m->ndense);
        parts[p].m = m; parts[p].ws = &wss[p]; parts[p].p = p; parts[p].P = P;
        parts[p].item_ids = item_ids; parts[p].starts = starts; parts[p].lens = lens;
        parts[p].order = order + (size_t)p * chunk; parts[p].n = chunk;
        parts[p].barrier = &barrier; parts[p].all = parts;
        sbo_rng_gen_seed(&m->rng, parts[p].seed);                   /* :97 */
    }
    if (P == 1) run_partition(&parts[0]);
    else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * P);  /* :100-101 par_iter_mut */
        for (int p = 0; p < P; ++p) pthread_create(&th[p], NULL, run_partition, &parts[p]);
        for (int p = 0; p < P; ++p) pthread_join(th[p], NULL);
        free(th);
    }
    float loss = 0.0f;
    for (int p = 0; p < P; ++p) { loss += parts[p].loss_value / (1.0f + (float)parts[p].examples); ws_free(&wss[p]); } /* :173-175 */
    pthread_barrier_destroy(&barrier);
    free(parts); free(wss); free(starts); free(lens); free(order);
    if (loss_out) *loss_out = loss;
    return SBO_OK;
}

/* ------------------------------------------------------------------------------------------
 * inference: sequence_model.rs:180-233, evaluation.rs:12-48
 * ---------------------------------------------------------------------------------------- */
int sbo_user_representation(sbo_model* m, const uint64_t* ids, size_t n, float* out) {
    const size_t D = m->D;
    uint64_t zero_id = 0;
    if (n > m->T) { ids += n - m->T; n = m->T; }         /* :188 last max_sequence_length ids */
    if (n == 0) { ids = &zero_id; n = 1; }               /* :197-200: hidden_states[0] with the default index 0 */
    float h[512], c[512], h2[512], c2[512], x[512], a[512], f[512], i_[512], g[512], o[512], tc[512];
    memset(h, 0, sizeof(h)); memset(c, 0, sizeof(c));
    const int lstm = m->h.model == SBO_MODEL_LSTM;
    if (!lstm) for (size_t d = 0; d < D; ++d) a[d] = sigmoidf_(m->dense[d]);
    for (size_t t = 0; t < n; ++t) {
        if (ids[t] >= m->N) return SBO_ERR_INVALID_ARGUMENT;
        memcpy(x, m->E + ids[t] * D, sizeof(float) * D);
        if (lstm) { lstm_cell(m, h, c, x, f, i_, g, o, c2, tc, h2); memcpy(c, c2, sizeof(float) * D); }
        else ewma_cell(a, D, t == 0, h, x, h2);
        memcpy(h, h2, sizeof(float) * D);
    }
    memcpy(out, h, sizeof(float) * D);
    return SBO_OK;
}

int sbo_predict(sbo_model* m, const float* user, const uint64_t* ids, size_t k, float* out) {
    for (size_t i = 0; i < k; ++i) {
        if (ids[i] >= m->N) return SBO_ERR_INVALID_ARGUMENT;
        float v = predict_single(m, user, ids[i]);       /* :220-223 */
        if (!isfinite(v)) return SBO_ERR_INVALID_PREDICTION; /* :225-229 */
        out[i] = v;
    }
    return SBO_OK;
}

int sbo_mrr_score(sbo_model* m, const uint64_t* user_ptr, const uint64_t* item_ids, size_t num_users, float* out) {
    const size_t N = m->N;
    float* pred = (float*)malloc(sizeof(float) * N);
    float user[512];
    double sum = 0.0; /* reference: f32 vector then sum()/len; double here only narrows rounding noise */
    float fsum = 0.0f; size_t cnt = 0;
    for (size_t u = 0; u < num_users; ++u) {
        size_t b = user_ptr[u], len = user_ptr[u + 1] - b;
        if (len < 2) continue;                                           /* evaluation.rs:20 */
        const uint64_t* ids = item_ids + b;
        size_t test_item = ids[len - 1];                                 /* :25 */
        int rc = sbo_user_representation(m, ids, len - 1, user);         /* :24,27 */
        if (rc) { free(pred); return rc; }
        for (size_t j = 0; j < N; ++j) {                                 /* :16,28 */
            pred[j] = predict_single(m, user, j);
            if (!isfinite(pred[j])) { free(pred); return SBO_ERR_INVALID_PREDICTION; }
        }
        for (size_t j = 0; j + 1 < len; ++j) pred[ids[j]] = -3.40282347e+38f; /* f32::MIN, :30-32 */
        float ts = pred[test_item]; size_t rank = 0;
        for (size_t j = 0; j < N; ++j) if (pred[j] >= ts) rank++;        /* :35-41 */
        fsum += 1.0f / (float)rank; sum += 1.0 / (double)rank; cnt++;    /* :43 */
    }
    free(pred);
    (void)sum;
    *out = cnt ? fsum / (float)cnt : NAN;                                /* :47 */
    return SBO_OK;
}
