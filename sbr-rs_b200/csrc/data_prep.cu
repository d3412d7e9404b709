// data_prep.cu -- Interactions::to_compressed on the device (data.rs:236-265).
//
// The reference sorts the interaction triples with a STABLE sort on (user, timestamp) (data.rs:240, comparator
// data.rs:213-221), counts interactions per user (data.rs:250) and prefix-sums the counts (data.rs:253-255).  ML-100K
// has 50,561 tied (user, timestamp) pairs, so stability decides the item order inside a sequence and with it every
// trained sequence.  On the device: two passes of CUB's LSD radix sort (stable by construction) -- first by timestamp,
// then by user -- over (key, original index) pairs, a gather of the item ids / timestamps through the final
// permutation, a histogram of users with 64-bit atomics and an inclusive scan.  All HBM-bound integer work:
//   algorithmic bytes per interaction ~ 3 x 8 B in, 2 sorts x (8 B key + 4 B value) x 8 radix passes, 2 x 8 + 4 B out.
// The host never sorts; it receives the finished CSR (and the narrowed item-id stream stays resident for fit()).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdint>
#include <string>

namespace sbr {

namespace {

__global__ void prep_validate_iota_kernel(const uint64_t* __restrict__ user, const uint64_t* __restrict__ item, size_t nnz, uint64_t num_users,
                                          uint64_t num_items, uint32_t* __restrict__ idx, int* __restrict__ bad) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nnz; i += (size_t)gridDim.x * blockDim.x) {
        idx[i] = (uint32_t)i;
        if (user[i] >= num_users) atomicOr(bad, 1);
        if (item[i] >= num_items) atomicOr(bad, 2);
    }
}
__global__ void prep_gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, size_t nnz, uint64_t* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nnz; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
// the CSR payload through the final permutation + the per-user histogram (user_ptr[u + 1] += 1, data.rs:250)
__global__ void prep_finish_kernel(const uint64_t* __restrict__ item, const uint64_t* __restrict__ ts, const uint64_t* __restrict__ sorted_user,
                                   const uint32_t* __restrict__ perm, size_t nnz, uint64_t* __restrict__ item_out, uint64_t* __restrict__ ts_out,
                                   uint32_t* __restrict__ item_u32_out, unsigned long long* __restrict__ user_ptr) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nnz; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t s = perm[i];
        const uint64_t it = item[s];
        item_out[i] = it; ts_out[i] = ts[s]; item_u32_out[i] = (uint32_t)it;
        atomicAdd(user_ptr + sorted_user[i] + 1, 1ull);
    }
}

struct Tmp {   // everything that is freed on every exit path
    void* p[16] = {};
    int n = 0;
    template <typename T> cudaError_t alloc(T** out, size_t bytes) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes ? bytes : 1);
        if (e == cudaSuccess) { p[n++] = q; *out = static_cast<T*>(q); }
        return e;
    }
    ~Tmp() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
};

}  // namespace

// Builds the CSR of `nnz` host triples on the device.  Host outputs: user_ptr[num_users + 1], item_out[nnz], ts_out[nnz].
// Device outputs (caller-allocated): d_item_u32[nnz] (the narrowed id stream fit() reads), d_user_ptr[num_users + 1].
// returns 0 ok, 1 CUDA error (*err), 2 an id out of range (*err)
int device_csr_build(const uint64_t* h_user, const uint64_t* h_item, const uint64_t* h_ts, size_t nnz, size_t num_users, size_t num_items,
                     uint64_t* h_user_ptr, uint64_t* h_item_out, uint64_t* h_ts_out, uint32_t* d_item_u32, uint64_t* d_user_ptr,
                     cudaStream_t st, std::string* err) {
#define DCU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { *err = std::string(#expr) + ": " + cudaGetErrorString(e__); return 1; } } while (0)
    if (nnz >= ((size_t)1 << 31)) { *err = "device CSR build handles fewer than 2^31 interactions per call"; return 2; }
    Tmp tmp;
    uint64_t *d_user, *d_item, *d_ts, *key_a, *key_b, *d_item_out, *d_ts_out;
    uint32_t *idx_a, *idx_b; int* d_bad;
    const size_t B8 = nnz * sizeof(uint64_t), B4 = nnz * sizeof(uint32_t);
    DCU(tmp.alloc(&d_user, B8)); DCU(tmp.alloc(&d_item, B8)); DCU(tmp.alloc(&d_ts, B8));
    DCU(tmp.alloc(&key_a, B8)); DCU(tmp.alloc(&key_b, B8));
    DCU(tmp.alloc(&idx_a, B4)); DCU(tmp.alloc(&idx_b, B4)); DCU(tmp.alloc(&d_bad, sizeof(int)));
    d_item_out = d_user;   // d_user is dead after the second sort's key gather: reuse for the outputs
    d_ts_out = key_a;
    DCU(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    DCU(cudaMemsetAsync(d_user_ptr, 0, (num_users + 1) * sizeof(uint64_t), st));
    if (nnz) {
        DCU(cudaMemcpyAsync(d_user, h_user, B8, cudaMemcpyHostToDevice, st));
        DCU(cudaMemcpyAsync(d_item, h_item, B8, cudaMemcpyHostToDevice, st));
        DCU(cudaMemcpyAsync(d_ts, h_ts, B8, cudaMemcpyHostToDevice, st));
        const int threads = 256;
        const int blocks = (int)std::min<size_t>((nnz + threads - 1) / threads, 148 * 16);
        prep_validate_iota_kernel<<<blocks, threads, 0, st>>>(d_user, d_item, nnz, num_users, num_items, idx_a, d_bad);
        DCU(cudaGetLastError());
        int bad = 0;
        DCU(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        DCU(cudaStreamSynchronize(st));
        if (bad) { *err = (bad & 1) ? "user id >= num_users" : "item id >= num_items"; return 2; }
        // pass 1: stable sort by timestamp        (key_a <- ts, sorted into key_b; idx_a -> idx_b)
        size_t cub_bytes = 0;
        DCU(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, d_ts, key_b, idx_a, idx_b, (int)nnz, 0, 64, st));
        void* cub_tmp = nullptr;
        DCU(tmp.alloc(&cub_tmp, cub_bytes));
        DCU(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, d_ts, key_b, idx_a, idx_b, (int)nnz, 0, 64, st));
        // pass 2: stable sort by user             (key_a <- user[idx_b], sorted into key_b; idx_b -> idx_a = permutation)
        prep_gather_u64_kernel<<<blocks, threads, 0, st>>>(d_user, idx_b, nnz, key_a);
        DCU(cudaGetLastError());
        size_t cub_bytes2 = 0;
        DCU(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes2, key_a, key_b, idx_b, idx_a, (int)nnz, 0, 64, st));
        if (cub_bytes2 > cub_bytes) { DCU(tmp.alloc(&cub_tmp, cub_bytes2)); cub_bytes = cub_bytes2; }
        DCU(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, key_a, key_b, idx_b, idx_a, (int)nnz, 0, 64, st));
        // payload through the permutation + histogram      (key_b = users in sorted order)
        prep_finish_kernel<<<blocks, threads, 0, st>>>(d_item, d_ts, key_b, idx_a, nnz, d_item_out, d_ts_out, d_item_u32,
                                                       reinterpret_cast<unsigned long long*>(d_user_ptr));
        DCU(cudaGetLastError());
        DCU(cudaMemcpyAsync(h_item_out, d_item_out, B8, cudaMemcpyDeviceToHost, st));
        DCU(cudaMemcpyAsync(h_ts_out, d_ts_out, B8, cudaMemcpyDeviceToHost, st));
    }
    {   // user_ptr[u] = sum of counts below u  (data.rs:253-255): inclusive scan of the shifted histogram, in place
        size_t scan_bytes = 0;
        DCU(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, d_user_ptr, d_user_ptr, (int)(num_users + 1), st));
        void* scan_tmp = nullptr;
        DCU(tmp.alloc(&scan_tmp, scan_bytes));
        DCU(cub::DeviceScan::InclusiveSum(scan_tmp, scan_bytes, d_user_ptr, d_user_ptr, (int)(num_users + 1), st));
    }
    DCU(cudaMemcpyAsync(h_user_ptr, d_user_ptr, (num_users + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    DCU(cudaStreamSynchronize(st));
    return 0;
#undef DCU
}

}  // namespace sbr

// ============================================================================================================
// sequence_model.rs:76-81 on the device: the sub-sequences of every user (data.rs:406-432: the FIRST chunk is the short
// one -- len % T items, if that is not 0 -- every later chunk has exactly T items), filtered to len > 2 (:81), in user
// order.  Three launches over the resident user_ptr: kept chunks per user -> exclusive scan -> one thread per kept chunk
// (the owning user found by binary search in the scanned offsets, so a user with a million interactions costs no more
// per thread than one with three).  Bit-identical to host_schedule() in api.cu (tests/test_gpu_data_prep.py).
// ============================================================================================================
namespace sbr {
namespace {

__device__ __forceinline__ void user_chunks(uint64_t len, uint64_t T, uint32_t* kept, uint32_t* first_kept, uint64_t* first) {
    uint64_t f = 0; uint32_t k = 0, fk = 0;
    if (len > 0) {
        f = (len >> 32) == 0 && (T >> 32) == 0 ? (uint64_t)((uint32_t)len % (uint32_t)T) : len % T;
        const uint64_t full = T > 2 ? (len - f) / T : 0;
        fk = f > 2 ? 1u : 0u;
        k = (uint32_t)full + fk;
    }
    *kept = k; *first_kept = fk; *first = f;
}

__global__ void sched_count_kernel(const uint64_t* __restrict__ user_ptr, size_t num_users, uint64_t T, uint32_t* __restrict__ counts) {
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < num_users; u += (size_t)gridDim.x * blockDim.x) {
        uint32_t k, fk; uint64_t f;
        user_chunks(user_ptr[u + 1] - user_ptr[u], T, &k, &fk, &f);
        counts[u] = k;
    }
}

// offsets = exclusive scan of counts, offsets[num_users] = nsub
__global__ void sched_fill_kernel(const uint64_t* __restrict__ user_ptr, const uint32_t* __restrict__ offsets, size_t num_users, uint64_t T,
                                  size_t nsub, uint64_t* __restrict__ seq_start, uint32_t* __restrict__ seq_len) {
    for (size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x; s < nsub; s += (size_t)gridDim.x * blockDim.x) {
        size_t lo = 0, hi = num_users;            // last user u with offsets[u] <= s  (users without kept chunks share offsets: take the last)
        while (hi - lo > 1) { const size_t mid = (lo + hi) >> 1; if (offsets[mid] <= s) lo = mid; else hi = mid; }
        const uint64_t b = user_ptr[lo], len = user_ptr[lo + 1] - b;
        uint32_t k, fk; uint64_t f;
        user_chunks(len, T, &k, &fk, &f);
        const uint32_t j = (uint32_t)(s - offsets[lo]);   // j-th kept chunk of this user
        if (fk && j == 0) { seq_start[s] = b; seq_len[s] = (uint32_t)f; }
        else { seq_start[s] = b + f + (uint64_t)(j - fk) * T; seq_len[s] = (uint32_t)T; }
    }
}

// usize ids -> u32 on the device (ids that arrived as raw 64-bit words from pinned host memory)
__global__ void narrow_ids_kernel(const uint64_t* __restrict__ src, size_t n, uint64_t bound, uint32_t* __restrict__ dst, int* __restrict__ bad) {
    int b = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t v = src[i];
        b |= (int)(v >= bound);
        dst[i] = (uint32_t)v;
    }
    if (b) atomicOr(bad, 1);
}

}  // namespace

// d_counts: [num_users + 1] u32 scratch (becomes the offsets), d_tmp: CUB scratch of *tmp_bytes (query with d_tmp == nullptr)
cudaError_t device_schedule(const uint64_t* d_user_ptr, size_t num_users, size_t T, uint32_t* d_counts, void* d_tmp, size_t* tmp_bytes, size_t nsub,
                            uint64_t* d_seq_start, uint32_t* d_seq_len, cudaStream_t st) {
    if (!d_tmp) return cub::DeviceScan::ExclusiveSum(nullptr, *tmp_bytes, d_counts, d_counts, (int)(num_users + 1), st);
    const int threads = 256;
    const int bu = (int)std::min<size_t>((num_users + threads - 1) / threads, 148 * 16);
    cudaError_t e = cudaMemsetAsync(d_counts + num_users, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (num_users) sched_count_kernel<<<std::max(bu, 1), threads, 0, st>>>(d_user_ptr, num_users, (uint64_t)T, d_counts);
    e = cub::DeviceScan::ExclusiveSum(d_tmp, *tmp_bytes, d_counts, d_counts, (int)(num_users + 1), st);
    if (e != cudaSuccess) return e;
    if (nsub) {
        const int bs = (int)std::min<size_t>((nsub + threads - 1) / threads, 148 * 16);
        sched_fill_kernel<<<bs, threads, 0, st>>>(d_user_ptr, d_counts, num_users, (uint64_t)T, nsub, d_seq_start, d_seq_len);
    }
    return cudaGetLastError();
}

cudaError_t launch_narrow_ids(const uint64_t* d_src, size_t n, uint64_t bound, uint32_t* d_dst, int* d_bad, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    narrow_ids_kernel<<<blocks, 256, 0, st>>>(d_src, n, bound, d_dst, d_bad);
    return cudaGetLastError();
}

}  // namespace sbr
