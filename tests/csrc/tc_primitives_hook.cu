// tc_primitives_hook.cu -- unit-test entry for the tcgen05 tile primitives of sbr-rs_b200/csrc/tc_tile.cuh.  Test
// infrastructure: built into its own tests/csrc/libtc_primitives_hook.so, never linked into libsbr_b200.so.
// Runs the three GEMM shapes of the tensor-core LSTM kernel on caller-provided row-major matrices:
//   mode 1: D[128x128] = Z[128x80](:, :64) . W[64x128]     tf32  (gates:  A K-major,  B K-major)
//   mode 2: D[128x64]  = Dl[128x128] . W[64x128]^T         bf16  (dz:     A K-major,  B K-major)
//   mode 3: D[128x80]  = Dl[128x128]^T . Z[128x80]         bf16  (dW^T:   A MN-major, B MN-major; same delta tile)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "tc_tile.cuh"

using namespace sbr::tc;

namespace {

__global__ void __launch_bounds__(128) tc_gemm_kernel(int mode, const float* __restrict__ Zg, const float* __restrict__ Dg,
                                                      const float* __restrict__ Wg, float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* zt = smem;                       // tf32 [128 seq][64 feat]        32768
    uint8_t* wt = zt + 128 * 64 * 4;          // tf32 [128 gd][64 feat]         32768
    uint8_t* db = wt + 128 * 64 * 4;          // bf16 [128 seq][128 gd]         32768
    uint8_t* zb = db + 128 * 128 * 2;         // bf16 [128 seq][80 feat]        20480
    uint8_t* wb = zb + 128 * 80 * 2;          // bf16 [64 feat][128 gd]         16384
    uint64_t* bar = reinterpret_cast<uint64_t*>(wb + 64 * 128 * 2);
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(bar + 1);
    const int r = threadIdx.x, warp = r >> 5;

    if (warp == 0) tmem_alloc<256>(tmem_base);
    if (r == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int c4 = 0; c4 < 16; ++c4) {  // Z row r (tf32, first 64 features)
        float4 v = *reinterpret_cast<const float4*>(Zg + (size_t)r * 80 + c4 * 4);
        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
        *reinterpret_cast<float4*>(zt + tile_chunk_off(r, c4, 16)) = v;
    }
    for (int c4 = 0; c4 < 16; ++c4) {  // W tile row gd = r: element (gd, feature) = W[feature][gd]
        float4 v;
        v.x = to_tf32(Wg[(size_t)(c4 * 4 + 0) * 128 + r]); v.y = to_tf32(Wg[(size_t)(c4 * 4 + 1) * 128 + r]);
        v.z = to_tf32(Wg[(size_t)(c4 * 4 + 2) * 128 + r]); v.w = to_tf32(Wg[(size_t)(c4 * 4 + 3) * 128 + r]);
        *reinterpret_cast<float4*>(wt + tile_chunk_off(r, c4, 16)) = v;
    }
    for (int c8 = 0; c8 < 16; ++c8) {  // delta row r (bf16)
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = Dg[(size_t)r * 128 + c8 * 8 + i];
        *reinterpret_cast<uint4*>(db + tile_chunk_off(r, c8, 16)) = pack_bf16x8(v);
    }
    for (int c8 = 0; c8 < 10; ++c8) {  // Z row r (bf16, 80 features)
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = Zg[(size_t)r * 80 + c8 * 8 + i];
        *reinterpret_cast<uint4*>(zb + tile_chunk_off(r, c8, 10)) = pack_bf16x8(v);
    }
    if (r < 64) {                       // Wb row feature = r: element (feature, gd) = W[feature][gd]
        for (int c8 = 0; c8 < 16; ++c8) {
            float v[8];
            for (int i = 0; i < 8; ++i) v[i] = Wg[(size_t)r * 128 + c8 * 8 + i];
            *reinterpret_cast<uint4*>(wb + tile_chunk_off(r, c8, 16)) = pack_bf16x8(v);
        }
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *tmem_base;
    if (r == 0) {
        const uint32_t za = smem_u32(zt), wa = smem_u32(wt), da = smem_u32(db), zba = smem_u32(zb), wba = smem_u32(wb);
        if (mode == 1) {
            const uint32_t idesc = make_idesc_tf32(128, 128, 0, 0);
            for (int k = 0; k < 8; ++k)   // K = 8 tf32 = 2 cores along K per instruction
                mma_tf32(tm, make_smem_desc(za + k * 256, 128, 16 * 128), make_smem_desc(wa + k * 256, 128, 16 * 128), idesc, k > 0);
        } else if (mode == 2) {
            const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
            for (int k = 0; k < 8; ++k)   // K = 16 bf16 = 2 cores along K per instruction
                mma_bf16(tm, make_smem_desc(da + k * 256, 128, 16 * 128), make_smem_desc(wba + k * 256, 128, 16 * 128), idesc, k > 0);
        } else {
            const uint32_t idesc = make_idesc_bf16(128, 80, 1, 1);
            for (int k = 0; k < 8; ++k)   // K = 16 sequences = 2 core-rows per instruction
                mma_bf16(tm, make_smem_desc(da + k * (2 * 16 * 128), 16 * 128, 128),
                         make_smem_desc(zba + k * (2 * 10 * 128), 10 * 128, 128), idesc, k > 0);
        }
        mma_commit(bar);
    }
    const int ncols = mode == 1 ? 128 : mode == 2 ? 64 : 80;
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < ncols; c += 8) {
        float v[8];
        tmem_ld8(lane_addr + c, v);
        for (int i = 0; i < 8; ++i) out[(size_t)r * ncols + c + i] = v[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tm);
}

}  // namespace

extern "C" int sbrdbg_tc_gemm(int mode, const float* Z /*128x80*/, const float* Dl /*128x128*/, const float* W /*64x128*/,
                              float* out) {
    const int ncols = mode == 1 ? 128 : mode == 2 ? 64 : 80;
    float *dZ, *dD, *dW, *dO;
    if (cudaMalloc(&dZ, 128 * 80 * 4) || cudaMalloc(&dD, 128 * 128 * 4) || cudaMalloc(&dW, 64 * 128 * 4) || cudaMalloc(&dO, 128 * ncols * 4))
        return (int)cudaGetLastError();
    cudaMemcpy(dZ, Z, 128 * 80 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, Dl, 128 * 128 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W, 64 * 128 * 4, cudaMemcpyHostToDevice);
    const int smem = 128 * 64 * 4 * 2 + 128 * 128 * 2 + 128 * 80 * 2 + 64 * 128 * 2 + 64;
    cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_gemm_kernel<<<1, 128, smem>>>(mode, dZ, dD, dW, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, dO, 128 * ncols * 4, cudaMemcpyDeviceToHost);
    cudaFree(dZ); cudaFree(dD); cudaFree(dW); cudaFree(dO);
    return (int)e;
}
