// Microbenchmark: what HBM bandwidth can a kernel with the map+index traffic shape reach?
// (6144 B read + 16384 B grid + 336 B index written per item, no transposition at all.)
// Variants: copy (1:1), pure write, mix with LSU stores, mix with TMA bulk stores from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_mix hbm_mix.cu && ./hbm_mix
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int64_t j = i + (int64_t)u * gridDim.x * blockDim.x; if (j < n) v[u] = __ldcs(src + j); }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int64_t j = i + (int64_t)u * gridDim.x * blockDim.x; if (j < n) __stcs(dst + j, v[u]); }
    }
}

__global__ void __launch_bounds__(256) k_write(float4* __restrict__ dst, int64_t n, float x) {
    const float4 v = make_float4(x, x, x, x);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) __stcs(dst + i, v);
}

// one item per CTA iteration: 384 float4 in, 1024 + 21 float4 out, LSU stores, next item's loads in flight
template <int PREFETCH>
__global__ void __launch_bounds__(256) k_mix_lsu(const float4* __restrict__ src, float4* __restrict__ grid, float4* __restrict__ idx, int64_t N) {
    const int tid = threadIdx.x;
    float4 a = make_float4(0, 0, 0, 0), b = a;
    int64_t it = blockIdx.x;
    if (it < N) { a = __ldcs(src + it * 384 + tid); if (tid < 128) b = __ldcs(src + it * 384 + 256 + tid); }
    for (; it < N; it += gridDim.x) {
        float4 na = a, nb = b;
        const int64_t nx = it + gridDim.x;
        if (PREFETCH && nx < N) { na = __ldcs(src + nx * 384 + tid); if (tid < 128) nb = __ldcs(src + nx * 384 + 256 + tid); }
        float4* g = grid + it * 1024;
        __stcs(g + tid, a);
        __stcs(g + 256 + tid, b);
        __stcs(g + 512 + tid, make_float4(0, 0, 0, 0));
        __stcs(g + 768 + tid, make_float4(0, 0, 0, 0));
        if (tid < 21) __stcs(idx + it * 21 + tid, a);
        if (!PREFETCH && nx < N) { na = __ldcs(src + nx * 384 + tid); if (tid < 128) nb = __ldcs(src + nx * 384 + 256 + tid); }
        a = na; b = nb;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// same traffic, grid written by TMA bulk copies (16 KB per item) out of a ring of shared-memory images
template <int RING>
__global__ void __launch_bounds__(256) k_mix_tma(const float4* __restrict__ src, float4* __restrict__ grid, float4* __restrict__ idx, int64_t N) {
    extern __shared__ __align__(128) float4 ring[];      // RING x 1024 float4
    const int tid = threadIdx.x;
    for (int i = tid; i < RING * 1024; i += 256) ring[i] = make_float4(0, 0, 0, 0);
    __syncthreads();
    float4 a = make_float4(0, 0, 0, 0), b = a;
    int64_t it = blockIdx.x;
    if (it < N) { a = __ldcs(src + it * 384 + tid); if (tid < 128) b = __ldcs(src + it * 384 + 256 + tid); }
    int slot = 0;
    for (; it < N; it += gridDim.x) {
        // slot must have been read by its previous bulk store: allow RING-1 groups in flight
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(RING - 1) : "memory");
        __syncthreads();
        float4* img = ring + slot * 1024;
        img[tid] = a;
        if (tid < 128) img[256 + tid] = b;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grid + it * 1024), "r"(smem_u32(img)), "n"(16384) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (tid < 21) __stcs(idx + it * 21 + tid, a);
        const int64_t nx = it + gridDim.x;
        if (nx < N) { a = __ldcs(src + nx * 384 + tid); if (tid < 128) b = __ldcs(src + nx * 384 + 256 + tid); }
        slot = slot + 1 == RING ? 0 : slot + 1;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_ms(F f, int reps = 10) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, tot = 0;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best; tot += ms;
    }
    CK(cudaGetLastError());
    printf("    best %.4f ms  mean %.4f ms\n", best, tot / reps);
    return tot / reps;
}

int main() {
    const int64_t N = 1000000;
    float4 *src, *grid, *idx;
    CK(cudaMalloc(&src, N * 6144)); CK(cudaMalloc(&grid, N * 16384)); CK(cudaMalloc(&idx, N * 336));
    CK(cudaMemset(src, 1, N * 6144));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const double mix_bytes = (double)N * (6144 + 16384 + 336);
    {   // copy 8 GB + 8 GB
        const int64_t n = N * 1024 / 2;
        printf("copy %.1f GB read + %.1f GB write\n", n * 16 / 1e9, n * 16 / 1e9);
        for (int mult : {8, 16, 32}) {
            printf("  grid %d x SMs\n", mult);
            float ms = time_ms([&] { k_copy<<<sms * mult, 256>>>(grid, grid + n, n); });
            printf("    -> %.0f GB/s\n", 2.0 * n * 16 / ms / 1e6);
        }
    }
    {
        const int64_t n = N * 1024;
        printf("pure write %.1f GB\n", n * 16 / 1e9);
        for (int mult : {8, 32}) {
            printf("  grid %d x SMs\n", mult);
            float ms = time_ms([&] { k_write<<<sms * mult, 256>>>(grid, n, 1.0f); });
            printf("    -> %.0f GB/s\n", 1.0 * n * 16 / ms / 1e6);
        }
    }
    printf("mix (map+index traffic shape, %.2f GB), LSU stores\n", mix_bytes / 1e9);
    for (int mult : {4, 8, 16}) {
        printf("  grid %d x SMs, prefetch\n", mult);
        float ms = time_ms([&] { k_mix_lsu<1><<<sms * mult, 256>>>(src, grid, idx, N); });
        printf("    -> %.0f GB/s\n", mix_bytes / ms / 1e6);
    }
    printf("  grid 8 x SMs, no prefetch\n");
    { float ms = time_ms([&] { k_mix_lsu<0><<<sms * 8, 256>>>(src, grid, idx, N); }); printf("    -> %.0f GB/s\n", mix_bytes / ms / 1e6); }
    printf("mix, TMA bulk stores\n");
    CK(cudaFuncSetAttribute(k_mix_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384));
    CK(cudaFuncSetAttribute(k_mix_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16384));
    CK(cudaFuncSetAttribute(k_mix_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384));
    for (int mult : {2, 3, 4}) {
        printf("  ring 2, grid %d x SMs\n", mult);
        float ms = time_ms([&] { k_mix_tma<2><<<sms * mult, 256, 2 * 16384>>>(src, grid, idx, N); });
        printf("    -> %.0f GB/s\n", mix_bytes / ms / 1e6);
    }
    for (int mult : {2, 3}) {
        printf("  ring 3, grid %d x SMs\n", mult);
        float ms = time_ms([&] { k_mix_tma<3><<<sms * mult, 256, 3 * 16384>>>(src, grid, idx, N); });
        printf("    -> %.0f GB/s\n", mix_bytes / ms / 1e6);
    }
    printf("  ring 4, grid 2 x SMs\n");
    { float ms = time_ms([&] { k_mix_tma<4><<<sms * 2, 256, 4 * 16384>>>(src, grid, idx, N); }); printf("    -> %.0f GB/s\n", mix_bytes / ms / 1e6); }
    return 0;
}
