// K1: batched d2xy / xy2d, and the width-generic (1/2/4/8-byte) permutation kernels.
// The 4-byte hot path lives in hq_tile.cu; these kernels are the dtype-preserving
// general path behind HilbertCurveMapper.map_to_2d / map_from_2d.
#include "hq_common.cuh"

__global__ void __launch_bounds__(256) k_d2xy(int log2n, int64_t d0, int64_t count, int32_t* __restrict__ x,
                                              int32_t* __restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t xx, yy;
        hq_d2xy(log2n, (uint64_t)(d0 + i), xx, yy);
        x[i] = (int32_t)xx;
        y[i] = (int32_t)yy;
    }
}

__global__ void __launch_bounds__(256) k_xy2d(int log2n, const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                                              int64_t count, int64_t* __restrict__ d) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        d[i] = (int64_t)hq_xy2d(log2n, (uint32_t)x[i], (uint32_t)y[i]);
}

// out[b, y, x] = d < D ? src[b, d] : 0 with d = xy2d(x, y).  Writes are coalesced; the
// reads stay inside the 2^j x 2^j locality of the curve, so they hit L1/L2.
template <typename T>
__global__ void __launch_bounds__(256) k_map_to_2d_any(const T* __restrict__ src, int64_t N, int64_t D, int64_t src_stride,
                                                       int log2n, T* __restrict__ dst, int64_t dst_stride) {
    const int64_t cells = (int64_t)1 << (2 * log2n);
    const int64_t total = N * cells;
    const uint32_t nm1 = (1u << log2n) - 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i >> (2 * log2n);
        const uint32_t cell = (uint32_t)(i & (cells - 1));
        const uint64_t d = hq_xy2d(log2n, cell & nm1, cell >> log2n);
        dst[b * dst_stride + cell] = (int64_t)d < D ? src[b * src_stride + (int64_t)d] : T(0);
    }
}

// out[b, d] = src[b, y, x] for d < D_out.
template <typename T>
__global__ void __launch_bounds__(256) k_map_from_2d_any(const T* __restrict__ src, int64_t N, int log2n, int64_t src_stride,
                                                         int64_t D_out, T* __restrict__ dst, int64_t dst_stride) {
    const int64_t total = N * D_out;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / D_out;
        const int64_t d = i - b * D_out;
        uint32_t x, y;
        hq_d2xy(log2n, (uint64_t)d, x, y);
        dst[b * dst_stride + d] = src[b * src_stride + ((int64_t)y << log2n) + x];
    }
}

static int grid_for(int64_t work_items, int threads) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" int hq_d2xy_batch(int n, int64_t d0, int64_t count, int32_t* x, int32_t* y, void* stream) {
    HQ_REQUIRE(hq_is_pow2(n), "Grid size must be a power of 2, got %d", n);
    HQ_REQUIRE(n <= (1 << 15), "grid size %d too large", n);
    HQ_REQUIRE(d0 >= 0 && count >= 0 && d0 + count <= (int64_t)n * n, "d range [%lld, %lld) outside grid", (long long)d0,
               (long long)(d0 + count));
    if (count == 0) return HQ_OK;
    HQ_REQUIRE(x && y, "null output");
    k_d2xy<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(hq_log2(n), d0, count, x, y);
    HQ_LAUNCH_OK("k_d2xy");
    return HQ_OK;
}

extern "C" int hq_xy2d_batch(int n, const int32_t* x, const int32_t* y, int64_t count, int64_t* d, void* stream) {
    HQ_REQUIRE(hq_is_pow2(n), "Grid size must be a power of 2, got %d", n);
    HQ_REQUIRE(n <= (1 << 15), "grid size %d too large", n);
    if (count <= 0) return HQ_OK;
    HQ_REQUIRE(x && y && d, "null pointer");
    k_xy2d<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(hq_log2(n), x, y, count, d);
    HQ_LAUNCH_OK("k_xy2d");
    return HQ_OK;
}

// fp32 hot path (hq_tile.cu)
int hq_tile_map_words(const uint32_t* src, int direction, int64_t N, int64_t D, int64_t src_stride, int n, uint32_t* dst,
                      int64_t dst_stride, cudaStream_t stream);

extern "C" int hq_map_to_2d(const void* src, int64_t N, int64_t D, int64_t src_stride, int n, int elem_bytes, void* dst,
                            int64_t dst_stride, void* stream) {
    HQ_REQUIRE(hq_is_pow2(n), "Dimension must be a power of 2, got %d", n);
    HQ_REQUIRE(n <= (1 << 15), "grid size %d too large", n);
    HQ_REQUIRE(D >= 0 && D <= (int64_t)n * n, "Too many parameters (%lld) for dimensions %dx%d", (long long)D, n, n);
    HQ_REQUIRE(N >= 0, "negative batch");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(dst && (src || D == 0), "null pointer");
    HQ_REQUIRE(src_stride >= D && dst_stride >= (int64_t)n * n, "stride smaller than row");
    const int lg = hq_log2(n);
    const int64_t total = N * (int64_t)n * n;
    cudaStream_t st = (cudaStream_t)stream;
    if (elem_bytes == 4 && n >= 4)
        return hq_tile_map_words((const uint32_t*)src, 0, N, D, src_stride, n, (uint32_t*)dst, dst_stride, st);
    switch (elem_bytes) {
        case 1: k_map_to_2d_any<uint8_t><<<grid_for(total, 256), 256, 0, st>>>((const uint8_t*)src, N, D, src_stride, lg, (uint8_t*)dst, dst_stride); break;
        case 2: k_map_to_2d_any<uint16_t><<<grid_for(total, 256), 256, 0, st>>>((const uint16_t*)src, N, D, src_stride, lg, (uint16_t*)dst, dst_stride); break;
        case 4: k_map_to_2d_any<uint32_t><<<grid_for(total, 256), 256, 0, st>>>((const uint32_t*)src, N, D, src_stride, lg, (uint32_t*)dst, dst_stride); break;
        case 8: k_map_to_2d_any<uint64_t><<<grid_for(total, 256), 256, 0, st>>>((const uint64_t*)src, N, D, src_stride, lg, (uint64_t*)dst, dst_stride); break;
        default: HQ_REQUIRE(false, "unsupported element width %d", elem_bytes);
    }
    HQ_LAUNCH_OK("k_map_to_2d_any");
    return HQ_OK;
}

extern "C" int hq_map_from_2d(const void* src, int64_t N, int n, int64_t src_stride, int64_t D_out, int elem_bytes, void* dst,
                              int64_t dst_stride, void* stream) {
    HQ_REQUIRE(hq_is_pow2(n), "Dimension must be a power of 2, got %d", n);
    HQ_REQUIRE(n <= (1 << 15), "grid size %d too large", n);
    HQ_REQUIRE(D_out >= 0 && D_out <= (int64_t)n * n, "D_out %lld outside grid", (long long)D_out);
    HQ_REQUIRE(N >= 0, "negative batch");
    if (N == 0 || D_out == 0) return HQ_OK;
    HQ_REQUIRE(src && dst, "null pointer");
    HQ_REQUIRE(src_stride >= (int64_t)n * n && dst_stride >= D_out, "stride smaller than row");
    const int lg = hq_log2(n);
    const int64_t total = N * D_out;
    cudaStream_t st = (cudaStream_t)stream;
    if (elem_bytes == 4 && n >= 4)
        return hq_tile_map_words((const uint32_t*)src, 1, N, D_out, src_stride, n, (uint32_t*)dst, dst_stride, st);
    switch (elem_bytes) {
        case 1: k_map_from_2d_any<uint8_t><<<grid_for(total, 256), 256, 0, st>>>((const uint8_t*)src, N, lg, src_stride, D_out, (uint8_t*)dst, dst_stride); break;
        case 2: k_map_from_2d_any<uint16_t><<<grid_for(total, 256), 256, 0, st>>>((const uint16_t*)src, N, lg, src_stride, D_out, (uint16_t*)dst, dst_stride); break;
        case 4: k_map_from_2d_any<uint32_t><<<grid_for(total, 256), 256, 0, st>>>((const uint32_t*)src, N, lg, src_stride, D_out, (uint32_t*)dst, dst_stride); break;
        case 8: k_map_from_2d_any<uint64_t><<<grid_for(total, 256), 256, 0, st>>>((const uint64_t*)src, N, lg, src_stride, D_out, (uint64_t*)dst, dst_stride); break;
        default: HQ_REQUIRE(false, "unsupported element width %d", elem_bytes);
    }
    HQ_LAUNCH_OK("k_map_from_2d_any");
    return HQ_OK;
}
