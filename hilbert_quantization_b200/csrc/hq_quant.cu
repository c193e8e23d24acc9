// a10: uint8 min/max quantise and its inverse (core/compressor.py:256-303), plus the
// general rectangular block-mean kernel used for index rows of non power-of-two images.
//
// Arithmetic is written with explicit round-to-nearest intrinsics so that no FMA
// contraction can change a result: the reference evaluates (v - min) / (max - min) * 255
// as three separate float32 operations and truncates, and the uint8 output must be
// bit-identical.
#include "hq_common.cuh"
#include <float.h>

namespace {

__device__ __forceinline__ uint32_t f2key(float f) {            // monotonic float -> uint
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ void block_minmax(float& mn, float& mx) {
    __shared__ float s_mn[32], s_mx[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) { s_mn[w] = mn; s_mx[w] = mx; }
    __syncthreads();
    mn = l < nw ? s_mn[l] : FLT_MAX;
    mx = l < nw ? s_mx[l] : -FLT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __syncthreads();
}

__device__ __forceinline__ uint8_t quant1(float v, float mn, float range, bool constant) {
    if (constant) return 128;
    const float t = __fmul_rn(__fdiv_rn(__fsub_rn(v, mn), range), 255.0f);
    return (uint8_t)(int)t;                                      // truncation, like ndarray.astype(uint8)
}

// One CTA per item; the item is read twice (second read hits L1/L2: items are <= 64 KB).
__global__ void __launch_bounds__(256) k_quant_small(const float* __restrict__ src, int64_t N, int64_t elems, int64_t src_stride,
                                                     uint8_t* __restrict__ dst, int64_t dst_stride, float* __restrict__ minmax) {
    for (int64_t item = blockIdx.x; item < N; item += gridDim.x) {
        const float* s = src + item * src_stride;
        float mn = FLT_MAX, mx = -FLT_MAX;
        for (int64_t i = threadIdx.x; i < elems; i += blockDim.x) {
            const float v = __ldg(s + i);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
        block_minmax(mn, mx);
        if (threadIdx.x == 0) { minmax[2 * item] = mn; minmax[2 * item + 1] = mx; }
        const bool constant = mn == mx;
        const float range = __fsub_rn(mx, mn);
        uint8_t* o = dst + item * dst_stride;
        for (int64_t i = threadIdx.x; i < elems; i += blockDim.x) o[i] = quant1(__ldg(s + i), mn, range, constant);
    }
}

__global__ void k_minmax_init(uint32_t* keys, int64_t N) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        keys[2 * i] = 0xffffffffu;       // running min (encoded)
        keys[2 * i + 1] = 0u;            // running max (encoded)
    }
}

__global__ void __launch_bounds__(256) k_minmax_big(const float* __restrict__ src, int64_t elems, int64_t src_stride, uint32_t* keys) {
    const int64_t item = blockIdx.y;
    const float* s = src + item * src_stride;
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(s + i);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    block_minmax(mn, mx);
    if (threadIdx.x == 0) {
        atomicMin(keys + 2 * item, f2key(mn));
        atomicMax(keys + 2 * item + 1, f2key(mx));
    }
}

__global__ void __launch_bounds__(256) k_quant_big(const float* __restrict__ src, int64_t elems, int64_t src_stride,
                                                   uint8_t* __restrict__ dst, int64_t dst_stride, const uint32_t* __restrict__ keys) {
    const int64_t item = blockIdx.y;
    const float mn = key2f(keys[2 * item]), mx = key2f(keys[2 * item + 1]);
    const bool constant = mn == mx;
    const float range = __fsub_rn(mx, mn);
    const float* s = src + item * src_stride;
    uint8_t* o = dst + item * dst_stride;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = quant1(__ldg(s + i), mn, range, constant);
}

__global__ void k_minmax_decode(uint32_t* keys, int64_t N) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 2 * N; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = __float_as_uint(key2f(keys[i]));
}

__global__ void __launch_bounds__(256) k_dequant(const uint8_t* __restrict__ src, int64_t N, int64_t elems, int64_t src_stride,
                                                 const float* __restrict__ minmax, float* __restrict__ dst, int64_t dst_stride) {
    const int64_t total = N * elems;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t item = i / elems, e = i - item * elems;
        const float mn = __ldg(minmax + 2 * item), mx = __ldg(minmax + 2 * item + 1);
        float r;
        if (mx == mn) r = mn;
        else r = __fadd_rn(__fmul_rn(__fdiv_rn((float)src[item * src_stride + e], 255.0f), __fsub_rn(mx, mn)), mn);
        dst[item * dst_stride + e] = r;
    }
}

// one warp per (item, section): mean of an sh x sw rectangle
__global__ void __launch_bounds__(256) k_block_means(const float* __restrict__ img, int64_t N, int W, int64_t img_stride, int sh, int sw,
                                                     const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int count,
                                                     float* __restrict__ out, int64_t out_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int cells = sh * sw;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < N * count; w += warps) {
        const int64_t item = w / count;
        const int s = (int)(w - item * count);
        const float* base = img + item * img_stride + (int64_t)rows[s] * sh * W + (int64_t)cols[s] * sw;
        float acc = 0.f;
        for (int c = lane; c < cells; c += 32) {
            const int r = c / sw, q = c - r * sw;
            acc += __ldg(base + (int64_t)r * W + q);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[item * out_stride + s] = acc / (float)cells;
    }
}

}  // namespace

extern "C" int hq_quantize_u8(const float* src, int64_t N, int64_t elems, int64_t src_stride, uint8_t* dst, int64_t dst_stride,
                              float* minmax, void* stream) {
    HQ_REQUIRE(N >= 0 && elems > 0, "bad shape");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(src && dst && minmax, "null pointer");
    HQ_REQUIRE(src_stride >= elems && dst_stride >= elems, "stride smaller than item");
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = hq_cached_sm_count();
    if (elems <= 16384) {
        int64_t blocks = N < (int64_t)sms * 8 ? N : (int64_t)sms * 8;
        k_quant_small<<<(unsigned)blocks, 256, 0, st>>>(src, N, elems, src_stride, dst, dst_stride, minmax);
        HQ_LAUNCH_OK("k_quant_small");
        return HQ_OK;
    }
    HQ_REQUIRE(N <= 65535, "too many large items in one call (%lld)", (long long)N);
    uint32_t* keys = reinterpret_cast<uint32_t*>(minmax);
    k_minmax_init<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(keys, N);
    int64_t bx = (elems + 256 * 16 - 1) / (256 * 16);
    const int64_t cap = ((int64_t)sms * 8 + N - 1) / N;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 g((unsigned)bx, (unsigned)N);
    k_minmax_big<<<g, 256, 0, st>>>(src, elems, src_stride, keys);
    k_quant_big<<<g, 256, 0, st>>>(src, elems, src_stride, dst, dst_stride, keys);
    k_minmax_decode<<<(unsigned)((2 * N + 255) / 256), 256, 0, st>>>(keys, N);
    hq_note_launch(3);
    HQ_LAUNCH_OK("k_quant_big");
    return HQ_OK;
}

extern "C" int hq_dequantize_u8(const uint8_t* src, int64_t N, int64_t elems, int64_t src_stride, const float* minmax, float* dst,
                                int64_t dst_stride, void* stream) {
    HQ_REQUIRE(N >= 0 && elems > 0, "bad shape");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(src && dst && minmax, "null pointer");
    HQ_REQUIRE(src_stride >= elems && dst_stride >= elems, "stride smaller than item");
    int64_t blocks = (N * elems + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_dequant<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, N, elems, src_stride, minmax, dst, dst_stride);
    HQ_LAUNCH_OK("k_dequant");
    return HQ_OK;
}

extern "C" int hq_block_means(const float* img, int64_t N, int H, int W, int64_t img_stride, int sh, int sw, const int32_t* rows,
                              const int32_t* cols, int count, float* out, int64_t out_stride, void* stream) {
    HQ_REQUIRE(N >= 0 && H > 0 && W > 0 && sh > 0 && sw > 0 && count >= 0, "bad shape");
    if (N == 0 || count == 0) return HQ_OK;
    HQ_REQUIRE(img && rows && cols && out, "null pointer");
    HQ_REQUIRE(img_stride >= (int64_t)H * W && out_stride >= count, "stride smaller than item");
    int64_t blocks = (N * count + 7) / 8;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_block_means<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(img, N, W, img_stride, sh, sw, rows, cols, count, out, out_stride);
    HQ_LAUNCH_OK("k_block_means");
    return HQ_OK;
}
