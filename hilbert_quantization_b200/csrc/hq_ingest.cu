// Shard ingest: ONE pass over a block of embeddings produces everything the search path keeps per row -- the
// hierarchical index row (run means over the Hilbert order, gathered by the variant's plan), the row norm and the
// unit-row bf16 operand of the rerank GEMM.  It replaces three passes over the fp32 rows (index-only item pass at
// 2 TB/s because it is issue bound, hq_row_norms, hq_to_bf16_unit) by one that is bound by HBM: read 4 D, write
// 2 D + 4 Lsum + 4 bytes per row.
//
// A run of 4^k consecutive curve positions is an aligned 2^k x 2^k block of the grid (what the item/tile kernels
// exploit), so without a grid to write the block means of every level are run means of the INPUT row: no
// coordinates at all.  One warp owns a row; lane l holds the 16-byte chunks j * 32 + l (j < J = D / 128):
//   level 1 (4 values)    in the thread:            ((x + y) + (z + w)) * 0.25
//   level 2 (16)          butterfly over lanes ^1, ^2
//   level 3 (64)          butterfly over lanes ^4, ^8         -> two values per j (lanes 0-15, 16-31)
//   level 4 (256)         lanes ^16 of chunk rows j, j + 1:   ((a + b) + (c + d)) * 0.25
//   level 5, 6            from the level below, in registers (uniform over the warp)
// which is exactly the 4-ary ((a + b) + (c + d)) * 0.25 tree of hq_item_pass.cuh / hq_tile.cu, so index rows are bit
// identical to hq_map_index's; the norm accumulates chunk by chunk with the fmaf order of k_row_norms and the bf16
// rows divide by it like k_to_bf16_unit, so those are bit identical too (tests/test_gpu_tensorcore.py).
//
// Reference: per document rag/embedding_generation/hierarchical_index_generator.py:23-342 (index rows) and the cosine
// normalisation of rag/search/engine.py:622-660, here for a whole shard at once.
#include "hq_common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int kWarps = 8;
constexpr int kStage = 64 + 16 + 4 + 1 + 3;          // levels 3..6 of a 64 x 64 grid, padded

template <int J>
__global__ void __launch_bounds__(kWarps * 32, (J <= 12 ? 3 : 1)) k_shard_ingest(const float* __restrict__ emb, int64_t N, int64_t stride,
                                                             const int32_t* __restrict__ plan, int Lsum, float* __restrict__ idx,
                                                             int64_t idx_pitch, float* __restrict__ norms,
                                                             __nv_bfloat16* __restrict__ unit, int64_t unit_pitch) {
    __shared__ float s_stage[kWarps][kStage];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* st = s_stage[warp];
    float* st3 = st;            // up to 64 values
    float* st4 = st + 64;       // up to 16
    float* st5 = st + 80;       // up to 4
    float* st6 = st + 84;       // 1
    constexpr int N3 = 2 * J, N4 = (J + 1) / 2, N5 = (J + 7) / 8, N6 = (J + 31) / 32;
    const int64_t warps = (int64_t)gridDim.x * kWarps;
    for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < N; row += warps) {
        const float4* src = reinterpret_cast<const float4*>(emb + row * stride) + lane;
        float4 v[J];
#pragma unroll
        for (int j = 0; j < J; ++j) v[j] = __ldcs(src + j * 32);
        float acc = 0.f;
        float m3[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            acc = fmaf(v[j].x, v[j].x, acc); acc = fmaf(v[j].y, v[j].y, acc);
            acc = fmaf(v[j].z, v[j].z, acc); acc = fmaf(v[j].w, v[j].w, acc);
            float t = ((v[j].x + v[j].y) + (v[j].z + v[j].w)) * 0.25f;
            t = t + __shfl_xor_sync(0xffffffffu, t, 1);
            t = t + __shfl_xor_sync(0xffffffffu, t, 2);
            t *= 0.25f;
            t = t + __shfl_xor_sync(0xffffffffu, t, 4);
            t = t + __shfl_xor_sync(0xffffffffu, t, 8);
            m3[j] = t * 0.25f;
        }
        float m4[N4];
#pragma unroll
        for (int i = 0; i < N4; ++i) {
            const float a = m3[2 * i] + __shfl_xor_sync(0xffffffffu, m3[2 * i], 16);
            float b = 0.f;
            if (2 * i + 1 < J) b = m3[2 * i + 1] + __shfl_xor_sync(0xffffffffu, m3[2 * i + 1], 16);
            m4[i] = (a + b) * 0.25f;
        }
        float m5[N5];
#pragma unroll
        for (int i = 0; i < N5; ++i) {
            const float a = m4[4 * i], b = 4 * i + 1 < N4 ? m4[4 * i + 1] : 0.f;
            const float c = 4 * i + 2 < N4 ? m4[4 * i + 2] : 0.f, d = 4 * i + 3 < N4 ? m4[4 * i + 3] : 0.f;
            m5[i] = ((a + b) + (c + d)) * 0.25f;
        }
        float m6;
        {
            const float a = m5[0], b = N5 > 1 ? m5[N5 > 1 ? 1 : 0] : 0.f;
            const float c = N5 > 2 ? m5[N5 > 2 ? 2 : 0] : 0.f, d = N5 > 3 ? m5[N5 > 3 ? 3 : 0] : 0.f;
            m6 = ((a + b) + (c + d)) * 0.25f;
        }
        __syncwarp();                                   // the previous row's gather has read the staging area
        if ((lane & 15) == 0) {
#pragma unroll
            for (int j = 0; j < J; ++j) st3[2 * j + (lane >> 4)] = m3[j];
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < N4; ++i) st4[i] = m4[i];
#pragma unroll
            for (int i = 0; i < N5; ++i) st5[i] = m5[i];
            st6[0] = m6;
        }
        // the norm, with the association of k_row_norms: per-lane fmaf chain over its chunks, then the xor butterfly
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const float nrm = sqrtf(acc);
        if (lane == 0 && norms) norms[row] = nrm;
        __syncwarp();
        for (int s = lane; s < Lsum; s += 32) {
            const int code = __ldg(plan + s), k = code >> 24, pos = code & 0xffffff;
            float val = 0.f;                            // positions past the row's data are structural zeros
            if (k == 3) { if (pos < N3) val = st3[pos]; }
            else if (k == 4) { if (pos < N4) val = st4[pos]; }
            else if (k == 5) { if (pos < N5) val = st5[pos]; }
            else if (k == 6) { if (pos < N6) val = st6[pos]; }
            idx[row * idx_pitch + s] = val;
        }
        if (unit) {
            uint2* dst = reinterpret_cast<uint2*>(unit + row * unit_pitch) + lane;
            const bool ok = nrm > 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float a = ok ? __fdiv_rn(v[j].x, nrm) : 0.f, b = ok ? __fdiv_rn(v[j].y, nrm) : 0.f;
                const float c = ok ? __fdiv_rn(v[j].z, nrm) : 0.f, d = ok ? __fdiv_rn(v[j].w, nrm) : 0.f;
                uint2 w;
                w.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
                w.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(c)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(d)) << 16);
                dst[j * 32] = w;
            }
        }
    }
}

bool ingest_chunks_ok(int64_t D) {
    if (D <= 0 || D % 128 != 0) return false;
    switch (D / 128) {
        case 2: case 4: case 6: case 8: case 12: case 16: case 24: case 32: return true;
        default: return false;
    }
}

}  // namespace

extern "C" int hq_shard_ingest_supported(int64_t D) { return ingest_chunks_ok(D) ? 1 : 0; }

extern "C" int hq_shard_ingest(const float* emb, int64_t N, int64_t D, int64_t stride, const int32_t* plan_codes, int Lsum,
                               float* idx, int64_t idx_pitch, float* norms, void* unit_bf16, int64_t unit_pitch, void* stream) {
    HQ_REQUIRE(N >= 0 && stride >= D, "bad shape");
    HQ_REQUIRE(ingest_chunks_ok(D), "hq_shard_ingest needs D / 128 in {2, 4, 6, 8, 12, 16, 24, 32} (D = %lld)", (long long)D);
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(emb && plan_codes && idx, "null pointer");
    HQ_REQUIRE(Lsum > 0 && idx_pitch >= Lsum, "bad index pitch");
    HQ_REQUIRE(stride % 4 == 0 && ((uintptr_t)emb & 15) == 0, "embedding rows must be 16-byte aligned");
    HQ_REQUIRE(!unit_bf16 || (unit_pitch >= D && unit_pitch % 4 == 0 && ((uintptr_t)unit_bf16 & 7) == 0), "bf16 rows must be 8-byte aligned");
    int64_t blocks = (N + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define HQ_INGEST(JJ)                                                                                                         \
    case JJ:                                                                                                                  \
        k_shard_ingest<JJ><<<(unsigned)blocks, kWarps * 32, 0, st>>>(emb, N, stride, plan_codes, Lsum, idx, idx_pitch, norms, \
                                                                     (__nv_bfloat16*)unit_bf16, unit_pitch);                  \
        break
    switch ((int)(D / 128)) {
        HQ_INGEST(2); HQ_INGEST(4); HQ_INGEST(6); HQ_INGEST(8); HQ_INGEST(12); HQ_INGEST(16); HQ_INGEST(24); HQ_INGEST(32);
        default: break;
    }
#undef HQ_INGEST
    HQ_LAUNCH_OK("k_shard_ingest");
    return HQ_OK;
}
