// a14: the comprehensive similarity blend of rag/search/engine.py:516-575
//     0.5 * hierarchical (weighted per-level (cos+1)/2 of the index rows, :994-1099, weights :1101-1138)
//   + 0.3 * embedding cosine of the grids                                (:622-660)
//   + 0.2 * spatial locality: mean (cos+1)/2 over ws x ws windows, stride ws/2  (:662-714)
// for (query frame, candidate frame) pairs.  Frames are "enhanced": rows [0, n) the n x n Hilbert
// grid, rows [n, n + L) the index rows (zero padded to n values).  One CTA per pair: both
// frames are staged in shared memory with 128-bit loads; warps 0..L-1 reduce one index level
// each, all threads share the grid cosine and the windows.  961 windows of 16 values per
// 64 x 64 pair = 46 K FMAs, so a pair costs about as much as reading it (17 KB): HBM bound
// for a full scan, latency bound for a shortlist.
#include "hq_common.cuh"

namespace {

constexpr int kBlendThreads = 256;

struct BlendParams {
    const float* frames;      // [N, frame_stride]
    int64_t N, frame_stride;
    const float* q_frames;    // [Q, q_stride]
    int64_t q_stride;
    int Q, n, L;
    float w[8];               // granularity weights (sum 1)
    const int64_t* cand_ids;  // optional [Q, M]; NULL: M == N, candidate m is row m
    int64_t M;
    float* out;               // [Q, M]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float cos01(float dot, float qq, float cc) {
    const float nq = sqrtf(qq), nc = sqrtf(cc);
    if (nq == 0.f || nc == 0.f) return 0.f;
    return __fmul_rn(__fadd_rn(__fdiv_rn(dot, __fmul_rn(nq, nc)), 1.0f), 0.5f);
}

__global__ void __launch_bounds__(kBlendThreads) k_blend(const BlendParams p) {
    extern __shared__ __align__(16) float sm[];
    const int n = p.n, L = p.L;
    const int frame_floats = (n + L) * n;
    float* s_q = sm;
    float* s_c = sm + frame_floats;
    __shared__ float s_red[3][kBlendThreads / 32];
    __shared__ float s_lvl[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pair = blockIdx.x;
    const int q = (int)(pair / p.M);
    const int64_t m = pair - (int64_t)q * p.M;
    const int64_t c = p.cand_ids ? p.cand_ids[pair] : m;
    if (c < 0 || c >= p.N) {
        if (tid == 0) p.out[pair] = -1.0f;
        return;
    }
    const float* gq = p.q_frames + (int64_t)q * p.q_stride;
    const float* gc = p.frames + c * p.frame_stride;
    const bool vec = (frame_floats % 4 == 0) && (p.q_stride % 4 == 0) && (p.frame_stride % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.q_frames) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.frames) & 15) == 0);
    if (vec) {
        for (int i = tid; i < frame_floats / 4; i += kBlendThreads) {
            reinterpret_cast<float4*>(s_q)[i] = __ldg(reinterpret_cast<const float4*>(gq) + i);
            reinterpret_cast<float4*>(s_c)[i] = __ldcs(reinterpret_cast<const float4*>(gc) + i);
        }
    } else {
        for (int i = tid; i < frame_floats; i += kBlendThreads) { s_q[i] = __ldg(gq + i); s_c[i] = __ldg(gc + i); }
    }
    __syncthreads();

    // ---- hierarchical part: one warp per index level ----
    for (int l = warp; l < L; l += kBlendThreads / 32) {
        const float* a = s_q + (n + l) * n;
        const float* b = s_c + (n + l) * n;
        float dot = 0.f, qq = 0.f, cc = 0.f;
        for (int i = lane; i < n; i += 32) { dot = fmaf(a[i], b[i], dot); qq = fmaf(a[i], a[i], qq); cc = fmaf(b[i], b[i], cc); }
        dot = warp_sum(dot); qq = warp_sum(qq); cc = warp_sum(cc);
        if (lane == 0) s_lvl[l] = p.w[l] * cos01(dot, qq, cc);
    }
    // ---- embedding cosine over the whole grid ----
    float dot = 0.f, qq = 0.f, cc = 0.f;
    for (int i = tid; i < n * n; i += kBlendThreads) { dot = fmaf(s_q[i], s_c[i], dot); qq = fmaf(s_q[i], s_q[i], qq); cc = fmaf(s_c[i], s_c[i], cc); }
    dot = warp_sum(dot); qq = warp_sum(qq); cc = warp_sum(cc);
    if (lane == 0) { s_red[0][warp] = dot; s_red[1][warp] = qq; s_red[2][warp] = cc; }
    __syncthreads();
    float emb = 0.f;
    {
        float d = 0.f, a = 0.f, b = 0.f;
        for (int i = 0; i < kBlendThreads / 32; ++i) { d += s_red[0][i]; a += s_red[1][i]; b += s_red[2][i]; }
        emb = cos01(d, a, b);
    }
    __syncthreads();
    // ---- spatial locality: ws x ws windows at stride ws / 2 ----
    int ws = n / 4 < 4 ? n / 4 : 4;
    float spatial = emb;
    if (ws >= 2) {
        const int step = ws / 2;
        const int per = (n - ws) / step + 1;
        float acc = 0.f;
        for (int wdx = tid; wdx < per * per; wdx += kBlendThreads) {
            const int wi = wdx / per, wj = wdx - wi * per;
            const float* a = s_q + (wi * step) * n + wj * step;
            const float* b = s_c + (wi * step) * n + wj * step;
            float d = 0.f, x = 0.f, y = 0.f;
            for (int r = 0; r < ws; ++r)
                for (int cidx = 0; cidx < ws; ++cidx) {
                    const float u = a[r * n + cidx], v = b[r * n + cidx];
                    d = fmaf(u, v, d); x = fmaf(u, u, x); y = fmaf(v, v, y);
                }
            acc += cos01(d, x, y);
        }
        acc = warp_sum(acc);
        if (lane == 0) s_red[0][warp] = acc;
        __syncthreads();
        float tot = 0.f;
        for (int i = 0; i < kBlendThreads / 32; ++i) tot += s_red[0][i];
        spatial = tot / (float)(per * per);
    }
    if (tid == 0) {
        float hier = 0.f;                                   // fixed order: results are run-to-run identical
        for (int l = 0; l < L; ++l) hier += s_lvl[l];
        p.out[pair] = 0.5f * hier + 0.3f * emb + 0.2f * spatial;
    }
}

}  // namespace

extern "C" int hq_comprehensive_scores(const float* frames, int64_t N, int n, int L, int64_t frame_stride, const float* q_frames, int Q,
                                       int64_t q_stride, const float* weights_host, const int64_t* cand_ids, int64_t M, float* out,
                                       void* stream) {
    HQ_REQUIRE(n >= 1 && n <= 128 && L >= 0 && L <= 8, "comprehensive similarity supports grids up to 128x128 and up to 8 index rows");
    HQ_REQUIRE(N >= 0 && Q >= 0 && M >= 0, "negative size");
    const int64_t frame_floats = (int64_t)(n + L) * n;
    HQ_REQUIRE(frame_stride >= frame_floats && q_stride >= frame_floats, "frame stride smaller than a frame");
    if (!cand_ids) HQ_REQUIRE(M == N, "without candidate ids M must equal N");
    if (Q == 0 || M == 0) return HQ_OK;
    HQ_REQUIRE(frames && q_frames && out && (L == 0 || weights_host), "null pointer");
    HQ_REQUIRE((int64_t)Q * M < ((int64_t)1 << 31), "too many pairs for one launch");
    BlendParams p{};
    p.frames = frames; p.N = N; p.frame_stride = frame_stride; p.q_frames = q_frames; p.q_stride = q_stride; p.Q = Q; p.n = n; p.L = L;
    for (int l = 0; l < 8; ++l) p.w[l] = l < L ? weights_host[l] : 0.f;
    p.cand_ids = cand_ids; p.M = M; p.out = out;
    const size_t smem = (size_t)2 * frame_floats * sizeof(float);
    static size_t smem_set = 48 * 1024;
    if (smem > smem_set) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_blend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    k_blend<<<(unsigned)((int64_t)Q * M), kBlendThreads, smem, (cudaStream_t)stream>>>(p);
    HQ_LAUNCH_OK("k_blend");
    return HQ_OK;
}

// ---------------------------------------------------------------------------------------
// f1: half-offset squares of PrecomputedHilbertIndexer (core/precomputed_hilbert_index.py:185-203).
// A square of side s at (s/2 + r*s, s/2 + c*s) is exactly four aligned (s/2)-blocks, so its mean is the
// mean of four entries of the next finer level of aligned means (or of four cells when s == 2):
//   out[item, r*(G/2-1) + c] = 0.25 * ((h[2r+1][2c+1] + h[2r+1][2c+2]) + (h[2r+2][2c+1] + h[2r+2][2c+2]))
// with h = the finer level as a row-major [G, G] image.
// ---------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_offset_squares(const float* __restrict__ half, int64_t N, int G, int64_t half_stride,
                                                        float* __restrict__ out, int64_t out_stride) {
    const int g1 = G / 2 - 1;
    const int64_t per = (int64_t)g1 * g1, total = N * per;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t item = t / per;
        const int e = (int)(t - item * per);
        const int r = e / g1, c = e - r * g1;
        const float* h = half + item * half_stride + (int64_t)(2 * r + 1) * G + (2 * c + 1);
        const float a = __ldg(h), b = __ldg(h + 1), cc = __ldg(h + G), d = __ldg(h + G + 1);
        out[item * out_stride + e] = ((a + b) + (cc + d)) * 0.25f;
    }
}
}  // namespace

extern "C" int hq_offset_square_means(const float* half, int64_t N, int G, int64_t half_stride, float* out, int64_t out_stride,
                                      void* stream) {
    HQ_REQUIRE(G >= 2 && G % 2 == 0 && N >= 0, "finer level must be an even-sided square");
    const int64_t per = (int64_t)(G / 2 - 1) * (G / 2 - 1);
    if (N == 0 || per == 0) return HQ_OK;
    HQ_REQUIRE(half && out && half_stride >= (int64_t)G * G && out_stride >= per, "null pointer or stride too small");
    int64_t blocks = (N * per + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_offset_squares<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(half, N, G, half_stride, out, out_stride);
    HQ_LAUNCH_OK("k_offset_squares");
    return HQ_OK;
}

// ---------------------------------------------------------------------------------------
// f4: video-path hierarchical similarity (core/video_storage.py:763-781): (pearson + 1) / 2 of two index
// vectors on their common prefix, clamped to [0, 1]; when a vector has zero variance: 1.0 if the two are
// np.allclose (|a - b| <= 1e-8 + 1e-5 |b|), else 0.0.  All pairs of A [M, S] x B [N, S] in float64 (np.corrcoef
// works in float64), one warp per pair, two passes (means, then centred products) like np.cov.
// ---------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) k_pearson01(const double* __restrict__ A, int64_t M, int64_t a_stride, const double* __restrict__ B,
                                                   int64_t N, int64_t b_stride, int S, double* __restrict__ out, int64_t out_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t pair = (((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; pair < M * N; pair += warps) {
        const int64_t m = pair / N, n = pair - m * N;
        const double* a = A + m * a_stride;
        const double* b = B + n * b_stride;
        double sa = 0.0, sb = 0.0;
        for (int i = lane; i < S; i += 32) { sa += a[i]; sb += b[i]; }
        const double ma = warp_sum_d(sa) / S, mb = warp_sum_d(sb) / S;
        double sab = 0.0, saa = 0.0, sbb = 0.0;
        int close = 1;
        for (int i = lane; i < S; i += 32) {
            const double x = a[i], y = b[i], da = x - ma, db = y - mb;
            sab += da * db; saa += da * da; sbb += db * db;
            if (!(fabs(x - y) <= 1e-8 + 1e-5 * fabs(y))) close = 0;
        }
        sab = warp_sum_d(sab); saa = warp_sum_d(saa); sbb = warp_sum_d(sbb);
        close = __all_sync(0xffffffffu, close);
        double r;
        if (saa == 0.0 || sbb == 0.0) r = close ? 1.0 : 0.0;
        else {
            double c = sab / sqrt(saa) / sqrt(sbb);
            c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);           // np.corrcoef clips to [-1, 1]
            r = (c + 1.0) / 2.0;
            r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
        }
        if (lane == 0) out[m * out_stride + n] = r;
    }
}
}  // namespace

extern "C" int hq_pearson01_matrix(const double* a, int64_t M, int64_t a_stride, const double* b, int64_t N, int64_t b_stride, int S,
                                   double* out, int64_t out_stride, void* stream) {
    HQ_REQUIRE(M >= 0 && N >= 0 && S >= 1, "bad shape");
    if (M == 0 || N == 0) return HQ_OK;
    HQ_REQUIRE(a && b && out && a_stride >= S && b_stride >= S && out_stride >= N, "null pointer or stride too small");
    int64_t blocks = (M * N + 7) / 8;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_pearson01<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, M, a_stride, b, N, b_stride, S, out, out_stride);
    HQ_LAUNCH_OK("k_pearson01");
    return HQ_OK;
}
