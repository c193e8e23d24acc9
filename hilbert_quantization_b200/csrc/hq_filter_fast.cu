// K5 fast path: the RAG progressive filter without a score matrix.
//
//   k_filter_bits     one pass over (queries x rows): the per-level threshold tests of ALL levels
//                     as three independent bit planes P_l[q][row] = (score_l(q,row) >= thr_l).
//                     A thread owns one row (all its index values in registers) and sweeps a
//                     128-query tile out of shared memory, four queries per 128-bit load.
//                     The test is the multiplication form  dot >= (x*_l * |q|) * |c|  of
//                     (dot / (|q||c|) + 1) / 2 >= thr_l  (x*_l = smallest float whose score
//                     reaches thr_l), so no division / square root per pair.
//   k_filter_cascade  one CTA per query walks the levels: alive &= P_l, counts, and ONLY when the
//                     ratio cut binds (count > cap) ranks the surviving rows exactly: their fp32
//                     scores are recomputed from the index rows (same arithmetic as the exact
//                     path in hq_search.cu), compacted to a scratch list, and the (count - cap)
//                     lowest (score asc, row id desc) are cleared with an exact radix select.
//
// Valid when every row / query index length equals the structural length of its level
// (dense data: hq_index_row_lengths == lvl_keff everywhere); otherwise the caller uses the
// exact per-level path (hq_filter_level / hq_filter_select).
// Reference semantics: rag/search/engine.py:178-287.
#include "hq_common.cuh"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kRows = 256;       // rows per CTA (one per thread)
constexpr int kQT = 128;         // queries per CTA

struct BitsParams {
    const float* idx;            // [N, Lsum]
    const float* rnorm;          // [N, L] row norms per level, NaN where the norm is 0
    int64_t N;
    hq_index_layout lay;
    const float* q_idx;          // [Q, Lsum]
    int Q;
    float xstar[3];
    uint32_t* bits;              // [L][Q][pitch]
    int64_t words, pitch;
    int q_tiles;
};

template <int K>
__device__ __forceinline__ void load_row(const float* __restrict__ p, int keff, float (&c)[K]) {
#pragma unroll
    for (int j = 0; j < K; j += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < keff) v = __ldg(reinterpret_cast<const float4*>(p + j));
        c[j] = v.x; c[j + 1] = v.y; c[j + 2] = v.z; c[j + 3] = v.w;
    }
}

template <int K>
__device__ __forceinline__ void dot4(const float (&c)[K], const float* __restrict__ sq /*[K][kQT]*/, int qq, float (&acc)[4]) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const float4 qv = *reinterpret_cast<const float4*>(sq + j * kQT + qq);
        acc[0] = fmaf(c[j], qv.x, acc[0]);
        acc[1] = fmaf(c[j], qv.y, acc[1]);
        acc[2] = fmaf(c[j], qv.z, acc[2]);
        acc[3] = fmaf(c[j], qv.w, acc[3]);
    }
}

template <int K0, int K1, int K2>
__global__ void __launch_bounds__(kRows, 2) k_filter_bits(const BitsParams p) {
    constexpr int KT = K0 + K1 + K2;
    extern __shared__ __align__(16) float sm[];
    float* s_q = sm;                         // [KT][kQT] transposed query tile, levels stacked
    float* s_tq = sm + KT * kQT;             // [3][kQT]   x*_l * |q_l|  (NaN when |q_l| == 0)

    const int tid = threadIdx.x, lane = tid & 31;
    const int q_tile = blockIdx.x % p.q_tiles;
    const int64_t row0 = (int64_t)(blockIdx.x / p.q_tiles) * kRows;
    const int q0 = q_tile * kQT;
    const int L = p.lay.L;

    // stage the query tile
    for (int e = tid; e < kQT * KT; e += kRows) {
        const int qq = e / KT, j = e - qq * KT;
        const int q = q0 + qq;
        int l = 0, jj = j;
        if (j >= K0) { l = 1; jj = j - K0; }
        if (j >= K0 + K1) { l = 2; jj = j - K0 - K1; }
        float v = 0.f;
        if (q < p.Q && l < L && jj < p.lay.lvl_keff[l]) v = __ldg(p.q_idx + (int64_t)q * p.lay.Lsum + p.lay.lvl_off[l] + jj);
        s_q[j * kQT + qq] = v;
    }
    __syncthreads();
    if (tid < kQT) {
        const int kk[3] = {K0, K1, K2};
        int base = 0;
        for (int l = 0; l < 3; ++l) {
            float c = 0.f;
            for (int j = 0; j < kk[l]; ++j) { const float v = s_q[(base + j) * kQT + tid]; c = fmaf(v, v, c); }
            const float nq = sqrtf(c);
            s_tq[l * kQT + tid] = nq > 0.f ? __fmul_rn(p.xstar[l], nq) : __int_as_float(0x7fc00000);
            base += kk[l];
        }
    }

    const int64_t row = row0 + tid;
    const bool in_range = row < p.N;
    const int64_t rr = in_range ? row : 0;
    float c0[K0], c1[K1 > 0 ? K1 : 4], c2[K2 > 0 ? K2 : 4];
    load_row<K0>(p.idx + rr * p.lay.Lsum + p.lay.lvl_off[0], p.lay.lvl_keff[0], c0);
    float n0 = __ldg(p.rnorm + rr * L), n1 = 0.f, n2 = 0.f;
    if constexpr (K1 > 0) { load_row<K1>(p.idx + rr * p.lay.Lsum + p.lay.lvl_off[1], p.lay.lvl_keff[1], c1); n1 = __ldg(p.rnorm + rr * L + 1); }
    if constexpr (K2 > 0) { load_row<K2>(p.idx + rr * p.lay.Lsum + p.lay.lvl_off[2], p.lay.lvl_keff[2], c2); n2 = __ldg(p.rnorm + rr * L + 2); }
    __syncthreads();

    const int64_t word = row >> 5;
    const bool word_ok = (row0 + (tid & ~31)) < p.N;
    uint32_t* b0 = p.bits;
    uint32_t* b1 = p.bits + (int64_t)p.Q * p.pitch;
    uint32_t* b2 = p.bits + 2 * (int64_t)p.Q * p.pitch;
    for (int qq = 0; qq < kQT; qq += 4) {
        if (q0 + qq >= p.Q) break;
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
        dot4<K0>(c0, s_q, qq, a0);
        if constexpr (K1 > 0) dot4<K1>(c1, s_q + K0 * kQT, qq, a1);
        if constexpr (K2 > 0) dot4<K2>(c2, s_q + (K0 + K1) * kQT, qq, a2);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = q0 + qq + u;
            if (q >= p.Q) break;
            const bool p0 = in_range && (a0[u] >= __fmul_rn(s_tq[qq + u], n0));
            const uint32_t w0 = __ballot_sync(0xffffffffu, p0);
            uint32_t w1 = 0, w2 = 0;
            if (K1 > 0) w1 = __ballot_sync(0xffffffffu, in_range && (a1[u] >= __fmul_rn(s_tq[kQT + qq + u], n1)));
            if (K2 > 0) w2 = __ballot_sync(0xffffffffu, in_range && (a2[u] >= __fmul_rn(s_tq[2 * kQT + qq + u], n2)));
            if (lane == 0 && word_ok) {
                b0[(int64_t)q * p.pitch + word] = w0;
                if (K1 > 0) b1[(int64_t)q * p.pitch + word] = w1;
                if (K2 > 0) b2[(int64_t)q * p.pitch + word] = w2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// row norms per level (sequential fmaf order, identical to the exact path); NaN marks zero
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_level_norms(const float* __restrict__ idx, int64_t N, hq_index_layout lay,
                                                     float* __restrict__ rnorm, int32_t* __restrict__ nonuniform,
                                                     const uint16_t* __restrict__ lens) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < N * lay.L; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / lay.L;
        const int l = (int)(t - row * lay.L);
        const float* r = idx + row * lay.Lsum + lay.lvl_off[l];
        float c = 0.f;
        for (int j = 0; j < lay.lvl_keff[l]; ++j) { const float v = __ldg(r + j); c = fmaf(v, v, c); }
        const float nrm = sqrtf(c);
        rnorm[t] = nrm > 0.f ? nrm : __int_as_float(0x7fc00000);
        if (lens && (int)lens[t] != lay.lvl_keff[l]) atomicOr(nonuniform, 1);
    }
}

__global__ void __launch_bounds__(256) k_query_norms(const float* __restrict__ q_idx, int Q, hq_index_layout lay, float* __restrict__ nq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Q * lay.L) return;
    const int l = t / Q, q = t - l * Q;
    const float* r = q_idx + (int64_t)q * lay.Lsum + lay.lvl_off[l];
    float c = 0.f;
    for (int j = 0; j < lay.lvl_keff[l]; ++j) { const float v = __ldg(r + j); c = fmaf(v, v, c); }
    nq[(int64_t)l * Q + q] = sqrtf(c);
}

// ---------------------------------------------------------------------------------------
// block-wide exact selection on a compact key list (k-th SMALLEST), warp-aggregated histograms
// ---------------------------------------------------------------------------------------
struct SelSmall { uint32_t key; uint32_t below; uint32_t equal; };   // #keys < key, #keys == key

__device__ SelSmall block_select_smallest(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ filt, uint32_t filt_val,
                                          uint32_t n, uint32_t k, uint32_t* hist /*2048*/, uint32_t* sh /*4*/) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    uint32_t prefix = 0, pmask = 0, remaining = k, below_total = 0, equal = 0;
    const int shifts[3] = {21, 10, 0};
    const int nbits[3] = {11, 11, 10};
    const uint32_t n_pad = (n + 31u) & ~31u;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = shifts[pass];
        const uint32_t nb = 1u << nbits[pass];
        for (uint32_t i = tid; i < nb; i += nt) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n_pad; i += nt) {
            bool valid = i < n;
            uint32_t key = 0;
            if (valid) {
                key = keys[i];
                valid = (key & pmask) == prefix && (!filt || filt[i] == filt_val);
            }
            const uint32_t bin = (key >> shift) & (nb - 1);
            const uint32_t act = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const uint32_t peers = __match_any_sync(act, bin);
                if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        if (tid < 32) {
            const uint32_t seg = nb / 32;
            uint32_t sum = 0;
            for (uint32_t b = 0; b < seg; ++b) sum += hist[tid * seg + b];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            const uint32_t excl = incl - sum;
            if (excl < remaining && remaining <= incl) {
                uint32_t below = excl, b = tid * seg;
                for (;; ++b) {
                    const uint32_t h = hist[b];
                    if (below + h >= remaining) break;
                    below += h;
                }
                sh[0] = b; sh[1] = remaining - below; sh[2] = hist[b]; sh[3] = below;
            }
        }
        __syncthreads();
        prefix |= sh[0] << shift;
        pmask |= (nb - 1) << shift;
        remaining = sh[1];
        equal = sh[2];
        below_total += sh[3];
        __syncthreads();
    }
    SelSmall r;
    r.key = prefix; r.below = below_total; r.equal = equal;
    return r;
}

struct CascadeParams {
    const uint32_t* bits;        // [L][Q][bits_pitch]
    int64_t words, bits_pitch;
    const float* idx;            // [N, Lsum]
    const float* lvl[3];         // optional per-level copies [N, lvl_pitch[l]] (small enough to stay in L2)
    int lvl_pitch[3];
    int64_t N;
    hq_index_layout lay;
    const float* q_idx;          // [Q, Lsum]
    int Q;
    double ratio[8];
    uint32_t key_lo[8];          // smallest score bit pattern a surviving row of the level can have
    uint32_t shift_a[8];         // right shift that maps (key - key_lo) of the level onto 2048 bins
    uint32_t* mask;              // [Q, mask_stride] out
    int64_t mask_stride;
    int32_t* counts;             // [L][3][Q] (n_alive, n_pass, n_out) or null
    int32_t* n_out;              // [Q]
    uint32_t* scratch_keys;      // [gridDim][N]
    uint32_t* scratch_rows;      // [gridDim][N]
    const int32_t* only;         // optional [Q]: process only queries with a non-zero flag
    const uint16_t* lens;        // optional [N, L] stored row lengths (rows shorter than lvl_keff exist)
};

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_warp[w] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int i = 0; i < nw; ++i) t += s_warp[i];
    __syncthreads();
    return t;
}

// exact fp32 score of one (query level row, database row) pair -- same arithmetic as hq_search.cu
// `m` < keff: the row's stored length (trailing zeros stripped) is shorter than the structural one -- the
// reference then takes the query norm over that prefix only (rag/search/engine.py:216-227)
__device__ __forceinline__ uint32_t level_key(const float* __restrict__ rp, const float* __restrict__ s_q, int keff, float nq, int m = 1 << 30) {
    if (m < keff) {
        float c = 0.f;
        for (int j = 0; j < m; ++j) c = fmaf(s_q[j], s_q[j], c);
        nq = sqrtf(c);
    }
    float dot = 0.f, cn2 = 0.f;
    for (int j = 0; j < keff; j += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(rp + j));
        dot = fmaf(v.x, s_q[j], dot); cn2 = fmaf(v.x, v.x, cn2);
        dot = fmaf(v.y, s_q[j + 1], dot); cn2 = fmaf(v.y, v.y, cn2);
        dot = fmaf(v.z, s_q[j + 2], dot); cn2 = fmaf(v.z, v.z, cn2);
        dot = fmaf(v.w, s_q[j + 3], dot); cn2 = fmaf(v.w, v.w, cn2);
    }
    const float nc = sqrtf(cn2);
    float sc = 0.f;
    if (nq != 0.f && nc != 0.f) sc = __fmul_rn(__fadd_rn(__fdiv_rn(dot, __fmul_rn(nq, nc)), 1.0f), 0.5f);
    if (sc < 0.f) sc = 0.f;
    return __float_as_uint(sc);
}

constexpr int kBufCap = 4096;     // candidates of the cut bin that are ranked in shared memory

__global__ void __launch_bounds__(1024, 1) k_filter_cascade(const CascadeParams p) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t sh[4];
    __shared__ uint32_t s_warp[32];
    __shared__ float s_q[64];
    __shared__ float s_nq;
    __shared__ uint32_t s_count, s_bufn;
    __shared__ uint32_t b_key[kBufCap], b_row[kBufCap];
    const int tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    const int L = p.lay.L;
    uint32_t* keys = p.scratch_keys + (int64_t)blockIdx.x * p.N;
    uint32_t* rows = p.scratch_rows + (int64_t)blockIdx.x * p.N;
    const int64_t words_pad = (p.words + 31) & ~(int64_t)31;

    for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
        if (p.only && !p.only[q]) continue;
        uint32_t* M = p.mask + (int64_t)q * p.mask_stride;
        int64_t n_alive = p.N;
        for (int l = 0; l < L; ++l) {
            // ---- alive &= P_l, count (4 independent word pairs in flight per thread) ----
            const uint32_t* P = p.bits + ((int64_t)l * p.Q + q) * p.bits_pitch;
            uint32_t c = 0;
            for (int64_t w0 = tid; w0 < p.words; w0 += 4 * (int64_t)nt) {
                uint32_t a[4], pb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t w = w0 + (int64_t)u * nt;
                    a[u] = 0; pb[u] = 0;
                    if (w < p.words) {
                        pb[u] = __ldg(P + w);
                        if (l == 0) {
                            const int64_t r0 = w * 32;
                            a[u] = r0 + 32 <= p.N ? 0xffffffffu : (r0 < p.N ? ((1u << (uint32_t)(p.N - r0)) - 1u) : 0u);
                        } else {
                            a[u] = __ldcg(M + w);            // rows cleared by other threads' atomics live in L2
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t w = w0 + (int64_t)u * nt;
                    if (w < p.words) { const uint32_t cand = a[u] & pb[u]; M[w] = cand; c += __popc(cand); }
                }
            }
            const uint32_t c_total = block_sum(c, s_warp);       // barriers inside: M is complete
            int64_t cap = (int64_t)((double)n_alive * p.ratio[l]);
            if (cap < 1) cap = 1;
            if (p.counts && tid == 0) {
                p.counts[((int64_t)l * 3 + 0) * p.Q + q] = (int32_t)n_alive;
                p.counts[((int64_t)l * 3 + 1) * p.Q + q] = (int32_t)c_total;
                p.counts[((int64_t)l * 3 + 2) * p.Q + q] = (int32_t)((int64_t)c_total > cap ? cap : c_total);
            }
            if ((int64_t)c_total <= cap) { n_alive = c_total; continue; }

            // ======== the ratio cut binds: drop the (c_total - cap) lowest (score asc, row id desc) ========
            const uint32_t drop = c_total - (uint32_t)cap;
            const int keff = p.lay.lvl_keff[l];
            if (tid < 64) s_q[tid] = tid < keff ? __ldg(p.q_idx + (int64_t)q * p.lay.Lsum + p.lay.lvl_off[l] + tid) : 0.f;
            if (tid == 0) { s_count = 0; s_bufn = 0; }
            for (int i = tid; i < 2048; i += nt) hist[i] = 0;
            __syncthreads();
            if (tid == 0) {
                float cq = 0.f;
                for (int j = 0; j < keff; ++j) cq = fmaf(s_q[j], s_q[j], cq);
                s_nq = sqrtf(cq);
            }
            // ---- A: compact the surviving row ids (bits only) ----
            for (int64_t w = tid; w < words_pad; w += nt) {
                uint32_t cand = w < p.words ? __ldcg(M + w) : 0u;
                const uint32_t cnt = __popc(cand);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t base = 0;
                if (lane == 0 && total) base = atomicAdd(&s_count, total);
                base = __shfl_sync(0xffffffffu, base, 0);
                uint32_t pos = base + incl - cnt;
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    rows[pos++] = (uint32_t)(w * 32 + b);
                }
            }
            __syncthreads();
            const uint32_t n_list = s_count;
            const float nq = s_nq;
            const float* lvl_base = l < 3 ? p.lvl[l] : nullptr;
            const int64_t lvl_pitch = l < 3 ? p.lvl_pitch[l] : 0;
            // ---- B: exact keys (4 independent rows in flight per thread) + histogram of the top digit ----
            // scores live in [thr - eps, 1]: bin on the offset from the smallest possible key so that the
            // 11-bit digit spreads; clamping only affects binning, never the order (true keys are kept)
            const uint32_t key_lo = p.key_lo[l], shiftA = p.shift_a[l];
            const uint32_t n_pad4 = (n_list + 4 * nt - 1) / (4 * nt) * (4 * nt);
            for (uint32_t i0 = tid; i0 < n_pad4; i0 += 4 * nt) {
                uint32_t k4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t i = i0 + u * nt;
                    k4[u] = 0;
                    if (i < n_list) {
                        const int64_t row = __ldcg(rows + i);
                        const float* rp = lvl_base ? lvl_base + row * lvl_pitch : p.idx + row * p.lay.Lsum + p.lay.lvl_off[l];
                        k4[u] = level_key(rp, s_q, keff, nq, p.lens ? (int)p.lens[row * L + l] : (1 << 30));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t i = i0 + u * nt;
                    const bool valid = i < n_list;
                    if (valid) keys[i] = k4[u];
                    uint32_t offk = k4[u] > key_lo ? k4[u] - key_lo : 0u;
                    uint32_t bin = offk >> shiftA;
                    if (bin > 2047u) bin = 2047u;
                    const uint32_t act = __ballot_sync(0xffffffffu, valid);
                    if (valid) {
                        const uint32_t peers = __match_any_sync(act, bin);
                        if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
                    }
                }
            }
            __syncthreads();
            // ---- locate the bin holding the drop-th smallest ----
            if (tid < 32) {
                uint32_t sum = 0;
                for (uint32_t b = 0; b < 64; ++b) sum += hist[tid * 64 + b];
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                const uint32_t excl = incl - sum;
                if (excl < drop && drop <= incl) {
                    uint32_t below = excl, b = tid * 64;
                    for (;; ++b) {
                        const uint32_t h = hist[b];
                        if (below + h >= drop) break;
                        below += h;
                    }
                    sh[0] = b; sh[1] = drop - below; sh[2] = hist[b];
                }
            }
            __syncthreads();
            const uint32_t cut_bin = sh[0], r_in_bin = sh[1], n_in_bin = sh[2];
            if (n_in_bin <= (uint32_t)kBufCap) {
                // ---- C: lower bins are dropped outright, the cut bin is ranked in shared memory ----
                for (uint32_t i0 = tid; i0 < n_pad4; i0 += 4 * nt) {
                    uint32_t k4[4], r4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t i = i0 + u * nt;
                        k4[u] = 0; r4[u] = 0;
                        if (i < n_list) { k4[u] = __ldcg(keys + i); r4[u] = __ldcg(rows + i); }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t i = i0 + u * nt;
                        if (i >= n_list) continue;
                        uint32_t offk = k4[u] > key_lo ? k4[u] - key_lo : 0u;
                        uint32_t bin = offk >> shiftA;
                        if (bin > 2047u) bin = 2047u;
                        if (bin < cut_bin) atomicAnd(&M[r4[u] >> 5], ~(1u << (r4[u] & 31)));
                        else if (bin == cut_bin) {
                            const uint32_t slot = atomicAdd(&s_bufn, 1u);
                            b_key[slot] = k4[u]; b_row[slot] = r4[u];
                        }
                    }
                }
                __syncthreads();
                const uint32_t nb = s_bufn;
                for (uint32_t e = tid; e < nb; e += nt) {
                    const uint32_t key = b_key[e], row = b_row[e];
                    uint32_t rank = 0;                                    // entries that go before this one
                    for (uint32_t j = 0; j < nb; ++j) {
                        const uint32_t kj = b_key[j], rj = b_row[j];
                        rank += (kj < key || (kj == key && rj > row)) ? 1u : 0u;
                    }
                    if (rank < r_in_bin) atomicAnd(&M[row >> 5], ~(1u << (row & 31)));
                }
            } else {
                // ---- heavily tied scores: generic exact selection over the whole list ----
                const SelSmall sel = block_select_smallest(keys, nullptr, 0, n_list, drop, hist, sh);
                const uint32_t t_d = drop - sel.below;                 // ties to drop (1..equal)
                uint32_t id_cut = 0;                                   // ties with row >= id_cut are dropped
                if (t_d < sel.equal) {
                    const SelSmall ids = block_select_smallest(rows, keys, sel.key, n_list, sel.equal - t_d + 1, hist, sh);
                    id_cut = ids.key;
                }
                for (uint32_t i = tid; i < n_list; i += nt) {
                    const uint32_t key = keys[i], row = rows[i];
                    if (key < sel.key || (key == sel.key && row >= id_cut)) atomicAnd(&M[row >> 5], ~(1u << (row & 31)));
                }
            }
            __syncthreads();
            n_alive = cap;
        }
        if (tid == 0) p.n_out[q] = (int32_t)n_alive;
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------
// Cascade over the candidate lists written by the tensor-core pass (hq_filter_tc.cu).
//
// k_filter_cascade re-ranks by GATHERING the index rows of ~200 K survivors per query out of
// L2 (4.2 ms per 1024-query batch, long-scoreboard bound).  The tensor-core pass already has
// every level's dot product in registers, so it appends (row, k1, k2) of the rows that pass
// the level-0 and level-1 thresholds to per-query lists; within a query the dot product
// orders rows exactly like the score ((dot / |q| + 1) / 2 is monotone), so the ratio cuts
// become streaming selections over compact, coalesced arrays:
//     level 1: keep the cap1 best of the list by k1 (ties -> lower row id)
//     level 2: of those, the rows that pass the level-2 threshold; keep the cap2 best by k2
// Selection = 2048-bin LINEAR histogram between the level's threshold and |q| (the largest
// possible dot product), then an exact ranking of the cut bin in shared memory; the result
// is a boundary (K, R): kept <=> k > K or (k == K and row < R).
// Queries it cannot handle (level-0 cut binds, list overflow, a cut bin larger than the
// buffer, L == 1) set fallback[q]; k_filter_cascade then runs for exactly those.
// ---------------------------------------------------------------------------------------
struct ListParams {
    const uint32_t* bits;        // [L][Q][bits_pitch] (plane 0 is counted)
    int64_t words, bits_pitch;
    int64_t N;
    int L, Q;
    double ratio[8];
    const float* tq;             // [3][Q] threshold in dot-product units
    const float* nq;             // [3][Q] |q_l|
    const uint32_t* l_rows;
    const float* l_k1;
    const float* l_k2;
    const int32_t* seg_n;
    int64_t seg_cap;
    int n_segs;
    uint32_t* mask;              // [Q, mask_stride], zero-filled by the caller
    int64_t mask_stride;
    int32_t* counts;             // optional [L][3][Q]
    int32_t* n_out;              // [Q]
    int32_t* fallback;           // [Q] out
    const int32_t* only;         // optional [Q]: process only queries with a non-zero flag (second stage of the window mode);
                                 // their mask rows are zero-filled here first
    float* tmp_keys;             // [gridDim][tmp_stride] compacted level-2 candidates of the query in flight
    uint32_t* tmp_rows;          // [gridDim][tmp_stride]
    int64_t tmp_stride;          // >= n_segs * seg_cap
};

constexpr int kMaxSegs = 320;
constexpr int kCutCap = 2048;
constexpr int kListThreads = 1024;    // one query per SM at a time (two 512-thread CTAs per SM measured 7 % slower at 1 M rows)
constexpr int64_t kListSmallShardRows = 400000;   // shards up to this many rows run two 512-thread CTAs per SM instead
constexpr int kU = 4;                 // independent list loads in flight per thread (128-bit each in the list passes)
constexpr int kU2 = 2;                // same for the compaction pass, which reads three arrays
constexpr uint32_t kChunk = kU * 128, kChunk2 = kU2 * 128;      // entries a warp takes at a time

struct Boundary { float K; uint32_t R; };      // kept <=> k > K || (k == K && row < R)
struct CutBin { float lo, scale; uint32_t cb, r_in_bin; bool on; };   // cut bin of a 2048-bin histogram, ranked later
__device__ __forceinline__ bool kept_by(const Boundary& b, float k, uint32_t row) { return k > b.K || (k == b.K && row < b.R); }

__device__ __forceinline__ uint32_t lin_bin(float k, float lo, float scale) {
    const float x = (k - lo) * scale;
    int b = (int)x;
    b = b < 0 ? 0 : b;
    return b > 2047 ? 2047u : (uint32_t)b;
}

// bin holding the drop-th smallest element of a 2048-bin histogram: sh[0] = bin, sh[1] = rank inside it (1-based),
// sh[2] = population of the bin
__device__ __forceinline__ void find_cut_bin(const uint32_t* hist, uint32_t drop, uint32_t* sh) {
    // Block-wide scan, every thread owns 2048 / blockDim.x consecutive bins (blockDim.x = 256, 512 or 1024).  The first
    // version let 32 threads sum 64 bins each (all lanes in one bank: 32-way conflicts) and one lane walk its 64 bins:
    // ~4 us per call with every other warp parked at the barrier, 10 % of the kernel on a 125 K-row shard.
    __shared__ uint32_t s_scan[32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int per = 2048 / nt;                         // <= 16: blockDim.x >= 128
    uint32_t v[16], sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i] = i < per ? hist[tid * per + i] : 0u;
        sum += v[i];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < nw ? s_scan[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_scan[lane] = wi - w;                         // exclusive prefix of the warp totals
    }
    __syncthreads();
    const uint32_t excl = s_scan[warp] + incl - sum;
    if (excl < drop && drop <= excl + sum) {
        uint32_t below = excl;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < per && below < drop) {
                if (below + v[i] >= drop) { sh[0] = (uint32_t)(tid * per + i); sh[1] = drop - below; sh[2] = v[i]; }
                below += v[i];
            }
        }
    }
    __syncthreads();
}

// exact boundary inside the cut bin: the r_in_bin lowest (key asc, row desc) entries of the buffer are dropped
__device__ __forceinline__ void rank_cut_bin(const float* b_key, const uint32_t* b_row, uint32_t nb, uint32_t r_in_bin, float* s_K,
                                             uint32_t* s_R) {
    for (uint32_t e = threadIdx.x; e < nb; e += blockDim.x) {
        const float key = b_key[e];
        const uint32_t row = b_row[e];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < nb; ++j) {
            const float kj = b_key[j];
            const uint32_t rj = b_row[j];
            rank += (kj < key || (kj == key && rj > row)) ? 1u : 0u;
        }
        if (rank == r_in_bin - 1) { *s_K = key; *s_R = row; }     // the last dropped entry
    }
    __syncthreads();
}

// Exact boundary of "drop the `drop` lowest (key asc, row desc) candidates": 2048-bin linear histogram over
// [lo, hi), then either an exact ranking of the cut bin in shared memory (<= kCutCap members) or a refinement
// of that bin into 2048 sub-bins (large shards put thousands of candidates into one bin).  `for_each(visit)`
// must call visit(key, token) for every candidate of the calling thread, `row_of(token)` returns its row id.
// Returns false when three rounds cannot isolate the boundary (heavily tied keys): the caller falls back.
// (scalars, not arrays: indexing arrays by the round number put the struct into local memory and an LDL into
// the per-candidate loops)
struct BinStack { float lo0, lo1, lo2, sc0, sc1, sc2; uint32_t cb0, cb1, cb2; int depth; };
__device__ __forceinline__ bool in_bins(const BinStack& s, float k) {
    bool ok = true;
    if (s.depth > 0) ok = lin_bin(k, s.lo0, s.sc0) == s.cb0;
    if (s.depth > 1) ok = ok && lin_bin(k, s.lo1, s.sc1) == s.cb1;
    if (s.depth > 2) ok = ok && lin_bin(k, s.lo2, s.sc2) == s.cb2;
    return ok;
}

template <class ForEach, class RowOf>
__device__ __forceinline__ bool select_boundary(ForEach for_each, RowOf row_of, uint32_t drop, float lo, float hi, bool have_hist0,
                                                uint32_t* hist, uint32_t* sh, float* b_key, uint32_t* b_row, uint32_t* s_bufn,
                                                float* s_K, uint32_t* s_R, Boundary& out, uint32_t cut_cap = (uint32_t)kCutCap) {
    const int tid = threadIdx.x, nt = blockDim.x;
    BinStack st{};
    float cur_lo = lo, cur_hi = hi;
    for (int round = 0; round < 3; ++round) {
        const float scale = cur_hi > cur_lo ? 2048.0f / (cur_hi - cur_lo) : 0.f;
        if (!(round == 0 && have_hist0)) {
            __syncthreads();
            for (int i = tid; i < 2048; i += nt) hist[i] = 0;
            __syncthreads();
            for_each([&](float k, uint32_t) {
                if (in_bins(st, k)) atomicAdd(&hist[lin_bin(k, cur_lo, scale)], 1u);
            });
            __syncthreads();
        }
        find_cut_bin(hist, drop, sh);
        const uint32_t cb = sh[0], r_in_bin = sh[1], n_in_bin = sh[2];
        if (round == 0) { st.lo0 = cur_lo; st.sc0 = scale; st.cb0 = cb; }
        else if (round == 1) { st.lo1 = cur_lo; st.sc1 = scale; st.cb1 = cb; }
        else { st.lo2 = cur_lo; st.sc2 = scale; st.cb2 = cb; }
        st.depth = round + 1;
        if (n_in_bin <= cut_cap) {
            if (tid == 0) *s_bufn = 0;
            __syncthreads();
            for_each([&](float k, uint32_t tok) {
                if (in_bins(st, k)) {
                    const uint32_t slot = atomicAdd(s_bufn, 1u);
                    b_key[slot] = k; b_row[slot] = row_of(tok);
                }
            });
            __syncthreads();
            rank_cut_bin(b_key, b_row, *s_bufn, r_in_bin, s_K, s_R);
            out.K = *s_K; out.R = *s_R;
            return true;
        }
        drop = r_in_bin;                                   // rank inside the cut bin
        const float w = (cur_hi - cur_lo) * (1.0f / 2048.0f);
        cur_lo = cur_lo + w * (float)cb;
        cur_hi = cur_lo + w;
        if (!(cur_hi > cur_lo)) return false;
    }
    return false;
}

// warp-aggregated append position in a block-wide compact list (count kept in shared memory)
__device__ __forceinline__ uint32_t compact_slot(bool take, uint32_t* s_count) {
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(s_count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kListThreads, 1) k_filter_cascade_lists(const ListParams p) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t sh[4];
    __shared__ uint32_t s_warp[32];
    __shared__ int32_t s_cnt[kMaxSegs];
    __shared__ float b_key[kCutCap];
    __shared__ uint32_t b_row[kCutCap];
    __shared__ float b_k2[kCutCap];
    __shared__ uint32_t s_bufn, s_flag, s_R, s_n2, s_nd, s_d2;
    __shared__ float s_K;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    // level-2 candidates of the query this CTA works on, compacted by the level-1 pass
    float* const c_k2 = p.tmp_keys + (int64_t)blockIdx.x * p.tmp_stride;
    uint32_t* const c_row = p.tmp_rows + (int64_t)blockIdx.x * p.tmp_stride;

    for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
        if (p.only) {
            if (!__ldg(p.only + q)) continue;
            for (int64_t w = tid; w < p.words; w += nt) p.mask[(int64_t)q * p.mask_stride + w] = 0u;
        }
        // ---- segment table, overflow check ----
        if (tid == 0) { s_flag = 0; s_n2 = 0; s_bufn = 0; }
        __syncthreads();
        uint32_t part = 0;
        for (int i = tid; i < p.n_segs; i += nt) {
            const int32_t n = p.seg_n[(int64_t)q * p.n_segs + i];
            if (n > p.seg_cap) s_flag = 1;
            s_cnt[i] = n;
            part += (uint32_t)n;
        }
        const uint32_t n1 = block_sum(part, s_warp);
        // ---- level 0: survivors of the threshold; its ratio cut must not bind ----
        // (the same pass counts the rows that pass EVERY level's threshold: when no ratio cut binds, or cut 1 drops only a
        // few rows, the survivor mask is the AND of the planes and the lists are not streamed at all -- see "lazy path")
        const uint32_t* P0 = p.bits + (int64_t)q * p.bits_pitch;
        const uint32_t* P1 = p.L > 1 ? p.bits + ((int64_t)1 * p.Q + q) * p.bits_pitch : nullptr;
        const uint32_t* P2 = p.L > 2 ? p.bits + ((int64_t)2 * p.Q + q) * p.bits_pitch : nullptr;
        const int64_t groups = (p.words + 3) >> 2;                    // plane rows are 32-byte aligned and padded to 8 words
        auto and_group = [&](int64_t g, uint32_t (&v0)[4], uint32_t (&va)[4]) {
            const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(P0) + g);
            uint4 a1 = make_uint4(~0u, ~0u, ~0u, ~0u), a2 = a1;
            if (P1) a1 = __ldg(reinterpret_cast<const uint4*>(P1) + g);
            if (P2) a2 = __ldg(reinterpret_cast<const uint4*>(P2) + g);
            const uint32_t x0[4] = {a0.x, a0.y, a0.z, a0.w}, x1[4] = {a1.x, a1.y, a1.z, a1.w}, x2[4] = {a2.x, a2.y, a2.z, a2.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t r0 = (g * 4 + i) * 32;
                const uint32_t tm = r0 + 32 <= p.N ? 0xffffffffu : (r0 < p.N ? ((1u << (uint32_t)(p.N - r0)) - 1u) : 0u);
                v0[i] = x0[i] & tm;
                va[i] = v0[i] & x1[i] & x2[i];
            }
        };
        uint32_t c = 0, ca = 0;
        for (int64_t g = tid; g < groups; g += nt) {
            uint32_t v0[4], va[4];
            and_group(g, v0, va);
#pragma unroll
            for (int i = 0; i < 4; ++i) { c += __popc(v0[i]); ca += __popc(va[i]); }
        }
        const uint32_t c0 = block_sum(c, s_warp);
        const uint32_t c_all = block_sum(ca, s_warp);                 // rows passing the thresholds of all levels
        int64_t cap0 = (int64_t)((double)p.N * p.ratio[0]);
        if (cap0 < 1) cap0 = 1;
        if (s_flag || (int64_t)c0 > cap0 || p.L < 2) {
            if (tid == 0) p.fallback[q] = 1;
            __syncthreads();
            continue;
        }
        const int64_t qbase = (int64_t)q * p.n_segs * p.seg_cap;
        const uint32_t cps = ((uint32_t)p.seg_cap + kChunk - 1) / kChunk, n_items = (uint32_t)p.n_segs * cps;
        const uint32_t cps2 = ((uint32_t)p.seg_cap + kChunk2 - 1) / kChunk2, n_items2 = (uint32_t)p.n_segs * cps2;
        const uint32_t* const L_rows = p.l_rows + qbase;
        const float* const L_k1 = p.l_k1 + qbase;
        const float* const L_k2 = p.l_k2 ? p.l_k2 + qbase : nullptr;

        // ---- level 1: ratio cut over the whole list by k1 ----
        int64_t cap1 = (int64_t)((double)c0 * p.ratio[1]);
        if (cap1 < 1) cap1 = 1;
        Boundary b1;
        b1.K = -INFINITY; b1.R = 0;
        bool failed = false;
        // every candidate of the query: a warp owns one chunk of a segment at a time (kU x 128 entries), each lane
        // issues kU independent 128-bit loads (4-byte loads kept only 16 KB in flight per SM: latency bound)
        auto each_l1 = [&](auto visit) {
            for (uint32_t item = warp; item < n_items; item += nw) {
                const uint32_t seg = item / cps, e_base = (item - seg * cps) * kChunk;
                const uint32_t cnt = (uint32_t)s_cnt[seg];
                if (e_base >= cnt) continue;
                const uint32_t off = seg * (uint32_t)p.seg_cap;
                float4 k[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t e = e_base + u * 128 + lane * 4;
                    k[u] = e < cnt ? __ldg(reinterpret_cast<const float4*>(L_k1 + off + e)) : make_float4(-1.f, -1.f, -1.f, -1.f);
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t e = e_base + u * 128 + lane * 4;
                    if (e < cnt) visit(k[u].x, off + e);
                    if (e + 1 < cnt) visit(k[u].y, off + e + 1);
                    if (e + 2 < cnt) visit(k[u].z, off + e + 2);
                    if (e + 3 < cnt) visit(k[u].w, off + e + 3);
                }
            }
        };
        // Ratio cut 1.  One histogram pass finds the cut bin; when that bin is small enough to be ranked in shared
        // memory (the normal case) its members are collected by the NEXT pass over the list, which classifies every
        // candidate as above / inside / below the cut bin -- no separate gather pass over the list.
        const bool cut1 = (int64_t)n1 > cap1;
        CutBin cb1{0.f, 0.f, 0u, 0u, false};
        if (cut1) {
            const float lo = __ldg(p.tq + (int64_t)1 * p.Q + q), hi = __ldg(p.nq + (int64_t)1 * p.Q + q);
            const float scale = hi > lo ? 2048.0f / (hi - lo) : 0.f;
            __syncthreads();
            for (int i = tid; i < 2048; i += nt) hist[i] = 0;
            __syncthreads();
            each_l1([&](float k, uint32_t) { atomicAdd(&hist[lin_bin(k, lo, scale)], 1u); });
            __syncthreads();
            find_cut_bin(hist, n1 - (uint32_t)cap1, sh);
            if (sh[2] <= (uint32_t)kCutCap) {
                cb1.lo = lo; cb1.scale = scale; cb1.cb = sh[0]; cb1.r_in_bin = sh[1]; cb1.on = true;
            } else {
                failed = !select_boundary(each_l1, [&](uint32_t tok) { return __ldg(L_rows + tok) & 0x7fffffffu; }, n1 - (uint32_t)cap1, lo,
                                          hi, true, hist, sh, b_key, b_row, &s_bufn, &s_K, &s_R, b1);
            }
        }
        if (failed) {
            if (tid == 0) p.fallback[q] = 1;
            __syncthreads();
            continue;
        }
        const int64_t out1 = cut1 ? cap1 : (int64_t)n1;
        if (p.counts && tid == 0) {
            p.counts[((int64_t)0 * 3 + 0) * p.Q + q] = (int32_t)p.N;
            p.counts[((int64_t)0 * 3 + 1) * p.Q + q] = (int32_t)c0;
            p.counts[((int64_t)0 * 3 + 2) * p.Q + q] = (int32_t)c0;
            p.counts[((int64_t)1 * 3 + 0) * p.Q + q] = (int32_t)c0;
            p.counts[((int64_t)1 * 3 + 1) * p.Q + q] = (int32_t)n1;
            p.counts[((int64_t)1 * 3 + 2) * p.Q + q] = (int32_t)out1;
        }
        uint32_t* M = p.mask + (int64_t)q * p.mask_stride;
        __syncthreads();
        if (tid == 0) s_bufn = 0;
        __syncthreads();
        // 2 = survives cut 1, 1 = member of the cut bin (decided after the pass), 0 = dropped
        auto class1 = [&](float k, uint32_t row) -> int {
            if (!cut1) return 2;
            if (cb1.on) {
                const uint32_t b = lin_bin(k, cb1.lo, cb1.scale);
                return b > cb1.cb ? 2 : (b == cb1.cb ? 1 : 0);
            }
            return kept_by(b1, k, row) ? 2 : 0;
        };
        // ---- lazy path ----
        // Without a binding ratio cut the survivors are exactly the rows that pass every threshold: M = P0 & P1 & P2, no list
        // is read (random-like data: level 2 never binds, level 1 binds for about half of the queries).  When cut 1 binds
        // but drops only a small part of the list, the dropped entries (below the cut bin, or the low ranks inside it) are
        // found by ONE more pass over k1 alone; their rows are cleared from the AND of the planes, and level 2 cannot bind
        // as long as (rows passing all thresholds) - (dropped rows that pass level 2) <= cap2.  Everything else (a deep cut
        // 1, a binding cut 2) takes the streaming path below.  It replaces, for these queries, a pass over all three list
        // arrays, the compaction of the level-2 candidates and one atomicOr per survivor (~45 K per query at 1 M rows).
        {
            const uint32_t drop1 = cut1 ? n1 - (uint32_t)cap1 : 0u;
            const float lo2l = p.L > 2 ? __ldg(p.tq + (int64_t)2 * p.Q + q) : 0.f;
            bool lazy = !cut1 || drop1 <= n1 / 8u;
            if (lazy && cut1) {
                if (tid == 0) { s_nd = 0; s_d2 = 0; }
                __syncthreads();
                auto dropped = [&](uint32_t row, float k2) {
                    c_row[atomicAdd(&s_nd, 1u)] = row;
                    if (p.L > 2 && k2 >= lo2l) atomicAdd(&s_d2, 1u);
                };
                each_l1([&](float k, uint32_t tok) {
                    int cls;
                    if (cb1.on) {
                        const uint32_t b = lin_bin(k, cb1.lo, cb1.scale);
                        cls = b > cb1.cb ? 2 : (b == cb1.cb ? 1 : 0);
                    } else {
                        cls = k > b1.K ? 2 : (k < b1.K ? 0 : 3);          // 3: ties with the boundary key, decided by the row id
                    }
                    if (cls == 2) return;
                    const uint32_t row = __ldg(L_rows + tok) & 0x7fffffffu;
                    const float k2 = L_k2 ? __ldg(L_k2 + tok) : 0.f;
                    if (cls == 3) cls = row < b1.R ? 2 : 0;
                    if (cls == 2) return;
                    if (cls == 1) {
                        const uint32_t slot = atomicAdd(&s_bufn, 1u);
                        b_key[slot] = k; b_row[slot] = row; b_k2[slot] = k2;
                    } else {
                        dropped(row, k2);
                    }
                });
                __syncthreads();
                if (cb1.on) {
                    const uint32_t nb = s_bufn;
                    rank_cut_bin(b_key, b_row, nb, cb1.r_in_bin, &s_K, &s_R);
                    const Boundary bb{s_K, s_R};
                    for (uint32_t e = tid; e < nb; e += nt)
                        if (!kept_by(bb, b_key[e], b_row[e])) dropped(b_row[e], b_k2[e]);
                    __syncthreads();
                }
                lazy = s_nd == drop1;                                      // (always; a mismatch would mean an inconsistent list)
            }
            int64_t n2l = 0, cap2l = 0;
            if (lazy && p.L > 2) {
                n2l = (int64_t)c_all - (cut1 ? (int64_t)s_d2 : 0);
                cap2l = (int64_t)((double)out1 * p.ratio[2]);
                if (cap2l < 1) cap2l = 1;
                lazy = n2l <= cap2l;
            }
            if (lazy) {
                for (int64_t g = tid; g < groups; g += nt) {
                    uint32_t v0[4], va[4];
                    and_group(g, v0, va);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (g * 4 + i < p.words) M[g * 4 + i] = va[i];
                }
                __syncthreads();
                if (cut1) {
                    const uint32_t nd = s_nd;
                    for (uint32_t e = tid; e < nd; e += nt) {
                        const uint32_t row = __ldcg(c_row + e);
                        atomicAnd(&M[row >> 5], ~(1u << (row & 31)));
                    }
                }
                if (tid == 0) {
                    if (p.counts && p.L > 2) {
                        p.counts[((int64_t)2 * 3 + 0) * p.Q + q] = (int32_t)out1;
                        p.counts[((int64_t)2 * 3 + 1) * p.Q + q] = (int32_t)n2l;
                        p.counts[((int64_t)2 * 3 + 2) * p.Q + q] = (int32_t)n2l;
                    }
                    p.n_out[q] = (int32_t)(p.L > 2 ? n2l : out1);
                }
                __syncthreads();
                continue;
            }
            __syncthreads();
            if (tid == 0) s_bufn = 0;
            __syncthreads();
        }
        if (p.L == 2) {
            for (uint32_t item = warp; item < n_items; item += nw) {
                const uint32_t seg = item / cps, e_base = (item - seg * cps) * kChunk;
                const uint32_t cnt = (uint32_t)s_cnt[seg];
                if (e_base >= cnt) continue;
                const uint32_t off = seg * (uint32_t)p.seg_cap;
                float4 k[kU];
                uint4 rw[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t e = e_base + u * 128 + lane * 4;
                    const bool ok = e < cnt;
                    k[u] = ok ? __ldcs(reinterpret_cast<const float4*>(L_k1 + off + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    rw[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(L_rows + off + e)) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t e = e_base + u * 128 + lane * 4;
                    const float kk[4] = {k[u].x, k[u].y, k[u].z, k[u].w};
                    const uint32_t rr[4] = {rw[u].x, rw[u].y, rw[u].z, rw[u].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t row = rr[i] & 0x7fffffffu;
                        const int cls = e + i < cnt ? class1(kk[i], row) : 0;
                        if (cls == 2) atomicOr(&M[row >> 5], 1u << (row & 31));
                        if (cls == 1) {
                            const uint32_t slot = atomicAdd(&s_bufn, 1u);
                            b_key[slot] = kk[i]; b_row[slot] = row;
                        }
                    }
                }
            }
            __syncthreads();
            if (cb1.on) {
                const uint32_t nb = s_bufn;
                rank_cut_bin(b_key, b_row, nb, cb1.r_in_bin, &s_K, &s_R);
                const Boundary bb{s_K, s_R};
                for (uint32_t e = tid; e < nb; e += nt)
                    if (kept_by(bb, b_key[e], b_row[e])) atomicOr(&M[b_row[e] >> 5], 1u << (b_row[e] & 31));
            }
            if (tid == 0) p.n_out[q] = (int32_t)out1;
            __syncthreads();
            continue;
        }

        // ---- level 2: survivors of cut 1 that pass the level-2 threshold are compacted; histogram of k2 on the fly ----
        const float lo2 = __ldg(p.tq + (int64_t)2 * p.Q + q), hi2 = __ldg(p.nq + (int64_t)2 * p.Q + q);
        const float scale2 = hi2 > lo2 ? 2048.0f / (hi2 - lo2) : 0.f;
        for (int i = tid; i < 2048; i += nt) hist[i] = 0;
        __syncthreads();
        for (uint32_t item = warp; item < n_items2; item += nw) {       // chunks of kU2 x 128 entries (three arrays are read)
            const uint32_t seg = item / cps2, e_base = (item - seg * cps2) * kChunk2;
            const uint32_t cnt = (uint32_t)s_cnt[seg];
            if (e_base >= cnt) continue;
            const uint32_t off = seg * (uint32_t)p.seg_cap;
            float4 k1[kU2], k2[kU2];
            uint4 rw[kU2];
#pragma unroll
            for (int u = 0; u < kU2; ++u) {
                const uint32_t e = e_base + u * 128 + lane * 4;
                const bool ok = e < cnt;
                rw[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(L_rows + off + e)) : make_uint4(0u, 0u, 0u, 0u);
                k1[u] = ok ? __ldcs(reinterpret_cast<const float4*>(L_k1 + off + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                k2[u] = ok ? __ldcs(reinterpret_cast<const float4*>(L_k2 + off + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // one shared-memory reservation per batch of kU2 x 128 candidates
            uint32_t bal[kU2 * 4], total = 0;
            bool take[kU2 * 4];
            float kk2[kU2 * 4];
            uint32_t rr[kU2 * 4];
#pragma unroll
            for (int u = 0; u < kU2; ++u) {
                const uint32_t e = e_base + u * 128 + lane * 4;
                const float a1[4] = {k1[u].x, k1[u].y, k1[u].z, k1[u].w};
                const float a2[4] = {k2[u].x, k2[u].y, k2[u].z, k2[u].w};
                const uint32_t ar[4] = {rw[u].x, rw[u].y, rw[u].z, rw[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int s = u * 4 + i;
                    kk2[s] = a2[i];
                    rr[s] = ar[i] & 0x7fffffffu;
                    // k2 < tq fails the level-2 threshold (the same fp32 comparison as the threshold pass; false for a
                    // NaN threshold); entries past the end of the segment are stale data
                    const int cls = e + i < cnt ? class1(a1[i], rr[s]) : 0;
                    if (cls == 1) {              // every member of the cut bin: cut 1 ranks them before level 2 looks at them
                        const uint32_t slot = atomicAdd(&s_bufn, 1u);
                        b_key[slot] = a1[i]; b_row[slot] = rr[s]; b_k2[slot] = a2[i];
                    }
                    take[s] = cls == 2 && a2[i] >= lo2;
                    bal[s] = __ballot_sync(0xffffffffu, take[s]);
                    total += (uint32_t)__popc(bal[s]);
                }
            }
            uint32_t base = 0;
            if (lane == 0 && total) base = atomicAdd(&s_n2, total);
            base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
            for (int s = 0; s < kU2 * 4; ++s) {
                if (take[s]) {
                    const uint32_t slot = base + (uint32_t)__popc(bal[s] & ((1u << lane) - 1u));
                    c_k2[slot] = kk2[s];
                    c_row[slot] = rr[s];
                    atomicAdd(&hist[lin_bin(kk2[s], lo2, scale2)], 1u);
                }
                base += (uint32_t)__popc(bal[s]);
            }
        }
        __syncthreads();
        if (cb1.on) {
            // cut 1 inside its cut bin (exact ranking of all members), then the level-2 threshold for the kept ones
            const uint32_t nb = s_bufn;
            rank_cut_bin(b_key, b_row, nb, cb1.r_in_bin, &s_K, &s_R);
            const Boundary bb{s_K, s_R};
            for (uint32_t e = tid; e < nb; e += nt) {
                if (kept_by(bb, b_key[e], b_row[e]) && b_k2[e] >= lo2) {
                    const uint32_t slot = atomicAdd(&s_n2, 1u);
                    c_k2[slot] = b_k2[e];
                    c_row[slot] = b_row[e];
                    atomicAdd(&hist[lin_bin(b_k2[e], lo2, scale2)], 1u);
                }
            }
            __syncthreads();
        }
        const uint32_t n2 = s_n2;
        int64_t cap2 = (int64_t)((double)out1 * p.ratio[2]);
        if (cap2 < 1) cap2 = 1;
        const bool cut2 = (int64_t)n2 > cap2;
        Boundary b2;
        b2.K = -INFINITY; b2.R = 0;
        CutBin cb2{0.f, 0.f, 0u, 0u, false};
        if (cut2) {
            find_cut_bin(hist, n2 - (uint32_t)cap2, sh);          // the histogram accumulated while compacting
            if (sh[2] <= (uint32_t)kCutCap) {
                cb2.lo = lo2; cb2.scale = scale2; cb2.cb = sh[0]; cb2.r_in_bin = sh[1]; cb2.on = true;
            } else {
                auto each_l2 = [&](auto visit) {
                    for (uint32_t e0 = tid; e0 < n2; e0 += kU * nt) {
                        float k[kU];
#pragma unroll
                        for (int u = 0; u < kU; ++u) k[u] = e0 + u * nt < n2 ? __ldcg(c_k2 + e0 + u * nt) : -1.0f;
#pragma unroll
                        for (int u = 0; u < kU; ++u)
                            if (e0 + u * nt < n2) visit(k[u], e0 + u * nt);
                    }
                };
                failed = !select_boundary(each_l2, [&](uint32_t tok) { return __ldcg(c_row + tok); }, n2 - (uint32_t)cap2, lo2, hi2, true,
                                          hist, sh, b_key, b_row, &s_bufn, &s_K, &s_R, b2);
            }
        }
        if (failed) {
            if (tid == 0) p.fallback[q] = 1;
            __syncthreads();
            continue;
        }
        const int64_t out2 = cut2 ? cap2 : (int64_t)n2;
        if (p.counts && tid == 0) {
            p.counts[((int64_t)2 * 3 + 0) * p.Q + q] = (int32_t)out1;
            p.counts[((int64_t)2 * 3 + 1) * p.Q + q] = (int32_t)n2;
            p.counts[((int64_t)2 * 3 + 2) * p.Q + q] = (int32_t)out2;
        }
        __syncthreads();
        if (tid == 0) s_bufn = 0;
        __syncthreads();
        for (uint32_t e0 = tid; e0 < n2; e0 += kU * nt) {
            float k[kU];
            uint32_t row[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const bool ok = e0 + u * nt < n2;
                k[u] = ok ? __ldcg(c_k2 + e0 + u * nt) : -INFINITY;
                row[u] = ok ? __ldcg(c_row + e0 + u * nt) : 0u;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (e0 + u * nt >= n2) continue;
                int cls = 2;
                if (cut2) {
                    if (cb2.on) {
                        const uint32_t b = lin_bin(k[u], cb2.lo, cb2.scale);
                        cls = b > cb2.cb ? 2 : (b == cb2.cb ? 1 : 0);
                    } else {
                        cls = kept_by(b2, k[u], row[u]) ? 2 : 0;
                    }
                }
                if (cls == 2) atomicOr(&M[row[u] >> 5], 1u << (row[u] & 31));
                if (cls == 1) {
                    const uint32_t slot = atomicAdd(&s_bufn, 1u);
                    b_key[slot] = k[u]; b_row[slot] = row[u];
                }
            }
        }
        __syncthreads();
        if (cb2.on) {
            const uint32_t nb = s_bufn;
            rank_cut_bin(b_key, b_row, nb, cb2.r_in_bin, &s_K, &s_R);
            const Boundary bb{s_K, s_R};
            for (uint32_t e = tid; e < nb; e += nt)
                if (kept_by(bb, b_key[e], b_row[e])) atomicOr(&M[b_row[e] >> 5], 1u << (b_row[e] & 31));
        }
        if (tid == 0) p.n_out[q] = (int32_t)out2;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Window mode.  On most data the ratio cuts bind far above the thresholds (the level-1 and level-2 index rows are block
// means of the level-0 row, so a row that passes level 0 almost always passes the other thresholds too: cut 1 keeps the
// best 50 % of the level-0 survivors, cut 2 the best 70 % of those), and 13 % of all (query, row) pairs went through the
// candidate lists only for the cascade to find two cut keys per query.  Instead:
//   1. a SAMPLE pass (the same tensor-core kernel over every s-th 64-row tile) writes the lists of a sample of the rows;
//   2. k_filter_predict simulates the cascade on the sample and brackets each cut key by a window [lo, hi) that contains
//      the true cut with ~5 sigma of the sampling error (distribution free: the error is binomial in rank space);
//   3. the WINDOW pass (k_filter_bits_tc<true>) classifies every pair against the windows with bit-mask arithmetic, writes
//      the plane of the rows that survive both cuts for sure, four counters per query, and lists only the window rows;
//   4. k_filter_cascade_win ranks the window rows exactly (same selection code as the streaming cascade) and checks the
//      prediction against the exact counters.  A query whose cut lies outside its window (or whose lists overflow, or
//      whose level-0 cut binds) is flagged and redone by the fallback: full-threshold planes for its query tile +
//      the generic gather cascade.
// ---------------------------------------------------------------------------------------
struct PredictParams {
    int Q, L;
    double ratio[3];
    float z;                     // window half width in standard deviations of the sample rank
    const float* tq;             // [3][Q]
    const float* nq;             // [3][Q]
    const float* l_k1;           // sample lists
    const float* l_k2;
    const int32_t* seg_n;
    int64_t seg_cap;
    int n_segs;
    const int32_t* c0_s;         // [Q] rows of the sample passing level 0
    float* win;                  // [4][Q] out
    int32_t* pflag;              // [Q] out: 1 = no prediction (sample lists overflowed)
};

// suffix[b] = sum of hist[b..2047] (256 threads, eight bins each)
__device__ __forceinline__ void suffix_sums_256(const uint32_t* hist, uint32_t* suffix, uint32_t* s_scan) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = hist[2047 - (tid * 8 + i)]; sum += v[i]; }      // reversed: thread 0 owns the top bins
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += s_scan[w];
    uint32_t run = base + incl - sum;
#pragma unroll
    for (int i = 0; i < 8; ++i) { run += v[i]; suffix[2047 - (tid * 8 + i)] = run; }
    __syncthreads();
}

// Window around the rank `r` (from the top) of a key distribution given as suffix sums over 2048 linear bins of
// [lo, hi): out_hi = lower edge of the lowest bin b with suffix[b] <= r - m (+inf if none), out_lo = lower edge of the
// highest bin b with suffix[b] >= r + m (lo if none).  sh_f[0] / sh_f[1] receive them.
__device__ __forceinline__ void window_from_suffix(const uint32_t* suffix, float r, float m, float lo, float hi, float* sh_f) {
    const int tid = threadIdx.x;
    const float w = (hi - lo) * (1.0f / 2048.0f);
    if (tid == 0) { sh_f[0] = INFINITY; sh_f[1] = lo; }
    __syncthreads();
    const float r_hi = r - m, r_lo = r + m;
    for (int b = tid; b < 2048; b += blockDim.x) {
        const float sb = (float)suffix[b];
        if (sb <= r_hi && (b == 0 || (float)suffix[b - 1] > r_hi)) sh_f[0] = lo + w * (float)b;
        if (sb >= r_lo && (b == 2047 || (float)suffix[b + 1] < r_lo)) sh_f[1] = lo + w * (float)b;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 4) k_filter_predict(const PredictParams p) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t suffix[2048];
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_warp[32];
    __shared__ float sh_f[2];
    __shared__ uint32_t s_cross;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int q = blockIdx.x;
    const float qnan = __int_as_float(0x7fc00000);
    uint32_t part = 0, over = 0;
    for (int i = tid; i < p.n_segs; i += nt) {
        const int32_t n = p.seg_n[(int64_t)q * p.n_segs + i];
        if (n > p.seg_cap) over = 1;
        part += (uint32_t)(n > p.seg_cap ? (int32_t)p.seg_cap : n);
    }
    const uint32_t n1 = block_sum(part, s_warp);
    const uint32_t any_over = block_sum(over, s_warp);
    const bool three = p.L > 2;
    const float t1 = __ldg(p.tq + (int64_t)1 * p.Q + q), t2 = three ? __ldg(p.tq + (int64_t)2 * p.Q + q) : -INFINITY;
    const float u1 = __ldg(p.nq + (int64_t)1 * p.Q + q), u2 = three ? __ldg(p.nq + (int64_t)2 * p.Q + q) : 0.f;
    const float t0 = __ldg(p.tq + q);
    if (any_over || !(t0 == t0) || !(t1 == t1) || !(t2 == t2)) {
        // NaN thresholds (a zero query level: nothing passes) give NaN windows = an empty result without a fallback
        if (tid < 4) p.win[(int64_t)tid * p.Q + q] = qnan;
        if (tid == 0) p.pflag[q] = any_over ? 1 : 0;
        return;
    }
    const float c0s = (float)p.c0_s[q];
    const int64_t qbase = (int64_t)q * p.n_segs * p.seg_cap;
    // the sample entries as 4-entry groups numbered across all segments (independent loads, in flight together: a warp per
    // segment was a chain of 30 DRAM round trips per pass for a single query's 262 segments)
    __shared__ int32_t s_pcnt[kMaxSegs];
    __shared__ int32_t s_pgoff[kMaxSegs + 1];
    const int lane = tid & 31;
    for (int i = tid; i < p.n_segs; i += nt) {
        const int32_t n = p.seg_n[(int64_t)q * p.n_segs + i];
        s_pcnt[i] = n > p.seg_cap ? (int32_t)p.seg_cap : n;
    }
    __syncthreads();
    if (tid < 32) {
        const int per = (p.n_segs + 31) / 32;
        int sum = 0;
        for (int i = 0; i < per; ++i) { const int sg = lane * per + i; if (sg < p.n_segs) sum += (s_pcnt[sg] + 3) >> 2; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - sum;
        for (int i = 0; i < per; ++i) {
            const int sg = lane * per + i;
            if (sg < p.n_segs) { s_pgoff[sg] = run; run += (s_pcnt[sg] + 3) >> 2; }
        }
        if (lane == 31) s_pgoff[p.n_segs] = incl;
    }
    __syncthreads();
    const int n_groups = s_pgoff[p.n_segs];
    auto each = [&](auto visit) {
        for (int g0 = tid; g0 < n_groups; g0 += 2 * nt) {
            float4 a[2], b[2];
            int n[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int g = g0 + u * nt;
                n[u] = 0;
                b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g < n_groups) {
                    int lo = 0, hi = p.n_segs;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (s_pgoff[mid] <= g) lo = mid; else hi = mid;
                    }
                    const int e = (g - s_pgoff[lo]) * 4;
                    const int64_t off = qbase + (int64_t)lo * p.seg_cap + e;
                    n[u] = min(4, s_pcnt[lo] - e);
                    a[u] = __ldg(reinterpret_cast<const float4*>(p.l_k1 + off));
                    if (three) b[u] = __ldg(reinterpret_cast<const float4*>(p.l_k2 + off));
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (n[u] > 0) visit(a[u].x, b[u].x);
                if (n[u] > 1) visit(a[u].y, b[u].y);
                if (n[u] > 2) visit(a[u].z, b[u].z);
                if (n[u] > 3) visit(a[u].w, b[u].w);
            }
        }
    };
    // ---- level 1 ----
    const float r1 = (float)p.ratio[1], r12 = (float)(p.ratio[1] * p.ratio[2]);
    const float cap1 = r1 * c0s;
    const float m1 = p.z * sqrtf(fmaxf(r1 * (1.0f - r1), 0.01f) * fmaxf(c0s, 1.0f)) + 4.0f;
    const float sc1 = u1 > t1 ? 2048.0f / (u1 - t1) : 0.f;
    for (int i = tid; i < 2048; i += nt) hist[i] = 0;
    __syncthreads();
    each([&](float k1, float) { atomicAdd(&hist[lin_bin(k1, t1, sc1)], 1u); });
    __syncthreads();
    suffix_sums_256(hist, suffix, s_scan);
    float lo1 = t1, hi1 = t1;
    const bool bind1 = (float)n1 + m1 > cap1;
    uint32_t cross_bin = 0;                            // entries in bins >= cross_bin play the survivors of cut 1 below
    if (bind1) {
        window_from_suffix(suffix, cap1, m1, t1, u1, sh_f);
        hi1 = sh_f[0]; lo1 = sh_f[1];
        if (tid == 0) s_cross = 0;
        __syncthreads();
        for (int b = tid; b < 2048; b += nt)
            if ((float)suffix[b] >= cap1 && (b == 2047 || (float)suffix[b + 1] < cap1)) s_cross = (uint32_t)b;
        __syncthreads();
        cross_bin = s_cross;
    }
    const float pop_cross = bind1 ? (float)hist[cross_bin] : 0.f;
    __syncthreads();
    if (!three) {                                      // two levels: no cut 2 (the window cascade treats level 2 as "everything passes")
        if (tid == 0) {
            p.win[(int64_t)0 * p.Q + q] = lo1; p.win[(int64_t)1 * p.Q + q] = hi1;
            p.win[(int64_t)2 * p.Q + q] = -INFINITY; p.win[(int64_t)3 * p.Q + q] = -INFINITY;
            p.pflag[q] = 0;
        }
        return;
    }
    // ---- level 2 on the sample's survivors of cut 1 ----
    const float out1 = fminf((float)n1, cap1);
    const float cap2 = (float)p.ratio[2] * out1;
    const float m2 = p.z * sqrtf(fmaxf(r12 * (1.0f - r12), 0.01f) * fmaxf(c0s, 1.0f)) + pop_cross + 4.0f;
    const float sc2 = u2 > t2 ? 2048.0f / (u2 - t2) : 0.f;
    for (int i = tid; i < 2048; i += nt) hist[i] = 0;
    __syncthreads();
    uint32_t c2 = 0;
    each([&](float k1, float k2) {
        if (lin_bin(k1, t1, sc1) >= cross_bin && k2 >= t2) { atomicAdd(&hist[lin_bin(k2, t2, sc2)], 1u); ++c2; }
    });
    const uint32_t n2 = block_sum(c2, s_warp);
    suffix_sums_256(hist, suffix, s_scan);
    float lo2 = t2, hi2 = t2;
    if ((float)n2 + m2 > cap2) {
        window_from_suffix(suffix, cap2, m2, t2, u2, sh_f);
        hi2 = sh_f[0]; lo2 = sh_f[1];
    }
    if (tid == 0) {
        p.win[(int64_t)0 * p.Q + q] = lo1; p.win[(int64_t)1 * p.Q + q] = hi1;
        p.win[(int64_t)2 * p.Q + q] = lo2; p.win[(int64_t)3 * p.Q + q] = hi2;
        p.pflag[q] = 0;
    }
}

struct WinParams {
    const uint32_t* alive;       // [Q][pitch] rows that survive both cuts for sure (window pass)
    int64_t words, pitch;
    int64_t N;
    int Q, L;
    double ratio[3];
    const float* tq;             // [3][Q]
    const float* nq;             // [3][Q]
    const float* win;            // [4][Q]
    const int32_t* wcnt;         // [4][Q]
    const int32_t* pflag;        // [Q]
    const uint32_t* l_rows;
    const float* l_k1;
    const float* l_k2;
    const int32_t* seg_n;
    int64_t seg_cap;
    int n_segs;
    uint32_t* mask;
    int64_t mask_stride;
    int32_t* n_out;
    int32_t* fallback;           // [Q] out
    int32_t* tile_flag;          // [ceil(Q / 128)] out: a query of the tile fell back
    float* tmp_keys;             // [gridDim][tmp_stride]
    uint32_t* tmp_rows;
    int64_t tmp_stride;
};

constexpr int kWinThreads = 128;      // eight CTAs per SM: 1184 queries in flight, a 1024-query batch is ONE wave
constexpr int kWinCutCap = 1024;      // cut-bin members ranked in shared memory (window histograms are fine grained)
constexpr int kWinCtasPerSm = 8;      // 64 registers per thread (256-thread CTAs: 4 per SM = 592 queries in flight, two waves)

template <int THREADS, int CTAS, int U>          // U: 4-entry groups a thread keeps in flight in the list passes
__global__ void __launch_bounds__(THREADS, CTAS) k_filter_cascade_win(const WinParams p) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t sh[4];
    __shared__ uint32_t s_warp[32];
    __shared__ int32_t s_cnt[kMaxSegs];
    __shared__ float b_key[kWinCutCap];
    __shared__ uint32_t b_row[kWinCutCap];
    __shared__ float b_k2[kWinCutCap];
    __shared__ uint32_t s_bufn, s_flag, s_R, s_n2, s_x1, s_x2;
    __shared__ float s_K;
    __shared__ int32_t s_goff[kMaxSegs + 1];          // exclusive prefix of the segments' 4-entry groups
    const int tid = threadIdx.x, nt = blockDim.x;
    float* const c_k2 = p.tmp_keys + (int64_t)blockIdx.x * p.tmp_stride;
    uint32_t* const c_row = p.tmp_rows + (int64_t)blockIdx.x * p.tmp_stride;

    for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
        if (tid == 0) { s_flag = 0; s_n2 = 0; s_bufn = 0; s_x1 = 0; s_x2 = 0; }
        __syncthreads();
        for (int i = tid; i < p.n_segs; i += nt) {
            const int32_t n = p.seg_n[(int64_t)q * p.n_segs + i];
            if (n > p.seg_cap) s_flag = 1;
            s_cnt[i] = n;
        }
        if (q + 1 < p.Q)                                 // padding words of the mask row (the last row ends at `words`)
            for (int64_t w = p.words + tid; w < p.mask_stride; w += nt) p.mask[(int64_t)q * p.mask_stride + w] = 0u;
        const int64_t c0 = __ldg(p.wcnt + q), nA1 = __ldg(p.wcnt + (int64_t)1 * p.Q + q);
        const int64_t nA1A2 = __ldg(p.wcnt + (int64_t)2 * p.Q + q), nAl = __ldg(p.wcnt + (int64_t)3 * p.Q + q);
        int64_t cap0 = (int64_t)((double)p.N * p.ratio[0]);
        if (cap0 < 1) cap0 = 1;
        __syncthreads();
        bool failed = s_flag || __ldg(p.pflag + q) || c0 > cap0;
        auto fail = [&]() {
            if (tid == 0) { p.fallback[q] = 1; p.tile_flag[q >> 7] = 1; }
            __syncthreads();
        };
        if (failed) { fail(); continue; }
        const bool three = p.L > 2;
        const float lo1 = __ldg(p.win + q), hi1 = __ldg(p.win + (int64_t)1 * p.Q + q);
        const float lo2 = __ldg(p.win + (int64_t)2 * p.Q + q), hi2 = __ldg(p.win + (int64_t)3 * p.Q + q);
        const float t1 = __ldg(p.tq + (int64_t)1 * p.Q + q), t2 = three ? __ldg(p.tq + (int64_t)2 * p.Q + q) : -INFINITY;
        uint32_t* M = p.mask + (int64_t)q * p.mask_stride;
        const uint32_t* AL = p.alive + (int64_t)q * p.pitch;
        if (AL != M) {   // M = alive plane (plane rows are 32-byte aligned and padded to 8 words; four 128-bit loads in flight per thread)
            const int64_t groups = (p.words + 3) >> 2;
            for (int64_t g0 = tid; g0 < groups; g0 += 4 * (int64_t)nt) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t g = g0 + (int64_t)u * nt;
                    v[u] = g < groups ? __ldg(reinterpret_cast<const uint4*>(AL) + g) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t w = (g0 + (int64_t)u * nt) * 4;
                    if (w < p.words) M[w] = v[u].x;
                    if (w + 1 < p.words) M[w + 1] = v[u].y;
                    if (w + 2 < p.words) M[w + 2] = v[u].z;
                    if (w + 3 < p.words) M[w + 3] = v[u].w;
                }
            }
        }
        if (!(lo1 == lo1)) {                            // NaN windows: a zero query level, nothing survives
            if (tid == 0) p.n_out[q] = 0;
            __syncthreads();
            continue;
        }
        const int64_t qbase = (int64_t)q * p.n_segs * p.seg_cap;
        const uint32_t* const L_rows = p.l_rows + qbase;
        const float* const L_k1 = p.l_k1 + qbase;
        const float* const L_k2 = three ? p.l_k2 + qbase : nullptr;
        // every listed entry of the query: visit(k1, token).  The entries are walked as 4-entry groups (segments are
        // 128-byte aligned) numbered across ALL segments, so that a thread's loads are independent and in flight together:
        // a loop over one short segment after the other (per block, then per warp) was a chain of 20-40 DRAM round trips
        // per pass -- the whole kernel at small shards.
        const int lane = tid & 31;
        if (tid < 32) {
            const int per = (p.n_segs + 31) / 32;
            int sum = 0;
            for (int i = 0; i < per; ++i) { const int sg = lane * per + i; if (sg < p.n_segs) sum += (s_cnt[sg] + 3) >> 2; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int run = incl - sum;
            for (int i = 0; i < per; ++i) {
                const int sg = lane * per + i;
                if (sg < p.n_segs) { s_goff[sg] = run; run += (s_cnt[sg] + 3) >> 2; }
            }
            if (lane == 31) s_goff[p.n_segs] = incl;
        }
        __syncthreads();
        const int n_groups = s_goff[p.n_segs];
        // A warp owns a contiguous range of groups, its lanes take consecutive groups (coalesced 512-byte reads); a lane's
        // groups ascend, so the segment of a group is found by walking forward from the lane's previous segment (a binary
        // search per group made the passes instruction bound on long lists).
        const int warp = tid >> 5, nw = nt >> 5;
        const int g_per_warp = (n_groups + nw - 1) / nw;
        const int g_begin = warp * g_per_warp, g_end = min(n_groups, g_begin + g_per_warp);
        int seg0 = 0;
        {
            int lo = 0, hi = p.n_segs;                               // last segment whose first group is <= g_begin
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_goff[mid] <= g_begin) lo = mid; else hi = mid;
            }
            seg0 = lo;
        }
        auto locate = [&](int g, int& seg, uint32_t& off, uint32_t& n) {
            while (s_goff[seg + 1] <= g) ++seg;
            const uint32_t e = (uint32_t)(g - s_goff[seg]) * 4u;
            off = (uint32_t)seg * (uint32_t)p.seg_cap + e;
            n = min(4u, (uint32_t)s_cnt[seg] - e);
        };
        auto each_all = [&](auto visit) {
            int seg = seg0;
            for (int g0 = g_begin + lane; g0 < g_end; g0 += U * 32) {
                float4 k[U];
                uint32_t off[U], n[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int g = g0 + u * 32;
                    n[u] = 0; off[u] = 0;
                    if (g < g_end) {
                        locate(g, seg, off[u], n[u]);
                        k[u] = __ldg(reinterpret_cast<const float4*>(L_k1 + off[u]));
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (n[u] > 0) visit(k[u].x, off[u]);
                    if (n[u] > 1) visit(k[u].y, off[u] + 1);
                    if (n[u] > 2) visit(k[u].z, off[u] + 2);
                    if (n[u] > 3) visit(k[u].w, off[u] + 3);
                }
            }
        };
        // the same with the rows and the level-2 keys loaded up front (the classification pass needs them for most entries)
        auto each_all3 = [&](auto visit) {
            int seg = seg0;
            for (int g0 = g_begin + lane; g0 < g_end; g0 += 2 * 32) {
                float4 k[2], k2[2];
                uint4 rw[2];
                uint32_t n[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int g = g0 + u * 32;
                    n[u] = 0;
                    k2[u] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
                    if (g < g_end) {
                        uint32_t off;
                        locate(g, seg, off, n[u]);
                        k[u] = __ldg(reinterpret_cast<const float4*>(L_k1 + off));
                        rw[u] = __ldg(reinterpret_cast<const uint4*>(L_rows + off));
                        if (three) k2[u] = __ldg(reinterpret_cast<const float4*>(L_k2 + off));
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (n[u] > 0) visit(k[u].x, rw[u].x & 0x7fffffffu, k2[u].x);
                    if (n[u] > 1) visit(k[u].y, rw[u].y & 0x7fffffffu, k2[u].y);
                    if (n[u] > 2) visit(k[u].z, rw[u].z & 0x7fffffffu, k2[u].z);
                    if (n[u] > 3) visit(k[u].w, rw[u].w & 0x7fffffffu, k2[u].w);
                }
            }
        };
        // Keys are relative to the lower window edges (the window pass folds -lo into the contraction): k1 in [0, w1) inside
        // window 1, >= w1 above it; k2 < 0 below window 2, >= w2 above it (w = hi - lo, the same float the pass tests against).
        const float w1 = hi1 - lo1, w2 = three ? hi2 - lo2 : -INFINITY;
        auto each_w1 = [&](auto visit) {
            each_all([&](float k, uint32_t tok) { if (k < w1) visit(k, tok); });
        };
        // ---- level 1: rows inside window 1, ranked below the nA1 rows above it ----
        const float top1 = (hi1 < INFINITY ? hi1 : __ldg(p.nq + (int64_t)1 * p.Q + q)) - lo1;       // relative upper edge
        const float scale1 = top1 > 0.f ? 2048.0f / top1 : 0.f;
        for (int i = tid; i < 2048; i += nt) hist[i] = 0;
        __syncthreads();
        uint32_t cw = 0;
        each_w1([&](float k, uint32_t) { atomicAdd(&hist[lin_bin(k, 0.f, scale1)], 1u); ++cw; });
        const int64_t n_w1 = block_sum(cw, s_warp);
        int64_t cap1 = (int64_t)((double)c0 * p.ratio[1]);
        if (cap1 < 1) cap1 = 1;
        const bool exact1 = lo1 == t1;                  // the window reaches down to the threshold: n1 is exact
        const int64_t n1 = nA1 + n_w1;
        if ((!exact1 && n1 < cap1) || (n1 > cap1 && nA1 > cap1)) { fail(); continue; }
        const bool cut1 = n1 > cap1;
        Boundary b1;
        b1.K = -INFINITY; b1.R = 0;
        CutBin cb1{0.f, 0.f, 0u, 0u, false};
        if (cut1) {
            find_cut_bin(hist, (uint32_t)(n1 - cap1), sh);
            if (sh[2] <= (uint32_t)kWinCutCap) {
                cb1.lo = 0.f; cb1.scale = scale1; cb1.cb = sh[0]; cb1.r_in_bin = sh[1]; cb1.on = true;
            } else {
                failed = !select_boundary(each_w1, [&](uint32_t tok) { return __ldg(L_rows + tok) & 0x7fffffffu; }, (uint32_t)(n1 - cap1),
                                          0.f, top1, true, hist, sh, b_key, b_row, &s_bufn, &s_K, &s_R, b1, (uint32_t)kWinCutCap);
            }
        }
        if (failed) { fail(); continue; }
        const int64_t out1 = cut1 ? cap1 : n1;
        int64_t cap2 = three ? (int64_t)((double)out1 * p.ratio[2]) : out1;
        if (cap2 < 1) cap2 = 1;
        __syncthreads();
        if (tid == 0) s_bufn = 0;
        const float top2 = three ? (hi2 < INFINITY ? hi2 : __ldg(p.nq + (int64_t)2 * p.Q + q)) - lo2 : 0.f;
        const float scale2 = top2 > 0.f ? 2048.0f / top2 : 0.f;
        for (int i = tid; i < 2048; i += nt) hist[i] = 0;
        __syncthreads();
        // a survivor of cut 1: dead below window 2, alive above it (window-1 rows are not in the plane yet), else a member
        auto level2 = [&](uint32_t row, float k2, bool inw1) {
            if (!(k2 >= 0.f)) return;
            if (inw1) atomicAdd(&s_x1, 1u);
            if (k2 >= w2) {
                if (inw1) { atomicOr(&M[row >> 5], 1u << (row & 31)); atomicAdd(&s_x2, 1u); }
                return;
            }
            const uint32_t slot = atomicAdd(&s_n2, 1u);
            c_k2[slot] = k2; c_row[slot] = row;
            atomicAdd(&hist[lin_bin(k2, 0.f, scale2)], 1u);
        };
        each_all3([&](float k1, uint32_t row, float k2) {          // (two levels: k2 = +inf, every survivor of cut 1 survives)
            const bool inw1 = k1 < w1;
            int cls = 2;
            if (inw1 && cut1) {
                if (cb1.on) {
                    const uint32_t b = lin_bin(k1, cb1.lo, cb1.scale);
                    cls = b > cb1.cb ? 2 : (b == cb1.cb ? 1 : 0);
                } else {
                    cls = kept_by(b1, k1, row) ? 2 : 0;
                }
            }
            if (cls == 0) return;
            if (cls == 1) {
                const uint32_t slot = atomicAdd(&s_bufn, 1u);
                b_key[slot] = k1; b_row[slot] = row; b_k2[slot] = k2;
                return;
            }
            level2(row, k2, inw1);
        });
        __syncthreads();
        if (cb1.on) {
            const uint32_t nb = s_bufn;
            rank_cut_bin(b_key, b_row, nb, cb1.r_in_bin, &s_K, &s_R);
            const Boundary bb{s_K, s_R};
            for (uint32_t e = tid; e < nb; e += nt)
                if (kept_by(bb, b_key[e], b_row[e])) level2(b_row[e], b_k2[e], true);
            __syncthreads();
        }
        const int64_t n_w2 = s_n2;
        const int64_t n2 = nA1A2 + (int64_t)s_x1, nAbove2 = nAl + (int64_t)s_x2;
        const bool exact2 = lo2 == t2;
        if (n2 - nAbove2 != n_w2 || (!exact2 && n2 < cap2) || (n2 > cap2 && nAbove2 > cap2)) { fail(); continue; }
        const bool cut2 = n2 > cap2;
        Boundary b2;
        b2.K = -INFINITY; b2.R = 0;
        CutBin cb2{0.f, 0.f, 0u, 0u, false};
        auto each_l2 = [&](auto visit) {
            for (uint32_t e0 = tid; e0 < (uint32_t)n_w2; e0 += 4 * nt) {
                float k[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) k[u] = e0 + u * nt < (uint32_t)n_w2 ? __ldcg(c_k2 + e0 + u * nt) : -1.0f;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (e0 + u * nt < (uint32_t)n_w2) visit(k[u], e0 + u * nt);
            }
        };
        if (cut2) {
            find_cut_bin(hist, (uint32_t)(n2 - cap2), sh);
            if (sh[2] <= (uint32_t)kWinCutCap) {
                cb2.lo = 0.f; cb2.scale = scale2; cb2.cb = sh[0]; cb2.r_in_bin = sh[1]; cb2.on = true;
            } else {
                failed = !select_boundary(each_l2, [&](uint32_t tok) { return __ldcg(c_row + tok); }, (uint32_t)(n2 - cap2), 0.f, top2, true,
                                          hist, sh, b_key, b_row, &s_bufn, &s_K, &s_R, b2, (uint32_t)kWinCutCap);
            }
        }
        if (failed) { fail(); continue; }
        __syncthreads();
        if (tid == 0) s_bufn = 0;
        __syncthreads();
        each_l2([&](float k, uint32_t tok) {
            const uint32_t row = __ldcg(c_row + tok);
            int cls = 2;
            if (cut2) {
                if (cb2.on) {
                    const uint32_t b = lin_bin(k, cb2.lo, cb2.scale);
                    cls = b > cb2.cb ? 2 : (b == cb2.cb ? 1 : 0);
                } else {
                    cls = kept_by(b2, k, row) ? 2 : 0;
                }
            }
            if (cls == 2) atomicOr(&M[row >> 5], 1u << (row & 31));
            if (cls == 1) {
                const uint32_t slot = atomicAdd(&s_bufn, 1u);
                b_key[slot] = k; b_row[slot] = row;
            }
        });
        __syncthreads();
        if (cb2.on) {
            const uint32_t nb = s_bufn;
            rank_cut_bin(b_key, b_row, nb, cb2.r_in_bin, &s_K, &s_R);
            const Boundary bb{s_K, s_R};
            for (uint32_t e = tid; e < nb; e += nt)
                if (kept_by(bb, b_key[e], b_row[e])) atomicOr(&M[b_row[e] >> 5], 1u << (b_row[e] & 31));
        }
        if (tid == 0) p.n_out[q] = (int32_t)(cut2 ? cap2 : n2);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Exceptional rows.  The fast path assumes every index row has its structural length; a block mean that
// happens to be exactly 0.0f at the end of a row shortens the row's stored length (one row in ~10 M at
// 768-D), and the reference then normalises the QUERY over that shorter prefix.  Such rows are masked out of
// the dense passes (validity words / overwritten plane bits) and scored here pair by pair with the exact
// path's arithmetic (k_filter_level); rows that pass levels 0 and 1 are appended to one extra list segment
// per query with score-equivalent keys ((2 s - 1) |q_l|), so the cascade ranks them with everybody else.
// ---------------------------------------------------------------------------------------
struct ExcParams {
    const float* idx;
    const uint16_t* lens;
    hq_index_layout lay;
    const float* q_idx;
    int Q;
    const int32_t* exc_rows;
    int n_exc;
    double thr[3];
    uint32_t* bits;
    int64_t bits_pitch;
    const float* nq;             // [3][Q] full query level norms
    const float* tq;             // [3][Q] thresholds in dot-product units (needed with lists)
    uint32_t* l_rows;            // lists (optional)
    float* l_k1;
    float* l_k2;
    int32_t* seg_n;
    int64_t seg_cap;
    int n_segs, extra_seg;
};

__global__ void __launch_bounds__(128) k_filter_exceptions(const ExcParams p) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)p.Q * p.n_exc) return;
    const int q = (int)(t / p.n_exc), e = (int)(t - (int64_t)q * p.n_exc);
    const int64_t row = p.exc_rows[e];
    const int L = p.lay.L;
    bool pass[3] = {false, false, false};
    float keq[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < L && l < 3; ++l) {
        const int keff = p.lay.lvl_keff[l];
        int m = (int)p.lens[row * L + l];
        m = m < keff ? m : keff;                                   // the query is dense: common prefix = row length
        const float* c = p.idx + row * p.lay.Lsum + p.lay.lvl_off[l];
        const float* qv = p.q_idx + (int64_t)q * p.lay.Lsum + p.lay.lvl_off[l];
        float dot = 0.f, nc2 = 0.f, nq2 = 0.f;
        for (int j = 0; j < keff; ++j) {
            const float cj = __ldg(c + j), qj = __ldg(qv + j);
            dot = fmaf(cj, qj, dot);
            nc2 = fmaf(cj, cj, nc2);
            if (j < m) nq2 = fmaf(qj, qj, nq2);
        }
        const float nq = sqrtf(nq2), nc = sqrtf(nc2);
        float sc = 0.f;
        if (nq != 0.f && nc != 0.f) sc = __fmul_rn(__fadd_rn(__fdiv_rn(dot, __fmul_rn(nq, nc)), 1.0f), 0.5f);
        pass[l] = (double)sc >= p.thr[l];
        keq[l] = (2.0f * sc - 1.0f) * __ldg(p.nq + (int64_t)l * p.Q + q);
        uint32_t* w = p.bits + ((int64_t)l * p.Q + q) * p.bits_pitch + (row >> 5);
        const uint32_t bit = 1u << (row & 31);
        atomicAnd(w, ~bit);
        if (pass[l]) atomicOr(w, bit);
    }
    if (p.l_rows && L >= 2 && pass[0] && pass[1]) {
        const int pos = atomicAdd(p.seg_n + (int64_t)q * p.n_segs + p.extra_seg, 1);
        if (pos < p.seg_cap) {
            const int64_t a = ((int64_t)q * p.n_segs + p.extra_seg) * p.seg_cap + pos;
            p.l_rows[a] = (uint32_t)row;
            p.l_k1[a] = keq[1];
            if (L > 2) {
                // the cascade repeats the level-2 test as k2 >= tq[2]: make the key say what the exact score said
                const float t2 = __ldg(p.tq + (int64_t)2 * p.Q + q);
                p.l_k2[a] = pass[2] ? fmaxf(keq[2], t2) : -INFINITY;
            }
        }
    }
}

// The same rows in window mode: classified against the query's windows like the window pass does (counters, alive
// plane, extra list segment).  Runs between the window pass and k_filter_cascade_win.
struct ExcWinParams {
    ExcParams e;
    const float* win;            // [4][Q]
    int32_t* wcnt;               // [4][Q]
    uint32_t* alive;             // [Q][alive_pitch]
    int64_t alive_pitch;
};

__global__ void __launch_bounds__(128) k_filter_exceptions_win(const ExcWinParams pw) {
    const ExcParams& p = pw.e;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)p.Q * p.n_exc) return;
    const int q = (int)(t / p.n_exc), e = (int)(t - (int64_t)q * p.n_exc);
    const int64_t row = p.exc_rows[e];
    const int L = p.lay.L;
    bool pass[3] = {false, false, false};
    float keq[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < L && l < 3; ++l) {
        const int keff = p.lay.lvl_keff[l];
        int m = (int)p.lens[row * L + l];
        m = m < keff ? m : keff;
        const float* c = p.idx + row * p.lay.Lsum + p.lay.lvl_off[l];
        const float* qv = p.q_idx + (int64_t)q * p.lay.Lsum + p.lay.lvl_off[l];
        float dot = 0.f, nc2 = 0.f, nq2 = 0.f;
        for (int j = 0; j < keff; ++j) {
            const float cj = __ldg(c + j), qj = __ldg(qv + j);
            dot = fmaf(cj, qj, dot);
            nc2 = fmaf(cj, cj, nc2);
            if (j < m) nq2 = fmaf(qj, qj, nq2);
        }
        const float nq = sqrtf(nq2), nc = sqrtf(nc2);
        float sc = 0.f;
        if (nq != 0.f && nc != 0.f) sc = __fmul_rn(__fadd_rn(__fdiv_rn(dot, __fmul_rn(nq, nc)), 1.0f), 0.5f);
        pass[l] = (double)sc >= p.thr[l];
        keq[l] = (2.0f * sc - 1.0f) * __ldg(p.nq + (int64_t)l * p.Q + q);
    }
    const float lo1 = __ldg(pw.win + q), hi1 = __ldg(pw.win + (int64_t)1 * p.Q + q);
    const float lo2 = __ldg(pw.win + (int64_t)2 * p.Q + q), hi2 = __ldg(pw.win + (int64_t)3 * p.Q + q);
    if (!(lo1 == lo1) || !pass[0]) return;
    const float t1 = __ldg(p.tq + (int64_t)1 * p.Q + q), t2 = L > 2 ? __ldg(p.tq + (int64_t)2 * p.Q + q) : 0.f;
    // keys that say what the exact scores said: a row that passes a threshold never sorts below it
    const float k1 = pass[1] ? fmaxf(keq[1], t1) : -INFINITY, k2 = pass[2] ? fmaxf(keq[2], t2) : -INFINITY;
    const bool A1 = k1 >= lo1, A2 = L < 3 || k2 >= lo2;
    // stored like the window pass stores them: relative to the lower window edge; the classes are then taken FROM the stored
    // keys, exactly as the cascade will repeat them
    const float k1s = A1 ? fmaxf(k1 - lo1, 0.f) : -INFINITY;
    const float k2s = L < 3 ? INFINITY : (A2 ? fmaxf(k2 - lo2, 0.f) : -INFINITY);
    const float w1 = hi1 - lo1, w2 = hi2 - lo2;
    const bool B1 = A1 && k1s >= w1, B2 = L < 3 || (A2 && k2s >= w2);
    atomicAdd(pw.wcnt + q, 1);
    if (B1) atomicAdd(pw.wcnt + (int64_t)1 * p.Q + q, 1);
    if (B1 && A2) atomicAdd(pw.wcnt + (int64_t)2 * p.Q + q, 1);
    if (B1 && B2) {
        atomicAdd(pw.wcnt + (int64_t)3 * p.Q + q, 1);
        atomicOr(pw.alive + (int64_t)q * pw.alive_pitch + (row >> 5), 1u << (row & 31));
    }
    if (A1 && (!B1 || (A2 && !B2))) {
        const int pos = atomicAdd(p.seg_n + (int64_t)q * p.n_segs + p.extra_seg, 1);
        if (pos < p.seg_cap) {
            const int64_t a = ((int64_t)q * p.n_segs + p.extra_seg) * p.seg_cap + pos;
            p.l_rows[a] = (uint32_t)row;
            p.l_k1[a] = k1s;
            if (L > 2) p.l_k2[a] = k2s;
        }
    }
}

template <int K0, int K1, int K2>
int launch_bits(const BitsParams& p, cudaStream_t st) {
    constexpr int KT = K0 + K1 + K2;
    const size_t smem = sizeof(float) * ((size_t)KT * kQT + 3 * kQT);
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_filter_bits<K0, K1, K2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int64_t row_tiles = (p.N + kRows - 1) / kRows;
    const int64_t blocks = row_tiles * p.q_tiles;
    HQ_REQUIRE(blocks < ((int64_t)1 << 31), "filter grid too large");
    k_filter_bits<K0, K1, K2><<<(unsigned)blocks, kRows, smem, st>>>(p);
    HQ_LAUNCH_OK("k_filter_bits");
    return HQ_OK;
}

}  // namespace

extern "C" int hq_filter_level_norms(const float* idx, const uint16_t* lens, int64_t N, const hq_index_layout* layout, float* rnorm,
                                     int32_t* nonuniform, void* stream) {
    HQ_REQUIRE(layout && layout->L >= 1 && layout->L <= 8, "bad index layout");
    HQ_REQUIRE(N >= 0, "negative N");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(idx && rnorm && nonuniform, "null pointer");
    int64_t blocks = (N * layout->L + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_level_norms<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, N, *layout, rnorm, nonuniform, lens);
    HQ_LAUNCH_OK("k_level_norms");
    return HQ_OK;
}

extern "C" int hq_filter_fast_supported(const hq_index_layout* layout) {
    if (!layout || layout->L < 1 || layout->L > 3 || layout->Lsum % 4 != 0) return 0;
    for (int l = 0; l < layout->L; ++l)
        if (layout->lvl_off[l] % 4 != 0 || layout->lvl_w[l] % 4 != 0) return 0;
    const int k0 = layout->lvl_keff[0], k1 = layout->L > 1 ? layout->lvl_keff[1] : 0, k2 = layout->L > 2 ? layout->lvl_keff[2] : 0;
    if (k0 > 64 || k1 > 16 || k2 > 4) return 0;
    return 1;
}

static inline int64_t plane_pitch(int64_t N) { return ((N + 31) / 32 + 7) & ~(int64_t)7; }      // words, 32-byte rows
static inline int64_t up16(int64_t b) { return (b + 15) & ~(int64_t)15; }
static inline int64_t up128(int64_t b) { return (b + 127) & ~(int64_t)127; }

struct ListGeom { int n_segs; int64_t seg_cap; bool on; };

// candidate-list geometry of the tensor-core pass for (N, Q): two segments per row range, each sized
// for a third of its rows (random data passes ~13-20 % of the rows through levels 0 and 1)
static ListGeom list_geom(int64_t N, int Q, const hq_index_layout* layout, bool extra_segment) {
    ListGeom g{0, 0, false};
    if (!layout || layout->L < 2 || !hq_filter_tc_supported(layout) || N <= 0 || Q <= 0) return g;
    int n_ranges = 0, tiles_per = 0;
    if (hq_filter_tc_plan(N, Q, &n_ranges, &tiles_per) != HQ_OK) return g;
    g.n_segs = 2 * n_ranges + (extra_segment ? 1 : 0);          // + one segment for exceptional rows
    if (g.n_segs > kMaxSegs) return g;
    g.seg_cap = (((int64_t)tiles_per * 32 + 2) / 3 + 31) & ~(int64_t)31;
    g.on = true;
    return g;
}

// Window mode (see k_filter_predict): sample stride, list geometry of the sample pass and of the window pass.
constexpr int64_t kWinMinRows = 65536;            // below this the sample would be a large part of the pass itself
constexpr int kWinMinQueries = 1;                 // (a single query against 1 M rows: 0.46 -> 0.39 ms through SearchGraph)
constexpr int kWinSampleTiles = 512;              // 64-row tiles the sample pass aims for (256 on shards below ~500 K rows)
struct WinGeom { bool on; int stride; int n_segs_s; int64_t seg_cap_s; int n_segs_w; int64_t seg_cap_w; };

static bool win_enabled() {
    static const int v = [] { const char* e = getenv("HQ_FILTER_WINDOW"); return e ? atoi(e) : 1; }();
    return v != 0;
}

static WinGeom win_geom(int64_t N, int Q, const hq_index_layout* layout) {
    WinGeom g{false, 1, 0, 0, 0, 0};
    static const int min_q = [] { const char* e = getenv("HQ_FILTER_WINDOW_MIN_Q"); return e ? atoi(e) : kWinMinQueries; }();
    if (!win_enabled() || !layout || layout->L < 2 || layout->L > 3 || !hq_filter_tc_supported(layout) || N < kWinMinRows || Q < min_q)
        return g;
    const int64_t tiles = (N + 63) / 64;
    static const int forced_tiles = [] { const char* e = getenv("HQ_FILTER_WINDOW_SAMPLE_TILES"); return e ? atoi(e) : 0; }();
    int64_t want = tiles / 16;                                              // ~6 % of the shard, within [256, 512] tiles
    want = want < 256 ? 256 : (want > kWinSampleTiles ? kWinSampleTiles : want);
    if (forced_tiles > 0) want = forced_tiles;
    int64_t stride = tiles / want;
    stride = stride < 2 ? 2 : (stride > 64 ? 64 : stride);
    int nr = 0, tp = 0;
    if (hq_filter_tc_plan_strided(N, Q, (int)stride, &nr, &tp) != HQ_OK) return g;
    g.stride = (int)stride;
    g.n_segs_s = 2 * nr;
    g.seg_cap_s = (((int64_t)tp * 32 + 2) / 3 + 31) & ~(int64_t)31;        // like the full lists: a third of the segment's rows
    if (hq_filter_tc_plan(N, Q, &nr, &tp) != HQ_OK) return g;
    g.n_segs_w = 2 * nr + 1;                                                // + the segment of the exceptional rows
    g.seg_cap_w = (((int64_t)tp * 32 + 7) / 8 + 31) & ~(int64_t)31;        // window rows: an eighth of the segment's rows
    if (g.seg_cap_w < 64) g.seg_cap_w = 64;
    if (g.n_segs_s > kMaxSegs || g.n_segs_w > kMaxSegs) return g;
    g.on = true;
    return g;
}

static inline int64_t win_list_entries(const WinGeom& g) {
    const int64_t a = (int64_t)g.n_segs_s * g.seg_cap_s, b = (int64_t)g.n_segs_w * g.seg_cap_w;
    return a > b ? a : b;
}
static inline int win_seg_slots(const WinGeom& g) { return g.n_segs_s > g.n_segs_w ? g.n_segs_s : g.n_segs_w; }

// window-mode scratch after the common part: alive plane | windows | counters | c0 of the sample, prediction flags,
// fallback tile flags | segment counts | rows, k1, k2
static int64_t win_extra_bytes(int64_t N, int Q, const WinGeom& g) {
    return 256 + up128((int64_t)Q * plane_pitch(N) * 4) + 2 * up128((int64_t)4 * Q * 4) + 3 * up128((int64_t)Q * 4) +
           up128((int64_t)Q * win_seg_slots(g) * 4) + 3 * up128((int64_t)Q * win_list_entries(g) * 4);
}

static void fill_cascade_keys(CascadeParams& cp, int L, const float* xstar, const double* ratio) {
    for (int l = 0; l < 8; ++l) {
        cp.ratio[l] = l < L ? ratio[l] : 1.0;
        // survivors of level l score >= (x*_l + 1) / 2 (up to rounding): offset keys from a little below it
        float lo = l < L && l < 3 ? (xstar[l] + 1.0f) * 0.5f : 0.f;
        if (!(lo > 0.f)) lo = 0.f;
        uint32_t klo;
        memcpy(&klo, &lo, 4);
        klo = klo > 64 ? klo - 64 : 0;
        const float one = 1.0f;
        uint32_t khi;
        memcpy(&khi, &one, 4);
        khi += 64;
        uint32_t range = khi > klo ? khi - klo : 1, sh_a = 0;
        while ((range >> sh_a) > 2047u) ++sh_a;
        cp.key_lo[l] = klo;
        cp.shift_a[l] = sh_a;
    }
}

extern "C" int64_t hq_filter_fast_scratch_bytes(int64_t N, int Q, const hq_index_layout* layout) {
    if (!layout || N <= 0 || Q <= 0) return 0;
    int grid = hq_cached_sm_count();
    if (grid > Q) grid = Q;
    // bit planes | generic cascade key / row lists | packed query operand, thresholds, norms, fallback flags |
    // candidate lists of the tensor-core pass (full lists, or the window-mode block when that is larger)
    int64_t b = (int64_t)layout->L * Q * plane_pitch(N) * 4 + up16((int64_t)grid * N * 8) + (int64_t)Q * 128 * 4 +
                3 * up16((int64_t)3 * Q * 4);
    const ListGeom g = list_geom(N, Q, layout, true);           // sized for the worst case
    int64_t lists = 0;
    if (g.on) lists = 256 + up128((int64_t)Q * g.n_segs * 4) + (int64_t)(layout->L > 2 ? 3 : 2) * Q * g.n_segs * g.seg_cap * 4;
    const WinGeom wg = win_geom(N, Q, layout);
    if (wg.on) { const int64_t w = win_extra_bytes(N, Q, wg); if (w > lists) lists = w; }
    return b + lists;
}

extern "C" int hq_filter_fast_mode(int64_t N, int Q, const hq_index_layout* layout) {
    if (!layout || N <= 0 || Q <= 0) return 0;
    if (win_geom(N, Q, layout).on) return 2;
    int grid = hq_cached_sm_count();
    if (grid > Q) grid = Q;
    const ListGeom lg = list_geom(N, Q, layout, false);
    return (lg.on && (int64_t)lg.n_segs * lg.seg_cap <= (int64_t)grid * N) ? 1 : 0;
}

extern "C" int64_t hq_filter_fast_fallback_offset(int64_t N, int Q, const hq_index_layout* layout) {
    if (!layout || N <= 0 || Q <= 0) return -1;
    int grid = hq_cached_sm_count();
    if (grid > Q) grid = Q;
    return (int64_t)layout->L * Q * plane_pitch(N) * 4 + up16((int64_t)grid * N * 8) + (int64_t)Q * 128 * 4 + 2 * up16((int64_t)3 * Q * 4);
}

extern "C" int hq_filter_fast_window_layout(int64_t N, int Q, const hq_index_layout* layout, int64_t* out) {
    HQ_REQUIRE(layout && out && N > 0 && Q > 0, "bad arguments");
    const WinGeom wg = win_geom(N, Q, layout);
    HQ_REQUIRE(wg.on, "window mode is not used for this shape");
    const int64_t fb = hq_filter_fast_fallback_offset(N, Q, layout);
    int64_t cur = fb + up16((int64_t)3 * Q * 4);
    // the block is 128-byte aligned relative to the scratch base (torch allocations are 512-byte aligned)
    cur = (cur + 127) & ~(int64_t)127;
    auto take = [&](int64_t bytes) { const int64_t r = cur; cur += up128(bytes); return r; };
    out[0] = take((int64_t)Q * plane_pitch(N) * 4);       // alive plane
    out[1] = take((int64_t)4 * Q * 4);                    // windows
    out[2] = take((int64_t)4 * Q * 4);                    // counters
    out[3] = take((int64_t)Q * 4);                        // c0 of the sample
    out[4] = take((int64_t)Q * 4);                        // prediction flags
    (void)take((int64_t)Q * 4);
    out[5] = fb + (int64_t)Q * 4;                         // query-tile flags (int32 [Q / 128]); second-stage failures follow at fb + 8 Q
    out[6] = take((int64_t)Q * win_seg_slots(wg) * 4);    // segment counts (window geometry after the search)
    out[7] = wg.n_segs_w; out[8] = wg.seg_cap_w; out[9] = wg.stride;
    return HQ_OK;
}

extern "C" int hq_filter_fast(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout, const float* q_idx, int Q,
                              const float* xstar, const double* ratio, const float* const* lvl_rows, const int32_t* lvl_pitch,
                              const float* db_packed, const uint32_t* valid, int64_t valid_pitch, const uint16_t* lens,
                              const int32_t* exc_rows, int n_exc, const double* thr, uint32_t* mask, int64_t mask_stride,
                              int32_t* n_out, int32_t* counts, void* scratch, int64_t scratch_bytes, void* stream) {
    return hq_filter_fast_rows(idx, rnorm, N, layout, q_idx, Q, xstar, ratio, lvl_rows, lvl_pitch, db_packed, nullptr, valid, valid_pitch, lens,
                               exc_rows, n_exc, thr, mask, mask_stride, n_out, counts, scratch, scratch_bytes, stream);
}

// hq_filter_fast with the scaled compact rows of hq_filter_rows_pack (`db_rows`, may be null): window-mode batches of up
// to hq_filter_rows_max_queries() queries run their window pass on the CUDA cores over those rows (hq_filter_rows.cu).
extern "C" int hq_filter_fast_rows(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout, const float* q_idx, int Q,
                                   const float* xstar, const double* ratio, const float* const* lvl_rows, const int32_t* lvl_pitch,
                                   const float* db_packed, const float* db_rows, const uint32_t* valid, int64_t valid_pitch,
                                   const uint16_t* lens, const int32_t* exc_rows, int n_exc, const double* thr, uint32_t* mask,
                                   int64_t mask_stride, int32_t* n_out, int32_t* counts, void* scratch, int64_t scratch_bytes,
                                   void* stream) {
    HQ_REQUIRE(hq_filter_fast_supported(layout), "index layout not supported by the fast filter");
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(idx && rnorm && q_idx && xstar && ratio && mask && n_out, "null pointer");
    HQ_REQUIRE(n_exc >= 0 && (n_exc == 0 || (exc_rows && lens && thr)), "exceptional rows need exc_rows, lens and thr");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    HQ_REQUIRE(mask_stride * 32 >= N, "mask stride too small");
    const int64_t need = hq_filter_fast_scratch_bytes(N, Q, layout);
    HQ_REQUIRE(scratch && scratch_bytes >= need, "scratch too small: need %lld bytes", (long long)need);
    const int64_t words = (N + 31) / 32, pitch = plane_pitch(N);
    const int L = layout->L;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = hq_cached_sm_count();
    if (grid > Q) grid = Q;
    uint32_t* const planes = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* const sc_keys = planes + (int64_t)L * Q * pitch;
    uint32_t* const sc_rows = sc_keys + (int64_t)grid * N;
    float* const q_packed = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sc_keys) + up16((int64_t)grid * N * 8));
    float* const tq = q_packed + (int64_t)Q * 128;
    float* const nq = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(tq) + up16((int64_t)3 * Q * 4));
    int32_t* const fallback = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(nq) + up16((int64_t)3 * Q * 4));
    ListGeom lg = db_packed ? list_geom(N, Q, layout, n_exc > 0) : ListGeom{0, 0, false};
    if (lg.on && n_exc > lg.seg_cap) lg.on = false;            // the extra segment could overflow: generic cascade
    if (lg.on && (int64_t)lg.n_segs * lg.seg_cap > (int64_t)grid * N) lg.on = false;   // tiny shards: no room to compact into
    HqFilterLists lists{};
    if (lg.on) {
        lists.n_segs = lg.n_segs; lists.seg_cap = lg.seg_cap;
        lists.seg_n = reinterpret_cast<int32_t*>((reinterpret_cast<uintptr_t>(fallback) + up16((int64_t)3 * Q * 4) + 127) & ~(uintptr_t)127);
        lists.rows = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(lists.seg_n) + up128((int64_t)Q * lg.n_segs * 4));
        lists.k1 = reinterpret_cast<float*>(lists.rows + (int64_t)Q * lg.n_segs * lg.seg_cap);
        lists.k2 = L > 2 ? lists.k1 + (int64_t)Q * lg.n_segs * lg.seg_cap : nullptr;
    }

    const WinGeom wg = (db_packed && valid && !counts) ? win_geom(N, Q, layout) : WinGeom{false, 1, 0, 0, 0, 0};
    if (wg.on) {
        // ================= window mode =================
        unsigned char* cur = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fallback) + up16((int64_t)3 * Q * 4) + 127) & ~(uintptr_t)127);
        auto take = [&](int64_t bytes) { unsigned char* r = cur; cur += up128(bytes); return r; };
        uint32_t* const alive_scratch = reinterpret_cast<uint32_t*>(take((int64_t)Q * pitch * 4));
        // mask rows padded like the planes (whole 32-byte sectors): the window pass writes the alive plane straight into the mask
        const bool in_place = mask_stride == pitch && (reinterpret_cast<uintptr_t>(mask) & 31) == 0;
        uint32_t* const alive = in_place ? mask : alive_scratch;
        float* const win = reinterpret_cast<float*>(take((int64_t)4 * Q * 4));
        int32_t* const wcnt = reinterpret_cast<int32_t*>(take((int64_t)4 * Q * 4));
        int32_t* const c0_s = reinterpret_cast<int32_t*>(take((int64_t)Q * 4));
        int32_t* const pflag = reinterpret_cast<int32_t*>(take((int64_t)Q * 4));
        (void)take((int64_t)Q * 4);                      // (slot kept so that the documented layout offsets stay valid)
        // flags that must survive the second stage (its full lists overwrite this window block): the fallback block has
        // room for three int32 per query -- [0, Q) window-stage failures, [Q, 2Q) query-tile flags, [2Q, 3Q) second-stage failures
        int32_t* const tile_flag = fallback + Q;
        int32_t* const fallback2 = fallback + 2 * (int64_t)Q;
        int32_t* const seg_n = reinterpret_cast<int32_t*>(take((int64_t)Q * win_seg_slots(wg) * 4));
        const int64_t ents = win_list_entries(wg);
        uint32_t* const l_rows = reinterpret_cast<uint32_t*>(take((int64_t)Q * ents * 4));
        float* const l_k1 = reinterpret_cast<float*>(take((int64_t)Q * ents * 4));
        float* const l_k2 = reinterpret_cast<float*>(take((int64_t)Q * ents * 4));
        HQ_REQUIRE(cur <= reinterpret_cast<unsigned char*>(scratch) + scratch_bytes, "internal: window-mode scratch layout exceeds the buffer");

        int rc = hq_filter_tc_prepare(layout, q_idx, Q, xstar, q_packed, tq, nq, st);
        if (rc != HQ_OK) return rc;
        // wcnt | c0_s | pflag | tile_flag are contiguous
        HQ_CUDA_OK(cudaMemsetAsync(wcnt, 0, (size_t)(reinterpret_cast<unsigned char*>(seg_n) - reinterpret_cast<unsigned char*>(wcnt)), st));
        HQ_CUDA_OK(cudaMemsetAsync(fallback, 0, (size_t)Q * 12, st));
        // 1. sample pass
        HqFilterLists ls{};
        ls.rows = l_rows; ls.k1 = l_k1; ls.k2 = L > 2 ? l_k2 : nullptr; ls.seg_n = seg_n; ls.seg_cap = wg.seg_cap_s; ls.n_segs = wg.n_segs_s;
        HqFtcOpts os{};
        os.tile_stride = wg.stride; os.c0_cnt = c0_s;
        rc = hq_filter_tc_pass(db_packed, valid, valid_pitch, N, layout, Q, q_packed, tq, nullptr, 0, &ls, &os, st);
        if (rc != HQ_OK) return rc;
        // 2. windows
        PredictParams pp{};
        pp.Q = Q; pp.L = L;
        for (int l = 0; l < 3; ++l) pp.ratio[l] = l < L ? ratio[l] : 1.0;
        // Window half width in standard deviations of the sample rank.  Measured on 24 batches of 1024 queries against
        // 1 M x 1536 rows (tools/window_fallback_rate.py): z = 3 -> 34 fallbacks in 24576 queries, 3.5 -> 6, 4 -> 1, 5 -> 0
        // (the "+ 4 ranks + bin population" slack makes the windows a little wider than z sigma).  z = 4 against z = 5:
        // window pass 1.45 -> 1.36 ms, window cascade 0.32 -> 0.27 ms; a fallback costs ~0.35 ms once per ~25 batches.
        // Large shards keep z = 5: a fallback ends in the generic gather cascade, whose cost grows with the shard (~12 ms for
        // one query against 12.5 M rows: more than the 5 ms per 4096-query batch the narrower windows save there).
        static const float zwin = [] { const char* e = getenv("HQ_FILTER_WINDOW_Z"); return e ? (float)atof(e) : 0.0f; }();
        pp.z = zwin > 0.f ? zwin : (N > 2000000 ? 5.0f : 4.0f);
        pp.tq = tq; pp.nq = nq; pp.l_k1 = l_k1; pp.l_k2 = l_k2; pp.seg_n = seg_n; pp.seg_cap = wg.seg_cap_s; pp.n_segs = wg.n_segs_s;
        pp.c0_s = c0_s; pp.win = win; pp.pflag = pflag;
        k_filter_predict<<<Q, 256, 0, st>>>(pp);
        HQ_LAUNCH_OK("k_filter_predict");
        // 3. window pass: a handful of queries over the scaled fp32 rows on the CUDA cores, a batch on the tensor cores (its
        //    thresholds folded into the query operand first)
        HQ_CUDA_OK(cudaMemsetAsync(seg_n, 0, (size_t)Q * wg.n_segs_w * 4, st));
        HqFilterLists lw{};
        lw.rows = l_rows; lw.k1 = l_k1; lw.k2 = L > 2 ? l_k2 : nullptr; lw.seg_n = seg_n; lw.seg_cap = wg.seg_cap_w; lw.n_segs = wg.n_segs_w;
        if (db_rows && Q <= hq_filter_rows_max_queries()) {
            rc = hq_filter_rows_pass(db_rows, valid, valid_pitch, N, layout, q_idx, Q, tq, win, alive, pitch, wcnt, &lw, st);
            if (rc != HQ_OK) return rc;
        } else {
            rc = hq_filter_tc_fold(layout, tq, win, Q, q_packed, st);
            if (rc != HQ_OK) return rc;
            HqFtcOpts ow{};
            ow.tile_stride = 1; ow.win = win; ow.wcnt = wcnt;
            rc = hq_filter_tc_pass(db_packed, valid, valid_pitch, N, layout, Q, q_packed, tq, alive, pitch, &lw, &ow, st);
            if (rc != HQ_OK) return rc;
        }
        ExcParams ep{};
        if (n_exc > 0) {
            ep.idx = idx; ep.lens = lens; ep.lay = *layout; ep.q_idx = q_idx; ep.Q = Q; ep.exc_rows = exc_rows; ep.n_exc = n_exc;
            for (int l = 0; l < 3; ++l) ep.thr[l] = l < L ? thr[l] : 2.0;
            ep.bits = planes; ep.bits_pitch = pitch; ep.nq = nq; ep.tq = tq;
            ep.l_rows = l_rows; ep.l_k1 = l_k1; ep.l_k2 = l_k2; ep.seg_n = seg_n; ep.seg_cap = wg.seg_cap_w; ep.n_segs = wg.n_segs_w;
            ep.extra_seg = wg.n_segs_w - 1;
            ExcWinParams ew{};
            ew.e = ep; ew.win = win; ew.wcnt = wcnt; ew.alive = alive; ew.alive_pitch = pitch;
            const int64_t pairs = (int64_t)Q * n_exc;
            k_filter_exceptions_win<<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(ew);
            HQ_LAUNCH_OK("k_filter_exceptions_win");
        }
        // 4. exact ranking of the window rows
        WinParams wp{};
        wp.alive = alive; wp.words = words; wp.pitch = pitch; wp.N = N; wp.Q = Q; wp.L = L;
        for (int l = 0; l < 3; ++l) wp.ratio[l] = l < L ? ratio[l] : 1.0;
        wp.tq = tq; wp.nq = nq; wp.win = win; wp.wcnt = wcnt; wp.pflag = pflag;
        wp.l_rows = l_rows; wp.l_k1 = l_k1; wp.l_k2 = l_k2; wp.seg_n = seg_n; wp.seg_cap = wg.seg_cap_w; wp.n_segs = wg.n_segs_w;
        wp.mask = mask; wp.mask_stride = mask_stride; wp.n_out = n_out; wp.fallback = fallback; wp.tile_flag = tile_flag;
        wp.tmp_stride = (int64_t)wg.n_segs_w * wg.seg_cap_w;
        int wgrid = hq_cached_sm_count() * kWinCtasPerSm;
        if (wgrid > Q) wgrid = Q;
        while (wgrid > 1 && (int64_t)wgrid * wp.tmp_stride > (int64_t)grid * N) --wgrid;
        HQ_REQUIRE((int64_t)wgrid * wp.tmp_stride <= (int64_t)grid * N, "internal: window lists larger than the cascade scratch");
        wp.tmp_keys = reinterpret_cast<float*>(sc_keys);
        wp.tmp_rows = sc_keys + (int64_t)wgrid * wp.tmp_stride;
        // a few queries: one large CTA per query (latency), a batch: eight small CTAs per SM (one wave)
        static const int variant = [] { const char* e = getenv("HQ_FILTER_WINDOW_CTA"); return e ? atoi(e) : 0; }();
        const bool big = variant == 256;       // (measured at 1 M rows: 341 us against 297 us for the small CTAs; 86 / 70 us at 125 K rows)
        const int tkw = hq_time_begin(2, st);
        if (Q <= hq_cached_sm_count()) k_filter_cascade_win<1024, 1, 2><<<wgrid, 1024, 0, st>>>(wp);
        else if (big) k_filter_cascade_win<256, 4, 4><<<wgrid > hq_cached_sm_count() * 4 ? hq_cached_sm_count() * 4 : wgrid, 256, 0, st>>>(wp);
        else k_filter_cascade_win<kWinThreads, kWinCtasPerSm, 2><<<wgrid, kWinThreads, 0, st>>>(wp);
        hq_time_end(2, tkw, st);
        HQ_LAUNCH_OK("k_filter_cascade_win");
        // 5. second stage for the flagged queries (a cut outside its window, a list overflow): the full-threshold pass with
        //    FULL candidate lists for their 128-query tiles, then the streaming list cascade for exactly those queries --
        //    the round-1 path, which needs no prediction.  Only what that cannot decide either (the level-0 cut binds, the
        //    full lists overflow) goes on to the generic gather cascade, whose cost grows with the shard (~12 ms per query
        //    against 12.5 M rows).  No flagged query: three launches that return at once.
        HqFtcOpts of{};
        of.tile_stride = 1; of.unit_only = tile_flag;
        const bool stage2_lists = lg.on;
        rc = hq_filter_tc_pass(db_packed, valid, valid_pitch, N, layout, Q, q_packed, tq, planes, pitch, stage2_lists ? &lists : nullptr, &of, st);
        if (rc != HQ_OK) return rc;
        if (n_exc > 0) {
            if (stage2_lists) {
                ep.l_rows = lists.rows; ep.l_k1 = lists.k1; ep.l_k2 = lists.k2; ep.seg_n = lists.seg_n; ep.seg_cap = lists.seg_cap;
                ep.n_segs = lists.n_segs; ep.extra_seg = lists.n_segs - 1;
                HQ_CUDA_OK(cudaMemset2DAsync(lists.seg_n + ep.extra_seg, (size_t)lists.n_segs * 4, 0, 4, (size_t)Q, st));
            } else {
                ep.l_rows = nullptr; ep.l_k1 = nullptr; ep.l_k2 = nullptr; ep.seg_n = nullptr;
            }
            const int64_t pairs = (int64_t)Q * n_exc;
            k_filter_exceptions<<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(ep);
            HQ_LAUNCH_OK("k_filter_exceptions");
        }
        const int32_t* generic_only = fallback;
        if (stage2_lists) {
            ListParams lp{};
            lp.bits = planes; lp.words = words; lp.bits_pitch = pitch; lp.N = N; lp.L = L; lp.Q = Q;
            for (int l = 0; l < 8; ++l) lp.ratio[l] = l < L ? ratio[l] : 1.0;
            lp.tq = tq; lp.nq = nq; lp.l_rows = lists.rows; lp.l_k1 = lists.k1; lp.l_k2 = lists.k2; lp.seg_n = lists.seg_n;
            lp.seg_cap = lists.seg_cap; lp.n_segs = lists.n_segs; lp.mask = mask; lp.mask_stride = mask_stride; lp.counts = nullptr;
            lp.n_out = n_out; lp.fallback = fallback2; lp.only = fallback;
            lp.tmp_stride = (int64_t)lists.n_segs * lists.seg_cap;
            int lgrid = hq_cached_sm_count();
            if (lgrid > Q) lgrid = Q;
            while (lgrid > 1 && (int64_t)lgrid * lp.tmp_stride > (int64_t)grid * N) --lgrid;
            lp.tmp_keys = reinterpret_cast<float*>(sc_keys);
            lp.tmp_rows = sc_keys + (int64_t)lgrid * lp.tmp_stride;
            k_filter_cascade_lists<<<lgrid, kListThreads, 0, st>>>(lp);
            HQ_LAUNCH_OK("k_filter_cascade_lists");
            generic_only = fallback2;
        }
        CascadeParams cp{};
        cp.bits = planes; cp.words = words; cp.bits_pitch = pitch; cp.idx = idx; cp.N = N; cp.lay = *layout; cp.q_idx = q_idx; cp.Q = Q;
        fill_cascade_keys(cp, L, xstar, ratio);
        for (int l = 0; l < 3; ++l) {
            cp.lvl[l] = (lvl_rows && lvl_pitch && l < L) ? lvl_rows[l] : nullptr;
            cp.lvl_pitch[l] = (lvl_rows && lvl_pitch && l < L) ? lvl_pitch[l] : 0;
            HQ_REQUIRE(!cp.lvl[l] || (cp.lvl_pitch[l] % 4 == 0 && cp.lvl_pitch[l] >= ((layout->lvl_keff[l] + 3) & ~3)), "bad level pitch");
        }
        cp.mask = mask; cp.mask_stride = mask_stride; cp.counts = nullptr; cp.n_out = n_out;
        cp.scratch_keys = sc_keys; cp.scratch_rows = sc_rows; cp.only = generic_only;
        cp.lens = n_exc > 0 ? lens : nullptr;
        k_filter_cascade<<<grid, 1024, 0, st>>>(cp);
        HQ_LAUNCH_OK("k_filter_cascade");
        return HQ_OK;
    }

    BitsParams bp{};
    bp.idx = idx; bp.rnorm = rnorm; bp.N = N; bp.lay = *layout; bp.q_idx = q_idx; bp.Q = Q;
    for (int l = 0; l < 3; ++l) bp.xstar[l] = l < L ? xstar[l] : 0.f;
    bp.bits = planes; bp.words = words; bp.pitch = pitch; bp.q_tiles = (Q + kQT - 1) / kQT;
    const int k0 = layout->lvl_keff[0], k1 = L > 1 ? layout->lvl_keff[1] : 0, k2 = L > 2 ? layout->lvl_keff[2] : 0;
    int rc;
    if (db_packed) {
        // tensor-core pass (hq_filter_tc.cu): same planes, 8x fewer SM cycles
        HQ_REQUIRE(valid, "the tensor-core filter needs the validity words of hq_filter_tc_valid");
        rc = hq_filter_bits_tc_launch(db_packed, valid, valid_pitch, N, layout, q_idx, Q, xstar, q_packed, tq, nq, planes, pitch,
                                      lg.on ? &lists : nullptr, st);
    } else if (L == 3) {
        if (k0 <= 24 && k1 <= 8) rc = launch_bits<24, 8, 4>(bp, st);
        else if (k0 <= 32 && k1 <= 8) rc = launch_bits<32, 8, 4>(bp, st);
        else rc = launch_bits<64, 16, 4>(bp, st);
    } else if (L == 2) {
        if (k0 <= 8 && k1 <= 4) rc = launch_bits<8, 4, 0>(bp, st);
        else if (k0 <= 16 && k1 <= 4) rc = launch_bits<16, 4, 0>(bp, st);
        else rc = launch_bits<64, 16, 0>(bp, st);
    } else {
        rc = k0 <= 16 ? launch_bits<16, 0, 0>(bp, st) : launch_bits<64, 0, 0>(bp, st);
    }
    if (rc != HQ_OK) return rc;

    if (n_exc > 0) {
        ExcParams ep{};
        ep.idx = idx; ep.lens = lens; ep.lay = *layout; ep.q_idx = q_idx; ep.Q = Q; ep.exc_rows = exc_rows; ep.n_exc = n_exc;
        for (int l = 0; l < 3; ++l) ep.thr[l] = l < L ? thr[l] : 2.0;
        ep.bits = planes; ep.bits_pitch = pitch; ep.nq = nq;
        if (!db_packed) {
            // the CUDA-core threshold pass does not compute the query norms: do it here
            k_query_norms<<<(Q * L + 255) / 256, 256, 0, st>>>(q_idx, Q, *layout, nq);
            HQ_LAUNCH_OK("k_query_norms");
        }
        if (lg.on) {
            ep.l_rows = lists.rows; ep.l_k1 = lists.k1; ep.l_k2 = lists.k2; ep.seg_n = lists.seg_n; ep.seg_cap = lists.seg_cap;
            ep.n_segs = lists.n_segs; ep.extra_seg = lists.n_segs - 1; ep.tq = tq;
            HQ_CUDA_OK(cudaMemset2DAsync(lists.seg_n + ep.extra_seg, (size_t)lists.n_segs * 4, 0, 4, (size_t)Q, st));
        }
        const int64_t pairs = (int64_t)Q * n_exc;
        k_filter_exceptions<<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(ep);
        HQ_LAUNCH_OK("k_filter_exceptions");
    }

    CascadeParams cp{};
    cp.bits = bp.bits; cp.words = words; cp.bits_pitch = pitch; cp.idx = idx; cp.N = N; cp.lay = *layout; cp.q_idx = q_idx; cp.Q = Q;
    fill_cascade_keys(cp, L, xstar, ratio);
    for (int l = 0; l < 3; ++l) {
        cp.lvl[l] = (lvl_rows && lvl_pitch && l < L) ? lvl_rows[l] : nullptr;
        cp.lvl_pitch[l] = (lvl_rows && lvl_pitch && l < L) ? lvl_pitch[l] : 0;
        HQ_REQUIRE(!cp.lvl[l] || (cp.lvl_pitch[l] % 4 == 0 && cp.lvl_pitch[l] >= ((layout->lvl_keff[l] + 3) & ~3)), "bad level pitch");
    }
    cp.mask = mask; cp.mask_stride = mask_stride; cp.counts = counts; cp.n_out = n_out;
    cp.scratch_keys = sc_keys;
    cp.scratch_rows = sc_rows;
    cp.only = nullptr;
    cp.lens = n_exc > 0 ? lens : nullptr;
    if (lg.on) {
        // streaming cascade over the candidate lists; the generic cascade then runs only for the queries it flags
        HQ_CUDA_OK(cudaMemsetAsync(fallback, 0, (size_t)Q * 4, st));
        HQ_CUDA_OK(cudaMemsetAsync(mask, 0, (size_t)((int64_t)(Q - 1) * mask_stride + words) * 4, st));
        ListParams lp{};
        lp.bits = planes; lp.words = words; lp.bits_pitch = pitch; lp.N = N; lp.L = L; lp.Q = Q;
        for (int l = 0; l < 8; ++l) lp.ratio[l] = l < L ? ratio[l] : 1.0;
        lp.tq = tq; lp.nq = nq; lp.l_rows = lists.rows; lp.l_k1 = lists.k1; lp.l_k2 = lists.k2; lp.seg_n = lists.seg_n;
        lp.seg_cap = lists.seg_cap; lp.n_segs = lists.n_segs; lp.mask = mask; lp.mask_stride = mask_stride; lp.counts = counts;
        lp.n_out = n_out; lp.fallback = fallback;
        // compaction scratch: shared with the generic cascade, which runs afterwards ([grid][N] keys + rows)
        lp.tmp_stride = (int64_t)lists.n_segs * lists.seg_cap;
        // Large shards: one 1024-thread CTA per SM (two 512-thread CTAs measured 7 % slower at 1 M rows).  Small
        // shards (a GPU's share of a row-sharded database): the per-query fixed work (segment table, plane popcount,
        // histogram clears / scans, cut-bin ranking: ~15 block-wide barriers) dominates the list streaming, so two
        // 512-thread CTAs per SM overlap one query's barriers with the other's loads.
        static const int forced_threads = [] { const char* e = getenv("HQ_LIST_CTA_THREADS"); return e ? atoi(e) : 0; }();
        int lthreads = forced_threads == 256 || forced_threads == 512 || forced_threads == 1024 ? forced_threads : (N <= kListSmallShardRows ? 512 : kListThreads);
        int lgrid = hq_cached_sm_count() * (kListThreads / lthreads);
        if (lgrid > Q) lgrid = Q;
        while (lgrid > 1 && (int64_t)lgrid * lp.tmp_stride > (int64_t)grid * N) --lgrid;
        HQ_REQUIRE(lp.tmp_stride <= (int64_t)grid * N, "internal: candidate lists larger than the cascade scratch");
        lp.tmp_keys = reinterpret_cast<float*>(sc_keys);
        lp.tmp_rows = sc_keys + (int64_t)lgrid * lp.tmp_stride;
        const int tkl = hq_time_begin(2, st);
        k_filter_cascade_lists<<<lgrid, lthreads, 0, st>>>(lp);
        hq_time_end(2, tkl, st);
        HQ_LAUNCH_OK("k_filter_cascade_lists");
        cp.only = fallback;
    }
    k_filter_cascade<<<grid, 1024, 0, st>>>(cp);
    HQ_LAUNCH_OK("k_filter_cascade");
    return HQ_OK;
}
