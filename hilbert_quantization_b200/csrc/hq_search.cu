// K5 coarse filter (per-level index-row cosine + threshold + exact ratio cut), the exact
// fp32 rerank path, per-query top-k and the multi-shard top-k merge.
//
// Reference semantics: rag/search/engine.py:178-287 (filter), :622-660 (cosine), :512 (sort).
#include "hq_common.cuh"
#include "hq_tc.cuh"
#include <cuda_bf16.h>
#include <float.h>

namespace {

// ------------------------------------------------------------------------------------
// stripped row lengths and row norms
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_lengths(const float* __restrict__ idx, int64_t N, hq_index_layout lay,
                                                     uint16_t* __restrict__ lens) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < N * lay.L; w += warps) {
        const int64_t row = w / lay.L;
        const int l = (int)(w - row * lay.L);
        const float* r = idx + row * lay.Lsum + lay.lvl_off[l];
        int last = 0;                                            // 1-based position of the last non-zero
        for (int i = lane; i < lay.lvl_w[l]; i += 32)
            if (__ldg(r + i) != 0.f) last = i + 1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        if (lane == 0) lens[row * lay.L + l] = (uint16_t)(last > 0 ? last : 1);
    }
}

__global__ void __launch_bounds__(256) k_row_norms(const float* __restrict__ x, int64_t N, int64_t D, int64_t stride,
                                                   float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec = (D % 4 == 0) && (stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < N; row += warps) {
        const float* r = x + row * stride;
        float acc = 0.f;
        if (vec) {
            for (int64_t i = lane; i < D / 4; i += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(r) + i);
                acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
            }
        } else {
            for (int64_t i = lane; i < D; i += 32) { const float v = __ldg(r + i); acc = fmaf(v, v, acc); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) norms[row] = sqrtf(acc);
    }
}

// ------------------------------------------------------------------------------------
// K5a: one filter level.  CTA = 256 rows x QT queries; a thread owns one row (K floats
// in registers) and sweeps the query tile four queries at a time out of shared memory.
// ------------------------------------------------------------------------------------
constexpr int kFilterRows = 256;
constexpr int kFilterQT = 64;

template <int K>
__global__ void __launch_bounds__(kFilterRows) k_filter_level(const float* __restrict__ idx, const uint16_t* __restrict__ lens,
                                                              int64_t N, hq_index_layout lay, int level, int keff,
                                                              const float* __restrict__ q_idx, const uint16_t* __restrict__ q_lens, int Q,
                                                              const uint32_t* __restrict__ mask_in, int64_t mask_stride, double thr,
                                                              float* __restrict__ scores, int64_t scores_stride,
                                                              uint32_t* __restrict__ mask_out, int32_t* __restrict__ n_alive,
                                                              int32_t* __restrict__ n_pass) {
    extern __shared__ __align__(16) unsigned char fl_smem[];
    float (*s_q)[kFilterQT] = reinterpret_cast<float (*)[kFilterQT]>(fl_smem);                       // [K][QT] transposed query tile
    float (*s_qcum)[K + 1] = reinterpret_cast<float (*)[K + 1]>(fl_smem + sizeof(float) * K * kFilterQT);   // [QT][K+1] prefix sums of squares
    float (*s_c)[K + 1] = reinterpret_cast<float (*)[K + 1]>(fl_smem + sizeof(float) * (K * kFilterQT + kFilterQT * (K + 1)));  // [rows][K+1]
    __shared__ int s_qlen[kFilterQT];
    __shared__ int s_alive[kFilterQT], s_pass[kFilterQT];

    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * kFilterRows;
    const int q0 = blockIdx.y * kFilterQT;
    const int off = lay.lvl_off[level];

    // stage rows (coalesced segments) and queries
    for (int e = tid; e < kFilterRows * K; e += kFilterRows) {
        const int r = e / K, j = e - r * K;
        const int64_t row = row0 + r;
        s_c[r][j] = (row < N && j < keff) ? __ldg(idx + row * lay.Lsum + off + j) : 0.f;
    }
    for (int e = tid; e < kFilterQT * K; e += kFilterRows) {
        const int qq = e / K, j = e - qq * K;
        const int q = q0 + qq;
        s_q[j][qq] = (q < Q && j < keff) ? __ldg(q_idx + (int64_t)q * lay.Lsum + off + j) : 0.f;
    }
    if (tid < kFilterQT) {
        const int q = q0 + tid;
        s_qlen[tid] = q < Q ? (int)q_lens[(int64_t)q * lay.L + level] : 1;
        s_alive[tid] = 0;
        s_pass[tid] = 0;
    }
    __syncthreads();
    if (tid < kFilterQT) {
        float c = 0.f;
        s_qcum[tid][0] = 0.f;
        for (int j = 0; j < K; ++j) { c = fmaf(s_q[j][tid], s_q[j][tid], c); s_qcum[tid][j + 1] = c; }
    }
    float c[K];
    float cn2 = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) { c[j] = s_c[tid][j]; cn2 = fmaf(c[j], c[j], cn2); }
    const int64_t row = row0 + tid;
    const bool in_range = row < N;
    int len_c = in_range ? (int)lens[row * lay.L + level] : 1;
    if (len_c > K) len_c = K;
    __syncthreads();

    for (int qq = 0; qq < kFilterQT; qq += 4) {
        if (q0 + qq >= Q) break;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float4 qv = *reinterpret_cast<const float4*>(&s_q[j][qq]);
            acc[0] = fmaf(c[j], qv.x, acc[0]);
            acc[1] = fmaf(c[j], qv.y, acc[1]);
            acc[2] = fmaf(c[j], qv.z, acc[2]);
            acc[3] = fmaf(c[j], qv.w, acc[3]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = q0 + qq + u;
            if (q >= Q) break;                                   // warp-uniform
            bool alive = in_range;
            if (alive && mask_in) alive = (__ldg(mask_in + (int64_t)q * mask_stride + (row >> 5)) >> (row & 31)) & 1u;
            int len_q = s_qlen[qq + u];
            if (len_q > K) len_q = K;
            float nq2, nc2;
            if (len_c <= len_q) {
                nc2 = cn2;
                nq2 = s_qcum[qq + u][len_c];
            } else {
                nq2 = s_qcum[qq + u][len_q];
                nc2 = 0.f;
#pragma unroll
                for (int j = 0; j < K; ++j) nc2 = j < len_q ? fmaf(c[j], c[j], nc2) : nc2;
            }
            const float nq = sqrtf(nq2), nc = sqrtf(nc2);
            float s = 0.f;
            if (nq != 0.f && nc != 0.f) s = __fmul_rn(__fadd_rn(__fdiv_rn(acc[u], __fmul_rn(nq, nc)), 1.0f), 0.5f);
            const bool pass = alive && ((double)s >= thr);
            if (in_range) scores[(int64_t)q * scores_stride + row] = alive ? s : -1.0f;
            const uint32_t b_alive = __ballot_sync(0xffffffffu, alive);
            const uint32_t b_pass = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) {
                if (row0 + (tid & ~31) < N) mask_out[(int64_t)q * mask_stride + ((row0 + tid) >> 5)] = b_pass;
                if (b_alive) atomicAdd(&s_alive[qq + u], __popc(b_alive));
                if (b_pass) atomicAdd(&s_pass[qq + u], __popc(b_pass));
            }
        }
    }
    __syncthreads();
    if (tid < kFilterQT && q0 + tid < Q) {
        if (s_alive[tid]) atomicAdd(n_alive + q0 + tid, s_alive[tid]);
        if (s_pass[tid]) atomicAdd(n_pass + q0 + tid, s_pass[tid]);
    }
}

// ------------------------------------------------------------------------------------
// exact k-th largest of the non-negative entries of s[0..N) (block-wide radix select on
// the float bit pattern, 11 + 11 + 10 bits).  Returns the key of the k-th largest and how
// many entries equal to it belong to the top k.
// ------------------------------------------------------------------------------------
struct SelectResult { uint32_t key; uint32_t take_ties; uint32_t ties_total; };

__device__ SelectResult block_radix_select(const float* __restrict__ s, int64_t N, uint32_t k, uint32_t* hist /*2048*/,
                                           uint32_t* sh /*4*/) {
    const int tid = threadIdx.x, nt = blockDim.x;
    uint32_t prefix = 0, pmask = 0, remaining = k;
    const int shifts[3] = {21, 10, 0};
    const int bits[3] = {11, 11, 10};
    uint32_t ties_total = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = shifts[pass];
        const uint32_t nb = 1u << bits[pass];
        for (uint32_t i = tid; i < nb; i += nt) hist[i] = 0;
        __syncthreads();
        for (int64_t i = tid; i < N; i += nt) {
            const float v = __ldg(s + i);
            if (v >= 0.f) {
                const uint32_t key = __float_as_uint(v);
                if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & (nb - 1)], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {
            // lane l owns bins [hi - (l+1)*seg, hi - l*seg) counted from the top
            const uint32_t seg = nb / 32;
            uint32_t sum = 0;
            const uint32_t top = nb - tid * seg;                 // exclusive upper bin of this lane's segment
            for (uint32_t b = 0; b < seg; ++b) sum += hist[top - 1 - b];
            uint32_t incl = sum;                                 // inclusive scan over lanes (lane 0 = highest bins)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            const uint32_t excl = incl - sum;
            const bool mine = excl < remaining && remaining <= incl;
            if (mine) {
                uint32_t above = excl, b = top;
                for (;;) {
                    --b;
                    const uint32_t h = hist[b];
                    if (above + h >= remaining) break;
                    above += h;
                }
                sh[0] = b;
                sh[1] = remaining - above;                       // rank inside the bin (1-based)
                sh[2] = hist[b];
            }
        }
        __syncthreads();
        const uint32_t b = sh[0];
        remaining = sh[1];
        ties_total = sh[2];
        prefix |= b << shift;
        pmask |= (nb - 1) << shift;
        __syncthreads();
    }
    SelectResult r;
    r.key = prefix;
    r.take_ties = remaining;
    r.ties_total = ties_total;
    return r;
}

// the same selection over keys produced by a functor: key_of(i, key) returns false for rows that do not take part
template <class KeyFn>
__device__ SelectResult block_radix_select_fn(KeyFn key_of, int64_t N, uint32_t k, uint32_t* hist /*2048*/, uint32_t* sh /*4*/) {
    const int tid = threadIdx.x, nt = blockDim.x;
    uint32_t prefix = 0, pmask = 0, remaining = k;
    const int shifts[3] = {21, 10, 0};
    const int bits[3] = {11, 11, 10};
    uint32_t ties_total = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = shifts[pass];
        const uint32_t nb = 1u << bits[pass];
        for (uint32_t i = tid; i < nb; i += nt) hist[i] = 0;
        __syncthreads();
        for (int64_t i = tid; i < N; i += nt) {
            uint32_t key;
            if (key_of(i, key) && (key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & (nb - 1)], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            const uint32_t seg = nb / 32;
            uint32_t sum = 0;
            const uint32_t top = nb - tid * seg;
            for (uint32_t b = 0; b < seg; ++b) sum += hist[top - 1 - b];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            const uint32_t excl = incl - sum;
            if (excl < remaining && remaining <= incl) {
                uint32_t above = excl, b = top;
                for (;;) {
                    --b;
                    const uint32_t h = hist[b];
                    if (above + h >= remaining) break;
                    above += h;
                }
                sh[0] = b;
                sh[1] = remaining - above;
                sh[2] = hist[b];
            }
        }
        __syncthreads();
        const uint32_t b = sh[0];
        remaining = sh[1];
        ties_total = sh[2];
        prefix |= b << shift;
        pmask |= (nb - 1) << shift;
        __syncthreads();
    }
    SelectResult r;
    r.key = prefix;
    r.take_ties = remaining;
    r.ties_total = ties_total;
    return r;
}

// block-wide exclusive prefix of a 0/1 flag for threads in row order; returns the rank
// and adds the block total to `running` (kept identical in every thread).
__device__ __forceinline__ uint32_t block_flag_rank(bool flag, uint32_t& running, uint32_t* s_warp /*32*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int i = 0; i < nw; ++i) {
        const uint32_t c = s_warp[i];
        if (i < w) before += c;
        total += c;
    }
    __syncthreads();
    const uint32_t rank = running + before + __popc(bal & ((1u << lane) - 1));
    running += total;
    return rank;
}

// K5b: ratio cut.  One CTA per query.
__global__ void __launch_bounds__(1024) k_filter_select(const float* __restrict__ scores, int64_t scores_stride, int64_t N,
                                                        const int32_t* __restrict__ n_alive, const int32_t* __restrict__ n_pass,
                                                        double ratio, uint32_t* __restrict__ mask, int64_t mask_stride,
                                                        int32_t* __restrict__ n_out) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t sh[4];
    __shared__ uint32_t s_warp[32];
    const int q = blockIdx.x;
    const int64_t na = n_alive[q], np = n_pass[q];
    int64_t cap = (int64_t)((double)na * ratio);                 // int(len * ratio), Python float semantics
    if (cap < 1) cap = 1;
    if (np <= cap) {
        if (threadIdx.x == 0) n_out[q] = (int32_t)np;
        return;
    }
    const float* s = scores + (int64_t)q * scores_stride;
    const SelectResult r = block_radix_select(s, N, (uint32_t)cap, hist, sh);
    uint32_t* m = mask + (int64_t)q * mask_stride;
    const bool ordered = r.take_ties < r.ties_total;
    uint32_t running = 0;
    const int lane = threadIdx.x & 31;
    for (int64_t base = 0; base < N; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const float v = i < N ? __ldg(s + i) : -1.f;
        const uint32_t key = __float_as_uint(v);
        bool keep = v >= 0.f && key > r.key;
        const bool tie = v >= 0.f && key == r.key;
        if (ordered) {
            const uint32_t rank = block_flag_rank(tie, running, s_warp);
            keep = keep || (tie && rank < r.take_ties);
        } else {
            keep = keep || tie;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0 && i < N) m[i >> 5] = bal;
    }
    if (threadIdx.x == 0) n_out[q] = (int32_t)cap;
}

// K5b with the REFERENCE's tie rule.  The reference sorts each level's candidate list with a stable sort
// (rag/search/engine.py:236), and the list arrives in the previous level's order: rows whose level score ties exactly at
// the cut are kept in the order of the previous level's score (then the level before that, then the row id).
// prev1 / prev0: the score matrices of levels l - 1 and l - 2 (null when absent).  Deeper ties (more than three
// levels) fall back to the row id.
__global__ void __launch_bounds__(1024) k_filter_select_prev(const float* __restrict__ scores, int64_t scores_stride, int64_t N,
                                                             const float* __restrict__ prev1, const float* __restrict__ prev0,
                                                             int64_t prev_stride, const int32_t* __restrict__ n_alive,
                                                             const int32_t* __restrict__ n_pass, double ratio,
                                                             uint32_t* __restrict__ mask, int64_t mask_stride,
                                                             int32_t* __restrict__ n_out) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t sh[4];
    __shared__ uint32_t s_warp[32];
    const int q = blockIdx.x;
    const int64_t na = n_alive[q], np = n_pass[q];
    int64_t cap = (int64_t)((double)na * ratio);
    if (cap < 1) cap = 1;
    if (np <= cap) {
        if (threadIdx.x == 0) n_out[q] = (int32_t)np;
        return;
    }
    const float* s = scores + (int64_t)q * scores_stride;
    const float* p1 = prev1 ? prev1 + (int64_t)q * prev_stride : nullptr;
    const float* p0 = prev0 ? prev0 + (int64_t)q * prev_stride : nullptr;
    const SelectResult r = block_radix_select(s, N, (uint32_t)cap, hist, sh);
    const bool ordered = r.take_ties < r.ties_total;
    SelectResult r1{0u, 0u, 0u}, r2{0u, 0u, 0u};
    const bool use1 = ordered && p1 != nullptr;
    bool use2 = false;
    if (use1) {
        r1 = block_radix_select_fn([&](int64_t i, uint32_t& key) {
                 const float v = __ldg(s + i);
                 if (!(v >= 0.f) || __float_as_uint(v) != r.key) return false;
                 key = __float_as_uint(fmaxf(__ldg(p1 + i), 0.f));
                 return true;
             }, N, r.take_ties, hist, sh);
        use2 = p0 != nullptr && r1.take_ties < r1.ties_total;
        if (use2)
            r2 = block_radix_select_fn([&](int64_t i, uint32_t& key) {
                     const float v = __ldg(s + i);
                     if (!(v >= 0.f) || __float_as_uint(v) != r.key) return false;
                     if (__float_as_uint(fmaxf(__ldg(p1 + i), 0.f)) != r1.key) return false;
                     key = __float_as_uint(fmaxf(__ldg(p0 + i), 0.f));
                     return true;
                 }, N, r1.take_ties, hist, sh);
    }
    uint32_t* m = mask + (int64_t)q * mask_stride;
    uint32_t running = 0;
    const int lane = threadIdx.x & 31;
    for (int64_t base = 0; base < N; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const float v = i < N ? __ldg(s + i) : -1.f;
        const uint32_t key = __float_as_uint(v);
        bool keep = v >= 0.f && key > r.key;
        const bool tie = v >= 0.f && key == r.key;
        if (ordered) {
            bool id_tie = tie;                       // rows still tied after every available level: lower row id first
            uint32_t take = r.take_ties;
            if (use1 && tie) {
                const uint32_t k1 = __float_as_uint(fmaxf(__ldg(p1 + i), 0.f));
                id_tie = false;
                if (k1 > r1.key) keep = true;
                else if (k1 == r1.key) {
                    if (!use2) id_tie = true;
                    else {
                        const uint32_t k0 = __float_as_uint(fmaxf(__ldg(p0 + i), 0.f));
                        if (k0 > r2.key) keep = true;
                        else if (k0 == r2.key) id_tie = true;
                    }
                }
            }
            if (use1) take = use2 ? r2.take_ties : r1.take_ties;
            const uint32_t rank = block_flag_rank(id_tie, running, s_warp);
            keep = keep || (id_tie && rank < take);
        } else {
            keep = keep || tie;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0 && i < N) m[i >> 5] = bal;
    }
    if (threadIdx.x == 0) n_out[q] = (int32_t)cap;
}

// ------------------------------------------------------------------------------------
// filter_scope = "global" (row-sharded database, the reference's SINGLE candidate list, rag/search/engine.py:272-287):
// the exact cut score of every query is found with two all-reduced histograms over the float32 bit pattern of the level
// scores (high 16 bits, then low 16 bits inside the winning high digit), ties at the cut score go to the lower GLOBAL row
// id.  These kernels are the per-shard halves; the host glue (distributed.global_ratio_cut_device) does the collectives on
// [Q, 65536] int32 histograms and [Q] vectors -- no [Q, N] boolean or int64 tensor is ever materialised (the first
// version went through eager PyTorch on such tensors: 51 GB per level for a 4096 x 12.5 M chunk).
// ------------------------------------------------------------------------------------
// hist[q][digit] += 1 for every passed row of a query that needs the cut; pass 1: digit = key >> 16, pass 2: digit = key &
// 0xffff for rows whose high digit equals d_hi[q].  grid = (chunks of rows, Q).
__global__ void __launch_bounds__(256) k_gcut_hist(const float* __restrict__ scores, int64_t scores_stride, int64_t N,
                                                   const uint32_t* __restrict__ mask, int64_t mask_stride,
                                                   const int32_t* __restrict__ need, const int32_t* __restrict__ d_hi,
                                                   int32_t* __restrict__ hist) {
    const int q = blockIdx.y;
    if (!need[q]) return;
    const float* s = scores + (int64_t)q * scores_stride;
    const uint32_t* m = mask + (int64_t)q * mask_stride;
    int32_t* h = hist + (int64_t)q * 65536;
    const int hi = d_hi ? d_hi[q] : -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        if (!((__ldg(m + (i >> 5)) >> (i & 31)) & 1u)) continue;
        const uint32_t key = __float_as_uint(fmaxf(__ldg(s + i), 0.f));
        if (d_hi) {
            if ((int)(key >> 16) == hi) atomicAdd(h + (key & 0xffffu), 1);
        } else {
            atomicAdd(h + (key >> 16), 1);
        }
    }
}

// largest digit d with #(digit >= d) >= want[q]; above[q] = #(digit > d).  One CTA of 1024 threads per query.
__global__ void __launch_bounds__(1024) k_gcut_scan(const int32_t* __restrict__ hist, const int32_t* __restrict__ need,
                                                    const int64_t* __restrict__ want, int64_t* __restrict__ d_out,
                                                    int64_t* __restrict__ above_out) {
    __shared__ long long s_part[1024];
    const int q = blockIdx.x, tid = threadIdx.x;
    if (!need[q]) return;
    const int32_t* h = hist + (int64_t)q * 65536;
    // thread t owns the 64 digits [65536 - 64 (t + 1), 65536 - 64 t): thread 0 the top ones
    const int top = 65536 - 64 * tid;
    long long sum = 0;
    for (int b = 0; b < 64; ++b) sum += h[top - 1 - b];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        long long run = 0;
        for (int t = 0; t < 1024; ++t) { const long long v = s_part[t]; s_part[t] = run; run += v; }   // exclusive, from the top
    }
    __syncthreads();
    const long long w = want[q];
    long long above = s_part[tid];
    if (above < w && w <= above + sum) {
        int b = top;
        for (;;) {
            --b;
            const long long c = h[b];
            if (above + c >= w) break;
            above += c;
        }
        d_out[q] = b;
        above_out[q] = above;
    }
}

// local number of passed rows whose score equals the cut score
__global__ void __launch_bounds__(256) k_gcut_ties(const float* __restrict__ scores, int64_t scores_stride, int64_t N,
                                                   const uint32_t* __restrict__ mask, int64_t mask_stride,
                                                   const int32_t* __restrict__ need, const int64_t* __restrict__ k_star,
                                                   int64_t* __restrict__ ties) {
    const int q = blockIdx.y;
    if (!need[q]) return;
    const float* s = scores + (int64_t)q * scores_stride;
    const uint32_t* m = mask + (int64_t)q * mask_stride;
    const uint32_t ks = (uint32_t)k_star[q];
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        if (((__ldg(m + (i >> 5)) >> (i & 31)) & 1u) && __float_as_uint(fmaxf(__ldg(s + i), 0.f)) == ks) ++c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(ties + q), c);
}

// keep = passed && (key > k*  ||  (key == k* && rank among this shard's ties (ascending row id) < quota[q]))
__global__ void __launch_bounds__(1024) k_gcut_apply(const float* __restrict__ scores, int64_t scores_stride, int64_t N,
                                                     uint32_t* __restrict__ mask, int64_t mask_stride,
                                                     const int32_t* __restrict__ need, const int64_t* __restrict__ k_star,
                                                     const int64_t* __restrict__ quota) {
    __shared__ uint32_t s_warp[32];
    const int q = blockIdx.x;
    if (!need[q]) return;
    const float* s = scores + (int64_t)q * scores_stride;
    uint32_t* m = mask + (int64_t)q * mask_stride;
    const uint32_t ks = (uint32_t)k_star[q];
    const long long qt = quota[q];
    uint32_t running = 0;
    const int lane = threadIdx.x & 31;
    for (int64_t base = 0; base < N; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        bool passed = false;
        uint32_t key = 0;
        if (i < N) {
            passed = (m[i >> 5] >> (i & 31)) & 1u;
            key = __float_as_uint(fmaxf(__ldg(s + i), 0.f));
        }
        const bool tie = passed && key == ks;
        const uint32_t rank = block_flag_rank(tie, running, s_warp);       // barriers inside: the word is read before it is rewritten
        const bool keep = passed && (key > ks || (tie && (long long)rank < qt));
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0 && i < N) m[i >> 5] = bal;
    }
}

// ------------------------------------------------------------------------------------
// exact fp32 rerank scores: tiled FMA GEMM (64 queries x 128 rows x 16 k per step)
// ------------------------------------------------------------------------------------
constexpr int kBM = 64, kBN = 128, kBK = 16;

__global__ void __launch_bounds__(256) k_rerank_scores(const float* __restrict__ db, const float* __restrict__ db_norm, int64_t N,
                                                       int64_t D, int64_t db_stride, const float* __restrict__ qm,
                                                       const float* __restrict__ q_norm, int Q, int64_t q_stride,
                                                       const uint32_t* __restrict__ mask, int64_t mask_stride,
                                                       float* __restrict__ scores, int64_t scores_stride, int vec) {
    __shared__ __align__(16) float sA[kBK][kBM + 4];
    __shared__ __align__(16) float sB[kBK][kBN + 4];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;                      // ty: 16 groups of 4 queries; tx: 16 groups of 2x4 rows
    const int64_t row0 = (int64_t)blockIdx.x * kBN;
    const int q0 = blockIdx.y * kBM;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int lr = tid >> 2, lk = (tid & 3) * 4;                 // loader: row lr (+64), k offset lk
    for (int64_t k0 = 0; k0 < D; k0 += kBK) {
        {   // queries
            const int q = q0 + lr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < Q) {
                const float* p = qm + (int64_t)q * q_stride + k0 + lk;
                if (vec && k0 + lk + 3 < D) v = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    if (k0 + lk < D) v.x = __ldg(p);
                    if (k0 + lk + 1 < D) v.y = __ldg(p + 1);
                    if (k0 + lk + 2 < D) v.z = __ldg(p + 2);
                    if (k0 + lk + 3 < D) v.w = __ldg(p + 3);
                }
            }
            sA[lk][lr] = v.x; sA[lk + 1][lr] = v.y; sA[lk + 2][lr] = v.z; sA[lk + 3][lr] = v.w;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 64 * h;
            const int64_t row = row0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < N) {
                const float* p = db + row * db_stride + k0 + lk;
                if (vec && k0 + lk + 3 < D) v = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    if (k0 + lk < D) v.x = __ldg(p);
                    if (k0 + lk + 1 < D) v.y = __ldg(p + 1);
                    if (k0 + lk + 2 < D) v.z = __ldg(p + 2);
                    if (k0 + lk + 3 < D) v.w = __ldg(p + 3);
                }
            }
            sB[lk][r] = v.x; sB[lk + 1][r] = v.y; sB[lk + 2][r] = v.z; sB[lk + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sB[k][64 + tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= Q) continue;
        const float nq = __ldg(q_norm + q);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t rbase = row0 + 64 * h + tx * 4;
            uint32_t mw = 0xffffffffu;
            if (mask && rbase < N) mw = __ldg(mask + (int64_t)q * mask_stride + (rbase >> 5));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t row = rbase + j;
                if (row >= N) continue;
                const bool alive = (mw >> (row & 31)) & 1u;
                const float nc = __ldg(db_norm + row);
                float s = 0.f;
                if (nq != 0.f && nc != 0.f) s = __fmul_rn(__fadd_rn(__fdiv_rn(acc[i][4 * h + j], __fmul_rn(nq, nc)), 1.0f), 0.5f);
                scores[(int64_t)q * scores_stride + row] = alive ? s : -1.0f;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// a13 for a handful of queries: exact fp32 cosine of the SURVIVING rows only.  One warp per (query, 32-row mask word):
// dead rows get -1, a surviving row is scored by the whole warp with the arithmetic of the tensor-core path's exact
// re-score (k_rerank_tc_merge: lane-strided fmaf chain, xor tree), so both paths return identical scores.  A single
// query against 1 M x 1536 reads ~45 K surviving rows (0.3 GB) instead of the 3 GB bf16 database.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rerank_sparse(const float* __restrict__ db, const float* __restrict__ db_norm, int64_t N,
                                                       int64_t D, int64_t db_stride, const float* __restrict__ qm,
                                                       const float* __restrict__ q_norm, int Q, int64_t q_stride,
                                                       const uint32_t* __restrict__ mask, int64_t mask_stride,
                                                       float* __restrict__ scores, int64_t scores_stride, int vec) {
    const int lane = threadIdx.x & 31;
    const int64_t words = (N + 31) / 32;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < (int64_t)Q * words; t += warps) {
        const int q = (int)(t / words);
        const int64_t w = t - (int64_t)q * words;
        uint32_t m = mask ? __ldg(mask + (int64_t)q * mask_stride + w) : 0xffffffffu;
        if (w * 32 + 32 > N) m &= (1u << (uint32_t)(N - w * 32)) - 1u;
        const float nq = __ldg(q_norm + q);
        const float* qv = qm + (int64_t)q * q_stride;
        float out = -1.0f;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int64_t row = w * 32 + b;
            const float* rv = db + row * db_stride;
            float acc = 0.f;
            if (vec) {
                for (int64_t i = lane; i < D / 4; i += 32) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(qv) + i);
                    const float4 c = __ldg(reinterpret_cast<const float4*>(rv) + i);
                    acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc); acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
                }
            } else {
                for (int64_t i = lane; i < D; i += 32) acc = fmaf(__ldg(qv + i), __ldg(rv + i), acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            const float nc = __ldg(db_norm + row);
            float sc = 0.f;
            if (nq != 0.f && nc != 0.f) sc = __fmul_rn(__fadd_rn(__fdiv_rn(acc, __fmul_rn(nq, nc)), 1.0f), 0.5f);
            if (lane == b) out = sc;
        }
        const int64_t row = w * 32 + lane;
        if (row < N) scores[(int64_t)q * scores_stride + row] = out;
    }
}

// ------------------------------------------------------------------------------------
// a13 + a15 for a handful of queries in ONE launch: exact fp32 cosine of the surviving rows (the arithmetic of k_rerank_sparse,
// bit for bit) with the top-k taken on the way -- no [Q, N] score row, no separate top-k launches.
//   grid (P, Q): the P CTAs of a query draw chunks of 16 mask words from a ticket counter (the survivors are a few per cent
//     of the rows and unevenly spread);
//   warp 0 of a CTA walks the mask and fetches every surviving row into a ring of shared-memory slots with ONE bulk copy
//     (cp.async.bulk + full / empty mbarriers, up to 64 rows in flight per SM); the other warps take the slots in turn and
//     score a row each, the query sits in shared memory.  (The D-strided loop of k_rerank_sparse waits for HBM once per 128
//     values -- 12 round trips per 1536-D row: ptxas sinks the loads next to their FMAs even when the source issues them
//     first; two buffers per warp with one copy ahead: 112 us for 57 K rows of 6 KB, slower than that loop.)
//   every warp keeps its best 32 (score desc, id asc) one per lane, the CTA ranks its warps' lists into a list of k, and the
//     LAST CTA of the query to finish (a done counter) merges the P lists: only the entries that reach tau = max over the
//     full lists of their minimum can be in the result.
// counters: [2 * Q] uint32, zero on entry (ticket, done); the last CTA leaves them zero again.
// ------------------------------------------------------------------------------------
constexpr int kSparseTopkThreads = 512;
constexpr int kSparseTopkMaxK = 32;
constexpr int kSparseTopkSlots = 64;           // most row slots of the shared-memory ring
constexpr int kSparseTopkSel = 1024;           // candidates the final merge ranks in shared memory

__device__ __forceinline__ bool cand_better(float s, int64_t id, float s2, int64_t id2) { return s > s2 || (s == s2 && id < id2); }

// B16: the rows of phase one are the shard's UNIT rows in bf16 (half the bytes: 57 K scattered rows arrive at ~4 TB/s whatever
// their size); the lists then hold the accumulators a = q . c16 (which order the rows like the cosine up to the bf16
// rounding of c), 32 per list, and the last CTA re-scores the best kSparseTopkR of the merged lists EXACTLY from the fp32
// rows (same arithmetic as the fp32 pass: identical scores) and checks that the result is proven:
//     t_k  >  max(a of everything that was not re-scored)  +  E,      E = |q| (dc_max + 1e-5 + 2.4e-7 D)
// (t_k = k-th best exact value q . c / |c|; |a - t| <= |q| |c16 - c / |c|| + accumulation error, the bound of
// k_rerank_tc_merge without the query's own rounding).  A query that fails the check -- a cluster of near duplicates around
// rank k -- raises flags[q]; the fp32 kernel launched right behind (`only_flagged`) redoes exactly those queries.
constexpr int kSparseTopkR = 32;               // rows re-scored exactly by the last CTA (B16)
constexpr int kSparseTopkSelB = 512;           // candidates the B16 merge ranks in shared memory

template <bool B16>
__global__ void __launch_bounds__(kSparseTopkThreads) k_rerank_sparse_topk(
    const float* __restrict__ db, const float* __restrict__ db_norm, int64_t N, int64_t D, int64_t db_stride,
    const __nv_bfloat16* __restrict__ db16, int64_t db16_pitch, float dc_max,
    const float* __restrict__ qm, const float* __restrict__ q_norm, int Q, int64_t q_stride, const uint32_t* __restrict__ mask,
    int64_t mask_stride, int k, int64_t id_base, int64_t* __restrict__ part_ids, float* __restrict__ part_scores,
    uint32_t* __restrict__ counters, uint32_t* __restrict__ flags, int only_flagged, int64_t* __restrict__ out_ids,
    float* __restrict__ out_scores, int S, int NP, int R) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    if (only_flagged && __ldcg(flags + blockIdx.y) == 0u) return;
    constexpr int NWMAX = kSparseTopkThreads / 32;
    __shared__ float s_ls[NWMAX * 32];
    __shared__ int64_t s_lid[NWMAX * 32];
    __shared__ uint32_t s_nvalid, s_last, s_m;
    __shared__ float s_wmax[NWMAX];
    __shared__ __align__(8) uint64_t s_full[kSparseTopkSlots], s_empty[kSparseTopkSlots];
    __shared__ int s_row[kSparseTopkSlots];
    __shared__ __align__(8) uint64_t s_rbar;
    __shared__ uint32_t s_scan[NWMAX];
    __shared__ int s_cut;
    const int q = blockIdx.y, P = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nw = nthr >> 5;
    const int NC = nw - NP, depth = S / NC;                                           // consumer warps, row slots of each (1, 2 or 4)
    const int NCg = NC / NP;                                                          // consumers served by one producer warp
    const uint32_t ldepth = depth == 4 ? 2u : (depth == 2 ? 1u : 0u);
    const int64_t words = (N + 31) / 32;
    const int n4 = (int)(D / 4);
    const int kl = B16 ? 32 : k;                                                      // entries per list
    const uint32_t row_bytes = B16 ? (uint32_t)D * 2u : (uint32_t)D * 4u;
    const uint32_t buf_bytes = (row_bytes + 127u) & ~127u;
    const uint32_t q_bytes = ((uint32_t)D * 4u + 127u) & ~127u;
    float4* qs = reinterpret_cast<float4*>(sm_raw);                                   // the query
    unsigned char* ring = sm_raw + q_bytes;                                           // S row slots
    for (int i = tid; i < n4; i += nthr) qs[i] = __ldg(reinterpret_cast<const float4*>(qm + (int64_t)q * q_stride) + i);
    if (tid == 0) { s_nvalid = 0; s_m = 0; hq_tc::mbar_init(&s_rbar, 1); }
    if (tid < S) { hq_tc::mbar_init(s_full + tid, 1); hq_tc::mbar_init(s_empty + tid, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    float my_s = -INFINITY;                   // lane j: the warp's j-th best so far
    int64_t my_id = INT64_MAX;
    if (warp < NP) {
        // ================= producers: walk the mask, one bulk copy per surviving row =================
        // Chunks of 16 mask words from the query's ticket counter; the ticket two chunks ahead and the words one chunk ahead
        // are in flight while the rows of the current chunk are issued.  Lane l owns 16 bits of the chunk (word l / 2, half
        // l % 2) and issues the copies of its own rows; the lanes of a round take consecutive sequence numbers.
        // A producer warp serves NCg consumer warps; its sequence number i belongs to its consumer i % NCg and to that
        // consumer's slot (i / NCg) % depth: the phases of a slot are
        // filled and drained strictly in order by ONE warp each (with a shared ring a warp could meet a slot two phases behind
        // -- bulk copies complete out of order -- and a parity wait cannot tell that from "ready").
        // (One lane issuing every copy, with the slot arithmetic as integer divisions: 870 cycles per row, 183 us per query;
        // one producer warp with 32 issuing lanes: 90 us, the consumers still waiting 40 % of the time.)
        const int64_t n_chunks = (words + 15) / 16;
        auto ticket = [&]() -> uint32_t { return lane == 0 ? atomicAdd(counters + 2 * q, 1u) : 0u; };
        auto load_half = [&](uint32_t ch) -> uint32_t {
            const int64_t w = 16 * (int64_t)ch + (lane >> 1);
            uint32_t mw = 0;
            if ((int64_t)ch < n_chunks && w < words) {
                mw = __ldg(mask + (int64_t)q * mask_stride + w);
                if (w * 32 + 32 > N) mw &= (1u << (uint32_t)(N - w * 32)) - 1u;
            }
            return (mw >> (16 * (lane & 1))) & 0xffffu;
        };
        uint32_t ch1 = __shfl_sync(0xffffffffu, ticket(), 0);
        uint32_t h1 = load_half(ch1);
        uint32_t t2 = ticket();
        uint32_t c_next = 0, turn_base = 0;                    // consumer and turn of the next sequence number
        const uint32_t dmask = (uint32_t)depth - 1u;           // depth is a power of two
        const uint32_t lower = (1u << lane) - 1u;
        const uint32_t cap = (uint32_t)(NCg * depth);          // slots of this producer
        for (;;) {
            const uint32_t ch = ch1;
            uint32_t h = h1;
            if ((int64_t)ch >= n_chunks) break;
            ch1 = __shfl_sync(0xffffffffu, t2, 0);
            h1 = load_half(ch1);
            t2 = ticket();
            const int64_t row_base = (16 * (int64_t)ch + (lane >> 1)) * 32 + 16 * (lane & 1);
            for (;;) {
                const bool act = h != 0;
                const uint32_t bal = __ballot_sync(0xffffffffu, act);
                if (!bal) break;
                // at most one row per slot and round: a lane never waits for a slot that a lane of the same round fills
                const uint32_t rank = (uint32_t)__popc(bal & lower);
                if (act && rank < cap) {
                    const int b = __ffs(h) - 1;
                    h &= h - 1;
                    uint32_t c = c_next + rank, turn = turn_base;
                    while (c >= (uint32_t)NCg) { c -= (uint32_t)NCg; ++turn; }
                    const uint32_t slot = (uint32_t)(warp * NCg) + c + (uint32_t)NC * (turn & dmask), use = turn >> ldepth;
                    if (use >= 1u) hq_tc::mbar_wait(s_empty + slot, (use - 1u) & 1u);
                    const int64_t row = row_base + b;
                    s_row[slot] = (int)row;
                    hq_tc::mbar_expect_tx(s_full + slot, row_bytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     hq_tc::smem_u32(ring + (size_t)slot * buf_bytes)),
                                 "l"(B16 ? reinterpret_cast<const void*>(db16 + row * db16_pitch) : reinterpret_cast<const void*>(db + row * db_stride)),
                                 "r"(row_bytes), "r"(hq_tc::smem_u32(s_full + slot))
                                 : "memory");
                }
                c_next += min((uint32_t)__popc(bal), cap);
                while (c_next >= (uint32_t)NCg) { c_next -= (uint32_t)NCg; ++turn_base; }
            }
        }
        if (lane < NCg) {                                       // one end marker per consumer warp of this producer
            uint32_t c = c_next + (uint32_t)lane, turn = turn_base;
            while (c >= (uint32_t)NCg) { c -= (uint32_t)NCg; ++turn; }
            const uint32_t slot = (uint32_t)(warp * NCg) + c + (uint32_t)NC * (turn & dmask), use = turn >> ldepth;
            if (use >= 1u) hq_tc::mbar_wait(s_empty + slot, (use - 1u) & 1u);
            s_row[slot] = -1;
            hq_tc::mbar_arrive(s_full + slot);
        }
        __syncwarp();
    } else {
        // ================= consumers: sequence numbers warp - 1, warp - 1 + NC, ... =================
        const float nq = __ldg(q_norm + q);
        float thr = -INFINITY;                    // value of lane kl - 1
        for (uint32_t turn = 0;; ++turn) {
            const uint32_t slot = (uint32_t)(warp - NP) + (uint32_t)NC * (turn & ((uint32_t)depth - 1u));
            hq_tc::mbar_wait(s_full + slot, (turn >> ldepth) & 1u);     // (a suspend-time hint here cost a microsecond per row)
            const int row = s_row[slot];
            if (row < 0) break;
            const float nc = __ldg(db_norm + row);
            float acc = 0.f;
            if constexpr (B16) {
                const uint4* rv = reinterpret_cast<const uint4*>(ring + (size_t)slot * buf_bytes);
#pragma unroll 2
                for (int i8 = lane; i8 < (n4 >> 1); i8 += 32) {
                    const uint4 w = rv[i8];
                    const float4 a0 = qs[2 * i8], a1 = qs[2 * i8 + 1];
                    acc = fmaf(a0.x, __uint_as_float(w.x << 16), acc); acc = fmaf(a0.y, __uint_as_float(w.x & 0xffff0000u), acc);
                    acc = fmaf(a0.z, __uint_as_float(w.y << 16), acc); acc = fmaf(a0.w, __uint_as_float(w.y & 0xffff0000u), acc);
                    acc = fmaf(a1.x, __uint_as_float(w.z << 16), acc); acc = fmaf(a1.y, __uint_as_float(w.z & 0xffff0000u), acc);
                    acc = fmaf(a1.z, __uint_as_float(w.w << 16), acc); acc = fmaf(a1.w, __uint_as_float(w.w & 0xffff0000u), acc);
                }
            } else {
                const float4* rv = reinterpret_cast<const float4*>(ring + (size_t)slot * buf_bytes);
#pragma unroll 4
                for (int i4 = lane; i4 < n4; i4 += 32) {      // lane-strided fmaf chain: the order of k_rerank_sparse
                    const float4 a = qs[i4];
                    const float4 c = rv[i4];
                    acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc); acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) hq_tc::mbar_arrive(s_empty + slot);                 // every lane's reads precede the shuffles
            float sc = 0.f;
            if constexpr (B16) {
                sc = nc != 0.f ? acc : -1.0e30f;               // zero rows (exact score 0) rank below every other row
            } else {
                if (nq != 0.f && nc != 0.f) sc = __fmul_rn(__fadd_rn(__fdiv_rn(acc, __fmul_rn(nq, nc)), 1.0f), 0.5f);
            }
            if (sc < thr) continue;                                   // warp-uniform
            const uint32_t bt = __ballot_sync(0xffffffffu, cand_better(sc, (int64_t)row, my_s, my_id));
            if (!bt) continue;
            const int pos = __ffs(bt) - 1;                            // the list is sorted: better than lanes pos .. 31
            const float up_s = __shfl_up_sync(0xffffffffu, my_s, 1);
            const int64_t up_id = __shfl_up_sync(0xffffffffu, my_id, 1);
            if (lane > pos) { my_s = up_s; my_id = up_id; }
            else if (lane == pos) { my_s = sc; my_id = row; }
            thr = __shfl_sync(0xffffffffu, my_s, kl - 1);
        }
    }
    // ---- the CTA's list: rank the warps' entries by counting ----
    s_ls[tid] = lane < kl ? my_s : -INFINITY;
    s_lid[tid] = lane < kl ? my_id : INT64_MAX;
    __syncthreads();
    {
        const float v = s_ls[tid];
        const int64_t id = s_lid[tid];
        if (id != INT64_MAX) {
            int rank = 0;
            for (int j = 0; j < nthr; ++j) rank += cand_better(s_ls[j], s_lid[j], v, id) ? 1 : 0;
            atomicAdd(&s_nvalid, 1u);
            if (rank < kl) {
                part_ids[((int64_t)q * P + blockIdx.x) * kl + rank] = id;
                part_scores[((int64_t)q * P + blockIdx.x) * kl + rank] = v;
            }
        }
    }
    __syncthreads();
    for (int j = (int)min(s_nvalid, (uint32_t)kl) + tid; j < kl; j += nthr) {
        part_ids[((int64_t)q * P + blockIdx.x) * kl + j] = -1;
        part_scores[((int64_t)q * P + blockIdx.x) * kl + j] = -1.0f;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(counters + 2 * q + 1, 1u) == (uint32_t)(P - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last CTA of the query: merge the P lists ----
    const int64_t* pi = part_ids + (int64_t)q * P * kl;
    const float* ps = part_scores + (int64_t)q * P * kl;
    const int M = P * kl;
    float tau = -INFINITY;
    for (int p = tid; p < P; p += nthr)
        if (__ldcg(pi + (int64_t)p * kl + kl - 1) >= 0) tau = fmaxf(tau, __ldcg(ps + (int64_t)p * kl + kl - 1));   // sorted: last = minimum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tau = fmaxf(tau, __shfl_xor_sync(0xffffffffu, tau, o));
    if (lane == 0) s_wmax[warp] = tau;
    __syncthreads();
    tau = -INFINITY;
    for (int w = 0; w < nw; ++w) tau = fmaxf(tau, s_wmax[w]);
    const int sel_cap = B16 ? kSparseTopkSelB : kSparseTopkSel;
    unsigned char* area = B16 ? sm_raw + q_bytes : sm_raw;                // fp32 pass: the query is no longer needed either
    int64_t* c_id = reinterpret_cast<int64_t*>(area);
    float* c_s = reinterpret_cast<float*>(c_id + sel_cap);
    // The candidates that can still matter: the best `need` entries of the pool (the result, or the shortlist and the entry
    // behind it), found with a 2048-bin histogram over [lo, hi] -- lo = tau (a full list proves that its whole length
    // reaches its minimum) or the pool's minimum.  (Everything >= tau: 1 100 of the 1 480 entries of a query's lists.)
    uint32_t* hist = reinterpret_cast<uint32_t*>(area + (((size_t)sel_cap * 12 + 127) & ~(size_t)127));   // [2048]
    const uint32_t need = B16 ? (uint32_t)R + 1u : (uint32_t)k;
    float vmax = -INFINITY, vmin = INFINITY;
    for (int e = tid; e < M; e += nthr)
        if (__ldcg(pi + e) >= 0) { const float v = __ldcg(ps + e); vmax = fmaxf(vmax, v); vmin = fminf(vmin, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    }
    __syncthreads();
    if (lane == 0) { s_wmax[warp] = vmax; s_ls[warp] = vmin; }
    for (int b = tid; b < 2048; b += nthr) hist[b] = 0;
    if (tid == 0) s_cut = 0;
    __syncthreads();
    vmax = -INFINITY; vmin = INFINITY;
    for (int w = 0; w < nw; ++w) { vmax = fmaxf(vmax, s_wmax[w]); vmin = fminf(vmin, s_ls[w]); }
    const float lo = tau > -INFINITY ? tau : vmin;
    const float bscale = vmax > lo ? 2047.0f / (vmax - lo) : 0.f;
    auto bin_of = [&](float v) -> int {
        const float x = (v - lo) * bscale;
        return x >= 2047.0f ? 2047 : (x > 0.f ? (int)x : 0);
    };
    for (int e = tid; e < M; e += nthr) {
        const float v = __ldcg(ps + e);
        if (__ldcg(pi + e) >= 0 && !(v < lo)) atomicAdd(&hist[bin_of(v)], 1u);
    }
    __syncthreads();
    {
        // thread t owns the bins [2048 - (t + 1) per, 2048 - t per): suffix counts by a scan over the threads, top bins first
        const int per = (2048 + nthr - 1) / nthr;
        const int b_hi = 2048 - tid * per, b_lo = max(0, b_hi - per);
        uint32_t sum = 0;
        for (int b = b_hi - 1; b >= b_lo; --b) sum += hist[b];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int w = 0; w < warp; ++w) run += s_scan[w];
        for (int b = b_hi - 1; b >= b_lo; --b) {
            const uint32_t before = run;
            run += hist[b];
            if (before < need && run >= need) s_cut = b;          // the highest bin whose suffix count reaches `need`
        }
    }
    __syncthreads();
    const int cut = s_cut;
    for (int e = tid; e < M; e += nthr) {
        const int64_t id = __ldcg(pi + e);
        const float v = __ldcg(ps + e);
        if (id >= 0 && !(v < lo) && bin_of(v) >= cut) {
            const uint32_t slot = atomicAdd(&s_m, 1u);
            if (slot < (uint32_t)sel_cap) { c_id[slot] = id; c_s[slot] = v; }
        }
    }
    for (int j = tid; j < k; j += nthr) { out_ids[(int64_t)q * k + j] = -1; out_scores[(int64_t)q * k + j] = -1.0f; }
    __syncthreads();
    const int m = (int)s_m;
    if constexpr (B16) {
        // shortlist = the best R candidates by accumulator; everything else is bounded by max(tau, the (R + 1)-th accumulator)
        __shared__ int32_t r_id[kSparseTopkR];
        __shared__ float r_s[kSparseTopkR], r_t[kSparseTopkR];
        __shared__ float s_next;
        __shared__ int s_pass;
        bool proven = m <= sel_cap;
        const int n_sel = m < R ? m : R;
        if (tid == 0) { s_next = -INFINITY; s_pass = 0; }
        __syncthreads();
        if (proven) {
            for (int t = tid; t < m; t += nthr) {
                const int64_t id = c_id[t];
                const float v = c_s[t];
                int rank = 0;
                for (int u = 0; u < m; ++u) rank += cand_better(c_s[u], c_id[u], v, id) ? 1 : 0;
                if (rank < R) r_id[rank] = (int32_t)id;
                else if (rank == R) s_next = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the ring's slots are about to be overwritten by bulk copies
            __syncthreads();
            // exact values of the shortlist: the fp32 rows as bulk copies behind the candidate area, a warp per row
            const uint32_t rb = (uint32_t)D * 4u;
            unsigned char* rows_sm = area + (((size_t)sel_cap * 12 + 127) & ~(size_t)127);
            if (tid == 0 && n_sel > 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                hq_tc::mbar_expect_tx(&s_rbar, (uint32_t)n_sel * rb);
                for (int i = 0; i < n_sel; ++i)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     hq_tc::smem_u32(rows_sm + (size_t)i * q_bytes)),
                                 "l"(db + (int64_t)r_id[i] * db_stride), "r"(rb), "r"(hq_tc::smem_u32(&s_rbar))
                                 : "memory");
            }
            if (n_sel > 0) hq_tc::mbar_wait(&s_rbar, 0u);
            const float nq = __ldg(q_norm + q);
            for (int i = warp; i < n_sel; i += nw) {
                const float4* rv = reinterpret_cast<const float4*>(rows_sm + (size_t)i * q_bytes);
                float acc = 0.f;
#pragma unroll 4
                for (int i4 = lane; i4 < n4; i4 += 32) {      // the arithmetic of the fp32 pass: identical scores
                    const float4 a = qs[i4];
                    const float4 c = rv[i4];
                    acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc); acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) {
                    const float nc = __ldg(db_norm + r_id[i]);
                    float sc = 0.f;
                    if (nq != 0.f && nc != 0.f) sc = __fmul_rn(__fadd_rn(__fdiv_rn(acc, __fmul_rn(nq, nc)), 1.0f), 0.5f);
                    r_s[i] = sc;
                    r_t[i] = nc != 0.f ? __fdiv_rn(acc, nc) : -1.0e30f;
                }
            }
            __syncthreads();
            if (warp == 0) {
                const float bound = fmaxf(tau, s_next);                    // -inf: every surviving row was re-scored
                int pass = 0;
                if (bound == -INFINITY) pass = 1;
                else if (n_sel >= k) {
                    bool mine = false;                                     // lane i: is r_t[i] the k-th best exact value?
                    if (lane < n_sel) {
                        int rank = 0;
                        for (int j = 0; j < n_sel; ++j) rank += (r_t[j] > r_t[lane] || (r_t[j] == r_t[lane] && j < lane)) ? 1 : 0;
                        mine = rank == k - 1;
                    }
                    const uint32_t who = __ballot_sync(0xffffffffu, mine);
                    const float tk = r_t[__ffs(who) - 1];
                    const float E = nq * (dc_max + 1.0e-5f + (float)D * 2.4e-7f);
                    pass = tk > bound + E ? 1 : 0;
                }
                if (lane == 0) s_pass = pass;
            }
            __syncthreads();
            proven = s_pass != 0;
        }
        if (proven) {
            if (tid < n_sel) {
                const int32_t id = r_id[tid];
                const float v = r_s[tid];
                int rank = 0;
                for (int j = 0; j < n_sel; ++j) rank += (r_s[j] > v || (r_s[j] == v && r_id[j] < id)) ? 1 : 0;
                if (rank < k) { out_ids[(int64_t)q * k + rank] = (int64_t)id + id_base; out_scores[(int64_t)q * k + rank] = v; }
            }
        } else if (tid == 0) {
            flags[q] = 1u;                                                 // the fp32 pass behind this launch redoes the query
        }
    } else {
        if (m <= sel_cap) {
            for (int t = tid; t < m; t += nthr) {
                const int64_t id = c_id[t];
                const float v = c_s[t];
                int rank = 0;
                for (int u = 0; u < m; ++u) rank += cand_better(c_s[u], c_id[u], v, id) ? 1 : 0;
                if (rank < k) { out_ids[(int64_t)q * k + rank] = id + id_base; out_scores[(int64_t)q * k + rank] = v; }
            }
        } else {
            // heavily tied scores: more candidates than the shared list holds -- rank them against the global lists
            for (int e = tid; e < M; e += nthr) {
                const int64_t id = __ldcg(pi + e);
                const float v = __ldcg(ps + e);
                if (id < 0 || v < lo || bin_of(v) < cut) continue;
                int rank = 0;
                for (int u = 0; u < M && rank < k; ++u) {
                    const int64_t idu = __ldcg(pi + u);
                    rank += (idu >= 0 && cand_better(__ldcg(ps + u), idu, v, id)) ? 1 : 0;
                }
                if (rank < k) { out_ids[(int64_t)q * k + rank] = id + id_base; out_scores[(int64_t)q * k + rank] = v; }
            }
        }
        if (only_flagged && tid == 0) flags[q] = 0u;
    }
    if (tid == 0) { counters[2 * q] = 0; counters[2 * q + 1] = 0; }
}

// ------------------------------------------------------------------------------------
// per-query top-k of a score row (exact; ties -> lower row id).  One CTA per query.
// ------------------------------------------------------------------------------------
constexpr int kMaxK = 1024;

// `parts` > 1: every query's row is cut into `parts` chunks of N scores (the last one clipped at n_total) and CTA
// (query * parts + part) returns the top-k of its chunk with row ids counted from the start of the row; k_topk_merge
// then merges the chunk lists.  One CTA walking a whole 1 M-score row took 0.65 ms of a 1.25 ms single-query search.
__global__ void __launch_bounds__(1024) k_topk_scores(const float* __restrict__ scores, int64_t scores_stride, int64_t N, int k,
                                                      int64_t id_base, int64_t* __restrict__ ids, float* __restrict__ out_scores,
                                                      int parts, int64_t n_total) {
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t sh[4];
    __shared__ uint32_t s_warp[32];
    __shared__ float s_val[kMaxK];
    __shared__ int64_t s_id[kMaxK];
    __shared__ uint32_t s_cnt;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int part = q % parts;
    const float* s = scores + (int64_t)(q / parts) * scores_stride + (int64_t)part * N;
    id_base += (int64_t)part * N;
    if (n_total - (int64_t)part * N < N) N = n_total - (int64_t)part * N > 0 ? n_total - (int64_t)part * N : 0;

    // how many live entries are there?  (needed when fewer than k survive)
    uint32_t live = 0;
    for (int64_t i = tid; i < N; i += blockDim.x) live += __ldg(s + i) >= 0.f ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
    if ((tid & 31) == 0) s_warp[tid >> 5] = live;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    uint32_t total_live = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total_live += s_warp[i];
    __syncthreads();
    const uint32_t kk = total_live < (uint32_t)k ? total_live : (uint32_t)k;
    for (int i = tid; i < kMaxK; i += blockDim.x) { s_val[i] = -2.f; s_id[i] = INT64_MAX; }
    __syncthreads();
    if (kk > 0) {
        const SelectResult r = block_radix_select(s, N, kk, hist, sh);
        const bool ordered = r.take_ties < r.ties_total;
        uint32_t running = 0;
        for (int64_t base = 0; base < N; base += blockDim.x) {
            const int64_t i = base + tid;
            const float v = i < N ? __ldg(s + i) : -1.f;
            const uint32_t key = __float_as_uint(v);
            bool keep = v >= 0.f && key > r.key;
            const bool tie = v >= 0.f && key == r.key;
            if (ordered) {
                const uint32_t rank = block_flag_rank(tie, running, s_warp);
                keep = keep || (tie && rank < r.take_ties);
            } else {
                keep = keep || tie;
            }
            if (keep) {
                const uint32_t slot = atomicAdd(&s_cnt, 1u);
                if (slot < (uint32_t)kMaxK) { s_val[slot] = v; s_id[slot] = i; }
            }
        }
        __syncthreads();
    }
    // rank by (score desc, id asc) -- k <= 1024, one thread per entry
    if (tid < kMaxK && tid < (int)kk) {
        const float v = s_val[tid];
        const int64_t id = s_id[tid];
        int rank = 0;
        for (uint32_t j = 0; j < kk; ++j) {
            const float vj = s_val[j];
            const int64_t idj = s_id[j];
            rank += (vj > v || (vj == v && idj < id)) ? 1 : 0;
        }
        ids[(int64_t)q * k + rank] = id + id_base;
        out_scores[(int64_t)q * k + rank] = v;
    }
    for (int i = (int)kk + tid; i < k; i += blockDim.x) {
        ids[(int64_t)q * k + i] = -1;
        out_scores[(int64_t)q * k + i] = -1.0f;
    }
}

// merge of P per-shard lists per query; ties -> lower id; id < 0 = empty slot.
// Rank by counting, but only among the candidates that can still make it: a list with k valid entries proves that k
// candidates score at least its minimum, so everything below tau = max over the full lists of their minimum is out
// (50 chunk lists of 10 for a single query: ~15 candidates are ranked instead of 500, 44 -> ~6 us).
__global__ void __launch_bounds__(1024) k_topk_merge(const int64_t* __restrict__ in_ids, const float* __restrict__ in_scores, int P,
                                                    int Q, int k, int64_t ids_shard_stride, int64_t scores_shard_stride,
                                                    int64_t q_stride, int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    extern __shared__ unsigned char smraw[];
    __shared__ float s_wmax[32];
    __shared__ int s_m;
    const int M = P * k;
    int64_t* s_id = reinterpret_cast<int64_t*>(smraw);
    float* s_val = reinterpret_cast<float*>(s_id + M);
    int* s_sel = reinterpret_cast<int*>(s_val + M);
    const int q = blockIdx.x;
    for (int e = threadIdx.x; e < M; e += blockDim.x) {
        const int p = e / k, j = e - p * k;
        s_id[e] = in_ids[p * ids_shard_stride + (int64_t)q * q_stride + j];
        s_val[e] = in_scores[p * scores_shard_stride + (int64_t)q * q_stride + j];
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) { out_ids[(int64_t)q * k + j] = -1; out_scores[(int64_t)q * k + j] = -1.0f; }
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    float tau = -INFINITY;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        float mn = INFINITY;
        bool full = true;
        for (int j = 0; j < k; ++j) {
            if (s_id[p * k + j] < 0) { full = false; break; }
            mn = fminf(mn, s_val[p * k + j]);
        }
        if (full) tau = fmaxf(tau, mn);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tau = fmaxf(tau, __shfl_xor_sync(0xffffffffu, tau, o));
    if ((threadIdx.x & 31) == 0) s_wmax[threadIdx.x >> 5] = tau;
    __syncthreads();
    tau = -INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tau = fmaxf(tau, s_wmax[w]);
    for (int e = threadIdx.x; e < M; e += blockDim.x)
        if (s_id[e] >= 0 && !(s_val[e] < tau)) s_sel[atomicAdd(&s_m, 1)] = e;
    __syncthreads();
    const int m = s_m;
    for (int t = threadIdx.x; t < m; t += blockDim.x) {
        const int e = s_sel[t];
        const int64_t id = s_id[e];
        const float v = s_val[e];
        int rank = 0;
        for (int u = 0; u < m; ++u) {
            const int j = s_sel[u];
            const int64_t idj = s_id[j];
            const float vj = s_val[j];
            rank += (vj > v || (vj == v && (idj < id || (idj == id && j < e)))) ? 1 : 0;
        }
        if (rank < k) { out_ids[(int64_t)q * k + rank] = id; out_scores[(int64_t)q * k + rank] = v; }
    }
}

// ------------------------------------------------------------------------------------
// a11: core per-level similarity (core/search_engine.py:151-189), float64 like NumPy
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_core_level_sims(const double* __restrict__ cand, int64_t N, int S, int64_t cand_stride,
                                                         const double* __restrict__ q, const int32_t* __restrict__ q_start,
                                                         const int32_t* __restrict__ c_start, const int32_t* __restrict__ lvl_len,
                                                         int n_levels, double* __restrict__ sims) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < N * n_levels; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / n_levels;
        const int l = (int)(t - row * n_levels);
        const int m = lvl_len[l];
        const double* c = cand + row * cand_stride + c_start[l];
        const double* qq = q + q_start[l];
        double out = 0.0;
        if (m > 0) {
            double qs = 0, cs = 0;
            for (int j = 0; j < m; ++j) { qs += qq[j]; cs += c[j]; }
            const double qm = qs / m, cm = cs / m;
            double qv = 0, cv = 0, q2 = 0, c2 = 0, mse = 0;
            for (int j = 0; j < m; ++j) {
                const double dq = qq[j] - qm, dc = c[j] - cm;
                qv += dq * dq; cv += dc * dc;
                q2 += qq[j] * qq[j]; c2 += c[j] * c[j];
                const double e = qq[j] - c[j];
                mse += e * e;
            }
            const double qstd = sqrt(qv / m), cstd = sqrt(cv / m);
            if (qstd == 0.0 && cstd == 0.0) out = fabs(qm - cm) < 1e-6 ? 1.0 : 0.0;
            else if (qstd == 0.0 || cstd == 0.0) out = 0.1;
            else {
                double corr = 0;
                for (int j = 0; j < m; ++j) corr += ((qq[j] - qm) / qstd) * ((c[j] - cm) / cstd);
                corr /= m;
                const double sim = (corr + 1.0) / 2.0;
                const double mx = q2 / m + c2 / m;
                double dist = 1.0;
                if (mx > 0) { dist = 1.0 - (mse / m) / mx; if (dist < 0) dist = 0; }
                out = 0.7 * sim + 0.3 * dist;
                out = out < 0 ? 0 : (out > 1 ? 1 : out);
            }
        }
        sims[row * n_levels + l] = out;
    }
}

int grid_cap(int64_t blocks, int per_sm) {
    const int64_t cap = (int64_t)hq_cached_sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

bool layout_ok(const hq_index_layout* l) {
    if (!l || l->L < 1 || l->L > 8 || l->Lsum < 1) return false;
    for (int i = 0; i < l->L; ++i)
        if (l->lvl_off[i] < 0 || l->lvl_w[i] < 1 || l->lvl_off[i] + l->lvl_w[i] > l->Lsum || l->lvl_keff[i] < 1 ||
            l->lvl_keff[i] > l->lvl_w[i])
            return false;
    return true;
}

}  // namespace

extern "C" int hq_index_row_lengths(const float* idx, int64_t N, const hq_index_layout* layout, uint16_t* lens, void* stream) {
    HQ_REQUIRE(layout_ok(layout), "bad index layout");
    HQ_REQUIRE(N >= 0, "negative N");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(idx && lens, "null pointer");
    k_row_lengths<<<grid_cap((N * layout->L + 7) / 8, 16), 256, 0, (cudaStream_t)stream>>>(idx, N, *layout, lens);
    HQ_LAUNCH_OK("k_row_lengths");
    return HQ_OK;
}

extern "C" int hq_row_norms(const float* x, int64_t N, int64_t D, int64_t stride, float* norms, void* stream) {
    HQ_REQUIRE(N >= 0 && D > 0 && stride >= D, "bad shape");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(x && norms, "null pointer");
    k_row_norms<<<grid_cap((N + 7) / 8, 16), 256, 0, (cudaStream_t)stream>>>(x, N, D, stride, norms);
    HQ_LAUNCH_OK("k_row_norms");
    return HQ_OK;
}

extern "C" int hq_filter_level(const float* idx, const uint16_t* lens, int64_t N, const hq_index_layout* layout, int level,
                               const float* q_idx, const uint16_t* q_lens, int Q, const uint32_t* mask_in, int64_t mask_stride,
                               double thr, float* scores, int64_t scores_stride, uint32_t* mask_out, int32_t* n_alive,
                               int32_t* n_pass, void* stream) {
    HQ_REQUIRE(layout_ok(layout), "bad index layout");
    HQ_REQUIRE(level >= 0 && level < layout->L, "level %d out of range", level);
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(idx && lens && q_idx && q_lens && scores && mask_out && n_alive && n_pass, "null pointer");
    HQ_REQUIRE(mask_stride * 32 >= N && scores_stride >= N, "mask/scores stride too small");
    const int keff = layout->lvl_keff[level];
    HQ_REQUIRE(keff <= 64, "filter level wider than 64 values (%d) is not supported", keff);
    dim3 g((unsigned)((N + kFilterRows - 1) / kFilterRows), (unsigned)((Q + kFilterQT - 1) / kFilterQT));
    cudaStream_t st = (cudaStream_t)stream;
#define HQ_FL(KK)                                                                                                              \
    do {                                                                                                                       \
        const size_t smem = sizeof(float) * ((size_t)KK * kFilterQT + (size_t)kFilterQT * (KK + 1) + (size_t)kFilterRows * (KK + 1)); \
        if (smem > 48 * 1024)                                                                                                  \
            HQ_CUDA_OK(cudaFuncSetAttribute(k_filter_level<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        k_filter_level<KK><<<g, kFilterRows, smem, st>>>(idx, lens, N, *layout, level, keff, q_idx, q_lens, Q, mask_in,         \
                                                         mask_stride, thr, scores, scores_stride, mask_out, n_alive, n_pass);  \
    } while (0)
    if (keff <= 4) HQ_FL(4);
    else if (keff <= 8) HQ_FL(8);
    else if (keff <= 16) HQ_FL(16);
    else if (keff <= 24) HQ_FL(24);
    else if (keff <= 32) HQ_FL(32);
    else if (keff <= 48) HQ_FL(48);
    else HQ_FL(64);
#undef HQ_FL
    HQ_LAUNCH_OK("k_filter_level");
    return HQ_OK;
}

extern "C" int hq_filter_select(const float* scores, int64_t scores_stride, int64_t N, int Q, const int32_t* n_alive,
                                const int32_t* n_pass, double ratio, uint32_t* mask, int64_t mask_stride, int32_t* n_out,
                                void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(scores && n_alive && n_pass && mask && n_out, "null pointer");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    k_filter_select<<<Q, 1024, 0, (cudaStream_t)stream>>>(scores, scores_stride, N, n_alive, n_pass, ratio, mask, mask_stride, n_out);
    HQ_LAUNCH_OK("k_filter_select");
    return HQ_OK;
}

extern "C" int hq_filter_select_prev(const float* scores, int64_t scores_stride, int64_t N, int Q, const float* prev1, const float* prev0,
                                     int64_t prev_stride, const int32_t* n_alive, const int32_t* n_pass, double ratio, uint32_t* mask,
                                     int64_t mask_stride, int32_t* n_out, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(scores && n_alive && n_pass && mask && n_out, "null pointer");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    HQ_REQUIRE(!prev0 || prev1, "prev0 (level l - 2) needs prev1 (level l - 1)");
    HQ_REQUIRE(!prev1 || prev_stride >= N, "prev stride too small");
    k_filter_select_prev<<<Q, 1024, 0, (cudaStream_t)stream>>>(scores, scores_stride, N, prev1, prev0, prev_stride, n_alive, n_pass, ratio,
                                                               mask, mask_stride, n_out);
    HQ_LAUNCH_OK("k_filter_select_prev");
    return HQ_OK;
}

extern "C" int hq_gcut_hist(const float* scores, int64_t scores_stride, int64_t N, int Q, const uint32_t* mask, int64_t mask_stride,
                            const int32_t* need, const int32_t* d_hi, int32_t* hist, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(scores && mask && need && hist, "null pointer");
    int64_t chunks = (N + 256 * 64 - 1) / (256 * 64);
    if (chunks > 64) chunks = 64;
    k_gcut_hist<<<dim3((unsigned)chunks, (unsigned)Q), 256, 0, (cudaStream_t)stream>>>(scores, scores_stride, N, mask, mask_stride, need, d_hi, hist);
    HQ_LAUNCH_OK("k_gcut_hist");
    return HQ_OK;
}

extern "C" int hq_gcut_scan(const int32_t* hist, int Q, const int32_t* need, const int64_t* want, int64_t* d_out, int64_t* above_out,
                            void* stream) {
    if (Q <= 0) return HQ_OK;
    HQ_REQUIRE(hist && need && want && d_out && above_out, "null pointer");
    k_gcut_scan<<<Q, 1024, 0, (cudaStream_t)stream>>>(hist, need, want, d_out, above_out);
    HQ_LAUNCH_OK("k_gcut_scan");
    return HQ_OK;
}

extern "C" int hq_gcut_ties(const float* scores, int64_t scores_stride, int64_t N, int Q, const uint32_t* mask, int64_t mask_stride,
                            const int32_t* need, const int64_t* k_star, int64_t* ties, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(scores && mask && need && k_star && ties, "null pointer");
    int64_t chunks = (N + 256 * 64 - 1) / (256 * 64);
    if (chunks > 64) chunks = 64;
    k_gcut_ties<<<dim3((unsigned)chunks, (unsigned)Q), 256, 0, (cudaStream_t)stream>>>(scores, scores_stride, N, mask, mask_stride, need, k_star, ties);
    HQ_LAUNCH_OK("k_gcut_ties");
    return HQ_OK;
}

extern "C" int hq_gcut_apply(const float* scores, int64_t scores_stride, int64_t N, int Q, uint32_t* mask, int64_t mask_stride,
                             const int32_t* need, const int64_t* k_star, const int64_t* quota, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(scores && mask && need && k_star && quota, "null pointer");
    k_gcut_apply<<<Q, 1024, 0, (cudaStream_t)stream>>>(scores, scores_stride, N, mask, mask_stride, need, k_star, quota);
    HQ_LAUNCH_OK("k_gcut_apply");
    return HQ_OK;
}

extern "C" int hq_rerank_scores_f32(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride, const float* q,
                                    const float* q_norm, int Q, int64_t q_stride, const uint32_t* mask, int64_t mask_stride,
                                    float* scores, int64_t scores_stride, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0 && D > 0, "bad shape");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(db && db_norm && q && q_norm && scores, "null pointer");
    HQ_REQUIRE(db_stride >= D && q_stride >= D && scores_stride >= N, "stride too small");
    HQ_REQUIRE(!mask || mask_stride * 32 >= N, "mask stride too small");
    const int vec = (db_stride % 4 == 0) && (q_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
    dim3 g((unsigned)((N + kBN - 1) / kBN), (unsigned)((Q + kBM - 1) / kBM));
    k_rerank_scores<<<g, 256, 0, (cudaStream_t)stream>>>(db, db_norm, N, D, db_stride, q, q_norm, Q, q_stride, mask, mask_stride,
                                                         scores, scores_stride, vec);
    HQ_LAUNCH_OK("k_rerank_scores");
    return HQ_OK;
}

// paired rows: out[i] = (cos(a_i, b_i) + 1) / 2 with the score arithmetic of k_rerank_scores (0 when a norm is 0).  One warp
// per pair; the windows of _calculate_spatial_locality_similarity (rag/search/engine.py:662-714) are one launch.
__global__ void __launch_bounds__(256) k_paired_cosine01(const float* __restrict__ a, const float* __restrict__ na,
                                                         const float* __restrict__ b, const float* __restrict__ nb, int64_t n,
                                                         int64_t D, int64_t stride, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const float* pa = a + i * stride;
    const float* pb = b + i * stride;
    float acc = 0.f;
    for (int64_t k = lane; k < D; k += 32) acc = fmaf(__ldg(pa + k), __ldg(pb + k), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float x = __ldg(na + i), y = __ldg(nb + i);
        out[i] = (x != 0.f && y != 0.f) ? __fmul_rn(__fadd_rn(__fdiv_rn(acc, __fmul_rn(x, y)), 1.0f), 0.5f) : 0.f;
    }
}

extern "C" int hq_paired_cosine01(const float* a, const float* a_norm, const float* b, const float* b_norm, int64_t n, int64_t D,
                                  int64_t stride, float* out, void* stream) {
    HQ_REQUIRE(n >= 0 && D > 0 && stride >= D, "bad shape");
    if (n == 0) return HQ_OK;
    HQ_REQUIRE(a && a_norm && b && b_norm && out, "null pointer");
    k_paired_cosine01<<<(unsigned)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, a_norm, b, b_norm, n, D, stride, out);
    HQ_LAUNCH_OK("k_paired_cosine01");
    return HQ_OK;
}

extern "C" int hq_rerank_scores_sparse_f32(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride, const float* q,
                                           const float* q_norm, int Q, int64_t q_stride, const uint32_t* mask, int64_t mask_stride,
                                           float* scores, int64_t scores_stride, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0 && D > 0, "bad shape");
    if (N == 0 || Q == 0) return HQ_OK;
    HQ_REQUIRE(db && db_norm && q && q_norm && scores, "null pointer");
    HQ_REQUIRE(db_stride >= D && q_stride >= D && scores_stride >= N, "stride too small");
    HQ_REQUIRE(!mask || mask_stride * 32 >= N, "mask stride too small");
    const int vec = (D % 4 == 0) && (db_stride % 4 == 0) && (q_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
    const int64_t tasks = (int64_t)Q * ((N + 31) / 32);
    k_rerank_sparse<<<grid_cap((tasks + 7) / 8, 16), 256, 0, (cudaStream_t)stream>>>(db, db_norm, N, D, db_stride, q, q_norm, Q, q_stride,
                                                                                   mask, mask_stride, scores, scores_stride, vec);
    HQ_LAUNCH_OK("k_rerank_sparse");
    return HQ_OK;
}

extern "C" int hq_rerank_sparse_topk_supported(int64_t D, int64_t db_stride, int64_t q_stride, int k) {
    return (D % 4 == 0 && db_stride % 4 == 0 && q_stride % 4 == 0 && D <= 8192 && k >= 1 && k <= kSparseTopkMaxK) ? 1 : 0;
}

static int sparse_topk_ctas(int Q) {
    const int sms = hq_cached_sm_count();          // one wave of 512-thread CTAs (one per SM), shared by the queries
    return Q <= 1 ? sms : (sms + Q - 1) / Q;
}

extern "C" int64_t hq_rerank_sparse_topk_scratch_bytes(int Q, int k) {
    if (Q <= 0 || k <= 0) return 0;
    const int kl = k > 32 ? k : 32;
    return (int64_t)Q * sparse_topk_ctas(Q) * kl * 12 + (int64_t)Q * 12 + 64;
}

namespace {
struct SparseGeom { int S, NP, NC, R; size_t smem; };

// shared memory of a launch: the query + S row slots (rows padded to 128 bytes), as many as fit 200 KB, one producer warp
// and up to 15 consumer warps with 1 / 2 / 4 slots each.  (2 / 3 / 4 producer warps with 14 / 12 / 12 consumers: 93-103 us
// for a query against 1 M x 1536 where this split takes 85-100 us -- 57 K scattered 6 KB rows arrive at ~4 TB/s either way;
// the bf16 variant with its 3 KB rows: 81-91 us with 1 to 4 producer warps, the fetch costs per ROW, not per byte.)
bool sparse_geom(int64_t D, bool b16, SparseGeom& g) {
    const size_t qb = ((size_t)D * 4 + 127) & ~(size_t)127;
    const size_t buf = b16 ? (((size_t)D * 2 + 127) & ~(size_t)127) : qb;
    const size_t budget = 200 * 1024;
    if (qb + buf > budget) return false;
    int S = (int)((budget - qb) / buf);
    if (S > kSparseTopkSlots) S = kSparseTopkSlots;
    if (S < 1) return false;
    g.NP = 1;
    g.NC = S < kSparseTopkThreads / 32 - 1 ? S : kSparseTopkThreads / 32 - 1;
    const int depth = S / g.NC >= 4 ? 4 : (S / g.NC >= 2 ? 2 : 1);                 // slots per consumer: a power of two
    g.S = depth * g.NC;
    g.smem = qb + (size_t)g.S * buf;
    g.R = 0;
    if (b16) {                                      // the last CTA's exact re-score: candidates + R fp32 rows behind the query
        const size_t cand = ((size_t)kSparseTopkSelB * 12 + 127) & ~(size_t)127;
        int R = (int)((budget - qb - cand) / qb);
        if (R > kSparseTopkR) R = kSparseTopkR;
        g.R = R;
        const size_t need = qb + cand + (size_t)R * qb;
        if (need > g.smem) g.smem = need;
    } else if (g.smem < (size_t)kSparseTopkSel * 12) {
        g.smem = (size_t)kSparseTopkSel * 12;
    }
    return true;
}
}  // namespace

// db_unit_bf16 (may be NULL): the shard's unit rows in bf16 (hq_to_bf16_unit) with dc_max = max |c16 - c / |c|| over the rows
// (hq_bf16_unit_error_max): phase one then reads half the bytes, the best 32 rows are re-scored exactly and a query whose
// result that does not prove is redone by the fp32 pass (second launch, a no-op otherwise).  Same ids and scores either way.
extern "C" int hq_rerank_sparse_topk(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride,
                                     const void* db_unit_bf16, int64_t db_pitch, float dc_max, const float* q, const float* q_norm,
                                     int Q, int64_t q_stride, const uint32_t* mask, int64_t mask_stride, int k, int64_t id_base,
                                     int64_t* ids, float* scores, void* scratch, int64_t scratch_bytes, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0 && D > 0, "bad shape");
    if (Q == 0) return HQ_OK;
    HQ_REQUIRE(db && db_norm && q && q_norm && mask && ids && scores, "null pointer");
    HQ_REQUIRE(hq_rerank_sparse_topk_supported(D, db_stride, q_stride, k), "needs D %% 4 == 0, D <= 8192, strides %% 4 == 0 and k <= %d",
               kSparseTopkMaxK);
    HQ_REQUIRE((reinterpret_cast<uintptr_t>(db) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0, "rows must be 16-byte aligned");
    HQ_REQUIRE(db_stride >= D && q_stride >= D && mask_stride * 32 >= N, "stride too small");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    HQ_REQUIRE(scratch && scratch_bytes >= hq_rerank_sparse_topk_scratch_bytes(Q, k), "scratch too small");
    const int P = sparse_topk_ctas(Q);
    const int kl = k > 32 ? k : 32;
    int64_t* p_ids = reinterpret_cast<int64_t*>(scratch);                        // [Q][P][kl]
    float* p_sc = reinterpret_cast<float*>(p_ids + (int64_t)Q * P * kl);
    uint32_t* counters = reinterpret_cast<uint32_t*>(p_sc + (((int64_t)Q * P * kl + 3) & ~(int64_t)3));   // [2 Q] tickets, done
    uint32_t* flags = counters + 2 * Q;                                                                   // [Q]
    cudaStream_t st = (cudaStream_t)stream;
    HQ_CUDA_OK(cudaMemsetAsync(counters, 0, (size_t)Q * 12, st));
    static bool attr = false;
    if (!attr) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_rerank_sparse_topk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        HQ_CUDA_OK(cudaFuncSetAttribute(k_rerank_sparse_topk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        attr = true;
    }
    SparseGeom g32{}, g16{};
    HQ_REQUIRE(sparse_geom(D, false, g32), "D too large");
    const __nv_bfloat16* b16 = reinterpret_cast<const __nv_bfloat16*>(db_unit_bf16);
    const bool use16 = b16 && D % 8 == 0 && db_pitch % 8 == 0 && db_pitch >= D && (reinterpret_cast<uintptr_t>(b16) & 15) == 0 &&
                       dc_max >= 0.f && dc_max < 1.f && sparse_geom(D, true, g16) && g16.R >= k + 8;
    if (use16) {
        k_rerank_sparse_topk<true><<<dim3(P, Q), (g16.NP + g16.NC) * 32, g16.smem, st>>>(
            db, db_norm, N, D, db_stride, b16, db_pitch, dc_max, q, q_norm, Q, q_stride, mask, mask_stride, k, id_base, p_ids, p_sc,
            counters, flags, 0, ids, scores, g16.S, g16.NP, g16.R);
        HQ_LAUNCH_OK("k_rerank_sparse_topk<bf16>");
    }
    k_rerank_sparse_topk<false><<<dim3(P, Q), (g32.NP + g32.NC) * 32, g32.smem, st>>>(
        db, db_norm, N, D, db_stride, nullptr, 0, 0.f, q, q_norm, Q, q_stride, mask, mask_stride, k, id_base, p_ids, p_sc, counters,
        flags, use16 ? 1 : 0, ids, scores, g32.S, g32.NP, 0);
    HQ_LAUNCH_OK("k_rerank_sparse_topk");
    return HQ_OK;
}

extern "C" int hq_topk_from_scores(const float* scores, int64_t scores_stride, int64_t N, int Q, int k, int64_t id_base,
                                   int64_t* ids, float* out_scores, void* stream) {
    HQ_REQUIRE(k >= 1 && k <= kMaxK, "k must be in [1, %d]", kMaxK);
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (Q == 0) return HQ_OK;
    HQ_REQUIRE(ids && out_scores && (scores || N == 0), "null pointer");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    k_topk_scores<<<Q, 1024, 0, (cudaStream_t)stream>>>(scores, scores_stride, N, k, id_base, ids, out_scores, 1, N);
    HQ_LAUNCH_OK("k_topk_scores");
    return HQ_OK;
}

// chunks per query of the two-level top-k (1 = single level): enough CTAs to fill the GPU, chunks of >= 4096 scores
static int topk_parts(int64_t N, int Q, int k) {
    int64_t parts = (2 * (int64_t)hq_cached_sm_count() + Q - 1) / Q;
    if (parts > 256) parts = 256;
    if (parts > N / 4096) parts = N / 4096;
    // the merge ranks the candidates that survive its tau bound against each other: a handful on ordinary data, but all
    // parts * k of them when every score ties (O(M^2) on one CTA: 244 x 10 candidates took 1.8 ms), so M stays bounded
    if (parts * (int64_t)k > 1024) parts = 1024 / k;
    return parts < 2 ? 1 : (int)parts;
}

extern "C" int64_t hq_topk_chunked_scratch_bytes(int64_t N, int Q, int k) {
    if (N <= 0 || Q <= 0 || k <= 0) return 0;
    return (int64_t)Q * topk_parts(N, Q, k) * k * 12;
}

// Same result as hq_topk_from_scores for FEW queries over LONG rows: two levels (chunk top-k, then merge).
extern "C" int hq_topk_from_scores_chunked(const float* scores, int64_t scores_stride, int64_t N, int Q, int k, int64_t id_base,
                                           int64_t* ids, float* out_scores, void* scratch, int64_t scratch_bytes, void* stream) {
    HQ_REQUIRE(k >= 1 && k <= kMaxK, "k must be in [1, %d]", kMaxK);
    HQ_REQUIRE(N >= 0 && Q >= 0, "negative size");
    if (Q == 0) return HQ_OK;
    HQ_REQUIRE(ids && out_scores && (scores || N == 0), "null pointer");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    const int parts = topk_parts(N, Q, k);
    if (parts == 1) return hq_topk_from_scores(scores, scores_stride, N, Q, k, id_base, ids, out_scores, stream);
    HQ_REQUIRE(scratch && scratch_bytes >= hq_topk_chunked_scratch_bytes(N, Q, k), "scratch too small");
    const int64_t chunk = ((N + parts - 1) / parts + 31) & ~(int64_t)31;        // trailing chunks may be empty
    int64_t* p_ids = reinterpret_cast<int64_t*>(scratch);                       // [Q][parts][k]
    float* p_sc = reinterpret_cast<float*>(p_ids + (int64_t)Q * parts * k);
    k_topk_scores<<<Q * parts, 1024, 0, (cudaStream_t)stream>>>(scores, scores_stride, chunk, k, id_base, p_ids, p_sc, parts, N);
    HQ_LAUNCH_OK("k_topk_scores");
    // "shard" p of query q starts at p * k + q * parts * k
    k_topk_merge<<<Q, parts * k > 256 ? 1024 : 256, (size_t)parts * k * 16, (cudaStream_t)stream>>>(p_ids, p_sc, parts, Q, k, (int64_t)k, (int64_t)k,
                                                                          (int64_t)parts * k, ids, out_scores);
    HQ_LAUNCH_OK("k_topk_merge");
    return HQ_OK;
}

extern "C" int64_t hq_rerank_scratch_bytes(int64_t N, int Q) { return N * (int64_t)Q * 4; }

extern "C" int hq_rerank_topk_f32(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride, const float* q,
                                  const float* q_norm, int Q, int64_t q_stride, const uint32_t* mask, int64_t mask_stride, int k,
                                  int64_t id_base, int64_t* ids, float* scores, void* scratch, int64_t scratch_bytes, void* stream) {
    HQ_REQUIRE(scratch && scratch_bytes >= hq_rerank_scratch_bytes(N, Q), "scratch too small");
    int rc = hq_rerank_scores_f32(db, db_norm, N, D, db_stride, q, q_norm, Q, q_stride, mask, mask_stride, (float*)scratch, N, stream);
    if (rc != HQ_OK) return rc;
    return hq_topk_from_scores((const float*)scratch, N, N, Q, k, id_base, ids, scores, stream);
}

extern "C" int hq_topk_merge_strided(const int64_t* in_ids, const float* in_scores, int P, int Q, int k, int64_t ids_shard_stride,
                                     int64_t scores_shard_stride, int64_t* out_ids, float* out_scores, void* stream) {
    HQ_REQUIRE(P >= 1 && Q >= 0 && k >= 1, "bad shape");
    if (Q == 0) return HQ_OK;
    HQ_REQUIRE(in_ids && in_scores && out_ids && out_scores, "null pointer");
    HQ_REQUIRE(ids_shard_stride >= (int64_t)Q * k && scores_shard_stride >= (int64_t)Q * k, "shard stride smaller than Q * k");
    const size_t smem = (size_t)P * k * 16;
    HQ_REQUIRE(smem <= 48 * 1024, "P*k too large for the merge kernel");
    k_topk_merge<<<Q, P * k > 256 ? 1024 : 256, smem, (cudaStream_t)stream>>>(in_ids, in_scores, P, Q, k, ids_shard_stride, scores_shard_stride, (int64_t)k,
                                                         out_ids, out_scores);
    HQ_LAUNCH_OK("k_topk_merge");
    return HQ_OK;
}

extern "C" int hq_topk_merge(const int64_t* in_ids, const float* in_scores, int P, int Q, int k, int64_t* out_ids, float* out_scores,
                             void* stream) {
    return hq_topk_merge_strided(in_ids, in_scores, P, Q, k, (int64_t)Q * k, (int64_t)Q * k, out_ids, out_scores, stream);
}

extern "C" int hq_core_level_sims(const double* cand, int64_t N, int S, int64_t cand_stride, const double* q, const int32_t* q_start,
                                  const int32_t* c_start, const int32_t* lvl_len, int n_levels, double* sims, void* stream) {
    HQ_REQUIRE(N >= 0 && S > 0 && n_levels >= 0 && cand_stride >= S, "bad shape");
    if (N == 0 || n_levels == 0) return HQ_OK;
    HQ_REQUIRE(cand && q && q_start && c_start && lvl_len && sims, "null pointer");
    k_core_level_sims<<<grid_cap((N * n_levels + 255) / 256, 16), 256, 0, (cudaStream_t)stream>>>(cand, N, S, cand_stride, q, q_start,
                                                                                                 c_start, lvl_len, n_levels, sims);
    HQ_LAUNCH_OK("k_core_level_sims");
    return HQ_OK;
}
