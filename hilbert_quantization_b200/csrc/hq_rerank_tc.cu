// K6: tensor-core rerank.  scores = Q x N x D cosine contraction on tcgen05 (bf16 in, fp32
// accumulate in TMEM) with the survivor mask and a streaming per-query top-k' fused into the
// epilogue, followed by an exact fp32 re-score of the k' shortlisted rows.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, 4 stages)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M=128 queries, N=256 rows, K=16)
//   warps 2..   : epilogue, EH threads per query row (EH = 1: four warps, 256 columns each; EH = 2: eight warps,
//                 128 columns each, for short rows whose MMAs finish before a four-warp epilogue does):
//                 tcgen05.ld 32 / EH columns at a time, apply mask / 1/|c|, keep the best KP (value, row) pairs
//                 in registers
// Work unit = (query tile of 128, contiguous range of 256-row tiles); units are laid out so
// that neighbouring CTAs stream the same database rows at the same time (L2 reuse), the
// query tile comes from L2.  Two TMEM accumulators (2 x 256 columns) overlap the epilogue of
// tile i with the MMAs of tile i+1.
#include "hq_tc.cuh"
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int UMMA_K = 16;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;      // 16 KB
constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;      // 32 KB
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr uint32_t TMEM_COLS = 512;

using namespace hq_tc;

constexpr uint32_t kIdesc = make_idesc(1 /*bf16*/, BM, BN);

struct TcParams {
    int64_t N;
    int Q, D;
    int m_tiles, n_tiles, n_ranges, tiles_per_range, num_units;
    const float* db_norm;
    const uint32_t* mask;
    int64_t mask_stride;
    int eh;               // epilogue threads per query row (1 or 2): partial lists per unit
    const int32_t* zero_rows;   // unit-row operand: ids of the zero-norm rows (ascending), appended by the merge when fewer
    int n_zero;                 // than k other survivors exist (they score exactly 0.0, the lowest possible score)
    float* part_val;      // [num_units][eh][128][KP]
    int32_t* part_idx;
};

// PS: the database operand holds unit rows (c / |c| rounded to bf16, hq_to_bf16_unit).  The accumulator then orders a
// query's rows like the cosine and the epilogue needs no 1/|c| per column (a shared-memory read, a select and a
// multiply: half of its instructions); zero-norm rows are cleared from the mask instead.
template <int KP, int EH, bool PS>
__global__ void __launch_bounds__(64 + 128 * EH, 1) k_rerank_tc(const __grid_constant__ CUtensorMap map_q,
                                                            const __grid_constant__ CUtensorMap map_db, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 128B-swizzled tiles need 1024-byte alignment in the shared window
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    float* s_inv = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);          // [2][BN]
    float* s_vals = s_inv + 2 * BN;                                                // [32 / EH][128 * EH] spill column per epilogue thread
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_vals + 32 * 128);             // [8 / EH][128 * EH]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_mask + 8 * 128);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k_blocks = (p.D + BK - 1) / BK;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_db)) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128 * EH); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
                const int m_tile = u % p.m_tiles, range = u / p.m_tiles;
                const int t0 = range * p.tiles_per_range;
                const int t1 = min(p.n_tiles, t0 + p.tiles_per_range);
                for (int t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(&map_q, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BK, m_tile * BM);
                        tma_load_2d(&map_db, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, kb * BK, t * BN);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        uint32_t stage = 0, phase = 0, it = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const int range = u / p.m_tiles;
            const int t0 = range * p.tiles_per_range;
            const int t1 = min(p.n_tiles, t0 + p.tiles_per_range);
            for (int t = t0; t < t1; ++t, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
                        const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * B_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_f16(tmem_base + acc * BN, da + (uint64_t)(k * UMMA_K * 2 >> 4), db + (uint64_t)(k * UMMA_K * 2 >> 4),
                                      kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);                       // frees the smem slot when the MMAs retire
                        if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]); // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= epilogue: streaming top-KP per query row =================
        // Code size matters here (the first version unrolled 256 insertion sites into 1.3 MB
        // of SASS and ran at 6 % tensor utilisation out of the instruction cache): the column
        // loop is rolled, each 32-column chunk computes a hit mask against the thread's current
        // k'-th best, and only chunks with hits spill their values to a private shared-memory
        // column and walk the set bits through ONE insertion site.
        constexpr int EPI = 128 * EH;                    // epilogue threads
        constexpr int CW = 32 / EH;                      // columns per chunk (the spill area holds CW x EPI values)
        constexpr int COLS = BN / EH;                    // columns per thread and tile
        constexpr int MW = COLS / 32;                    // mask words per thread and tile
        const int ew = warp & 3;                         // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                // which COLS columns of the tile (0 when EH == 1)
        const int col0 = half * COLS;
        const int row_in_tile = ew * 32 + lane;
        const int et = (warp - 2) * 32 + lane;           // index among the epilogue threads
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const int m_tile = u % p.m_tiles, range = u / p.m_tiles;
            const int t0 = range * p.tiles_per_range;
            const int t1 = min(p.n_tiles, t0 + p.tiles_per_range);
            const int q = m_tile * BM + row_in_tile;
            const bool q_ok = q < p.Q;
            float bv[KP];
            int32_t bi[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) { bv[j] = -FLT_MAX; bi[j] = -1; }
            float thr = -FLT_MAX;
            // The mask words (thread-private) and row norms of a tile are fetched one tile ahead: issued right after
            // the barrier, they travel while this tile's columns are processed (loading them at the top of the tile
            // put a global-load latency plus a barrier in front of every epilogue: 768-D ran at 60 % of the tensor peak).
            static_assert(BN % EPI == 0 && BN / EPI <= 2, "row norms: at most two per thread");
            uint32_t m_next[MW];
            float n_next[BN / EPI];
            auto fetch = [&](int t) {
#pragma unroll
                for (int h = 0; h < BN / EPI; ++h) {
                    const int64_t r = (int64_t)t * BN + et + EPI * h;
                    n_next[h] = r < p.N ? __ldg(p.db_norm + r) : 0.f;
                }
#pragma unroll
                for (int w = 0; w < MW; ++w) {
                    // the loaded word is not touched here (the tail of the last word is cut where it is consumed):
                    // any use would wait for the load right away
                    const int64_t r0 = (int64_t)t * BN + col0 + 32 * w;
                    const bool in = q_ok && r0 < p.N;
                    uint32_t m = in ? 0xffffffffu : 0u;
                    if (in && p.mask) m = __ldg(p.mask + (int64_t)q * p.mask_stride + (r0 >> 5));
                    m_next[w] = m;
                }
            };
            if (t0 < t1) fetch(t0);
            for (int t = t0; t < t1; ++t, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                // 1/|c| of this tile's rows (0 marks a zero-norm or out-of-range row)
                float* inv = s_inv + acc * BN;
                uint32_t* nzw = reinterpret_cast<uint32_t*>(inv);       // PS: bit r = row r of the tile has a non-zero norm
#pragma unroll
                for (int h = 0; h < BN / EPI; ++h) {
                    if constexpr (PS) {
                        const uint32_t b = __ballot_sync(0xffffffffu, n_next[h] > 0.f);
                        if (lane == 0) nzw[((et + EPI * h) >> 5)] = b;
                    } else {
                        inv[et + EPI * h] = n_next[h] > 0.f ? 1.0f / n_next[h] : 0.f;
                    }
                }
#pragma unroll
                for (int w = 0; w < MW; ++w) {                                       // thread-private slots, read back per chunk
                    const int64_t r0 = (int64_t)t * BN + col0 + 32 * w;
                    uint32_t m = m_next[w];
                    if (r0 + 32 > p.N) m = r0 < p.N ? (m & ((1u << (uint32_t)(p.N - r0)) - 1u)) : 0u;
                    s_mask[w * EPI + et] = m;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
                if (t + 1 < t1) fetch(t + 1);
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < COLS / CW; ++c) {
                    uint32_t r[CW];
                    tmem_ld(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + col0 + CW * c, r);
                    tmem_ld_wait();
                    uint32_t m = s_mask[((CW * c) >> 5) * EPI + et];
                    if constexpr (CW < 32) m = (m >> ((CW * c) & 31)) & ((1u << CW) - 1u);
                    uint32_t hits = 0;
                    if constexpr (PS) {
                        uint32_t nz = nzw[(col0 + CW * c) >> 5];
                        if constexpr (CW < 32) nz = (nz >> ((CW * c) & 31)) & ((1u << CW) - 1u);
                        m &= nz;
#pragma unroll
                        for (int j = 0; j < CW; ++j) hits |= (__uint_as_float(r[j]) > thr ? 1u : 0u) << j;
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            const float iv = inv[col0 + CW * c + j];
                            const float v = iv > 0.f ? __uint_as_float(r[j]) * iv : -FLT_MAX * 0.5f;
                            r[j] = __float_as_uint(v);
                            hits |= (v > thr ? 1u : 0u) << j;
                        }
                    }
                    hits &= m;
                    if (hits) {
#pragma unroll
                        for (int j = 0; j < CW; ++j) s_vals[j * EPI + et] = __uint_as_float(r[j]);
                        while (hits) {
                            const int j = __ffs(hits) - 1;
                            hits &= hits - 1;
                            const float v = s_vals[j * EPI + et];
                            if (v > bv[KP - 1]) {
                                bv[KP - 1] = v;
                                bi[KP - 1] = t * BN + col0 + CW * c + j;
#pragma unroll
                                for (int s = KP - 1; s > 0; --s) {
                                    if (bv[s] > bv[s - 1]) {
                                        const float tv = bv[s]; bv[s] = bv[s - 1]; bv[s - 1] = tv;
                                        const int32_t ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
                                    }
                                }
                            }
                        }
                        thr = bv[KP - 1];
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
            }
            if (q_ok) {
                float* pv = p.part_val + (((int64_t)u * EH + half) * BM + row_in_tile) * KP;
                int32_t* pi = p.part_idx + (((int64_t)u * EH + half) * BM + row_in_tile) * KP;
#pragma unroll
                for (int j = 0; j < KP; ++j) { pv[j] = bv[j]; pi[j] = bi[j]; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// merge the per-unit shortlists of one query, re-score the best KP exactly in fp32, emit top-k
template <int KP>
__global__ void __launch_bounds__(128) k_rerank_tc_merge(const TcParams p, const float* __restrict__ db_f32, int64_t db_stride,
                                                         const float* __restrict__ q_f32, int64_t q_stride,
                                                         const float* __restrict__ q_norm, int k, int64_t id_base,
                                                         int64_t* __restrict__ ids, float* __restrict__ scores) {
    extern __shared__ unsigned char sm[];
    const int M = p.n_ranges * p.eh * KP;
    float* c_val = reinterpret_cast<float*>(sm);
    int32_t* c_idx = reinterpret_cast<int32_t*>(c_val + M);
    __shared__ float top_val[KP];
    __shared__ int32_t top_idx[KP];
    __shared__ float ex_val[KP];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m_tile = q / BM, row = q % BM;
    for (int e = tid; e < M; e += blockDim.x) {
        const int r = e / KP, j = e - r * KP;                   // r = range * eh + half
        const int64_t u = (int64_t)(r / p.eh) * p.m_tiles + m_tile;
        const int64_t a = (((u * p.eh + r % p.eh) * BM + row) * KP) + j;
        c_val[e] = p.part_val[a];
        c_idx[e] = p.part_idx[a];
    }
    if (tid < KP) { top_val[tid] = -FLT_MAX; top_idx[tid] = -1; }
    __syncthreads();
    // Shortlist = best KP of the M candidates (value desc, id asc), in two levels so that no round needs a block-wide
    // barrier: every warp extracts the best KP of its own quarter of the candidates (KP rounds of a warp arg-best over
    // shared memory, __syncwarp only), then warp 0 merges the 4 x KP winners held two per lane in registers.  (The
    // first version ranked every candidate against every other one, O(M^2): 0.6 ms of a 1.0 ms single-query search;
    // the second ran KP block-wide rounds with two __syncthreads and a serial pick each: 60 us per 1024-query batch.)
    __shared__ float w_val[4 * KP];
    __shared__ int32_t w_idx[4 * KP];
    {
        const int per = (M + 3) / 4, e0 = warp * per, e1 = min(M, e0 + per);
        for (int round = 0; round < KP; ++round) {
            float bv = -FLT_MAX;
            int32_t bi = 0x7fffffff, bp = -1;
            for (int e = e0 + lane; e < e1; e += 32) {
                const int32_t id = c_idx[e];
                if (id < 0) continue;
                const float v = c_val[e];
                if (bp < 0 || v > bv || (v == bv && id < bi)) { bv = v; bi = id; bp = e; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o), op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (op >= 0 && (bp < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; bp = op; }
            }
            if (lane == 0) {
                w_val[warp * KP + round] = bp >= 0 ? bv : -FLT_MAX;
                w_idx[warp * KP + round] = bp >= 0 ? bi : -1;
                if (bp >= 0) c_idx[bp] = -1;                   // taken
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int PER = (4 * KP + 31) / 32;
        float mv[PER];
        int32_t mi[PER];
#pragma unroll
        for (int s = 0; s < PER; ++s) {
            const int e = lane + 32 * s;
            mv[s] = e < 4 * KP ? w_val[e] : -FLT_MAX;
            mi[s] = e < 4 * KP ? w_idx[e] : -1;
        }
        for (int round = 0; round < KP; ++round) {
            float bv = -FLT_MAX;
            int32_t bi = -1, bs = -1;
#pragma unroll
            for (int s = 0; s < PER; ++s)
                if (mi[s] >= 0 && (bi < 0 || mv[s] > bv || (mv[s] == bv && mi[s] < bi))) { bv = mv[s]; bi = mi[s]; bs = s; }
            int bl = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
                if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && (oi < bi || (oi == bi && ol < bl))))) { bv = ov; bi = oi; bl = ol; }
            }
            if (bi >= 0 && bl == lane) {
#pragma unroll
                for (int s = 0; s < PER; ++s)
                    if (s == bs) mi[s] = -1;
            }
            if (lane == 0) { top_val[round] = bi >= 0 ? bv : -FLT_MAX; top_idx[round] = bi; }
        }
    }
    __syncthreads();
    // exact fp32 cosine of the shortlisted rows (one warp per candidate)
    const float nq = q_norm[q];
    const float* qv = q_f32 + (int64_t)q * q_stride;
    const bool vec = (p.D % 4 == 0) && (db_stride % 4 == 0) && (q_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(db_f32) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(q_f32) & 15) == 0);
    for (int c = warp; c < KP; c += 4) {
        const int32_t id = top_idx[c];
        float s = -1.0f;
        if (id >= 0) {
            const float* rv = db_f32 + (int64_t)id * db_stride;
            float acc = 0.f;
            if (vec) {
                for (int i = lane; i < p.D / 4; i += 32) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(qv) + i);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(rv) + i);
                    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
                }
            } else {
                for (int i = lane; i < p.D; i += 32) acc = fmaf(__ldg(qv + i), __ldg(rv + i), acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            const float nc = __ldg(p.db_norm + id);
            s = 0.f;
            if (nq != 0.f && nc != 0.f) s = __fmul_rn(__fadd_rn(__fdiv_rn(acc, __fmul_rn(nq, nc)), 1.0f), 0.5f);
        }
        if (lane == 0) ex_val[c] = s;
    }
    __syncthreads();
    if (tid < KP) {
        const int32_t id = top_idx[tid];
        const float v = ex_val[tid];
        if (id >= 0) {
            int rank = 0;
            for (int j = 0; j < KP; ++j) {
                const int32_t idj = top_idx[j];
                if (idj < 0) continue;
                rank += (ex_val[j] > v || (ex_val[j] == v && idj < id)) ? 1 : 0;
            }
            if (rank < k) { ids[(int64_t)q * k + rank] = (int64_t)id + id_base; scores[(int64_t)q * k + rank] = v; }
        }
    }
    // fewer than k survivors: fill the tail
    if (tid == 0) {
        int cnt = 0;
        for (int j = 0; j < KP; ++j) cnt += top_idx[j] >= 0 ? 1 : 0;
        // unit-row operand: zero-norm rows never reach the shortlists; they score 0.0 and follow everybody else
        for (int z = 0; z < p.n_zero && cnt < k; ++z) {
            const int32_t id = p.zero_rows[z];
            if (p.mask && !((p.mask[(int64_t)q * p.mask_stride + (id >> 5)] >> (id & 31)) & 1u)) continue;
            ids[(int64_t)q * k + cnt] = (int64_t)id + id_base;
            scores[(int64_t)q * k + cnt] = 0.f;
            ++cnt;
        }
        for (int j = cnt; j < k; ++j) { ids[(int64_t)q * k + j] = -1; scores[(int64_t)q * k + j] = -1.0f; }
    }
}

__global__ void __launch_bounds__(256) k_to_bf16(const float* __restrict__ src, int64_t N, int64_t D, int64_t src_stride,
                                                 __nv_bfloat16* __restrict__ dst, int64_t dst_pitch) {
    const int64_t total = N * dst_pitch;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / dst_pitch, c = i - r * dst_pitch;
        dst[i] = __float2bfloat16_rn(c < D ? __ldg(src + r * src_stride + c) : 0.f);
    }
}

__global__ void __launch_bounds__(256) k_to_bf16_unit(const float* __restrict__ src, int64_t N, int64_t D, int64_t src_stride,
                                                      const float* __restrict__ norms, __nv_bfloat16* __restrict__ dst, int64_t dst_pitch) {
    const int64_t total = N * dst_pitch;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / dst_pitch, c = i - r * dst_pitch;
        const float n = __ldg(norms + r);
        dst[i] = __float2bfloat16_rn((c < D && n > 0.f) ? __fdiv_rn(__ldg(src + r * src_stride + c), n) : 0.f);
    }
}

// Database build: eight elements per thread (two 128-bit loads, one 128-bit store), one 64-bit division per chunk instead of
// one per element (the scalar kernel above ran at 1.7 TB/s: 5.5 ms per 1 M x 1536 rows).  Same arithmetic: IEEE division by the
// row norm, round to nearest even.  Needs 16-byte aligned rows on both sides.
__global__ void __launch_bounds__(256) k_to_bf16_unit_vec(const float* __restrict__ src, int64_t N, int64_t D, int64_t src_stride,
                                                          const float* __restrict__ norms, __nv_bfloat16* __restrict__ dst, int64_t dst_pitch) {
    const int64_t cpr = dst_pitch >> 3, total = N * cpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cpr, c = (i - r * cpr) << 3;
        const float n = __ldg(norms + r);
        const float* s = src + r * src_stride + c;
        float v[8];
        if (c + 8 <= D) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(s)), b = __ldcs(reinterpret_cast<const float4*>(s) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = c + j < D ? __ldg(s + j) : 0.f;
        }
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lo = n > 0.f ? __fdiv_rn(v[2 * j], n) : 0.f, hi = n > 0.f ? __fdiv_rn(v[2 * j + 1], n) : 0.f;
            w[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
        }
        *reinterpret_cast<uint4*>(dst + r * dst_pitch + c) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows) {
    return make_map_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, pitch_elems, BK, box_rows);
}

void plan_units(int64_t N, int Q, int sms, TcParams& p) {
    p.m_tiles = (Q + BM - 1) / BM;
    p.n_tiles = (int)((N + BN - 1) / BN);
    int best_s = 1;
    double best_eff = -1.0;
    for (int waves = 1; waves <= 4; ++waves) {
        int s = (waves * sms) / p.m_tiles;
        if (s < 1) s = 1;
        if (s > p.n_tiles) s = p.n_tiles;
        const int tiles_per = (p.n_tiles + s - 1) / s;
        s = (p.n_tiles + tiles_per - 1) / tiles_per;                 // ranges actually needed
        const int units = s * p.m_tiles;
        const int rounds = (units + sms - 1) / sms;
        const double eff = (double)p.n_tiles * p.m_tiles / ((double)rounds * tiles_per * sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_s = s; }
    }
    p.n_ranges = best_s;
    p.tiles_per_range = (p.n_tiles + best_s - 1) / best_s;
    p.n_ranges = (p.n_tiles + p.tiles_per_range - 1) / p.tiles_per_range;
    p.num_units = p.n_ranges * p.m_tiles;
}

int pick_kp(int k) { return k <= 10 ? 16 : (k <= 20 ? 32 : 0); }

template <int KP, int EH, bool PS>
int launch_tc(const CUtensorMap& mq, const CUtensorMap& mdb, const TcParams& p, const float* db_f32, int64_t db_stride,
              const float* q_f32, int64_t q_stride, const float* q_norm, int k, int64_t id_base, int64_t* ids, float* scores,
              cudaStream_t st) {
    const size_t smem = STAGES * STAGE_BYTES + (2 * BN + 32 * 128 + 8 * 128) * sizeof(float) + (2 * STAGES + 4) * sizeof(uint64_t) + 16 + 1024;
    static bool attr = false;
    if (!attr) {
        HQ_CUDA_OK(cudaFuncSetAttribute((k_rerank_tc<KP, EH, PS>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    int grid = hq_cached_sm_count();
    if (grid > p.num_units) grid = p.num_units;
    k_rerank_tc<KP, EH, PS><<<grid, 64 + 128 * EH, smem, st>>>(mq, mdb, p);
    HQ_LAUNCH_OK("k_rerank_tc");
    const size_t msm = (size_t)p.n_ranges * EH * KP * 8;
    if (msm > 48 * 1024) HQ_CUDA_OK(cudaFuncSetAttribute(k_rerank_tc_merge<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
    k_rerank_tc_merge<KP><<<p.Q, 128, msm, st>>>(p, db_f32, db_stride, q_f32, q_stride, q_norm, k, id_base, ids, scores);
    HQ_LAUNCH_OK("k_rerank_tc_merge");
    return HQ_OK;
}

}  // namespace

extern "C" int hq_to_bf16(const float* src, int64_t N, int64_t D, int64_t src_stride, void* dst, int64_t dst_pitch, void* stream) {
    HQ_REQUIRE(N >= 0 && D > 0 && src_stride >= D && dst_pitch >= D, "bad shape");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(src && dst, "null pointer");
    int64_t blocks = (N * dst_pitch + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_to_bf16<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, N, D, src_stride, (__nv_bfloat16*)dst, dst_pitch);
    HQ_LAUNCH_OK("k_to_bf16");
    return HQ_OK;
}

extern "C" int64_t hq_rerank_bf16_scratch_bytes(int64_t N, int Q, int k) {
    const int kp = pick_kp(k);
    if (kp == 0 || N <= 0 || Q <= 0) return 0;
    TcParams p{};
    plan_units(N, Q, hq_cached_sm_count(), p);
    return (int64_t)p.num_units * 2 * BM * kp * 8;            // sized for two partial lists per unit
}

static int rerank_topk_bf16(bool unit_rows, const int32_t* zero_rows, int n_zero, const void* db_bf16, int64_t db_pitch,
                            const float* db_f32, int64_t db_stride, const float* db_norm,
                            int64_t N, int64_t D, const void* q_bf16, int64_t q_pitch, const float* q_f32, int64_t q_stride,
                            const float* q_norm, int Q, const uint32_t* mask, int64_t mask_stride, int k, int64_t id_base,
                            int64_t* ids, float* scores, void* scratch, int64_t scratch_bytes, void* stream) {
    HQ_REQUIRE(N >= 0 && Q >= 0 && D > 0, "bad shape");
    const int kp = pick_kp(k);
    HQ_REQUIRE(k >= 1 && kp != 0, "k must be in [1, 20] for the tensor-core rerank (got %d)", k);
    if (Q == 0) return HQ_OK;
    HQ_REQUIRE(ids && scores, "null output");
    if (N == 0) return hq_topk_from_scores(nullptr, 0, 0, Q, k, id_base, ids, scores, stream);
    HQ_REQUIRE(db_bf16 && db_f32 && db_norm && q_bf16 && q_f32 && q_norm, "null pointer");
    HQ_REQUIRE(db_pitch % 8 == 0 && q_pitch % 8 == 0 && db_pitch >= D && q_pitch >= D, "bf16 row pitch must be a multiple of 8 and >= D");
    HQ_REQUIRE((reinterpret_cast<uintptr_t>(db_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_bf16) & 15) == 0, "bf16 operands must be 16-byte aligned");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    HQ_REQUIRE(!mask || mask_stride * 32 >= N, "mask stride too small");
    TcParams p{};
    p.N = N; p.Q = Q; p.D = (int)D; p.db_norm = db_norm; p.mask = mask; p.mask_stride = mask_stride;
    p.zero_rows = zero_rows; p.n_zero = n_zero;
    plan_units(N, Q, hq_cached_sm_count(), p);
    // Rows of up to 1024 values: the MMAs of a tile take less time than a four-warp epilogue (768-D ran at 60 % of
    // the tensor peak), so the tile's columns are split over eight epilogue warps.
    // (measured faster at 1536-D too: 2.40 -> 2.25 ms per 1024 x 1 M batch)
    p.eh = 2;
    if (const char* e = getenv("HQ_RERANK_EH")) { if (e[0] == '1' && !unit_rows) p.eh = 1; }
    const int64_t need = (int64_t)p.num_units * p.eh * BM * kp * 8;
    HQ_REQUIRE(scratch && scratch_bytes >= need, "scratch too small: need %lld bytes", (long long)need);
    p.part_val = reinterpret_cast<float*>(scratch);
    p.part_idx = reinterpret_cast<int32_t*>(p.part_val + (int64_t)p.num_units * p.eh * BM * kp);
    CUtensorMap mq, mdb;
    int rc = make_map(&mq, q_bf16, Q, D, q_pitch, BM);
    if (rc != HQ_OK) return rc;
    rc = make_map(&mdb, db_bf16, N, D, db_pitch, BN);
    if (rc != HQ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
#define HQ_TC_LAUNCH(KP_, EH_, PS_) \
    return launch_tc<KP_, EH_, PS_>(mq, mdb, p, db_f32, db_stride, q_f32, q_stride, q_norm, k, id_base, ids, scores, st)
    if (unit_rows) {
        if (kp == 16) HQ_TC_LAUNCH(16, 2, true);
        HQ_TC_LAUNCH(32, 2, true);
    }
    if (kp == 16 && p.eh == 2) HQ_TC_LAUNCH(16, 2, false);
    if (kp == 16) HQ_TC_LAUNCH(16, 1, false);
    if (p.eh == 2) HQ_TC_LAUNCH(32, 2, false);
    HQ_TC_LAUNCH(32, 1, false);
#undef HQ_TC_LAUNCH
}

extern "C" int hq_rerank_topk_bf16(const void* db_bf16, int64_t db_pitch, const float* db_f32, int64_t db_stride, const float* db_norm,
                                   int64_t N, int64_t D, const void* q_bf16, int64_t q_pitch, const float* q_f32, int64_t q_stride,
                                   const float* q_norm, int Q, const uint32_t* mask, int64_t mask_stride, int k, int64_t id_base,
                                   int64_t* ids, float* scores, void* scratch, int64_t scratch_bytes, void* stream) {
    return rerank_topk_bf16(false, nullptr, 0, db_bf16, db_pitch, db_f32, db_stride, db_norm, N, D, q_bf16, q_pitch, q_f32, q_stride, q_norm,
                            Q, mask, mask_stride, k, id_base, ids, scores, scratch, scratch_bytes, stream);
}

extern "C" int hq_rerank_topk_unit_bf16(const void* db_unit_bf16, int64_t db_pitch, const float* db_f32, int64_t db_stride,
                                        const float* db_norm, const int32_t* zero_rows, int n_zero, int64_t N, int64_t D,
                                        const void* q_bf16, int64_t q_pitch, const float* q_f32, int64_t q_stride, const float* q_norm,
                                        int Q, const uint32_t* mask, int64_t mask_stride, int k, int64_t id_base, int64_t* ids,
                                        float* scores, void* scratch, int64_t scratch_bytes, void* stream) {
    HQ_REQUIRE(n_zero >= 0 && (n_zero == 0 || zero_rows), "zero_rows missing");
    return rerank_topk_bf16(true, zero_rows, n_zero, db_unit_bf16, db_pitch, db_f32, db_stride, db_norm, N, D, q_bf16, q_pitch, q_f32,
                            q_stride, q_norm, Q, mask, mask_stride, k, id_base, ids, scores, scratch, scratch_bytes, stream);
}

extern "C" int hq_to_bf16_unit(const float* src, int64_t N, int64_t D, int64_t src_stride, const float* norms, void* dst,
                               int64_t dst_pitch, void* stream) {
    HQ_REQUIRE(N >= 0 && D > 0 && src_stride >= D && dst_pitch >= D, "bad shape");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(src && dst && norms, "null pointer");
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (dst_pitch % 8 == 0 && src_stride % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0) {
        int64_t blocks = (N * (dst_pitch / 8) + 255) / 256;
        if (blocks > cap) blocks = cap;
        k_to_bf16_unit_vec<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, N, D, src_stride, norms, (__nv_bfloat16*)dst, dst_pitch);
        HQ_LAUNCH_OK("k_to_bf16_unit_vec");
        return HQ_OK;
    }
    int64_t blocks = (N * dst_pitch + 255) / 256;
    if (blocks > cap) blocks = cap;
    k_to_bf16_unit<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, N, D, src_stride, norms, (__nv_bfloat16*)dst, dst_pitch);
    HQ_LAUNCH_OK("k_to_bf16_unit");
    return HQ_OK;
}
