// K6: tensor-core rerank.  scores = Q x N x D cosine contraction on tcgen05 (bf16 in, fp32
// accumulate in TMEM) with the survivor mask and a streaming per-query top-k' fused into the
// epilogue, followed by an exact fp32 re-score of a shortlist (k_rerank_tc_merge) whose sufficiency is
// PROVEN per query with a rounding-error bound; queries that cannot be proven are re-scored exactly over
// all their surviving rows (k_guard_rescore / k_guard_merge), so the returned top-k is the exact path's.
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

constexpr int R_MAX = 64;        // rows re-scored exactly per query at most (in chunks of R_CHUNK, until the guard passes)
constexpr int R_CHUNK = 16;
constexpr int FB_SLICES = 128;   // exact fallback: slices of a flagged query's mask row, one CTA each
constexpr int FB_K = 20;         // the tensor-core rerank supports k <= 20

__device__ __forceinline__ uint32_t ord_of(float v) {            // order-preserving float -> uint32 (> 0 for finite values)
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_of_ord(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// dot(q, row) by one warp: lane-strided fmaf chain + xor tree, every lane returns the sum.  fp32 rows: the arithmetic of
// k_rerank_sparse (identical scores on both paths).  bf16 rows (a bf16-only database): the stored values, exactly.
template <bool B16>
__device__ __forceinline__ float warp_row_dot(const float* __restrict__ qv, const void* __restrict__ row, int D, bool vec, int lane) {
    float acc = 0.f;
    if constexpr (!B16) {
        const float* rv = reinterpret_cast<const float*>(row);
        if (vec) {
            for (int i = lane; i < D / 4; i += 32) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(qv) + i);
                const float4 b = __ldg(reinterpret_cast<const float4*>(rv) + i);
                acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
            }
        } else {
            for (int i = lane; i < D; i += 32) acc = fmaf(__ldg(qv + i), __ldg(rv + i), acc);
        }
    } else {
        const __nv_bfloat16* rv = reinterpret_cast<const __nv_bfloat16*>(row);
        if (vec) {
            for (int i = lane; i < D / 8; i += 32) {
                const uint4 w = __ldg(reinterpret_cast<const uint4*>(rv) + i);
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(qv) + 2 * i);
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(qv) + 2 * i + 1);
                acc = fmaf(a0.x, __uint_as_float(w.x << 16), acc); acc = fmaf(a0.y, __uint_as_float(w.x & 0xffff0000u), acc);
                acc = fmaf(a0.z, __uint_as_float(w.y << 16), acc); acc = fmaf(a0.w, __uint_as_float(w.y & 0xffff0000u), acc);
                acc = fmaf(a1.x, __uint_as_float(w.z << 16), acc); acc = fmaf(a1.y, __uint_as_float(w.z & 0xffff0000u), acc);
                acc = fmaf(a1.z, __uint_as_float(w.w << 16), acc); acc = fmaf(a1.w, __uint_as_float(w.w & 0xffff0000u), acc);
            }
        } else {
            for (int i = lane; i < D; i += 32) acc = fmaf(__ldg(qv + i), __bfloat162float(rv[i]), acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// the same dot products for up to four rows at once (loads of the four rows in flight together; every row's own
// accumulation order is that of warp_row_dot, so the values are bit identical)
template <bool B16>
__device__ __forceinline__ void warp_rows_dot4(const float* __restrict__ qv, const void* const (&row)[4], int n_rows, int D, bool vec, int lane,
                                               float (&out)[4]) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (!vec) {
        for (int r = 0; r < n_rows; ++r) out[r] = warp_row_dot<B16>(qv, row[r], D, false, lane);
        return;
    }
    if constexpr (!B16) {
        for (int i = lane; i < D / 4; i += 32) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(qv) + i);
            float4 b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) b[r] = r < n_rows ? __ldg(reinterpret_cast<const float4*>(row[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r] = fmaf(a.x, b[r].x, acc[r]); acc[r] = fmaf(a.y, b[r].y, acc[r]);
                acc[r] = fmaf(a.z, b[r].z, acc[r]); acc[r] = fmaf(a.w, b[r].w, acc[r]);
            }
        }
    } else {
        for (int i = lane; i < D / 8; i += 32) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(qv) + 2 * i);
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(qv) + 2 * i + 1);
            uint4 w[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) w[r] = r < n_rows ? __ldg(reinterpret_cast<const uint4*>(row[r]) + i) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r] = fmaf(a0.x, __uint_as_float(w[r].x << 16), acc[r]); acc[r] = fmaf(a0.y, __uint_as_float(w[r].x & 0xffff0000u), acc[r]);
                acc[r] = fmaf(a0.z, __uint_as_float(w[r].y << 16), acc[r]); acc[r] = fmaf(a0.w, __uint_as_float(w[r].y & 0xffff0000u), acc[r]);
                acc[r] = fmaf(a1.x, __uint_as_float(w[r].z << 16), acc[r]); acc[r] = fmaf(a1.y, __uint_as_float(w[r].z & 0xffff0000u), acc[r]);
                acc[r] = fmaf(a1.z, __uint_as_float(w[r].w << 16), acc[r]); acc[r] = fmaf(a1.w, __uint_as_float(w[r].w & 0xffff0000u), acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        out[r] = acc[r];
    }
}

__device__ __forceinline__ float score_of(float acc, float nq, float nc) {
    return (nq != 0.f && nc != 0.f) ? __fmul_rn(__fadd_rn(__fdiv_rn(acc, __fmul_rn(nq, nc)), 1.0f), 0.5f) : 0.f;
}

// fewer than k results: zero-norm rows (they never reach the shortlists, score exactly 0.0 and follow everybody else), then -1
__device__ void fill_tail(const TcParams& p, int q, int cnt, int k, int64_t id_base, int64_t* ids, float* scores) {
    for (int z = 0; z < p.n_zero && cnt < k; ++z) {
        const int32_t id = p.zero_rows[z];
        if (p.mask && !((p.mask[(int64_t)q * p.mask_stride + (id >> 5)] >> (id & 31)) & 1u)) continue;
        ids[(int64_t)q * k + cnt] = (int64_t)id + id_base;
        scores[(int64_t)q * k + cnt] = 0.f;
        ++cnt;
    }
    for (int j = cnt; j < k; ++j) { ids[(int64_t)q * k + j] = -1; scores[(int64_t)q * k + j] = -1.0f; }
}

// Merge the per-unit shortlists of one query, re-score the best rows exactly, PROVE that no other row can belong to the
// top-k, emit.
//
// Pool: n_sub = n_ranges * eh sublists of KP (value, row) pairs, each the best KP accumulator values of one epilogue thread's
// rows, sorted.  The merge walks the pool in (value desc, row asc) order with a tournament over the sublist heads and
// re-scores the rows it takes in chunks of R_CHUNK (one warp per row, exact).  After every chunk the guard is evaluated:
//   every row NOT re-scored has accumulator value a <= bound = max(last value of every FULL sublist, value of the next head)
//   and true value t = q . c / |c| <= a + E, so if the k-th best exact t of the re-scored rows exceeds bound + E, the exact
//   top-k lies inside the re-scored set.  E bounds |a - t| rigorously:
//     a - t = (q16 - q) . c16  +  q . (c16 - c / |c|)  +  accumulation error of the tensor pipe
//     |a - t| <= |q16 - q| * |c16|  +  |q| * dc_max  +  D * 2^-22 * |q| * |c16|          (Cauchy-Schwarz; |c16| <= 1 + 2^-8)
//   with |q16 - q| computed here from the query, dc_max = max over the rows of |c16 - c / |c|| computed when the database was
//   built (0 for a bf16-only database, whose rows ARE c16), plus 1e-5 |q| for the fp32 rounding of the exact values themselves.
// A query whose guard still fails after R_MAX rows (near-duplicate clusters around the k-th score) is appended to the flag
// list; k_guard_rescore / k_guard_merge then re-score ALL its surviving rows exactly and overwrite its results.
template <int KP, bool B16>
__global__ void __launch_bounds__(128) k_rerank_tc_merge(const TcParams p, const float* __restrict__ db_f32, int64_t db_stride,
                                                         const __nv_bfloat16* __restrict__ db_b16, int64_t db_pitch,
                                                         const float* __restrict__ q_f32, int64_t q_stride,
                                                         const float* __restrict__ q_norm, int k, int64_t id_base, float dc_max,
                                                         int32_t* __restrict__ guard, int64_t* __restrict__ ids,
                                                         float* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int n_sub = p.n_ranges * p.eh, M = n_sub * KP;
    uint64_t* c_key = reinterpret_cast<uint64_t*>(sm);           // [M]  (ord(value) << 32) | ~row, 0 = empty slot
    __shared__ int32_t sel_id[R_MAX];
    __shared__ float ex_s[R_MAX], ex_t[R_MAX];
    __shared__ uint32_t s_B;
    __shared__ float s_red[4];
    __shared__ float s_tk;
    __shared__ int s_state;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m_tile = q / BM, row = q % BM;
    if (tid == 0) { s_B = 0; s_state = 0; }
    __syncthreads();
    for (int e = tid; e < M; e += blockDim.x) {
        const int r = e / KP, j = e - r * KP;                   // r = range * eh + half
        const int64_t u = (int64_t)(r / p.eh) * p.m_tiles + m_tile;
        const int64_t a = (((u * p.eh + r % p.eh) * BM + row) * KP) + j;
        const int32_t id = p.part_idx[a];
        const uint32_t o = ord_of(p.part_val[a]);
        c_key[e] = id >= 0 ? (((uint64_t)o << 32) | (uint32_t)(0xffffffffu - (uint32_t)id)) : 0ull;
        if (j == KP - 1 && id >= 0) atomicMax(&s_B, o);         // a full sublist may have dropped rows up to its last value
    }
    // |bf16(q) - q|
    const float nq = q_norm[q];
    const float* qv = q_f32 + (int64_t)q * q_stride;
    float qe = 0.f;
    for (int i = tid; i < p.D; i += blockDim.x) {
        const float x = __ldg(qv + i);
        const float d = __bfloat162float(__float2bfloat16_rn(x)) - x;
        qe = fmaf(d, d, qe);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qe += __shfl_xor_sync(0xffffffffu, qe, o);
    if (lane == 0) s_red[warp] = qe;
    __syncthreads();
    const float q_err = sqrtf(s_red[0] + s_red[1] + s_red[2] + s_red[3]) * 1.0001f;
    const float E = q_err * 1.00391f + nq * (dc_max + 1.0e-5f + (float)p.D * 2.4e-7f);

    const bool vec = B16 ? ((p.D % 8 == 0) && (db_pitch % 8 == 0) && (q_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(db_b16) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(q_f32) & 15) == 0))
                         : ((p.D % 4 == 0) && (db_stride % 4 == 0) && (q_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(db_f32) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(q_f32) & 15) == 0));
    // ---- merge: ONE warp walks the pool in (value desc, row asc) order, a chunk of R_CHUNK rows at a time, with a
    // tournament over the sublist heads that needs no block-wide barrier (the first version took one element per round
    // of the whole block with two barriers each: ~95 us per query whatever the shard size; ranking every pool entry by
    // binary searches over all sublists was worse: the pool holds up to ~2000 entries in ~120 sublists).
    uint8_t* head = reinterpret_cast<uint8_t*>(c_key + M);       // [n_sub] next untaken entry of every sublist
    __shared__ int s_cnt;
    __shared__ unsigned long long s_next;
    for (int sb = tid; sb < n_sub; sb += blockDim.x) head[sb] = 0;
    __syncthreads();
    int cnt = 0, done = 0;
    while (true) {
        if (warp == 0) {
            int c = cnt;
            // (a first chunk of KP * 3 / 2 rows -- the 17th accumulator is rarely far enough below the 10th exact value, the
            // 25th is -- made the kernel slower, 0.136 -> 0.166 ms: a chunk of 24 is two passes of the four warps, and the kernel
            // is a chain of round trips per query, not a bandwidth problem)
            const int target = min(done + R_CHUNK, R_MAX);
            uint64_t nk;
            while (true) {
                uint64_t bk = 0;
                int bs = -1;
                for (int sb = lane; sb < n_sub; sb += 32) {
                    const int h = head[sb];
                    if (h < KP) {
                        const uint64_t key = c_key[sb * KP + h];
                        if (key > bk) { bk = key; bs = sb; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint64_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
                    const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                    if (ok > bk) { bk = ok; bs = os; }
                }
                nk = bk;
                if (nk == 0ull || c == target) break;
                if (lane == 0) {
                    sel_id[c] = (int32_t)(0xffffffffu - (uint32_t)(nk & 0xffffffffull));
                    head[bs] = (uint8_t)(head[bs] + 1);
                }
                ++c;
                __syncwarp();
            }
            if (lane == 0) { s_cnt = c; s_next = nk; }
        }
        __syncthreads();
        cnt = s_cnt;
        const uint64_t next_key = s_next;
        const bool exhausted = next_key == 0ull;
        {
            // Every 128-byte line of the chunk's rows goes in flight NOW (prefetch.global.L2, fire and forget): the re-score loop
            // below walks a row in D / 128 dependent steps (the loads of a step issue behind the FMAs of the step before), each a
            // DRAM round trip of 1-3 us under the load of a thousand such chains; with the lines already on their way the
            // later steps meet them in L2.
            const int n_new = cnt - done;
            const int64_t row_bytes = (int64_t)p.D * (B16 ? 2 : 4);
            const int lines = (int)((row_bytes + 127) / 128) + 1;        // + the line of the row's last byte (unaligned rows)
            for (int i = tid; i < n_new * lines; i += blockDim.x) {
                const int r = i / lines, l = i - r * lines;
                const int64_t id = sel_id[done + r];
                const char* base = B16 ? reinterpret_cast<const char*>(db_b16 + id * db_pitch) : reinterpret_cast<const char*>(db_f32 + id * db_stride);
                const int64_t off = l + 1 < lines ? (int64_t)l * 128 : row_bytes - 1;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            }
        }
        // ---- exact values of the chunk: a warp takes four rows at once ----
        for (int c0 = done + 4 * warp; c0 < cnt; c0 += 16) {
            const int nr = min(4, cnt - c0);
            const void* rows[4];
            int32_t idr[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                idr[r] = sel_id[c0 + (r < nr ? r : 0)];
                rows[r] = B16 ? reinterpret_cast<const void*>(db_b16 + (int64_t)idr[r] * db_pitch)
                              : reinterpret_cast<const void*>(db_f32 + (int64_t)idr[r] * db_stride);
            }
            float acc[4];
            warp_rows_dot4<B16>(qv, rows, nr, p.D, vec, lane, acc);
            if (lane < nr) {
                float a = acc[0];
                int32_t id = idr[0];
#pragma unroll
                for (int r = 1; r < 4; ++r)
                    if (lane == r) { a = acc[r]; id = idr[r]; }
                const float nc = __ldg(p.db_norm + id);
                ex_s[c0 + lane] = score_of(a, nq, nc);
                ex_t[c0 + lane] = nc != 0.f ? __fdiv_rn(a, nc) : 0.f;
            }
        }
        done = cnt;
        __syncthreads();
        // ---- guard ----
        if (warp == 0) {
            const uint32_t next_o = exhausted ? 0u : (uint32_t)(next_key >> 32);
            const uint32_t bound_o = s_B > next_o ? s_B : next_o;
            int state;
            if (bound_o == 0u) {
                state = 1;                                      // nothing was left out: every surviving row has been re-scored
            } else if (cnt >= k) {
                // k-th largest exact value among the re-scored rows (rank by counting, cnt <= 64)
                for (int i = lane; i < cnt; i += 32) {
                    const float ti = ex_t[i];
                    int rank = 0;
                    for (int j = 0; j < cnt; ++j) rank += (ex_t[j] > ti || (ex_t[j] == ti && j < i)) ? 1 : 0;
                    if (rank == k - 1) s_tk = ti;
                }
                __syncwarp();
                state = (s_tk > float_of_ord(bound_o) + E) ? 1 : ((exhausted || cnt >= R_MAX) ? 2 : 0);
            } else {
                state = (exhausted || cnt >= R_MAX) ? 2 : 0;
            }
            if (lane == 0) s_state = state;
        }
        __syncthreads();
        if (s_state != 0) break;
    }
    // ---- emit: rank the re-scored rows by (exact score desc, row asc) ----
    if (tid < cnt) {
        const int32_t id = sel_id[tid];
        const float v = ex_s[tid];
        int rank = 0;
        for (int j = 0; j < cnt; ++j) rank += (ex_s[j] > v || (ex_s[j] == v && sel_id[j] < id)) ? 1 : 0;
        if (rank < k) { ids[(int64_t)q * k + rank] = (int64_t)id + id_base; scores[(int64_t)q * k + rank] = v; }
    }
    if (tid == 0) {
        fill_tail(p, q, cnt < k ? cnt : k, k, id_base, ids, scores);
        if (guard) {
            atomicAdd(&guard[1], cnt);
            if (s_state == 2) guard[4 + atomicAdd(&guard[0], 1)] = q;
        }
    }
}

// ---- exact fallback of the flagged queries -----------------------------------------------------------------------
// Persistent over (flagged query, mask slice) items read from the device-side flag list; no flagged query = immediate exit.
// One warp per surviving row, exact score; per-warp top-k lists in shared memory, merged per CTA into fb[q][slice][FB_K].
template <bool B16>
__global__ void __launch_bounds__(256) k_guard_rescore(const TcParams p, const float* __restrict__ db_f32, int64_t db_stride,
                                                       const __nv_bfloat16* __restrict__ db_b16, int64_t db_pitch,
                                                       const float* __restrict__ q_f32, int64_t q_stride,
                                                       const float* __restrict__ q_norm, int k, const int32_t* __restrict__ guard,
                                                       float* __restrict__ fb_val, int32_t* __restrict__ fb_idx) {
    __shared__ float l_val[8][FB_K];
    __shared__ int32_t l_idx[8][FB_K];
    const int n_f = guard[0];
    if (n_f <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t words = (p.N + 31) / 32, wps = (words + FB_SLICES - 1) / FB_SLICES;
    const bool vec = B16 ? ((p.D % 8 == 0) && (db_pitch % 8 == 0) && (q_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(db_b16) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(q_f32) & 15) == 0))
                         : ((p.D % 4 == 0) && (db_stride % 4 == 0) && (q_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(db_f32) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(q_f32) & 15) == 0));
    for (int64_t item = blockIdx.x; item < (int64_t)n_f * FB_SLICES; item += gridDim.x) {
        const int f = (int)(item / FB_SLICES), sl = (int)(item - (int64_t)f * FB_SLICES);
        const int q = guard[4 + f];
        const float nq = q_norm[q];
        const float* qv = q_f32 + (int64_t)q * q_stride;
        if (lane < FB_K) { l_val[warp][lane] = -1.f; l_idx[warp][lane] = -1; }
        __syncwarp();
        const int64_t w0 = sl * wps, w1 = min(words, w0 + wps);
        for (int64_t w = w0 + warp; w < w1; w += 8) {
            uint32_t bits = p.mask ? __ldg(p.mask + (int64_t)q * p.mask_stride + w) : 0xffffffffu;
            if (w * 32 + 32 > p.N) bits &= (1u << (uint32_t)(p.N - w * 32)) - 1u;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const int32_t id = (int32_t)(w * 32 + b);
                const float nc = __ldg(p.db_norm + id);
                if (nc == 0.f) continue;                          // zero-norm rows: appended by the merge's tail rule
                const void* rv = B16 ? reinterpret_cast<const void*>(db_b16 + (int64_t)id * db_pitch)
                                     : reinterpret_cast<const void*>(db_f32 + (int64_t)id * db_stride);
                const float s = score_of(warp_row_dot<B16>(qv, rv, p.D, vec, lane), nq, nc);
                const float wv = l_val[warp][k - 1];
                const int32_t wi = l_idx[warp][k - 1];
                if (wi < 0 || s > wv || (s == wv && id < wi)) {
                    if (lane == 0) {
                        int pos = k - 1;
                        while (pos > 0) {
                            const float pv = l_val[warp][pos - 1];
                            const int32_t pi = l_idx[warp][pos - 1];
                            if (pi >= 0 && (pv > s || (pv == s && pi < id))) break;
                            l_val[warp][pos] = pv; l_idx[warp][pos] = pi;
                            --pos;
                        }
                        l_val[warp][pos] = s; l_idx[warp][pos] = id;
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        // CTA top-k of the 8 warp lists: rank by counting (<= 8 * FB_K entries)
        float* ov = fb_val + ((int64_t)q * FB_SLICES + sl) * FB_K;
        int32_t* oi = fb_idx + ((int64_t)q * FB_SLICES + sl) * FB_K;
        if (tid < FB_K) { ov[tid] = -1.f; oi[tid] = -1; }
        __syncthreads();
        if (tid < 8 * k) {
            const int wl = tid / k, j = tid - wl * k;
            const float v = l_val[wl][j];
            const int32_t id = l_idx[wl][j];
            if (id >= 0) {
                int rank = 0;
                for (int a = 0; a < 8; ++a)
                    for (int c = 0; c < k; ++c) {
                        const int32_t ic = l_idx[a][c];
                        if (ic < 0) continue;
                        const float vc = l_val[a][c];
                        rank += (vc > v || (vc == v && ic < id)) ? 1 : 0;
                    }
                if (rank < k) { ov[rank] = v; oi[rank] = id; }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_guard_merge(const TcParams p, int k, int64_t id_base, const int32_t* __restrict__ guard,
                                                     float* __restrict__ fb_val, int32_t* __restrict__ fb_idx,
                                                     int64_t* __restrict__ ids, float* __restrict__ scores) {
    __shared__ float w_val[8];
    __shared__ int32_t w_idx[8];
    __shared__ int w_pos[8];
    const int n_f = guard[0];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int f = blockIdx.x; f < n_f; f += gridDim.x) {
        const int q = guard[4 + f];
        float* cv = fb_val + (int64_t)q * FB_SLICES * FB_K;
        int32_t* ci = fb_idx + (int64_t)q * FB_SLICES * FB_K;
        int cnt = 0;
        for (int round = 0; round < k; ++round) {
            float bv = -2.f;
            int32_t bi = -1;
            int bp = -1;
            for (int e = tid; e < FB_SLICES * FB_K; e += blockDim.x) {
                const int32_t id = ci[e];
                if (id < 0) continue;
                const float v = cv[e];
                if (bp < 0 || v > bv || (v == bv && id < bi)) { bv = v; bi = id; bp = e; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (op >= 0 && (bp < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; bp = op; }
            }
            if (lane == 0) { w_val[warp] = bv; w_idx[warp] = bi; w_pos[warp] = bp; }
            __syncthreads();
            if (tid == 0) {
                int best = -1;
                for (int w = 0; w < 8; ++w) {
                    if (w_pos[w] < 0) continue;
                    if (best < 0 || w_val[w] > w_val[best] || (w_val[w] == w_val[best] && w_idx[w] < w_idx[best])) best = w;
                }
                if (best >= 0) {
                    ids[(int64_t)q * k + round] = (int64_t)w_idx[best] + id_base;
                    scores[(int64_t)q * k + round] = w_val[best];
                    ci[w_pos[best]] = -1;                          // taken
                    w_pos[0] = 1;
                } else {
                    w_pos[0] = -1;
                }
            }
            __syncthreads();
            const bool got = w_pos[0] >= 0;
            __syncthreads();
            if (!got) break;
            ++cnt;
        }
        if (tid == 0) fill_tail(p, q, cnt, k, id_base, ids, scores);
        __syncthreads();
    }
}

// max over the rows of | bf16 unit row - c / |c| |_2  (the database-side term of the guard's error bound)
__global__ void __launch_bounds__(256) k_bf16_unit_error_max(const float* __restrict__ src, int64_t N, int64_t D, int64_t src_stride,
                                                             const float* __restrict__ norms, const __nv_bfloat16* __restrict__ unit,
                                                             int64_t pitch, float* __restrict__ out_max) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float worst = 0.f;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < N; r += warps) {
        const float n = __ldg(norms + r);
        float e = 0.f;
        for (int64_t c = lane; c < D; c += 32) {
            const float x = n > 0.f ? __fdiv_rn(__ldg(src + r * src_stride + c), n) : 0.f;
            const float d = __bfloat162float(unit[r * pitch + c]) - x;
            e = fmaf(d, d, e);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        worst = fmaxf(worst, e);
    }
    if (lane == 0 && worst > 0.f) atomicMax(reinterpret_cast<int*>(out_max), __float_as_int(sqrtf(worst) * 1.0001f + 1e-7f));
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

struct GuardArgs {
    float dc_max;
    int32_t* guard;        // [4 + Q] int32: [0] flagged queries, [1] rows re-scored exactly, [4..] the flagged queries
    float* fb_val;         // [Q][FB_SLICES][FB_K] fallback partial lists
    int32_t* fb_idx;
};

template <int KP, int EH, bool B16>
int launch_tc(const CUtensorMap& mq, const CUtensorMap& mdb, const TcParams& p, const float* db_f32, int64_t db_stride,
              const __nv_bfloat16* db_b16, int64_t db_pitch, const float* q_f32, int64_t q_stride, const float* q_norm, int k,
              int64_t id_base, const GuardArgs& g, int64_t* ids, float* scores, cudaStream_t st) {
    const size_t smem = STAGES * STAGE_BYTES + (2 * BN + 32 * 128 + 8 * 128) * sizeof(float) + (2 * STAGES + 4) * sizeof(uint64_t) + 16 + 1024;
    static bool attr = false;
    if (!attr) {
        HQ_CUDA_OK(cudaFuncSetAttribute((k_rerank_tc<KP, EH, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    int grid = hq_cached_sm_count();
    if (grid > p.num_units) grid = p.num_units;
    HQ_CUDA_OK(cudaMemsetAsync(g.guard, 0, 4 * sizeof(int32_t), st));
    const int tk = hq_time_begin(0, st);
    k_rerank_tc<KP, EH, true><<<grid, 64 + 128 * EH, smem, st>>>(mq, mdb, p);
    hq_time_end(0, tk, st);
    HQ_LAUNCH_OK("k_rerank_tc");
    const size_t msm = (size_t)p.n_ranges * EH * KP * 8 + (size_t)p.n_ranges * EH + 16;
    static size_t msm_set = 48 * 1024;
    if (msm > msm_set) {
        HQ_REQUIRE(msm <= 220 * 1024, "too many partial shortlists per query (%zu bytes of shared memory)", msm);
        HQ_CUDA_OK(cudaFuncSetAttribute((k_rerank_tc_merge<KP, B16>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
        msm_set = msm;
    }
    const int tm = hq_time_begin(3, st);
    k_rerank_tc_merge<KP, B16><<<p.Q, 128, msm, st>>>(p, db_f32, db_stride, db_b16, db_pitch, q_f32, q_stride, q_norm, k, id_base,
                                                       g.dc_max, g.guard, ids, scores);
    hq_time_end(3, tm, st);
    HQ_LAUNCH_OK("k_rerank_tc_merge");
    // exact fallback of the flagged queries (both kernels return at once when the flag list is empty)
    k_guard_rescore<B16><<<hq_cached_sm_count() * 2, 256, 0, st>>>(p, db_f32, db_stride, db_b16, db_pitch, q_f32, q_stride, q_norm, k,
                                                                   g.guard, g.fb_val, g.fb_idx);
    HQ_LAUNCH_OK("k_guard_rescore");
    k_guard_merge<<<64, 256, 0, st>>>(p, k, id_base, g.guard, g.fb_val, g.fb_idx, ids, scores);
    HQ_LAUNCH_OK("k_guard_merge");
    return HQ_OK;
}

int64_t parts_bytes(const TcParams& p, int kp) { return (int64_t)p.num_units * 2 * BM * kp * 8; }

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
    return parts_bytes(p, kp) + (int64_t)Q * FB_SLICES * FB_K * 8;     // partial shortlists + fallback partial lists
}

extern "C" int hq_rerank_topk_unit_bf16(const void* db_unit_bf16, int64_t db_pitch, const float* db_f32, int64_t db_stride,
                                        const float* db_norm, const int32_t* zero_rows, int n_zero, int64_t N, int64_t D,
                                        const void* q_bf16, int64_t q_pitch, const float* q_f32, int64_t q_stride, const float* q_norm,
                                        int Q, const uint32_t* mask, int64_t mask_stride, int k, int64_t id_base, float dc_max,
                                        int64_t* ids, float* scores, int32_t* guard, void* scratch, int64_t scratch_bytes,
                                        void* stream) {
    HQ_REQUIRE(n_zero >= 0 && (n_zero == 0 || zero_rows), "zero_rows missing");
    HQ_REQUIRE(N >= 0 && Q >= 0 && D > 0, "bad shape");
    const int kp = pick_kp(k);
    HQ_REQUIRE(k >= 1 && kp != 0, "k must be in [1, 20] for the tensor-core rerank (got %d)", k);
    if (Q == 0) return HQ_OK;
    HQ_REQUIRE(ids && scores, "null output");
    if (N == 0) return hq_topk_from_scores(nullptr, 0, 0, Q, k, id_base, ids, scores, stream);
    HQ_REQUIRE(db_unit_bf16 && db_norm && q_bf16 && q_f32 && q_norm && guard, "null pointer");
    HQ_REQUIRE(dc_max >= 0.f && dc_max < 1.f, "dc_max out of range");
    HQ_REQUIRE(db_pitch % 8 == 0 && q_pitch % 8 == 0 && db_pitch >= D && q_pitch >= D, "bf16 row pitch must be a multiple of 8 and >= D");
    HQ_REQUIRE((reinterpret_cast<uintptr_t>(db_unit_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_bf16) & 15) == 0, "bf16 operands must be 16-byte aligned");
    HQ_REQUIRE(N < ((int64_t)1 << 31), "N too large for one shard");
    HQ_REQUIRE(!mask || mask_stride * 32 >= N, "mask stride too small");
    TcParams p{};
    p.N = N; p.Q = Q; p.D = (int)D; p.db_norm = db_norm; p.mask = mask; p.mask_stride = mask_stride;
    p.zero_rows = zero_rows; p.n_zero = n_zero;
    plan_units(N, Q, hq_cached_sm_count(), p);
    // Eight epilogue warps (two threads per query row): the MMAs of a 768-D tile take less time than a four-warp epilogue
    // (60 % of the tensor peak), and it measured faster at 1536-D too (2.40 -> 2.25 ms per 1024 x 1 M batch).
    p.eh = 2;
    const int64_t parts = (int64_t)p.num_units * p.eh * BM * kp * 8;
    const int64_t need = parts + (int64_t)Q * FB_SLICES * FB_K * 8;
    HQ_REQUIRE(scratch && scratch_bytes >= need, "scratch too small: need %lld bytes", (long long)need);
    p.part_val = reinterpret_cast<float*>(scratch);
    p.part_idx = reinterpret_cast<int32_t*>(p.part_val + (int64_t)p.num_units * p.eh * BM * kp);
    GuardArgs g{};
    g.dc_max = dc_max;
    g.guard = guard;
    g.fb_val = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + parts);
    g.fb_idx = reinterpret_cast<int32_t*>(g.fb_val + (int64_t)Q * FB_SLICES * FB_K);
    CUtensorMap mq, mdb;
    int rc = make_map(&mq, q_bf16, Q, D, q_pitch, BM);
    if (rc != HQ_OK) return rc;
    rc = make_map(&mdb, db_unit_bf16, N, D, db_pitch, BN);
    if (rc != HQ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* b16 = reinterpret_cast<const __nv_bfloat16*>(db_unit_bf16);
#define HQ_TC_LAUNCH(KP_, B16_) \
    return launch_tc<KP_, 2, B16_>(mq, mdb, p, db_f32, db_stride, b16, db_pitch, q_f32, q_stride, q_norm, k, id_base, g, ids, scores, st)
    if (db_f32) {
        HQ_REQUIRE(db_stride >= D, "fp32 row stride too small");
        if (kp == 16) HQ_TC_LAUNCH(16, false);
        HQ_TC_LAUNCH(32, false);
    }
    if (kp == 16) HQ_TC_LAUNCH(16, true);
    HQ_TC_LAUNCH(32, true);
#undef HQ_TC_LAUNCH
}

extern "C" int hq_bf16_unit_error_max(const float* src, int64_t N, int64_t D, int64_t src_stride, const float* norms,
                                      const void* unit_bf16, int64_t pitch, float* out_max, void* stream) {
    HQ_REQUIRE(N >= 0 && D > 0 && src_stride >= D && pitch >= D, "bad shape");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(src && norms && unit_bf16 && out_max, "null pointer");
    int64_t blocks = (N * 32 + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_bf16_unit_error_max<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, N, D, src_stride, norms,
                                                                             (const __nv_bfloat16*)unit_bf16, pitch, out_max);
    HQ_LAUNCH_OK("k_bf16_unit_error_max");
    return HQ_OK;
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
