// K5 on the tensor cores: the per-level threshold tests of the RAG progressive filter
// (rag/search/engine.py:178-287) for Q queries x N rows as one tcgen05 pass that writes bit
// planes P_l[q][row] = (score_l(q, row) >= thr_l).  It replaces k_filter_bits (CUDA-core FMAs,
// 4.0 ms for 1024 x 1 M x 36 values, issue bound) and feeds the same cascade kernel.
//
// Accuracy.  The test must agree with the fp32 reference except for rows whose score lies
// within ~2e-6 of the threshold, so a plain tf32 contraction (10-bit mantissa) is not enough.
// Every operand value x is split exactly: hi = x with the low 13 mantissa bits cleared (a tf32
// number), lo = (x - hi) truncated the same way; the contraction runs over the three products
// hi*hi + hi*lo + lo*hi.  Both operands hold  [x_hi | x_lo]  per level (2 x kp floats, kp = level width padded to 8); the
// three products are three MMA k-step sequences over (A_hi, B_hi), (A_hi, B_lo), (A_lo, B_hi) -- the descriptors simply
// point at the same hi block twice.  (Round 1 stored A = [q_hi | q_hi | q_lo], B = [c_hi | c_lo | c_hi]: 512 bytes per
// database row instead of 384 for 1536-D, 256 for 768-D.)
// The dropped terms are below 2^-21 relative per product (score error < 3e-7).  Database
// rows are pre-scaled by 1/|c_l|, so the test is  dot >= x*_l * |q_l|  with a per-thread
// constant; zero-norm rows are all-zero operands and are cleared through a per-level validity
// word, zero-norm queries carry a NaN threshold.
//
//   warp 0     : TMA producer (query tile once per unit, 64-row database tiles, 4 stages)
//   warp 1     : TMEM allocator + tcgen05.mma.kind::tf32 issuer: M = 128 queries, N = 64 rows,
//                K = 8 per instruction; one [128 x 64] accumulator per level, two sets
//   warps 2..9 : epilogue, thread = (query, 32-row half of the tile): tcgen05.ld 32 columns per
//                level, FADD + funnel shift per element builds the 32-bit pass word, words are
//                staged in shared memory and flushed as 32-byte runs (8 words = 4 tiles)
#include "hq_tc.cuh"
#include <string.h>

using namespace hq_tc;

namespace {

constexpr int FM = 128;                        // queries per tile (MMA M)
constexpr int FR = 64;                         // database rows per tile (MMA N)
constexpr int KS = 128;                        // most packed floats per row: 4 slabs of 32 (128 B); a layout uses n_slabs of them
constexpr int F_STAGES = 3;
constexpr uint32_t A_SLAB_BYTES = FM * 128;    // 16 KB
constexpr uint32_t B_SLAB_BYTES = FR * 128;    // 8 KB
constexpr uint32_t A_BYTES = 4 * A_SLAB_BYTES;
constexpr uint32_t B_STAGE_BYTES = 4 * B_SLAB_BYTES;
constexpr int F_THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr uint32_t ACC_COLS = 3 * FR;          // one accumulator set: 3 levels x 64 columns
constexpr uint32_t F_TMEM_COLS = 512;
constexpr uint32_t kIdescTf32 = make_idesc(2 /*tf32*/, FM, FR);
constexpr int TILE_GROUP = 4;                  // tiles per flush: 8 words = one 32-byte sector per query and level
constexpr int STAGE_WORDS = 3 * 2 * FM * 4;      // pass words of one flush group: [level][chunk of 4 words][query][4]
constexpr int LBUF = 16;                       // per-thread staging slots of the candidate lists (flushed 8 at a time)

struct Segs {
    int L;
    int off[3];        // first packed column of the level
    int kp[3];         // level width padded to 8 (one tf32 k-step)
    int n_slabs;       // 128-byte slabs of a DATABASE row: the level blocks
    int n_slabs_q;     // slabs of a QUERY row: the level blocks + three threshold blocks (window mode's folded thresholds)
    int thr_off;       // first packed column after the level blocks: A holds its threshold blocks there
    int ones_in_row;   // the B side of the folded thresholds, a block (1, 1, 0 ...): 1 = in the padding of the row's fourth slab
                       // (columns thr_off ..), 0 = a constant slab in shared memory (the stage's unused fourth slab)
};

bool make_segs(const hq_index_layout* lay, Segs& s) {
    if (!lay || lay->L < 1 || lay->L > 3) return false;
    int off = 0;
    s.L = lay->L;
    for (int l = 0; l < 3; ++l) { s.off[l] = 0; s.kp[l] = 0; }
    for (int l = 0; l < lay->L; ++l) {
        if (lay->lvl_keff[l] < 1) return false;
        s.kp[l] = (lay->lvl_keff[l] + 7) & ~7;
        s.off[l] = off;
        off += 2 * s.kp[l];                     // [hi | lo]
    }
    s.thr_off = off;
    if (off + 8 * 3 > KS) return false;
    s.n_slabs = (off + 31) / 32;
    s.n_slabs_q = (off + 8 * 3 + 31) / 32;
    s.ones_in_row = s.n_slabs == 4 ? 1 : 0;     // off + 8 <= KS holds (checked above): the block fits the padding
    return true;
}

struct FtcParams {
    int64_t N;
    int Q, L;
    int seg_off[3], ksteps[3];  // first packed column of the level's hi block; k-steps of 8 floats per block (kp / 8)
    int n_slabs, n_slabs_q, thr_off, ones_in_row;
    int m_tiles, n_tiles, n_ranges, tiles_per_range, num_units;
    const float* tq;            // [3][Q]  x*_l * |q_l| (NaN when |q_l| == 0)
    const uint32_t* valid;      // [L][valid_pitch] bit r: row r has a non-zero level norm
    int64_t valid_pitch;
    uint32_t* bits;             // [L][Q][bits_pitch]
    int64_t words, bits_pitch;
    int bits_vec;               // planes are 32-byte aligned with a pitch of whole sectors: 256-bit stores
    // optional candidate lists (L >= 2): rows passing levels 0 and 1, one private segment per
    // (query, row range, tile half) so that no atomics are needed and rows stay ascending
    uint32_t* l_rows;           // [Q][n_segs][seg_cap]  row (the level-2 test is k2 >= tq[2], repeated by the cascade)
    float* l_k1;                // [Q][n_segs][seg_cap]  level-1 dot product (order == score order within a query)
    float* l_k2;                // [Q][n_segs][seg_cap]  level-2 dot product (L == 3)
    int32_t* seg_n;             // [Q][n_segs] entries produced (may exceed seg_cap: overflow, list invalid)
    int64_t seg_cap;
    int n_segs;
    int tile_stride;            // sample pass: only every tile_stride-th 64-row tile is visited (no planes are written)
    const int32_t* unit_only;   // optional [m_tiles]: query tiles with a zero flag are skipped (fallback pass)
    int32_t* c0_cnt;            // optional [Q]: += rows passing the level-0 threshold (sample pass)
    // window mode (k_filter_bits_tc<true>): per-query key windows around the predicted ratio cuts
    const float* win;           // [4][Q] lo1, hi1, lo2, hi2 (dot-product units; window = lo <= k < hi)
    int32_t* wcnt;              // [4][Q] += c0 | rows above window 1 | of those, k2 >= lo2 | of those, k2 >= hi2 (= alive plane)
};

constexpr int LBUF_W = 8;                       // window mode: ring of eight entries per array, flushed when full
constexpr int KSTAGE_BYTES = 8 * EPI_THREADS * 16;   // window mode: k1 / k2 of 16 rows per epilogue thread, 16-byte chunks

__device__ __forceinline__ uint32_t pass_word(const uint32_t (&r)[32], float tq) {
    // bit j = (r[j] >= tq): the sign of (r[j] - tq) is shifted in element by element; four independent
    // chains of eight so that the funnel shifts do not serialise on one register
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float d = __fadd_rn(__uint_as_float(r[c * 8 + j]), -tq);
            w[c] = __funnelshift_l(__float_as_uint(d), w[c], 1);
        }
    }
    const uint32_t all = (w[0] << 24) | (w[1] << 16) | (w[2] << 8) | w[3];      // bit 31 - j = sign of element j
    return ~__brev(all);
}

__device__ __forceinline__ uint32_t sign_word(const uint32_t (&r)[32]) {
    // bit j = (r[j] >= 0) as a float, i.e. its sign bit is clear: the funnel shifts of pass_word without the subtraction
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = __funnelshift_l(r[c * 8 + j], w[c], 1);
    }
    const uint32_t all = (w[0] << 24) | (w[1] << 16) | (w[2] << 8) | w[3];
    return ~__brev(all);
}

__device__ __forceinline__ void st_global_256(uint32_t* dst, const uint32_t (&v)[8]) {       // one full 32-byte sector
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}


// One unit (128-query tile x row range) of the epilogue in WINDOW mode.  The ratio cuts of levels 1 and 2 keep a fixed
// fraction of the rows that pass level 0, so on most data they bind far above the thresholds and ~13 % of all (query, row)
// pairs used to be appended to the candidate lists only for the cascade to find two cut keys among them.  Here every query
// carries two key windows [lo1, hi1) and [lo2, hi2) that a sample pass predicted around its cuts (k_filter_predict):
//   k1 >= hi1 : survives cut 1 for sure           k2 >= hi2 : survives cut 2 for sure
//   k1 <  lo1 : dropped by cut 1 for sure         k2 <  lo2 : dropped by cut 2 (or fails the threshold) for sure
// so the pass writes ONE plane (rows alive for sure), four counters per query and appends only the rows inside a window
// (~1.5 % of the pairs instead of 13 %); the cascade ranks those exactly and checks the prediction against the counters
// (a query whose cut falls outside its window is redone by the fallback pass).  The five sign tests are bit-mask
// arithmetic; the rare appends walk the set bits and fetch k1 / k2 from a per-thread staging area in shared memory
// (register arrays cannot be indexed by a run-time bit position).
__device__ __forceinline__ void win_unit(const FtcParams& p, int m_tile, int range, int t0, int t1, int q, bool q_ok, bool warp_has_q,
                                         float tq0, int ew, int half, int q_in, int et, int lane, uint32_t& it, uint32_t tmem_base,
                                         uint32_t (&vw_next)[3], uint32_t* s_stage, uint32_t* s_lists, uint8_t* s_kstage,
                                         uint64_t* tfull_bar, uint64_t* tempty_bar) {
    const float qnan = __int_as_float(0x7fc00000);
    const float lo1 = q_ok ? __ldg(p.win + (int64_t)0 * p.Q + q) : qnan, hi1 = q_ok ? __ldg(p.win + (int64_t)1 * p.Q + q) : qnan;
    const float lo2 = q_ok ? __ldg(p.win + (int64_t)2 * p.Q + q) : qnan, hi2 = q_ok ? __ldg(p.win + (int64_t)3 * p.Q + q) : qnan;
    const bool three = p.L > 2;                                            // two-level layouts (32 x 32 grids) have no cut 2
    const bool okw = (tq0 == tq0) && (lo1 == lo1) && (hi1 == hi1) && (!three || ((lo2 == lo2) && (hi2 == hi2)));
    const float w1 = hi1 - lo1, w2 = three ? hi2 - lo2 : 0.f;            // upper window edges relative to the folded lower ones
    const bool lists = p.l_rows != nullptr && q_ok;
    const int64_t seg_base = ((int64_t)q * p.n_segs + (range * 2 + half)) * p.seg_cap;
    const int seg_cap32 = (int)min(p.seg_cap, (int64_t)0x7ffffff0);
    uint32_t run = 0, flushed = 0;
    uint32_t c_c0 = 0, c_a1 = 0, c_a1a2 = 0, c_al = 0;
    uint32_t* const ring = s_lists + et * LBUF_W;                          // [3][EPI_THREADS][LBUF_W]
    constexpr int RING_W = EPI_THREADS * LBUF_W;
    float* const kst = reinterpret_cast<float*>(s_kstage + et * 16);       // chunk c at kst + c * EPI_THREADS * 4 floats
    constexpr int KCH = EPI_THREADS * 4;
    for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const int tl = (t - t0) & (TILE_GROUP - 1);
        uint32_t vw[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) vw[l] = vw_next[l];
        if (t + 1 < t1) {
#pragma unroll
            for (int l = 0; l < 3; ++l) vw_next[l] = __ldg(p.valid + (int64_t)l * p.valid_pitch + 2 * (int64_t)(t + 1) * p.tile_stride + half);
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        if (!warp_has_q) {
            mbar_arrive(&tempty_bar[acc]);
            continue;
        }
        tc_fence_after();
        uint32_t r0[32], r1[32], r2[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * ACC_COLS + half * 32;
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + FR, r1);
        if (three) tmem_ld32(taddr + 2 * FR, r2);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        // The accumulators hold dot0 - t0, dot1 - lo1, dot2 - lo2 (thresholds folded into the contraction): the level-0 test and
        // the lower window edges are SIGN tests (one funnel shift per element), the upper edges one subtraction more
        // (k - (hi - lo) >= 0).  Classes, counters, the alive word and the word E of the window rows are bit-mask arithmetic.
        uint32_t W0 = 0, A1 = 0, B1 = 0, A2 = 0xffffffffu, B2 = 0xffffffffu;
        if (okw) {
            W0 = sign_word(r0) & vw[0];
            A1 = sign_word(r1) & vw[1];
            B1 = pass_word(r1, w1) & A1;
            if (three) { A2 = sign_word(r2) & vw[2]; B2 = pass_word(r2, w2) & A2; }
        }
        uint32_t alive = 0, E = 0;
        if (okw) {
            const uint32_t S1 = W0 & B1;
            alive = S1 & B2;
            E = W0 & A1 & (~B1 | (A2 & ~B2));
            c_c0 += (uint32_t)__popc(W0);
            c_a1 += (uint32_t)__popc(S1);
            c_a1a2 += (uint32_t)__popc(S1 & A2);
            c_al += (uint32_t)__popc(alive);
        }
        if (lists && E) {
            const uint32_t row0 = (uint32_t)t * (uint32_t)p.tile_stride * FR + (uint32_t)half * 32u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t m = (E >> (16 * h)) & 0xffffu;
                if (m) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        *reinterpret_cast<uint4*>(kst + c * KCH) = make_uint4(r1[16 * h + 4 * c], r1[16 * h + 4 * c + 1], r1[16 * h + 4 * c + 2], r1[16 * h + 4 * c + 3]);
                        if (three) *reinterpret_cast<uint4*>(kst + (4 + c) * KCH) = make_uint4(r2[16 * h + 4 * c], r2[16 * h + 4 * c + 1], r2[16 * h + 4 * c + 2], r2[16 * h + 4 * c + 3]);
                    }
                    // (two bits per turn with both rows' keys fetched before the first append: 1.46 ms instead of 1.39 ms)
                    while (m) {
                        const int j = __ffs((int)m) - 1;
                        m &= m - 1;
                        const float* src = kst + (j >> 2) * KCH + (j & 3);
                        // the keys are RELATIVE to the lower window edges (k1 = dot1 - lo1, k2 = dot2 - lo2); the cascade repeats
                        // the class tests on them (k1 < hi1 - lo1: inside window 1; k2 >= hi2 - lo2: above window 2)
                        const uint32_t k1b = __float_as_uint(src[0]), k2b = __float_as_uint(src[4 * KCH]);
                        const uint32_t roww = row0 + 16u * h + (uint32_t)j;
                        const uint32_t slot = run & (LBUF_W - 1);
                        ring[slot] = roww;
                        ring[RING_W + slot] = k1b;
                        ring[2 * RING_W + slot] = k2b;
                        ++run;
                        if ((run & (LBUF_W - 1)) == 0) {                    // ring full: one 32-byte sector per array
                            if ((int)flushed + 8 <= seg_cap32) {
                                uint32_t v[3][8];
#pragma unroll
                                for (int a = 0; a < 3; ++a) {
                                    const uint4 x = *reinterpret_cast<const uint4*>(ring + a * RING_W);
                                    const uint4 y = *reinterpret_cast<const uint4*>(ring + a * RING_W + 4);
                                    v[a][0] = x.x; v[a][1] = x.y; v[a][2] = x.z; v[a][3] = x.w;
                                    v[a][4] = y.x; v[a][5] = y.y; v[a][6] = y.z; v[a][7] = y.w;
                                }
                                st_global_256(p.l_rows + seg_base + flushed, v[0]);
                                st_global_256(reinterpret_cast<uint32_t*>(p.l_k1 + seg_base + flushed), v[1]);
                                if (three) st_global_256(reinterpret_cast<uint32_t*>(p.l_k2 + seg_base + flushed), v[2]);
                            }
                            flushed += 8;
                        }
                    }
                }
            }
        }
        s_stage[((tl >> 1) * FM + q_in) * 4 + (tl & 1) * 2 + half] = alive;
        if (tl == TILE_GROUP - 1 || t == t1 - 1) {
            asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");
            const int64_t wb = 2 * ((int64_t)(t - tl) * p.tile_stride);
            const int n_w = 2 * (tl + 1);
            if (half == 0) {
                const int qq = ew * 32 + lane;
                const int gq = m_tile * FM + qq;
                if (gq < p.Q) {
                    const uint4 a = *reinterpret_cast<const uint4*>(s_stage + (0 * FM + qq) * 4);
                    const uint4 b = *reinterpret_cast<const uint4*>(s_stage + (1 * FM + qq) * 4);
                    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    uint32_t* dst = p.bits + (int64_t)gq * p.bits_pitch + wb;
                    if (p.bits_vec && n_w == 8 && wb + 8 <= p.words) {
                        st_global_256(dst, v);
                    } else {
#pragma unroll
                        for (int wi = 0; wi < 8; ++wi)
                            if (wi < n_w && wb + wi < p.words) dst[wi] = v[wi];
                    }
                }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");
        }
    }
    if (lists) {
        for (uint32_t e = flushed; e < run && (int64_t)e < p.seg_cap; ++e) {
            const uint32_t slot = e & (LBUF_W - 1);
            p.l_rows[seg_base + e] = ring[slot];
            p.l_k1[seg_base + e] = __uint_as_float(ring[RING_W + slot]);
            if (three) p.l_k2[seg_base + e] = __uint_as_float(ring[2 * RING_W + slot]);
        }
        p.seg_n[(int64_t)q * p.n_segs + range * 2 + half] = (int)run;
    }
    if (q_ok && p.wcnt) {
        if (c_c0) atomicAdd(p.wcnt + (int64_t)0 * p.Q + q, (int)c_c0);
        if (c_a1) atomicAdd(p.wcnt + (int64_t)1 * p.Q + q, (int)c_a1);
        if (c_a1a2) atomicAdd(p.wcnt + (int64_t)2 * p.Q + q, (int)c_a1a2);
        if (c_al) atomicAdd(p.wcnt + (int64_t)3 * p.Q + q, (int)c_al);
    }
}

template <bool WIN>
__global__ void __launch_bounds__(F_THREADS, 1) k_filter_bits_tc(const __grid_constant__ CUtensorMap map_q,
                                                                 const __grid_constant__ CUtensorMap map_db, const FtcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_a = smem;                                        // query tile, resident for a unit
    uint8_t* smem_b = smem + A_BYTES;                              // F_STAGES database tiles
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(smem_b + F_STAGES * B_STAGE_BYTES);   // [3][2][128][4] ([1][2][128][4] in window mode)
    uint32_t* s_lists = s_stage + (WIN ? STAGE_WORDS / 3 : STAGE_WORDS);                  // [3][EPI_THREADS][LBUF], 16-byte chunks swizzled
    // window mode: [3][EPI_THREADS][LBUF_W] rings, then the k1 / k2 staging area [8][EPI_THREADS][4]
    uint8_t* s_kstage = reinterpret_cast<uint8_t*>(s_lists + 3 * LBUF_W * EPI_THREADS);
    uint64_t* bars = WIN ? reinterpret_cast<uint64_t*>(s_kstage + KSTAGE_BYTES) : reinterpret_cast<uint64_t*>(s_lists + 3 * LBUF * EPI_THREADS);
    uint64_t* full_bar = bars;                      // [F_STAGES]
    uint64_t* empty_bar = bars + F_STAGES;          // [F_STAGES]
    uint64_t* tfull_bar = bars + 2 * F_STAGES;      // [2]
    uint64_t* tempty_bar = bars + 2 * F_STAGES + 2; // [2]
    uint64_t* qfull_bar = bars + 2 * F_STAGES + 4;
    uint64_t* qempty_bar = bars + 2 * F_STAGES + 5;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * F_STAGES + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_db)) : "memory");
        for (int i = 0; i < F_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], EPI_THREADS); }
        mbar_init(qfull_bar, 1);
        mbar_init(qempty_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(F_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if constexpr (WIN) {
        if (!p.ones_in_row) {
            // constant B block of the folded thresholds: the fourth slab of every stage (never written by the loads, n_slabs <= 3)
            // holds (1, 1, 0 ...) in the first 16-byte chunk of each row, at its 128-byte-swizzled place (chunk ^ (row & 7))
            for (int i = threadIdx.x; i < F_STAGES * FR * 8; i += F_THREADS) {
                const int st = i / (FR * 8), r = (i >> 3) % FR, c = i & 7;
                const float one = c == (r & 7) ? 1.0f : 0.f;
                *reinterpret_cast<float4*>(smem_b + st * B_STAGE_BYTES + 3 * B_SLAB_BYTES + r * 128 + c * 16) = make_float4(one, one, 0.f, 0.f);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, uq = 0;
            for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
                const int m_tile = u % p.m_tiles, range = u / p.m_tiles;
                if (p.unit_only && !__ldg(p.unit_only + m_tile)) continue;
                const int t0 = range * p.tiles_per_range;
                const int t1 = min(p.n_tiles, t0 + p.tiles_per_range);
                mbar_wait_relaxed(qempty_bar, (uq & 1u) ^ 1u);                  // MMAs of the previous unit have retired
                ++uq;
                mbar_expect_tx(qfull_bar, (uint32_t)p.n_slabs_q * A_SLAB_BYTES);
                for (int s = 0; s < p.n_slabs_q; ++s) tma_load_2d(&map_q, qfull_bar, smem_a + s * A_SLAB_BYTES, s * 32, m_tile * FM);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], (uint32_t)p.n_slabs * B_SLAB_BYTES);
                    for (int s = 0; s < p.n_slabs; ++s)
                        tma_load_2d(&map_db, &full_bar[stage], smem_b + stage * B_STAGE_BYTES + s * B_SLAB_BYTES, s * 32, t * p.tile_stride * FR);
                    if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        uint32_t stage = 0, phase = 0, it = 0, uq = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const int range = u / p.m_tiles;
            if (p.unit_only && !__ldg(p.unit_only + u % p.m_tiles)) continue;
            const int t0 = range * p.tiles_per_range;
            const int t1 = min(p.n_tiles, t0 + p.tiles_per_range);
            mbar_wait_relaxed(qfull_bar, uq & 1u);
            ++uq;
            tc_fence_after();
            for (int t = t0; t < t1; ++t, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait_relaxed(&tempty_bar[acc], acc_phase ^ 1);
                mbar_wait_relaxed(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(smem_a), b0 = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        if (l < p.L) {
                            // hi*hi, hi*lo, lo*hi: (A block, B block) = (0, 0), (0, 1), (1, 0)
#pragma unroll
                            for (int pr = 0; pr < 3; ++pr) {
                                const uint32_t a_blk = pr == 2 ? 1u : 0u, b_blk = pr == 1 ? 1u : 0u;
                                for (int s = 0; s < p.ksteps[l]; ++s) {
                                    const uint32_t offa = (uint32_t)p.seg_off[l] + a_blk * 8u * p.ksteps[l] + 8u * s;
                                    const uint32_t offb = (uint32_t)p.seg_off[l] + b_blk * 8u * p.ksteps[l] + 8u * s;
                                    const uint64_t da = make_smem_desc(a0 + (offa >> 5) * A_SLAB_BYTES + (offa & 31u) * 4u);
                                    const uint64_t db = make_smem_desc(b0 + (offb >> 5) * B_SLAB_BYTES + (offb & 31u) * 4u);
                                    umma_tf32(tmem_base + acc * ACC_COLS + l * FR, da, db, kIdescTf32, (pr > 0 || s > 0) ? 1u : 0u);
                                }
                            }
                            if constexpr (WIN) {
                                // window mode: one more k-step makes the accumulator  dot_l - threshold_l  (-t0, -lo1, -lo2 in
                                // the query operand's threshold blocks against a (1, 1, 0 ...) block: the padding of the
                                // rows' fourth slab, or the constant slab behind a stage's three data slabs)
                                const uint32_t offt = (uint32_t)p.thr_off + 8u * l, offo = (uint32_t)p.thr_off;
                                const uint64_t da = make_smem_desc(a0 + (offt >> 5) * A_SLAB_BYTES + (offt & 31u) * 4u);
                                const uint64_t db = make_smem_desc(p.ones_in_row ? b0 + (offo >> 5) * B_SLAB_BYTES + (offo & 31u) * 4u
                                                                                 : b0 + 3u * B_SLAB_BYTES);
                                umma_tf32(tmem_base + acc * ACC_COLS + l * FR, da, db, kIdescTf32, 1u);
                            }
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    umma_commit(&tfull_bar[acc]);
                    if (t == t1 - 1) umma_commit(qempty_bar);
                }
                __syncwarp();
                if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue =================
        const int ew = warp & 3;                         // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                // which 32 rows of the 64-row tile
        const int q_in = ew * 32 + lane;
        const int et = (warp - 2) * 32 + lane;           // 0..255
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const int m_tile = u % p.m_tiles, range = u / p.m_tiles;
            if (p.unit_only && !__ldg(p.unit_only + m_tile)) continue;
            const int t0 = range * p.tiles_per_range;
            const int t1 = min(p.n_tiles, t0 + p.tiles_per_range);
            const int q = m_tile * FM + q_in;
            const bool q_ok = q < p.Q;
            // A batch of a few queries fills one lane quarter of the last (or only) query tile: the warps of the other
            // quarters skip the epilogue (single query against 1 M rows: 141 -> ~90 us, the pass becomes HBM bound).
            const bool warp_has_q = m_tile * FM + ew * 32 < p.Q;
            float tq[3];
#pragma unroll
            for (int l = 0; l < 3; ++l) tq[l] = (q_ok && l < p.L) ? __ldg(p.tq + (int64_t)l * p.Q + q) : __int_as_float(0x7fc00000);
            const bool lists = p.l_rows != nullptr && p.L >= 2 && q_ok;
            const int64_t seg_base = ((int64_t)q * p.n_segs + (range * 2 + half)) * p.seg_cap;
            int seg_pos = 0, seg_flushed = 0;
            // Byte offset of the next sector of this thread's segment, kept in a register pair and advanced by 32 per flush
            // (opaque to the compiler: it used to rebuild (q * n_segs + seg) * seg_cap + flushed from the constant bank with
            // 64-bit multiplies in every flush), and the capacity as a 32-bit value (one compare instead of a 64-bit pair).
            uint64_t goff = (uint64_t)seg_base * 4u;
            asm volatile("" : "+l"(goff));
            const int seg_cap32 = (int)min(p.seg_cap, (int64_t)0x7ffffff0);
            // Per-thread ring of LBUF entries per array, 64 bytes per thread, so that a flush reads its eight
            // entries with two 128-bit loads per array (8 scalar loads per array made the flush 22 % of the
            // kernel's instructions).  The 16-byte chunk index is XORed with (lane >> 1) & 3: the eight lanes of
            // a quarter warp then hit eight different bank groups.
            constexpr uint32_t RING_A_BYTES = EPI_THREADS * LBUF * 4;
            const uint32_t swz16 = (uint32_t)((lane >> 1) & 3) << 4;
            const uint32_t sl_addr = smem_u32(s_lists + et * LBUF);             // 64-byte aligned
            const uint32_t sl_addr_x = sl_addr ^ swz16;
            uint32_t vw_next[3];
#pragma unroll
            for (int l = 0; l < 3; ++l) vw_next[l] = (l < p.L && t0 < t1) ? __ldg(p.valid + (int64_t)l * p.valid_pitch + 2 * (int64_t)t0 * p.tile_stride + half) : 0u;
            if constexpr (WIN) {
                win_unit(p, m_tile, range, t0, t1, q, q_ok, warp_has_q, tq[0], ew, half, q_in, et, lane, it, tmem_base, vw_next, s_stage, s_lists, s_kstage,
                         tfull_bar, tempty_bar);
                continue;
            }
            uint32_t c0_acc = 0;
            for (int t = t0; t < t1; ++t, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                const int tl = (t - t0) & (TILE_GROUP - 1);
                uint32_t vw[3];
#pragma unroll
                for (int l = 0; l < 3; ++l) vw[l] = vw_next[l];
                if (t + 1 < t1) {                        // the next tile's validity words travel during this tile
#pragma unroll
                    for (int l = 0; l < 3; ++l) vw_next[l] = l < p.L ? __ldg(p.valid + (int64_t)l * p.valid_pitch + 2 * (int64_t)(t + 1) * p.tile_stride + half) : 0u;
                }
                mbar_wait(&tfull_bar[acc], acc_phase);
                if (!warp_has_q) {                       // no query in this warp's 32 rows of the tile: release the accumulator, nothing to test
                    mbar_arrive(&tempty_bar[acc]);
                    continue;
                }
                tc_fence_after();
                uint32_t r0[32], r1[32], r2[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * ACC_COLS + half * 32;
                tmem_ld32(taddr, r0);
                if (p.L > 1) tmem_ld32(taddr + FR, r1);
                if (p.L > 2) tmem_ld32(taddr + 2 * FR, r2);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);           // the values are in registers: the accumulators are free
                uint32_t w0 = pass_word(r0, tq[0]) & vw[0], w1 = 0, w2 = 0;
                if (!(tq[0] == tq[0])) w0 = 0;
                c0_acc += (uint32_t)__popc(w0);
                s_stage[((0 * 2 + (tl >> 1)) * FM + q_in) * 4 + (tl & 1) * 2 + half] = w0;
                if (p.L > 1) {
                    w1 = pass_word(r1, tq[1]) & vw[1];
                    if (!(tq[1] == tq[1])) w1 = 0;
                    s_stage[((1 * 2 + (tl >> 1)) * FM + q_in) * 4 + (tl & 1) * 2 + half] = w1;
                }
                if (p.L > 2) {
                    w2 = pass_word(r2, tq[2]) & vw[2];
                    if (!(tq[2] == tq[2])) w2 = 0;
                    s_stage[((2 * 2 + (tl >> 1)) * FM + q_in) * 4 + (tl & 1) * 2 + half] = w2;
                }
                if (lists) {
                    // Rows that pass the level-0 and level-1 thresholds.  Entries are staged in per-thread
                    // shared-memory slots and leave as whole 32-byte sectors (8 entries per array): storing
                    // them one 4-byte value at a time took this kernel from 0.8 to 11.4 ms, 16-byte halves 2.7 ms
                    // (partial-sector writes are read-modify-write in the ECC-protected L2).
                    const uint32_t t1w = w0 & w1;
                    if (t1w) {
                        const uint32_t row0 = (uint32_t)t * (uint32_t)p.tile_stride * FR + (uint32_t)half * 32u;
                        uint32_t run = (uint32_t)seg_pos << 2;         // entry counter in bytes of one ring array
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            // Branch-free appends: every bit position is visited with predicated stores; the ring slot
                            // comes from a running predicated counter (a prefix popcount per bit cost 4 more
                            // instructions; a branch per bit ran at 1.7 IPC).
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                const int j = g * 8 + jj;
                                const uint32_t bit = t1w & (1u << j);
                                const uint32_t roww = row0 + j;
                                // ring slot = base ^ (run & mask) as ONE lop3 (the compiler split it into an AND and an XOR)
                                asm volatile(
                                    "{\n\t.reg .pred p;\n\t.reg .b32 a;\n\tsetp.ne.u32 p, %1, 0;\n\t"
                                    "lop3.b32 a, %0, %8, %2, 0x6a;\n\t"
                                    "@p st.shared.u32 [a], %3;\n\t"
                                    "@p st.shared.u32 [a+%6], %4;\n\t"
                                    "@p st.shared.u32 [a+%7], %5;\n\t"
                                    "@p add.u32 %0, %0, 4;\n\t}"
                                    : "+r"(run)
                                    : "r"(bit), "r"(sl_addr_x), "r"(roww), "r"(r1[j]), "r"(r2[j]), "n"(RING_A_BYTES), "n"(2 * RING_A_BYTES),
                                      "n"(LBUF * 4 - 1)
                                    : "memory");
                            }
                            seg_pos = (int)(run >> 2);
                            if (seg_pos - seg_flushed >= 8) {
                                if (seg_flushed + 8 <= seg_cap32) {
                                    const uint32_t f0 = sl_addr_x ^ (((uint32_t)seg_flushed * 4u) & (LBUF * 4 - 1));   // chunk of entries f..f+3
                                    const uint32_t f1 = f0 ^ 16u;                                                      // f+4..f+7 (f is a multiple of 8)
                                    // all six loads first, into separate registers: re-using the eight registers of one
                                    // array for the next made every load wait for the previous 256-bit store to read them
                                    uint32_t v[3][8];
#pragma unroll
                                    for (int a = 0; a < 3; ++a) {
                                        if (a == 2 && p.L < 3) break;
                                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[a][0]), "=r"(v[a][1]), "=r"(v[a][2]), "=r"(v[a][3]) : "r"(f0 + a * RING_A_BYTES));
                                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[a][4]), "=r"(v[a][5]), "=r"(v[a][6]), "=r"(v[a][7]) : "r"(f1 + a * RING_A_BYTES));
                                    }
                                    st_global_256(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(p.l_rows) + goff), v[0]);
                                    st_global_256(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(p.l_k1) + goff), v[1]);
                                    if (p.L > 2) st_global_256(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(p.l_k2) + goff), v[2]);
                                }
                                seg_flushed += 8;
                                goff += 32u;
                            }
                        }
                    }
                }
                if (p.bits && (tl == TILE_GROUP - 1 || t == t1 - 1)) {
                    // the two warps that share a TMEM lane quarter (row halves 0 and 1 of the same 32 queries) write
                    // those queries' words out between themselves: a 64-thread named barrier instead of all 256
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");
                    const int64_t wb = 2 * (int64_t)(t - tl);
                    const int n_w = 2 * (tl + 1);
                    // one (level, query) row of the group per thread: two 128-bit reads, one 32-byte sector out
                    // (word-by-word copies were 10 % of the kernel's instructions)
                    for (int e = half * 32 + lane; e < 3 * 32; e += 64) {
                        const int l = e >> 5, qq = ew * 32 + (e & 31);
                        const int gq = m_tile * FM + qq;
                        if (l < p.L && gq < p.Q) {
                            const uint4 a = *reinterpret_cast<const uint4*>(s_stage + ((l * 2 + 0) * FM + qq) * 4);
                            const uint4 b = *reinterpret_cast<const uint4*>(s_stage + ((l * 2 + 1) * FM + qq) * 4);
                            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                            uint32_t* dst = p.bits + ((int64_t)l * p.Q + gq) * p.bits_pitch + wb;
                            if (p.bits_vec && n_w == 8 && wb + 8 <= p.words) {
                                st_global_256(dst, v);
                            } else {
#pragma unroll
                                for (int wi = 0; wi < 8; ++wi)
                                    if (wi < n_w && wb + wi < p.words) dst[wi] = v[wi];
                            }
                        }
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + ew) : "memory");
                }
            }
            if (lists) {
                // tail of the segment (< 8 entries)
                for (int e = seg_flushed; e < seg_pos && e < p.seg_cap; ++e) {
                    const uint32_t* sl = s_lists + et * LBUF + (((((uint32_t)e * 4u) & (LBUF * 4 - 1)) ^ swz16) >> 2);
                    p.l_rows[seg_base + e] = sl[0];
                    p.l_k1[seg_base + e] = __uint_as_float(sl[RING_A_BYTES / 4]);
                    if (p.L > 2) p.l_k2[seg_base + e] = __uint_as_float(sl[2 * RING_A_BYTES / 4]);
                }
                p.seg_n[(int64_t)q * p.n_segs + range * 2 + half] = seg_pos;
            }
            if (p.c0_cnt && q_ok && c0_acc) atomicAdd(p.c0_cnt + q, (int)c0_acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F_TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------
// operand packing
// ---------------------------------------------------------------------------------------
struct PackParams {
    const float* idx;       // [N, Lsum]
    const float* rnorm;     // [N, L] (NaN = zero norm); null for queries
    int64_t N;
    hq_index_layout lay;
    int is_query;
    int seg_off[3], kp[3];
    float* out;             // [N, pitch]
    int pitch;              // packed floats per row (whole slabs)
    int thr_off;            // rows with padding in their fourth slab (ones_in_row): columns thr_off, thr_off + 1 hold 1.0 (the B side
                            // of the folded thresholds); -1 otherwise and for queries
};

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(256) k_pack_rows(const PackParams p) {
    const int64_t total = p.N * p.pitch;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / p.pitch;
        const int c = (int)(i - row * p.pitch);
        float outv = (p.thr_off >= 0 && (c == p.thr_off || c == p.thr_off + 1)) ? 1.0f : 0.f;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            if (l < p.lay.L && c >= p.seg_off[l] && c < p.seg_off[l] + 2 * p.kp[l]) {
                const int rel = c - p.seg_off[l];
                const int part = rel / p.kp[l], j = rel - part * p.kp[l];
                float x = j < p.lay.lvl_keff[l] ? __ldg(p.idx + row * p.lay.Lsum + p.lay.lvl_off[l] + j) : 0.f;
                if (!p.is_query) {
                    const float nrm = __ldg(p.rnorm + row * p.lay.L + l);
                    x = (nrm == nrm) ? __fdiv_rn(x, nrm) : 0.f;
                }
                const float hi = trunc_tf32(x);
                const float lo = trunc_tf32(__fadd_rn(x, -hi));
                outv = part == 1 ? lo : hi;
            }
        }
        p.out[i] = outv;
    }
}

// tq[l][q] = x*_l * |q_l| (sequential fmaf norm, like the CUDA-core path); NaN for a zero query level
__global__ void __launch_bounds__(256) k_query_tq(const float* __restrict__ q_idx, int Q, hq_index_layout lay, float x0, float x1, float x2,
                                                  float* __restrict__ tq, float* __restrict__ nq_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Q * lay.L) return;
    const int l = t / Q, q = t - l * Q;
    const float* r = q_idx + (int64_t)q * lay.Lsum + lay.lvl_off[l];
    float c = 0.f;
    for (int j = 0; j < lay.lvl_keff[l]; ++j) { const float v = __ldg(r + j); c = fmaf(v, v, c); }
    const float nq = sqrtf(c);
    const float xs = l == 0 ? x0 : (l == 1 ? x1 : x2);
    tq[(int64_t)l * Q + q] = nq > 0.f ? __fmul_rn(xs, nq) : __int_as_float(0x7fc00000);
    nq_out[(int64_t)l * Q + q] = nq;
}

// Window mode folds its lower thresholds into the contraction: the query operand carries three 8-column blocks after the
// level blocks -- [-t0], [-lo1], [-lo2], each as an exact tf32 (hi, lo) pair in its first two columns -- and every database
// row the pair (1, 1) at the same place, so one extra k-step per level makes the accumulator  dot - threshold  and the
// epilogue's test is the SIGN of the accumulator (one funnel shift per element instead of FADD + shift).  Runs between
// k_filter_predict and the window pass.  (Tried: five single-buffered accumulators with the upper edges folded as well, so
// that every test is a sign test: 1.83 ms instead of 1.39 ms -- 26 MMAs per tile from one issuing thread and no second
// accumulator set to hide them.)
__global__ void __launch_bounds__(256) k_fold_thresholds(const float* __restrict__ tq, const float* __restrict__ win, int Q, int L,
                                                         int thr_off, int pitch, float* __restrict__ q_packed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Q * 3) return;
    const int q = t / 3, b = t - q * 3;
    float v;
    if (b == 0) v = __ldg(tq + q);                                           // t0
    else if (b == 1) v = __ldg(win + (int64_t)0 * Q + q);                    // lo1
    else v = L > 2 ? __ldg(win + (int64_t)2 * Q + q) : 0.f;                  // lo2
    const float hi = trunc_tf32(v);
    const float lo = trunc_tf32(__fadd_rn(v, -hi));
    float* dst = q_packed + (int64_t)q * pitch + thr_off + 8 * b;
    dst[0] = -hi;
    dst[1] = -lo;
}

__global__ void __launch_bounds__(256) k_valid_bits(const float* __restrict__ rnorm, int64_t N, int L, uint32_t* __restrict__ valid,
                                                    int64_t valid_pitch) {
    const int64_t n_pad = valid_pitch * 32;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_pad * L; t += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(t / n_pad);
        const int64_t row = t - (int64_t)l * n_pad;
        bool ok = false;
        if (row < N) { const float v = __ldg(rnorm + row * L + l); ok = v == v; }
        const uint32_t w = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0) valid[(int64_t)l * valid_pitch + (row >> 5)] = w;
    }
}

void plan_units(int64_t n_tiles, int Q, int sms, FtcParams& p) {
    p.m_tiles = (Q + FM - 1) / FM;
    p.n_tiles = (int)n_tiles;
    int best_tp = p.n_tiles;
    double best_eff = -1.0;
    for (int waves = 1; waves <= 4; ++waves) {
        int s = (waves * sms) / p.m_tiles;
        if (s < 1) s = 1;
        int tiles_per = (p.n_tiles + s - 1) / s;
        tiles_per = (tiles_per + TILE_GROUP - 1) / TILE_GROUP * TILE_GROUP;     // flush groups stay sector aligned
        s = (p.n_tiles + tiles_per - 1) / tiles_per;
        const int units = s * p.m_tiles;
        const int rounds = (units + sms - 1) / sms;
        const double eff = (double)p.n_tiles * p.m_tiles / ((double)rounds * tiles_per * sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_tp = tiles_per; }
    }
    p.tiles_per_range = best_tp;
    p.n_ranges = (p.n_tiles + best_tp - 1) / best_tp;
    p.num_units = p.n_ranges * p.m_tiles;
}

}  // namespace

extern "C" int hq_filter_tc_supported(const hq_index_layout* layout) {
    Segs s;
    return make_segs(layout, s) ? 1 : 0;
}

extern "C" int hq_filter_tc_packed_cols(const hq_index_layout* layout) {
    Segs s;
    return make_segs(layout, s) ? s.n_slabs * 32 : 0;
}

extern "C" int64_t hq_filter_tc_valid_pitch(int64_t N) {
    // words, rounded up to whole 64-row tiles
    return (N + FR - 1) / FR * 2;
}

extern "C" int hq_filter_tc_pack(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout, int is_query,
                                 float* packed, void* stream) {
    Segs s;
    HQ_REQUIRE(make_segs(layout, s), "index layout not supported by the tensor-core filter");
    HQ_REQUIRE(N >= 0, "negative N");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(idx && packed && (is_query || rnorm), "null pointer");
    PackParams p{};
    p.idx = idx; p.rnorm = rnorm; p.N = N; p.lay = *layout; p.is_query = is_query ? 1 : 0; p.out = packed;
    p.pitch = (is_query ? s.n_slabs_q : s.n_slabs) * 32;
    p.thr_off = (!is_query && s.ones_in_row) ? s.thr_off : -1;
    for (int l = 0; l < 3; ++l) { p.seg_off[l] = s.off[l]; p.kp[l] = s.kp[l]; }
    int64_t blocks = (N * p.pitch + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    k_pack_rows<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    HQ_LAUNCH_OK("k_pack_rows");
    return HQ_OK;
}

extern "C" int hq_filter_tc_valid(const float* rnorm, int64_t N, const hq_index_layout* layout, uint32_t* valid, int64_t valid_pitch,
                                  void* stream) {
    HQ_REQUIRE(layout && layout->L >= 1 && layout->L <= 3, "bad index layout");
    HQ_REQUIRE(N >= 0 && valid_pitch >= hq_filter_tc_valid_pitch(N), "valid pitch too small");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(rnorm && valid, "null pointer");
    int64_t blocks = (valid_pitch * 32 * layout->L + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    k_valid_bits<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rnorm, N, layout->L, valid, valid_pitch);
    HQ_LAUNCH_OK("k_valid_bits");
    return HQ_OK;
}

extern "C" int hq_filter_tc_plan(int64_t N, int Q, int* n_ranges, int* tiles_per_range) {
    HQ_REQUIRE(N > 0 && Q > 0 && n_ranges && tiles_per_range, "bad arguments");
    FtcParams p{};
    plan_units((N + FR - 1) / FR, Q, hq_cached_sm_count(), p);
    *n_ranges = p.n_ranges;
    *tiles_per_range = p.tiles_per_range;
    return HQ_OK;
}

// the same for a pass that visits every tile_stride-th 64-row tile (sample pass)
int hq_filter_tc_plan_strided(int64_t N, int Q, int tile_stride, int* n_ranges, int* tiles_per_range) {
    HQ_REQUIRE(N > 0 && Q > 0 && tile_stride >= 1 && n_ranges && tiles_per_range, "bad arguments");
    FtcParams p{};
    const int64_t tiles = (N + FR - 1) / FR;
    plan_units((tiles + tile_stride - 1) / tile_stride, Q, hq_cached_sm_count(), p);
    *n_ranges = p.n_ranges;
    *tiles_per_range = p.tiles_per_range;
    return HQ_OK;
}

// Query-side preparation of a batch: packed operand [Q, 128], thresholds tq [3, Q] and level norms nq [3, Q].
int hq_filter_tc_prepare(const hq_index_layout* layout, const float* q_idx, int Q, const float* xstar, float* q_packed, float* tq,
                         float* nq, cudaStream_t st) {
    Segs s;
    HQ_REQUIRE(make_segs(layout, s), "index layout not supported by the tensor-core filter");
    HQ_REQUIRE(q_idx && q_packed && tq && nq && xstar, "null pointer");
    int rc = hq_filter_tc_pack(q_idx, nullptr, Q, layout, 1, q_packed, st);
    if (rc != HQ_OK) return rc;
    k_query_tq<<<(Q * layout->L + 255) / 256, 256, 0, st>>>(q_idx, Q, *layout, xstar[0], layout->L > 1 ? xstar[1] : 0.f,
                                                            layout->L > 2 ? xstar[2] : 0.f, tq, nq);
    HQ_LAUNCH_OK("k_query_tq");
    return HQ_OK;
}

// Window mode: write the folded thresholds (-t0, -lo1, -lo2, -hi1, -hi2 as tf32 pairs) into the query operand.
int hq_filter_tc_fold(const hq_index_layout* layout, const float* tq, const float* win, int Q, float* q_packed, cudaStream_t st) {
    Segs s;
    HQ_REQUIRE(make_segs(layout, s) && s.L >= 2, "index layout not supported by the window mode");
    HQ_REQUIRE(tq && win && q_packed && Q > 0, "bad arguments");
    k_fold_thresholds<<<(Q * 3 + 255) / 256, 256, 0, st>>>(tq, win, Q, s.L, s.thr_off, s.n_slabs_q * 32, q_packed);
    HQ_LAUNCH_OK("k_fold_thresholds");
    return HQ_OK;
}

// The threshold pass over a prepared batch.  bits may be null (sample pass: lists and the level-0 count only); `o`
// selects the sample stride, the query-tile subset and the window mode (see FtcParams).
int hq_filter_tc_pass(const float* db_packed, const uint32_t* valid, int64_t valid_pitch, int64_t N, const hq_index_layout* layout, int Q,
                      const float* q_packed, const float* tq, uint32_t* bits, int64_t bits_pitch, const HqFilterLists* lists,
                      const HqFtcOpts* o, cudaStream_t st) {
    Segs s;
    HQ_REQUIRE(make_segs(layout, s), "index layout not supported by the tensor-core filter");
    HQ_REQUIRE(db_packed && valid && q_packed && tq, "null pointer");
    HQ_REQUIRE((reinterpret_cast<uintptr_t>(db_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_packed) & 15) == 0,
               "packed operands must be 16-byte aligned");
    HQ_REQUIRE(valid_pitch >= hq_filter_tc_valid_pitch(N), "valid pitch too small");
    const int64_t words = (N + 31) / 32;
    HQ_REQUIRE(!bits || bits_pitch >= words, "bit plane pitch too small");
    const int stride = o && o->tile_stride > 1 ? o->tile_stride : 1;
    const bool win = o && o->win;
    HQ_REQUIRE(!(stride > 1 && bits), "the sample pass does not write planes");
    HQ_REQUIRE(!win || (s.L >= 2 && stride == 1 && bits && o->wcnt && lists && lists->rows && (s.L < 3 || lists->k2)),
               "window mode needs two or three levels, a plane, counters and lists");

    FtcParams p{};
    p.N = N; p.Q = Q; p.L = s.L; p.n_slabs = s.n_slabs; p.n_slabs_q = s.n_slabs_q; p.thr_off = s.thr_off; p.ones_in_row = s.ones_in_row;
    for (int l = 0; l < 3; ++l) { p.seg_off[l] = s.off[l]; p.ksteps[l] = s.kp[l] / 8; }
    p.tq = tq; p.valid = valid; p.valid_pitch = valid_pitch; p.bits = bits; p.words = words; p.bits_pitch = bits_pitch;
    p.bits_vec = (bits && (reinterpret_cast<uintptr_t>(bits) & 31) == 0 && bits_pitch % 8 == 0) ? 1 : 0;
    p.tile_stride = stride;
    if (o) { p.unit_only = o->unit_only; p.c0_cnt = o->c0_cnt; p.win = o->win; p.wcnt = o->wcnt; }
    const int64_t tiles = (N + FR - 1) / FR;
    plan_units((tiles + stride - 1) / stride, Q, hq_cached_sm_count(), p);
    if (lists && lists->rows) {
        HQ_REQUIRE(lists->n_segs >= 2 * p.n_ranges && lists->n_segs <= 2 * p.n_ranges + 1 && lists->seg_cap > 0 && lists->seg_cap % 8 == 0 &&
                       lists->k1 && lists->seg_n && (s.L < 3 || lists->k2),
                   "candidate list geometry does not match hq_filter_tc_plan");
        p.l_rows = lists->rows; p.l_k1 = lists->k1; p.l_k2 = lists->k2; p.seg_n = lists->seg_n; p.seg_cap = lists->seg_cap;
        p.n_segs = lists->n_segs;
    }
    CUtensorMap mq, mdb;
    const int ks = s.n_slabs * 32, ksq = s.n_slabs_q * 32;          // packed floats per database / query row
    int rc = make_map_2d(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, q_packed, Q, ksq, ksq, 32, FM);
    if (rc != HQ_OK) return rc;
    rc = make_map_2d(&mdb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, db_packed, N, ks, ks, 32, FR);
    if (rc != HQ_OK) return rc;
    const size_t smem_full = 1024 + A_BYTES + F_STAGES * B_STAGE_BYTES + (STAGE_WORDS + 3 * LBUF * EPI_THREADS) * sizeof(uint32_t) +
                             (2 * F_STAGES + 6) * sizeof(uint64_t) + 16;
    const size_t smem_win = 1024 + A_BYTES + F_STAGES * B_STAGE_BYTES + (STAGE_WORDS / 3 + 3 * LBUF_W * EPI_THREADS) * sizeof(uint32_t) +
                            KSTAGE_BYTES + (2 * F_STAGES + 6) * sizeof(uint64_t) + 16;
    static bool attr = false;
    if (!attr) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_filter_bits_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full));
        HQ_CUDA_OK(cudaFuncSetAttribute(k_filter_bits_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_win));
        attr = true;
    }
    int grid = hq_cached_sm_count();
    if (grid > p.num_units) grid = p.num_units;
    const int tk = (win || (stride == 1 && !p.unit_only)) ? hq_time_begin(1, st) : -1;      // the main pass, not sample / fallback
    if (win) k_filter_bits_tc<true><<<grid, F_THREADS, smem_win, st>>>(mq, mdb, p);
    else k_filter_bits_tc<false><<<grid, F_THREADS, smem_full, st>>>(mq, mdb, p);
    hq_time_end(1, tk, st);
    HQ_LAUNCH_OK("k_filter_bits_tc");
    return HQ_OK;
}

// Bit planes (and optionally candidate lists) for a query batch.  q_packed [Q, 128], tq [3, Q] and
// nq [3, Q] are caller-provided scratch; lists may be null.
int hq_filter_bits_tc_launch(const float* db_packed, const uint32_t* valid, int64_t valid_pitch, int64_t N,
                             const hq_index_layout* layout, const float* q_idx, int Q, const float* xstar, float* q_packed,
                             float* tq, float* nq, uint32_t* bits, int64_t bits_pitch, const HqFilterLists* lists, cudaStream_t st) {
    HQ_REQUIRE(bits, "null pointer");
    int rc = hq_filter_tc_prepare(layout, q_idx, Q, xstar, q_packed, tq, nq, st);
    if (rc != HQ_OK) return rc;
    return hq_filter_tc_pass(db_packed, valid, valid_pitch, N, layout, Q, q_packed, tq, bits, bits_pitch, lists, nullptr, st);
}
