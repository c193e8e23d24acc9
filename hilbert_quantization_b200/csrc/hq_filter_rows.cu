// K5r: the WINDOW pass of the fast filter (rag/search/engine.py:178-287, see k_filter_bits_tc<true> in hq_filter_tc.cu) for
// a HANDFUL of queries, on the CUDA cores.
//
// The tensor-core pass contracts 128 queries x 64 rows per tile whatever the batch size: for one query it is bound by its
// own MMA / TMEM round trip per tile (105 us against 1 M x 1536 rows, 384 MB of [hi | lo] operand at 3.6 TB/s) with 127
// of the 128 accumulator lanes unused.  A single query needs 34 multiply-adds per row, so here the rows stream through
// shared memory as plain fp32 -- a compact copy of the index rows with every level scaled by 1 / |c_l| and padded to whole
// float4 (36 floats = 144 bytes per 1536-D row, 16 floats for 768-D: 2.7x fewer bytes than the split operand) -- and a
// thread scores its row against every query of the batch:
//   CTA = one row range of the tensor pass's plan (the list segments keep their meaning), 512-row tiles arrive by bulk
//   copies (cp.async.bulk + mbarrier, two to four buffers); lane = row, so the five class words of a 32-row group are warp
//   ballots and the counters, the alive word and the appends of the window rows need no bit walking.
// Outputs exactly what the tensor pass writes in window mode (alive plane, four counters per query, lists of the window
// rows with keys relative to the lower window edges, segment counts); the dot products are fp32 FMA chains over the
// scaled rows (the tensor pass: tf32 hi/lo products, error < 3e-7), both within the 2e-6 band of the filter's parity bar.
#include "hq_tc.cuh"

namespace {

constexpr int R_THREADS = 512;         // 16 warps: with one warp per scheduler (128 threads, 128-row tiles) every LDS / FMA latency of a
constexpr int R_TILE = 512;            // row's chain was exposed: 64 us per query against 1 M x 1536 rows.  Rows per tile = threads
constexpr int R_BUFS = 4;              // most tile buffers (as many as fit 200 KB, at least two)
constexpr int R_MAX_Q = 8;

struct RowsParams {
    const float* rows;          // [N][KC] scaled compact index rows
    int64_t N;
    int KC, L;
    int blk[4];                 // first float4 of level l inside a compact row; blk[L] = KC / 4
    int q_off[3], keff[3];      // level l inside a q_idx row (layout offsets), its width
    int Lsum;
    const float* q_idx;         // [Q][Lsum]
    int Q;
    const float* tq;            // [3][Q]
    const float* win;           // [4][Q]
    const uint32_t* valid;      // [L][valid_pitch]
    int64_t valid_pitch;
    uint32_t* bits;             // alive plane [Q][bits_pitch]
    int64_t bits_pitch, words;
    int32_t* wcnt;              // [4][Q]
    uint32_t* l_rows;
    float* l_k1;
    float* l_k2;
    int32_t* seg_n;
    int64_t seg_cap;
    int n_segs;
    int n_ranges, tiles_per_range;      // the tensor pass's plan (64-row tiles); segment = range * 2 + half
    int nbuf;                           // tile buffers in shared memory
};

__global__ void __launch_bounds__(R_THREADS) k_filter_win_rows(const RowsParams p) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t s_full[R_BUFS];
    __shared__ uint32_t s_cnt[R_MAX_Q][2];
    __shared__ float s_thr[R_MAX_Q][6];         // t0, lo1, w1, lo2, w2, ok
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = warp & 1;
    const int KC4 = p.KC >> 2;
    const uint32_t tile_bytes = (uint32_t)R_TILE * (uint32_t)p.KC * 4u;
    float4* qs = reinterpret_cast<float4*>(sm_raw);                                   // [Q][KC4], zero padded
    unsigned char* bufs = sm_raw + (((size_t)p.Q * p.KC * 4 + 127) & ~(size_t)127);   // R_BUFS tiles
    const bool three = p.L > 2;
    for (int i = tid; i < p.Q * p.KC; i += R_THREADS) {
        const int q = i / p.KC, c = i - q * p.KC;
        float v = 0.f;
        for (int l = 0; l < p.L; ++l) {
            const int j = c - 4 * p.blk[l];
            if (j >= 0 && j < p.keff[l] && c < 4 * p.blk[l + 1]) v = __ldg(p.q_idx + (int64_t)q * p.Lsum + p.q_off[l] + j);
        }
        reinterpret_cast<float*>(qs)[i] = v;
    }
    if (tid < p.Q) {
        const int q = tid;
        const float t0 = __ldg(p.tq + q);
        const float lo1 = __ldg(p.win + q), hi1 = __ldg(p.win + (int64_t)p.Q + q);
        const float lo2 = __ldg(p.win + (int64_t)2 * p.Q + q), hi2 = __ldg(p.win + (int64_t)3 * p.Q + q);
        const bool ok = (t0 == t0) && (lo1 == lo1) && (hi1 == hi1) && (!three || ((lo2 == lo2) && (hi2 == hi2)));
        s_thr[q][0] = t0; s_thr[q][1] = lo1; s_thr[q][2] = hi1 - lo1;
        s_thr[q][3] = three ? lo2 : 0.f; s_thr[q][4] = three ? hi2 - lo2 : 0.f; s_thr[q][5] = ok ? 1.f : 0.f;
        s_cnt[q][0] = 0; s_cnt[q][1] = 0;
    }
    if (tid < R_BUFS) hq_tc::mbar_init(s_full + tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int range = blockIdx.x;
    const int64_t row_begin = (int64_t)range * p.tiles_per_range * 64;
    int64_t row_end = row_begin + (int64_t)p.tiles_per_range * 64;
    const int64_t n_pad = (p.N + 63) / 64 * 64;
    if (row_end > n_pad) row_end = n_pad;
    const int n_tiles = row_end > row_begin ? (int)((row_end - row_begin + R_TILE - 1) / R_TILE) : 0;
    const int NB = p.nbuf;
    auto issue = [&](int t) {                      // thread 0: tile t of this range into buffer t % NB (rows clipped at N)
        const int64_t r0 = row_begin + (int64_t)t * R_TILE;
        int64_t nr = p.N - r0;
        if (nr > R_TILE) nr = R_TILE;
        uint64_t* bar = s_full + (t % NB);
        if (nr <= 0) { hq_tc::mbar_arrive(bar); return; }
        const uint32_t bytes = (uint32_t)nr * (uint32_t)p.KC * 4u;
        hq_tc::mbar_expect_tx(bar, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         hq_tc::smem_u32(bufs + (size_t)(t % NB) * tile_bytes)),
                     "l"(p.rows + r0 * p.KC), "r"(bytes), "r"(hq_tc::smem_u32(bar))
                     : "memory");
    };
    if (tid == 0)
        for (int t = 0; t < NB - 1 && t < n_tiles; ++t) issue(t);

    // counters of this warp: lane q holds query q's four
    uint32_t c_c0 = 0, c_a1 = 0, c_a1a2 = 0, c_al = 0;
    const uint32_t lower = (1u << lane) - 1u;
    const int seg = range * 2 + half;
    for (int t = 0; t < n_tiles; ++t) {
        if (tid == 0 && t + NB - 1 < n_tiles) issue(t + NB - 1);               // into the buffer released by the barrier below
        const int64_t g0 = row_begin + (int64_t)t * R_TILE + 32 * warp;          // first row of this warp's group
        const bool in_range = g0 < row_end;
        uint32_t vw[3] = {0u, 0u, 0xffffffffu};
        if (in_range) {
            vw[0] = __ldg(p.valid + (g0 >> 5));
            vw[1] = __ldg(p.valid + p.valid_pitch + (g0 >> 5));
            if (three) vw[2] = __ldg(p.valid + 2 * p.valid_pitch + (g0 >> 5));
        }
        hq_tc::mbar_wait(s_full + (t % NB), (uint32_t)(t / NB) & 1u);
        if (in_range) {
            const float4* r4 = reinterpret_cast<const float4*>(bufs + (size_t)(t % NB) * tile_bytes) + (size_t)(32 * warp + lane) * KC4;
            const uint32_t row = (uint32_t)(g0 + lane);
            for (int q = 0; q < p.Q; ++q) {
                float d0 = 0.f, d1 = 0.f, d2 = 0.f;
                const float4* q4 = qs + q * KC4;
                for (int j = 0; j < KC4; ++j) {
                    const float4 c = r4[j], a = q4[j];
                    if (j < p.blk[1]) { d0 = fmaf(a.x, c.x, d0); d0 = fmaf(a.y, c.y, d0); d0 = fmaf(a.z, c.z, d0); d0 = fmaf(a.w, c.w, d0); }
                    else if (j < p.blk[2]) { d1 = fmaf(a.x, c.x, d1); d1 = fmaf(a.y, c.y, d1); d1 = fmaf(a.z, c.z, d1); d1 = fmaf(a.w, c.w, d1); }
                    else { d2 = fmaf(a.x, c.x, d2); d2 = fmaf(a.y, c.y, d2); d2 = fmaf(a.z, c.z, d2); d2 = fmaf(a.w, c.w, d2); }
                }
                const bool ok = s_thr[q][5] != 0.f;
                const float k0 = d0 - s_thr[q][0], k1 = d1 - s_thr[q][1], k2 = d2 - s_thr[q][3];
                const float w1 = s_thr[q][2], w2 = s_thr[q][4];
                const uint32_t W0 = __ballot_sync(0xffffffffu, ok && k0 >= 0.f) & vw[0];
                const uint32_t A1 = __ballot_sync(0xffffffffu, ok && k1 >= 0.f) & vw[1];
                const uint32_t B1 = __ballot_sync(0xffffffffu, ok && (k1 - w1) >= 0.f) & A1;
                uint32_t A2 = 0xffffffffu, B2 = 0xffffffffu;
                if (three) {
                    A2 = __ballot_sync(0xffffffffu, ok && k2 >= 0.f) & vw[2];
                    B2 = __ballot_sync(0xffffffffu, ok && (k2 - w2) >= 0.f) & A2;
                }
                const uint32_t S1 = W0 & B1;
                const uint32_t alive = S1 & B2;
                const uint32_t E = W0 & A1 & (~B1 | (A2 & ~B2));
                if (lane == q) {
                    c_c0 += (uint32_t)__popc(W0); c_a1 += (uint32_t)__popc(S1);
                    c_a1a2 += (uint32_t)__popc(S1 & A2); c_al += (uint32_t)__popc(alive);
                    if ((g0 >> 5) < p.words) p.bits[(int64_t)q * p.bits_pitch + (g0 >> 5)] = alive;
                }
                if (E && p.l_rows) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&s_cnt[q][half], (uint32_t)__popc(E));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if ((E >> lane) & 1u) {
                        const int64_t slot = (int64_t)base + __popc(E & lower);
                        if (slot < p.seg_cap) {
                            const int64_t a = ((int64_t)q * p.n_segs + seg) * p.seg_cap + slot;
                            p.l_rows[a] = row;
                            p.l_k1[a] = k1;
                            if (three) p.l_k2[a] = k2;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (lane < p.Q && p.wcnt) {
        if (c_c0) atomicAdd(p.wcnt + lane, (int)c_c0);
        if (c_a1) atomicAdd(p.wcnt + (int64_t)p.Q + lane, (int)c_a1);
        if (c_a1a2) atomicAdd(p.wcnt + (int64_t)2 * p.Q + lane, (int)c_a1a2);
        if (c_al) atomicAdd(p.wcnt + (int64_t)3 * p.Q + lane, (int)c_al);
    }
    if (tid < 2 * p.Q && p.seg_n) {
        const int q = tid >> 1, h = tid & 1;
        p.seg_n[(int64_t)q * p.n_segs + range * 2 + h] = (int32_t)s_cnt[q][h];
    }
}

struct RowPackParams {
    const float* idx;           // [N][Lsum]
    const float* rnorm;         // [N][L] level norms (NaN / 0: no such level)
    int64_t N;
    hq_index_layout lay;
    int KC;
    int blk[4];
    float* out;                 // [N][KC]
};

__global__ void __launch_bounds__(256) k_filter_rows_pack(const RowPackParams p) {
    const int64_t total = p.N * p.KC;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / p.KC;
        const int c = (int)(i - row * p.KC);
        float v = 0.f;
        for (int l = 0; l < p.lay.L; ++l) {
            const int j = c - 4 * p.blk[l];
            if (j >= 0 && j < p.lay.lvl_keff[l] && c < 4 * p.blk[l + 1]) {
                const float nrm = __ldg(p.rnorm + row * p.lay.L + l);
                const float x = __ldg(p.idx + row * p.lay.Lsum + p.lay.lvl_off[l] + j);
                v = (nrm == nrm) ? __fdiv_rn(x, nrm) : 0.f;          // the scaling of the tensor pass's operand (k_pack_rows)
            }
        }
        p.out[i] = v;
    }
}

bool rows_blocks(const hq_index_layout* lay, int (&blk)[4], int& KC) {
    if (!lay || lay->L < 2 || lay->L > 3) return false;
    int off4 = 0;
    for (int l = 0; l < 4; ++l) blk[l] = 0;
    for (int l = 0; l < lay->L; ++l) {
        if (lay->lvl_keff[l] < 1) return false;
        blk[l] = off4;
        off4 += (lay->lvl_keff[l] + 3) / 4;
    }
    for (int l = lay->L; l < 4; ++l) blk[l] = off4;
    KC = off4 * 4;
    return KC <= 128;
}

}  // namespace

extern "C" int hq_filter_rows_max_queries(void) { return R_MAX_Q; }

extern "C" int hq_filter_rows_cols(const hq_index_layout* layout) {
    int blk[4], KC = 0;
    return rows_blocks(layout, blk, KC) ? KC : 0;
}

extern "C" int hq_filter_rows_pack(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout, float* rows,
                                   void* stream) {
    RowPackParams p{};
    HQ_REQUIRE(rows_blocks(layout, p.blk, p.KC), "index layout not supported by the row pass");
    HQ_REQUIRE(N >= 0, "negative size");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(idx && rnorm && rows, "null pointer");
    p.idx = idx; p.rnorm = rnorm; p.N = N; p.lay = *layout; p.out = rows;
    int64_t blocks = (N * p.KC + 255) / 256;
    const int64_t cap = (int64_t)hq_cached_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    k_filter_rows_pack<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    HQ_LAUNCH_OK("k_filter_rows_pack");
    return HQ_OK;
}

// The window pass over scaled compact rows (hq_filter_rows_pack): same outputs as hq_filter_tc_pass in window mode.
int hq_filter_rows_pass(const float* rows, const uint32_t* valid, int64_t valid_pitch, int64_t N, const hq_index_layout* layout,
                        const float* q_idx, int Q, const float* tq, const float* win, uint32_t* alive, int64_t alive_pitch,
                        int32_t* wcnt, const HqFilterLists* lists, cudaStream_t st) {
    RowsParams p{};
    HQ_REQUIRE(rows_blocks(layout, p.blk, p.KC), "index layout not supported by the row pass");
    HQ_REQUIRE(Q >= 1 && Q <= R_MAX_Q && N > 0, "the row pass takes 1..%d queries", R_MAX_Q);
    HQ_REQUIRE(rows && valid && q_idx && tq && win && alive && wcnt && lists && lists->rows && lists->k1 && lists->seg_n, "null pointer");
    HQ_REQUIRE((reinterpret_cast<uintptr_t>(rows) & 15) == 0, "rows must be 16-byte aligned");
    int nr = 0, tp = 0;
    int rc = hq_filter_tc_plan(N, Q, &nr, &tp);
    if (rc != HQ_OK) return rc;
    HQ_REQUIRE(lists->n_segs >= 2 * nr, "list geometry does not match the pass plan");
    p.rows = rows; p.N = N; p.L = layout->L; p.Lsum = layout->Lsum;
    for (int l = 0; l < 3; ++l) { p.q_off[l] = layout->lvl_off[l]; p.keff[l] = l < layout->L ? layout->lvl_keff[l] : 0; }
    p.q_idx = q_idx; p.Q = Q; p.tq = tq; p.win = win; p.valid = valid; p.valid_pitch = valid_pitch;
    p.bits = alive; p.bits_pitch = alive_pitch; p.words = (N + 31) / 32; p.wcnt = wcnt;
    p.l_rows = lists->rows; p.l_k1 = lists->k1; p.l_k2 = layout->L > 2 ? lists->k2 : nullptr;
    p.seg_n = lists->seg_n; p.seg_cap = lists->seg_cap; p.n_segs = lists->n_segs;
    p.n_ranges = nr; p.tiles_per_range = tp;
    HQ_REQUIRE(layout->L < 3 || p.l_k2, "null pointer");
    const size_t qbytes = ((size_t)Q * p.KC * 4 + 127) & ~(size_t)127, tile = (size_t)R_TILE * p.KC * 4;
    int nbuf = (int)((200 * 1024 - qbytes) / tile);
    if (nbuf > R_BUFS) nbuf = R_BUFS;
    HQ_REQUIRE(nbuf >= 2, "rows too wide for the row pass");
    p.nbuf = nbuf;
    const size_t smem = qbytes + (size_t)nbuf * tile;
    static bool attr = false;
    if (!attr) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_filter_win_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    k_filter_win_rows<<<nr, R_THREADS, smem, st>>>(p);
    HQ_LAUNCH_OK("k_filter_win_rows");
    return HQ_OK;
}
