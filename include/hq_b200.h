/*
 * hq_b200.h -- C ABI of libhq_b200.so, the sm_100a implementation of the
 * hilbert-quantization hot path (Hilbert map/unmap -> hierarchical index (+u8
 * quantise) -> progressive search).
 *
 * Rules of the boundary
 *   - every entry point is extern "C", returns an int status (0 = HQ_OK,
 *     negative = error, text via hq_last_error()), never aborts, never
 *     allocates device memory, never takes ownership;
 *   - all data pointers are DEVICE pointers unless the name says host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     calls are asynchronous on that stream;
 *   - element strides are counted in elements, not bytes.
 *
 * Each declaration cites the reference interface it replaces (paths relative
 * to the reference checkout, Tylerlhess/hilbert-quantization v1.3.0).
 */
#ifndef HQ_B200_H
#define HQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HQ_OK            0
#define HQ_EINVAL       (-1)   /* bad argument (message in hq_last_error)      */
#define HQ_ECUDA        (-2)   /* CUDA runtime error                           */
#define HQ_EUNSUPPORTED (-3)   /* valid request this build does not implement  */

#define HQ_ABI_VERSION  1

/* ---- library ---------------------------------------------------------- */
int         hq_version(void);
const char* hq_last_error(void);              /* thread-local, never NULL   */
int         hq_sm_count(void);                /* SMs of the current device  */
int64_t     hq_launch_count(int reset);       /* kernels launched by this library so far */

/* ---- a1/a2: Hilbert coordinates ---------------------------------------
 * hilbert_quantization/core/hilbert_mapper.py:17-66 (generate_hilbert_coordinates,
 * _hilbert_index_to_xy) and :68-90 (_xy_to_hilbert_index); RAG twin
 * rag/embedding_generation/hilbert_mapper.py:122-230.
 * x[i], y[i] = d2xy(n, d0 + i);  d[i] = xy2d(n, x[i], y[i]).  n power of two. */
int hq_d2xy_batch(int n, int64_t d0, int64_t count, int32_t* x, int32_t* y, void* stream);
int hq_xy2d_batch(int n, const int32_t* x, const int32_t* y, int64_t count, int64_t* d, void* stream);

/* ---- a3/a4: map_to_2d / map_from_2d (pure permutation + zero pad) -----
 * core/hilbert_mapper.py:115-174 (map_to_2d) and :176-205 (map_from_2d); RAG twin
 * rag/embedding_generation/hilbert_mapper.py:16-120.  Any element width in
 * {1,2,4,8} bytes (dtype preserved, tests/test_hilbert_mapper.py:359-378).
 *   map_to_2d  : src [N, D] (row stride src_stride) -> dst [N, n, n] (item stride dst_stride)
 *   map_from_2d: src [N, n, n]                      -> dst [N, D_out] (first D_out curve cells) */
int hq_map_to_2d(const void* src, int64_t N, int64_t D, int64_t src_stride, int n, int elem_bytes,
                 void* dst, int64_t dst_stride, void* stream);
int hq_map_from_2d(const void* src, int64_t N, int n, int64_t src_stride, int64_t D_out, int elem_bytes,
                   void* dst, int64_t dst_stride, void* stream);

/* ---- a3/a4 + a6/a7/a8 fused: one pass over fp32 items -------------------
 * Replaces the map_to_2d -> generate_*_indices -> embed sequence of
 * core/pipeline.py:129-140 and the RAG sequence
 * rag/embedding_generation/hilbert_mapper.py:16 ->
 * rag/embedding_generation/hierarchical_index_generator.py:103 (variant C),
 * core/streaming_index_builder.py:315 (variant B), core/index_generator.py:313 (variant A).
 *
 * direction 0: src = streams [N, D]  (map):   optional grid_out
 * direction 1: src = grids  [N, n, n] (unmap): optional stream_out (first D values)
 * Index values are a gather (`plan`, length plan_len, device int32) out of the
 * per-item run-mean pyramid: plan[i] < 0 -> 0.0; plan[i] < n*n -> grid cell
 * plan[i] (row-major); otherwise pyramid level k>=1, position j encoded as
 * n*n + level_base(k) + j with level_base(1)=0, level_base(k+1)=level_base(k)+n*n/4^k.
 * pyr_mode 0: float32 pairwise tree means -> idx_out float32 (variants A, C);
 * pyr_mode 1: float64 strict ((a+b)+c)+d)*0.25 -> idx_out float64 (variant B, bit-exact).
 * Any of grid_out / stream_out / idx_out may be NULL.  n <= 64 runs as one
 * fused launch; larger n is tiled 64x64 and finished by a second small launch
 * that needs `scratch` (hq_fused_scratch_bytes). */
int64_t hq_fused_scratch_bytes(int64_t N, int n, int pyr_mode);
/* same, when the caller knows the lowest pyramid level (>= 1) its plan references */
int64_t hq_fused_scratch_bytes_min_level(int64_t N, int n, int pyr_mode, int min_level);
int hq_map_index_fused_ml(const float* src, int direction, int64_t N, int64_t D, int64_t src_stride, int n,
                          float* grid_out, int64_t grid_stride,
                          float* stream_out, int64_t stream_stride,
                          const int32_t* plan, int plan_len, int pyr_mode, int min_level,
                          void* idx_out, int64_t idx_stride,
                          void* scratch, int64_t scratch_bytes, void* stream);
/* Parameter stream (BASELINE config 4; core/streaming_processor.py:539-582 chunking, core/pipeline.py:325-349
 * zero padding, then map_to_2d + index per chunk): `total` consecutive float32 values -> ceil(total / n^2) grids of
 * n x n (n > 64, src and grid_out 16-byte aligned, total % 4 == 0), the last grid zero padded, ONE launch.
 * grid_out [N, n, n] dense (or NULL); idx_out / plan / scratch as for hq_map_index_fused_ml with N = ceil(total / n^2). */
int hq_map_index_stream(const float* src, int64_t total, int n, float* grid_out, const int32_t* plan, int plan_len,
                        int pyr_mode, int min_level, void* idx_out, int64_t idx_stride, void* scratch,
                        int64_t scratch_bytes, void* stream);
int hq_map_index_fused(const float* src, int direction, int64_t N, int64_t D, int64_t src_stride, int n,
                       float* grid_out, int64_t grid_stride,
                       float* stream_out, int64_t stream_stride,
                       const int32_t* plan, int plan_len, int pyr_mode,
                       void* idx_out, int64_t idx_stride,
                       void* scratch, int64_t scratch_bytes, void* stream);

/* ---- a8 general: block means over arbitrary (H, W) images -------------
 * rag/embedding_generation/hierarchical_index_generator.py:204-244 for images
 * that are not power-of-two squares.  out[item, s] = mean(img[r0:r0+sh, c0:c0+sw])
 * with (r0, c0) = (rows[s]*sh, cols[s]*sw). */
int hq_block_means(const float* img, int64_t N, int H, int W, int64_t img_stride, int sh, int sw,
                   const int32_t* rows, const int32_t* cols, int count,
                   float* out, int64_t out_stride, void* stream);

/* ---- a3 + a6/a7/a8 + a9 + a10 fused: map_to_2d, hierarchical index, embed, uint8 normalise in ONE launch ----
 * core/pipeline.py:129-146 (map_to_2d -> generate_optimized_indices -> embed_indices_in_image) followed by
 * core/compressor.py:256-280 (_normalize_for_compression of the enhanced frame).  Per item: 4 D bytes in;
 * n^2 + plan_len uint8 frame bytes (grid rows, then the index slots of `plan` in order: one row of S = n values for
 * variants A / B, L rows of n for variant C), the (min, max) pair and optionally the float index values out.
 * The frame's min / max are reduced from the grid cells (every index value is a cell or a mean of cells, and rounding
 * is monotonic) plus 0 when `frame_zero` (padded index slots) or D < n^2 (padding cells): bit-identical to quantising
 * the materialised frame with hq_quantize_u8.  Grids of 32 x 32 and 64 x 64, rows of D % 4 == 0 values, 16-byte aligned;
 * other shapes return HQ_EUNSUPPORTED (run hq_map_index_fused + hq_quantize_u8).  u8_stride: bytes per frame, % 16 == 0. */
int hq_map_index_quant(const float* src, int64_t N, int64_t D, int64_t src_stride, int n, const int32_t* plan, int plan_len,
                       int pyr_mode, int min_level, int frame_zero, uint8_t* u8_out, int64_t u8_stride, float* minmax,
                       void* idx_out, int64_t idx_stride, void* stream);

/* ---- a10: uint8 quantise / dequantise ---------------------------------
 * core/compressor.py:256-280 (_normalize_for_compression: truncating cast,
 * constant image -> 128) and :282-303 (_denormalize_from_compression).
 * minmax [N, 2] float32 is written by quantise and read by dequantise. */
int hq_quantize_u8(const float* src, int64_t N, int64_t elems, int64_t src_stride,
                   uint8_t* dst, int64_t dst_stride, float* minmax, void* stream);
int hq_dequantize_u8(const uint8_t* src, int64_t N, int64_t elems, int64_t src_stride,
                     const float* minmax, float* dst, int64_t dst_stride, void* stream);

/* ---- a12: RAG progressive filter ---------------------------------------
 * rag/search/engine.py:51-95 (progressive_hierarchical_search), :178-241
 * (_filter_candidates_at_level), :243-287 (_apply_progressive_threshold),
 * :1025-1051 (_compare_single_level_indices).
 *
 * Index rows are stored compact: idx [N, Lsum] float32, level l occupying
 * columns [lvl_off[l], lvl_off[l] + lvl_w[l]); lvl_keff[l] <= lvl_w[l] bounds the
 * non-zero prefix of every row and query of the level (columns beyond it are
 * structurally zero because d >= D is padding).  lens [N, L] uint16 are the
 * trailing-zero-stripped lengths (>= 1) written by hq_index_row_lengths.
 *
 * hq_filter_level scores ONE level for Q queries against N rows, restricted to
 * rows alive in mask_in (bit `row` of the query's mask_stride uint32 words;
 * NULL = all alive):
 *   score = (cos(q[:m], c[:m]) + 1) / 2,  m = min(len_q, len_c), 0 if a norm is 0
 * It writes scores [Q, N] (float32, -1 for dead rows), the tentative survivor
 * mask (alive AND score >= thr) into mask_out, and per-query counts n_alive[q],
 * n_pass[q] (both must be zeroed by the caller).
 * hq_filter_select applies the ratio cut: cap[q] = max(1, int(n_alive[q]*ratio));
 * queries with n_pass <= cap keep the tentative mask, the others keep exactly the
 * cap best rows (score desc, ties -> lower row id) found by an exact radix select. */
typedef struct hq_index_layout {
    int32_t L;              /* number of levels (<= 8)              */
    int32_t Lsum;           /* row pitch of idx in floats           */
    int32_t lvl_off[8];
    int32_t lvl_w[8];
    int32_t lvl_keff[8];
} hq_index_layout;

int hq_index_row_lengths(const float* idx, int64_t N, const hq_index_layout* layout,
                         uint16_t* lens, void* stream);
int hq_filter_level(const float* idx, const uint16_t* lens, int64_t N, const hq_index_layout* layout, int level,
                    const float* q_idx, const uint16_t* q_lens, int Q,
                    const uint32_t* mask_in, int64_t mask_stride, double thr,
                    float* scores, int64_t scores_stride, uint32_t* mask_out,
                    int32_t* n_alive, int32_t* n_pass, void* stream);
int hq_filter_select(const float* scores, int64_t scores_stride, int64_t N, int Q,
                     const int32_t* n_alive, const int32_t* n_pass, double ratio,
                     uint32_t* mask, int64_t mask_stride, int32_t* n_out, void* stream);
/* The same cut with the reference's tie rule (stable sort of a list that arrives in the previous level's order,
 * rag/search/engine.py:236): rows tied exactly at the cut score are kept in the order of the previous level's score
 * (prev1 [Q, prev_stride], level l - 1), then of the level before (prev0, level l - 2), then by row id.  prev1 / prev0 may
 * be null (first levels). */
int hq_filter_select_prev(const float* scores, int64_t scores_stride, int64_t N, int Q, const float* prev1, const float* prev0,
                          int64_t prev_stride, const int32_t* n_alive, const int32_t* n_pass, double ratio, uint32_t* mask,
                          int64_t mask_stride, int32_t* n_out, void* stream);

/* Fast path of the same filter (no score matrix): one pass computes the threshold tests of
 * all levels as bit planes, then one CTA per query walks the levels and ranks rows exactly
 * (fp32 scores recomputed from the index rows, exact radix select, ties -> lower row id) only
 * where the ratio cut binds.  Requires L <= 3, level widths multiples of 4 and every stored /
 * query index length equal to lvl_keff (hq_filter_level_norms reports violations through
 * *nonuniform); otherwise use hq_filter_level / hq_filter_select.
 *   rnorm [N, L]   per-level row norms (NaN where 0), written by hq_filter_level_norms
 *   xstar [L]      smallest float x with ((x + 1) / 2 evaluated in float32) >= thr_l
 *   counts         optional [L][3][Q] (n_alive, n_pass, n_out) per level
 */
int hq_filter_fast_supported(const hq_index_layout* layout);
int hq_filter_level_norms(const float* idx, const uint16_t* lens, int64_t N, const hq_index_layout* layout,
                          float* rnorm, int32_t* nonuniform, void* stream);
int64_t hq_filter_fast_scratch_bytes(int64_t N, int Q, const hq_index_layout* layout);
/* Diagnostics: which variant hq_filter_fast runs for (N, Q) with a packed tensor-core operand and no `counts`
 * (2 = window mode: sample pass + predicted cut windows, 1 = full candidate lists, 0 = planes + generic cascade), and the
 * byte offset inside `scratch` of the int32 [Q] flags of the queries that were redone by the per-query fallback. */
int hq_filter_fast_mode(int64_t N, int Q, const hq_index_layout* layout);
int64_t hq_filter_fast_fallback_offset(int64_t N, int Q, const hq_index_layout* layout);
/* Window mode only: byte offsets inside `scratch` (a 128-byte aligned buffer) of out[0] the alive plane, out[1] the
 * windows float [4][Q], out[2] the counters int32 [4][Q], out[3] the sample's level-0 counts, out[4] prediction flags,
 * out[5] fallback tile flags, out[6] the segment counts int32 [Q][out[7]]; out[8] = segment capacity, out[9] = sample stride. */
int hq_filter_fast_window_layout(int64_t N, int Q, const hq_index_layout* layout, int64_t* out);
/* lvl_rows / lvl_pitch: optional HOST arrays [L] of device pointers / pitches of per-level
 * copies [N, pitch] of the index rows (pitch % 4 == 0); they are small enough to stay in the
 * 126 MB L2 while every query re-ranks its candidates. */
int hq_filter_fast(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout,
                   const float* q_idx, int Q, const float* xstar, const double* ratio,
                   const float* const* lvl_rows, const int32_t* lvl_pitch,
                   const float* db_packed, const uint32_t* valid, int64_t valid_pitch,
                   const uint16_t* lens, const int32_t* exc_rows, int n_exc, const double* thr,
                   uint32_t* mask, int64_t mask_stride, int32_t* n_out, int32_t* counts,
                   void* scratch, int64_t scratch_bytes, void* stream);
/* The same filter with one more database operand for the latency path: `db_rows` [N, hq_filter_rows_cols] float32 =
 * the used part of every index row with each level scaled by 1 / |c_l| and padded to whole float4
 * (hq_filter_rows_pack: 144 bytes per 1536-D row, 64 for 768-D).  Window-mode batches of at most
 * hq_filter_rows_max_queries() queries then run their window pass as plain fp32 FMAs over those rows on the CUDA cores
 * (a thread per row, rows staged by bulk copies; the tensor-core pass spends a 128-query tile on them): same outputs,
 * same 2e-6 parity band.  db_rows == NULL: hq_filter_fast. */
int hq_filter_rows_max_queries(void);
int hq_filter_rows_cols(const hq_index_layout* layout);
int hq_filter_rows_pack(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout, float* rows,
                        void* stream);
int hq_filter_fast_rows(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout,
                        const float* q_idx, int Q, const float* xstar, const double* ratio,
                        const float* const* lvl_rows, const int32_t* lvl_pitch,
                        const float* db_packed, const float* db_rows, const uint32_t* valid, int64_t valid_pitch,
                        const uint16_t* lens, const int32_t* exc_rows, int n_exc, const double* thr,
                        uint32_t* mask, int64_t mask_stride, int32_t* n_out, int32_t* counts,
                        void* scratch, int64_t scratch_bytes, void* stream);
/* exc_rows [n_exc] (device int32): rows whose stored length differs from lvl_keff at some level (a block mean
 * that is exactly 0 at the end of an index row: ~1 row in 10 M at 768-D).  The reference normalises the query
 * over the shorter common prefix for them (rag/search/engine.py:216-227); they are excluded from the dense
 * passes (clear their bits in `valid`) and scored pair by pair with the exact path's arithmetic, using
 * `lens` [N, L] and the HOST thresholds `thr` [L]; the cascade ranks them together with all other rows. */

/* Tensor-core form of the threshold pass of hq_filter_fast (db_packed != NULL): the per-level
 * dot products run as ONE tcgen05 tf32 contraction per 128-query x 64-row tile with every
 * value split exactly into tf32 hi + lo parts (three products per term, score error < 3e-7),
 * rows pre-scaled by 1 / |c_l|, and the comparison + bit packing done in the TMEM epilogue.
 * Supported when 2 * sum_l pad8(lvl_keff[l]) <= 128 and L <= 3.
 *   hq_filter_tc_packed_cols : packed floats per operand row ([x_hi | x_lo] per level, whole 128-byte slabs): 96 for
 *                        1536-D (64 x 64 grids), 64 for 768-D (32 x 32); 0 when the layout is not supported
 *   hq_filter_tc_pack  : idx [N, Lsum] (+ rnorm) -> packed [N, hq_filter_tc_packed_cols] float32 operand rows
 *                        (is_query = 1: unscaled; 0: scaled by 1 / |c_l|); the three products hi*hi + hi*lo + lo*hi are
 *                        three MMA sequences over the same blocks
 *   hq_filter_tc_valid : valid [L][valid_pitch] bit r = row r has a non-zero level norm;
 *                        valid_pitch >= hq_filter_tc_valid_pitch(N) (whole 64-row tiles)
 *   hq_filter_tc_plan  : work split of the pass for (N, Q) on the current device: n_ranges row ranges of
 *                        tiles_per_range 64-row tiles (the candidate lists have 2 segments per range)
 * With L >= 2 the pass also appends, per query, the rows that pass levels 0 and 1 together with their
 * level-1 / level-2 dot products to candidate lists inside `scratch`; the ratio cuts then run as
 * streaming selections over those lists (no index-row gathers), and the generic cascade only handles
 * the queries that path cannot (level-0 cut binds, list overflow, heavily tied cut bin). */
int hq_filter_tc_supported(const hq_index_layout* layout);
int hq_filter_tc_plan(int64_t N, int Q, int* n_ranges, int* tiles_per_range);
int64_t hq_filter_tc_valid_pitch(int64_t N);
int hq_filter_tc_packed_cols(const hq_index_layout* layout);
int hq_filter_tc_pack(const float* idx, const float* rnorm, int64_t N, const hq_index_layout* layout, int is_query,
                      float* packed, void* stream);
int hq_filter_tc_valid(const float* rnorm, int64_t N, const hq_index_layout* layout, uint32_t* valid,
                       int64_t valid_pitch, void* stream);

/* ---- a13/a15: cosine rerank + top-k --------------------------------------
 * rag/search/engine.py:622-660 (_calculate_embedding_cosine_similarity),
 * :512 / :778-781 (stable descending sort, first k).
 * score = (dot(q, c) / (|q| |c|) + 1) / 2, 0 if a norm is 0.  Rows dead in
 * `mask` are skipped.  Ties -> lower row id.  ids are int64 row ids + id_base
 * (-1 = no result), scores float32.
 *   hq_rerank_scores_f32 : exact fp32 FMA scores [Q, N] (-1 for dead rows)
 *   hq_topk_from_scores  : per-query top-k of a score matrix
 *   hq_rerank_topk_f32   : the two above through `scratch` (Q*N floats)            */
int hq_row_norms(const float* x, int64_t N, int64_t D, int64_t stride, float* norms, void* stream);
int hq_rerank_scores_f32(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride,
                         const float* q, const float* q_norm, int Q, int64_t q_stride,
                         const uint32_t* mask, int64_t mask_stride,
                         float* scores, int64_t scores_stride, void* stream);
/* out[i] = (cos(a_i, b_i) + 1) / 2 of n PAIRED rows (0 when a norm is 0): all windows of
 * _calculate_spatial_locality_similarity (rag/search/engine.py:662-714) in one launch.     */
int hq_paired_cosine01(const float* a, const float* a_norm, const float* b, const float* b_norm, int64_t n, int64_t D,
                       int64_t stride, float* out, void* stream);
/* same scores, but only the rows alive in `mask` are read (one warp per surviving row): the latency path for a
 * handful of queries, where the survivors (a few per cent of the rows) are far less data than the whole database */
int hq_rerank_scores_sparse_f32(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride,
                                const float* q, const float* q_norm, int Q, int64_t q_stride,
                                const uint32_t* mask, int64_t mask_stride,
                                float* scores, int64_t scores_stride, void* stream);
/* rag/search/engine.py:622-660 + :512 for a handful of queries in ONE launch: the exact fp32 cosine of the rows alive in `mask`
 * (the scores of hq_rerank_scores_sparse_f32, bit for bit) with the top-k (ties -> lower row id) taken on the way: no
 * [Q, N] score row, no separate top-k launches.  k <= 32, D % 4 == 0, 16-byte aligned rows
 * (hq_rerank_sparse_topk_supported); scratch: hq_rerank_sparse_topk_scratch_bytes(Q, k).  ids / scores [Q, k], -1 / -1.0f
 * when fewer than k rows survive.  db_unit_bf16 (optional, with dc_max as for hq_rerank_topk_unit_bf16): the survivors are
 * first ranked by their bf16 unit rows (half the bytes), the best 32 re-scored exactly, and a query whose top-k that does not
 * PROVE (error bound of the tensor-core rerank) is redone over the fp32 rows by a second launch: same results. */
int hq_rerank_sparse_topk_supported(int64_t D, int64_t db_stride, int64_t q_stride, int k);
int64_t hq_rerank_sparse_topk_scratch_bytes(int Q, int k);
int hq_rerank_sparse_topk(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride,
                          const void* db_unit_bf16, int64_t db_pitch, float dc_max,
                          const float* q, const float* q_norm, int Q, int64_t q_stride,
                          const uint32_t* mask, int64_t mask_stride, int k, int64_t id_base,
                          int64_t* ids, float* scores, void* scratch, int64_t scratch_bytes, void* stream);
int hq_topk_from_scores(const float* scores, int64_t scores_stride, int64_t N, int Q, int k, int64_t id_base,
                        int64_t* ids, float* out_scores, void* stream);
/* hq_topk_from_scores for FEW queries over LONG rows: the row is cut into chunks (one CTA each), the chunk lists are
 * merged; identical results (ties -> lower row id).  scratch: hq_topk_chunked_scratch_bytes(N, Q, k). */
int64_t hq_topk_chunked_scratch_bytes(int64_t N, int Q, int k);
int hq_topk_from_scores_chunked(const float* scores, int64_t scores_stride, int64_t N, int Q, int k, int64_t id_base,
                                int64_t* ids, float* out_scores, void* scratch, int64_t scratch_bytes, void* stream);
int64_t hq_rerank_scratch_bytes(int64_t N, int Q);
int hq_rerank_topk_f32(const float* db, const float* db_norm, int64_t N, int64_t D, int64_t db_stride,
                       const float* q, const float* q_norm, int Q, int64_t q_stride,
                       const uint32_t* mask, int64_t mask_stride, int k, int64_t id_base,
                       int64_t* ids, float* scores, void* scratch, int64_t scratch_bytes, void* stream);

/* Tensor-core rerank (the one true dense GEMM of the path): bf16 tcgen05 contraction with
 * the survivor mask and a streaming top-k' (k' = 16 for k <= 10, 32 for k <= 20) per epilogue thread fused
 * into the epilogue, then an exact re-score of a shortlist and a per-query PROOF that the shortlist contains
 * the exact top-k (see hq_rerank_topk_unit_bf16).  q_bf16: row-major bf16 copy of the queries
 * (hq_to_bf16), row pitch a multiple of 8. */
int hq_to_bf16(const float* src, int64_t N, int64_t D, int64_t src_stride, void* dst, int64_t dst_pitch, void* stream);
int64_t hq_rerank_bf16_scratch_bytes(int64_t N, int Q, int k);
/* Database operand of UNIT rows (c / |c| rounded to bf16): the epilogue needs no 1/|c| per column.  Zero-norm
 * rows score exactly 0.0 (rag/search/engine.py:640-643) and cannot be told apart in this operand: the caller passes
 * their ids (ascending) and they are appended after every other survivor, as the reference's stable sort orders them. */
int hq_to_bf16_unit(const float* src, int64_t N, int64_t D, int64_t src_stride, const float* norms, void* dst,
                    int64_t dst_pitch, void* stream);
/* Shard ingest: ONE pass over a block of embeddings [N, D] writes what the search path keeps per row -- the compact
 * hierarchical index row (rag/embedding_generation/hierarchical_index_generator.py:23-342 per document; identical bits to
 * hq_map_index with no grid output), the row norm (hq_row_norms) and the unit-row bf16 operand (hq_to_bf16_unit; may be
 * NULL).  Without a grid to write, block means over the Hilbert grid are run means of the input row, so the pass is
 * bound by HBM (read 4 D, write 2 D + 4 Lsum + 4 bytes per row).  `plan_codes[s]` = (pyramid level k << 24) | position in
 * curve order for index slot s (levels 3..6: runs of 64 .. 4096 values).  D / 128 must be in {2,4,6,8,12,16,24,32}
 * (hq_shard_ingest_supported), rows 16-byte aligned. */
int hq_shard_ingest_supported(int64_t D);
int hq_shard_ingest(const float* emb, int64_t N, int64_t D, int64_t stride, const int32_t* plan_codes, int Lsum,
                    float* idx, int64_t idx_pitch, float* norms, void* unit_bf16, int64_t unit_pitch, void* stream);
/* rag/search/engine.py:622-660 + :512 for a whole query batch.  The database operand holds UNIT rows (c / |c| rounded to
 * bf16), so the accumulator a of a (query, row) pair orders the rows like the cosine.  Per query the rows are taken in
 * accumulator order and re-scored exactly (fp32 FMA over db_f32, the arithmetic of hq_rerank_scores_sparse_f32) in chunks of
 * 16 until the GUARD holds:   (k-th best exact q.c/|c| so far)  >  (largest accumulator of any row not re-scored) + E,
 *   E = |bf16(q) - q| (1 + 2^-8) + |q| (dc_max + D 2^-22 + 1e-5)   >=   |a - q.c/|c||   for every row,
 * which proves that the exact top-k lies among the re-scored rows.  dc_max = max over the rows of |unit_bf16 - c / |c||_2
 * (hq_bf16_unit_error_max).  Queries still unproven after 64 rows (near-duplicate clusters around the k-th score) are listed in
 * `guard` and re-scored exactly over ALL their surviving rows by a second pair of kernels in the same call, so the result
 * equals the exact path's for every query.
 *   db_f32 == NULL : bf16-only database.  The stored rows ARE the bf16 unit rows: scores are (cos(q, stored row) + 1) / 2,
 *                    exact for the stored values (they differ from the fp32 rows' scores by the bf16 rounding of the rows,
 *                    <= 2^-8 relative per component); db_norm = norms of the stored rows, dc_max = 0.
 *   guard [4 + Q] int32 (device): [0] queries sent to the exact fallback, [1] rows re-scored exactly, [4..] the flagged queries. */
int hq_rerank_topk_unit_bf16(const void* db_unit_bf16, int64_t db_pitch, const float* db_f32, int64_t db_stride,
                             const float* db_norm, const int32_t* zero_rows, int n_zero, int64_t N, int64_t D,
                             const void* q_bf16, int64_t q_pitch, const float* q_f32, int64_t q_stride,
                             const float* q_norm, int Q, const uint32_t* mask, int64_t mask_stride,
                             int k, int64_t id_base, float dc_max, int64_t* ids, float* scores, int32_t* guard,
                             void* scratch, int64_t scratch_bytes, void* stream);
/* out_max (device float, initialised by the caller, e.g. 0) = max(out_max, max_r |unit_bf16[r] - src[r] / norms[r]|_2) */
int hq_bf16_unit_error_max(const float* src, int64_t N, int64_t D, int64_t src_stride, const float* norms,
                           const void* unit_bf16, int64_t pitch, float* out_max, void* stream);

/* ---- a14: comprehensive similarity blend -----------------------------------
 * rag/search/engine.py:516-575 (_calculate_comprehensive_similarity):
 *   0.5 * weighted per-level (cos+1)/2 of the index rows (:994-1099; weights :1101-1138, passed in)
 * + 0.3 * (cos+1)/2 of the grids (:622-660)
 * + 0.2 * mean (cos+1)/2 over ws x ws grid windows at stride ws/2, ws = min(4, n/4) (:662-714)
 * frames / q_frames: enhanced frames, rows [0, n) = n x n grid, rows [n, n+L) = index rows.
 * cand_ids: optional device [Q, M] row ids per query (-1 = none -> score -1); NULL = every row (M == N).
 * weights_host: HOST array [L].  out [Q, M] float32. */
int hq_comprehensive_scores(const float* frames, int64_t N, int n, int L, int64_t frame_stride,
                            const float* q_frames, int Q, int64_t q_stride, const float* weights_host,
                            const int64_t* cand_ids, int64_t M, float* out, void* stream);

/* ---- f1: PrecomputedHilbertIndexer half-offset squares ------------------------
 * core/precomputed_hilbert_index.py:185-203.  The aligned square means of every level come out of
 * hq_map_index_fused (they are run means of the Hilbert stream); a square offset by half its side is
 * the mean of four entries of the next finer level `half` (row-major [G, G] per item):
 * out [N, (G/2 - 1)^2], row-major over (row, col) like the reference's loop. */
int hq_offset_square_means(const float* half, int64_t N, int G, int64_t half_stride,
                           float* out, int64_t out_stride, void* stream);

/* ---- f4: video-path hierarchical similarity ------------------------------------
 * core/video_storage.py:763-781 (_calculate_hierarchical_similarity), used by _traditional_search (:741-761),
 * _sort_frames_by_hierarchical_indices (:1203-1277) and _find_optimal_insertion_position (:1751-1803):
 * out[m, n] = clamp((pearson(a_m[:S], b_n[:S]) + 1) / 2, 0, 1) in float64; zero-variance vectors score 1.0 when
 * np.allclose, else 0.0. */
int hq_pearson01_matrix(const double* a, int64_t M, int64_t a_stride, const double* b, int64_t N, int64_t b_stride,
                        int S, double* out, int64_t out_stride, void* stream);

/* ---- a15 multi-GPU: merge of per-shard top-k ---------------------------
 * in_ids/in_scores [P, Q, k] (all-gathered) -> out [Q, k]; ties -> lower id;
 * entries with id < 0 are empty. */
int hq_topk_merge(const int64_t* in_ids, const float* in_scores, int P, int Q, int k,
                  int64_t* out_ids, float* out_scores, void* stream);
/* Same merge for results gathered as ONE packed block per shard ([Q, k] int64 ids followed by [Q, k] float32
 * scores, a single all-gather): shard p's lists start at in_ids + p * ids_shard_stride (int64 elements) and
 * in_scores + p * scores_shard_stride (float elements). */
int hq_topk_merge_strided(const int64_t* in_ids, const float* in_scores, int P, int Q, int k, int64_t ids_shard_stride,
                          int64_t scores_shard_stride, int64_t* out_ids, float* out_scores, void* stream);

/* ---- a11: core progressive search, per-level similarity ------------------
 * core/search_engine.py:111-189 (compare_indices_at_level): sims [N, n_levels]
 * float64; level l compares q[q_start[l] : +lvl_len[l]] with
 * cand[row, c_start[l] : +lvl_len[l]] (lvl_len = common length of the level). */
int hq_core_level_sims(const double* cand, int64_t N, int S, int64_t cand_stride, const double* q,
                       const int32_t* q_start, const int32_t* c_start, const int32_t* lvl_len, int n_levels,
                       double* sims, void* stream);

/* filter_scope = "global" on a row-sharded database: the per-shard halves of the reference's single-list ratio cut
 * (rag/search/engine.py:272-287).  scores [Q, scores_stride] are the level scores of hq_filter_level, mask [Q, mask_stride] its
 * pass bits, need [Q] marks the queries whose global pass count exceeds the cap.  hq_gcut_hist adds the passed rows to
 * hist [Q, 65536] by the high (d_hi == NULL) or the low 16 bits of the score's bit pattern (rows whose high digit is
 * d_hi[q]); the caller all-reduces hist; hq_gcut_scan finds the digit of the want[q]-th best and the count above it;
 * hq_gcut_ties counts this shard's rows AT the cut score k_star[q]; hq_gcut_apply rewrites the mask: rows above the cut
 * score and the first quota[q] tied rows by ascending row id. */
int hq_gcut_hist(const float* scores, int64_t scores_stride, int64_t N, int Q, const uint32_t* mask, int64_t mask_stride,
                 const int32_t* need, const int32_t* d_hi, int32_t* hist, void* stream);
int hq_gcut_scan(const int32_t* hist, int Q, const int32_t* need, const int64_t* want, int64_t* d_out, int64_t* above_out,
                 void* stream);
int hq_gcut_ties(const float* scores, int64_t scores_stride, int64_t N, int Q, const uint32_t* mask, int64_t mask_stride,
                 const int32_t* need, const int64_t* k_star, int64_t* ties, void* stream);
int hq_gcut_apply(const float* scores, int64_t scores_stride, int64_t N, int Q, uint32_t* mask, int64_t mask_stride,
                  const int32_t* need, const int64_t* k_star, const int64_t* quota, void* stream);

/* Per-kernel timing for the benchmark's roofline (CUDA events on the launching stream around the launches of the four
 * search kernels): hq_kernel_timing(1) arms it, hq_kernel_timing_read synchronises and returns, per slot, the summed
 * milliseconds and the number of launches since it was armed (slot 0 k_rerank_tc, 1 k_filter_bits_tc main pass,
 * 2 k_filter_cascade_win / _lists, 3 k_rerank_tc_merge), then disarms. */
int hq_kernel_timing(int on);
int hq_kernel_timing_read(float* ms_out, int32_t* n_out, int slots);

#ifdef __cplusplus
}
#endif
#endif /* HQ_B200_H */
