// Shared device helpers and host-side error plumbing for libhq_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/hq_b200.h"

void hq_set_error(const char* fmt, ...);
void hq_note_launch(int n);     // bookkeeping for hq_launch_count (bench.py reports it)

#define HQ_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) {                                \
            hq_set_error(__VA_ARGS__);                \
            return HQ_EINVAL;                         \
        }                                             \
    } while (0)

#define HQ_CUDA_OK(expr)                                                            \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            hq_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));           \
            return HQ_ECUDA;                                                        \
        }                                                                           \
    } while (0)

#define HQ_LAUNCH_OK(name)                                                          \
    do {                                                                            \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            hq_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));  \
            return HQ_ECUDA;                                                        \
        }                                                                           \
        hq_note_launch(1);                                                          \
    } while (0)

static inline bool hq_is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }
static inline int hq_log2(int64_t n) { int k = 0; while ((int64_t(1) << k) < n) ++k; return k; }
int hq_cached_sm_count();
// per-kernel timing hooks for bench.py (hq_lib.cu): no-ops unless hq_kernel_timing(1) armed them
int hq_time_begin(int slot, cudaStream_t st);
void hq_time_end(int slot, int token, cudaStream_t st);
// tensor-core bit-plane pass of the coarse filter (hq_filter_tc.cu), used by hq_filter_fast
struct HqFilterLists {          // candidate lists written by the pass (see FtcParams in hq_filter_tc.cu)
    uint32_t* rows;
    float* k1;
    float* k2;
    int32_t* seg_n;
    int64_t seg_cap;
    int n_segs;
};
int hq_filter_bits_tc_launch(const float* db_packed, const uint32_t* valid, int64_t valid_pitch, int64_t N,
                             const hq_index_layout* layout, const float* q_idx, int Q, const float* xstar, float* q_packed,
                             float* tq, float* nq, uint32_t* bits, int64_t bits_pitch, const HqFilterLists* lists, cudaStream_t st);
extern "C" int hq_filter_tc_plan(int64_t N, int Q, int* n_ranges, int* tiles_per_range);
int hq_filter_tc_plan_strided(int64_t N, int Q, int tile_stride, int* n_ranges, int* tiles_per_range);
struct HqFtcOpts {              // variants of the threshold pass (see FtcParams in hq_filter_tc.cu)
    int tile_stride;            // > 1: sample pass over every tile_stride-th 64-row tile (lists + level-0 counts, no planes)
    const int32_t* unit_only;   // optional [ceil(Q / 128)]: only query tiles with a non-zero flag are processed
    int32_t* c0_cnt;            // optional [Q]: += rows passing the level-0 threshold
    const float* win;           // window mode: [4][Q] lo1, hi1, lo2, hi2
    int32_t* wcnt;              // window mode: [4][Q] counters
};
int hq_filter_tc_prepare(const hq_index_layout* layout, const float* q_idx, int Q, const float* xstar, float* q_packed, float* tq,
                         float* nq, cudaStream_t st);
int hq_filter_rows_pass(const float* rows, const uint32_t* valid, int64_t valid_pitch, int64_t N, const hq_index_layout* layout,
                        const float* q_idx, int Q, const float* tq, const float* win, uint32_t* alive, int64_t alive_pitch,
                        int32_t* wcnt, const HqFilterLists* lists, cudaStream_t st);
extern "C" int hq_filter_rows_max_queries(void);
int hq_filter_tc_fold(const hq_index_layout* layout, const float* tq, const float* win, int Q, float* q_packed, cudaStream_t st);
int hq_filter_tc_pass(const float* db_packed, const uint32_t* valid, int64_t valid_pitch, int64_t N, const hq_index_layout* layout, int Q,
                      const float* q_packed, const float* tq, uint32_t* bits, int64_t bits_pitch, const HqFilterLists* lists,
                      const HqFtcOpts* o, cudaStream_t st);

// ---- Hilbert curve, the reference's variant (core/hilbert_mapper.py:42-113) ----
// d -> (x, y), low bit-pairs first.
__host__ __device__ __forceinline__ void hq_d2xy(int log2n, uint64_t d, uint32_t& x, uint32_t& y) {
    uint32_t xx = 0, yy = 0;
    uint64_t t = d;
    for (int i = 0; i < log2n; ++i) {
        const uint32_t s = 1u << i;
        const uint32_t rx = 1u & (uint32_t)(t >> 1);
        const uint32_t ry = 1u & ((uint32_t)t ^ rx);
        if (ry == 0) {
            if (rx == 1) { xx = s - 1 - xx; yy = s - 1 - yy; }
            const uint32_t tmp = xx; xx = yy; yy = tmp;
        }
        xx += s * rx;
        yy += s * ry;
        t >>= 2;
    }
    x = xx; y = yy;
}

// (x, y) -> d, high bits first (core/hilbert_mapper.py:68-90).
__host__ __device__ __forceinline__ uint64_t hq_xy2d(int log2n, uint32_t x, uint32_t y) {
    uint64_t d = 0;
    for (int i = log2n - 1; i >= 0; --i) {
        const uint32_t s = 1u << i;
        const uint32_t rx = (x & s) ? 1u : 0u;
        const uint32_t ry = (y & s) ? 1u : 0u;
        d += (uint64_t)s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = s - 1 - x; y = s - 1 - y; }
            const uint32_t tmp = x; x = y; y = tmp;
        }
    }
    return d;
}
