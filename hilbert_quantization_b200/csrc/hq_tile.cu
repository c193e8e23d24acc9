// K2/K3/K4: fused Hilbert map / unmap + run-mean pyramid + index gather (fp32 items).
//
// Work unit ("chunk") = 4096 curve positions = 1024 quads of 4 consecutive positions:
//   n <= 64 : 4096/(n*n) whole items per chunk (one 64x64 item, four 32x32 items, ...)
//   n  > 64 : one 64x64 tile of one item; any aligned 4096-run of the curve is an aligned
//             64x64 tile whose inner order is the n=64 base order composed with a
//             (swap, flip) pair accumulated from the upper bit-pairs of d (the rotate
//             step of core/hilbert_mapper.py:92-113 applied to whole tiles).
// A thread owns quads tid, tid+256, tid+512, tid+768 of the chunk, so global accesses on
// the curve side are 128-bit and fully coalesced; the 2-D side goes through a shared
// memory image of the tile (row pitch n+4 floats) and is written / read as 128-bit row
// segments.  The run-mean pyramid (levels 1..3) is reduced in registers with warp
// shuffles straight from the loaded quads; levels >= 4 are finished from shared memory.
#include "hq_tc.cuh"
#include <float.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kThreads = 256;
constexpr int kChunkQuads = 1024;
constexpr int kQuadsPerThread = kChunkQuads / kThreads;

struct TileParams {
    const float* src;
    int direction;          // 0 stream->grid, 1 grid->stream
    int64_t N, D, src_stride;
    int64_t D_last;         // values of the LAST item (parameter streams: the tail grid is partially filled); == D otherwise
    int log2n;              // whole-grid side
    int log2t;              // tile side (min(n, 64))
    float* grid_out; int64_t grid_stride;
    float* stream_out; int64_t stream_stride;
    const int32_t* plan; int plan_len;
    void* idx_out; int64_t idx_stride;
    void* scratch;          // big grids: per item levels [min_level .. top] concatenated
    int min_level;          // lowest pyramid level the plan references (>= 1), 99 = none
    int64_t num_chunks;
    int vec_src, vec_grid, vec_stream;   // 128-bit access legal?
    // fused uint8 quantisation of the enhanced frame (k_item_pass_bulk<.., QUANT>): frame = n^2 grid bytes + plan_len index bytes
    uint8_t* u8_out; int64_t u8_stride;  // bytes per item frame (multiple of 16)
    float* mm_out;                       // [N, 2] (min, max) of every frame
    int frame_zero;                      // the frame holds structural zeros (padding cells / padded index slots)
};

__device__ __forceinline__ uint32_t quad_entry(int log2t, uint32_t q) {
    // cells of the four positions 4q..4q+3 inside a tile of side 2^log2t (base orientation)
    uint32_t x[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) hq_d2xy(log2t, 4ull * q + i, x[i], y[i]);
    const uint32_t x0 = x[0] & ~1u, y0 = y[0] & ~1u;
    uint32_t e = x0 | (y0 << 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) e |= (((x[i] & 1u) | ((y[i] & 1u) << 1)) << (16 + 2 * i));
    return e;
}

// (swap, flip) accumulated over the bit-pairs of the tile index t (upper part of d), and
// the tile origin.  Mirrors the loop of _hilbert_index_to_xy for s >= 64.
__device__ __forceinline__ void tile_frame(int upper_bits, uint64_t t, uint32_t& X, uint32_t& Y, uint32_t& swp, uint32_t& flp) {
    uint32_t xx = 0, yy = 0, a = 0, b = 0;
    for (int i = 0; i < upper_bits; ++i) {
        const uint32_t s = 1u << i;
        const uint32_t rx = 1u & (uint32_t)(t >> 1);
        const uint32_t ry = 1u & ((uint32_t)t ^ rx);
        if (ry == 0) {
            if (rx == 1) { xx = s - 1 - xx; yy = s - 1 - yy; b ^= 1u; }
            const uint32_t tmp = xx; xx = yy; yy = tmp;
            a ^= 1u;
        }
        xx += s * rx;
        yy += s * ry;
        t >>= 2;
    }
    X = xx; Y = yy; swp = a; flp = b;
}

template <int MODE> struct PyrT { using type = float; };
template <> struct PyrT<1> { using type = double; };

template <int MODE>
__device__ __forceinline__ typename PyrT<MODE>::type mean4(float x, float y, float z, float w) {
    if (MODE == 0) return (typename PyrT<MODE>::type)(((x + y) + (z + w)) * 0.25f);
    return (typename PyrT<MODE>::type)(((((double)x + (double)y) + (double)z) + (double)w) * 0.25);
}

// mean of the 4 lanes {g, g+stride, g+2*stride, g+3*stride}; every lane of the group gets it (MODE 0)
// or only the group leader holds a valid value (MODE 1, strict left-to-right order).
template <int MODE>
__device__ __forceinline__ typename PyrT<MODE>::type group_mean(typename PyrT<MODE>::type v, int stride) {
    if (MODE == 0) {
        float t = (float)v;
        t = t + __shfl_xor_sync(0xffffffffu, t, stride);
        t = t + __shfl_xor_sync(0xffffffffu, t, 2 * stride);
        return (typename PyrT<MODE>::type)(t * 0.25f);
    } else {
        const double a = (double)v;
        const double b = __shfl_down_sync(0xffffffffu, a, stride);
        const double c = __shfl_down_sync(0xffffffffu, a, 2 * stride);
        const double d = __shfl_down_sync(0xffffffffu, a, 3 * stride);
        return (typename PyrT<MODE>::type)((((a + b) + c) + d) * 0.25);
    }
}

template <int DIR, int MODE>
__global__ void __launch_bounds__(kThreads) k_tile_pass(const TileParams p) {
    using P = typename PyrT<MODE>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int log2n = p.log2n, log2t = p.log2t;
    const uint32_t T = 1u << log2t;
    const uint32_t pitch = T + (T >= 32 ? 4u : 0u);            // floats per smem row
    const uint32_t item_cells = 1u << (2 * log2t);              // cells per tile / small item
    const uint32_t qpi = item_cells >> 2;                       // quads per tile / small item
    const int log2qpi = 2 * log2t - 2;
    const uint32_t ipc = kChunkQuads / qpi;                     // items per chunk (1 for T == 64)
    const bool tiled = log2n > log2t;
    const int upper_bits = log2n - log2t;
    const uint32_t pyr_items = (item_cells - 1) / 3;            // pyramid values per tile (levels 1..log2t)
    const int top_level = log2t;                                // level with one value per tile

    float* s_grid = reinterpret_cast<float*>(smem_raw);                                   // ipc * T * pitch
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_grid + (size_t)ipc * T * pitch);      // qpi
    P* s_pyr = reinterpret_cast<P*>(s_tab + qpi + ((qpi & 1) ? 1 : 0));                   // ipc * pyr_items (8B aligned)
    __shared__ uint32_t s_frame[4];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool want_pyr = p.plan_len > 0 && p.min_level <= 32;
    const bool need_grid_smem = (DIR == 1) || p.grid_out != nullptr || (p.plan_len > 0 && !tiled);

    // one-time per CTA: quad table, zeroed tile image and pyramid (cells/runs at d >= D are never written)
    for (uint32_t q = tid; q < qpi; q += kThreads) s_tab[q] = quad_entry(log2t, q);
    for (uint32_t i = tid; i < ipc * T * pitch; i += kThreads) s_grid[i] = 0.f;
    for (uint32_t i = tid; i < ipc * pyr_items; i += kThreads) s_pyr[i] = (P)0;
    __syncthreads();

    const int64_t tiles_per_item = tiled ? ((int64_t)1 << (2 * upper_bits)) : 1;
    const int64_t n_cells = (int64_t)1 << (2 * log2n);

    for (int64_t chunk = blockIdx.x; chunk < p.num_chunks; chunk += gridDim.x) {
        // ---- chunk frame ----
        int64_t item0; uint64_t tile = 0; uint32_t X = 0, Y = 0, swp = 0, flp = 0;
        if (tiled) {
            item0 = chunk / tiles_per_item;
            tile = (uint64_t)(chunk - item0 * tiles_per_item);
            if (tid == 0) {
                tile_frame(upper_bits, tile, X, Y, swp, flp);
                s_frame[0] = X; s_frame[1] = Y; s_frame[2] = swp; s_frame[3] = flp;
            }
            __syncthreads();
            X = s_frame[0]; Y = s_frame[1]; swp = s_frame[2]; flp = s_frame[3];
        } else {
            item0 = chunk * ipc;
        }
        const int64_t d_tile0 = (int64_t)tile << (2 * log2t);   // first curve position of this tile

        float4 v[kQuadsPerThread];

        if (DIR == 0) {
            // ---- load quads from the stream (coalesced 128-bit) ----
#pragma unroll
            for (int r = 0; r < kQuadsPerThread; ++r) {
                const uint32_t qi = tid + r * kThreads;
                const uint32_t il = qi >> log2qpi;
                const uint32_t q = qi & (qpi - 1);
                const int64_t item = item0 + il;
                const int64_t d = d_tile0 + 4 * (int64_t)q;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (item < p.N && d < p.D) {
                    const float* s = p.src + item * p.src_stride + d;
                    if (p.vec_src && d + 3 < p.D) {
                        val = __ldg(reinterpret_cast<const float4*>(s));
                    } else {
                        val.x = __ldg(s);
                        if (d + 1 < p.D) val.y = __ldg(s + 1);
                        if (d + 2 < p.D) val.z = __ldg(s + 2);
                        if (d + 3 < p.D) val.w = __ldg(s + 3);
                    }
                }
                v[r] = val;
            }
            // ---- scatter into the tile image ----
            if (need_grid_smem) {
#pragma unroll
                for (int r = 0; r < kQuadsPerThread; ++r) {
                    const uint32_t qi = tid + r * kThreads;
                    const uint32_t il = qi >> log2qpi;
                    const uint32_t q = qi & (qpi - 1);
                    const int64_t d = d_tile0 + 4 * (int64_t)q;
                    if (!tiled && d >= p.D) continue;            // never written -> stays zero
                    const uint32_t e = s_tab[q];
                    const uint32_t x0 = e & 0xffu, y0 = (e >> 8) & 0xffu;
                    const float vals[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
                    float* base = s_grid + (size_t)il * T * pitch;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t x = x0 + ((e >> (16 + 2 * i)) & 1u);
                        uint32_t y = y0 + ((e >> (17 + 2 * i)) & 1u);
                        if (flp) { x = T - 1 - x; y = T - 1 - y; }
                        if (swp) { const uint32_t t2 = x; x = y; y = t2; }
                        base[y * pitch + x] = vals[i];
                    }
                }
            }
        } else {
            // ---- load the 2-D tile (coalesced 128-bit row segments) into the tile image ----
            const uint32_t vec_per_row = T >> 2;
            const uint32_t total_vec = ipc * T * vec_per_row;     // == 1024
#pragma unroll
            for (int r = 0; r < kQuadsPerThread; ++r) {
                const uint32_t vi = tid + r * kThreads;
                if (vi >= total_vec) break;
                const uint32_t il = vi >> log2qpi;
                const uint32_t rem = vi & (qpi - 1);
                const uint32_t y = rem >> (log2t - 2);
                const uint32_t x = (rem & (vec_per_row - 1)) << 2;
                const int64_t item = item0 + il;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (item < p.N) {
                    const float* s = p.src + item * p.src_stride + (((int64_t)Y << log2t) + y << log2n) + ((int64_t)X << log2t) + x;
                    if (p.vec_src) val = __ldg(reinterpret_cast<const float4*>(s));
                    else { val.x = __ldg(s); val.y = __ldg(s + 1); val.z = __ldg(s + 2); val.w = __ldg(s + 3); }
                }
                *reinterpret_cast<float4*>(s_grid + (size_t)il * T * pitch + y * pitch + x) = val;
            }
            __syncthreads();
            // ---- gather quads in curve order ----
#pragma unroll
            for (int r = 0; r < kQuadsPerThread; ++r) {
                const uint32_t qi = tid + r * kThreads;
                const uint32_t il = qi >> log2qpi;
                const uint32_t q = qi & (qpi - 1);
                const uint32_t e = s_tab[q];
                const uint32_t x0 = e & 0xffu, y0 = (e >> 8) & 0xffu;
                const float* base = s_grid + (size_t)il * T * pitch;
                float vals[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t x = x0 + ((e >> (16 + 2 * i)) & 1u);
                    uint32_t y = y0 + ((e >> (17 + 2 * i)) & 1u);
                    if (flp) { x = T - 1 - x; y = T - 1 - y; }
                    if (swp) { const uint32_t t2 = x; x = y; y = t2; }
                    vals[i] = base[y * pitch + x];
                }
                v[r] = make_float4(vals[0], vals[1], vals[2], vals[3]);
                if (p.stream_out) {
                    const int64_t item = item0 + il;
                    const int64_t d = d_tile0 + 4 * (int64_t)q;
                    if (item < p.N && d < p.D) {
                        float* o = p.stream_out + item * p.stream_stride + d;
                        if (p.vec_stream && d + 3 < p.D) {
                            *reinterpret_cast<float4*>(o) = v[r];
                        } else {
                            o[0] = vals[0];
                            if (d + 1 < p.D) o[1] = vals[1];
                            if (d + 2 < p.D) o[2] = vals[2];
                            if (d + 3 < p.D) o[3] = vals[3];
                        }
                    }
                }
            }
        }

        // ---- pyramid levels 1..3 in registers / shuffles ----
        if (want_pyr) {
#pragma unroll
            for (int r = 0; r < kQuadsPerThread; ++r) {
                const uint32_t qi = tid + r * kThreads;
                const uint32_t il = qi >> log2qpi;
                const uint32_t q = qi & (qpi - 1);
                const bool live = tiled || DIR == 1 || (d_tile0 + 4 * (int64_t)(q & ~31u)) < p.D;   // warp-uniform
                if (!live) continue;
                P* pyr = s_pyr + (size_t)il * pyr_items;
                const P m1 = mean4<MODE>(v[r].x, v[r].y, v[r].z, v[r].w);
                uint32_t base = 0;
                if (p.min_level <= 1) pyr[base + q] = m1;
                if (top_level >= 2) {
                    base += qpi;
                    const P m2 = group_mean<MODE>(m1, 1);
                    if ((lane & 3) == 0 && p.min_level <= 2) pyr[base + (q >> 2)] = m2;
                    if (top_level >= 3) {
                        base += qpi >> 2;
                        const P m3 = group_mean<MODE>(m2, 4);
                        if ((lane & 15) == 0) pyr[base + (q >> 4)] = m3;
                    } else if ((lane & 3) == 0) {
                        pyr[base + (q >> 2)] = m2;
                    }
                } else {
                    pyr[base + q] = m1;
                }
            }
        }
        __syncthreads();

        // ---- levels >= 4 from shared memory ----
        if (want_pyr && top_level >= 4) {
            uint32_t base_prev = qpi + (qpi >> 2);               // start of level 3
            uint32_t cnt_prev = qpi >> 4;                        // values per item at level 3
            for (int k = 4; k <= top_level; ++k) {
                const uint32_t cnt = cnt_prev >> 2;
                const uint32_t base = base_prev + cnt_prev;
                for (uint32_t t = tid; t < ipc * cnt; t += kThreads) {
                    const uint32_t il = t / cnt, j = t - il * cnt;
                    P* pyr = s_pyr + (size_t)il * pyr_items;
                    const P a = pyr[base_prev + 4 * j], b = pyr[base_prev + 4 * j + 1];
                    const P c = pyr[base_prev + 4 * j + 2], d = pyr[base_prev + 4 * j + 3];
                    pyr[base + j] = MODE == 0 ? (P)(((a + b) + (c + d)) * (P)0.25) : (P)((((a + b) + c) + d) * (P)0.25);
                }
                __syncthreads();
                base_prev = base;
                cnt_prev = cnt;
            }
        }

        // ---- 2-D side out (map direction) ----
        if (DIR == 0 && p.grid_out) {
            const uint32_t vec_per_row = T >> 2;
            const uint32_t total_vec = ipc * T * vec_per_row;
#pragma unroll
            for (int r = 0; r < kQuadsPerThread; ++r) {
                const uint32_t vi = tid + r * kThreads;
                if (vi >= total_vec) break;
                const uint32_t il = vi >> log2qpi;
                const uint32_t rem = vi & (qpi - 1);
                const uint32_t y = rem >> (log2t - 2);
                const uint32_t x = (rem & (vec_per_row - 1)) << 2;
                const int64_t item = item0 + il;
                if (item >= p.N) continue;
                const float4 val = *reinterpret_cast<const float4*>(s_grid + (size_t)il * T * pitch + y * pitch + x);
                float* o = p.grid_out + item * p.grid_stride + (((int64_t)Y << log2t) + y << log2n) + ((int64_t)X << log2t) + x;
                if (p.vec_grid) __stcs(reinterpret_cast<float4*>(o), val);
                else { o[0] = val.x; o[1] = val.y; o[2] = val.z; o[3] = val.w; }
            }
        }

        // ---- index values ----
        if (want_pyr || (p.plan_len > 0 && !tiled)) {
            if (!tiled) {
                const uint32_t total = ipc * (uint32_t)p.plan_len;
                for (uint32_t t = tid; t < total; t += kThreads) {
                    const uint32_t il = t / (uint32_t)p.plan_len, i = t - il * (uint32_t)p.plan_len;
                    const int64_t item = item0 + il;
                    if (item >= p.N) continue;
                    const int32_t off = __ldg(p.plan + i);
                    P val = (P)0;
                    if (off >= 0) {
                        if ((uint32_t)off < item_cells) {
                            const uint32_t y = (uint32_t)off >> log2t, x = (uint32_t)off & (T - 1);
                            val = (P)s_grid[(size_t)il * T * pitch + y * pitch + x];
                        } else {
                            val = s_pyr[(size_t)il * pyr_items + ((uint32_t)off - item_cells)];
                        }
                    }
                    reinterpret_cast<P*>(p.idx_out)[item * p.idx_stride + i] = val;
                }
            } else {
                // big grid: publish in-tile levels [min_level .. 6] to the per-item scratch pyramid
                P* sc = reinterpret_cast<P*>(p.scratch);
                int64_t sc_item = 0;                               // values per item in scratch
                for (int k = p.min_level; k <= log2n; ++k) sc_item += n_cells >> (2 * k);
                P* dst = sc + item0 * sc_item;
                int64_t lvl_base_g = 0;                            // base of level k inside scratch
                uint32_t lvl_base_s = 0;                           // base of level k inside s_pyr
                for (int k = 1; k <= top_level; ++k) {
                    const uint32_t cnt = item_cells >> (2 * k);    // values of this tile at level k
                    if (k >= p.min_level) {
                        for (uint32_t t = tid; t < cnt; t += kThreads)
                            dst[lvl_base_g + (int64_t)tile * cnt + t] = s_pyr[lvl_base_s + t];
                        lvl_base_g += n_cells >> (2 * k);
                    }
                    lvl_base_s += cnt;
                }
            }
        }
        __syncthreads();     // tile image / pyramid are reused by the next chunk
    }
}

// Big grids, second launch: finish levels above the tile level and apply the gather plan.
template <int MODE>
__global__ void __launch_bounds__(256) k_pyramid_top(const TileParams p) {
    using P = typename PyrT<MODE>::type;
    const int log2n = p.log2n;
    const int64_t n_cells = (int64_t)1 << (2 * log2n);
    int64_t sc_item = 0;
    for (int k = p.min_level; k <= log2n; ++k) sc_item += n_cells >> (2 * k);
    const int64_t item = blockIdx.x;
    P* sc = reinterpret_cast<P*>(p.scratch) + item * sc_item;

    // levels log2t+1 .. log2n, each from the level below (one CTA per item, block-wide steps)
    int64_t base_prev = 0;
    for (int k = p.min_level; k < p.log2t; ++k) base_prev += n_cells >> (2 * k);
    for (int k = p.log2t + 1; k <= log2n; ++k) {
        const int64_t cnt_prev = n_cells >> (2 * (k - 1));
        const int64_t cnt = cnt_prev >> 2;
        const int64_t base = base_prev + cnt_prev;
        for (int64_t j = threadIdx.x; j < cnt; j += blockDim.x) {
            const P a = sc[base_prev + 4 * j], b = sc[base_prev + 4 * j + 1];
            const P c = sc[base_prev + 4 * j + 2], d = sc[base_prev + 4 * j + 3];
            sc[base + j] = MODE == 0 ? (P)(((a + b) + (c + d)) * (P)0.25) : (P)((((a + b) + c) + d) * (P)0.25);
        }
        __syncthreads();
        base_prev = base;
    }
    // plan
    int64_t skip = 0;                                    // pyramid values below min_level (not stored)
    for (int k = 1; k < p.min_level; ++k) skip += n_cells >> (2 * k);
    for (int i = threadIdx.x; i < p.plan_len; i += blockDim.x) {
        const int32_t off = p.plan[i];
        P val = (P)0;
        if (off >= 0) {
            if ((int64_t)off < n_cells) {
                const uint32_t y = (uint32_t)off >> log2n, x = (uint32_t)off & ((1u << log2n) - 1);
                if (p.direction == 1) {
                    val = (P)p.src[item * p.src_stride + off];
                } else {
                    const int64_t d = (int64_t)hq_xy2d(log2n, x, y);
                    val = d < (item == p.N - 1 ? p.D_last : p.D) ? (P)p.src[item * p.src_stride + d] : (P)0;
                }
            } else {
                val = sc[(int64_t)off - n_cells - skip];
            }
        }
        reinterpret_cast<P*>(p.idx_out)[item * p.idx_stride + i] = val;
    }
}

#include "hq_item_pass.cuh"

size_t tile_smem_bytes(int log2t, int mode) {
    const uint32_t T = 1u << log2t;
    const uint32_t pitch = T + (T >= 32 ? 4u : 0u);
    const uint32_t cells = T * T, qpi = cells / 4, ipc = kChunkQuads / qpi;
    const uint32_t pyr_items = (cells - 1) / 3;
    size_t b = (size_t)ipc * T * pitch * 4 + (size_t)(qpi + (qpi & 1)) * 4;
    b += (size_t)ipc * pyr_items * (mode ? 8 : 4) + 16;
    return b;
}

template <int DIR, int MODE>
int launch_tile(const TileParams& p, cudaStream_t st) {
    const size_t smem = tile_smem_bytes(p.log2t, MODE);
    static bool attr_done[2][2] = {{false, false}, {false, false}};
    if (!attr_done[DIR][MODE]) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_tile_pass<DIR, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_done[DIR][MODE] = true;
    }
    int per_sm = 0;
    HQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_pass<DIR, MODE>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)hq_cached_sm_count() * per_sm;
    if (blocks > p.num_chunks) blocks = p.num_chunks;
    k_tile_pass<DIR, MODE><<<(unsigned)blocks, kThreads, smem, st>>>(p);
    HQ_LAUNCH_OK("k_tile_pass");
    return HQ_OK;
}

int64_t scratch_values_per_item(int log2n, int min_level) {
    const int64_t n_cells = (int64_t)1 << (2 * log2n);
    int64_t v = 0;
    for (int k = min_level; k <= log2n; ++k) v += n_cells >> (2 * k);
    return v;
}

}  // namespace

extern "C" int64_t hq_fused_scratch_bytes(int64_t N, int n, int pyr_mode) {
    if (!hq_is_pow2(n) || n <= 64 || N <= 0) return 0;
    // worst case: the plan references level 1
    return N * scratch_values_per_item(hq_log2(n), 1) * (pyr_mode ? 8 : 4);
}

// The plan lives in device memory; its minimum referenced level is passed through
// plan_len's sign-free companion: callers that know it is >= k can shrink the scratch
// with hq_fused_scratch_bytes_min_level.
extern "C" int64_t hq_fused_scratch_bytes_min_level(int64_t N, int n, int pyr_mode, int min_level) {
    if (!hq_is_pow2(n) || n <= 64 || N <= 0) return 0;
    if (min_level < 1) min_level = 1;
    return N * scratch_values_per_item(hq_log2(n), min_level) * (pyr_mode ? 8 : 4);
}

static int fused_impl(const float* src, int direction, int64_t N, int64_t D, int64_t src_stride, int n, float* grid_out,
                      int64_t grid_stride, float* stream_out, int64_t stream_stride, const int32_t* plan, int plan_len,
                      int pyr_mode, int min_level, void* idx_out, int64_t idx_stride, void* scratch, int64_t scratch_bytes,
                      cudaStream_t st, int64_t D_last = -1) {
    HQ_REQUIRE(hq_is_pow2(n) && n >= 4 && n <= (1 << 15), "fused path needs a power-of-2 grid side in [4, 32768], got %d", n);
    HQ_REQUIRE(direction == 0 || direction == 1, "direction must be 0 or 1");
    HQ_REQUIRE(pyr_mode == 0 || pyr_mode == 1, "pyr_mode must be 0 or 1");
    const int64_t cells = (int64_t)n * n;
    HQ_REQUIRE(D >= 0 && D <= cells, "Too many parameters (%lld) for dimensions %dx%d", (long long)D, n, n);
    HQ_REQUIRE(N >= 0, "negative batch");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(src != nullptr || D == 0, "null src");
    HQ_REQUIRE(plan_len >= 0 && (plan_len == 0 || (plan && idx_out)), "plan/idx_out missing");
    HQ_REQUIRE(direction == 0 ? src_stride >= D : src_stride >= cells, "src stride smaller than row");
    HQ_REQUIRE(!grid_out || grid_stride >= cells, "grid stride smaller than grid");
    HQ_REQUIRE(!stream_out || stream_stride >= D, "stream stride smaller than D");
    HQ_REQUIRE(!(direction == 0 && stream_out), "stream_out is only produced by direction 1");
    HQ_REQUIRE(!(direction == 1 && grid_out), "grid_out is only produced by direction 0");

    TileParams p{};
    p.src = src; p.direction = direction; p.N = N; p.D = D; p.src_stride = src_stride;
    p.D_last = D_last < 0 ? D : D_last;
    p.log2n = hq_log2(n); p.log2t = p.log2n > 6 ? 6 : p.log2n;
    p.grid_out = grid_out; p.grid_stride = grid_stride;
    p.stream_out = stream_out; p.stream_stride = stream_stride;
    p.plan = plan; p.plan_len = plan_len; p.idx_out = idx_out; p.idx_stride = idx_stride;
    p.scratch = scratch; p.min_level = plan_len > 0 ? (min_level < 1 ? 1 : min_level) : 99;
    const bool tiled = p.log2n > 6;
    const uint32_t qpi = (1u << (2 * p.log2t)) / 4;
    p.num_chunks = tiled ? N * ((int64_t)1 << (2 * (p.log2n - 6))) : (N + (kChunkQuads / qpi) - 1) / (kChunkQuads / qpi);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.vec_src = al16(src) && (src_stride % 4 == 0);
    p.vec_grid = grid_out && al16(grid_out) && (grid_stride % 4 == 0);
    p.vec_stream = stream_out && al16(stream_out) && (stream_stride % 4 == 0);
    if (tiled && plan_len > 0) {
        const int64_t need = N * scratch_values_per_item(p.log2n, p.min_level) * (pyr_mode ? 8 : 4);
        HQ_REQUIRE(scratch && scratch_bytes >= need, "scratch too small: need %lld bytes", (long long)need);
        if (p.min_level > 6) p.min_level = 6;   // the tile kernel always publishes its own top level
        const int64_t need2 = N * scratch_values_per_item(p.log2n, p.min_level) * (pyr_mode ? 8 : 4);
        HQ_REQUIRE(scratch_bytes >= need2, "scratch too small: need %lld bytes", (long long)need2);
    }
    int rc;
    // grids up to 64x64 with 128-bit aligned rows take the register-resident fast path
    const bool fast = !tiled && (D % 4 == 0) && p.vec_src && (direction == 0 ? (!grid_out || p.vec_grid) : (!stream_out || p.vec_stream)) &&
                      N * (direction == 0 ? (grid_out ? grid_stride : src_stride) : src_stride) < ((int64_t)1 << 62) &&
                      (int64_t)256 * (src_stride > grid_stride ? src_stride : grid_stride) < ((int64_t)1 << 31) &&
                      (int64_t)256 * stream_stride < ((int64_t)1 << 31);
    if (fast) {
        if (direction == 0 && item_pass::bulk_eligible(p)) rc = pyr_mode ? item_pass::launch_bulk<1>(p, st) : item_pass::launch_bulk<0>(p, st);
        else if (direction == 0) rc = pyr_mode ? item_pass::launch<0, 1>(p, st) : item_pass::launch<0, 0>(p, st);
        else rc = pyr_mode ? item_pass::launch<1, 1>(p, st) : item_pass::launch<1, 0>(p, st);
    } else if (direction == 0 && item_pass::tile_bulk_eligible(p) && p.D_last % 4 == 0) {
        rc = pyr_mode ? item_pass::launch_tile_bulk<1>(p, st) : item_pass::launch_tile_bulk<0>(p, st);
    } else if (p.D_last != p.D) {
        HQ_REQUIRE(false, "a ragged parameter stream needs the bulk tile path (grid side > 64, 16-byte aligned buffers, length %% 4 == 0)");
    } else if (direction == 0) rc = pyr_mode ? launch_tile<0, 1>(p, st) : launch_tile<0, 0>(p, st);
    else rc = pyr_mode ? launch_tile<1, 1>(p, st) : launch_tile<1, 0>(p, st);
    if (rc != HQ_OK) return rc;
    if (tiled && plan_len > 0) {
        if (pyr_mode) k_pyramid_top<1><<<(unsigned)N, 256, 0, st>>>(p);
        else k_pyramid_top<0><<<(unsigned)N, 256, 0, st>>>(p);
        HQ_LAUNCH_OK("k_pyramid_top");
    }
    return HQ_OK;
}

extern "C" int hq_map_index_fused(const float* src, int direction, int64_t N, int64_t D, int64_t src_stride, int n,
                                  float* grid_out, int64_t grid_stride, float* stream_out, int64_t stream_stride,
                                  const int32_t* plan, int plan_len, int pyr_mode, void* idx_out, int64_t idx_stride,
                                  void* scratch, int64_t scratch_bytes, void* stream) {
    return fused_impl(src, direction, N, D, src_stride, n, grid_out, grid_stride, stream_out, stream_stride, plan, plan_len,
                      pyr_mode, 1, idx_out, idx_stride, scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int hq_map_index_fused_ml(const float* src, int direction, int64_t N, int64_t D, int64_t src_stride, int n,
                                     float* grid_out, int64_t grid_stride, float* stream_out, int64_t stream_stride,
                                     const int32_t* plan, int plan_len, int pyr_mode, int min_level, void* idx_out,
                                     int64_t idx_stride, void* scratch, int64_t scratch_bytes, void* stream) {
    return fused_impl(src, direction, N, D, src_stride, n, grid_out, grid_stride, stream_out, stream_stride, plan, plan_len,
                      pyr_mode, min_level, idx_out, idx_stride, scratch, scratch_bytes, (cudaStream_t)stream);
}

// map_to_2d + hierarchical index + uint8 quantisation of the enhanced frame in ONE launch (grids of 32 x 32 and 64 x 64,
// 16-byte aligned rows with D % 4 == 0); HQ_EUNSUPPORTED otherwise (callers run hq_map_index_fused + hq_quantize_u8).
extern "C" int hq_map_index_quant(const float* src, int64_t N, int64_t D, int64_t src_stride, int n, const int32_t* plan, int plan_len,
                                  int pyr_mode, int min_level, int frame_zero, uint8_t* u8_out, int64_t u8_stride, float* minmax,
                                  void* idx_out, int64_t idx_stride, void* stream) {
    HQ_REQUIRE(hq_is_pow2(n) && n >= 4 && n <= (1 << 15), "fused path needs a power-of-2 grid side in [4, 32768], got %d", n);
    HQ_REQUIRE(pyr_mode == 0 || pyr_mode == 1, "pyr_mode must be 0 or 1");
    const int64_t cells = (int64_t)n * n;
    HQ_REQUIRE(D >= 0 && D <= cells, "Too many parameters (%lld) for dimensions %dx%d", (long long)D, n, n);
    HQ_REQUIRE(N >= 0 && plan_len >= 0, "negative size");
    if (N == 0) return HQ_OK;
    HQ_REQUIRE(src && u8_out && minmax && (plan_len == 0 || plan), "null pointer");
    HQ_REQUIRE(src_stride >= D && u8_stride >= cells + plan_len, "stride smaller than row / frame");
    TileParams p{};
    p.src = src; p.direction = 0; p.N = N; p.D = D; p.src_stride = src_stride; p.D_last = D;
    p.log2n = hq_log2(n); p.log2t = p.log2n;
    p.plan = plan; p.plan_len = plan_len; p.idx_out = idx_out; p.idx_stride = idx_stride;
    p.min_level = plan_len > 0 ? (min_level < 1 ? 1 : min_level) : 99;
    p.u8_out = u8_out; p.u8_stride = u8_stride; p.mm_out = minmax; p.frame_zero = (frame_zero || D < cells) ? 1 : 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.vec_src = al16(src) && (src_stride % 4 == 0);
    const bool ok = (n == 32 || n == 64) && p.vec_src && D % 4 == 0 && D >= 4 && al16(u8_out) && u8_stride % 16 == 0 &&
                    (int64_t)256 * src_stride < ((int64_t)1 << 31);
    if (!ok) {
        hq_set_error("hq_map_index_quant: shape not covered by the fused kernel (n = %d, D = %lld)", n, (long long)D);
        return HQ_EUNSUPPORTED;
    }
    const uint32_t qpi = (1u << (2 * p.log2t)) / 4;
    p.num_chunks = (N + (kChunkQuads / qpi) - 1) / (kChunkQuads / qpi);
    return pyr_mode ? item_pass::launch_bulk_quant<1>(p, (cudaStream_t)stream) : item_pass::launch_bulk_quant<0>(p, (cudaStream_t)stream);
}

// Parameter stream (BASELINE config 4): `total` consecutive float32 values cut into ceil(total / n^2) grids of n x n,
// the last one zero padded -- the chunking of core/streaming_processor.py:539-582 + _pad_parameters
// (core/pipeline.py:325-349) + map_to_2d + index, as ONE launch over all grids (the partially filled tail grid used to
// be a second pair of launches).
extern "C" int hq_map_index_stream(const float* src, int64_t total, int n, float* grid_out, const int32_t* plan, int plan_len,
                                   int pyr_mode, int min_level, void* idx_out, int64_t idx_stride, void* scratch,
                                   int64_t scratch_bytes, void* stream) {
    HQ_REQUIRE(total >= 0, "negative stream length");
    HQ_REQUIRE(hq_is_pow2(n) && n > 64, "the stream entry point needs a power-of-2 grid side > 64, got %d", n);
    if (total == 0) return HQ_OK;
    const int64_t cells = (int64_t)n * n;
    const int64_t N = (total + cells - 1) / cells;
    const int64_t D_last = total - (N - 1) * cells;
    return fused_impl(src, 0, N, cells, cells, n, grid_out, cells, nullptr, 0, plan, plan_len, pyr_mode, min_level, idx_out, idx_stride,
                      scratch, scratch_bytes, (cudaStream_t)stream, D_last);
}

// 32-bit words, no arithmetic: the dtype-preserving 4-byte map / unmap.
int hq_tile_map_words(const uint32_t* src, int direction, int64_t N, int64_t D, int64_t src_stride, int n, uint32_t* dst,
                      int64_t dst_stride, cudaStream_t stream) {
    if (direction == 0)
        return fused_impl(reinterpret_cast<const float*>(src), 0, N, D, src_stride, n, reinterpret_cast<float*>(dst), dst_stride,
                          nullptr, 0, nullptr, 0, 0, 99, nullptr, 0, nullptr, 0, stream);
    return fused_impl(reinterpret_cast<const float*>(src), 1, N, D, src_stride, n, nullptr, 0, reinterpret_cast<float*>(dst),
                      dst_stride, nullptr, 0, 0, 99, nullptr, 0, nullptr, 0, stream);
}
