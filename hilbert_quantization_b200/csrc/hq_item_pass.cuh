// Fast path of the fused map / unmap + index kernel for grids up to 64x64 (whole items per
// chunk, 128-bit aligned rows).  Same results as k_tile_pass; the differences are all about
// instruction count and latency hiding (ncu on the first version: 5100 warp instructions
// per 64x64 item, issue-bound at 52 % of HBM peak):
//   * every per-thread address (quad position on the curve, the four scatter slots in the
//     tile image, the row segment it stores) is loop invariant and lives in registers;
//   * the tile image and the pyramid are double buffered in shared memory, so one
//     __syncthreads per chunk is enough, and the next chunk's 128-bit loads are issued
//     before the current chunk is stored;
//   * levels >= 4 and the index gather are done by warp 0 while the other warps store.
#pragma once

namespace item_pass {

constexpr int kThreads = 256;
constexpr int kQPT = 4;               // quads per thread per chunk (1024 quads = 4096 positions)

template <int MODE> struct PT { using type = float; };
template <> struct PT<1> { using type = double; };

template <int DIR, int MODE>
__global__ void __launch_bounds__(kThreads, 3) k_item_pass(const TileParams p) {
    using P = typename PT<MODE>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int log2t = p.log2t;
    const uint32_t T = 1u << log2t;
    const uint32_t pitch = T + (T >= 32 ? 4u : 0u);
    const uint32_t cells = 1u << (2 * log2t);
    const uint32_t qpi = cells >> 2;
    const int log2qpi = 2 * log2t - 2;
    const uint32_t ipc = 1024u / qpi;
    const uint32_t pyr_items = (cells - 1) / 3;
    const int top_level = log2t;
    const uint32_t img_floats = ipc * T * pitch;
    const uint32_t pyr_vals = ipc * pyr_items;

    float* const s_img0 = reinterpret_cast<float*>(smem_raw);                       // two tile images
    P* const s_pyr0 = reinterpret_cast<P*>(s_img0 + 2 * img_floats + ((2 * img_floats) & 1));   // two pyramids

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_pyr = p.plan_len > 0 && p.min_level <= 32;
    const uint32_t plan_total = ipc * (uint32_t)p.plan_len;
    const bool plan_by_all = plan_total > 128;

    // ---- loop-invariant per-thread geometry ----
    uint32_t il_[kQPT], q_[kQPT];
    uint32_t slot01[kQPT], slot23[kQPT];      // tile-image offsets of the quad's four cells (16 bit each)
    uint32_t row_s[kQPT];                     // tile-image offset of the row segment this thread stores / loads
    int32_t cur_off[kQPT], row_g[kQPT];       // global offsets (curve side / 2-D side) relative to the chunk's first item
    bool live[kQPT];
#pragma unroll
    for (int r = 0; r < kQPT; ++r) {
        const uint32_t qi = tid + r * kThreads;
        const uint32_t il = qi >> log2qpi, q = qi & (qpi - 1);
        il_[r] = il; q_[r] = q;
        uint32_t x[4], y[4], o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            hq_d2xy(log2t, 4ull * q + i, x[i], y[i]);
            o[i] = il * T * pitch + y[i] * pitch + x[i];
        }
        slot01[r] = o[0] | (o[1] << 16);
        slot23[r] = o[2] | (o[3] << 16);
        const uint32_t vy = q >> (log2t - 2), vx = (q & ((T >> 2) - 1)) << 2;
        row_s[r] = il * T * pitch + vy * pitch + vx;
        const int64_t side2d_stride = DIR == 0 ? p.grid_stride : p.src_stride;
        const int64_t curve_stride = DIR == 0 ? p.src_stride : p.stream_stride;
        row_g[r] = (int32_t)(il * side2d_stride + vy * T + vx);
        cur_off[r] = (int32_t)(il * curve_stride + 4 * q);
        live[r] = DIR == 1 || (int64_t)4 * q < p.D;                  // cells at d >= D are never written (stay zero)
    }

    for (uint32_t i = tid; i < 2 * img_floats; i += kThreads) s_img0[i] = 0.f;
    for (uint32_t i = tid; i < 2 * pyr_vals; i += kThreads) s_pyr0[i] = (P)0;
    __syncthreads();

    auto load_chunk = [&](int64_t chunk, float4 (&v)[kQPT]) {
        const int64_t item0 = chunk * ipc;
        const int64_t left = p.N - item0;
#pragma unroll
        for (int r = 0; r < kQPT; ++r) {
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((int64_t)il_[r] < left) {
                if (DIR == 0) {
                    if (live[r]) val = __ldcs(reinterpret_cast<const float4*>(p.src + item0 * p.src_stride + cur_off[r]));
                } else {
                    val = __ldcs(reinterpret_cast<const float4*>(p.src + item0 * p.src_stride + row_g[r]));
                }
            }
            v[r] = val;
        }
    };

    float4 v[kQPT];
    int64_t chunk = blockIdx.x;
    if (chunk < p.num_chunks) load_chunk(chunk, v);
    int buf = 0;
    for (; chunk < p.num_chunks; chunk += gridDim.x, buf ^= 1) {
        const int64_t item0 = chunk * ipc;
        const int64_t left = p.N - item0;
        float* img = s_img0 + buf * img_floats;
        P* pyrb = s_pyr0 + buf * pyr_vals;

        if (DIR == 0) {
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                if (!live[r]) continue;
                img[slot01[r] & 0xffffu] = v[r].x;
                img[slot01[r] >> 16] = v[r].y;
                img[slot23[r] & 0xffffu] = v[r].z;
                img[slot23[r] >> 16] = v[r].w;
            }
        } else {
#pragma unroll
            for (int r = 0; r < kQPT; ++r) *reinterpret_cast<float4*>(img + row_s[r]) = v[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                v[r] = make_float4(img[slot01[r] & 0xffffu], img[slot01[r] >> 16], img[slot23[r] & 0xffffu], img[slot23[r] >> 16]);
                if (p.stream_out && (int64_t)il_[r] < left && (int64_t)4 * q_[r] < p.D)
                    __stcs(reinterpret_cast<float4*>(p.stream_out + item0 * p.stream_stride + cur_off[r]), v[r]);
            }
        }

        if (want_pyr) {
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                const uint32_t q = q_[r];
                const bool warp_live = DIR == 1 || (int64_t)4 * (q & ~31u) < p.D;      // warp-uniform
                if (!warp_live) continue;
                P* pyr = pyrb + (size_t)il_[r] * pyr_items;
                const P m1 = mean4<MODE>(v[r].x, v[r].y, v[r].z, v[r].w);
                uint32_t base = 0;
                if (top_level < 2) { pyr[q] = m1; continue; }
                if (p.min_level <= 1) pyr[q] = m1;
                base += qpi;
                const P m2 = group_mean<MODE>(m1, 1);
                if (top_level < 3) { if ((lane & 3) == 0) pyr[base + (q >> 2)] = m2; continue; }
                if ((lane & 3) == 0 && p.min_level <= 2) pyr[base + (q >> 2)] = m2;
                base += qpi >> 2;
                const P m3 = group_mean<MODE>(m2, 4);
                if ((lane & 15) == 0) pyr[base + (q >> 4)] = m3;
            }
        }
        __syncthreads();

        // next chunk's loads go out now; they land while this chunk is finished and stored
        float4 nv[kQPT];
        const int64_t next = chunk + gridDim.x;
        if (next < p.num_chunks) load_chunk(next, nv);

        if (want_pyr && warp == 0) {
            uint32_t base_prev = qpi + (qpi >> 2), cnt_prev = qpi >> 4;
            for (int k = 4; k <= top_level; ++k) {
                const uint32_t cnt = cnt_prev >> 2, base = base_prev + cnt_prev;
                for (uint32_t t = lane; t < ipc * cnt; t += 32) {
                    const uint32_t il = t / cnt, j = t - il * cnt;
                    P* pyr = pyrb + (size_t)il * pyr_items;
                    const P a = pyr[base_prev + 4 * j], b = pyr[base_prev + 4 * j + 1];
                    const P c = pyr[base_prev + 4 * j + 2], d = pyr[base_prev + 4 * j + 3];
                    pyr[base + j] = MODE == 0 ? (P)(((a + b) + (c + d)) * (P)0.25) : (P)((((a + b) + c) + d) * (P)0.25);
                }
                __syncwarp();
                base_prev = base;
                cnt_prev = cnt;
            }
        }
        if (p.plan_len > 0 && plan_by_all) __syncthreads();
        if (p.plan_len > 0 && (plan_by_all || warp == 0)) {
            const uint32_t step = plan_by_all ? kThreads : 32;
            for (uint32_t t = plan_by_all ? tid : lane; t < plan_total; t += step) {
                const uint32_t il = t / (uint32_t)p.plan_len, i = t - il * (uint32_t)p.plan_len;
                if ((int64_t)il >= left) continue;
                const int32_t off = __ldg(p.plan + i);
                P val = (P)0;
                if (off >= 0) {
                    if ((uint32_t)off < cells) val = (P)img[il * T * pitch + ((uint32_t)off >> log2t) * pitch + ((uint32_t)off & (T - 1))];
                    else val = pyrb[(size_t)il * pyr_items + ((uint32_t)off - cells)];
                }
                reinterpret_cast<P*>(p.idx_out)[(item0 + il) * p.idx_stride + i] = val;
            }
        }

        if (DIR == 0 && p.grid_out) {
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                if ((int64_t)il_[r] >= left) continue;
                const float4 val = *reinterpret_cast<const float4*>(img + row_s[r]);
                __stcs(reinterpret_cast<float4*>(p.grid_out + item0 * p.grid_stride + row_g[r]), val);
            }
        }
#pragma unroll
        for (int r = 0; r < kQPT; ++r) v[r] = nv[r];
    }
}

inline size_t smem_bytes(int log2t, int mode) {
    const uint32_t T = 1u << log2t;
    const uint32_t pitch = T + (T >= 32 ? 4u : 0u);
    const uint32_t cells = T * T, qpi = cells / 4, ipc = 1024 / qpi;
    return (size_t)2 * ipc * T * pitch * 4 + 8 + (size_t)2 * ipc * ((cells - 1) / 3) * (mode ? 8 : 4) + 16;
}

template <int DIR, int MODE>
int launch(const TileParams& p, cudaStream_t st) {
    const size_t smem = smem_bytes(p.log2t, MODE);
    static bool attr_done = false;
    if (!attr_done) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_item_pass<DIR, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_done = true;
    }
    static int per_sm_cache[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int& per_sm = per_sm_cache[p.log2t & 7];
    if (per_sm == 0) {
        HQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_item_pass<DIR, MODE>, kThreads, smem));
        if (per_sm < 1) per_sm = 1;
    }
    int64_t blocks = (int64_t)hq_cached_sm_count() * per_sm;
    if (blocks > p.num_chunks) blocks = p.num_chunks;
    k_item_pass<DIR, MODE><<<(unsigned)blocks, kThreads, smem, st>>>(p);
    HQ_LAUNCH_OK("k_item_pass");
    return HQ_OK;
}

}  // namespace item_pass
