// Fast path of the fused map / unmap + index kernel for grids up to 64x64 (whole items per
// chunk, 128-bit aligned rows).  Same results as k_tile_pass; the differences are all about
// instruction count and latency hiding.  ncu history of the 64x64 / D=1536 / variant-C case:
//   v1 k_tile_pass            5100 warp instructions per item, issue bound, 52 % of HBM peak
//   v2 runtime-geometry fast  3950 (integer address math + parameter reloads dominated)
//   v3 this file: the grid side is a template parameter, every per-thread address is a
//      loop-invariant 32-bit offset from one per-chunk base pointer, full chunks skip all
//      bounds checks, the tile image / pyramid are double buffered (one __syncthreads per
//      chunk), the next chunk's 128-bit loads are issued before the current chunk is stored,
//      and levels >= 4 plus the index gather run on warp 0 while the other warps store.
#pragma once

namespace item_pass {

constexpr int kThreads = 256;          // data threads (own the quads / row segments)
constexpr int kBlock = kThreads + 32;  // + one index warp: levels >= 4 and the entries that read them
constexpr int kQPT = 4;               // quads per thread per chunk (1024 quads = 4096 positions)
constexpr int kPlanCap = 1024;        // gather plans up to this many entries are staged in shared memory

template <int MODE> struct PT { using type = float; };
template <> struct PT<1> { using type = double; };

template <int LOG2T>
struct Geo {
    static constexpr uint32_t T = 1u << LOG2T;
    static constexpr uint32_t pitch = T + (T >= 32 ? 4u : 0u);
    static constexpr uint32_t cells = T * T;
    static constexpr uint32_t qpi = cells / 4;
    static constexpr int log2qpi = 2 * LOG2T - 2;
    static constexpr uint32_t ipc = 1024u / qpi;
    static constexpr uint32_t pyr_items = (cells - 1) / 3;
    static constexpr uint32_t img_floats = ipc * T * pitch;
    static constexpr uint32_t pyr_vals = ipc * pyr_items;
    static constexpr uint32_t base2 = qpi, base3 = qpi + qpi / 4;
};


// Levels 1..3 of the run-mean pyramid straight from the quads a warp holds in registers: a
// thread's quad is level 1, 4 / 16 consecutive lanes are levels 2 / 3 (two shuffle steps each,
// same association as the shared-memory tree: ((a+b)+(c+d)) fp32, (((a+b)+c)+d) fp64).  Dead
// quads (d >= D) hold zeros, exactly what the zero padding contributes.  `warp_live` bit r is
// set when any lane of the warp holds data for quad slot r (warp-uniform skip).
template <int MODE, int LOG2T>
__device__ __forceinline__ void pyramid_levels_123(const float4 (&v)[kQPT], uint32_t live_bits, uint32_t warp_live,
                                                   typename PT<MODE>::type* pyrb, int tid, int lane, int min_level = 1) {
    using P = typename PT<MODE>::type;
    using G = Geo<LOG2T>;
#pragma unroll
    for (int r = 0; r < kQPT; ++r) {
        if (!((warp_live >> r) & 1u)) continue;
        const uint32_t qi = tid + r * kThreads;
        const uint32_t il = qi >> G::log2qpi, q = qi & (G::qpi - 1);
        const bool live = (live_bits >> r) & 1u;
        P* pyr = pyrb + il * G::pyr_items;
        const P m1 = mean4<MODE>(v[r].x, v[r].y, v[r].z, v[r].w);
        // levels below the lowest one the plan references are only needed in registers (variant C reads levels >= 3)
        if (live && min_level <= 1) pyr[q] = m1;
        const P m2 = group_mean<MODE>(m1, 1);
        if (live && (lane & 3) == 0 && min_level <= 2) pyr[G::base2 + (q >> 2)] = m2;
        if (LOG2T >= 3) {
            const P m3 = group_mean<MODE>(m2, 4);
            if (live && (lane & 15) == 0) pyr[G::base3 + (q >> 4)] = m3;
        }
    }
}

// Levels >= 4 (from level 3 in shared memory) by one warp, all items of the chunk in one flat loop.
template <int MODE, int LOG2T>
__device__ __forceinline__ void pyramid_levels_top(typename PT<MODE>::type* pyrb, uint32_t live1, int lane) {
    using P = typename PT<MODE>::type;
    using G = Geo<LOG2T>;
    if (LOG2T < 4) return;
    uint32_t base_prev = G::base3, cnt_prev = G::qpi >> 4;
    uint32_t live_prev = (((live1 + 3) >> 2) + 3) >> 2;          // live runs of level 3
#pragma unroll
    for (int k = 4; k <= LOG2T; ++k) {
        const uint32_t cnt = cnt_prev >> 2, base = base_prev + cnt_prev;
        const uint32_t live = (live_prev + 3) >> 2;
        for (uint32_t e = lane; e < G::ipc * live; e += 32) {
            const uint32_t il = e / live, j = e - il * live;
            P* pyr = pyrb + il * G::pyr_items;
            const P a = pyr[base_prev + 4 * j], b = pyr[base_prev + 4 * j + 1];
            const P c = pyr[base_prev + 4 * j + 2], d = pyr[base_prev + 4 * j + 3];
            pyr[base + j] = MODE == 0 ? (P)(((a + b) + (c + d)) * (P)0.25) : (P)((((a + b) + c) + d) * (P)0.25);
        }
        __syncwarp();
        base_prev = base;
        cnt_prev = cnt;
        live_prev = live;
    }
}

template <int DIR, int MODE, int LOG2T>
__global__ void __launch_bounds__(kBlock, 3) k_item_pass(const TileParams p) {
    using P = typename PT<MODE>::type;
    using G = Geo<LOG2T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* const s_img0 = reinterpret_cast<float*>(smem_raw);                                   // two tile images
    P* const s_pyr0 = reinterpret_cast<P*>(s_img0 + 2 * G::img_floats + ((2 * G::img_floats) & 1));   // two pyramids
    int32_t* const s_plan = reinterpret_cast<int32_t*>(s_pyr0 + 2 * G::pyr_vals);                 // gather plan (<= kPlanCap)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Warp 8 never touches global data rows: it finishes the pyramid of chunk i (levels >= 4) and
    // writes the index entries that read those levels while warps 0-7 already store chunk i and
    // scatter chunk i+1.  ncu showed the data warps waiting ~17 % of the time at the chunk barrier
    // for the one warp that carried this serial tail.
    const bool idx_warp = warp == kThreads / 32;
    const bool want_pyr = p.plan_len > 0 && p.min_level <= 32;
    const int64_t side2d_stride = DIR == 0 ? p.grid_stride : p.src_stride;
    const int64_t curve_stride = DIR == 0 ? p.src_stride : p.stream_stride;

    // ---- loop-invariant per-thread geometry ----
    uint32_t slot01[kQPT], slot23[kQPT];      // tile-image offsets of the quad's four cells (16 bit each)
    uint32_t row_s[kQPT];                     // tile-image offset of the row segment this thread stores / loads
    int32_t cur_off[kQPT], row_g[kQPT];       // global offsets (curve side / 2-D side) from the chunk's first item
    uint32_t live_bits = 0;                   // bit r: quad r holds data (d < D)
#pragma unroll
    for (int r = 0; r < kQPT; ++r) {
        const uint32_t qi = tid + r * kThreads;
        const uint32_t il = qi >> G::log2qpi, q = qi & (G::qpi - 1);
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t x, y;
            hq_d2xy(LOG2T, 4ull * q + i, x, y);
            o[i] = il * G::T * G::pitch + y * G::pitch + x;
        }
        slot01[r] = o[0] | (o[1] << 16);
        slot23[r] = o[2] | (o[3] << 16);
        const uint32_t vy = q >> (LOG2T - 2), vx = (q & ((G::T >> 2) - 1)) << 2;
        row_s[r] = il * G::T * G::pitch + vy * G::pitch + vx;
        row_g[r] = (int32_t)(il * side2d_stride + vy * G::T + vx);
        cur_off[r] = (int32_t)(il * curve_stride + 4 * q);
        if (!idx_warp && (DIR == 1 || (int64_t)4 * q < p.D)) live_bits |= 1u << r;
    }

    uint32_t warp_live = 0;
#pragma unroll
    for (int r = 0; r < kQPT; ++r)
        if (__any_sync(0xffffffffu, (live_bits >> r) & 1u)) warp_live |= 1u << r;

    const bool plan_in_smem = p.plan_len <= kPlanCap;
    if (plan_in_smem)
        for (int i = tid; i < p.plan_len; i += kBlock) s_plan[i] = __ldg(p.plan + i);
    for (uint32_t i = tid; i < 2 * G::img_floats; i += kBlock) s_img0[i] = 0.f;
    for (uint32_t i = tid; i < 2 * G::pyr_vals; i += kBlock) s_pyr0[i] = (P)0;
    __syncthreads();

    // 2-D side (DIR 1) or curve side (DIR 0) quads of one chunk; `left` = items remaining
    auto load_chunk = [&](int64_t chunk, float4 (&v)[kQPT]) {
        const int64_t item0 = chunk * G::ipc;
        const int64_t left = p.N - item0;
        const float* base = p.src + item0 * p.src_stride;
        const bool full = left >= (int64_t)G::ipc;
#pragma unroll
        for (int r = 0; r < kQPT; ++r) {
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t il = (tid + r * kThreads) >> G::log2qpi;
            if (((live_bits >> r) & 1u) && (full || (int64_t)il < left))
                val = __ldcs(reinterpret_cast<const float4*>(base + (DIR == 0 ? cur_off[r] : row_g[r])));
            v[r] = val;
        }
    };

    // hand-off counters between the data warps and the index warp (monotonic, per CTA):
    //   s_done  = number of chunks whose index entries the index warp has finished
    // s_free[b]: the index warp has finished the chunk that used pyramid / image buffer b (one phase per
    // use).  An mbarrier, not a polled flag: __nanosleep granularity (~1 us) used to set the CTA's
    // iteration time whenever the data warps arrived a little early.
    __shared__ __align__(8) uint64_t s_free[2];
    if (tid == 0) {
        hq_tc::mbar_init(&s_free[0], 1);
        hq_tc::mbar_init(&s_free[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    float4 v[kQPT];
    int64_t chunk = blockIdx.x;
    if (chunk < p.num_chunks) load_chunk(chunk, v);
    uint32_t buf = 0, iter = 0;
    for (; chunk < p.num_chunks; chunk += gridDim.x, buf ^= 1u, ++iter) {
        const int64_t item0 = chunk * G::ipc;
        const int64_t left = p.N - item0;
        const bool full = left >= (int64_t)G::ipc;
        float* img = s_img0 + buf * G::img_floats;
        P* pyrb = s_pyr0 + buf * G::pyr_vals;

        // the buffers of this chunk were last used by chunk iter-2: its index entries must be out
        if (!idx_warp && p.plan_len > 0 && iter >= 2) hq_tc::mbar_wait(&s_free[iter & 1u], ((iter >> 1) - 1u) & 1u);
        // ---- data warps: curve-side quads -> tile image (or tile image -> quads for DIR 1) ----
        if (DIR == 0) {
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                if ((live_bits >> r) & 1u) {
                    img[slot01[r] & 0xffffu] = v[r].x;
                    img[slot01[r] >> 16] = v[r].y;
                    img[slot23[r] & 0xffffu] = v[r].z;
                    img[slot23[r] >> 16] = v[r].w;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < kQPT; ++r)
                if (!idx_warp) *reinterpret_cast<float4*>(img + row_s[r]) = v[r];
        }
        if (!idx_warp) asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");   // image complete (data warps only)

        // next chunk's loads go out now; they land while this chunk is finished and stored
        float4 nv[kQPT];
        const int64_t next = chunk + gridDim.x;
        if (next < p.num_chunks && !idx_warp) load_chunk(next, nv);

        if (!idx_warp) {
            // ---- 2-D / curve side out first (fire and forget), the pyramid chain afterwards ----
            if (DIR == 0) {
                if (p.grid_out) {
                    float* gbase = p.grid_out + item0 * p.grid_stride;
#pragma unroll
                    for (int r = 0; r < kQPT; ++r) {
                        const uint32_t il = (tid + r * kThreads) >> G::log2qpi;
                        if (full || (int64_t)il < left) {
                            const float4 val = *reinterpret_cast<const float4*>(img + row_s[r]);
                            __stcs(reinterpret_cast<float4*>(gbase + row_g[r]), val);
                        }
                    }
                }
            } else {
                float* sbase = p.stream_out ? p.stream_out + item0 * p.stream_stride : nullptr;
#pragma unroll
                for (int r = 0; r < kQPT; ++r) {
                    v[r] = make_float4(img[slot01[r] & 0xffffu], img[slot01[r] >> 16], img[slot23[r] & 0xffffu], img[slot23[r] >> 16]);
                    const uint32_t qi = tid + r * kThreads;
                    const uint32_t il = qi >> G::log2qpi, q = qi & (G::qpi - 1);
                    if (sbase && (full || (int64_t)il < left) && (int64_t)4 * q < p.D)
                        __stcs(reinterpret_cast<float4*>(sbase + cur_off[r]), v[r]);
                }
            }
            // levels 1..3 in registers / shuffles on the data warps (they have slack: the kernel is bound by
            // the serial tail of the index warp otherwise -- "index only" used to take as long as "map only")
            if (want_pyr) pyramid_levels_123<MODE, LOG2T>(v, live_bits, warp_live, pyrb, tid, lane);
            // levels 1..3 (and the tile image) of this chunk are in shared memory: hand them to the index
            // warp without waiting for it
            // Hand-off = named barrier, alternating ids 1 / 3 per chunk: the data warps arrive without
            // waiting (bar.arrive orders their shared-memory writes for the thread that syncs), the index
            // warp blocks in hardware -- no polling, no fence on the data warps.  Two ids are enough
            // because the s_done wait above keeps the data warps at most two chunks ahead.
            if (p.plan_len > 0) {
                if (iter & 1u) asm volatile("bar.arrive 3, %0;" ::"n"(kBlock) : "memory");
                else asm volatile("bar.arrive 1, %0;" ::"n"(kBlock) : "memory");
            }
        } else if (p.plan_len > 0) {
            // ---- index warp: wait for the 8 data warps of this chunk, finish the pyramid, write the entries ----
            if (iter & 1u) asm volatile("bar.sync 3, %0;" ::"n"(kBlock) : "memory");
            else asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");
            // runs at d >= D are never touched: they stay zero from the one-time clear
            if (want_pyr) pyramid_levels_top<MODE, LOG2T>(pyrb, DIR == 1 ? G::qpi : (uint32_t)((p.D + 3) >> 2), lane);
            P* out = reinterpret_cast<P*>(p.idx_out) + item0 * p.idx_stride;
            const uint32_t n_items = full ? G::ipc : (uint32_t)left;
            for (uint32_t e = lane; e < n_items * (uint32_t)p.plan_len; e += 32) {
                const uint32_t il = e / (uint32_t)p.plan_len, i = e - il * (uint32_t)p.plan_len;
                const float* im = img + il * G::T * G::pitch;
                const P* pyr = pyrb + il * G::pyr_items;
                const int32_t off = plan_in_smem ? s_plan[i] : __ldg(p.plan + i);
                P val = (P)0;
                if (off >= 0) {
                    if ((uint32_t)off < G::cells) val = (P)im[((uint32_t)off >> LOG2T) * G::pitch + ((uint32_t)off & (G::T - 1))];
                    else val = pyr[(uint32_t)off - G::cells];
                }
                __stcs(out + il * p.idx_stride + i, val);
            }
            __syncwarp();
            if (lane == 0) hq_tc::mbar_arrive(&s_free[iter & 1u]);
        }
#pragma unroll
        for (int r = 0; r < kQPT; ++r) v[r] = nv[r];
    }
}


// ---------------------------------------------------------------------------------------
// v5/v6 (map direction, dense grid output): both global sides go through the bulk-copy engine.
//
// v5: the 2-D side leaves the SM as ONE 16 KB bulk copy per chunk (cp.async.bulk shared ->
// global).  A microbenchmark with this kernel's traffic shape (tools/microbench/hbm_mix.cu:
// 6144 B in, 16720 B out per item) reaches 5.8-6.0 TB/s with LSU stores and 6.25 TB/s with
// bulk stores.  The tile image is kept in plain global order (row pitch T, no padding): the
// scatter pays 4-5-way bank conflicts, which at ~0.2 shared-memory wavefronts per clock is
// irrelevant.  (A first attempt with 128B-swizzled tensor stores, 2 boxes of 32 x 64 floats,
// was slower than v3: 4.6 TB/s -- 128-byte inner extents are a poor shape for the TMA store
// path.)  Three images: the store of chunk i has finished reading its image before chunk
// i + 3 scatters into it, and that knowledge rides on the existing barrier.
//
// v6: v5 alone stayed at 5.4 TB/s -- with one chunk prefetched in registers per CTA there
// are only ~3 source rows in flight per SM against ~1.5 us of loaded HBM latency.  The
// source rows now arrive by bulk loads (global -> shared, mbarrier completion) through a
// ring of `stages` staging buffers that one thread keeps full, so the loads in flight no
// longer cost registers or depend on the CTA count.
// ---------------------------------------------------------------------------------------
constexpr int kMaxStages = 8;

// QUANT (north star (2): "quantization and encoding of those indices fused into the same kernel"): the uint8 min/max
// quantisation of the ENHANCED frame (grid + index rows, core/compressor.py:256-280 applied to the output of
// embed_indices_in_image, core/pipeline.py:140-146) leaves the same pass.  The frame's min / max are those of the grid
// cells (plus 0 when the frame holds structural zeros: padding cells, padded index slots): every index value is either a
// cell or a mean of cells, and rounding is monotonic, so fl(mean) stays inside [min, max] of its cells.  The data warps
// reduce min / max from the quads they hold anyway, quantise the dense tile image (16 cells per thread, one 128-bit store)
// and the index warp quantises the index rows it gathers: 4 D bytes in, n^2 + plan_len + 8 bytes out per item, no second pass.
// The reference evaluates ((v - min) / (max - min) * 255).astype(uint8) as three float32 operations and a truncation
// (core/compressor.py:274).  An IEEE division per cell (~10 instructions) made the fused pass issue bound at twice the
// instruction count of the index pass, so the correctly rounded quotient comes from the item's correctly rounded reciprocal
// y = RN(1 / range), computed once, and one Newton correction with an exact FMA residual (Markstein: q0 = RN(d y) is faithful,
// rem = d - range q0 is exact, RN(q0 + rem y) = RN(d / range) when nothing under- or overflows).  `y == 0` selects the plain
// division (ranges below 1e-30 or above 1e30, where y or the residual could leave the normal range).
__device__ __forceinline__ float quant_recip(float range) {
    return (range > 1.0e-30f && range < 1.0e30f) ? __frcp_rn(range) : 0.f;
}
__device__ __forceinline__ uint32_t quant_u8(float v, float mn, float range, float y, bool constant) {
    if (constant) return 128u;
    const float d = __fsub_rn(v, mn);
    float a;
    if (y != 0.f) {
        const float q0 = __fmul_rn(d, y);
        a = __fmaf_rn(__fmaf_rn(-range, q0, d), y, q0);
    } else {
        a = __fdiv_rn(d, range);
    }
    // truncation, like astype(uint8): 0 <= a * 255 <= 255, so adding 2^23 with round-toward-zero leaves floor(a * 255) in the
    // low mantissa bits (a full-rate FADD instead of a quarter-rate F2I per cell)
    return __float_as_uint(__fadd_rz(__fmul_rn(a, 255.0f), 8388608.0f)) & 0xffu;
}

template <int MODE, int LOG2T, bool QUANT>
__global__ void __launch_bounds__(kBlock, 3) k_item_pass_bulk(const TileParams p, const int kRing, const int stages, const uint32_t stage_floats,
                                                              const int plan_cap) {
    using P = typename PT<MODE>::type;
    using G = Geo<LOG2T>;
    constexpr uint32_t kImg = 4096;                                 // floats per chunk image (ipc items of T x T)
    constexpr int IPC = (int)G::ipc;                                // items per chunk (1 for 64 x 64, 4 for 32 x 32)
    __shared__ float s_mm[QUANT ? 2 : 1][QUANT ? IPC : 1][QUANT ? kThreads / 32 : 1][2];   // per-warp (min, max) partials, two chunks
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* const s_img0 = reinterpret_cast<float*>(smem_raw);                                    // kRing dense tile images
    float* const s_stage0 = s_img0 + kRing * kImg;                                               // `stages` source staging buffers
    P* const s_pyr0 = reinterpret_cast<P*>(s_stage0 + (size_t)stages * stage_floats);            // two pyramids
    int32_t* const s_plan = reinterpret_cast<int32_t*>(s_pyr0 + 2 * G::pyr_vals);
    __shared__ __align__(8) uint64_t s_full[kMaxStages];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool idx_warp = warp == kThreads / 32;
    const bool want_pyr = p.plan_len > 0 && p.min_level <= 32;
    const uint32_t row_bytes = (uint32_t)p.D * 4u;

    uint32_t slot01[kQPT], slot23[kQPT];
    uint32_t st_off[kQPT];                     // float offset of the quad inside a staging buffer
    uint32_t live_bits = 0;
#pragma unroll
    for (int r = 0; r < kQPT; ++r) {
        const uint32_t qi = tid + r * kThreads;
        const uint32_t il = qi >> G::log2qpi, q = qi & (G::qpi - 1);
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t x, y;
            hq_d2xy(LOG2T, 4ull * q + i, x, y);
            o[i] = il * G::cells + y * G::T + x;
        }
        slot01[r] = o[0] | (o[1] << 16);
        slot23[r] = o[2] | (o[3] << 16);
        st_off[r] = il * (uint32_t)p.D + 4 * q;
        if (!idx_warp && (int64_t)4 * q < p.D) live_bits |= 1u << r;
    }

    uint32_t warp_live = 0;
#pragma unroll
    for (int r = 0; r < kQPT; ++r)
        if (__any_sync(0xffffffffu, (live_bits >> r) & 1u)) warp_live |= 1u << r;

    const bool plan_in_smem = p.plan_len <= plan_cap;
    if (plan_in_smem)
        for (int i = tid; i < p.plan_len; i += kBlock) s_plan[i] = __ldg(p.plan + i);
    for (uint32_t i = tid; i < (uint32_t)kRing * kImg; i += kBlock) s_img0[i] = 0.f;
    for (uint32_t i = tid; i < 2 * G::pyr_vals; i += kBlock) s_pyr0[i] = (P)0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zero fill is read by the bulk stores too

    // one thread keeps the staging ring full: source rows of chunk `chunk` -> stage buffer `st`
    auto issue_loads = [&](int64_t chunk, int st) {
        const int64_t item0 = chunk * G::ipc;
        const int64_t left = p.N - item0;
        const uint32_t n_items = left >= (int64_t)G::ipc ? G::ipc : (uint32_t)left;
        const uint32_t bar = hq_tc::smem_u32(&s_full[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n_items * row_bytes) : "memory");
        float* dst = s_stage0 + (size_t)st * stage_floats;
        if (p.src_stride == p.D || n_items == 1) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             hq_tc::smem_u32(dst)),
                         "l"(p.src + item0 * p.src_stride), "r"(n_items * row_bytes), "r"(bar)
                         : "memory");
        } else {
            for (uint32_t il = 0; il < n_items; ++il)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 hq_tc::smem_u32(dst + il * (uint32_t)p.D)),
                             "l"(p.src + (item0 + il) * p.src_stride), "r"(row_bytes), "r"(bar)
                             : "memory");
        }
    };

    __shared__ __align__(8) uint64_t s_free[2];         // see k_item_pass
    if (tid == 0) {
        hq_tc::mbar_init(&s_free[0], 1);
        hq_tc::mbar_init(&s_free[1], 1);
        for (int i = 0; i < stages; ++i) hq_tc::mbar_init(&s_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        int64_t c = blockIdx.x;
        for (int i = 0; i < stages && c < p.num_chunks; ++i, c += gridDim.x) issue_loads(c, i);
    }

    uint32_t ring = 0, iter = 0, st = 0, st_phase = 0;
    for (int64_t chunk = blockIdx.x; chunk < p.num_chunks; chunk += gridDim.x, ring = (ring + 1 == (uint32_t)kRing ? 0 : ring + 1), ++iter) {
        const int64_t item0 = chunk * G::ipc;
        const int64_t left = p.N - item0;
        const bool full = left >= (int64_t)G::ipc;
        float* img = s_img0 + ring * kImg;
        P* pyrb = s_pyr0 + (iter & 1u) * G::pyr_vals;
        float4 v[kQPT];

        if (!idx_warp) {
            // ---- this chunk's source rows have landed in the staging buffer ----
            hq_tc::mbar_wait(&s_full[st], st_phase);
            const float* stg = s_stage0 + (size_t)st * stage_floats;
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t il = (tid + r * kThreads) >> G::log2qpi;
                if (((live_bits >> r) & 1u) && (full || (int64_t)il < left)) val = *reinterpret_cast<const float4*>(stg + st_off[r]);
                v[r] = val;
            }
            // the index warp reads pyramid (iter & 1): it must be done with chunk iter-2 (the image of
            // chunk iter-3 was released even earlier)
            if (p.plan_len > 0 && iter >= 2) hq_tc::mbar_wait(&s_free[iter & 1u], ((iter >> 1) - 1u) & 1u);
            if constexpr (QUANT) {
                // (min, max) of the live quads per item; quad slot r belongs to item (r * kThreads) >> log2qpi
                float mn[IPC], mx[IPC];
#pragma unroll
                for (int i = 0; i < IPC; ++i) { mn[i] = FLT_MAX; mx[i] = -FLT_MAX; }
#pragma unroll
                for (int r = 0; r < kQPT; ++r) {
                    constexpr int dummy = 0; (void)dummy;
                    const int il = (r * kThreads) >> G::log2qpi;
                    if ((live_bits >> r) & 1u) {
                        mn[il] = fminf(mn[il], fminf(fminf(v[r].x, v[r].y), fminf(v[r].z, v[r].w)));
                        mx[il] = fmaxf(mx[il], fmaxf(fmaxf(v[r].x, v[r].y), fmaxf(v[r].z, v[r].w)));
                    }
                }
#pragma unroll
                for (int i = 0; i < IPC; ++i) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        mn[i] = fminf(mn[i], __shfl_xor_sync(0xffffffffu, mn[i], o));
                        mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
                    }
                    if (lane == 0) { s_mm[iter & 1u][i][warp][0] = mn[i]; s_mm[iter & 1u][i][warp][1] = mx[i]; }
                }
            }
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                if ((live_bits >> r) & 1u) {
                    img[slot01[r] & 0xffffu] = v[r].x;
                    img[slot01[r] >> 16] = v[r].y;
                    img[slot23[r] & 0xffffu] = v[r].z;
                    img[slot23[r] >> 16] = v[r].w;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            // store(iter - kRing + 1) has finished reading its image before anybody passes this barrier,
            // so chunk iter+1 may scatter into it
            if (tid == 0) {
                if (kRing == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else if (kRing == 3) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else if (kRing == 4) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            }
            asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");
            if (tid == 0) {
                if (p.grid_out) {
                    const uint32_t bytes = full ? kImg * 4u : (uint32_t)left * G::cells * 4u;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.grid_out + item0 * (int64_t)G::cells),
                                 "r"(hq_tc::smem_u32(img)), "r"(bytes)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                // every data thread has read staging buffer st: refill it
                const int64_t nx = chunk + (int64_t)stages * gridDim.x;
                if (nx < p.num_chunks) issue_loads(nx, (int)st);
            }
            if constexpr (QUANT) {
                // 16 consecutive cells of the dense image per thread -> 16 bytes, one 128-bit store
                const uint32_t c0 = (uint32_t)tid * 16u, il = c0 >> (2 * LOG2T), cell = c0 & (G::cells - 1);
                if (full || (int64_t)il < left) {
                    float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
                    for (int w = 0; w < kThreads / 32; ++w) { mn = fminf(mn, s_mm[iter & 1u][il][w][0]); mx = fmaxf(mx, s_mm[iter & 1u][il][w][1]); }
                    if (p.frame_zero) { mn = fminf(mn, 0.f); mx = fmaxf(mx, 0.f); }
                    const bool constant = mn == mx;
                    const float range = __fsub_rn(mx, mn), y = quant_recip(range);
                    uint32_t w4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 c = *reinterpret_cast<const float4*>(img + c0 + 4 * j);
                        w4[j] = quant_u8(c.x, mn, range, y, constant) | (quant_u8(c.y, mn, range, y, constant) << 8) |
                                (quant_u8(c.z, mn, range, y, constant) << 16) | (quant_u8(c.w, mn, range, y, constant) << 24);
                    }
                    uint8_t* frame = p.u8_out + (item0 + il) * p.u8_stride;
                    __stcs(reinterpret_cast<uint4*>(frame + cell), make_uint4(w4[0], w4[1], w4[2], w4[3]));
                    if (cell == 0) { p.mm_out[2 * (item0 + il)] = mn; p.mm_out[2 * (item0 + il) + 1] = mx; }
                }
            }
            if (want_pyr) pyramid_levels_123<MODE, LOG2T>(v, live_bits, warp_live, pyrb, tid, lane);
            if (p.plan_len > 0) {
                if (iter & 1u) asm volatile("bar.arrive 3, %0;" ::"n"(kBlock) : "memory");
                else asm volatile("bar.arrive 1, %0;" ::"n"(kBlock) : "memory");
            }
        } else if (p.plan_len > 0) {
            if (iter & 1u) asm volatile("bar.sync 3, %0;" ::"n"(kBlock) : "memory");
            else asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");
            if (want_pyr) pyramid_levels_top<MODE, LOG2T>(pyrb, (uint32_t)((p.D + 3) >> 2), lane);
            P* out = reinterpret_cast<P*>(p.idx_out) + item0 * p.idx_stride;
            const uint32_t n_items = full ? G::ipc : (uint32_t)left;
            float q_mn[IPC], q_rg[IPC];
            bool q_const[IPC];
            if constexpr (QUANT) {
#pragma unroll
                for (int i = 0; i < IPC; ++i) {
                    float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
                    for (int w = 0; w < kThreads / 32; ++w) { mn = fminf(mn, s_mm[iter & 1u][i][w][0]); mx = fmaxf(mx, s_mm[iter & 1u][i][w][1]); }
                    if (p.frame_zero) { mn = fminf(mn, 0.f); mx = fmaxf(mx, 0.f); }
                    q_mn[i] = mn; q_rg[i] = __fsub_rn(mx, mn); q_const[i] = mn == mx;
                }
            }
            for (uint32_t e = lane; e < n_items * (uint32_t)p.plan_len; e += 32) {
                const uint32_t il = e / (uint32_t)p.plan_len, i = e - il * (uint32_t)p.plan_len;
                const int32_t off = plan_in_smem ? s_plan[i] : __ldg(p.plan + i);
                P val = (P)0;
                if (off >= 0) {
                    if ((uint32_t)off < G::cells) val = (P)img[il * G::cells + off];
                    else val = pyrb[il * G::pyr_items + (uint32_t)off - G::cells];
                }
                if (!QUANT || p.idx_out) __stcs(out + il * p.idx_stride + i, val);
                if constexpr (QUANT) {
                    // the index row is embedded in the image's dtype (float32) before the frame is normalised
                    float mn = q_mn[0], rg = q_rg[0];
                    bool cst = q_const[0];
#pragma unroll
                    for (int j = 1; j < IPC; ++j)
                        if ((int)il == j) { mn = q_mn[j]; rg = q_rg[j]; cst = q_const[j]; }
                    p.u8_out[(item0 + il) * p.u8_stride + G::cells + i] = (uint8_t)quant_u8((float)val, mn, rg, quant_recip(rg), cst);
                }
            }
            __syncwarp();
            if (lane == 0) hq_tc::mbar_arrive(&s_free[iter & 1u]);
        }
        if (++st == (uint32_t)stages) { st = 0; st_phase ^= 1u; }
    }
    // the images must outlive the stores that read them
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// true when the bulk-copy variant can take this call (no grid output, or a dense 16-byte aligned one;
// rows that are whole 16-byte units; few enough items per chunk that one thread can issue their loads)
inline bool bulk_eligible(const TileParams& p) {
    return p.direction == 0 && p.vec_src && p.log2t >= 5 && p.D >= 4 &&
           (!p.grid_out || (p.vec_grid && p.grid_stride == ((int64_t)1 << (2 * p.log2t))));
}

inline int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <int MODE, int LOG2T, bool QUANT = false>
int launch_bulk_t(const TileParams& p, cudaStream_t st) {
    using G = Geo<LOG2T>;
    // images: kRing chunk images when a grid is written (the stores read them asynchronously), two otherwise
    // staging: as many chunk-sized source buffers as fit the per-CTA budget
    // Measured on B200 (tools/bench_item_pass.py, 262144 items): what matters is CTAs per SM (4 x 55 KB:
    // 6.2-6.3 TB/s for 1536-D/64x64 and 768-D/32x32; 2 x 110 KB with 7 stages: 5.3 TB/s), so the
    // default is the smallest footprint: two images, two staging buffers.  Tuning knobs, not API.
    static const int ring_cfg = env_int("HQ_ITEM_RING", 2), stages_cfg = env_int("HQ_ITEM_STAGES", 0),
                     budget_kb = env_int("HQ_ITEM_SMEM_KB", 55), ctas_cap = env_int("HQ_ITEM_CTAS", 8);
    const int ring = p.grid_out ? ring_cfg : 2;
    const uint32_t stage_floats = (uint32_t)(G::ipc * p.D + 31) & ~31u;
    const int plan_cap = (p.plan_len <= kPlanCap ? (p.plan_len + 31) & ~31 : 0);
    const size_t fixed = (size_t)ring * 4096 * 4 + (size_t)2 * G::pyr_vals * (MODE ? 8 : 4) + (size_t)plan_cap * 4 + 64;
    const size_t budget = (size_t)budget_kb * 1024;
    int stages = budget > fixed ? (int)((budget - fixed) / ((size_t)stage_floats * 4)) : 2;
    if (stages_cfg > 0) stages = stages_cfg;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    const size_t smem = fixed + (size_t)stages * stage_floats * 4;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        HQ_CUDA_OK(cudaFuncSetAttribute((k_item_pass_bulk<MODE, LOG2T, QUANT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    int per_sm = 0;
    HQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k_item_pass_bulk<MODE, LOG2T, QUANT>), kBlock, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > ctas_cap) per_sm = ctas_cap;
    int64_t blocks = (int64_t)hq_cached_sm_count() * per_sm;
    if (blocks > p.num_chunks) blocks = p.num_chunks;
    k_item_pass_bulk<MODE, LOG2T, QUANT><<<(unsigned)blocks, kBlock, smem, st>>>(p, ring, stages, stage_floats, plan_cap);
    HQ_LAUNCH_OK(QUANT ? "k_item_pass_bulk<quant>" : "k_item_pass_bulk");
    return HQ_OK;
}

template <int MODE>
int launch_bulk(const TileParams& p, cudaStream_t st) {
    return p.log2t == 5 ? launch_bulk_t<MODE, 5>(p, st) : launch_bulk_t<MODE, 6>(p, st);
}

template <int MODE>
int launch_bulk_quant(const TileParams& p, cudaStream_t st) {
    return p.log2t == 5 ? launch_bulk_t<MODE, 5, true>(p, st) : launch_bulk_t<MODE, 6, true>(p, st);
}

// ---------------------------------------------------------------------------------------
// Big grids (n > 64), map direction: the same bulk-copy pipeline per 64x64 TILE.  A chunk is an aligned
// 4096-run of one item's stream = one tile whose inner order is the n = 64 order composed with the
// (swap, flip) pair of tile_frame.  Source: one 16 KB bulk load (clipped at D).  Tile out: ONE 2-D tensor
// store (box 64 x 64 floats, rows 256 B apart by n * 4 B in global memory).  The pyramid of the tile
// (levels 1..6) is published to the per-item scratch for k_pyramid_top, like k_tile_pass does.
// k_tile_pass (no prefetch, five barriers per tile) ran the 494 M-parameter stream of config C4 at
// 2.8 TB/s.
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock, 3) k_tile_pass_bulk(const __grid_constant__ CUtensorMap map_grid, const TileParams p,
                                                              const int kRing, const int stages) {
    using P = typename PT<MODE>::type;
    using G = Geo<6>;
    constexpr uint32_t kImg = 4096;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    // The tile images are the source of 128-byte-SWIZZLED tensor stores (two 32-column halves per tile): the scatter of a
    // warp's 32 quads into the dense image cost 4.75 shared-memory wavefronts per store (the cells of 128 consecutive curve
    // positions share 8-16 columns = banks), 2.0 with the 16-byte chunks of a row XORed by (row & 7).  1024-byte aligned.
    unsigned char* const smem_raw = smem_dyn + ((1024u - (hq_tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
    float* const s_img0 = reinterpret_cast<float*>(smem_raw);                     // kRing tile images (swizzled halves)
    float* const s_stage0 = s_img0 + kRing * kImg;                                // `stages` source staging buffers
    P* const s_pyr0 = reinterpret_cast<P*>(s_stage0 + (size_t)stages * kImg);     // two pyramids
    __shared__ __align__(8) uint64_t s_full[kMaxStages];
    __shared__ __align__(8) uint64_t s_free[2];
    // byte offset of cell (x, y) inside a tile image: half (x >> 5) | row y of 128 bytes | 16-byte chunk ^ (y & 7) | (x & 3)
    auto cell_off = [](uint32_t x, uint32_t y) -> uint32_t {
        const uint32_t xx = x & 31u;
        return (x >> 5) * 8192u + y * 128u + ((((xx >> 2) ^ (y & 7u)) << 4) | ((xx & 3u) << 2));
    };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool idx_warp = warp == kThreads / 32;
    const bool want_pyr = p.plan_len > 0 && p.min_level <= 32;
    const int upper_bits = p.log2n - 6;
    const int tile_shift = 2 * upper_bits;                         // tiles per item = 4^upper_bits: shifts, no 64-bit divisions
    const int64_t tile_mask = ((int64_t)1 << tile_shift) - 1;
    const int64_t n_cells = (int64_t)1 << (2 * p.log2n);

    // tile-image offsets of the quad's four cells, unswapped (y * 64 + x) and swapped (x * 64 + y)
    uint32_t slotA01[kQPT], slotA23[kQPT], slotB01[kQPT], slotB23[kQPT];
#pragma unroll
    for (int r = 0; r < kQPT; ++r) {
        const uint32_t q = tid + r * kThreads;
        uint32_t a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t x, y;
            hq_d2xy(6, 4ull * q + i, x, y);
            a[i] = cell_off(x, y);                     // BYTE offsets inside the tile image (< 16384: 16 bits each)
            b[i] = cell_off(y, x);
        }
        slotA01[r] = a[0] | (a[1] << 16); slotA23[r] = a[2] | (a[3] << 16);
        slotB01[r] = b[0] | (b[1] << 16); slotB23[r] = b[2] | (b[3] << 16);
    }
    for (uint32_t i = tid; i < 2 * G::pyr_vals; i += kBlock) s_pyr0[i] = (P)0;

    // floats of the chunk that exist in the source (the rest of the tile is zero padding)
    auto chunk_valid = [&](int64_t chunk) -> uint32_t {
        const int64_t item = chunk >> tile_shift, tile = chunk & tile_mask;
        const int64_t left = (item == p.N - 1 ? p.D_last : p.D) - tile * 4096;
        return left <= 0 ? 0u : (left >= 4096 ? 4096u : (uint32_t)left);
    };
    auto issue_load = [&](int64_t chunk, int st) {
        const int64_t item = chunk >> tile_shift, tile = chunk & tile_mask;
        const uint32_t bytes = chunk_valid(chunk) * 4u;
        const uint32_t bar = hq_tc::smem_u32(&s_full[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        if (bytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             hq_tc::smem_u32(s_stage0 + (size_t)st * kImg)),
                         "l"(p.src + item * p.src_stride + tile * 4096), "r"(bytes), "r"(bar)
                         : "memory");
    };

    // (swap, flip, tile origin) of a chunk: a loop over the upper bit pairs of the tile index.  Every thread of the CTA
    // used to run it for every chunk (together with two 64-bit divisions: half of the kernel's 4300 warp instructions
    // per tile, 67 % issue utilisation at 68 % of the DRAM peak); now thread 32 computes the NEXT chunk's frame while
    // the others scatter, and everybody reads four words from shared memory.
    __shared__ uint32_t s_frame[2][4];
    auto put_frame = [&](int64_t chunk, uint32_t slot) {
        uint32_t X, Y, swp, flp;
        tile_frame(upper_bits, (uint64_t)(chunk & tile_mask), X, Y, swp, flp);
        s_frame[slot][0] = X; s_frame[slot][1] = Y; s_frame[slot][2] = swp; s_frame[slot][3] = flp;
    };
    if (tid == 32 && (int64_t)blockIdx.x < p.num_chunks) put_frame(blockIdx.x, 0);
    if (tid == 0) {
        hq_tc::mbar_init(&s_free[0], 1);
        hq_tc::mbar_init(&s_free[1], 1);
        for (int i = 0; i < stages; ++i) hq_tc::mbar_init(&s_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        int64_t c = blockIdx.x;
        for (int i = 0; i < stages && c < p.num_chunks; ++i, c += gridDim.x) issue_load(c, i);
    }

    int64_t sc_item = 0;                                           // scratch values per item (levels min_level .. log2n)
    for (int k = p.min_level; k <= p.log2n; ++k) sc_item += n_cells >> (2 * k);

    uint32_t ring = 0, iter = 0, st = 0, st_phase = 0;
    for (int64_t chunk = blockIdx.x; chunk < p.num_chunks; chunk += gridDim.x, ring = (ring + 1 == (uint32_t)kRing ? 0 : ring + 1), ++iter) {
        const int64_t item = chunk >> tile_shift;
        const uint64_t tile = (uint64_t)(chunk & tile_mask);
        const uint32_t X = s_frame[iter & 1u][0], Y = s_frame[iter & 1u][1], swp = s_frame[iter & 1u][2], flp = s_frame[iter & 1u][3];
        const uint32_t valid = chunk_valid(chunk);
        float* img = s_img0 + ring * kImg;
        P* pyrb = s_pyr0 + (iter & 1u) * G::pyr_vals;

        if (!idx_warp) {
            hq_tc::mbar_wait(&s_full[st], st_phase);
            const float* stg = s_stage0 + (size_t)st * kImg;
            float4 v[kQPT];
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                const uint32_t q = tid + r * kThreads;
                v[r] = 4 * q < valid ? *reinterpret_cast<const float4*>(stg + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (p.plan_len > 0 && iter >= 2) hq_tc::mbar_wait(&s_free[iter & 1u], ((iter >> 1) - 1u) & 1u);
            // the frame of the next chunk (read after this iteration's barrier; slot (iter + 1) & 1 was last read in
            // iteration iter - 1, before that iteration's barrier)
            if (tid == 32 && chunk + (int64_t)gridDim.x < p.num_chunks) put_frame(chunk + gridDim.x, (iter + 1u) & 1u);
            // every cell of the tile is written (zeros beyond D): the images are reused by tiles with other fills.
            // A flipped tile mirrors cell (x, y) to (63 - x, 63 - y): half, row and (x & 3) flip, the swizzled chunk stays
            // (both of its XOR terms flip) -- byte offset ^ 0x3f8c: one LOP3 per cell.
            const uint32_t fx = flp ? 0x3f8cu : 0u;
            char* const imgb = reinterpret_cast<char*>(img);
#pragma unroll
            for (int r = 0; r < kQPT; ++r) {
                const uint32_t s01 = swp ? slotB01[r] : slotA01[r], s23 = swp ? slotB23[r] : slotA23[r];
                *reinterpret_cast<float*>(imgb + ((s01 & 0xffffu) ^ fx)) = v[r].x;
                *reinterpret_cast<float*>(imgb + ((s01 >> 16) ^ fx)) = v[r].y;
                *reinterpret_cast<float*>(imgb + ((s23 & 0xffffu) ^ fx)) = v[r].z;
                *reinterpret_cast<float*>(imgb + ((s23 >> 16) ^ fx)) = v[r].w;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tid == 0) {
                if (kRing == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");
            if (tid == 0) {
                if (p.grid_out) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&map_grid)),
                                     "r"(hq_tc::smem_u32(img) + 8192u * h), "r"((int32_t)(X * 64 + 32 * h)),
                                     "r"((int32_t)(item * ((int64_t)1 << p.log2n) + Y * 64))
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                const int64_t nx = chunk + (int64_t)stages * gridDim.x;
                if (nx < p.num_chunks) issue_load(nx, (int)st);
            }
            if (want_pyr) pyramid_levels_123<MODE, 6>(v, 0xfu, 0xfu, pyrb, tid, lane, p.min_level);
            if (p.plan_len > 0) {
                if (iter & 1u) asm volatile("bar.arrive 3, %0;" ::"n"(kBlock) : "memory");
                else asm volatile("bar.arrive 1, %0;" ::"n"(kBlock) : "memory");
            }
        } else if (p.plan_len > 0) {
            if (iter & 1u) asm volatile("bar.sync 3, %0;" ::"n"(kBlock) : "memory");
            else asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");
            if (want_pyr) {
                pyramid_levels_top<MODE, 6>(pyrb, G::qpi, lane);
                // publish the tile's levels [min_level .. 6] to the per-item scratch pyramid
                P* dst = reinterpret_cast<P*>(p.scratch) + item * sc_item;
                int64_t lvl_base_g = 0;
                uint32_t lvl_base_s = 0;
                for (int k = 1; k <= 6; ++k) {
                    const uint32_t cnt = 4096u >> (2 * k);
                    if (k >= p.min_level) {
                        for (uint32_t t = lane; t < cnt; t += 32) dst[lvl_base_g + (int64_t)tile * cnt + t] = pyrb[lvl_base_s + t];
                        lvl_base_g += n_cells >> (2 * k);
                    }
                    lvl_base_s += cnt;
                }
            }
            __syncwarp();
            if (lane == 0) hq_tc::mbar_arrive(&s_free[iter & 1u]);
        }
        if (++st == (uint32_t)stages) { st = 0; st_phase ^= 1u; }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// big grids: dense 16-byte aligned grid output (or none), whole 16-byte units in the source
inline bool tile_bulk_eligible(const TileParams& p) {
    return p.direction == 0 && p.log2n > 6 && p.log2n <= 15 && p.vec_src && (p.D % 4 == 0) && (p.src_stride % 4 == 0) &&
           (!p.grid_out || (p.vec_grid && p.grid_stride == ((int64_t)1 << (2 * p.log2n)) && hq_tc::encode_tiled_fn() != nullptr &&
                            p.N * ((int64_t)1 << p.log2n) < ((int64_t)1 << 31)));
}

template <int MODE>
int launch_tile_bulk(const TileParams& p, cudaStream_t st) {
    using G = Geo<6>;
    static const int ring = env_int("HQ_TILE_RING", 2), stages = env_int("HQ_TILE_STAGES", 2);
    const size_t smem = (size_t)(ring + stages) * 4096 * 4 + (size_t)2 * G::pyr_vals * (MODE ? 8 : 4) + 64 + 1024;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_tile_pass_bulk<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    int per_sm = 0;
    HQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_pass_bulk<MODE>, kBlock, smem));
    if (per_sm < 1) per_sm = 1;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (p.grid_out) {
        const int64_t n = (int64_t)1 << p.log2n;
        const int rc = hq_tc::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.grid_out, p.N * n, n, n, 32, 64);   // 128-byte swizzle
        if (rc != HQ_OK) return rc;
    }
    int64_t blocks = (int64_t)hq_cached_sm_count() * per_sm;
    if (blocks > p.num_chunks) blocks = p.num_chunks;
    k_tile_pass_bulk<MODE><<<(unsigned)blocks, kBlock, smem, st>>>(map, p, ring, stages);
    HQ_LAUNCH_OK("k_tile_pass_bulk");
    return HQ_OK;
}

template <int LOG2T>
inline size_t smem_bytes(int mode) {
    using G = Geo<LOG2T>;
    return (size_t)2 * G::img_floats * 4 + 8 + (size_t)2 * G::pyr_vals * (mode ? 8 : 4) + (size_t)kPlanCap * 4 + 16;
}

template <int DIR, int MODE, int LOG2T>
int launch_t(const TileParams& p, cudaStream_t st) {
    const size_t smem = smem_bytes<LOG2T>(MODE);
    static int per_sm = 0;
    if (per_sm == 0) {
        HQ_CUDA_OK(cudaFuncSetAttribute(k_item_pass<DIR, MODE, LOG2T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        HQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_item_pass<DIR, MODE, LOG2T>, kBlock, smem));
        if (per_sm < 1) per_sm = 1;
    }
    int64_t blocks = (int64_t)hq_cached_sm_count() * per_sm;
    if (blocks > p.num_chunks) blocks = p.num_chunks;
    k_item_pass<DIR, MODE, LOG2T><<<(unsigned)blocks, kBlock, smem, st>>>(p);
    HQ_LAUNCH_OK("k_item_pass");
    return HQ_OK;
}

template <int DIR, int MODE>
int launch(const TileParams& p, cudaStream_t st) {
    switch (p.log2t) {
        case 2: return launch_t<DIR, MODE, 2>(p, st);
        case 3: return launch_t<DIR, MODE, 3>(p, st);
        case 4: return launch_t<DIR, MODE, 4>(p, st);
        case 5: return launch_t<DIR, MODE, 5>(p, st);
        default: return launch_t<DIR, MODE, 6>(p, st);
    }
}

}  // namespace item_pass
