// Library-level entry points: version, error text, device properties.
#include "hq_common.cuh"
#include <mutex>
#include <atomic>

static thread_local char g_err[512] = "";

void hq_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int hq_cached_sm_count() {
    static int sms[64];
    static std::once_flag once;
    std::call_once(once, [] { for (int i = 0; i < 64; ++i) sms[i] = 0; });
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

static std::atomic<long long> g_launches{0};
void hq_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t hq_launch_count(int reset) {
    return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

extern "C" int hq_version(void) { return HQ_ABI_VERSION; }
extern "C" const char* hq_last_error(void) { return g_err; }
extern "C" int hq_sm_count(void) { return hq_cached_sm_count(); }

// ---- per-kernel timing for bench.py (roofline of ONE kernel: CUDA events on the launching stream around its launches) ----
// hq_kernel_timing(1) arms it; launchers bracket the kernels they name with hq_time_begin / hq_time_end (an event pair from
// a small pool per slot); hq_kernel_timing_read synchronises the events and returns the summed milliseconds and the launch
// count per slot, then disarms.  Slots: 0 = k_rerank_tc, 1 = k_filter_bits_tc (window or full pass), 2 = k_filter_cascade_win
// / k_filter_cascade_lists, 3 = k_rerank_tc_merge.  Off (the default) costs one relaxed load per launch.
namespace {
constexpr int kTimeSlots = 4, kTimePairs = 512;
std::atomic<int> g_time_on{0};
cudaEvent_t g_ev[kTimeSlots][kTimePairs][2];
int g_ev_n[kTimeSlots];
bool g_ev_made = false;
std::mutex g_time_mu;
}  // namespace

extern "C" int hq_kernel_timing(int on) {
    std::lock_guard<std::mutex> lk(g_time_mu);
    if (on && !g_ev_made) {
        for (int s = 0; s < kTimeSlots; ++s)
            for (int i = 0; i < kTimePairs; ++i) {
                HQ_CUDA_OK(cudaEventCreate(&g_ev[s][i][0]));
                HQ_CUDA_OK(cudaEventCreate(&g_ev[s][i][1]));
            }
        g_ev_made = true;
    }
    for (int s = 0; s < kTimeSlots; ++s) g_ev_n[s] = 0;
    g_time_on.store(on ? 1 : 0, std::memory_order_relaxed);
    return HQ_OK;
}

int hq_time_begin(int slot, cudaStream_t st) {
    if (!g_time_on.load(std::memory_order_relaxed) || slot < 0 || slot >= kTimeSlots) return -1;
    std::lock_guard<std::mutex> lk(g_time_mu);
    if (g_ev_n[slot] >= kTimePairs) return -1;
    const int i = g_ev_n[slot]++;
    cudaEventRecord(g_ev[slot][i][0], st);
    return i;
}

void hq_time_end(int slot, int token, cudaStream_t st) {
    if (token < 0) return;
    cudaEventRecord(g_ev[slot][token][1], st);
}

extern "C" int hq_kernel_timing_read(float* ms_out, int32_t* n_out, int slots) {
    HQ_REQUIRE(ms_out && n_out && slots >= 0, "bad arguments");
    std::lock_guard<std::mutex> lk(g_time_mu);
    g_time_on.store(0, std::memory_order_relaxed);
    for (int s = 0; s < slots; ++s) {
        ms_out[s] = 0.f; n_out[s] = 0;
        if (s >= kTimeSlots) continue;
        for (int i = 0; i < g_ev_n[s]; ++i) {
            float ms = 0.f;
            HQ_CUDA_OK(cudaEventSynchronize(g_ev[s][i][1]));
            HQ_CUDA_OK(cudaEventElapsedTime(&ms, g_ev[s][i][0], g_ev[s][i][1]));
            ms_out[s] += ms;
            ++n_out[s];
        }
        g_ev_n[s] = 0;
    }
    return HQ_OK;
}
