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
