// Development aids, compiled only by `python -m sba_gan_b200.build --dev` (-DSBA_DEV_AIDS): tuning knobs from the
// environment and a globaltimer timeline of the kernels of consecutive ABI calls (tools/timeline.py).  The product
// library contains none of this: no getenv, no device synchronisation, no extra symbols.
#ifdef SBA_DEV_AIDS
#include <cstdlib>
#include <mutex>

#include "kernels.h"
#include "tc5_common.cuh"

namespace sba {
namespace tc5 {

const DevTuning& dev_tuning() {
    static const DevTuning t = [] {
        auto get = [](const char* name) { const char* v = getenv(name); return v ? atoi(v) : -1; };
        return DevTuning{get("SBA_TC5_CTAS_PER_SM"), get("SBA_TC5_STATIC"), get("SBA_TC5_CHUNK"), get("SBA_TC5_LATE_TRIGGER"), get("SBA_TC5_VARIANT")};
    }();
    return t;
}

namespace {
constexpr int kCalls = 256, kStamps = 16;
std::mutex g_mu;
unsigned long long* g_buf = nullptr;
int g_next = 0;
bool g_on = false;
}  // namespace

unsigned long long* timeline_slot() {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_on || g_next >= kCalls) return nullptr;
    return g_buf + (size_t)(g_next++) * kStamps;
}

}  // namespace tc5
}  // namespace sba

extern "C" {
static void tl_init_values() {
    using namespace sba::tc5;
    static unsigned long long init[kCalls * kStamps];
    for (int i = 0; i < kCalls * kStamps; ++i) init[i] = (i & 1) ? 0ull : ~0ull;
    cudaDeviceSynchronize();
    cudaMemcpy(g_buf, init, sizeof(init), cudaMemcpyHostToDevice);
}
// start handing out timeline slots to the next (up to 256) ABI calls - eager launches or launches being captured
__attribute__((visibility("default"))) int sba_dev_timeline_start(void) {
    using namespace sba::tc5;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_buf == nullptr && cudaMalloc(&g_buf, kCalls * kStamps * sizeof(unsigned long long)) != cudaSuccess) return -1;
    tl_init_values();
    g_next = 0;
    g_on = true;
    return 0;
}
// stop handing out slots (calls made from now on are not stamped); returns the number of slots handed out
__attribute__((visibility("default"))) int sba_dev_timeline_stop(void) {
    using namespace sba::tc5;
    std::lock_guard<std::mutex> lock(g_mu);
    g_on = false;
    return g_next;
}
// clear the stamps (e.g. between the capture of a CUDA graph and the replay that is to be measured)
__attribute__((visibility("default"))) int sba_dev_timeline_clear(void) {
    using namespace sba::tc5;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_buf == nullptr) return -1;
    tl_init_values();
    return 0;
}
// copy the stamps of the first n slots to `out` (16 values per call)
__attribute__((visibility("default"))) int sba_dev_timeline_read(unsigned long long* out, int n) {
    using namespace sba::tc5;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_buf == nullptr) return -1;
    if (n > kCalls) n = kCalls;
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_buf, (size_t)n * kStamps * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return n;
}
}
#endif
