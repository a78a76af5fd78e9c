// Error text, version and device probing of the C ABI.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace orb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
HostCallWorkspace& host_call_workspace() {
    static thread_local HostCallWorkspace ws;
    return ws;
}
}  // namespace orb

extern "C" {
const char* orb_last_error(void) { return orb::g_err; }
const char* orb_version(void) { return "orb_b200 0.1 (sm_100a)"; }
long long orb_launch_count(void) { return orb::g_launches.load(); }
int orb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
}
