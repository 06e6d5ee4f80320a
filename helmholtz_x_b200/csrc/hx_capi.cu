// Library-wide state of libhx_b200: error text, version, launch counter.
#include "hx_common.cuh"

namespace hx {
thread_local char g_err[512] = "";
int64_t g_launches = 0;
}  // namespace hx

extern "C" const char* hx_last_error(void) { return hx::g_err; }
extern "C" int hx_version(void) { return 100; }
extern "C" int64_t hx_launch_count(void) { return hx::g_launches; }
extern "C" void hx_launch_count_reset(void) { hx::g_launches = 0; }
extern "C" void hx_launch_count_add(int64_t n) { hx::g_launches += n; }
