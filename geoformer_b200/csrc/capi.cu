// C-ABI plumbing: error state, init, launch counter.
#include "common.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>

namespace gf {
int init_driver(int device);
extern std::atomic<int64_t> g_launches;
}  // namespace gf

static thread_local char g_err[512] = "";

int gf_set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
  return code;
}

extern "C" int gf_abi_version(void) { return GF_ABI_VERSION; }
extern "C" const char* gf_last_error(void) { return g_err; }
extern "C" int gf_init(int device) { return gf::init_driver(device); }
extern "C" int64_t gf_launch_count(void) { return gf::g_launches.load(); }
