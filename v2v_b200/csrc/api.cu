// Library-level entry points: version, error reporting, device query.
#include "common.cuh"

#include <cstring>

namespace v2v {
namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return V2V_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace v2v

extern "C" int v2v_abi_version(void) { return V2V_ABI_VERSION; }

extern "C" const char* v2v_last_error(void) { return v2v::g_err; }

extern "C" long long v2v_launch_count(void) { return v2v::g_launches.load(std::memory_order_relaxed); }

extern "C" int v2v_device_info(int device, int* cc_major, int* cc_minor, int* sm_count) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    (void)cudaGetLastError();
    v2v::set_error("no CUDA device %d (count %d)", device, n);
    return V2V_ERR_NO_DEVICE;
  }
  cudaDeviceProp p;
  V2V_CUDA(cudaGetDeviceProperties(&p, device));
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  return V2V_OK;
}
