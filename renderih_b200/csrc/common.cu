#include "common.cuh"
#include <stdarg.h>

namespace rih {
static thread_local char g_err[1024] = "";
int g_pdl = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel launch %s failed: %s", what, cudaGetErrorString(e));
    return 3;
  }
  return 0;
}
}  // namespace rih

RIH_API const char* rih_last_error(void) { return rih::g_err; }
RIH_API int rih_version(void) { return 200; }

// Programmatic dependent launch for every kernel of the library (see common.cuh): 1 = on, 0 = plain stream-ordered launches.  Scheduling only:
// results are unchanged (no reference counterpart).
RIH_API int rih_set_pdl(int on) { rih::g_pdl = on ? 1 : 0; return 0; }

// Device properties probe (used by the host side to fail loudly on a non-sm_100 device).
RIH_API int rih_device_info(int device, int* cc_major, int* cc_minor, int* sm_count, size_t* smem_optin) {
  cudaDeviceProp p;
  RIH_CUDA(cudaGetDeviceProperties(&p, device));
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
  return 0;
}
