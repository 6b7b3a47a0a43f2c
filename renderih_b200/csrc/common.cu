#include "common.cuh"
#include <stdarg.h>

namespace rih {
static thread_local char g_err[1024] = "";
int g_pdl = 0;
int g_reverse = 0;
int g_l2_hints = 0;

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

// Traversal direction of the NEXT launches: 1 = the tensor-core GEMM / convolution kernels walk their output tiles, and the BatchNorm passes
// their rows, from the last to the first.  A consumer that starts where its producer finished finds the most recently written part of an
// activation (whatever of it fits in the 126 MB L2) still cached; alternating the direction kernel by kernel ("serpentine") keeps that true
// along a whole chain of layers.  Scheduling only: results are unchanged up to summation order in atomically accumulated statistics.
RIH_API int rih_set_traversal(int reverse) { rih::g_reverse = reverse ? 1 : 0; return 0; }

// 1 = activations that a GEMM / 1x1 convolution / BatchNorm pass streams through once are loaded with the L2 evict-first priority, so that they do
// not displace the previous kernel's output (which the next kernel is about to read) from the L2.  Scheduling only.
RIH_API int rih_set_l2_hints(int on) { rih::g_l2_hints = on ? 1 : 0; return 0; }

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
