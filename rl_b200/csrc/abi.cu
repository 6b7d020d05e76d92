// abi.cu -- error plumbing and device queries shared by the C-ABI entry points.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace rlb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return RLB_OK;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return RLB_ENODEV;
  return RLB_ECUDA;
}

int check_launch(const char *kernel) { return check_cuda(cudaPeekAtLastError(), kernel); }

int sm_count() {
  // cached per device ordinal (one process per GPU is the deployment model, but be correct anyway)
  static int cache[64];
  int dev = 0;
  if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice") != RLB_OK) return -1;
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    if (check_cuda(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute") !=
        RLB_OK)
      return -1;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace rlb

extern "C" {

int rlb_version(void) { return RLB_VERSION; }

const char *rlb_last_error(void) { return rlb::g_err; }

int rlb_device_sm_count(void) { return rlb::sm_count(); }

}  // extern "C"
