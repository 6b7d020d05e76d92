// abi.cu -- error plumbing and device queries shared by the C-ABI entry points.
#include <stdarg.h>
#include <stdlib.h>
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

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("RLB_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

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

int rlb_l2_persist(const void *ptr, size_t bytes, rlb_stream_t stream) {
  using namespace rlb;
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  int max_persist = 0, max_window = 0;
  rc = check_cuda(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev),
                  "cudaDeviceGetAttribute(MaxPersistingL2CacheSize)");
  if (rc) return rc;
  rc = check_cuda(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev),
                  "cudaDeviceGetAttribute(MaxAccessPolicyWindowSize)");
  if (rc) return rc;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (ptr == nullptr || bytes == 0 || max_persist <= 0 || max_window <= 0) {
    attr.accessPolicyWindow.num_bytes = 0;  // clears the window
    check_cuda(cudaStreamSetAttribute(as_stream(stream), cudaStreamAttributeAccessPolicyWindow, &attr),
               "cudaStreamSetAttribute(clear)");
    return 0;
  }
  size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
  rc = check_cuda(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want), "cudaDeviceSetLimit(PersistingL2)");
  if (rc) return rc;
  const size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
  attr.accessPolicyWindow.base_ptr = const_cast<void *>(ptr);
  attr.accessPolicyWindow.num_bytes = win;
  attr.accessPolicyWindow.hitRatio = want >= win ? 1.0f : (float)want / (float)win;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  rc = check_cuda(cudaStreamSetAttribute(as_stream(stream), cudaStreamAttributeAccessPolicyWindow, &attr),
                  "cudaStreamSetAttribute(AccessPolicyWindow)");
  if (rc) return rc;
  return (int)(want >> 20) + 1;  // > 0: MiB of L2 set aside (+1)
}

}  // extern "C"
