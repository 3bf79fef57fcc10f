// ABI plumbing: version, thread-local error text, device properties.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

int kge_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int kge_num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached_sms = n;
    cached_dev = dev;
  }
  return cached_sms;
}

extern "C" int kge_abi_version(void) { return KGE_ABI_VERSION; }
extern "C" const char* kge_last_error(void) { return g_last_error; }

extern "C" int kge_copy_h2d_async(void* dst_device, const void* src_host, int64_t nbytes, kge_stream_t stream) {
  KGE_REQUIRE(nbytes >= 0 && (nbytes == 0 || (dst_device && src_host)), KGE_E_ARG, "bad copy arguments");
  if (nbytes == 0) return 0;
  KGE_CUDA(cudaMemcpyAsync(dst_device, src_host, (size_t)nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
