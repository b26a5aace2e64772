// Error plumbing and device probe of the C ABI.
#include "common.cuh"
#include <cstring>

static thread_local char g_err[512] = "";

void mmda_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int mmda_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  mmda_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return MMDA_ERR_CUDA;
}

extern "C" {

const char* mmda_last_error(void) { return g_err; }

int mmda_abi_version(void) { return 1; }

// out[0]=SM count, out[1]=max opt-in smem/block, out[2]=cc major, out[3]=cc minor, out[4]=L2 bytes
int mmda_device_info(int* out5) {
  int dev = 0;
  MMDA_CUDA(cudaGetDevice(&dev));
  MMDA_CUDA(cudaDeviceGetAttribute(&out5[0], cudaDevAttrMultiProcessorCount, dev));
  MMDA_CUDA(cudaDeviceGetAttribute(&out5[1], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  MMDA_CUDA(cudaDeviceGetAttribute(&out5[2], cudaDevAttrComputeCapabilityMajor, dev));
  MMDA_CUDA(cudaDeviceGetAttribute(&out5[3], cudaDevAttrComputeCapabilityMinor, dev));
  MMDA_CUDA(cudaDeviceGetAttribute(&out5[4], cudaDevAttrL2CacheSize, dev));
  return MMDA_OK;
}

}  // extern "C"
