// Status / error plumbing of libtasr_kernels.so.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";
unsigned long long g_tasr_launches = 0;
const unsigned long long* g_tasr_seed_ptr = nullptr;

extern "C" int tasr_set_dropout_seed_ptr(const uint64_t* dev_ptr) {
  g_tasr_seed_ptr = reinterpret_cast<const unsigned long long*>(dev_ptr);
  return TASR_OK;
}

extern "C" const uint64_t* tasr_get_dropout_seed_ptr(void) {
  return reinterpret_cast<const uint64_t*>(g_tasr_seed_ptr);
}

extern "C" uint64_t tasr_launch_count(void) { return g_tasr_launches; }

int tasr_set_cuda_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
  return TASR_ERR_CUDA;
}

extern "C" const char* tasr_last_error(void) { return g_err; }

extern "C" const char* tasr_status_string(int status) {
  switch (status) {
    case TASR_OK: return "ok";
    case TASR_ERR_SHAPE: return "unsupported or inconsistent shape";
    case TASR_ERR_ALIGN: return "misaligned pointer or leading dimension";
    case TASR_ERR_ARCH: return "device is not sm_100 (B200)";
    case TASR_ERR_CUDA: return "CUDA launch/driver error";
    case TASR_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

extern "C" int tasr_version(void) { return 1; }

extern "C" int tasr_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  return major == 10 ? TASR_OK : TASR_ERR_ARCH;
}

extern "C" int tasr_mel_init(void);

extern "C" int tasr_init(void) {
  int rc = tasr_check_device();
  if (rc) return rc;
  return tasr_mel_init();
}
