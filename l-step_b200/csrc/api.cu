// Status / error plumbing of the C ABI (include/lstep_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace lstep {

static thread_local char g_cuda_err[256] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
}

bool pdl_enabled() {
  static const bool on = getenv("LSTEP_NO_PDL") == nullptr;
  return on;
}

int check_launch(const char* where) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return LSTEP_OK;
  set_cuda_error(e, where);
  return LSTEP_ERR_CUDA;
}

}  // namespace lstep

extern "C" const char* lstep_strerror(int status) {
  switch (status) {
    case LSTEP_OK: return "ok";
    case LSTEP_ERR_INVALID_ARG: return "invalid argument";
    case LSTEP_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
    case LSTEP_ERR_WORKSPACE: return "workspace too small";
    case LSTEP_ERR_CUDA: return "CUDA runtime error";
    case LSTEP_ERR_ID_RANGE: return "id does not fit the int32 device encoding";
    default: return "unknown status";
  }
}

extern "C" const char* lstep_last_cuda_error(void) { return lstep::g_cuda_err; }

extern "C" int lstep_abi_version(void) { return LSTEP_ABI_VERSION; }

extern "C" int lstep_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}
