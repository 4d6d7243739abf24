// api.cu — library identification / error strings for the C-ABI in include/goslam_b200.h.
#include "common.cuh"

static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void gs_note_cuda_error(cudaError_t e) { g_last_cuda_error = e; }

extern "C" {

const char* goslam_last_cuda_error(void) {
  return g_last_cuda_error == cudaSuccess ? "" : cudaGetErrorString(g_last_cuda_error);
}

int goslam_version(void) { return 100; }

int goslam_sm_arch(void) {
#if defined(GOSLAM_SM_ARCH)
  return GOSLAM_SM_ARCH;
#else
  return 100;
#endif
}

const char* goslam_strerror(int code) {
  switch (code) {
    case GOSLAM_OK: return "ok";
    case GOSLAM_EINVAL: return "invalid argument or shape";
    case GOSLAM_ELAUNCH: return "CUDA launch failed";
    case GOSLAM_EWORKSPACE: return "workspace missing or too small";
    case GOSLAM_EUNSUPPORTED: return "entry point not supported (training-only backward)";
    default: return "unknown error";
  }
}

}  // extern "C"
