// Library-level entry points of the C ABI (include/comat_b200.h).
#include "common.cuh"

static thread_local int g_last_cuda_error = 0;

extern "C" void comat_set_cuda_error(int e) { g_last_cuda_error = e; }
extern "C" int comat_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" int comat_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* comat_strerror(int status) {
  switch (status) {
    case COMAT_OK: return "ok";
    case COMAT_ERR_INVALID: return "invalid argument";
    case COMAT_ERR_UNSUPPORTED: return "unsupported shape/dtype";
    case COMAT_ERR_CUDA: return "CUDA error (see comat_last_cuda_error)";
    case COMAT_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}
