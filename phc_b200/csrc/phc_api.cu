// Error reporting / version entry points of libphc_b200.so (include/phc_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/phc_b200.h"

static thread_local char g_err[512] = "";

extern "C" void phc_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

extern "C" int phc_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PHC_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return PHC_ERR_CUDA;
}

extern "C" const char* phc_last_error(void) { return g_err; }
extern "C" int phc_version(void) { return 100; }   /* 0.1.0 */
extern "C" int phc_compiled_sm(void) { return 100; }
