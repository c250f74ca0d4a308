/* Minimal C host of libphc_b200.so: proves the boundary is a plain C ABI (no C++ / torch types in the signatures).
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -o /tmp/c_host -Lphc_b200/lib -lphc_b200 -Wl,-rpath,$PWD/phc_b200/lib && /tmp/c_host
 *
 * Without a GPU it exercises the entry points that do no device work (sizes, strides, error reporting); with one it also
 * runs phc_gae on a tiny rollout through the CUDA runtime API the caller owns.  A Go (cgo), Rust (bindgen) or Julia host
 * binds the same header the same way; the reference itself is Python, so its binding is ctypes (phc_b200/_lib.py). */
#include <stdio.h>
#include <string.h>

#include "phc_b200.h"

int main(void) {
  printf("phc_b200 version %d, compiled for sm_%d\n", phc_version(), phc_compiled_sm());
  /* shapes of the shipped SMPL configuration: obs 358 + 576, AMP 196 */
  const uint32_t flags = PHC_FLAG_UPRIGHT | PHC_FLAG_LOCAL_ROOT_OBS | PHC_FLAG_ROOT_HEIGHT_OBS;
  const int self_dim = phc_self_obs_dim(24, flags), task_dim = phc_task_obs_dim(24, 1), amp_dim = phc_amp_obs_dim(19, 4, flags);
  printf("self obs %d, task obs %d, amp obs %d, body record stride %d floats\n", self_dim, task_dim, amp_dim, phc_motion_body_stride(24));
  if (self_dim != 358 || task_dim != 576 || amp_dim != 196 || phc_motion_body_stride(24) != 312) return 1;
  /* error convention: negative code + thread-local message, nothing thrown */
  const int rc = phc_env_step(NULL, NULL);
  printf("phc_env_step(NULL) -> %d (%s)\n", rc, phc_last_error());
  if (rc != PHC_ERR_INVALID_ARG || strlen(phc_last_error()) == 0) return 2;
  PhcStepArgs a;
  memset(&a, 0, sizeof a);
  a.num_envs = 0;              /* an empty batch is a successful no-op */
  if (phc_env_step(&a, NULL) != PHC_OK) return 3;
  printf("sizeof(PhcStepArgs) = %zu, sizeof(PhcMotionLib) = %zu\n", sizeof(PhcStepArgs), sizeof(PhcMotionLib));
  return 0;
}
