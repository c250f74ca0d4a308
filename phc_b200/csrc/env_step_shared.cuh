// Pieces shared by the fused env-step kernels (env_step.cu: one warp per env; env_step_packed.cu: four envs per three warps).
#pragma once
#include <stdint.h>

#include "../../include/phc_b200.h"
#include "phc_math.cuh"

namespace phc {

constexpr int kBodyRec = 13;

// Experiment builds (tools/timeline_env.py): per-env %globaltimer stamps at the phase boundaries of a kernel, 8 slots per env.
// Each translation unit keeps its own buffer pointer (static: no relocatable device code in this library).
#ifdef PHC_EXP_TIMELINE
static __device__ unsigned long long* g_timeline = nullptr;
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define PHC_TL_IF(cond, env, k) do { if ((cond) && g_timeline) g_timeline[(size_t)(env) * 8 + (k)] = gtimer(); } while (0)
#else
#define PHC_TL_IF(cond, env, k) do { } while (0)
#endif
#define PHC_TL(k) PHC_TL_IF(lane == 0, env, k)

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }


struct BodyRec { V3 p; Q4 q; V3 v; V3 w; };

__device__ __forceinline__ BodyRec load_body(const float* s) {
  BodyRec b;
  b.p = v3(s[0], s[1], s[2]);
  b.q = q4(s[3], s[4], s[5], s[6]);
  b.v = v3(s[7], s[8], s[9]);
  b.w = v3(s[10], s[11], s[12]);
  return b;
}

// two-frame blend of one body: lerp pos(+offset)/vel/angvel, slerp rot (motion_lib_base.py:474-488)
__device__ __forceinline__ BodyRec blend_body(const float* s0, const float* s1, float bl, V3 off) {
  const BodyRec a = load_body(s0), b = load_body(s1);
  const float omb = 1.0f - bl;
  BodyRec r;
  r.p = lerp3(a.p, b.p, omb, bl) + off;
  r.v = lerp3(a.v, b.v, omb, bl);
  r.w = lerp3(a.w, b.w, omb, bl);
  r.q = slerp(a.q, b.q, bl);
  return r;
}

__device__ __forceinline__ void st3(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
__device__ __forceinline__ void st6(float* d, TanNorm t) { st3(d, t.t); st3(d + 3, t.n); }

// The steady-state launch of the shipped SMPL configuration (phc_env_step checks every condition): flags exactly kFastFlags,
// pose cache on, no env mask, per-env motion records given, every row movable as a TMA bulk copy, no ref_* side buffers.
constexpr uint32_t kFastFlags = PHC_FLAG_UPRIGHT | PHC_FLAG_LOCAL_ROOT_OBS | PHC_FLAG_ROOT_HEIGHT_OBS | PHC_FLAG_POWER_REWARD |
                                PHC_FLAG_EARLY_TERM | PHC_FLAG_REWARD_FROM_CACHE;

}  // namespace phc
