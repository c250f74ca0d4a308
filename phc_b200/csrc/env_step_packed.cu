// Fused env step, packed mapping: FOUR environments per THREE warps (96 threads = 4 x 24 bodies, every lane carries a body).
//
// The one-warp-per-env kernel of env_step.cu leaves 8 of 32 lanes idle at 24 bodies: 25 % of every issue slot of an
// instruction-issue bound kernel (profiles/env_step_r1c_ncu.md: 22.0 of 32 lanes active per instruction).  This kernel is the
// compile-time specialised steady state of the shipped SMPL configuration (kFastFlags, J = 24, T = 1, pose cache, AMP ring slot,
// rows as TMA bulk copies -- exactly the launches phc_env_step sends to env_step_kernel<1, 24, false, true>) with thread
// t of a CTA = body t % 24 of env 4 * blockIdx.x + t / 24.  The per-body arithmetic is the same code (env_step_shared.cuh,
// phc_math.cuh), operation for operation; what changes is everything that used to be warp-scoped:
//   * staging: one shared-memory region, two mbarriers and one elected thread (body 0) per ENV; regions are 24 words (mod 32)
//     apart so that the stride-13 record reads of the two envs that share a warp fall into disjoint banks;
//   * the four tracking-error sums, the power sum and the termination vote cross warp boundaries: an 8-lane butterfly (segments
//     start at multiples of 8 lanes) leaves per-octet partials, which go through shared memory and are added in a fixed order by
//     the env's elected thread -- deterministic, but a different association than the 32-lane butterfly of env_step.cu
//     (the reward agrees to rounding, not bit for bit; everything per body IS bit-identical);
//   * three CTA barriers replace the __syncwarp()s: inputs staged -> phase A, phase A done (reductions published, the obs row may
//     overwrite the consumed blocks) -> phase B, rows staged -> bulk stores.
// Reference functions replaced: as env_step.cu (include/phc_b200.h, PhcStepArgs).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/phc_b200.h"
#include "phc_common.cuh"
#include "phc_math.cuh"
#include "env_step_shared.cuh"

namespace phc {
namespace packed {

constexpr int kJ = 24;                        // bodies per env = threads per env
constexpr int kEnvsPerCta = 4;
constexpr int kThreads = kJ * kEnvsPerCta;    // 96 = 3 full warps
constexpr int kCtasPerSm = 7;                 // 28 envs resident per SM, as in env_step.cu: 4096 envs = one wave on 148 SMs
constexpr int kBS = 312;                      // body_stride = round4(13 * 24)
constexpr int kD = 69;                        // 3 * (J - 1) dofs
constexpr int kSelfDim = 1 + 15 * kJ - 3;     // 358
constexpr int kObsDim = kSelfDim + 24 * kJ;   // 934
constexpr int kObsPad = 936;
// per-env shared-memory region (floats).  The observation row is staged over [cache | dof | state] once phase A has consumed them.
constexpr int kOffCache = 0;                  // cached reference pose of the reward time (one frame record)
constexpr int kOffDof = kBS;                  // (pos, vel) pairs, 2 * 69 -> 140 floats (inside the row's footprint)
constexpr int kOffState = 2 * kBS;            // simulator block
constexpr int kOffOslots = 3 * kBS;           // observation bracket: 2 frame slots; slot 0 becomes the pose-cache row
constexpr int kOffAmp = 5 * kBS;              // AMP vector (round4(amp_dim) floats), then 2 mbarriers (4 floats)
constexpr int kRedWords = 64;                 // CTA scratch behind the env regions: [4 envs][5 sums][3 octets] + 3 ballots

__host__ __device__ inline int env_stride(int amp_dim) {      // region size, padded to 24 (mod 32) words (bank layout above)
  const int need = kOffAmp + round4(amp_dim) + 4;
  return need + ((24 - (need & 31)) & 31);
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
env_step_packed_kernel(const __grid_constant__ PhcStepArgs a, const int amp_dim, const int stride) {
  extern __shared__ __align__(128) float smem[];
  const int t = threadIdx.x;
  const int e = t / kJ;
  const int j = t - e * kJ;
  const int lane = t & 31;
  const int env = blockIdx.x * kEnvsPerCta + e;            // the launcher guarantees num_envs % 4 == 0
  const bool lead = j == 0;                                // the env's elected thread
  float* const w_base = smem + e * stride;
  float* const s_cache = w_base + kOffCache;
  float* const s_dof = w_base + kOffDof;
  float* const s_state = w_base + kOffState;
  float* const s_oslots = w_base + kOffOslots;
  float* const s_amp = w_base + kOffAmp;
  float* const s_obs = w_base;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(s_amp + round4(amp_dim));   // state + cached reference pose
  uint64_t* const bar_o = bar + 1;                                              // observation bracket
  float* const s_red = smem + kEnvsPerCta * stride;
  uint32_t* const s_vote = reinterpret_cast<uint32_t*>(s_red + 60);

  constexpr uint32_t kBlockBytes = kBS * 4u;
  if (lead) {
    mbar_init(bar, 1);
    mbar_init(bar_o, 1);
    mbar_init_fence();
  }
  grid_dependency_wait();          // PDL: nothing above touches global memory (see env_step.cu)
  PHC_TL_IF(lead, env, 0);
#ifdef PHC_EXP_TIMELINE
  if (lead && g_timeline) g_timeline[(size_t)env * 8 + 7] = smid();
#endif

  // ---- every load that depends only on the env index, the few scalar requests first ------------------------------------------
  const int64_t progress = a.progress[env];
  const float t_start = a.start_times[env], t_off = a.start_offsets[env];
  const V3 goff = v3(a.global_offset[3 * env + 0], a.global_offset[3 * env + 1], a.global_offset[3 * env + 2]);
  const int4 em = *reinterpret_cast<const int4*>(a.env_motion + env);
  const float m_len = __int_as_float(em.x), m_dt = __int_as_float(em.y);
  const int m_nf = em.z;
  const int64_t m_start = em.w;
  if (lead) {
    mbar_arrive_expect_tx(bar, 2u * kBlockBytes);
    bulk_g2s(s_state, a.body_state + (size_t)env * a.bodies_per_env * kBodyRec, kBlockBytes, bar);
    bulk_g2s(s_cache, a.ref_cache + (size_t)env * kBS, kBlockBytes, bar);
  }
  const float2* g_dof = reinterpret_cast<const float2*>(a.dof_state) + (size_t)env * kD;
  const float* g_force = a.dof_force + (size_t)env * kD;
  float2 dof_pv[3];
  float dof_f[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int d = j + kJ * u;
    dof_pv[u] = (d < kD) ? g_dof[d] : make_float2(0.f, 0.f);
    dof_f[u] = (d < kD) ? g_force[d] : 0.f;
  }

  // ---- observation bracket at the NEXT motion time (humanoid_im.py:744-752); the reward pose of this step is the cached one ----
  const float tn = PHC_ADD(PHC_ADD(PHC_MUL((float)(progress + 1), a.dt), t_start), t_off);
  const Bracket32 bo = frame_bracket32(tn, m_len, m_nf, m_dt);
  const float bl_o = bo.blend;
  const bool two = bo.i1 != bo.i0;                         // the last frame of a clip brackets itself: one copy, both slots alias
  const float* const po0 = s_oslots;
  const float* const po1 = two ? s_oslots + kBS : s_oslots;
  if (lead) {
    mbar_arrive_expect_tx(bar_o, two ? 2u * kBlockBytes : kBlockBytes);
    bulk_g2s(s_oslots, a.lib.frames_body + (size_t)(m_start + bo.i0) * kBS, kBlockBytes, bar_o);
    if (two) bulk_g2s(s_oslots + kBS, a.lib.frames_body + (size_t)(m_start + bo.i1) * kBS, kBlockBytes, bar_o);
  }
  PHC_TL_IF(lead, env, 1);

  // ---- while the copies fly: dof state / force (power reward + AMP joint inputs) ---------------------------------------------
  float power = 0.0f;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int d = j + kJ * u;
    if (d < kD) {
      s_dof[2 * d] = dof_pv[u].x;
      s_dof[2 * d + 1] = dof_pv[u].y;
      power += fabsf(dof_f[u] * dof_pv[u].y);
    }
  }
  __syncthreads();               // barriers initialised + dof pairs staged, for every thread of the env (they span warps)
  mbar_wait(bar, 0);
  PHC_TL_IF(lead, env, 2);

  // ================= phase A: everything that reads the cached reference pose / simulator block ================================
  const BodyRec sim = load_body(s_state + j * kBodyRec);
  const V3 root_p = v3(s_state[0], s_state[1], s_state[2]);
  const Q4 root_q = q4(s_state[3], s_state[4], s_state[5], s_state[6]);
  const float heading = heading_angle(root_q);
  const Q4 hq = quat_about_z(heading);
  const Q4 hinv = q4(0.0f, 0.0f, -hq.z, hq.w);

  float dist;
  {
    const BodyRec ref = load_body(s_cache + j * kBodyRec);
    const V3 dp = ref.p - sim.p;
    const float sp = dp.x * dp.x + dp.y * dp.y + dp.z * dp.z;
    float e_pos = sp / 3.0f;
    const float ang = quat_angle(qmul(ref.q, qconj(sim.q)));
    float e_rot = ang * ang;
    const V3 dv = ref.v - sim.v, dw = ref.w - sim.w;
    float e_vel = (dv.x * dv.x + dv.y * dv.y + dv.z * dv.z) / 3.0f;
    float e_ang = (dw.x * dw.x + dw.y * dw.y + dw.z * dw.z) / 3.0f;
    dist = sqrtf(sp);
    // 8-lane butterfly: four values in 4 shuffles (halves exchange two values, quarters one), the power sum in 3
    const bool h4 = (lane & 4) != 0;
    float k0 = h4 ? e_vel : e_pos, k1 = h4 ? e_ang : e_rot;
    const float x0 = h4 ? e_pos : e_vel, x1 = h4 ? e_rot : e_ang;
    k0 += __shfl_xor_sync(0xffffffffu, x0, 4);
    k1 += __shfl_xor_sync(0xffffffffu, x1, 4);
    const bool h2 = (lane & 2) != 0;
    float k = h2 ? k1 : k0;
    const float x = h2 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, x, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);       // lanes 8o + {0,1}: e_pos of octet o; {2,3}: e_rot; {4,5}: e_vel; {6,7}: e_ang
    power += __shfl_xor_sync(0xffffffffu, power, 4);
    power += __shfl_xor_sync(0xffffffffu, power, 2);
    power += __shfl_xor_sync(0xffffffffu, power, 1);
    const int oct = j >> 3;                        // octet of this thread inside its env (segments start at multiples of 8 lanes)
    if ((lane & 1) == 0) s_red[(e * 5 + ((lane & 6) >> 1)) * 3 + oct] = k;
    else if ((lane & 7) == 1) s_red[(e * 5 + 4) * 3 + oct] = power;
  }
  {
    const uint32_t vote = __ballot_sync(0xffffffffu, dist > a.term_thresh[j]);
    if (lane == 0) s_vote[t >> 5] = vote;
  }

  // AMP observation of the simulated character (build_amp_observations_smpl) -> its own staging row
  {
    const int nj = a.num_amp_joints, nk = a.num_key_bodies;
    float* o = s_amp + 1;
    if (lead) s_amp[0] = root_p.z;
    if (j < nj) {                                  // nj <= 23: one joint per thread
      const int jid = a.amp_joints[j];
      const float* dj = s_dof + 6 * jid;
      st6(o + 12 + 6 * j, tan_norm(exp_map_to_quat(v3(dj[0], dj[2], dj[4]))));
      st3(o + 12 + 6 * nj + 3 * j, v3(dj[1], dj[3], dj[5]));
    }
    if (j < nk) {
      const float* kb = s_state + a.key_bodies[j] * kBodyRec;
      st3(o + 12 + 9 * nj + 3 * j, qrot_z(hinv, v3(kb[0], kb[1], kb[2]) - root_p));
    }
  }
  __syncthreads();               // cache / dof / state consumed by every thread of the env: the obs row may overwrite them;
  PHC_TL_IF(lead, env, 3);       // reduction partials and votes published

  // ================= phase B: observation row ===================================================================================
  if (lead) s_obs[0] = root_p.z;
  {
    float* o_pos = s_obs + 1;
    float* o_rot = o_pos + 3 * (kJ - 1);
    float* o_vel = o_rot + 6 * kJ;
    float* o_ang = o_vel + 3 * kJ;
    if (j > 0) st3(o_pos + 3 * (j - 1), qrot_z(hinv, sim.p - root_p));
    const TanNorm tnm = tan_norm(qmul_zl(hinv, sim.q));
    const V3 lv = qrot_z(hinv, sim.v), lw = qrot_z(hinv, sim.w);
    st6(o_rot + 6 * j, tnm);
    st3(o_vel + 3 * j, lv);
    st3(o_ang + 3 * j, lw);
    if (lead) {                                    // AMP root columns = the root's self-observation entries (upright start)
      float* o = s_amp + 1;
      st6(o, tnm); st3(o + 6, lv); st3(o + 9, lw);
    }
  }
  mbar_wait(bar_o, 0);
  PHC_TL_IF(lead, env, 4);
  {
    const BodyRec ref = blend_body(po0 + j * kBodyRec, po1 + j * kBodyRec, bl_o, goff);
    // thread j has consumed records j of both frames: slot 0 of the bracket becomes the row of the pose cache
    float* c = s_oslots + j * kBodyRec;
    st3(c, ref.p); c[3] = ref.q.x; c[4] = ref.q.y; c[5] = ref.q.z; c[6] = ref.q.w; st3(c + 7, ref.v); st3(c + 10, ref.w);
    float* tb = s_obs + kSelfDim;
    st3(tb + 3 * j, qrot_z(hinv, ref.p - sim.p));
    st6(tb + 3 * kJ + 6 * j, tan_norm(qmul_zr(qmul_zl(hinv, qmul(ref.q, qconj(sim.q))), hq)));
    st3(tb + 9 * kJ + 3 * j, qrot_z(hinv, ref.v - sim.v));
    st3(tb + 12 * kJ + 3 * j, qrot_z(hinv, ref.w - sim.w));
    st3(tb + 15 * kJ + 3 * j, qrot_z(hinv, ref.p - root_p));
    st6(tb + 18 * kJ + 6 * j, tan_norm(qmul_zl(hinv, ref.q)));
  }
  if (j < kObsPad - kObsDim) s_obs[kObsDim + j] = 0.f;      // the row's pad columns are written as zeros
  fence_async_smem();
  __syncthreads();
  PHC_TL_IF(lead, env, 5);
  if (!lead) return;

  // ---- the env's elected thread: rows leave as TMA bulk stores, then reward / reset while they drain --------------------------
  bulk_s2g(a.amp_out + (size_t)env * a.amp_out_stride + (a.ring_head ? (size_t)(*a.ring_head) * (size_t)amp_dim : (size_t)0), s_amp,
           (uint32_t)amp_dim * 4u);
  bulk_s2g(a.obs + (size_t)env * a.obs_stride, s_obs, (uint32_t)kObsPad * 4u);
  bulk_s2g(a.ref_cache + (size_t)env * kBS, s_oslots, kBlockBytes);
  bulk_commit();
  {
    const float* r = s_red + e * 15;
    const float r_pos = expf(-a.k_pos * (((r[0] + r[1]) + r[2]) / (float)kJ));
    const float r_rot = expf(-a.k_rot * (((r[3] + r[4]) + r[5]) / (float)kJ));
    const float r_vel = expf(-a.k_vel * (((r[6] + r[7]) + r[8]) / (float)kJ));
    const float r_ang = expf(-a.k_ang_vel * (((r[9] + r[10]) + r[11]) / (float)kJ));
    const float pw = (r[12] + r[13]) + r[14];
    float rew = a.w_pos * r_pos + a.w_rot * r_rot + a.w_vel * r_vel + a.w_ang_vel * r_ang;
    float* raw = a.reward_raw + (size_t)env * 5;
    raw[0] = r_pos; raw[1] = r_rot; raw[2] = r_vel; raw[3] = r_ang;
    float pr = -a.power_coef * pw;
    if (progress <= 3) pr = 0.0f;
    rew = rew + pr;
    raw[4] = pr;
    a.rew[env] = rew;
    // compute_humanoid_im_reset + the is_recovery override: the env's 24 vote bits out of the CTA's 96
    const uint32_t v0 = s_vote[0], v1 = s_vote[1], v2 = s_vote[2];
    const uint32_t mine = e == 0 ? (v0 & 0xffffffu) : (e == 1 ? ((v0 >> 24) | ((v1 & 0xffffu) << 8)) : (e == 2 ? ((v1 >> 16) | ((v2 & 0xffu) << 16)) : (v2 >> 8)));
    const bool fallen = mine != 0u;
    const float t_now = PHC_ADD(PHC_ADD(PHC_MUL((float)progress, a.dt), t_start), t_off);
    const bool pass_time = t_now >= m_len;
    int64_t terminated = (fallen && (progress > 1)) ? 1 : 0;
    int64_t reset = pass_time ? 1 : terminated;
    if (a.cycle_counter && !pass_time && a.cycle_counter[env] > 0) { reset = 0; terminated = 0; }
    a.reset[env] = reset;
    a.terminate[env] = terminated;
  }
  bulk_wait_read0();             // the shared-memory rows must outlive the bulk reads
  PHC_TL_IF(lead, env, 6);
}

}  // namespace packed
}  // namespace phc

// ------------------------------------------------------------------------------------------------------------
// launch (called by phc_env_step in env_step.cu once it has checked that the launch is the shipped steady state)
// ------------------------------------------------------------------------------------------------------------
extern "C" int phc_check_cuda(cudaError_t e, const char* what);
extern "C" void phc_count_launches(int n);

extern "C" int phc_env_step_packed_launch(const PhcStepArgs* a, int amp_dim, int pdl, void* stream) {
  using namespace phc::packed;
  const int stride = env_stride(amp_dim);
  const size_t smem = ((size_t)kEnvsPerCta * stride + kRedWords) * sizeof(float);
  static size_t smem_limit = 48 * 1024;
  if (smem > smem_limit) {
    cudaError_t e = cudaFuncSetAttribute(env_step_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return phc_check_cuda(e, "cudaFuncSetAttribute(env_step_packed_kernel)");
    smem_limit = smem;
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(a->num_envs / kEnvsPerCta));
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = la;
  lc.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&lc, env_step_packed_kernel, *a, amp_dim, stride);
  if (e != cudaSuccess) return phc_check_cuda(e, "cudaLaunchKernelEx(env_step_packed_kernel)");
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "env_step_packed_kernel launch");
}

#ifdef PHC_EXP_TIMELINE
extern "C" PHC_API int phc_exp_set_timeline_packed(void* buf) {
  unsigned long long* p = static_cast<unsigned long long*>(buf);
  return phc_check_cuda(cudaMemcpyToSymbol(phc::g_timeline, &p, sizeof(p)), "phc_exp_set_timeline_packed");
}
#endif
