"""TEST INFRASTRUCTURE: assemble and compile the CPU emulation of the fused env-step kernel FROM THE PRODUCT SOURCES.

The translation unit is: the emulation prelude (tests/emu/cuda_emu_prelude.h, phc_common_emu.cuh) + `warp_sum` / `warp_sum4`
cut verbatim out of phc_b200/csrc/phc_common.cuh + include/phc_b200.h + phc_b200/csrc/phc_math.cuh + the `namespace phc { ... }`
part of phc_b200/csrc/env_step.cu (layout helpers and `env_step_kernel`, verbatim) + a launcher that runs one warp (32
std::threads) per env.  Nothing of the kernel is restated here."""
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "phc_b200", "csrc")

LAUNCHER = r'''
namespace phc { alignas(128) float smem[1 << 16]; }      // the kernel's `extern __shared__ float smem[]`

template <int T_MAX, int JT, bool GETUP, bool FAST>
static void emu_launch(const PhcStepArgs& a, int obs_dim, int self_dim, int amp_dim, bool alias_obs, bool state_bulk_ok) {
  for (int env = 0; env < a.num_envs; ++env) {
    EmuWarp warp;
    emu_warp = &warp;
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; ++lane)
      lanes.emplace_back([&, lane] {
        emu_lane = lane;
        threadIdx.x = (unsigned)((env % phc::kWarpsPerCta) * 32 + lane); threadIdx.y = threadIdx.z = 0;
        blockIdx.x = (unsigned)(env / phc::kWarpsPerCta); blockIdx.y = blockIdx.z = 0;
        phc::env_step_kernel<T_MAX, JT, GETUP, FAST>(a, obs_dim, self_dim, amp_dim, alias_obs, state_bulk_ok);
      });
    for (auto& t : lanes) t.join();
  }
  emu_warp = nullptr;
}

extern "C" int emu_env_step(const PhcStepArgs* a, int obs_dim, int self_dim, int amp_dim, int alias_obs, int state_bulk_ok, int variant) {
  switch (variant) {
    case 0: emu_launch<1, 24, false, false>(*a, obs_dim, self_dim, amp_dim, alias_obs, state_bulk_ok); return 0;
    case 1: emu_launch<1, 24, false, true>(*a, obs_dim, self_dim, amp_dim, alias_obs, state_bulk_ok); return 0;
    case 2: emu_launch<1, 24, true, false>(*a, obs_dim, self_dim, amp_dim, alias_obs, state_bulk_ok); return 0;
    case 3: emu_launch<1, 0, false, false>(*a, obs_dim, self_dim, amp_dim, alias_obs, state_bulk_ok); return 0;
    case 4: emu_launch<4, 0, false, false>(*a, obs_dim, self_dim, amp_dim, alias_obs, state_bulk_ok); return 0;
    case 5: emu_launch<1, 0, true, false>(*a, obs_dim, self_dim, amp_dim, alias_obs, state_bulk_ok); return 0;
  }
  return -1;
}
extern "C" uint32_t emu_fast_flags(void) { return phc::kFastFlags; }
'''


def assemble() -> str:
    common = open(os.path.join(CSRC, "phc_common.cuh")).read()
    a = common.index("__device__ __forceinline__ float warp_sum(float v) {")
    b = common.index("}  // namespace phc")
    reductions = common[a:b]
    step = open(os.path.join(CSRC, "env_step.cu")).read()
    k0 = step.index("namespace phc {")
    k1 = step.index("}  // namespace phc") + len("}  // namespace phc")
    kernel = step[k0:k1]
    assert "env_step_kernel(" in kernel and "<<<" not in kernel
    return "\n".join([
        '#include "cuda_emu_prelude.h"', '#include "phc_common_emu.cuh"',
        "namespace phc {", reductions, "}",
        f'#include "{os.path.join(ROOT, "include", "phc_b200.h")}"', f'#include "{os.path.join(CSRC, "phc_math.cuh")}"',
        kernel, LAUNCHER])


def build(out_dir: str) -> str:
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not available")
    src = os.path.join(out_dir, "env_step_emu.cpp")
    with open(src, "w") as f:
        f.write(assemble())
    so = os.path.join(out_dir, "libenv_step_emu.so")
    r = subprocess.run([gxx, "-O1", "-std=c++20", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + HERE, src, "-o", so, "-lm"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stderr[:6000])
    return so


if __name__ == "__main__":
    import tempfile
    d = tempfile.mkdtemp()
    print(build(d))
