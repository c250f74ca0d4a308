// TEST INFRASTRUCTURE -- a minimal host emulation of the CUDA constructs phc_b200/csrc/env_step.cu uses, so that the KERNEL
// SOURCE ITSELF (not a restatement) can be compiled with g++ and run on the CPU against the goldens of the unmodified
// reference (tests/test_env_step_emu_cpu.py).  One warp = 32 std::threads; warp collectives (__shfl*_sync, __any_sync,
// __syncwarp) are barrier + exchange array (the kernel only uses them in warp-uniform code with a full mask); mbarriers are
// two atomics (pending arrivals, pending bytes); TMA bulk copies are memcpy executed by the issuing lane.  What this checks:
// indexing, staging / aliasing of shared memory, the bracket / dedupe logic, every arithmetic expression (with the host's libm,
// -ffp-contract=off).  What it cannot check: alignment / async-proxy rules of the real TMA, timing, register pressure.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __align__(x)
#define __shared__

struct EmuDim3 { unsigned x, y, z; };
static thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;
static thread_local int emu_lane;

struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
struct int4 { int x, y, z, w; };
struct float4 { float x, y, z, w; };
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }

struct EmuWarp {
  std::barrier<> bar{32};
  uint64_t xch[32];
};
static thread_local EmuWarp* emu_warp = nullptr;     // the warp this (lane) thread belongs to
static std::barrier<>* emu_block_bar = nullptr;      // __syncthreads() of the block being emulated (block-level kernels only)
static inline void __syncthreads() { emu_block_bar->arrive_and_wait(); }

// PHC_EMU_CHAOS=1: every lane dawdles for a random few microseconds after each collective, so lanes drift as far apart as the
// program allows -- missing __syncwarp()s then show up as wrong results instead of hiding behind near-lock-step execution.
#include <chrono>
#include <cstdlib>
#include <random>
static inline void emu_chaos() {
  static const bool on = [] { const char* v = std::getenv("PHC_EMU_CHAOS"); return v && v[0] == '1'; }();
  if (!on) return;
  static thread_local std::minstd_rand rng{std::random_device{}()};
  const unsigned r = rng() % 8;
  if (r < 3) std::this_thread::sleep_for(std::chrono::microseconds(20 * (r + 1)));
  else if (r < 5) std::this_thread::yield();
}
static inline void __syncwarp() { emu_warp->bar.arrive_and_wait(); emu_chaos(); }
template <class T>
static inline T emu_exchange(T v, int src) {        // a lane outside 0..31 reads its own value (shfl_down / shfl_up at the edge)
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles");
  uint64_t u = 0;
  std::memcpy(&u, &v, sizeof(T));
  emu_warp->xch[emu_lane] = u;
  emu_warp->bar.arrive_and_wait();
  const uint64_t r = emu_warp->xch[(src < 0 || src > 31) ? emu_lane : src];
  emu_warp->bar.arrive_and_wait();
  emu_chaos();
  T o;
  std::memcpy(&o, &r, sizeof(T));
  return o;
}
static inline float __shfl_down_sync(unsigned, float v, int o) { return emu_exchange(v, emu_lane + o); }
static inline double __shfl_xor_sync(unsigned, double v, int o) { return emu_exchange(v, emu_lane ^ o); }
static inline float __shfl_xor_sync(unsigned, float v, int o) { return emu_exchange(v, emu_lane ^ o); }
static inline int __shfl_xor_sync(unsigned, int v, int o) { return emu_exchange(v, emu_lane ^ o); }
static inline float __shfl_sync(unsigned, float v, int src) { return emu_exchange(v, src); }
static inline int __shfl_sync(unsigned, int v, int src) { return emu_exchange(v, src); }
static inline bool __any_sync(unsigned, bool p) {
  emu_warp->xch[emu_lane] = p ? 1u : 0u;
  emu_warp->bar.arrive_and_wait();
  uint64_t any = 0;
  for (int i = 0; i < 32; ++i) any |= emu_warp->xch[i];
  emu_warp->bar.arrive_and_wait();
  return any != 0;
}

static inline unsigned __ballot_sync(unsigned, bool p) {
  emu_warp->xch[emu_lane] = p ? 1u : 0u;
  emu_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(emu_warp->xch[i] & 1u) << i;
  emu_warp->bar.arrive_and_wait();
  emu_chaos();
  return m;
}

// atomicAdd: one global lock (the emulation is about values, not contention)
static std::mutex emu_atomic_mutex;
template <class T>
static inline T atomicAdd(T* p, T v) { std::lock_guard<std::mutex> g(emu_atomic_mutex); const T o = *p; *p = o + v; return o; }

// round-to-nearest intrinsics: plain operations under -ffp-contract=off
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
