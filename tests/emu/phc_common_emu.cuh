// TEST INFRASTRUCTURE -- host stand-ins for phc_b200/csrc/phc_common.cuh (same names, same contracts) used when the kernel
// source is compiled for the CPU emulation: mbarrier = {pending arrivals, pending bytes}, bulk copies = memcpy by the issuer.
#pragma once
#include "cuda_emu_prelude.h"

namespace phc {

struct EmuMBar { std::atomic<int32_t> arrivals; std::atomic<int32_t> tx; };
static_assert(sizeof(EmuMBar) == 8, "an mbarrier is one 8-byte shared-memory word");
static inline EmuMBar* emu_bar(uint64_t* b) { return reinterpret_cast<EmuMBar*>(b); }

static inline void mbar_init(uint64_t* bar, uint32_t count) { emu_bar(bar)->arrivals.store((int32_t)count); emu_bar(bar)->tx.store(0); }
static inline void mbar_init_fence() {}
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { emu_bar(bar)->tx.fetch_add((int32_t)bytes); emu_bar(bar)->arrivals.fetch_sub(1); }
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_bar(bar)->tx.fetch_add((int32_t)bytes); }
static inline void mbar_arrive(uint64_t* bar) { emu_bar(bar)->arrivals.fetch_sub(1); }
static inline bool mbar_try_wait(uint64_t* bar, uint32_t) { return emu_bar(bar)->arrivals.load() <= 0 && emu_bar(bar)->tx.load() <= 0; }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) std::this_thread::yield(); emu_chaos(); }
static inline void fence_async_smem() {}
static inline void grid_dependency_wait() {}      // programmatic dependent launch: nothing precedes the emulated launch
static inline void bulk_s2g(void* dst, const void* src, uint32_t bytes) { std::memcpy(dst, src, bytes); }
static inline void bulk_commit() {}
static inline void bulk_wait_read0() {}
// the real copy needs 16-byte aligned addresses and a multiple of 16 bytes: keep that contract visible in the emulation
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15)) std::abort();
  std::memcpy(dst, src, bytes);
  emu_bar(bar)->tx.fetch_sub((int32_t)bytes);
}

}  // namespace phc
