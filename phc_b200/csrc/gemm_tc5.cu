// Blackwell-native GEMM for the MLPs: tcgen05.mma (kind::tf32) issued by one thread, operands staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle), accumulator in TMEM, epilogue warps read it back with tcgen05.ld.
// Same contract as phc_gemm (gemm.cu) -- C[M,N] (+)= epi(alpha * sum_k A(m,k) B(n,k)), both operand major-nesses --
// and the same fp32-equivalent numerics: 3xTF32, but with the split done ONCE per operand in global memory
// (phc_split_tf32: x -> hi = rna_tf32(x), lo = rna_tf32(x - hi)) because UMMA reads its operands straight from shared
// memory; per k-step the issuing thread launches D += A_lo*B_hi, D += A_hi*B_lo, D += A_hi*B_hi (small terms first).
//
// Two tile configurations share the code (template parameter CTAS):
//   CTAS = 1: 128 x 128 tile per CTA (small / skinny problems);
//   CTAS = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) computes a 256 x 256 tile: each CTA stages its own 128 rows
//             of A and its own 128 of the 256 B rows (same 64 KB per stage), the leader CTA issues M = 256, N = 256 MMAs
//             that read both CTAs' shared memory, each CTA keeps its 128 x 256 accumulator half in its own TMEM.
//             Twice the flops per byte fetched from L2 -- the 128 x 128 tile is L2-bandwidth bound at ~50 % of the MMA rate.
// CTA = 192 threads: warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA producer, warp 5 TMEM owner + MMA
// issuer.  Tile 128 x 128, BLOCK_K = 32 floats (one 128-byte swizzle atom), 3 pipeline stages x 4 operand tiles x 16 KB.
// Shared-memory operand layouts (what the descriptors encode, cf. cute/atom/mma_traits_sm100.hpp make_umma_desc):
//   k-contiguous operand : tile [128 rows][32 floats], SW128; atoms of 8 rows x 128 B, SBO = 1024 B; one UMMA
//                          (K = 8 floats = 32 B) advances the start address by 32 B inside the atom;
//   mn-contiguous operand: for 32-bit (tf32) data the only UMMA layout is SWIZZLE_128B_BASE32B (32-byte swizzle
//                          granularity, cute Layout_MN_SW128_32B_Atom = Swizzle<2,5,2>, atoms of 4 k-rows x 128 B), the
//                          TMA counterpart is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  4 TMA boxes [32 k-rows][32 floats],
//                          each 4096 B: LBO = 4096 B between MN atoms, SBO = 512 B between groups of 4 k-rows; one UMMA
//                          (K = 8 = two groups) advances the start address by 1024 B.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/phc_b200.h"
#include "phc_common.cuh"
#include "tc5_common.cuh"

extern "C" void phc_set_error(const char* msg);
extern "C" int phc_check_cuda(cudaError_t e, const char* what);
extern "C" void phc_count_launches(int n);

namespace phc {
namespace tc5 {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;               // 16 KB per operand tile (BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct Args {
  float* C; float* C_hi; float* C_lo;       // optional pre-split copies of the result for the next GEMM (same ldc)
  const float* bias; float* mask;        // mask = `aux` of the C ABI (read for the backward modes, written by SiLU forward)
  int M, N, K;
  int64_t ldc, ldmask;
  float alpha;
  int relu, accumulate, k_splits;
};

// epilogue of one 32-column chunk of row m held in r[] (alpha, bias, ReLU, ReLU-backward mask, store / atomic add, optional
// pre-split copies).  n_base = first column of the chunk.
__device__ __forceinline__ void epilogue_store(const Args& g, const uint32_t (&r)[32], int m, int n_base, bool add_bias) {
  const int n0 = n_base, c0 = 0;
        float* crow = g.C + (int64_t)m * g.ldc;
        float* mrow = g.mask ? g.mask + (int64_t)m * g.ldmask : nullptr;
        const int act = g.relu;                               // PHC_ACT_*
        // 16-byte vector path when the row segment is aligned and fully inside the matrix (always for interior tiles)
        const bool vec = !g.accumulate && ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                         (!mrow || (((g.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0)));
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + c0 + j;
          if (n >= g.N) break;
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = g.alpha * __uint_as_float(r[j + e]);
            if (g.bias && add_bias && n + e < g.N) x += g.bias[n + e];
            if (act == PHC_ACT_RELU) x = fmaxf(x, 0.f);
            v[e] = x;
          }
          if (vec && n + 3 < g.N) {
            if (act == PHC_ACT_SILU) {
              if (mrow) *reinterpret_cast<float4*>(mrow + n) = make_float4(v[0], v[1], v[2], v[3]);     // pre-activation
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = silu_f(v[e]);
            } else if (mrow) {
              const float4 mk = *reinterpret_cast<const float4*>(mrow + n);
              if (act == PHC_ACT_SILU_BWD) {
                v[0] *= silu_grad_f(mk.x); v[1] *= silu_grad_f(mk.y); v[2] *= silu_grad_f(mk.z); v[3] *= silu_grad_f(mk.w);
              } else {
                v[0] = mk.x > 0.f ? v[0] : 0.f; v[1] = mk.y > 0.f ? v[1] : 0.f;
                v[2] = mk.z > 0.f ? v[2] : 0.f; v[3] = mk.w > 0.f ? v[3] : 0.f;
              }
            }
            *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
            if (g.C_hi) {
              float hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                uint32_t h, l;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v[e]));
                const float res = v[e] - __uint_as_float(h);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(res));
                hi[e] = __uint_as_float(h); lo[e] = __uint_as_float(l);
              }
              *reinterpret_cast<float4*>(g.C_hi + (int64_t)m * g.ldc + n) = make_float4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<float4*>(g.C_lo + (int64_t)m * g.ldc + n) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (n + e >= g.N) break;
              float x = v[e];
              if (act == PHC_ACT_SILU) {
                if (mrow) mrow[n + e] = x;
                x = silu_f(x);
              } else if (mrow) {
                x = (act == PHC_ACT_SILU_BWD) ? x * silu_grad_f(mrow[n + e]) : ((mrow[n + e] > 0.f) ? x : 0.f);
              }
              if (g.accumulate) atomicAdd(crow + n + e, x);
              else {
                crow[n + e] = x;
                if (g.C_hi) {
                  uint32_t h, l;
                  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
                  const float res = x - __uint_as_float(h);
                  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(res));
                  g.C_hi[(int64_t)m * g.ldc + n + e] = __uint_as_float(h);
                  g.C_lo[(int64_t)m * g.ldc + n + e] = __uint_as_float(l);
                }
              }
            }
          }
        }
}

template <bool A_K, bool B_K, int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const __grid_constant__ Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TN = BN * CTAS;                            // accumulator columns per CTA (= MMA N)
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // CTAS == 2: blockIdx.x enumerates the CTAs of the pairs along M; both CTAs of a pair share the 256-column N range
  const int m0 = (CTAS == 2) ? ((int)(blockIdx.x >> 1) * 2 * BM + (int)rank * BM) : (int)blockIdx.y * BM;
  const int n0 = (CTAS == 2) ? (int)blockIdx.y * TN : (int)blockIdx.x * BN;
  const int nb0 = n0 + (int)rank * BN;                     // first B row this CTA stages
  const int kb_total = (g.K + BK - 1) / BK;
  const int kb_per = (kb_total + g.k_splits - 1) / g.k_splits;
  const int kb_begin = blockIdx.z * kb_per;
  const int kb_end = min(kb_total, kb_begin + kb_per);
  const int nkb = kb_end - kb_begin;
  if (nkb <= 0) return;                                   // uniform per CTA

  if (threadIdx.x == 0) {
    // full barrier: the pair's two producers arrive (the leader's arrival carries the byte count of both CTAs)
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, CTAS); mbar_init(empty_bar + s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) { if (CTAS == 2) tmem_alloc_2cta(tmem_slot, TN); else tmem_alloc(tmem_slot, TN); }   // whole warp, .sync.aligned
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();                       // peer barriers initialised, both TMEM halves allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(empty_bar + s, ((i / STAGES) - 1) & 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        const int k0 = (kb_begin + i) * BK;
        if (leader) mbar_expect_tx(full_bar + s, CTAS * STAGE_BYTES);
        else mbar_arrive_remote_leader(full_bar + s);
        auto ld = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
          if (CTAS == 2) tma_load_2d_2cta(dst, tm, full_bar + s, c0, c1);
          else tma_load_2d(dst, tm, full_bar + s, c0, c1);
        };
        if (A_K) {
          ld(st, &tmAh, k0, m0);
          ld(st + TILE_BYTES, &tmAl, k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) {
            ld(st + j * 4096, &tmAh, m0 + 32 * j, k0);
            ld(st + TILE_BYTES + j * 4096, &tmAl, m0 + 32 * j, k0);
          }
        }
        if (B_K) {
          ld(st + 2 * TILE_BYTES, &tmBh, k0, nb0);
          ld(st + 3 * TILE_BYTES, &tmBl, k0, nb0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) {
            ld(st + 2 * TILE_BYTES + j * 4096, &tmBh, nb0 + 32 * j, k0);
            ld(st + 3 * TILE_BYTES + j * 4096, &tmBl, nb0 + 32 * j, k0);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = instr_desc(!A_K, !B_K, BM * CTAS, TN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbar_wait(full_bar + s, (i / STAGES) & 1);
        tc_fence_after();
        const uint32_t st = s32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint32_t a_off = A_K ? kk * 32 : kk * 1024;
          const uint32_t b_off = B_K ? kk * 32 : kk * 1024;
          const uint32_t a_lbo = A_K ? 16 : 4096, b_lbo = B_K ? 16 : 4096;
          const uint32_t a_sbo = A_K ? 1024 : 512, b_sbo = B_K ? 1024 : 512;
          const uint32_t a_lt = A_K ? 2 : 1, b_lt = B_K ? 2 : 1;
          const uint64_t dAh = smem_desc(st + a_off, a_lbo, a_sbo, a_lt);
          const uint64_t dAl = smem_desc(st + TILE_BYTES + a_off, a_lbo, a_sbo, a_lt);
          const uint64_t dBh = smem_desc(st + 2 * TILE_BYTES + b_off, b_lbo, b_sbo, b_lt);
          const uint64_t dBl = smem_desc(st + 3 * TILE_BYTES + b_off, b_lbo, b_sbo, b_lt);
          if (CTAS == 2) {
            umma_tf32_2cta(tmem_base, dAl, dBh, idesc, (i > 0 || kk > 0) ? 1u : 0u);
            umma_tf32_2cta(tmem_base, dAh, dBl, idesc, 1u);
            umma_tf32_2cta(tmem_base, dAh, dBh, idesc, 1u);
          } else {
            umma_tf32(tmem_base, dAl, dBh, idesc, (i > 0 || kk > 0) ? 1u : 0u);
            umma_tf32(tmem_base, dAh, dBl, idesc, 1u);
            umma_tf32(tmem_base, dAh, dBh, idesc, 1u);
          }
        }
        if (CTAS == 2) umma_commit_2cta(empty_bar + s); else umma_commit(empty_bar + s);   // frees the stage (in both CTAs)
      }
      if (CTAS == 2) umma_commit_2cta(tmem_full); else umma_commit(tmem_full);             // accumulator complete
    }
  } else {
    // ===================== epilogue warps 0..3: TMEM -> registers -> global =====================
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int m = m0 + row;
#pragma unroll 1
    for (int c0 = 0; c0 < TN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      if (m < g.M) epilogue_store(g, r, m, n0 + c0, blockIdx.z == 0);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();                       // the peer may still be reading / the leader still issuing
  if (warp == 5) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_2cta(tmem_base, TN); else tmem_dealloc(tmem_base, TN);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant (single-CTA 128 x 128 tiles): one CTA per SM walks a static list of tiles.  The shared-memory operand
// pipeline runs continuously across tiles and the accumulator is double-buffered in TMEM (2 x 128 columns), so the
// epilogue of tile i (TMEM -> registers -> global) overlaps the MMAs of tile i+1:
//   TMA warp  : for every (tile, k-block): wait empty[s] -> arm full[s] -> bulk-tensor loads
//   MMA thread: for every tile: wait tmem_empty[acc] -> for every k-block: wait full[s] -> 12 UMMAs -> commit empty[s];
//               commit tmem_full[acc]
//   epilogue  : wait tmem_full[acc] -> tcgen05.ld / epilogue math / stores -> arrive tmem_empty[acc] (one lane per warp)
// Tiles are ordered n fastest (see decode below).
// ------------------------------------------------------------------------------------------------------------------

template <bool A_K, bool B_K>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc5_persist_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                        const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                        const __grid_constant__ Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  const int kb_total = (g.K + BK - 1) / BK;
  const int kb_per = (kb_total + g.k_splits - 1) / g.k_splits;
  const int num_tiles = tiles_m * tiles_n * g.k_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (z, m, n): n fastest.  The CTAs that run at the same time then share a few A row blocks (read once from HBM,
  // re-used out of L2 by the other n tiles right away) and sweep the B operand, which for the MLP shapes is the weight
  // matrix and stays L2 resident.  (m fastest made every concurrent tile stream a different A block and re-read the
  // whole activation matrix from HBM once per n tile: 372 vs 433 TFLOP/s against the one-tile-per-CTA launch.)
  auto decode = [&](int t, int& m0, int& n0, int& kb_begin, int& nkb, int& z) {
    const int ni = t % tiles_n;
    const int r = t / tiles_n;
    const int mi = r % tiles_m;
    z = r / tiles_m;
    m0 = mi * BM; n0 = ni * BN;
    kb_begin = z * kb_per;
    const int kb_end = min(kb_total, kb_begin + kb_per);
    nkb = max(0, kb_end - kb_begin);
  };

  if (warp == 4) {
    if (lane == 0) {
      uint32_t it = 0;                                        // global k-block counter (continues across tiles)
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m0, n0, kb_begin, nkb, z;
        decode(t, m0, n0, kb_begin, nkb, z);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(empty_bar + s, ((it / STAGES) - 1) & 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int k0 = (kb_begin + i) * BK;
          mbar_expect_tx(full_bar + s, STAGE_BYTES);
          if (A_K) {
            tma_load_2d(st, &tmAh, full_bar + s, k0, m0);
            tma_load_2d(st + TILE_BYTES, &tmAl, full_bar + s, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) {
              tma_load_2d(st + j * 4096, &tmAh, full_bar + s, m0 + 32 * j, k0);
              tma_load_2d(st + TILE_BYTES + j * 4096, &tmAl, full_bar + s, m0 + 32 * j, k0);
            }
          }
          if (B_K) {
            tma_load_2d(st + 2 * TILE_BYTES, &tmBh, full_bar + s, k0, n0);
            tma_load_2d(st + 3 * TILE_BYTES, &tmBl, full_bar + s, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
              tma_load_2d(st + 2 * TILE_BYTES + j * 4096, &tmBh, full_bar + s, n0 + 32 * j, k0);
              tma_load_2d(st + 3 * TILE_BYTES + j * 4096, &tmBl, full_bar + s, n0 + 32 * j, k0);
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc(!A_K, !B_K, BM, BN);
      uint32_t it = 0, lt = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
        int m0, n0, kb_begin, nkb, z;
        decode(t, m0, n0, kb_begin, nkb, z);
        const uint32_t acc = lt & 1, use = lt >> 1;
        if (use > 0) mbar_wait(tmem_empty + acc, (use - 1) & 1);    // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(full_bar + s, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t st = s32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t a_off = A_K ? kk * 32 : kk * 1024;
            const uint32_t b_off = B_K ? kk * 32 : kk * 1024;
            const uint32_t a_lbo = A_K ? 16 : 4096, b_lbo = B_K ? 16 : 4096;
            const uint32_t a_sbo = A_K ? 1024 : 512, b_sbo = B_K ? 1024 : 512;
            const uint32_t a_lt = A_K ? 2 : 1, b_lt = B_K ? 2 : 1;
            const uint64_t dAh = smem_desc(st + a_off, a_lbo, a_sbo, a_lt);
            const uint64_t dAl = smem_desc(st + TILE_BYTES + a_off, a_lbo, a_sbo, a_lt);
            const uint64_t dBh = smem_desc(st + 2 * TILE_BYTES + b_off, b_lbo, b_sbo, b_lt);
            const uint64_t dBl = smem_desc(st + 3 * TILE_BYTES + b_off, b_lbo, b_sbo, b_lt);
            umma_tf32(tmem_d, dAl, dBh, idesc, (i > 0 || kk > 0) ? 1u : 0u);
            umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
            umma_tf32(tmem_d, dAh, dBh, idesc, 1u);
          }
          umma_commit(empty_bar + s);
        }
        umma_commit(tmem_full + acc);        // also fires (immediately) for an empty k-range: the epilogue then sees nkb == 0
      }
    }
  } else {
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      int m0, n0, kb_begin, nkb, z;
      decode(t, m0, n0, kb_begin, nkb, z);
      const uint32_t acc = lt & 1, use = lt >> 1;
      mbar_wait(tmem_full + acc, use & 1);
      tc_fence_after();
      const int m = m0 + warp * 32 + lane;
      if (nkb > 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + acc * BN + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
          if (m < g.M) epilogue_store(g, r, m, n0 + c0, z == 0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
    }
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent CTA-pair variant: a cluster of two CTAs walks a static list of 256 x 256 tiles with cta_group::2 MMAs
// (M = 256 split over the two SMs, N = 256).  Each CTA stages only ITS 128 rows of A and ITS 128 of the 256 B rows per
// k-block -- half the L2 -> shared-memory traffic per flop of the 128 x 128 kernel, which is what bounds 3xTF32 (four
// operand tiles feed three MMAs).  Accumulators are double-buffered in TMEM (2 x 256 of the 512 columns), so the
// epilogue of tile i overlaps the MMAs of tile i+1 as in the single-CTA persistent kernel.
//   TMA warp (both CTAs): wait own empty[s] -> leader arms full[s] with the bytes of BOTH CTAs, the peer arrives remotely
//                         -> bulk-tensor loads that credit the leader's barrier
//   MMA thread (leader) : wait tmem_empty[acc] (8 arrivals: 4 epilogue warps x 2 CTAs) -> per k-block: wait full[s] ->
//                         12 UMMAs -> multicast commit to empty[s] of both CTAs; multicast commit to tmem_full[acc]
//   epilogue (both CTAs): wait own tmem_full[acc] -> tcgen05.ld / math / stores of its 128 rows -> arrive on the leader's
//                         tmem_empty[acc]
// ------------------------------------------------------------------------------------------------------------------
template <bool A_K, bool B_K>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc5_persist_pair_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                             const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                             const __grid_constant__ Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2] (the leader's copies are the ones waited on)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  constexpr int PM = 2 * BM, PN = 2 * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_m = (g.M + PM - 1) / PM, tiles_n = (g.N + PN - 1) / PN;
  const int kb_total = (g.K + BK - 1) / BK;
  const int kb_per = (kb_total + g.k_splits - 1) / g.k_splits;
  const int num_tiles = tiles_m * tiles_n * g.k_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 2); mbar_init(empty_bar + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) tmem_alloc_2cta(tmem_slot, 2 * PN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (z, m, n): n fastest (see the single-CTA kernel).  m0 = first row of THIS CTA, n0 = first column of the pair
  auto decode = [&](int t, int& m0, int& n0, int& kb_begin, int& nkb, int& z) {
    const int ni = t % tiles_n;
    const int r = t / tiles_n;
    const int mi = r % tiles_m;
    z = r / tiles_m;
    m0 = mi * PM + (int)rank * BM; n0 = ni * PN;
    kb_begin = z * kb_per;
    const int kb_end = min(kb_total, kb_begin + kb_per);
    nkb = max(0, kb_end - kb_begin);
  };

  if (warp == 4) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        int m0, n0, kb_begin, nkb, z;
        decode(t, m0, n0, kb_begin, nkb, z);
        const int nb0 = n0 + (int)rank * BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(empty_bar + s, ((it / STAGES) - 1) & 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int k0 = (kb_begin + i) * BK;
          if (leader) mbar_expect_tx(full_bar + s, 2 * STAGE_BYTES);
          else mbar_arrive_remote_leader(full_bar + s);
          if (A_K) {
            tma_load_2d_2cta(st, &tmAh, full_bar + s, k0, m0);
            tma_load_2d_2cta(st + TILE_BYTES, &tmAl, full_bar + s, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) {
              tma_load_2d_2cta(st + j * 4096, &tmAh, full_bar + s, m0 + 32 * j, k0);
              tma_load_2d_2cta(st + TILE_BYTES + j * 4096, &tmAl, full_bar + s, m0 + 32 * j, k0);
            }
          }
          if (B_K) {
            tma_load_2d_2cta(st + 2 * TILE_BYTES, &tmBh, full_bar + s, k0, nb0);
            tma_load_2d_2cta(st + 3 * TILE_BYTES, &tmBl, full_bar + s, k0, nb0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
              tma_load_2d_2cta(st + 2 * TILE_BYTES + j * 4096, &tmBh, full_bar + s, nb0 + 32 * j, k0);
              tma_load_2d_2cta(st + 3 * TILE_BYTES + j * 4096, &tmBl, full_bar + s, nb0 + 32 * j, k0);
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = instr_desc(!A_K, !B_K, PM, PN);
      uint32_t it = 0, lt = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs, ++lt) {
        int m0, n0, kb_begin, nkb, z;
        decode(t, m0, n0, kb_begin, nkb, z);
        const uint32_t acc = lt & 1, use = lt >> 1;
        if (use > 0) mbar_wait(tmem_empty + acc, (use - 1) & 1);    // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * PN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(full_bar + s, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t st = s32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t a_off = A_K ? kk * 32 : kk * 1024;
            const uint32_t b_off = B_K ? kk * 32 : kk * 1024;
            const uint32_t a_lbo = A_K ? 16 : 4096, b_lbo = B_K ? 16 : 4096;
            const uint32_t a_sbo = A_K ? 1024 : 512, b_sbo = B_K ? 1024 : 512;
            const uint32_t a_lt = A_K ? 2 : 1, b_lt = B_K ? 2 : 1;
            const uint64_t dAh = smem_desc(st + a_off, a_lbo, a_sbo, a_lt);
            const uint64_t dAl = smem_desc(st + TILE_BYTES + a_off, a_lbo, a_sbo, a_lt);
            const uint64_t dBh = smem_desc(st + 2 * TILE_BYTES + b_off, b_lbo, b_sbo, b_lt);
            const uint64_t dBl = smem_desc(st + 3 * TILE_BYTES + b_off, b_lbo, b_sbo, b_lt);
            umma_tf32_2cta(tmem_d, dAl, dBh, idesc, (i > 0 || kk > 0) ? 1u : 0u);
            umma_tf32_2cta(tmem_d, dAh, dBl, idesc, 1u);
            umma_tf32_2cta(tmem_d, dAh, dBh, idesc, 1u);
          }
          umma_commit_2cta(empty_bar + s);       // frees the stage in both CTAs
        }
        umma_commit_2cta(tmem_full + acc);
      }
    }
  } else {
    uint32_t lt = 0;
    for (int t = pair_id; t < num_tiles; t += num_pairs, ++lt) {
      int m0, n0, kb_begin, nkb, z;
      decode(t, m0, n0, kb_begin, nkb, z);
      const uint32_t acc = lt & 1, use = lt >> 1;
      mbar_wait(tmem_full + acc, use & 1);
      tc_fence_after();
      const int m = m0 + warp * 32 + lane;
      if (nkb > 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < PN; c0 += 32) {
          if (n0 + c0 >= g.N) break;
          uint32_t r[32];
          tmem_ld32(tmem_base + acc * PN + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
          if (m < g.M) epilogue_store(g, r, m, n0 + c0, z == 0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(tmem_empty + acc); else mbar_arrive_remote_leader(tmem_empty + acc); }
    }
  }
  __syncthreads();
  cluster_sync_all();                            // the peer may still be reading / the leader still issuing
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 2 * PN);
  }
}

// x -> hi = rna_tf32(x), lo = rna_tf32(x - hi) over a strided [rows, cols] block
__global__ void split_tf32_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t ldo) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = x[r * ldx + c];
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float res = v - __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(res));
    hi[r * ldo + c] = __uint_as_float(h);
    lo[r * ldo + c] = __uint_as_float(l);
  }
}

// 2-D fp32 tensor map; inner dimension = the contiguous one.  kmajor: dims {K, rows}, box {32, 128};
// otherwise dims {rows(MN), K}, box {32, 32}.
static bool make_map(CUtensorMap* tm, const float* base, int64_t ld, int rows_mn, int K, bool kmajor) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (kmajor) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows_mn; box[0] = 32; box[1] = 128; }
  else { dims[0] = (cuuint64_t)rows_mn; dims[1] = (cuuint64_t)K; box[0] = 32; box[1] = 32; }
  strides[0] = (cuuint64_t)ld * 4;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tc5
}  // namespace phc

extern "C" int phc_split_tf32(const float* x, int64_t ldx, int64_t rows, int32_t cols, float* hi, float* lo, int64_t ldo,
                              void* stream) {
  if (!x || !hi || !lo || rows < 0 || cols < 1 || ldx < cols || ldo < cols) { phc_set_error("phc_split_tf32: bad arguments"); return PHC_ERR_INVALID_ARG; }
  if (rows == 0) return PHC_OK;
  int64_t g = (rows * cols + 255) / 256; if (g > 148 * 8) g = 148 * 8;
  phc::tc5::split_tf32_kernel<<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, rows, cols, hi, lo, ldo); phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "split_tf32_kernel");
}

extern "C" int phc_gemm_tc5(const float* A_hi, const float* A_lo, int64_t lda, int32_t a_kmajor, const float* B_hi,
                            const float* B_lo, int64_t ldb, int32_t b_kmajor, float* C, float* C_hi, float* C_lo, int64_t ldc,
                            int32_t M, int32_t N, int32_t K, float alpha, const float* bias, int32_t relu, float* mask,
                            int64_t ldmask, int32_t accumulate, int32_t k_splits, void* stream) {
  using namespace phc::tc5;
  if (!A_hi || !A_lo || !B_hi || !B_lo || !C || M < 0 || N < 0 || K < 1) { phc_set_error("phc_gemm_tc5: bad arguments"); return PHC_ERR_INVALID_ARG; }
  if (M == 0 || N == 0) return PHC_OK;
  if ((lda & 3) || (ldb & 3)) { phc_set_error("phc_gemm_tc5: leading dimensions must be multiples of 4 floats (TMA strides are 16-byte multiples)"); return PHC_ERR_INVALID_ARG; }
  for (const float* p : {A_hi, A_lo, B_hi, B_lo})
    if (reinterpret_cast<uintptr_t>(p) & 15) { phc_set_error("phc_gemm_tc5: operands must be 16-byte aligned"); return PHC_ERR_INVALID_ARG; }
  if (k_splits < 1) k_splits = 1;
  if (relu < 0 || relu > PHC_ACT_SILU_BWD || (relu == PHC_ACT_SILU_BWD && !mask)) { phc_set_error("phc_gemm_tc5: bad activation code"); return PHC_ERR_INVALID_ARG; }
  if (k_splits > 1 && (!accumulate || relu || mask)) { phc_set_error("phc_gemm_tc5: split-K needs accumulate=1 and a linear epilogue"); return PHC_ERR_INVALID_ARG; }
  CUtensorMap tAh, tAl, tBh, tBl;
  if (!make_map(&tAh, A_hi, lda, M, K, a_kmajor) || !make_map(&tAl, A_lo, lda, M, K, a_kmajor) ||
      !make_map(&tBh, B_hi, ldb, N, K, b_kmajor) || !make_map(&tBl, B_lo, ldb, N, K, b_kmajor)) {
    phc_set_error("phc_gemm_tc5: cuTensorMapEncodeTiled failed"); return PHC_ERR_CUDA;
  }
  Args g;
  if ((C_hi == nullptr) != (C_lo == nullptr) || (C_hi && accumulate)) { phc_set_error("phc_gemm_tc5: C_hi/C_lo come as a pair and not with accumulate"); return PHC_ERR_INVALID_ARG; }
  g.C = C; g.C_hi = C_hi; g.C_lo = C_lo; g.bias = bias; g.mask = mask; g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.ldmask = ldmask; g.alpha = alpha;
  g.relu = relu; g.accumulate = accumulate; g.k_splits = k_splits;
  // CTA-pair tiles (256 x 256) when the problem is big enough in both dimensions, else single-CTA 128 x 128 tiles
  // measured on the PPO shapes (bench.py): single-CTA tiles with the persistent, epilogue-overlapped kernel beat the pair
  // tiles, so the pair path is opt-in (PHC_TC5_PAIR=1) and the persistent kernel is the default (PHC_TC5_PERSIST=0 disables)
  const bool pair = (M > BM) && (N > BN + BN / 2) && getenv("PHC_TC5_PAIR") && getenv("PHC_TC5_PAIR")[0] == '1';
  const bool pairp = !pair && (M > BM) && (N > BN) && getenv("PHC_TC5_PAIRP") && getenv("PHC_TC5_PAIRP")[0] == '1';
  const bool persist = !pair && !pairp && !(getenv("PHC_TC5_PERSIST") && getenv("PHC_TC5_PERSIST")[0] == '0');
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (pair) {
    cfg.gridDim = dim3(2 * ((M + 2 * BM - 1) / (2 * BM)), (N + 2 * BN - 1) / (2 * BN), k_splits);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3((N + BN - 1) / BN, (M + BM - 1) / BM, k_splits);
    cfg.attrs = nullptr; cfg.numAttrs = 0;
  }
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
#define PHC_TC5_LAUNCH(AK, BK_, NC)                                                                                     \
  do {                                                                                                                  \
    static bool done = false;                                                                                           \
    if (!done) {                                                                                                        \
      e = cudaFuncSetAttribute(gemm_tc5_kernel<AK, BK_, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);  \
      if (e != cudaSuccess) return phc_check_cuda(e, "cudaFuncSetAttribute(gemm_tc5)");                                 \
      done = true;                                                                                                      \
    }                                                                                                                   \
    e = cudaLaunchKernelEx(&cfg, gemm_tc5_kernel<AK, BK_, NC>, tAh, tAl, tBh, tBl, g);                                  \
    if (e != cudaSuccess) return phc_check_cuda(e, "cudaLaunchKernelEx(gemm_tc5)");                                     \
    phc_count_launches(1);                                                                                              \
  } while (0)
#define PHC_TC5_DISPATCH(NC)                                       \
  do {                                                             \
    if (a_kmajor && b_kmajor) PHC_TC5_LAUNCH(true, true, NC);      \
    else if (a_kmajor && !b_kmajor) PHC_TC5_LAUNCH(true, false, NC); \
    else if (!a_kmajor && b_kmajor) PHC_TC5_LAUNCH(false, true, NC); \
    else PHC_TC5_LAUNCH(false, false, NC);                         \
  } while (0)
  if (pairp) {
    static int num_sms2 = 0;
    if (!num_sms2) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms2, cudaDevAttrMultiProcessorCount, dev); if (num_sms2 <= 0) num_sms2 = 148; }
    const int tiles2 = ((M + 2 * BM - 1) / (2 * BM)) * ((N + 2 * BN - 1) / (2 * BN)) * k_splits;
    const int pairs = tiles2 < num_sms2 / 2 ? tiles2 : num_sms2 / 2;
    cfg.gridDim = dim3(2 * pairs);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
#define PHC_TC5_PPLAUNCH(AK, BK_)                                                                                                \
  do {                                                                                                                           \
    static bool done = false;                                                                                                    \
    if (!done) {                                                                                                                 \
      e = cudaFuncSetAttribute(gemm_tc5_persist_pair_kernel<AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);  \
      if (e != cudaSuccess) return phc_check_cuda(e, "cudaFuncSetAttribute(gemm_tc5_persist_pair)");                             \
      done = true;                                                                                                               \
    }                                                                                                                            \
    e = cudaLaunchKernelEx(&cfg, gemm_tc5_persist_pair_kernel<AK, BK_>, tAh, tAl, tBh, tBl, g);                                  \
    if (e != cudaSuccess) return phc_check_cuda(e, "cudaLaunchKernelEx(gemm_tc5_persist_pair)");                                 \
    phc_count_launches(1);                                                                                                       \
  } while (0)
    if (a_kmajor && b_kmajor) PHC_TC5_PPLAUNCH(true, true);
    else if (a_kmajor && !b_kmajor) PHC_TC5_PPLAUNCH(true, false);
    else if (!a_kmajor && b_kmajor) PHC_TC5_PPLAUNCH(false, true);
    else PHC_TC5_PPLAUNCH(false, false);
#undef PHC_TC5_PPLAUNCH
  } else if (persist) {
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * k_splits;
    cfg.gridDim = dim3(tiles < num_sms ? tiles : num_sms);
#define PHC_TC5_PLAUNCH(AK, BK_)                                                                                            \
  do {                                                                                                                      \
    static bool done = false;                                                                                               \
    if (!done) {                                                                                                            \
      e = cudaFuncSetAttribute(gemm_tc5_persist_kernel<AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);  \
      if (e != cudaSuccess) return phc_check_cuda(e, "cudaFuncSetAttribute(gemm_tc5_persist)");                             \
      done = true;                                                                                                          \
    }                                                                                                                       \
    e = cudaLaunchKernelEx(&cfg, gemm_tc5_persist_kernel<AK, BK_>, tAh, tAl, tBh, tBl, g);                                  \
    if (e != cudaSuccess) return phc_check_cuda(e, "cudaLaunchKernelEx(gemm_tc5_persist)");                                 \
    phc_count_launches(1);                                                                                                  \
  } while (0)
    if (a_kmajor && b_kmajor) PHC_TC5_PLAUNCH(true, true);
    else if (a_kmajor && !b_kmajor) PHC_TC5_PLAUNCH(true, false);
    else if (!a_kmajor && b_kmajor) PHC_TC5_PLAUNCH(false, true);
    else PHC_TC5_PLAUNCH(false, false);
#undef PHC_TC5_PLAUNCH
  } else if (pair) PHC_TC5_DISPATCH(2);
  else PHC_TC5_DISPATCH(1);
#undef PHC_TC5_DISPATCH
#undef PHC_TC5_LAUNCH
  return phc_check_cuda(cudaGetLastError(), "gemm_tc5_kernel launch");
}
