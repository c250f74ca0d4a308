// Grouped, persistent 3xTF32 GEMM on tcgen05 / TMEM / TMA, WIDE tiles: 128 x 256 x 16 per CTA.
//
// Same contract, roles and numerics as gemm_tc5s.cu (raw fp32 tiles by TMA, lo = rna_tf32(x - trunc_tf32(x)) made by eight splitter
// warps in shared memory, three kind::tf32 products per k-step, TMA store / reduce-add epilogue, up to 8 problems per launch, tiles
// drawn from a global counter).  What changes is the tile, because the 128 x 128 x 32 kernel is bound by operand BYTES, not by the
// tensor pipe (profiles/gemm_r2_ab_layout.md): per k-block it moves 32 KB L2 -> shared memory, its splitters read and write another
// 64 KB of shared memory, and the MMAs fetch 96 KB of operands from it.  A 128 x 256 tile does twice the flops on 1.5x the bytes:
//   L2 -> smem bytes / flop  -25 %,   splitter traffic / flop  -25 %,   MMA operand fetches / flop  -25 %  (N = 256 instructions).
// The price is shared memory: a 32-deep k-block of such a tile would be 48 KB raw + 48 KB lo, too few stages.  So the k-block is 16
// deep (64-byte rows, SWIZZLE_64B for k-contiguous operands; 16-row boxes of the 128-byte MN-major layout otherwise):
//   6 raw stages x 24 KB + 2 lo slots x 24 KB + 32 KB epilogue staging = 224 KB.
// TMEM: two 256-column fp32 accumulators = all 512 columns (one CTA per SM is guaranteed by the shared-memory footprint).
// The tile width is a per-problem property (256, or 128 for N <= 128: heads), and the MMA's N is trimmed to the 32-column multiple
// that covers the valid part of a ragged last tile (N = 934: 166 -> 192 columns instead of 256).
// CTA = 448 threads: warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA producer + tile scheduler, warp 5 TMEM owner + MMA
// issuer, warps 6-13 splitters.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/phc_b200.h"
#include "phc_common.cuh"
#include "tc5_common.cuh"

extern "C" void phc_set_error(const char* msg);
extern "C" int phc_check_cuda(cudaError_t e, const char* what);
extern "C" void phc_count_launches(int n);

namespace phc {
namespace tc5 {
namespace wide_tile {

constexpr int BM = 128, BN = 256, BK = 16;
constexpr int A_TILE = BM * BK * 4;                   // 8 KB
constexpr int B_TILE = BN * BK * 4;                   // 16 KB
constexpr int STAGE = A_TILE + B_TILE;                // 24 KB: raw stage and lo slot alike
constexpr int RAW_STAGES = 6, LO_STAGES = 2;
constexpr int NUM_SPLIT_WARPS = 8;
constexpr int NUM_THREADS = (6 + NUM_SPLIT_WARPS) * 32;
constexpr int EPI_BUF = 32 * 32 * 4;                  // one 32 x 32 fp32 chunk per warp
constexpr int EPI_BYTES = 4 * 2 * EPI_BUF;            // 4 warps, double buffered
constexpr int MAX_PROBLEMS = PHC_GEMM_GROUP_MAX;
constexpr int SCHED_DEPTH = 4;
constexpr int SMEM = RAW_STAGES * STAGE + LO_STAGES * STAGE + EPI_BYTES + 1024 /*align slack*/ + 384 /*barriers, scheduler ring*/;
static_assert(SMEM <= 232448, "shared memory budget of one CTA");

struct Prob {
  const float* bias;
  float* aux;
  long long ldaux;
  int M, N, K;
  float alpha;
  int act, accumulate, k_splits, a_k, b_k;
  int bn;               // tile width of this problem: 256 or 128
  int tiles_m, tiles_n, kb_total, kb_per, tile_begin, tile_count;
};

struct alignas(64) Params {
  CUtensorMap tmA[MAX_PROBLEMS];
  CUtensorMap tmB[MAX_PROBLEMS];
  CUtensorMap tmC[MAX_PROBLEMS];
  Prob p[MAX_PROBLEMS];
  int count, total_tiles;
  unsigned int* sched;  // {next tile, CTAs done}, both zero between launches; NULL = static striding
};

struct Tile { int g, m0, n0, kb_begin, nkb, z; };

__device__ __forceinline__ Tile decode(const Params& P, int t) {
  int g = 0;
#pragma unroll 1
  while (g + 1 < P.count && t >= P.p[g].tile_begin + P.p[g].tile_count) ++g;
  const Prob& q = P.p[g];
  const int tl = t - q.tile_begin;
  const int ni = tl % q.tiles_n;
  const int r = tl / q.tiles_n;
  const int mi = r % q.tiles_m;
  Tile o;
  o.g = g;
  o.z = r / q.tiles_m;
  o.m0 = mi * BM;
  o.n0 = ni * q.bn;
  o.kb_begin = o.z * q.kb_per;
  const int kb_end = min(q.kb_total, o.kb_begin + q.kb_per);
  o.nkb = max(0, kb_end - o.kb_begin);
  return o;
}

// lo = rna_tf32(x - trunc_tf32(x)), integer arithmetic on the bit pattern (see gemm_tc5s.cu)
__device__ __forceinline__ float split_lo(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float d = x - hi;
  return __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xFFFFE000u);
}

template <bool SINGLE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc5w_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lo_smem = smem + RAW_STAGES * STAGE;
  uint8_t* epi_smem = lo_smem + LO_STAGES * STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);     // [R] TMA bytes landed
  uint64_t* raw_empty = full_bar + RAW_STAGES;                               // [R] MMAs that read the raw stage are done
  uint64_t* lo_full = raw_empty + RAW_STAGES;                                // [L] lo slot written by all splitter warps
  uint64_t* lo_empty = lo_full + LO_STAGES;                                  // [L] MMAs that read the lo slot are done
  uint64_t* tmem_full = lo_empty + LO_STAGES;       // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint64_t* sched_full = tmem_empty + 2;            // [SCHED_DEPTH] tile id published by the producer warp
  uint64_t* sched_empty = sched_full + SCHED_DEPTH; // [SCHED_DEPTH] every consumer warp has read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_empty + SCHED_DEPTH);
  volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [SCHED_DEPTH]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;     // shfl: provably warp-uniform
  const bool elected = elect_one();
  const int unit = blockIdx.x, num_units = gridDim.x;
  const bool dyn = P.sched != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(raw_empty + s, 1); }
    for (int s = 0; s < LO_STAGES; ++s) { mbar_init(lo_full + s, NUM_SPLIT_WARPS); mbar_init(lo_empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 4); }
    for (int a = 0; a < SCHED_DEPTH; ++a) { mbar_init(sched_full + a, 1); mbar_init(sched_empty + a, SINGLE ? 5 : 5 + NUM_SPLIT_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(tmem_slot, 2 * BN);
  if (warp == 4 && lane == 0)
    for (int g = 0; g < P.count; ++g) { prefetch_tensormap(&P.tmA[g]); prefetch_tensormap(&P.tmB[g]); prefetch_tensormap(&P.tmC[g]); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // i-th tile of this CTA for a consumer role: the whole warp waits for the id, one lane acknowledges
  auto consumer_tile = [&](uint32_t i) -> int {
    if (!dyn) return unit + (int)i * num_units;
    const uint32_t a = i % SCHED_DEPTH;
    mbar_wait(sched_full + a, (i / SCHED_DEPTH) & 1);
    const int t = sched_tile[a];
    __syncwarp();
    if (elected) mbar_arrive(sched_empty + a);
    return t;
  };

  if (warp == 4) {
    // ===================== TMA producer + tile scheduler =====================
    uint32_t it = 0;                                          // k-block counter, continues across tiles
    auto draw = [&]() -> int {
      int v = 0;
      if (elected) v = (int)atomicAdd(P.sched, 1u);
      return __shfl_sync(0xffffffffu, v, __ffs(__ballot_sync(0xffffffffu, elected)) - 1);
    };
    int t_next = dyn ? draw() : 0;
    for (uint32_t ti = 0;; ++ti) {
      int t;
      if (dyn) {
        t = t_next;
        const uint32_t a = ti % SCHED_DEPTH;
        if (ti >= (uint32_t)SCHED_DEPTH) mbar_wait(sched_empty + a, ((ti / SCHED_DEPTH) - 1) & 1);
        if (elected) { sched_tile[a] = t; mbar_arrive(sched_full + a); }      // (the terminating id is published too)
        __syncwarp();
        if (t >= P.total_tiles) break;
        t_next = draw();                                      // in flight while this tile's loads are issued
      } else {
        t = unit + (int)ti * num_units;
        if (t >= P.total_tiles) break;
      }
      const Tile tl = decode(P, t);
      const Prob& q = P.p[tl.g];
      const CUtensorMap* tA = &P.tmA[tl.g];
      const CUtensorMap* tB = &P.tmB[tl.g];
      const bool ak = q.a_k != 0, bk = q.b_k != 0;
      const int bn = q.bn;
      const uint32_t tx = (uint32_t)(A_TILE + bn * BK * 4);
      for (int i = 0; i < tl.nkb; ++i, ++it) {
        const int s = it % RAW_STAGES;
        if (it >= (uint32_t)RAW_STAGES) mbar_wait(raw_empty + s, ((it / RAW_STAGES) - 1) & 1);
        uint8_t* st = smem + s * STAGE;
        const int k0 = (tl.kb_begin + i) * BK;
        if (elected) {
          mbar_expect_tx(full_bar + s, tx);
          if (ak) tma_load_2d(st, tA, full_bar + s, k0, tl.m0);                      // box {16 k, 128 rows}, 64-byte swizzle
          else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(st + j * 2048, tA, full_bar + s, tl.m0 + 32 * j, k0);     // boxes {32 m, 16 k}
          }
          if (bk) tma_load_2d(st + A_TILE, tB, full_bar + s, k0, tl.n0);             // box {16 k, bn rows}
          else {
#pragma unroll 1
            for (int j = 0; j < bn / 32; ++j) tma_load_2d(st + A_TILE + j * 2048, tB, full_bar + s, tl.n0 + 32 * j, k0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    uint32_t it = 0;
    for (uint32_t lt = 0;; ++lt) {
      const int t = consumer_tile(lt);
      if (t >= P.total_tiles) break;
      const Tile tl = decode(P, t);
      const Prob& q = P.p[tl.g];
      const bool ak = q.a_k != 0, bk = q.b_k != 0;
      int n_mma = (min(q.bn, q.N - tl.n0) + 31) & ~31;           // N of the instruction: the valid columns, rounded up to 32
      const uint32_t idesc = instr_desc(!ak, !bk, BM, n_mma);
      // k-contiguous: 64-byte rows, SWIZZLE_64B (layout 4), 8-row groups 512 B apart; MN-contiguous: 32-float rows, 128B / 32B-atom
      // swizzle (layout 1), 16-row boxes of 2 KB per 32 columns
      const uint32_t a_lbo = ak ? 16 : 2048, b_lbo = bk ? 16 : 2048;
      const uint32_t a_sbo = 512, b_sbo = 512;
      const uint32_t a_lt = ak ? 4 : 1, b_lt = bk ? 4 : 1;
      const uint32_t a_step = ak ? 32 : 1024, b_step = bk ? 32 : 1024;
      const uint32_t acc = lt & 1, use = lt >> 1;
      if (use > 0) mbar_wait(tmem_empty + acc, (use - 1) & 1);    // the epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int i = 0; i < tl.nkb; ++i, ++it) {
        const int s = it % RAW_STAGES, l = it % LO_STAGES;
        mbar_wait(full_bar + s, (it / RAW_STAGES) & 1);
        if (!SINGLE) mbar_wait(lo_full + l, (it / LO_STAGES) & 1);
        tc_fence_after();
        const uint32_t st = s32(smem + s * STAGE), sl = s32(lo_smem + l * STAGE);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t dAh = smem_desc(st + kk * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t dAl = smem_desc(sl + kk * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t dBh = smem_desc(st + A_TILE + kk * b_step, b_lbo, b_sbo, b_lt);
          const uint64_t dBl = smem_desc(sl + A_TILE + kk * b_step, b_lbo, b_sbo, b_lt);
          const uint32_t first = (i > 0 || kk > 0) ? 1u : 0u;
          if (elected) {
            if (SINGLE) {
              umma_tf32(tmem_d, dAh, dBh, idesc, first);
            } else {
              umma_tf32(tmem_d, dAl, dBh, idesc, first);
              umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
              umma_tf32(tmem_d, dAh, dBh, idesc, 1u);
            }
          }
        }
        if (elected) { umma_commit(raw_empty + s); if (!SINGLE) umma_commit(lo_empty + l); }
        __syncwarp();
      }
      if (elected) umma_commit(tmem_full + acc);
      __syncwarp();
    }
  } else if (warp >= 6) {
    // ===================== splitters: lo tiles of every landed stage =====================
    const int tid = threadIdx.x - 6 * 32;
    constexpr int NT = NUM_SPLIT_WARPS * 32;                                  // 256 threads x 16 B = 4 KB per pass
    constexpr int PER_A = A_TILE / 16 / NT;                                   // 2
    constexpr int PER_B = B_TILE / 16 / NT;                                   // 4 (2 used by a 128-wide tile)
    uint32_t it = 0;
    for (uint32_t lt = 0; !SINGLE; ++lt) {
      const int t = consumer_tile(lt);
      if (t >= P.total_tiles) break;
      const Tile tl = decode(P, t);
      const bool wide = P.p[tl.g].bn > 128;
      for (int i = 0; i < tl.nkb; ++i, ++it) {
        const int s = it % RAW_STAGES, l = it % LO_STAGES;
        mbar_wait(full_bar + s, (it / RAW_STAGES) & 1);
        const uint32_t raw = s32(smem + s * STAGE) + (uint32_t)tid * 16u;
        const uint32_t lo = s32(lo_smem + l * STAGE) + (uint32_t)tid * 16u;
        float4 v[PER_A + PER_B];
#pragma unroll
        for (int j = 0; j < PER_A + PER_B / 2; ++j) v[j] = lds128(raw + j * NT * 16);
        if (wide) {
#pragma unroll
          for (int j = PER_A + PER_B / 2; j < PER_A + PER_B; ++j) v[j] = lds128(raw + j * NT * 16);
        }
        if (it >= (uint32_t)LO_STAGES) mbar_wait(lo_empty + l, ((it / LO_STAGES) - 1) & 1);   // the MMAs of k-block it - 2 are done
#pragma unroll
        for (int j = 0; j < PER_A + PER_B / 2; ++j)
          sts128(lo + j * NT * 16, split_lo(v[j].x), split_lo(v[j].y), split_lo(v[j].z), split_lo(v[j].w));
        if (wide) {
#pragma unroll
          for (int j = PER_A + PER_B / 2; j < PER_A + PER_B; ++j)
            sts128(lo + j * NT * 16, split_lo(v[j].x), split_lo(v[j].y), split_lo(v[j].z), split_lo(v[j].w));
        }
        fence_proxy_async_smem();                              // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (elected) mbar_arrive(lo_full + l);
      }
    }
  } else {
    // ===================== epilogue warps 0..3: TMEM -> registers -> swizzled smem chunk -> TMA store =====================
    uint8_t* my_buf = epi_smem + warp * 2 * EPI_BUF;
    uint32_t chunk = 0;
    for (uint32_t lt = 0;; ++lt) {
      const int t = consumer_tile(lt);
      if (t >= P.total_tiles) break;
      const Tile tl = decode(P, t);
      const Prob& q = P.p[tl.g];
      const CUtensorMap* tC = &P.tmC[tl.g];
      const uint32_t acc = lt & 1, use = lt >> 1;
      mbar_wait(tmem_full + acc, use & 1);
      tc_fence_after();
      const int m = tl.m0 + warp * 32 + lane;
      const bool row_ok = m < q.M;
      const int act = q.act;
      const float alpha = q.alpha;
      const float* bias = (q.bias && tl.z == 0) ? q.bias : nullptr;
      float* arow = (q.aux && act < PHC_ACT_RELU_BITS) ? q.aux + (long long)m * q.ldaux : nullptr;
      const bool aux_vec = arow && ((q.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.aux) & 15) == 0);
      const int bn = q.bn;
      if (tl.nkb > 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < bn; c0 += 32) {
          const int nb = tl.n0 + c0;
          if (nb >= q.N) break;                                  // warp-uniform
          const float bias_l = (bias && nb + lane < q.N) ? bias[nb + lane] : 0.0f;
          uint32_t* bits = (act >= PHC_ACT_RELU_BITS && q.aux && row_ok)
                               ? reinterpret_cast<uint32_t*>(q.aux) + (long long)m * q.ldaux + (nb >> 5) : nullptr;
          const uint32_t mbits = (act == PHC_ACT_MASK_BITS && bits) ? *bits : 0u;
          uint32_t r[32];
          tmem_ld32(tmem_base + acc * BN + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
          if (c0 + 32 >= bn || nb + 32 >= q.N) {                 // last chunk read: hand the accumulator back before the math / store
            tc_fence_before();
            __syncwarp();
            if (elected) mbar_arrive(tmem_empty + acc);
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = alpha * __uint_as_float(r[j]);
            if (bias) x += __shfl_sync(0xffffffffu, bias_l, j);
            if (act == PHC_ACT_RELU || act == PHC_ACT_RELU_BITS) x = fmaxf(x, 0.f);
            v[j] = x;
          }
          if (act == PHC_ACT_RELU_BITS) {                        // ReLU forward: 1 bit per element for the backward pass
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
            if (bits) *bits = w;
          } else if (act == PHC_ACT_MASK_BITS) {                 // ReLU backward from the saved bits
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((mbits >> j) & 1u) ? v[j] : 0.f;
          } else if (act == PHC_ACT_SILU) {
            if (arow && row_ok) {                                // pre-activation out
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (aux_vec && nb + j + 3 < q.N) *reinterpret_cast<float4*>(arow + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                else for (int e = 0; e < 4; ++e) if (nb + j + e < q.N) arow[nb + j + e] = v[j + e];
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
          } else if (arow) {                                     // ReLU backward mask / SiLU backward factor
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float mk[4] = {0.f, 0.f, 0.f, 0.f};
              if (row_ok) {
                if (aux_vec && nb + j + 3 < q.N) {
                  const float4 t4 = *reinterpret_cast<const float4*>(arow + nb + j);
                  mk[0] = t4.x; mk[1] = t4.y; mk[2] = t4.z; mk[3] = t4.w;
                } else {
                  for (int e = 0; e < 4; ++e) if (nb + j + e < q.N) mk[e] = arow[nb + j + e];
                }
              }
#pragma unroll
              for (int e = 0; e < 4; ++e)
                v[j + e] = (act == PHC_ACT_SILU_BWD) ? v[j + e] * silu_grad_f(mk[e]) : (mk[e] > 0.f ? v[j + e] : 0.f);
            }
          }
          // stage the 32 x 32 chunk (row = lane) in the 128-byte-swizzled layout the C tensor map expects
          uint8_t* buf = my_buf + (chunk & 1) * EPI_BUF;
          ++chunk;
          if (elected) bulk_wait_group_read<1>();                // the store issued two chunks ago has read this buffer
          __syncwarp();
          const uint32_t brow = s32(buf) + (uint32_t)lane * 128u;
#pragma unroll
          for (int c = 0; c < 8; ++c) sts128(brow + ((uint32_t)(c ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (elected) {
            if (q.accumulate) tma_reduce_add_2d(tC, buf, nb, tl.m0 + warp * 32);
            else tma_store_2d(tC, buf, nb, tl.m0 + warp * 32);
            bulk_commit_group();
          }
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (elected) mbar_arrive(tmem_empty + acc);
      }
    }
    if (elected) bulk_wait_group_read<0>();                    // staging buffers must outlive the stores' reads
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
  // the last CTA to get here puts both counters back to zero for the next launch (every CTA has drawn its terminating tile by now)
  if (dyn && threadIdx.x == 0 && atomicInc(P.sched + 1, gridDim.x - 1) == gridDim.x - 1) { __threadfence(); P.sched[0] = 0u; }
}

// 2-D fp32 tensor map; inner dimension = the contiguous one
static bool make_map(CUtensorMap* tm, const float* base, int64_t ld, uint64_t inner, uint64_t outer, uint32_t box_inner,
                     uint32_t box_outer, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {inner, outer}, strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {box_inner, box_outer}, estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int SCHED_SLOTS = 64;
__device__ unsigned int g_sched_counters[SCHED_SLOTS][2];

}  // namespace wide_tile
}  // namespace tc5
}  // namespace phc

// the wide-tile implementation behind phc_gemm_group (dispatch in gemm_tc5s.cu); problems are validated by the caller
extern "C" int phc_gemm_group_wide(const PhcGemmDesc* d, int32_t count, int32_t single_pass, int32_t dynamic_sched, void* stream) {
  using namespace phc::tc5::wide_tile;
  static Params P;      // host staging (the struct is copied at launch)
  memset(&P.p, 0, sizeof(P.p));
  int tiles = 0, n = 0;
  for (int i = 0; i < count; ++i) {
    const PhcGemmDesc& g = d[i];
    if (g.M == 0 || g.N == 0) continue;
    int ks = g.k_splits < 1 ? 1 : g.k_splits;
    Prob& q = P.p[n];
    q.bias = g.bias; q.aux = g.aux; q.ldaux = g.ldaux; q.M = g.M; q.N = g.N; q.K = g.K; q.alpha = g.alpha; q.act = g.act;
    q.accumulate = g.accumulate ? 1 : 0; q.a_k = g.a_kmajor ? 1 : 0; q.b_k = g.b_kmajor ? 1 : 0;
    q.bn = g.N > 128 ? 256 : 128;
    q.tiles_m = (g.M + BM - 1) / BM;
    q.tiles_n = (g.N + q.bn - 1) / q.bn;
    q.kb_total = (g.K + BK - 1) / BK;
    if (ks > q.kb_total) ks = q.kb_total;
    q.k_splits = ks;
    q.kb_per = (q.kb_total + ks - 1) / ks;
    q.tile_begin = tiles;
    q.tile_count = q.tiles_m * q.tiles_n * ks;
    tiles += q.tile_count;
    bool ok = q.a_k ? make_map(&P.tmA[n], g.A, g.lda, (uint64_t)g.K, (uint64_t)g.M, BK, BM, CU_TENSOR_MAP_SWIZZLE_64B)
                    : make_map(&P.tmA[n], g.A, g.lda, (uint64_t)g.M, (uint64_t)g.K, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    ok = ok && (q.b_k ? make_map(&P.tmB[n], g.B, g.ldb, (uint64_t)g.K, (uint64_t)g.N, BK, (uint32_t)q.bn, CU_TENSOR_MAP_SWIZZLE_64B)
                      : make_map(&P.tmB[n], g.B, g.ldb, (uint64_t)g.N, (uint64_t)g.K, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    ok = ok && make_map(&P.tmC[n], g.C, g.ldc, (uint64_t)g.N, (uint64_t)g.M, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) { phc_set_error("phc_gemm_group: cuTensorMapEncodeTiled failed (wide tiles)"); return PHC_ERR_CUDA; }
    ++n;
  }
  if (n == 0) return PHC_OK;
  P.count = n; P.total_tiles = tiles;
  P.sched = nullptr;
  if (dynamic_sched) {
    static unsigned int* base = nullptr;
    static unsigned int launch_no = 0;
    if (!base) {
      void* p = nullptr;
      cudaError_t es = cudaGetSymbolAddress(&p, g_sched_counters);
      if (es != cudaSuccess) return phc_check_cuda(es, "cudaGetSymbolAddress(g_sched_counters)");
      base = static_cast<unsigned int*>(p);
    }
    P.sched = base + 2 * (launch_no++ % SCHED_SLOTS);
  }
  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles < num_sms ? tiles : num_sms));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cfg.attrs = nullptr; cfg.numAttrs = 0;
  cfg.dynamicSmemBytes = SMEM;
  using Kernel = void (*)(const Params);
  static const Kernel kernels[2] = {gemm_tc5w_kernel<false>, gemm_tc5w_kernel<true>};
  static bool smem_set[2] = {false, false};
  const int ki = single_pass ? 1 : 0;
  cudaError_t e;
  if (!smem_set[ki]) {
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernels[ki]), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return phc_check_cuda(e, "cudaFuncSetAttribute(gemm_tc5w)");
    smem_set[ki] = true;
  }
  e = cudaLaunchKernelEx(&cfg, kernels[ki], P);
  if (e != cudaSuccess) return phc_check_cuda(e, "cudaLaunchKernelEx(gemm_tc5w)");
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "gemm_tc5w_kernel launch");
}
