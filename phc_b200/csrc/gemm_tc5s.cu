// Grouped, persistent 3xTF32 GEMM on tcgen05 / TMEM / TMA with the operand split done IN SHARED MEMORY.
//
// Same contract as phc_gemm (gemm.cu): C[M,N] (+)= epi(alpha * sum_k A(m,k) B(n,k)), both operand major-nesses, fp32 in and
// out, fp32-equivalent numerics.  Differences to gemm_tc5.cu (operands pre-split into hi / lo arrays in global memory):
//   * TMA stages the RAW fp32 tiles; eight "splitter" warps read each landed stage and write lo = rna_tf32(x - trunc_tf32(x))
//     into a second shared-memory tile of the same (swizzled) layout.  tcgen05.mma kind::tf32 ignores the 13 low mantissa
//     bits of its operands, so the raw tile IS the hi operand (hi = trunc_tf32(x), hi + lo = x to 2^-22 |x|); per k-step the
//     issuing thread launches D += A_lo*B_hi, D += A_hi*B_lo, D += A_hi*B_hi.  The dropped A_lo*B_lo term is <= 2^-20 |ab|
//     with mean 2^-22 (truncation makes it one-signed: a relative bias of 2.4e-7 on the result, far below the 1e-5 bar).
//     Half the L2 -> shared-memory bytes per flop, no hi / lo copies of activations or weights in HBM, no split passes, and the
//     epilogue writes ONE output instead of three.
//   * The epilogue goes TMEM -> registers -> (bias, activation) -> a 128-byte-swizzled shared-memory tile per warp -> TMA
//     tensor store (or TMA reduce-add for accumulate / split-K): 4 KB coalesced bulk writes instead of 16-byte row pieces.
//   * One launch takes up to 6 independent problems (the same layer of actor, critic and discriminator; dW and dX of one
//     layer): their tiles form one list walked by one persistent CTA (or CTA pair) per SM -- fewer launches, and the tail
//     quantisation (tiles mod 148) is paid once per group instead of once per GEMM.
// CTAS = 1: 128 x 128 tile per CTA, 3 stages of 64 KB (A raw, B raw, A lo, B lo).
// CTAS = 2: a CTA pair (cluster of 2, cta_group::2) computes a 256 x 128 tile: each CTA stages its 128 rows of A and 64 of
//           the 128 B rows, 4 stages of 48 KB; the leader issues M = 256, N = 128 MMAs that read both CTAs' shared memory.
// CTA = 448 threads: warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA producer, warp 5 TMEM owner + MMA issuer,
// warps 6-13 splitters.  Barriers per stage: full (TMA bytes landed, local), split (lo tile written: 8 local warps, or 16
// incl. the peer's in pair mode -- on the leader), empty (tcgen05.commit, multicast to both CTAs in pair mode).
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
namespace smem_split {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int A_TILE = BM * BK * 4;                   // 16 KB
constexpr int NUM_SPLIT_WARPS = 8;
constexpr int NUM_THREADS = (6 + NUM_SPLIT_WARPS) * 32;
constexpr int EPI_BUF = 32 * 32 * 4;                  // one 32 x 32 fp32 chunk per warp
constexpr int EPI_BYTES = 4 * 2 * EPI_BUF;            // 4 warps, double buffered
constexpr int MAX_PROBLEMS = PHC_GEMM_GROUP_MAX;

constexpr int SCHED_DEPTH = 4;      // tile ids in flight between the producer warp and the slowest consumer role

template <int CTAS>
struct Cfg {
  static constexpr int B_ROWS = BN / CTAS;
  static constexpr int B_TILE = B_ROWS * BK * 4;
  static constexpr int RAW = A_TILE + B_TILE;         // bytes of the raw A and B tiles of a stage
  // Default layout: the TMA ring holds the raw tiles only (4 / 6 stages of 32 / 24 KB), both lo tiles go to the two-slot lo ring (a lo
  // tile is needed only from its split to the MMAs of its k-block).
  // -DPHC_TC5S_BLO_IN_STAGE (A/B build, measured and not adopted -- profiles/gemm_r2_ab_layout.md): a stage is [A raw | B raw | B lo]
  // and the B lo tile may arrive by TMA from a pre-split copy of the weights (PhcGemmDesc.B_lo) instead of being made by the
  // splitters; the lo ring then holds A lo only.  Shared memory leaves 3 (5) such stages: single GEMMs gain 4-8 % from the halved
  // splitter traffic, the grouped launches of the learner lose 2-4 % to the shallower ring and the 50 % higher L2 -> shared-memory
  // traffic.  In the default layout B_lo is accepted and ignored.
#ifdef PHC_TC5S_BLO_IN_STAGE
  static constexpr bool BLO_IN_STAGE = true;
#else
  static constexpr bool BLO_IN_STAGE = false;
#endif
  static constexpr int STAGE = BLO_IN_STAGE ? A_TILE + 2 * B_TILE : RAW;
  static constexpr int LO_SLOT = BLO_IN_STAGE ? A_TILE : RAW;
  static constexpr int RAW_STAGES = BLO_IN_STAGE ? (CTAS == 1 ? 3 : 5) : (CTAS == 1 ? 4 : 6);
  static constexpr int LO_STAGES = 2;
  static constexpr int SMEM = RAW_STAGES * STAGE + LO_STAGES * LO_SLOT + EPI_BYTES + 1024 /*align slack*/ + 320 /*barriers, scheduler ring*/;
};

struct Prob {
  const float* bias;
  float* aux;
  long long ldaux;
  int M, N, K;
  float alpha;
  int act, accumulate, k_splits, a_k, b_k, has_blo;
  int tiles_m, tiles_n, kb_total, kb_per, tile_begin, tile_count;
};

struct alignas(64) Params {
  CUtensorMap tmA[MAX_PROBLEMS];
  CUtensorMap tmB[MAX_PROBLEMS];
  CUtensorMap tmBlo[MAX_PROBLEMS];      // pre-split lo copy of B (weights), when the problem has one
  CUtensorMap tmC[MAX_PROBLEMS];
  Prob p[MAX_PROBLEMS];
  int count, total_tiles;
  unsigned int* sched;  // dynamic tile scheduler: {next tile, CTAs done} in global memory (both zero between launches); NULL = static striding
  int single_pass;      // 1: one tensor-core product per fp32 product (plain TF32, ~1e-3 relative): no lo tiles, no splitters
};

struct Tile { int g, m0, n0, kb_begin, nkb, z; };

template <int CTAS>
__device__ __forceinline__ Tile decode(const Params& P, int t, int rank) {
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
  o.m0 = mi * (BM * CTAS) + rank * BM;
  o.n0 = ni * BN;
  o.kb_begin = o.z * q.kb_per;
  const int kb_end = min(q.kb_total, o.kb_begin + q.kb_per);
  o.nkb = max(0, kb_end - o.kb_begin);
  return o;
}

// lo = rna_tf32(x - trunc_tf32(x)): x - trunc is exact in fp32 (13 significant bits), round-to-nearest (ties away) to the 11
// bits a tf32 operand keeps, done with integer arithmetic on the bit pattern
__device__ __forceinline__ float split_lo(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float d = x - hi;
#ifdef PHC_TC5S_TRUNC_LO
  return d;
#else
  return __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xFFFFE000u);
#endif
}

// SINGLE: one tensor-core pass per product (PHC_GEMM_TF32_SINGLE_PASS) -- a compile-time variant so the 3xTF32 issue loop carries
// no run-time mode test
template <int CTAS, bool SINGLE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc5s_kernel(const __grid_constant__ Params P) {
  using C = Cfg<CTAS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lo_smem = smem + C::RAW_STAGES * C::STAGE;          // A lo ring
  uint8_t* epi_smem = lo_smem + C::LO_STAGES * C::LO_SLOT;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);     // [R] TMA bytes landed (local)
  uint64_t* raw_empty = full_bar + C::RAW_STAGES;                            // [R] MMAs that read the raw stage are done
  uint64_t* lo_full = raw_empty + C::RAW_STAGES;                             // [L] lo slot written (all splitter warps; leader's copy)
  uint64_t* lo_empty = lo_full + C::LO_STAGES;                               // [L] MMAs that read the lo slot are done
  uint64_t* tmem_full = lo_empty + C::LO_STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2] (pair mode: the leader's copies are the ones waited on)
  uint64_t* sched_full = tmem_empty + 2;            // [SCHED_DEPTH] tile id published by the producer warp
  uint64_t* sched_empty = sched_full + SCHED_DEPTH; // [SCHED_DEPTH] every consumer warp has read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_empty + SCHED_DEPTH);
  volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [SCHED_DEPTH]
  // Tile order.  Static: unit u takes tiles u, u + units, ...  Dynamic (one-CTA tiles): the producer warp draws the next tile from a
  // global counter and hands it to the other roles through a small shared-memory ring, so a CTA that got long tiles (the K = 1960
  // discriminator layer next to K = 934 actor / critic tiles in one launch) or shares its SM with another stream's kernel simply
  // draws fewer of them.
  const bool dyn = CTAS == 1 && P.sched != nullptr;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;     // shfl: provably warp-uniform
  const bool elected = elect_one();                 // the one lane of each warp that issues TMA / tcgen05 / barrier arrivals
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = blockIdx.x / CTAS, num_units = gridDim.x / CTAS;     // a unit = one CTA or one CTA pair

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::RAW_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(raw_empty + s, 1); }
    for (int s = 0; s < C::LO_STAGES; ++s) { mbar_init(lo_full + s, NUM_SPLIT_WARPS * CTAS); mbar_init(lo_empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 4 * CTAS); }
    for (int a = 0; a < SCHED_DEPTH; ++a) { mbar_init(sched_full + a, 1); mbar_init(sched_empty + a, SINGLE ? 5 : 5 + NUM_SPLIT_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) { if (CTAS == 2) tmem_alloc_2cta(tmem_slot, 2 * BN); else tmem_alloc(tmem_slot, 2 * BN); }
  if (warp == 4 && lane == 0) {
    for (int g = 0; g < P.count; ++g) {
      prefetch_tensormap(&P.tmA[g]); prefetch_tensormap(&P.tmB[g]); prefetch_tensormap(&P.tmC[g]);
      if (P.p[g].has_blo) prefetch_tensormap(&P.tmBlo[g]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();                // peer barriers initialised, both TMEM halves allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // i-th tile of this CTA for a consumer role (MMA, splitters, epilogue): the whole warp waits, one lane acknowledges
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
    // ===================== TMA producer: raw fp32 tiles (whole warp walks the loop, the elected lane issues) =====================
    uint32_t it = 0;                                          // k-block counter, continues across tiles
    auto draw = [&]() -> int {                                // next tile from the global counter (one atomic per warp)
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
      const Tile tl = decode<CTAS>(P, t, (int)rank);
      const Prob& q = P.p[tl.g];
      const CUtensorMap* tA = &P.tmA[tl.g];
      const CUtensorMap* tB = &P.tmB[tl.g];
      const CUtensorMap* tBlo = &P.tmBlo[tl.g];
      const int nb0 = tl.n0 + (int)rank * C::B_ROWS;
      const bool ak = q.a_k != 0, bk = q.b_k != 0, blo = q.has_blo != 0 && !SINGLE && C::BLO_IN_STAGE;
      for (int i = 0; i < tl.nkb; ++i, ++it) {
        const int s = it % C::RAW_STAGES;
        if (it >= (uint32_t)C::RAW_STAGES) mbar_wait(raw_empty + s, ((it / C::RAW_STAGES) - 1) & 1);
        uint8_t* st = smem + s * C::STAGE;
        const int k0 = (tl.kb_begin + i) * BK;
        if (elected) {
          mbar_expect_tx(full_bar + s, blo ? C::RAW + C::B_TILE : C::RAW);
          if (ak) tma_load_2d(st, tA, full_bar + s, k0, tl.m0);
          else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(st + j * 4096, tA, full_bar + s, tl.m0 + 32 * j, k0);
          }
          if (bk) tma_load_2d(st + A_TILE, tB, full_bar + s, k0, nb0);
          else {
#pragma unroll
            for (int j = 0; j < C::B_ROWS / 32; ++j) tma_load_2d(st + A_TILE + j * 4096, tB, full_bar + s, nb0 + 32 * j, k0);
          }
          if (blo) {                                     // the pre-split lo tile of the weights, same boxes, into the stage's B lo area
            uint8_t* sb = st + A_TILE + C::B_TILE;
            if (bk) tma_load_2d(sb, tBlo, full_bar + s, k0, nb0);
            else {
#pragma unroll
              for (int j = 0; j < C::B_ROWS / 32; ++j) tma_load_2d(sb + j * 4096, tBlo, full_bar + s, nb0 + 32 * j, k0);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (leader CTA; whole warp walks the loop, the elected lane issues) =====================
    if (leader) {
      uint32_t it = 0;
      for (uint32_t lt = 0;; ++lt) {
        const int t = consumer_tile(lt);
        if (t >= P.total_tiles) break;
        const Tile tl = decode<CTAS>(P, t, 0);
        const Prob& q = P.p[tl.g];
        const bool ak = q.a_k != 0, bk = q.b_k != 0;
        const uint32_t idesc = instr_desc(!ak, !bk, BM * CTAS, BN);
        const uint32_t a_lbo = ak ? 16 : 4096, b_lbo = bk ? 16 : 4096;
        const uint32_t a_sbo = ak ? 1024 : 512, b_sbo = bk ? 1024 : 512;
        const uint32_t a_lt = ak ? 2 : 1, b_lt = bk ? 2 : 1;
        const uint32_t a_step = ak ? 32 : 1024, b_step = bk ? 32 : 1024;
        const uint32_t acc = lt & 1, use = lt >> 1;
        if (use > 0) mbar_wait(tmem_empty + acc, (use - 1) & 1);    // the epilogue(s) drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int i = 0; i < tl.nkb; ++i, ++it) {
          const int s = it % C::RAW_STAGES, l = it % C::LO_STAGES;
          mbar_wait(full_bar + s, (it / C::RAW_STAGES) & 1);        // own raw tiles (the peer's are implied by its splitters)
          if (!SINGLE) mbar_wait(lo_full + l, (it / C::LO_STAGES) & 1);   // lo tiles of both CTAs written and fenced
          tc_fence_after();
          const uint32_t st = s32(smem + s * C::STAGE), sl = s32(lo_smem + l * C::LO_SLOT);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t dAh = smem_desc(st + kk * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t dAl = smem_desc(sl + kk * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t dBh = smem_desc(st + A_TILE + kk * b_step, b_lbo, b_sbo, b_lt);
            const uint64_t dBl = smem_desc((C::BLO_IN_STAGE ? st + A_TILE + C::B_TILE : sl + A_TILE) + kk * b_step, b_lbo, b_sbo, b_lt);
            const uint32_t first = (i > 0 || kk > 0) ? 1u : 0u;
            if (elected) {
              if (SINGLE) {
                if (CTAS == 2) umma_tf32_2cta(tmem_d, dAh, dBh, idesc, first); else umma_tf32(tmem_d, dAh, dBh, idesc, first);
              } else if (CTAS == 2) {
                umma_tf32_2cta(tmem_d, dAl, dBh, idesc, first);
                umma_tf32_2cta(tmem_d, dAh, dBl, idesc, 1u);
                umma_tf32_2cta(tmem_d, dAh, dBh, idesc, 1u);
              } else {
                umma_tf32(tmem_d, dAl, dBh, idesc, first);
                umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
                umma_tf32(tmem_d, dAh, dBh, idesc, 1u);
              }
            }
          }
          if (elected) {
            if (CTAS == 2) { umma_commit_2cta(raw_empty + s); if (!SINGLE) umma_commit_2cta(lo_empty + l); }   // frees both slots (in both CTAs)
            else { umma_commit(raw_empty + s); if (!SINGLE) umma_commit(lo_empty + l); }
          }
          __syncwarp();
        }
        if (elected) { if (CTAS == 2) umma_commit_2cta(tmem_full + acc); else umma_commit(tmem_full + acc); }
        __syncwarp();
      }
    }
  } else if (warp >= 6) {
    // ===================== splitters: lo tile of every landed stage =====================
    const int tid = threadIdx.x - 6 * 32;
    constexpr int NT = NUM_SPLIT_WARPS * 32;
    constexpr int PER_A = A_TILE / 16 / NT, PER_B = C::B_TILE / 16 / NT;      // float4 per thread: 4 of A, 4 (2) of B
    static_assert(PER_A * NT * 16 == A_TILE && PER_B * NT * 16 == C::B_TILE, "tiles do not divide over the splitter threads");
    uint32_t it = 0;
    for (uint32_t lt = 0; !SINGLE; ++lt) {
      const int t = consumer_tile(lt);
      if (t >= P.total_tiles) break;
      const Tile tl = decode<CTAS>(P, t, (int)rank);
      const bool split_b = P.p[tl.g].has_blo == 0 || !C::BLO_IN_STAGE;          // activations as B: their lo tile is made here, into the stage itself
      for (int i = 0; i < tl.nkb; ++i, ++it) {
        const int s = it % C::RAW_STAGES, l = it % C::LO_STAGES;
        mbar_wait(full_bar + s, (it / C::RAW_STAGES) & 1);
        const uint32_t raw = s32(smem + s * C::STAGE) + (uint32_t)tid * 16u;
        const uint32_t lo = s32(lo_smem + l * C::LO_SLOT) + (uint32_t)tid * 16u;
        float4 va[PER_A], vb[PER_B];
#pragma unroll
        for (int j = 0; j < PER_A; ++j) va[j] = lds128(raw + j * NT * 16);
        if (split_b) {
#pragma unroll
          for (int j = 0; j < PER_B; ++j) vb[j] = lds128(raw + A_TILE + j * NT * 16);
          if (C::BLO_IN_STAGE) {
#pragma unroll
            for (int j = 0; j < PER_B; ++j)                // the stage is this k-block's own: no further wait
              sts128(raw + A_TILE + C::B_TILE + j * NT * 16, split_lo(vb[j].x), split_lo(vb[j].y), split_lo(vb[j].z), split_lo(vb[j].w));
          }
        }
        if (it >= (uint32_t)C::LO_STAGES) mbar_wait(lo_empty + l, ((it / C::LO_STAGES) - 1) & 1);   // the MMAs of k-block it - 2 are done
        if (!C::BLO_IN_STAGE) {
#pragma unroll
          for (int j = 0; j < PER_B; ++j)
            sts128(lo + A_TILE + j * NT * 16, split_lo(vb[j].x), split_lo(vb[j].y), split_lo(vb[j].z), split_lo(vb[j].w));
        }
#pragma unroll
        for (int j = 0; j < PER_A; ++j)
          sts128(lo + j * NT * 16, split_lo(va[j].x), split_lo(va[j].y), split_lo(va[j].z), split_lo(va[j].w));
        fence_proxy_async_smem();                              // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (elected) { if (CTAS == 2) mbar_arrive_remote_leader(lo_full + l); else mbar_arrive(lo_full + l); }
      }
    }
  } else {
    // ===================== epilogue warps 0..3: TMEM -> registers -> swizzled smem chunk -> TMA store =====================
    uint8_t* my_buf = epi_smem + warp * 2 * EPI_BUF;
    uint32_t chunk = 0;
    for (uint32_t lt = 0;; ++lt) {
      const int t = consumer_tile(lt);
      if (t >= P.total_tiles) break;
      const Tile tl = decode<CTAS>(P, t, (int)rank);
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
      if (tl.nkb > 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int nb = tl.n0 + c0;
          if (nb >= q.N) break;                                  // warp-uniform
          // loads that do not depend on the accumulator go first: the bias slice of this chunk (one coalesced load, handed
          // round by shuffles) and, for the bit-mask modes, this row's 32 sign bits
          const float bias_l = (bias && nb + lane < q.N) ? bias[nb + lane] : 0.0f;
          uint32_t* bits = (act >= PHC_ACT_RELU_BITS && q.aux && row_ok)
                               ? reinterpret_cast<uint32_t*>(q.aux) + (long long)m * q.ldaux + (nb >> 5) : nullptr;
          const uint32_t mbits = (act == PHC_ACT_MASK_BITS && bits) ? *bits : 0u;
          uint32_t r[32];
          tmem_ld32(tmem_base + acc * BN + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
          if (c0 + 32 >= BN || nb + 32 >= q.N) {                 // last chunk read: hand the accumulator back before the math / store
            tc_fence_before();
            __syncwarp();
            if (elected) { if (CTAS == 2 && !leader) mbar_arrive_remote_leader(tmem_empty + acc); else mbar_arrive(tmem_empty + acc); }
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
          if (elected) bulk_wait_group_read<1>();                // the store issued two chunks ago has read this buffer (same lane issues and waits)
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
        if (elected) { if (CTAS == 2 && !leader) mbar_arrive_remote_leader(tmem_empty + acc); else mbar_arrive(tmem_empty + acc); }
      }
    }
    if (elected) bulk_wait_group_read<0>();                    // staging buffers must outlive the stores' reads
    tc_fence_before();
  }
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();                             // the peer may still be reading / the leader still issuing
  if (warp == 5) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_2cta(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
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

// 0: not decided yet (env PHC_TC5S_CTAS = 1 | 2).  Default 1: measured on the PPO shapes (profiles/gemm_microbench_r2c.log) the
// 128 x 128 one-CTA tiles reach 560-680 TF/s of tensor work, the 256 x 128 CTA-pair tiles 350-410 -- with N = 128 the pair
// saves no operand traffic worth its cross-CTA barrier round trips; it stays in the tree as an opt-in, parity-tested variant.
static int g_ctas = 0;
static int g_single_pass = 0;
static int g_tile = 0;        // 0: not decided yet (env PHC_TC5_TILE = 128 | 256; default 256 = gemm_tc5w.cu), else the tile width
static int g_sched = -1;      // -1: not decided yet (env PHC_TC5S_SCHED = static | dynamic; default dynamic), 0 static, 1 dynamic
// {next tile, CTAs done} pairs of the dynamic scheduler, zero at load time and put back to zero by every launch's last CTA; launches
// rotate through them so that two launches in flight on different streams do not share a pair
constexpr int SCHED_SLOTS = 64;
__device__ unsigned int g_sched_counters[SCHED_SLOTS][2];

}  // namespace smem_split
}  // namespace tc5
}  // namespace phc

extern "C" int phc_gemm_group_wide(const PhcGemmDesc* d, int32_t count, int32_t single_pass, int32_t dynamic_sched, void* stream);

static int validate_group(const PhcGemmDesc* d, int32_t count) {
  using namespace phc::tc5::smem_split;
  if (!d || count < 1 || count > MAX_PROBLEMS) { phc_set_error("phc_gemm_group: 1 <= count <= PHC_GEMM_GROUP_MAX problems"); return PHC_ERR_INVALID_ARG; }
  for (int i = 0; i < count; ++i) {
    const PhcGemmDesc& g = d[i];
    if (!g.A || !g.B || !g.C || g.M < 0 || g.N < 0 || g.K < 1) { phc_set_error("phc_gemm_group: bad problem (NULL operand or negative size)"); return PHC_ERR_INVALID_ARG; }
    if (g.M == 0 || g.N == 0) continue;
    if ((g.lda & 3) || (g.ldb & 3) || (g.ldc & 3)) { phc_set_error("phc_gemm_group: leading dimensions must be multiples of 4 floats (TMA strides are 16-byte multiples)"); return PHC_ERR_INVALID_ARG; }
    for (const void* p : {(const void*)g.A, (const void*)g.B, (const void*)g.C})
      if (reinterpret_cast<uintptr_t>(p) & 15) { phc_set_error("phc_gemm_group: A, B, C must be 16-byte aligned"); return PHC_ERR_INVALID_ARG; }
    const int ks = g.k_splits < 1 ? 1 : g.k_splits;
    if (g.act < 0 || g.act > PHC_ACT_MASK_BITS || ((g.act == PHC_ACT_SILU_BWD || g.act == PHC_ACT_MASK_BITS) && !g.aux)) { phc_set_error("phc_gemm_group: bad activation code"); return PHC_ERR_INVALID_ARG; }
    if (g.act >= PHC_ACT_RELU_BITS && g.aux && ((reinterpret_cast<uintptr_t>(g.aux) & 3) || g.ldaux < (g.N + 31) / 32)) { phc_set_error("phc_gemm_group: bit-mask aux needs ldaux >= ceil(N / 32) words"); return PHC_ERR_INVALID_ARG; }
    if (ks > 1 && (!g.accumulate || g.act || g.aux)) { phc_set_error("phc_gemm_group: split-K needs accumulate=1 and a linear epilogue"); return PHC_ERR_INVALID_ARG; }
    if (g.B_lo && (reinterpret_cast<uintptr_t>(g.B_lo) & 15)) { phc_set_error("phc_gemm_group: B_lo must be 16-byte aligned"); return PHC_ERR_INVALID_ARG; }
  }
  return PHC_OK;
}

extern "C" int phc_gemm_group(const PhcGemmDesc* d, int32_t count, void* stream) {
  using namespace phc::tc5::smem_split;
  const int vrc = validate_group(d, count);
  if (vrc != PHC_OK) return vrc;
  if (!g_ctas) { const char* v = getenv("PHC_TC5S_CTAS"); g_ctas = (v && v[0] == '2') ? 2 : 1; }
  if (g_sched < 0) { const char* v = getenv("PHC_TC5S_SCHED"); g_sched = (v && v[0] == 's') ? 0 : 1; }
  if (!g_tile) { const char* v = getenv("PHC_TC5_TILE"); g_tile = (v && v[0] == '1') ? 128 : 256; }
  const int ctas = g_ctas;
  if (g_tile == 256 && ctas == 1) return phc_gemm_group_wide(d, count, g_single_pass, g_sched, stream);      // gemm_tc5w.cu
  static Params P;      // host staging (launches are serialised by the caller's stream order; the struct is copied at launch)
  memset(&P.p, 0, sizeof(P.p));
  int tiles = 0, n = 0;
  for (int i = 0; i < count; ++i) {
    const PhcGemmDesc& g = d[i];
    if (g.M == 0 || g.N == 0) continue;
    int ks = g.k_splits < 1 ? 1 : g.k_splits;
    Prob& q = P.p[n];
    q.bias = g.bias; q.aux = g.aux; q.ldaux = g.ldaux; q.M = g.M; q.N = g.N; q.K = g.K; q.alpha = g.alpha; q.act = g.act;
    q.accumulate = g.accumulate ? 1 : 0; q.a_k = g.a_kmajor ? 1 : 0; q.b_k = g.b_kmajor ? 1 : 0; q.has_blo = g.B_lo ? 1 : 0;
    q.tiles_m = (g.M + BM * ctas - 1) / (BM * ctas);
    q.tiles_n = (g.N + BN - 1) / BN;
    q.kb_total = (g.K + BK - 1) / BK;
    if (ks > q.kb_total) ks = q.kb_total;
    q.k_splits = ks;
    q.kb_per = (q.kb_total + ks - 1) / ks;
    q.tile_begin = tiles;
    q.tile_count = q.tiles_m * q.tiles_n * ks;
    tiles += q.tile_count;
    const uint32_t brows = BN / ctas;
    bool ok = q.a_k ? make_map(&P.tmA[n], g.A, g.lda, (uint64_t)g.K, (uint64_t)g.M, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B)
                    : make_map(&P.tmA[n], g.A, g.lda, (uint64_t)g.M, (uint64_t)g.K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    ok = ok && (q.b_k ? make_map(&P.tmB[n], g.B, g.ldb, (uint64_t)g.K, (uint64_t)g.N, 32, brows, CU_TENSOR_MAP_SWIZZLE_128B)
                      : make_map(&P.tmB[n], g.B, g.ldb, (uint64_t)g.N, (uint64_t)g.K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    if (g.B_lo)
      ok = ok && (q.b_k ? make_map(&P.tmBlo[n], g.B_lo, g.ldb, (uint64_t)g.K, (uint64_t)g.N, 32, brows, CU_TENSOR_MAP_SWIZZLE_128B)
                        : make_map(&P.tmBlo[n], g.B_lo, g.ldb, (uint64_t)g.N, (uint64_t)g.K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    ok = ok && make_map(&P.tmC[n], g.C, g.ldc, (uint64_t)g.N, (uint64_t)g.M, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) { phc_set_error("phc_gemm_group: cuTensorMapEncodeTiled failed"); return PHC_ERR_CUDA; }
    ++n;
  }
  if (n == 0) return PHC_OK;
  P.count = n; P.total_tiles = tiles; P.single_pass = g_single_pass;
  P.sched = nullptr;
  if (g_sched == 1 && ctas == 1) {
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
  const int units = num_sms / ctas;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((unsigned)((tiles < units ? tiles : units) * ctas));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (ctas == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = Cfg<2>::SMEM;
  } else {
    cfg.attrs = nullptr; cfg.numAttrs = 0;
    cfg.dynamicSmemBytes = Cfg<1>::SMEM;
  }
  using Kernel = void (*)(const Params);
  static const Kernel kernels[4] = {gemm_tc5s_kernel<1, false>, gemm_tc5s_kernel<1, true>, gemm_tc5s_kernel<2, false>, gemm_tc5s_kernel<2, true>};
  static bool smem_set[4] = {false, false, false, false};
  const int ki = (ctas == 2 ? 2 : 0) + (g_single_pass ? 1 : 0);
  if (!smem_set[ki]) {
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernels[ki]), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    if (e != cudaSuccess) return phc_check_cuda(e, "cudaFuncSetAttribute(gemm_tc5s)");
    smem_set[ki] = true;
  }
  e = cudaLaunchKernelEx(&cfg, kernels[ki], P);
  if (e != cudaSuccess) return phc_check_cuda(e, "cudaLaunchKernelEx(gemm_tc5s)");
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "gemm_tc5s_kernel launch");
}

extern "C" int phc_gemm_tc5s(const float* A, int64_t lda, int32_t a_kmajor, const float* B, int64_t ldb, int32_t b_kmajor, float* C,
                             int64_t ldc, int32_t M, int32_t N, int32_t K, float alpha, const float* bias, int32_t act, float* aux,
                             int64_t ldaux, int32_t accumulate, int32_t k_splits, void* stream) {
  PhcGemmDesc d;
  d.A = A; d.lda = lda; d.a_kmajor = a_kmajor; d.B = B; d.ldb = ldb; d.b_kmajor = b_kmajor; d.C = C; d.ldc = ldc;
  d.M = M; d.N = N; d.K = K; d.alpha = alpha; d.bias = bias; d.act = act; d.aux = aux; d.ldaux = ldaux;
  d.accumulate = accumulate; d.k_splits = k_splits; d.B_lo = nullptr;
  return phc_gemm_group(&d, 1, stream);
}

namespace phc { namespace tc5 { namespace smem_split {
__global__ void split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) lo[i] = split_lo(x[i]);
}
}}}

extern "C" int phc_split_lo(const float* x, float* lo, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!x || !lo))) { phc_set_error("phc_split_lo: NULL buffer"); return PHC_ERR_INVALID_ARG; }
  if (n == 0) return PHC_OK;
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  phc::tc5::smem_split::split_lo_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(x, lo, n);
  phc_count_launches(1);
  return phc_check_cuda(cudaGetLastError(), "split_lo_kernel launch");
}

extern "C" int phc_gemm_set_precision(int32_t mode) {      // PHC_GEMM_FP32_3XTF32 (default) | PHC_GEMM_TF32_SINGLE_PASS
  if (mode != PHC_GEMM_FP32_3XTF32 && mode != PHC_GEMM_TF32_SINGLE_PASS) { phc_set_error("phc_gemm_set_precision: unknown mode"); return PHC_ERR_INVALID_ARG; }
  phc::tc5::smem_split::g_single_pass = mode == PHC_GEMM_TF32_SINGLE_PASS;
  return PHC_OK;
}

extern "C" int phc_gemm_tc5s_set_tile(int32_t width) {     // 128: this file's 128 x 128 x 32 tiles; 256: gemm_tc5w.cu's 128 x 256 x 16; 0 = default
  if (width != 0 && width != 128 && width != 256) { phc_set_error("phc_gemm_tc5s_set_tile: 0, 128 or 256"); return PHC_ERR_INVALID_ARG; }
  phc::tc5::smem_split::g_tile = width;
  return PHC_OK;
}

extern "C" int phc_gemm_tc5s_set_sched(int32_t mode) {     // tile order of the one-CTA kernel: 0 static striding, 1 dynamic (global counter), -1 default
  if (mode < -1 || mode > 1) { phc_set_error("phc_gemm_tc5s_set_sched: -1, 0 or 1"); return PHC_ERR_INVALID_ARG; }
  phc::tc5::smem_split::g_sched = mode;
  return PHC_OK;
}

extern "C" int phc_gemm_tc5s_set_ctas(int32_t ctas) {      // A/B switch for tests and tools (1 or 2; 0 = back to the environment default)
  if (ctas < 0 || ctas > 2) { phc_set_error("phc_gemm_tc5s_set_ctas: 0, 1 or 2"); return PHC_ERR_INVALID_ARG; }
  phc::tc5::smem_split::g_ctas = ctas;
  return PHC_OK;
}
