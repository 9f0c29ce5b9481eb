// Fused ResidualUnit on a CTA PAIR (tcgen05 cta_group::2), C = 64 in split precision, sm_100a.
//
//   y = x + W1 * snake2( W7 (*) snake1(x) + b7 ) + b1          (vq/module.py:74-89)
//
// Why a pair.  ru_persist.cu keeps both weight images (bf16 hi + lo: 131 KB at C = 64) resident in every SM, which
// leaves room for ONE activation slot: staging a tile and the K-tap MMA chain of the previous one serialise, and the
// tile period is their sum (phase trace: 3.6 k + 4.5 k of 8.8 k cycles).  A cluster of two CTAs issues ONE
// tcgen05.mma.cta_group::2 of M = 256: each SM contributes the 128 rows of ITS tile as the A operand and HALF of the
// B operand's rows, so
//   * every SM holds half of the weights (94 KB stacked / 65 KB plain) -> a second activation slot fits and staging
//     overlaps the MMA chain of the previous tile;
//   * the tensor core fetches half of the B rows per SM and MMA -- the kernels are bound by shared-memory bandwidth,
//     most of it operand fetches.
//
// Roles per CTA are those of ru_persist.cu (LOAD x8 / MMA / MID x4 / STORE x4 warps); only the LEADER CTA's MMA warp
// issues.  Cross-CTA protocol (every barrier exists at the same shared-memory offset in both CTAs):
//   producers -> MMA   (slab staged, A2 tile written, accumulator drained): the warps of BOTH CTAs arrive on the
//                      LEADER's barrier (mbarrier.arrive.release.cluster on the mapa-translated address);
//   MMA -> consumers   (slot free, accumulator ready): tcgen05.commit ... multicast::cluster to the same barrier in
//                      both CTAs; everybody waits on its local copy.
// Tiles are handed out in pairs (2i, 2i+1): rank r of the pair takes tile 2i + r; an odd tile count gives rank 1 one
// phantom tile (zero slab, no stores) so that both CTAs run the same number of MMA rounds.
//
// Weight images (host-packed per rank, ops.pack_pair_weights):
//   STACK  K-tap conv as two MMAs per (tap, 16-channel group): a_hi x [w_hi | w_lo] (N = 2C: rank 0 holds the w_hi rows,
//          rank 1 the w_lo rows) and a_lo x w_hi (N = C: rank r holds rows [r*C/2, (r+1)*C/2)).
//   plain  three MMAs of N = C (hi*hi, hi*lo, lo*hi), rank r holding half of the rows of w_hi and of w_lo -- used when
//          the stacked image does not leave room for two activation slots (dilation 9).
//   The 1x1 conv always runs plain.
//
// Measured (B200, 8 x 30 s clips, C = 64, split precision; ru_persist in brackets): dilation 1 / 3 (stacked) 363 / 366 us
// [442 / 438], dilation 9 (plain) 375 us [450]: 20 % faster.  The first build was no faster than ru_persist (472 / 466 /
// 441 us): its cross-CTA arrives and waits asked for cluster-scope release / acquire, which compiles to MEMBAR.ALL.GPU +
// CCTL.IVALL -- an L1 invalidation per hand-off that the loaders' cached activation reads paid for.  The hand-offs only
// carry shared-memory data, for which the default CTA-scope semantics are enough (tc_common.cuh, BC_PAIR_STRONG).
#include "common.cuh"
#include "tc_common.cuh"

#ifndef BC_RU_ILP     // 1: branch-free staging batches (what made the conv_stream producers 6-15 % faster); measured 1-2 % SLOWER here -> 0: one guarded block per item
#define BC_RU_ILP 0
#endif
namespace {
using namespace bc::tc;

constexpr int BM = 128;
constexpr int LOAD_WARPS = 8;
constexpr int LOAD_THREADS = LOAD_WARPS * 32;
constexpr int LOAD_GROUPS = 2;                      // loader groups alternate tiles: two tiles' HBM loads in flight
constexpr int GROUP_WARPS = LOAD_WARPS / LOAD_GROUPS;
constexpr int MMA_WARP = LOAD_WARPS;
constexpr int MID_WARP0 = MMA_WARP + 4;             // the MMA warp shares its warpgroup with three idle warps (setmaxnreg works on warpgroups)
constexpr int MID_WARPS = 4;
constexpr int EPI_WARP0 = MID_WARP0 + MID_WARPS;
constexpr int RP_WARPS = EPI_WARP0 + 4;
constexpr int RP_THREADS = RP_WARPS * 32;
constexpr int REG_MMA = 56, REG_LOAD = 112, REG_STORE = 104;
constexpr int LD_BATCH = 5;
constexpr int EPI_LD = 36;
constexpr size_t STAGE_BYTES = (size_t)4 * 32 * EPI_LD * sizeof(float);

struct RpParams {
  const float* x;
  float* y;
  const uint8_t* w7;       // [2 ranks][w7_rank_bytes]
  const uint8_t* w1;       // [2 ranks][w1_rank_bytes]
  const float* b7;
  const float* b1;
  const float* sa1;
  const float* sib1;
  const float* sa2;
  const float* sib2;
  int B, T, C, K, dil, pad_left;
  int slab_rows, tiles_per_item, total_tiles;
  uint32_t w7_rank_bytes, w1_rank_bytes;
};

enum { B_A_FULL = 0, B_A_EMPTY = 2, B_ACC1_FULL = 4, B_ACC1_EMPTY = 6, B_A2_FULL = 8, B_A2_EMPTY = 9,
       B_ACC2_FULL = 10, B_ACC2_EMPTY = 12, B_W_FULL = 14, N_BARS = 15 };

// r[0 .. 32) += the 32 accumulator columns at `taddr` (the a_hi * w_lo half), 16 columns at a time to bound registers
__device__ __forceinline__ void add_lo_half32(uint32_t taddr, uint32_t r[32]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t t[32];
    tmem_load(taddr + 16u * h, 2, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float x0, x1;
      unpack2(add2(pack2(__uint_as_float(r[16 * h + 2 * e]), __uint_as_float(r[16 * h + 2 * e + 1])),
                   pack2(__uint_as_float(t[2 * e]), __uint_as_float(t[2 * e + 1]))), x0, x1);
      r[16 * h + 2 * e] = __float_as_uint(x0);
      r[16 * h + 2 * e + 1] = __float_as_uint(x1);
    }
  }
}

__host__ __device__ inline uint32_t rp_plane_bytes(int slab_rows) {
  uint32_t b = (uint32_t)slab_rows * 16u;
  while (b % 128u != 16u) b += 16u;
  return b;
}

// (item, tile-in-item) of tile index `tile` (may be == total_tiles: the phantom tile of an odd count)
struct TileAt {
  int b, tt;
  bool valid;
};
__device__ __forceinline__ TileAt tile_at(int tile, const RpParams& p) {
  TileAt t;
  t.valid = tile < p.total_tiles;
  const int tl = t.valid ? tile : 0;
  t.b = tl / p.tiles_per_item;
  t.tt = tl - t.b * p.tiles_per_item;
  return t;
}

template <bool STACK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RP_THREADS, 1) ru_pair_kernel(const RpParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int C = 64, GROUPS = 4, planes = 8, NSLOT = 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  constexpr uint32_t ACC1_COLS = STACK ? 2 * C : C;            // TMEM columns of one K-tap accumulator stage
  constexpr uint32_t ACC2_BASE = 2u * ACC1_COLS;
  constexpr uint32_t TMEM_COLS = 512;
  const uint32_t plane_bytes = rp_plane_bytes(p.slab_rows);
  const uint32_t a_split = planes * plane_bytes;
  const uint32_t a_slot = (a_split * 2u + 127u) & ~127u;
  constexpr uint32_t a2_plane = BM * 16u;
  constexpr uint32_t a2_split = planes * a2_plane;
  constexpr uint32_t a2_slot = a2_split * 2u;

  uint8_t* sW7 = smem_raw;
  uint8_t* sW1 = sW7 + p.w7_rank_bytes;
  uint8_t* sA = sW1 + p.w1_rank_bytes;
  uint8_t* sA2 = sA + (size_t)a_slot * NSLOT;
  float* sPar = reinterpret_cast<float*>(sA2 + a2_slot);        // b7 | sa2 | sib2 | b1
  float* sStage = sPar + 4 * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 4 * 32 * EPI_LD);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_A_FULL + s), 2 * GROUP_WARPS);            // both CTAs' loader groups
      mbar_init(BAR(B_A_EMPTY + s), 1);
      mbar_init(BAR(B_ACC1_FULL + s), 1);
      mbar_init(BAR(B_ACC1_EMPTY + s), 2 * MID_WARPS);
      mbar_init(BAR(B_ACC2_FULL + s), 1);
      mbar_init(BAR(B_ACC2_EMPTY + s), 2 * 4);
    }
    mbar_init(BAR(B_A2_FULL), 2 * MID_WARPS);
    mbar_init(BAR(B_A2_EMPTY), 1);
    mbar_init(BAR(B_W_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t wb = p.w7_rank_bytes + p.w1_rank_bytes;
    mbar_expect_tx(BAR(B_W_FULL), wb);
    const uint8_t* s7 = p.w7 + (size_t)rank * p.w7_rank_bytes;
    for (uint32_t off = 0; off < p.w7_rank_bytes; off += 32768u)
      bulk_g2s_notx(smem_u32(sW7) + off, s7 + off, min(32768u, p.w7_rank_bytes - off), BAR(B_W_FULL));
    const uint8_t* s1 = p.w1 + (size_t)rank * p.w1_rank_bytes;
    bulk_g2s_notx(smem_u32(sW1), s1, p.w1_rank_bytes, BAR(B_W_FULL));
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < C; i += RP_THREADS) {
    sPar[i] = __ldg(p.b7 + i);
    sPar[C + i] = __ldg(p.sa2 + i);
    sPar[2 * C + i] = __ldg(p.sib2 + i);
    sPar[3 * C + i] = __ldg(p.b1 + i);
  }
  if (tid == 0) mbar_wait(BAR(B_W_FULL), 0);      // this CTA's half of the weights has landed ...
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // ... and the peer's: barriers initialised, TMEM allocated, weights resident in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pair schedule: pair `pi` of `npairs` takes the tile pairs pi, pi + npairs, ...; this CTA the tile 2 * pair + rank
  const int pi = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int total_pairs = (p.total_tiles + 1) / 2;
  int n_my = 0;
  for (int q = pi; q < total_pairs; q += npairs) ++n_my;

  if (warp < LOAD_WARPS) {
    // ======================= LOAD =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_LOAD));
    const int items = planes * p.slab_rows;
    constexpr int gthreads = LOAD_THREADS / LOAD_GROUPS;
    const int grp = tid / gthreads;
    const int gtid = tid - grp * gthreads;
    const int pl = gtid & (planes - 1);
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.sa1 + pl * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.sa1 + pl * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.sib1 + pl * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.sib1 + pl * 8) + 1);
    for (int it = grp; it < n_my; it += LOAD_GROUPS) {       // group g always fills slot g
      const int slot = it & 1, use = it >> 1;
      const TileAt ta = tile_at(2 * (pi + it * npairs) + (int)rank, p);
      const int g0 = ta.tt * BM - p.pad_left;
      const float* xcol = p.x + (size_t)ta.b * p.T * C + pl * 8;
      uint8_t* dstA = sA + (size_t)slot * a_slot + (size_t)pl * plane_bytes;
      bool waited = false;
      for (int i0 = gtid; i0 < items; i0 += gthreads * LD_BATCH) {
        float4 lo4[LD_BATCH], hi4[LD_BATCH];
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          const int i = i0 + j * gthreads;
          const int g = g0 + (i >> 3);
          if (ta.valid && i < items && g >= 0 && g < p.T) {
            const float4* src = reinterpret_cast<const float4*>(xcol + (size_t)g * C);
            lo4[j] = __ldg(src);
            hi4[j] = __ldg(src + 1);
          } else {
            lo4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            hi4[j] = lo4[j];
          }
        }
        if (!waited) {  // the global loads above are already in flight while we wait for the slot
          mbar_wait(BAR(B_A_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
          waited = true;
        }
#if BC_RU_ILP
        // arithmetic of the whole batch first, branch-free (items beyond the slab hold zeros), then the predicated stores:
        // LD_BATCH independent SnakeBeta -> split chains interleave instead of one guarded block per item
        uint4 hq[LD_BATCH], lq[LD_BATCH];
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          float v[8] = {lo4[j].x, lo4[j].y, lo4[j].z, lo4[j].w, hi4[j].x, hi4[j].y, hi4[j].z, hi4[j].w};
          snake8<2>(v, a0, a1, b0, b1);
          split8<2>(v, hq[j], lq[j]);
        }
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          const int i = i0 + j * gthreads;
          if (i < items) {
            *reinterpret_cast<uint4*>(dstA + (size_t)(i >> 3) * 16) = hq[j];
            if (2 == 2) *reinterpret_cast<uint4*>(dstA + (size_t)(i >> 3) * 16 + a_split) = lq[j];
          }
        }
#else
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          const int i = i0 + j * gthreads;
          if (i < items) {
            float v[8] = {lo4[j].x, lo4[j].y, lo4[j].z, lo4[j].w, hi4[j].x, hi4[j].y, hi4[j].z, hi4[j].w};
            snake8<2>(v, a0, a1, b0, b1);
            split_store<2>(v, dstA + (size_t)(i >> 3) * 16, a_split);
          }
        }
#endif
      }
      if (!waited) mbar_wait(BAR(B_A_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(BAR(B_A_FULL + slot));
    }
  } else if (warp < MID_WARP0) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_MMA));
   if (warp == MMA_WARP && rank == 0) {
    // ======================= MMA issue (leader CTA only) =======================
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t uW7 = smem0, uW1 = uW7 + p.w7_rank_bytes, uA = uW1 + p.w1_rank_bytes, uA2 = uA + a_slot * NSLOT;
    const uint32_t hi_d = desc_hi(128u);
    const uint32_t a_g = (2u * plane_bytes) >> 4, a_k = (uint32_t)p.dil, a_sp = a_split >> 4;
    const uint32_t a2_g = (2u * a2_plane) >> 4;
    // per-rank weight regions.  STACK: X = [K*GROUPS][2 planes][C rows] then Y = [K*GROUPS][2][C/2];
    // plain: Hh = [K*GROUPS][2][C/2] (this rank's rows of w_hi) then Hl (of w_lo).  1x1: Hh [GROUPS][2][C/2], Hl.
    const uint32_t half_g = ((uint32_t)(C / 2) * 32u) >> 4;              // one (tap, group) block of C/2 rows, 16-byte units
    const uint32_t full_g = ((uint32_t)C * 32u) >> 4;
    const uint32_t x_lo0 = desc_lo(uW7, (uint32_t)C * 16u);
    const uint32_t y_lo0 = desc_lo(uW7 + (uint32_t)p.K * GROUPS * C * 32u, (uint32_t)(C / 2) * 16u);
    const uint32_t hh_lo0 = desc_lo(uW7, (uint32_t)(C / 2) * 16u);
    const uint32_t hl_lo0 = desc_lo(uW7 + (uint32_t)p.K * GROUPS * (C / 2) * 32u, (uint32_t)(C / 2) * 16u);
    const uint32_t w1h_lo0 = desc_lo(uW1, (uint32_t)(C / 2) * 16u);
    const uint32_t w1l_lo0 = desc_lo(uW1 + (uint32_t)GROUPS * (C / 2) * 32u, (uint32_t)(C / 2) * 16u);
    const uint32_t idesc_wide = idesc_bf16_m256(2 * C), idesc_n = idesc_bf16_m256(C);
    for (int it = 0; it <= n_my; ++it) {
      if (it < n_my) {  // K-tap conv of tile pair `it`
        const int slot = it & 1, use = it >> 1, as = it & 1, ause = it >> 1;
        mbar_wait_cluster(BAR(B_A_FULL + slot), (uint32_t)(use & 1));
        mbar_wait_cluster(BAR(B_ACC1_EMPTY + as), (uint32_t)((ause & 1) ^ 1));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)as * ACC1_COLS;
        const uint32_t a_lo0 = desc_lo(uA + (uint32_t)slot * a_slot, plane_bytes);
        if (elect_one()) {
          // taps as a rolled loop (descriptor offsets advance incrementally): fully unrolled, the 28 loop-invariant
          // offset pairs get hoisted out of the tile loop and spill past the MMA warpgroup's 56 registers
          uint32_t a_k0 = a_lo0, kg_full = 0, kg_half = 0;
#pragma unroll 1
          for (int k = 0; k < p.K; ++k, a_k0 += a_k) {
#pragma unroll
            for (int g = 0; g < GROUPS; ++g, kg_full += full_g, kg_half += half_g) {
              const uint32_t a_lo = a_k0 + (uint32_t)g * a_g;
              const uint32_t first = (uint32_t)(k | g);
              if (STACK) {
                mma2_bf16_rt(d, a_lo, x_lo0 + kg_full, hi_d, hi_d, idesc_wide, first);                       // a_hi * [w_hi | w_lo]
                mma2_bf16_raw<true>(d, a_lo + a_sp, y_lo0 + kg_half, hi_d, hi_d, idesc_n);                    // a_lo * w_hi
              } else {
                mma2_bf16_rt(d, a_lo, hh_lo0 + kg_half, hi_d, hi_d, idesc_n, first);                          // a_hi * w_hi
                mma2_bf16_raw<true>(d, a_lo, hl_lo0 + kg_half, hi_d, hi_d, idesc_n);                          // a_hi * w_lo
                mma2_bf16_raw<true>(d, a_lo + a_sp, hh_lo0 + kg_half, hi_d, hi_d, idesc_n);                   // a_lo * w_hi
              }
            }
          }
          umma_commit_pair(BAR(B_A_EMPTY + slot));
          umma_commit_pair(BAR(B_ACC1_FULL + as));
        }
        __syncwarp();
      }
      if (it >= 1) {  // 1x1 conv of tile pair `it - 1`
        const int j = it - 1;
        const int as = j & 1, ause = j >> 1;
        mbar_wait_cluster(BAR(B_A2_FULL), (uint32_t)(j & 1));
        mbar_wait_cluster(BAR(B_ACC2_EMPTY + as), (uint32_t)((ause & 1) ^ 1));
        tc_fence_after();
        const uint32_t d = tmem_base + ACC2_BASE + (uint32_t)(as * C);
        const uint32_t a_lo0 = desc_lo(uA2, a2_plane);
        if (elect_one()) {
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) {
            const uint32_t a_lo = a_lo0 + (uint32_t)g * a2_g;
            if (g == 0) mma2_bf16_raw<false>(d, a_lo, w1h_lo0 + (uint32_t)g * half_g, hi_d, hi_d, idesc_n);
            else        mma2_bf16_raw<true>(d, a_lo, w1h_lo0 + (uint32_t)g * half_g, hi_d, hi_d, idesc_n);
            mma2_bf16_raw<true>(d, a_lo, w1l_lo0 + (uint32_t)g * half_g, hi_d, hi_d, idesc_n);
            mma2_bf16_raw<true>(d, a_lo + (a2_split >> 4), w1h_lo0 + (uint32_t)g * half_g, hi_d, hi_d, idesc_n);
          }
          umma_commit_pair(BAR(B_A2_EMPTY));
          umma_commit_pair(BAR(B_ACC2_FULL + as));
        }
        __syncwarp();
      }
    }
   }
  } else if (warp < EPI_WARP0) {
    // ======================= MID: acc1 -> snake2 -> bf16 A2 tile =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    for (int it = 0; it < n_my; ++it) {
      const int as = it & 1, ause = it >> 1;
      mbar_wait(BAR(B_ACC1_FULL + as), (uint32_t)(ause & 1));
      tc_fence_after();
      mbar_wait(BAR(B_A2_EMPTY), (uint32_t)((it & 1) ^ 1));
      uint8_t* dst = sA2 + (size_t)row * 16;
      const uint32_t taddr = tmem_base + (uint32_t)as * ACC1_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        tmem_load32(taddr + (uint32_t)c0, r);
        if (STACK) add_lo_half32(taddr + (uint32_t)(C + c0), r);   // + a_hi * w_lo
        if (c0 + 32 >= C) {  // this warp's share of the accumulator is read: hand it back before the math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(BAR(B_ACC1_EMPTY + as));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + 8 * j;
          const float4 bi0 = *reinterpret_cast<const float4*>(sPar + c), bi1 = *reinterpret_cast<const float4*>(sPar + c + 4);
          const float4 s0 = *reinterpret_cast<const float4*>(sPar + C + c), s1 = *reinterpret_cast<const float4*>(sPar + C + c + 4);
          const float4 i0 = *reinterpret_cast<const float4*>(sPar + 2 * C + c), i1 = *reinterpret_cast<const float4*>(sPar + 2 * C + c + 4);
          float v[8];
          acc_bias8(r + 8 * j, bi0, bi1, v);
          snake8<2>(v, s0, s1, i0, i1);
          split_store<2>(v, dst + (size_t)(c / 8) * a2_plane, a2_split);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(BAR(B_A2_FULL));
    }
  } else {
    // ======================= STORE: acc2 + b1 + x -> y =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_STORE));
    const int q = warp & 3;
    float* sT = sStage + (size_t)(warp - EPI_WARP0) * (32 * EPI_LD);
    const int crow = lane >> 3, cchunk = (lane & 7) * 4;     // coalesced mapping: rows crow + 4*i, 4 floats at cchunk
    for (int it = 0; it < n_my; ++it) {
      const TileAt ta = tile_at(2 * (pi + it * npairs) + (int)rank, p);
      const int as = it & 1, ause = it >> 1;
      const int trow0 = ta.tt * BM + q * 32;                 // first row of this warp's block
      const size_t off0 = ((size_t)ta.b * p.T + trow0 + crow) * C + cchunk;
      const float* rp = p.x + off0;
      float* yp = p.y + off0;
      const size_t istep = (size_t)4 * C;
      const int rows_ok = ta.valid ? p.T - trow0 - crow : 0;  // row 4*i of this lane is valid iff 4*i < rows_ok
      const uint32_t taddr = tmem_base + ACC2_BASE + (uint32_t)(as * C) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 32) {
        float4 res4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          res4[i] = 4 * i < rows_ok ? __ldg(reinterpret_cast<const float4*>(rp + c0 + i * istep)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 == 0) {
          mbar_wait(BAR(B_ACC2_FULL + as), (uint32_t)(ause & 1));
          tc_fence_after();
        }
        uint32_t r[32];
        tmem_load32(taddr + (uint32_t)c0, r);
        if (c0 + 32 >= C) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(BAR(B_ACC2_EMPTY + as));
        }
        float* own = sT + lane * EPI_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(sPar + 3 * C + c0 + 4 * j);
          *reinterpret_cast<float4*>(own + 4 * j) = acc_bias4(r + 4 * j, bb);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(sT + (4 * i + crow) * EPI_LD + cchunk);
          if (4 * i < rows_ok) __stcs(reinterpret_cast<float4*>(yp + c0 + i * istep), add4(v, res4[i]));
        }
        __syncwarp();
      }
    }
  }
#undef BAR
  // ---- teardown: nobody frees TMEM / exits while the peer may still be reading or signalling ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

struct RpPlan {
  bool stack;
  uint32_t w7_rank_bytes, w1_rank_bytes;
  size_t smem;
};

bool rp_plan(int C, int K, int dilation, int precision, RpPlan* pl) {
  if (C != 64 || K < 1 || K > 7 || precision != BC_PREC_BF16X3) return false;
  const int slab_rows = (BM - 1) + (K - 1) * dilation + 1;
  const size_t a_slot = ((size_t)2 * (C / 8) * rp_plane_bytes(slab_rows) + 127) & ~size_t(127);
  const size_t a2_slot = (size_t)2 * (C / 8) * BM * 16;
  const size_t fixed = 2 * a_slot + a2_slot + 4 * C * sizeof(float) + STAGE_BYTES + N_BARS * 8 + 64;
  const size_t w1 = (size_t)2 * (C / 16) * 2 * (C / 2) * 16;                     // Hh + Hl
  const size_t w7_stack = (size_t)K * (C / 16) * 2 * (C + C / 2) * 16;           // X + Y
  const size_t w7_plain = (size_t)K * (C / 16) * 2 * (C / 2) * 16 * 2;           // Hh + Hl
  if ((size_t)rp_plane_bytes(slab_rows) * 2 >= (1u << 18)) return false;
  pl->w1_rank_bytes = (uint32_t)w1;
  if (fixed + w1 + w7_stack <= 227 * 1024) { pl->stack = true; pl->w7_rank_bytes = (uint32_t)w7_stack; }
  else if (fixed + w1 + w7_plain <= 227 * 1024) { pl->stack = false; pl->w7_rank_bytes = (uint32_t)w7_plain; }
  else return false;
  pl->smem = fixed + w1 + pl->w7_rank_bytes;
  return true;
}

}  // namespace

namespace bc {

// 0 = no pair plan; 1 = plain weight image; 2 = stacked
int ru_pair_layout(int C, int K, int dilation, int precision) {
  if (!policy().ru_pair) return 0;
  RpPlan pl;
  if (!rp_plan(C, K, dilation, precision, &pl)) return 0;
  return pl.stack ? 2 : 1;
}

int resunit_pair_fwd(const float* x, const void* w7_pair, const float* b7, const float* sa1, const float* sib1,
                     const void* w1_pair, const float* b1, const float* sa2, const float* sib2, float* y, int B, int T,
                     int C, int K, int dilation, int pad_left, int precision, cudaStream_t st) {
  RpPlan pl;
  if (!rp_plan(C, K, dilation, precision, &pl))
    return fail(BC_EUNSUPPORTED, "resunit(pair): C=%d K=%d dil=%d precision=%d not supported", C, K, dilation, precision);
  RpParams p;
  p.x = x; p.y = y; p.w7 = reinterpret_cast<const uint8_t*>(w7_pair); p.w1 = reinterpret_cast<const uint8_t*>(w1_pair);
  p.b7 = b7; p.b1 = b1; p.sa1 = sa1; p.sib1 = sib1; p.sa2 = sa2; p.sib2 = sib2;
  p.B = B; p.T = T; p.C = C; p.K = K; p.dil = dilation; p.pad_left = pad_left;
  p.slab_rows = (BM - 1) + (K - 1) * dilation + 1;
  p.tiles_per_item = (T + BM - 1) / BM;
  const long long total = (long long)p.tiles_per_item * B;
  if (total > 2147483647ll - 2) return fail(BC_EINVAL, "resunit(pair): too many tiles");
  p.total_tiles = (int)total;
  p.w7_rank_bytes = pl.w7_rank_bytes;
  p.w1_rank_bytes = pl.w1_rank_bytes;
  void (*kern)(const RpParams) = pl.stack ? ru_pair_kernel<true> : ru_pair_kernel<false>;
  static bool configured[64][2] = {{false}};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 0 || dev >= 64 || !configured[dev][pl.stack ? 1 : 0]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_check(e, "cudaFuncSetAttribute(ru_pair)");
    if (dev >= 0 && dev < 64) configured[dev][pl.stack ? 1 : 0] = true;
  }
  const long long pairs = (total + 1) / 2;
  const int max_pairs = sms / 2;
  const int grid = 2 * (int)(pairs < max_pairs ? pairs : max_pairs);
  kern<<<grid, RP_THREADS, pl.smem, st>>>(p);
  BC_LAUNCH_CHECK("ru_pair_kernel");
  return BC_OK;
}

}  // namespace bc
