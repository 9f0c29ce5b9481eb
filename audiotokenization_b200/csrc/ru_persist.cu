// Persistent, warp-specialised fused ResidualUnit for the high-rate layers (C <= 64), sm_100a.
//
//   y = x + W1 * snake2( W7 (*) snake1(x) + b7 ) + b1          (vq/module.py:74-89)
//
// One CTA per SM stays resident for the whole launch: both weight images are bulk-copied into shared
// memory ONCE, then the CTA walks its share of the (item, 128-step) tiles through a 4-stage pipeline
// whose stages run on different warps and overlap across consecutive tiles:
//
//   warps 0-7   LOAD   x (fp32, HBM) -> snake1 -> bf16 hi[/lo] -> K-major slab in smem            (a_full)
//   warp  8     MMA    conv7: K x C/16 [x3] tcgen05.mma into acc1[s] (TMEM)                         (acc1_full)
//                      conv1: C/16 [x3] tcgen05.mma on the re-quantised tile into acc2[s]          (acc2_full)
//   warps 9-16  MID    acc1 -> +b7 -> snake2 -> bf16 hi[/lo] -> smem A2 tile                        (a2_full)
//   warps 17-20 STORE  acc2 -> +b1 -> +x (residual) -> y (fp32, HBM)
// The 8 LOAD warps form two groups that take alternate tiles, so the HBM latency of tile i+1 is
// hidden behind the activation math of tile i.
//
// Stage hand-offs are mbarriers; the MMA warp issues conv7 of tile i+1 before conv1 of tile i so the
// tensor core never waits for the MID stage.  The activation makes exactly one HBM read (plus the
// L2-resident residual re-read) and one HBM write; the intermediate never leaves the SM.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

#ifndef BC_RU_ILP     // 1: branch-free staging batches (what made the conv_stream producers 6-15 % faster); measured 1-2 % SLOWER here -> 0: one guarded block per item
#define BC_RU_ILP 0
#endif
namespace {
using namespace bc::tc;

constexpr int BM = 128;
#ifndef BC_RU_LOAD_WARPS
#define BC_RU_LOAD_WARPS 8
#endif
#ifndef BC_RU_LD_BATCH
#define BC_RU_LD_BATCH 5
#endif
constexpr int LOAD_WARPS = BC_RU_LOAD_WARPS;
constexpr int LOAD_THREADS = LOAD_WARPS * 32;
constexpr int LOAD_GROUPS = 2;                      // loader groups alternate tiles: two tiles' HBM loads in flight
constexpr int GROUP_WARPS = LOAD_WARPS / LOAD_GROUPS;
constexpr int GROUP_THREADS = GROUP_WARPS * 32;
constexpr int MMA_WARP = LOAD_WARPS;
constexpr int MID_WARP0 = MMA_WARP + 4;             // the MMA warp shares its warpgroup with three idle warps (setmaxnreg works on warpgroups)
#ifndef BC_RU_MID_WARPS
#define BC_RU_MID_WARPS 4
#endif
constexpr int MID_WARPS = BC_RU_MID_WARPS;   // 4 (one per TMEM lane quarter) measured 5-13 % faster than 8: 544 threads leave 96+ registers per thread, no spills
constexpr int EPI_WARP0 = MID_WARP0 + MID_WARPS;    // 4 warps
constexpr int RU_WARPS = EPI_WARP0 + 4;
constexpr int RU_THREADS = RU_WARPS * 32;
static_assert(LOAD_WARPS % 4 == 0 && MID_WARPS % 4 == 0, "roles must fill whole warpgroups");
// Register budget after launch (96 per thread for 20 warps): the MMA warpgroup keeps 40, the loaders (raw tile in
// registers while they wait for the slot) take 120, the store warps 104.  (5 x 96 = 2 x 120 + 40 + 96 + 104)
constexpr int REG_MMA = 56, REG_LOAD = 112, REG_STORE = 104;
constexpr int LD_BATCH = BC_RU_LD_BATCH;
constexpr int EPI_LD = 36;                           // staging row stride in floats (32 + 4: conflict-free 16-byte accesses both ways)
constexpr size_t STAGE_BYTES = (size_t)4 * 32 * EPI_LD * sizeof(float);   // one [32 rows][32 + 4] block per store warp

struct RuParams {
  const float* x;
  float* y;
  const uint4* w7;
  const uint4* w1;
  const float* b7;
  const float* b1;
  const float* sa1;
  const float* sib1;
  const float* sa2;
  const float* sib2;
  int B, T, C, K, dil, pad_left;
  int slab_rows, nslot, n_pow2, tiles_per_item, total_tiles;
  uint32_t idesc;
  long long* trace;   // debug: [tile_it < 64][16 events] clock64 stamps of CTA 0 (NULL = off)
};

// per-stage clock stamps are compiled in only with -DBC_TRACE (python -m audiotokenization_b200.build with BC_TRACE=1):
// the tiles of the narrow layers are so short that even the disabled checks were ~7 % of the executed instructions
#ifdef BC_TRACE
#define TRACE(ev) do { if (p.trace && blockIdx.x == 0 && it < 64 && lane == 0) p.trace[it * 16 + (ev)] = clock64(); } while (0)
#else
#define TRACE(ev) do { } while (0)
#endif
// ring slot / use count of pipeline step `i` for nslot in {1, 2} without a runtime division
#define SLOT_OF(i) (p.nslot == 2 ? ((i) & 1) : 0)
#define USE_OF(i) (p.nslot == 2 ? ((i) >> 1) : (i))

enum { B_A_FULL = 0, B_A_EMPTY = 2, B_ACC1_FULL = 4, B_ACC1_EMPTY = 6, B_A2_FULL = 8, B_A2_EMPTY = 10,
       B_ACC2_FULL = 12, B_ACC2_EMPTY = 14, B_W_FULL = 16, N_BARS = 17 };

// r[0 .. 8*n8) += the accumulator columns at `taddr` (the a_hi * w_lo half), 16 columns at a time to bound registers
__device__ __forceinline__ void add_lo_half(uint32_t taddr, int n8, uint32_t r[32]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (2 * h < n8) {
      uint32_t t[32];
      tmem_load(taddr + 16u * h, min(2, n8 - 2 * h), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (e / 4 + 2 * h < n8) {   // pairs (2e, 2e+1) of this 16-column piece
          float x0, x1;
          unpack2(add2(pack2(__uint_as_float(r[16 * h + 2 * e]), __uint_as_float(r[16 * h + 2 * e + 1])),
                       pack2(__uint_as_float(t[2 * e]), __uint_as_float(t[2 * e + 1]))), x0, x1);
          r[16 * h + 2 * e] = __float_as_uint(x0);
          r[16 * h + 2 * e + 1] = __float_as_uint(x1);
        }
      }
    }
  }
}

// Stride between the 8-channel planes of the activation slab: rows * 16 B, padded to 16 (mod 128) so that the up to 8
// planes a warp writes side by side (one 16-byte row chunk each) fall into different banks.
__host__ __device__ inline uint32_t ru_plane_bytes(int slab_rows) {
  uint32_t b = (uint32_t)slab_rows * 16u;
  while (b % 128u != 16u) b += 16u;
  return b;
}

template <int SPLIT, int GROUPS>
__global__ void __launch_bounds__(RU_THREADS, 1) ru_persist_kernel(const RuParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C;
  const int planes = C / 8;
  // Split precision, K-tap conv: w_hi and w_lo sit side by side as ONE B operand of 2C rows, so a_hi meets both in a
  // single MMA of width 2C (accumulator columns [0,C) and [C,2C)) and a_lo * w_hi is a second MMA of width C into
  // [0,C): with one activation slot the tile period is LOAD-convert + the K-tap MMA chain (trace), and 28 x (64 + 48)
  // tensor cycles replace 84 x 48 at C = 64.  The MID stage, which is off that critical path, adds the two halves.
  // The 1x1 conv keeps three MMAs of width C: the STORE stage would otherwise become the longest.
  constexpr int ACCW = SPLIT == 2 ? 2 : 1;
  const uint32_t acc1_stage = (uint32_t)(ACCW * p.n_pow2);   // TMEM columns of one K-tap accumulator stage
  const uint32_t acc2_base = 2u * acc1_stage;                 // the two 1x1 accumulator stages follow
  uint32_t tmem_cols = 32;
  while (tmem_cols < acc2_base + 2u * (uint32_t)p.n_pow2) tmem_cols <<= 1;
  const uint32_t w7_split = (uint32_t)p.K * C * C * 2u;
  const uint32_t w1_split = (uint32_t)C * C * 2u;
  const uint32_t plane_bytes = ru_plane_bytes(p.slab_rows);
  const uint32_t a_split = planes * plane_bytes;
  const uint32_t a_slot = (a_split * SPLIT + 127u) & ~127u;
  const uint32_t a2_plane = BM * 16u;
  const uint32_t a2_split = planes * a2_plane;
  const uint32_t a2_slot = a2_split * SPLIT;

  uint8_t* sW7 = smem_raw;
  uint8_t* sW1 = sW7 + w7_split * SPLIT;
  uint8_t* sA = sW1 + w1_split * SPLIT;
  uint8_t* sA2 = sA + (size_t)a_slot * p.nslot;
  float* sPar = reinterpret_cast<float*>(sA2 + (size_t)a2_slot * p.nslot);  // b7 | sa2 | sib2 | b1
  float* sStage = sPar + 4 * C;                      // 4 x [32][EPI_LD] fp32: transposes accumulator rows into whole HBM lines
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 4 * 32 * EPI_LD);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

  // ---- one-time setup ----
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_A_FULL + s), p.nslot == 2 ? GROUP_WARPS : LOAD_WARPS);
      mbar_init(BAR(B_A_EMPTY + s), 1);
      mbar_init(BAR(B_ACC1_FULL + s), 1);
      mbar_init(BAR(B_ACC1_EMPTY + s), MID_WARPS);
      mbar_init(BAR(B_A2_FULL + s), MID_WARPS);
      mbar_init(BAR(B_A2_EMPTY + s), 1);
      mbar_init(BAR(B_ACC2_FULL + s), 1);
      mbar_init(BAR(B_ACC2_EMPTY + s), 4);
    }
    mbar_init(BAR(B_W_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // resident weights: one expect_tx; copies in <= 32 KB pieces, except the split-precision K-tap image, which goes
    // piece by piece (hi/lo, tap-group, k-plane) so that it lands as [tap-group][k-plane][hi rows | lo rows]
    const uint32_t w7_bytes = w7_split * SPLIT, w1_bytes = w1_split * SPLIT;
    mbar_expect_tx(BAR(B_W_FULL), w7_bytes + w1_bytes);
    if (SPLIT == 1) {
      for (uint32_t off = 0; off < w7_bytes; off += 32768u)
        bulk_g2s_notx(smem_u32(sW7) + off, reinterpret_cast<const uint8_t*>(p.w7) + off, min(32768u, w7_bytes - off), BAR(B_W_FULL));
    } else {
      const uint32_t piece = (uint32_t)C * 16u;                 // one k-plane of one tap-group: C rows x 16 B
      for (int sp = 0; sp < 2; ++sp)
        for (int kg = 0; kg < p.K * GROUPS; ++kg) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w7) + (size_t)sp * w7_split + (size_t)kg * 2u * piece;
          const uint32_t dst = smem_u32(sW7) + (uint32_t)kg * 4u * piece + (uint32_t)sp * piece;
          bulk_g2s_notx(dst, src, piece, BAR(B_W_FULL));                       // k-plane 0
          bulk_g2s_notx(dst + 2u * piece, src + piece, piece, BAR(B_W_FULL));  // k-plane 1
        }
    }
    for (uint32_t off = 0; off < w1_bytes; off += 32768u)
      bulk_g2s_notx(smem_u32(sW1) + off, reinterpret_cast<const uint8_t*>(p.w1) + off, min(32768u, w1_bytes - off), BAR(B_W_FULL));
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < C; i += RU_THREADS) {
    sPar[i] = __ldg(p.b7 + i);
    sPar[C + i] = __ldg(p.sa2 + i);
    sPar[2 * C + i] = __ldg(p.sib2 + i);
    sPar[3 * C + i] = __ldg(p.b1 + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int first = blockIdx.x, step = gridDim.x;

  if (warp < LOAD_WARPS) {
    // ======================= LOAD =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_LOAD));
    const int items = planes * p.slab_rows;
    // two slots: the loader groups take alternate tiles (group g always fills slot g, so a producer is never
    // more than one mbarrier phase ahead); one slot: all 8 warps stage every tile together
    const int ngroups = p.nslot == 2 ? LOAD_GROUPS : 1;
    const int gthreads = LOAD_THREADS / ngroups;
    const int grp = tid / gthreads;
    const int gtid = tid - grp * gthreads;
    const int pshift = 31 - __clz(planes);       // planes is a power of two (C in {16,32,64})
    const int pl = gtid & (planes - 1);
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.sa1 + pl * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.sa1 + pl * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.sib1 + pl * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.sib1 + pl * 8) + 1);
    int it = grp;
    // (item, tile-in-item) advance incrementally: one division per CTA instead of one per tile
    int b = (first + grp * step) / p.tiles_per_item, tt = (first + grp * step) - b * p.tiles_per_item;
    for (int tile = first + grp * step; tile < p.total_tiles; tile += ngroups * step, it += ngroups, tt += ngroups * step) {
      while (tt >= p.tiles_per_item) { tt -= p.tiles_per_item; ++b; }
      const int slot = SLOT_OF(it), use = USE_OF(it);
      const int g0 = tt * BM - p.pad_left;
      const float* xcol = p.x + (size_t)b * p.T * C + pl * 8;
      uint8_t* dstA = sA + (size_t)slot * a_slot + (size_t)pl * plane_bytes;
      bool waited = false;
      if (gtid == 0) TRACE(0);
      for (int i0 = gtid; i0 < items; i0 += gthreads * LD_BATCH) {
        float4 lo4[LD_BATCH], hi4[LD_BATCH];
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          const int i = i0 + j * gthreads;
          const int g = g0 + (i >> pshift);
          if (i < items && g >= 0 && g < p.T) {
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
          if (gtid == 0) TRACE(1);
        }
#if BC_RU_ILP
        // arithmetic of the whole batch first, branch-free (items beyond the slab hold zeros), then the predicated stores:
        // LD_BATCH independent SnakeBeta -> split chains interleave instead of one guarded block per item
        uint4 hq[LD_BATCH], lq[LD_BATCH];
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          float v[8] = {lo4[j].x, lo4[j].y, lo4[j].z, lo4[j].w, hi4[j].x, hi4[j].y, hi4[j].z, hi4[j].w};
          snake8<SPLIT>(v, a0, a1, b0, b1);
          split8<SPLIT>(v, hq[j], lq[j]);
        }
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          const int i = i0 + j * gthreads;
          if (i < items) {
            *reinterpret_cast<uint4*>(dstA + (size_t)(i >> pshift) * 16) = hq[j];
            if (SPLIT == 2) *reinterpret_cast<uint4*>(dstA + (size_t)(i >> pshift) * 16 + a_split) = lq[j];
          }
        }
#else
#pragma unroll
        for (int j = 0; j < LD_BATCH; ++j) {
          const int i = i0 + j * gthreads;
          if (i < items) {
            float v[8] = {lo4[j].x, lo4[j].y, lo4[j].z, lo4[j].w, hi4[j].x, hi4[j].y, hi4[j].z, hi4[j].w};
            snake8<SPLIT>(v, a0, a1, b0, b1);
            split_store<SPLIT>(v, dstA + (size_t)(i >> pshift) * 16, a_split);
          }
        }
#endif
      }
      if (!waited) mbar_wait(BAR(B_A_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
      fence_async_smem();
      __syncwarp();
      if (gtid == 0) TRACE(2);
      if (lane == 0) mbar_arrive(BAR(B_A_FULL + slot));
    }
  } else if (warp < MID_WARP0) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_MMA));
   if (warp == MMA_WARP) {
    // ======================= MMA issue =======================
    // All address arithmetic below is built from kernel parameters, blockIdx and loop counters only, so it
    // stays in uniform registers; one elected lane issues a whole tile's MMAs back to back.
    {
      int n_my = 0;
      for (int tile = first; tile < p.total_tiles; tile += step) ++n_my;
      mbar_wait(BAR(B_W_FULL), 0);
      const uint32_t smem0 = smem_u32(smem_raw);
      const uint32_t uW7 = smem0, uW1 = uW7 + w7_split * SPLIT, uA = uW1 + w1_split * SPLIT, uA2 = uA + a_slot * p.nslot;
      const uint32_t w7_lo0 = desc_lo(uW7, (uint32_t)(ACCW * C) * 16u), w1_lo0 = desc_lo(uW1, (uint32_t)C * 16u);
      const uint32_t idesc_wide = idesc_bf16_m128(ACCW * C);    // [w_hi | w_lo] operand of the K-tap conv
      const uint32_t hi_d = desc_hi(128u);
      const uint32_t a_g = (2u * plane_bytes) >> 4, a_k = (uint32_t)p.dil, a_sp = a_split >> 4;
      const uint32_t b_g = ((uint32_t)C * 32u) >> 4, b7_g = ((uint32_t)(ACCW * C) * 32u) >> 4;
      const uint32_t a2_g = (2u * a2_plane) >> 4;
      for (int it = 0; it <= n_my; ++it) {
        if (it < n_my) {  // K-tap conv of tile `it`
          const int slot = SLOT_OF(it), use = USE_OF(it), as = it & 1, ause = it >> 1;
          mbar_wait(BAR(B_A_FULL + slot), (uint32_t)(use & 1));
          mbar_wait(BAR(B_ACC1_EMPTY + as), (uint32_t)((ause & 1) ^ 1));
          tc_fence_after();
          TRACE(3);
          const uint32_t d = tmem_base + (uint32_t)as * acc1_stage;
          const uint32_t a_lo0 = desc_lo(uA + (uint32_t)slot * a_slot, plane_bytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
              if (k < p.K) {
#pragma unroll
                for (int g = 0; g < GROUPS; ++g) {
                  const uint32_t a_lo = a_lo0 + (uint32_t)k * a_k + (uint32_t)g * a_g;
                  const uint32_t b_lo = w7_lo0 + (uint32_t)(k * GROUPS + g) * b7_g;
                  if (k == 0 && g == 0) mma_bf16_raw<false>(d, a_lo, b_lo, hi_d, hi_d, idesc_wide);   // a_hi * [w_hi | w_lo]
                  else                  mma_bf16_raw<true>(d, a_lo, b_lo, hi_d, hi_d, idesc_wide);
                  if (SPLIT == 2) mma_bf16_raw<true>(d, a_lo + a_sp, b_lo, hi_d, hi_d, p.idesc);           // a_lo * w_hi
                }
              }
            }
            umma_commit(BAR(B_A_EMPTY + slot));
            umma_commit(BAR(B_ACC1_FULL + as));
          }
          __syncwarp();
          TRACE(4);
        }
        if (it >= 1) {  // 1x1 conv of tile `it - 1`
          const int j = it - 1;
          const int slot = SLOT_OF(j), use = USE_OF(j), as = j & 1, ause = j >> 1;
          mbar_wait(BAR(B_A2_FULL + slot), (uint32_t)(use & 1));
          mbar_wait(BAR(B_ACC2_EMPTY + as), (uint32_t)((ause & 1) ^ 1));
          tc_fence_after();
          { const int it = j; TRACE(5); }
          const uint32_t d = tmem_base + acc2_base + (uint32_t)(as * p.n_pow2);
          const uint32_t a_lo0 = desc_lo(uA2 + (uint32_t)slot * a2_slot, a2_plane);
          if (elect_one()) {
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
              const uint32_t a_lo = a_lo0 + (uint32_t)g * a2_g, b_lo = w1_lo0 + (uint32_t)g * b_g;
              if (g == 0) mma_bf16_raw<false>(d, a_lo, b_lo, hi_d, hi_d, p.idesc);
              else        mma_bf16_raw<true>(d, a_lo, b_lo, hi_d, hi_d, p.idesc);
              if (SPLIT == 2) {
                mma_bf16_raw<true>(d, a_lo, b_lo + (w1_split >> 4), hi_d, hi_d, p.idesc);
                mma_bf16_raw<true>(d, a_lo + (a2_split >> 4), b_lo, hi_d, hi_d, p.idesc);
              }
            }
            umma_commit(BAR(B_A2_EMPTY + slot));
            umma_commit(BAR(B_ACC2_FULL + as));
          }
          __syncwarp();
          { const int it = j; TRACE(6); }
        }
      }
    }
   }
  } else if (warp < EPI_WARP0) {
    // ======================= MID: acc1 -> snake2 -> bf16 A2 tile =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int chalf = (warp - MID_WARP0) >> 2;       // which half of the channels this warp converts
    const int cbeg = chalf * (C / (MID_WARPS / 4)), cend = cbeg + C / (MID_WARPS / 4);
    int it = 0;
    for (int tile = first; tile < p.total_tiles; tile += step, ++it) {
      const int slot = SLOT_OF(it), use = USE_OF(it), as = it & 1, ause = it >> 1;
      mbar_wait(BAR(B_ACC1_FULL + as), (uint32_t)(ause & 1));
      tc_fence_after();
      mbar_wait(BAR(B_A2_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
      if (warp == MID_WARP0) TRACE(7);
      uint8_t* dst = sA2 + (size_t)slot * a2_slot + (size_t)row * 16;
      const uint32_t taddr = tmem_base + (uint32_t)as * acc1_stage + ((uint32_t)(q * 32) << 16);
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        const int n8 = min(4, (cend - c0) / 8);
        uint32_t r[32];
        tmem_load(taddr + (uint32_t)c0, n8, r);
        if (SPLIT == 2) add_lo_half(taddr + (uint32_t)(C + c0), n8, r);   // + a_hi * w_lo
        if (c0 + 32 >= cend) {  // this warp's share of the accumulator is read: hand it back before the math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ACC1_EMPTY + as));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < n8) {
            const int c = c0 + 8 * j;
            const float4 bi0 = *reinterpret_cast<const float4*>(sPar + c), bi1 = *reinterpret_cast<const float4*>(sPar + c + 4);
            const float4 s0 = *reinterpret_cast<const float4*>(sPar + C + c), s1 = *reinterpret_cast<const float4*>(sPar + C + c + 4);
            const float4 i0 = *reinterpret_cast<const float4*>(sPar + 2 * C + c), i1 = *reinterpret_cast<const float4*>(sPar + 2 * C + c + 4);
            float v[8];
            acc_bias8(r + 8 * j, bi0, bi1, v);
            snake8<SPLIT>(v, s0, s1, i0, i1);
            split_store<SPLIT>(v, dst + (size_t)(c / 8) * a2_plane, a2_split);
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (warp == MID_WARP0) TRACE(8);
      if (lane == 0) mbar_arrive(BAR(B_A2_FULL + slot));
    }
  } else {
    // ======================= STORE: acc2 + b1 + x -> y =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_STORE));
    // The accumulator arrives one row per lane; a lane-per-row global access costs 32 L1 wavefronts per instruction
    // (the L1 data pipe was 92 % busy and the bound of this kernel).  Each warp owns a padded [32 rows][32 + 4] fp32
    // block: bias-added rows go in, and leave with 8 lanes per row (4 whole lines per instruction); the residual is
    // fetched in that same coalesced mapping and added from registers on the way out.
    const int q = warp & 3;
    float* sT = sStage + (size_t)(warp - EPI_WARP0) * (32 * EPI_LD);
    const int crow = lane >> 3, cchunk = (lane & 7) * 4;     // coalesced mapping: rows crow + 4*i, 4 floats at cchunk
    int it = 0;
    int b = first / p.tiles_per_item, tt = first - b * p.tiles_per_item;
    for (int tile = first; tile < p.total_tiles; tile += step, ++it, tt += step) {
      while (tt >= p.tiles_per_item) { tt -= p.tiles_per_item; ++b; }
      const int as = it & 1, ause = it >> 1;
      const int trow0 = tt * BM + q * 32;                    // first row of this warp's block
      const size_t off0 = ((size_t)b * p.T + trow0 + crow) * C + cchunk;
      const float* rp = p.x + off0;
      float* yp = p.y + off0;
      const size_t istep = (size_t)4 * C;
      const int rows_ok = p.T - trow0 - crow;                // row 4*i of this lane is valid iff 4*i < rows_ok
      const uint32_t taddr = tmem_base + acc2_base + (uint32_t)(as * p.n_pow2) + ((uint32_t)(q * 32) << 16);
      if (warp == EPI_WARP0) TRACE(9);
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int ncol = min(32, C - c0), n8 = ncol / 8;
        const bool col_ok = cchunk < ncol;
        float4 res4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          res4[i] = (col_ok && 4 * i < rows_ok) ? __ldg(reinterpret_cast<const float4*>(rp + c0 + i * istep)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 == 0) {
          mbar_wait(BAR(B_ACC2_FULL + as), (uint32_t)(ause & 1));
          tc_fence_after();
          if (warp == EPI_WARP0) TRACE(10);
        }
        uint32_t r[32];
        tmem_load(taddr + (uint32_t)c0, n8, r);
        if (c0 + 32 >= C) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ACC2_EMPTY + as));
        }
        float* own = sT + lane * EPI_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < 2 * n8) {
            const float4 bb = *reinterpret_cast<const float4*>(sPar + 3 * C + c0 + 4 * j);
            *reinterpret_cast<float4*>(own + 4 * j) = acc_bias4(r + 4 * j, bb);
          }
        }
        __syncwarp();
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(sT + (4 * i + crow) * EPI_LD + cchunk);
            if (4 * i < rows_ok) __stcs(reinterpret_cast<float4*>(yp + c0 + i * istep), add4(v, res4[i]));
          }
        }
        __syncwarp();
      }
      if (warp == EPI_WARP0) TRACE(11);
    }
  }
#undef BAR
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

size_t ru_smem_bytes(int C, int K, int dil, int split, int nslot) {
  const size_t slab_rows = (BM - 1) + (size_t)(K - 1) * dil + 1;
  const size_t w = (size_t)split * ((size_t)K * C * C * 2 + (size_t)C * C * 2);
  const size_t a_slot = ((size_t)split * (C / 8) * ru_plane_bytes((int)slab_rows) + 127) & ~size_t(127);
  const size_t a2_slot = (size_t)split * (C / 8) * BM * 16;
  return w + nslot * (a_slot + a2_slot) + 4 * C * sizeof(float) + STAGE_BYTES + N_BARS * 8 + 64;
}

}  // namespace

namespace bc {

long long* g_ru_trace = nullptr;   // shared with ru_group.cu (debug)

// 0 = not applicable, else number of smem operand slots the persistent kernel would use
int ru_persist_slots(int C, int K, int dilation, int precision) {
  if ((C != 16 && C != 32 && C != 64) || K > 7) return 0;   // power-of-two plane count; weights must stay resident
  if (!policy().ru_persist) return 0;
  const int split = precision == BC_PREC_BF16X3 ? 2 : 1;
  if (ru_smem_bytes(C, K, dilation, split, 2) <= 227 * 1024) return 2;
  if (ru_smem_bytes(C, K, dilation, split, 1) <= 227 * 1024) return 1;
  return 0;
}

int resunit_persist_fwd(const float* x, const float* w7, const float* b7, const float* sa1, const float* sib1,
                        const float* w1, const float* b1, const float* sa2, const float* sib2, float* y, int B, int T,
                        int C, int K, int dilation, int pad_left, int precision, cudaStream_t st) {
  const int nslot = ru_persist_slots(C, K, dilation, precision);
  if (nslot == 0 || K > 7) return fail(BC_EUNSUPPORTED, "resunit(persistent): C=%d K=%d dil=%d not supported", C, K, dilation);
  RuParams p;
  p.x = x; p.y = y; p.w7 = reinterpret_cast<const uint4*>(w7); p.w1 = reinterpret_cast<const uint4*>(w1);
  p.b7 = b7; p.b1 = b1; p.sa1 = sa1; p.sib1 = sib1; p.sa2 = sa2; p.sib2 = sib2;
  p.B = B; p.T = T; p.C = C; p.K = K; p.dil = dilation; p.pad_left = pad_left;
  p.slab_rows = (BM - 1) + (K - 1) * dilation + 1;
  p.nslot = nslot;
  p.n_pow2 = C <= 32 ? 32 : 64;
  p.tiles_per_item = (T + BM - 1) / BM;
  const long long total = (long long)p.tiles_per_item * B;
  if (total > 2147483647ll) return fail(BC_EINVAL, "resunit(persistent): too many tiles");
  p.total_tiles = (int)total;
  p.idesc = idesc_bf16_m128(C);
  p.trace = g_ru_trace;
  const int split = precision == BC_PREC_BF16X3 ? 2 : 1;
  const size_t smem = ru_smem_bytes(C, K, dilation, split, nslot);
  if ((size_t)ru_plane_bytes(p.slab_rows) * 2 >= (1u << 18)) return fail(BC_EUNSUPPORTED, "resunit(persistent): descriptor offset overflow");
  void (*kern)(const RuParams) = nullptr;
  const int gi = C == 16 ? 0 : (C == 32 ? 1 : 2);
  if (split == 1) kern = gi == 0 ? ru_persist_kernel<1, 1> : (gi == 1 ? ru_persist_kernel<1, 2> : ru_persist_kernel<1, 4>);
  else            kern = gi == 0 ? ru_persist_kernel<2, 1> : (gi == 1 ? ru_persist_kernel<2, 2> : ru_persist_kernel<2, 4>);
  static bool configured[64][6] = {{false}};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 0 || dev >= 64 || !configured[dev][(split - 1) * 3 + gi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_check(e, "cudaFuncSetAttribute(ru_persist)");
    if (dev >= 0 && dev < 64) configured[dev][(split - 1) * 3 + gi] = true;
  }
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  kern<<<grid, RU_THREADS, smem, st>>>(p);
  BC_LAUNCH_CHECK("ru_persist_kernel");
  return BC_OK;
}

}  // namespace bc

// debug hook (not part of the product path): device buffer of 64*16 int64 that receives clock64 stamps
extern "C" int bc_debug_set_ru_trace(void* device_buffer) {
  bc::g_ru_trace = reinterpret_cast<long long*>(device_buffer);
  return BC_OK;
}
