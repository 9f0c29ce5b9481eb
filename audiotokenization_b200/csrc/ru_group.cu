// Fused ResidualUnit for the narrow layers, "one warpgroup per tile" form (sm_100a).
//
//   y = x + W1 * snake2( W7 (*) snake1(x) + b7 ) + b1          (vq/module.py:74-89)
//
// ru_persist.cu splits the unit into roles (LOAD / MMA / MID / STORE warps) that hand a tile from one to the next;
// measured, the narrow layers are bound by the CUDA-core activation math and that pipeline keeps the SM's issue
// slots only ~40 % busy (each role idles while it waits for its neighbours).  Here every warpgroup (4 warps, 128
// threads = the 128 TMEM lanes of an accumulator) takes a tile through ALL the steps by itself:
//
//   stage   x (fp32, HBM) -> snake1 -> bf16 hi[/lo] -> K-major slab in the group's shared memory
//   mma7    one elected thread issues the K-tap conv into the group's acc1 (TMEM), commit -> mbarrier
//   mid     acc1 -> +b7 -> snake2 -> bf16 hi[/lo] -> the group's A2 tile
//   mma1    one elected thread issues the 1x1 conv into acc2, commit -> mbarrier
//   store   acc2 + b1 + x -> y, through a padded fp32 staging block (the group's idle operand buffers) so that the
//           residual rows are read and the output rows written as whole 128-byte lines
//
// and G such groups (up to 4) per CTA work on different tiles, out of phase: while one waits for HBM or for its
// MMAs, the others fill the issue slots.  Both weight images are resident in shared memory, shared by the groups.
// 512 threads leave 128 registers per thread (no spills).
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

#ifndef BC_RU_ILP     // 1: branch-free staging batches (conv_stream.cu); measured 1 % slower here -> 0: one guarded block per item
#define BC_RU_ILP 0
#endif
namespace {
using namespace bc::tc;

constexpr int BM = 128;
constexpr int GT = 128;          // threads per group
constexpr int MAX_G = 4;
constexpr int RG_THREADS = MAX_G * GT;

struct RgParams {
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
  int slab_rows, n_pow2, tiles_per_item, total_tiles, G, tmem_cols;
  uint32_t idesc, group_bytes, a_bytes;
  long long* trace;   // debug (-DBC_TRACE): [tile iteration < 64][16] clock64 stamps of group 0 of CTA 0
};
#ifdef BC_TRACE
#define GTRACE(ev) do { if (p.trace && blockIdx.x == 0 && g == 0 && gt == 0 && titer < 64) p.trace[titer * 16 + (ev)] = clock64(); } while (0)
#else
#define GTRACE(ev) do { } while (0)
#endif

__device__ __forceinline__ void group_sync(int g) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory");
}

template <int SPLIT, int GROUPS>
__global__ void __launch_bounds__(RG_THREADS, 1) ru_group_kernel(const RgParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int C = GROUPS * 16;
  constexpr int planes = C / 8;
  constexpr int LPR = C / 4;                 // lanes (16-byte chunks) per fp32 row
  constexpr int RPI = GT / LPR;              // rows covered by one coalesced group-wide instruction
  constexpr int SLD = C + 4;                 // staging row stride in floats (conflict-free 16-byte accesses both ways)
  // split precision, K-tap conv: [w_hi | w_lo] as ONE B operand of 2C rows -> a_hi * both in one MMA of width 2C, a_lo * w_hi
  // in a second of width C (two A fetches per tap-group instead of three: the kernel is bound by shared-memory bandwidth,
  // most of it the tensor core's operand reads); the mid step adds the two accumulator halves
  constexpr int ACCW = SPLIT == 2 ? 2 : 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = warp >> 2, gt = tid & (GT - 1), gw = warp & 3;
  const uint32_t w7_split = (uint32_t)p.K * C * C * 2u;
  const uint32_t w1_split = (uint32_t)C * C * 2u;
  const uint32_t plane_bytes = (uint32_t)p.slab_rows * 16u;
  const uint32_t a_split = planes * plane_bytes;
  const uint32_t a2_plane = BM * 16u;
  const uint32_t a2_split = planes * a2_plane;

  uint8_t* sW7 = smem_raw;
  uint8_t* sW1 = sW7 + w7_split * SPLIT;
  uint8_t* sG = sW1 + w1_split * SPLIT + (size_t)g * p.group_bytes;   // this group's [A slab | A2 tile]
  uint8_t* sA = sG;
  uint8_t* sA2 = sG + p.a_bytes;
  float* sPar = reinterpret_cast<float*>(sW1 + w1_split * SPLIT + (size_t)p.G * p.group_bytes);   // b7 | sa2 | sib2 | b1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPar + 4 * C);       // [0] weights, [1 + 2g] mma7 done, [2 + 2g] mma1 done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * MAX_G);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_w = bar0, bar7 = bar0 + 8u * (1 + 2 * g), bar1 = bar7 + 8u;

  if (tid == 0) {
    for (int i = 0; i < 1 + 2 * MAX_G; ++i) mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t w7_bytes = w7_split * SPLIT, w1_bytes = w1_split * SPLIT;
    mbar_expect_tx(bar_w, w7_bytes + w1_bytes);
    if (SPLIT == 1) {
      for (uint32_t off = 0; off < w7_bytes; off += 32768u)
        bulk_g2s_notx(smem_u32(sW7) + off, reinterpret_cast<const uint8_t*>(p.w7) + off, min(32768u, w7_bytes - off), bar_w);
    } else {   // piece by piece, so that the resident image is [tap-group][k-plane][hi rows | lo rows]
      const uint32_t piece = (uint32_t)C * 16u;
      for (int sp = 0; sp < 2; ++sp)
        for (int kg = 0; kg < p.K * GROUPS; ++kg) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w7) + (size_t)sp * w7_split + (size_t)kg * 2u * piece;
          const uint32_t dst = smem_u32(sW7) + (uint32_t)kg * 4u * piece + (uint32_t)sp * piece;
          bulk_g2s_notx(dst, src, piece, bar_w);
          bulk_g2s_notx(dst + 2u * piece, src + piece, piece, bar_w);
        }
    }
    for (uint32_t off = 0; off < w1_bytes; off += 32768u)
      bulk_g2s_notx(smem_u32(sW1) + off, reinterpret_cast<const uint8_t*>(p.w1) + off, min(32768u, w1_bytes - off), bar_w);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < C; i += RG_THREADS) {
    sPar[i] = __ldg(p.b7 + i);
    sPar[C + i] = __ldg(p.sa2 + i);
    sPar[2 * C + i] = __ldg(p.sib2 + i);
    sPar[3 * C + i] = __ldg(p.b1 + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (g < p.G) {
    const uint32_t acc1 = tmem_base + (uint32_t)((ACCW + 1) * g * p.n_pow2), acc2 = acc1 + (uint32_t)(ACCW * p.n_pow2);
    const uint32_t taddr_lane = (uint32_t)(gw * 32) << 16;      // this warp's TMEM lane quarter
    const int row = gw * 32 + lane;                              // accumulator row of this thread
    // staging geometry (LOAD): thread -> (plane, row offset)
    const int pl = gt & (planes - 1);
    const int prow = gt / planes;                                // rows prow + (GT / planes) * j
    constexpr int PRS = GT / planes;
    constexpr int SB = 6;                                        // 32-byte items a thread keeps in flight while staging
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.sa1 + pl * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.sa1 + pl * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.sib1 + pl * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.sib1 + pl * 8) + 1);
    // coalesced fp32 row mapping (STORE): thread -> (row offset, 16-byte chunk)
    const int crow = gt / LPR, cchunk = (gt % LPR) * 4;
    float* sT = reinterpret_cast<float*>(sG);
    const uint32_t hi_d = desc_hi(128u);
    const uint32_t w7_lo0 = desc_lo(smem_u32(sW7), (uint32_t)(ACCW * C) * 16u), w1_lo0 = desc_lo(smem_u32(sW1), (uint32_t)C * 16u);
    const uint32_t idesc_wide = idesc_bf16_m128(ACCW * C);
    const uint32_t a_lo0 = desc_lo(smem_u32(sA), plane_bytes), a2_lo0 = desc_lo(smem_u32(sA2), a2_plane);
    const uint32_t a_g = (2u * plane_bytes) >> 4, a_k = (uint32_t)p.dil, a_sp = a_split >> 4;
    const uint32_t b_g = ((uint32_t)C * 32u) >> 4, b7_g = ((uint32_t)(ACCW * C) * 32u) >> 4, a2_g = (2u * a2_plane) >> 4;

    const int tstep = (int)gridDim.x * p.G;
    int tile = (int)blockIdx.x * p.G + g;
    int b = tile / p.tiles_per_item, tt = tile - b * p.tiles_per_item;
    uint32_t ph = 0;
    bool w_ready = false;
    int titer = 0;
    (void)titer;
    for (; tile < p.total_tiles; tile += tstep, tt += tstep, ph ^= 1u, ++titer) {
      while (tt >= p.tiles_per_item) { tt -= p.tiles_per_item; ++b; }
      const int t0 = tt * BM;
      const int g0 = t0 - p.pad_left;
      const float* xb = p.x + (size_t)b * p.T * C;
      // ---- stage: x -> snake1 -> bf16 hi[/lo] slab ----
      GTRACE(0);
      {
        const float* xcol = xb + pl * 8;
        uint8_t* dst = sA + (size_t)pl * plane_bytes;
        for (int r0 = prow; r0 < p.slab_rows; r0 += PRS * SB) {   // SB items in flight: the whole slab at C = 32
          float4 lo4[SB], hi4[SB];
#pragma unroll
          for (int j = 0; j < SB; ++j) {
            const int r = r0 + PRS * j, gr = g0 + r;
            if (r < p.slab_rows && gr >= 0 && gr < p.T) {
              const float4* src = reinterpret_cast<const float4*>(xcol + (size_t)gr * C);
              lo4[j] = __ldg(src);
              hi4[j] = __ldg(src + 1);
            } else {
              lo4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              hi4[j] = lo4[j];
            }
          }
#if BC_RU_ILP
          // all the arithmetic of the batch first, branch-free (rows beyond the slab hold zeros), then the predicated stores:
          // the compiler interleaves the SB independent SnakeBeta -> split chains instead of running one guarded block per item
          uint4 hq[SB], lq[SB];
#pragma unroll
          for (int j = 0; j < SB; ++j) {
            float v[8] = {lo4[j].x, lo4[j].y, lo4[j].z, lo4[j].w, hi4[j].x, hi4[j].y, hi4[j].z, hi4[j].w};
            snake8<SPLIT>(v, a0, a1, b0, b1);
            split8<SPLIT>(v, hq[j], lq[j]);
          }
#pragma unroll
          for (int j = 0; j < SB; ++j) {
            const int r = r0 + PRS * j;
            if (r < p.slab_rows) {
              *reinterpret_cast<uint4*>(dst + (size_t)r * 16) = hq[j];
              if (SPLIT == 2) *reinterpret_cast<uint4*>(dst + (size_t)r * 16 + a_split) = lq[j];
            }
          }
#else
#pragma unroll
          for (int j = 0; j < SB; ++j) {
            const int r = r0 + PRS * j;
            if (r < p.slab_rows) {
              float v[8] = {lo4[j].x, lo4[j].y, lo4[j].z, lo4[j].w, hi4[j].x, hi4[j].y, hi4[j].z, hi4[j].w};
              snake8<SPLIT>(v, a0, a1, b0, b1);
              split_store<SPLIT>(v, dst + (size_t)r * 16, a_split);
            }
          }
#endif
        }
      }
      GTRACE(1);
      fence_async_smem();
      group_sync(g);
      GTRACE(2);
      // ---- mma7 ----
      if (gw == 0) {
        if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            if (k < p.K) {
#pragma unroll
              for (int gg = 0; gg < GROUPS; ++gg) {
                const uint32_t a_lo = a_lo0 + (uint32_t)k * a_k + (uint32_t)gg * a_g;
                const uint32_t b_lo = w7_lo0 + (uint32_t)(k * GROUPS + gg) * b7_g;
                if (k == 0 && gg == 0) mma_bf16_raw<false>(acc1, a_lo, b_lo, hi_d, hi_d, idesc_wide);   // a_hi * [w_hi | w_lo]
                else                   mma_bf16_raw<true>(acc1, a_lo, b_lo, hi_d, hi_d, idesc_wide);
                if (SPLIT == 2) mma_bf16_raw<true>(acc1, a_lo + a_sp, b_lo, hi_d, hi_d, p.idesc);            // a_lo * w_hi
              }
            }
          }
          umma_commit(bar7);
        }
        __syncwarp();
      }
      // ---- residual rows for the store step: requested now, consumed after both MMA chains ----
      constexpr int NRES = BM / RPI > 8 ? 8 : BM / RPI;          // rows per thread held in registers (first NRES*RPI rows)
      float4 res4[NRES];
      {
        const float* rp = xb + (size_t)(t0 + crow) * C + cchunk;
#pragma unroll
        for (int i = 0; i < NRES; ++i)
          res4[i] = (t0 + crow + RPI * i < p.T) ? __ldg(reinterpret_cast<const float4*>(rp + (size_t)RPI * i * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // ---- mid: acc1 -> +b7 -> snake2 -> A2 ----
      GTRACE(3);
      mbar_wait(bar7, ph);
      tc_fence_after();
      GTRACE(4);
      {
        uint8_t* dst = sA2 + (size_t)row * 16;
#pragma unroll
        for (int c0 = 0; c0 < C; c0 += 32) {
          uint32_t r[32];
          tmem_load32(acc1 + taddr_lane + (uint32_t)c0, r);
          if (SPLIT == 2) {   // + a_hi * w_lo, 16 columns at a time
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t t[32];
              tmem_load(acc1 + taddr_lane + (uint32_t)(C + c0 + 16 * h), 2, t);
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
#pragma unroll
          for (int j = 0; j < 4; ++j) {
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
      GTRACE(5);
      tc_fence_before();
      fence_async_smem();
      group_sync(g);
      GTRACE(6);
      // ---- mma1 ----
      if (gw == 0) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int gg = 0; gg < GROUPS; ++gg) {
            const uint32_t a_lo = a2_lo0 + (uint32_t)gg * a2_g, b_lo = w1_lo0 + (uint32_t)gg * b_g;
            if (gg == 0) mma_bf16_raw<false>(acc2, a_lo, b_lo, hi_d, hi_d, p.idesc);
            else         mma_bf16_raw<true>(acc2, a_lo, b_lo, hi_d, hi_d, p.idesc);
            if (SPLIT == 2) {
              mma_bf16_raw<true>(acc2, a_lo, b_lo + (w1_split >> 4), hi_d, hi_d, p.idesc);
              mma_bf16_raw<true>(acc2, a_lo + (a2_split >> 4), b_lo, hi_d, hi_d, p.idesc);
            }
          }
          umma_commit(bar1);
        }
        __syncwarp();
      }
      // ---- store: acc2 + b1 + x -> y ----
      GTRACE(7);
      mbar_wait(bar1, ph);          // also: the 1x1 conv has finished reading A2, the K-tap conv the slab -> staging is free
      tc_fence_after();
      GTRACE(8);
      // The residual never goes through the staging block: it was fetched in the coalesced (row group, 16-byte
      // chunk) mapping, which is also the mapping of the final store, so it is added there, from registers.
      {
        float* own = sT + (size_t)row * SLD;
#pragma unroll
        for (int c0 = 0; c0 < C; c0 += 32) {
          uint32_t r[32];
          tmem_load32(acc2 + taddr_lane + (uint32_t)c0, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = *reinterpret_cast<const float4*>(sPar + 3 * C + c0 + 4 * j);
            *reinterpret_cast<float4*>(own + c0 + 4 * j) = acc_bias4(r + 4 * j, bb);
          }
        }
      }
      GTRACE(9);
      tc_fence_before();
      group_sync(g);
      GTRACE(10);
      {
        float* yp = p.y + ((size_t)b * p.T + t0 + crow) * C + cchunk;
        const float* rp = xb + (size_t)(t0 + crow) * C + cchunk;
#pragma unroll
        for (int i = 0; i < BM / RPI; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(sT + (size_t)(crow + RPI * i) * SLD + cchunk);
          if (t0 + crow + RPI * i < p.T) {
            const float4 xr = i < NRES ? res4[i] : __ldg(reinterpret_cast<const float4*>(rp + (size_t)RPI * i * C));   // rows beyond the register prefetch
            __stcs(reinterpret_cast<float4*>(yp + (size_t)RPI * i * C), add4(xr, v));
          }
        }
      }
      GTRACE(11);
      group_sync(g);               // staging reads done before the next tile's slab overwrites it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

struct RgPlan {
  int G, split;
  uint32_t a_bytes, group_bytes;
  size_t smem;
};

bool rg_plan(int C, int K, int dilation, int precision, RgPlan* pl) {
  if ((C != 32 && C != 64) || K > 7 || K < 1) return false;
  if (precision != BC_PREC_BF16 && precision != BC_PREC_BF16X3) return false;
  const int split = precision == BC_PREC_BF16X3 ? 2 : 1;
  const size_t slab_rows = (BM - 1) + (size_t)(K - 1) * dilation + 1;
  const size_t w = (size_t)split * ((size_t)K * C * C * 2 + (size_t)C * C * 2);
  size_t a = ((size_t)split * (C / 8) * slab_rows * 16 + 127) & ~size_t(127);
  const size_t a2 = (size_t)split * (C / 8) * BM * 16;
  const size_t stage = (size_t)BM * (C + 4) * 4;                 // fp32 staging of the store step aliases [A | A2]
  if (a + a2 < stage) a = ((stage - a2) + 127) & ~size_t(127);
  const size_t group = a + a2;
  const size_t fixed = w + 4 * C * sizeof(float) + (1 + 2 * MAX_G) * 8 + 64;
  int G = (int)((225 * 1024 - fixed) / group);
  if (G > MAX_G) G = MAX_G;
  const int n_pow2 = C <= 32 ? 32 : 64;
  if (G > 512 / ((split + 1) * n_pow2)) G = 512 / ((split + 1) * n_pow2);   // K-tap accumulator (x2 in split precision: hi | lo halves) + 1x1 accumulator per group in 512 TMEM columns
  if (G < 3) return false;                                       // with fewer groups the role pipeline of ru_persist is better
  pl->G = G; pl->split = split; pl->a_bytes = (uint32_t)a; pl->group_bytes = (uint32_t)group;
  pl->smem = fixed + (size_t)G * group;
  return true;
}

}  // namespace

namespace bc {
extern long long* g_ru_trace;

// 0 = not applicable, else the number of warpgroups per CTA
int ru_group_groups(int C, int K, int dilation, int precision) {
  if (!policy().ru_group) return 0;
  RgPlan pl;
  return rg_plan(C, K, dilation, precision, &pl) ? pl.G : 0;
}

int resunit_group_fwd(const float* x, const float* w7, const float* b7, const float* sa1, const float* sib1,
                      const float* w1, const float* b1, const float* sa2, const float* sib2, float* y, int B, int T, int C,
                      int K, int dilation, int pad_left, int precision, cudaStream_t st) {
  RgPlan pl;
  if (!rg_plan(C, K, dilation, precision, &pl))
    return fail(BC_EUNSUPPORTED, "resunit(group): C=%d K=%d dil=%d not supported", C, K, dilation);
  RgParams p;
  p.trace = g_ru_trace;
  p.x = x; p.y = y; p.w7 = reinterpret_cast<const uint4*>(w7); p.w1 = reinterpret_cast<const uint4*>(w1);
  p.b7 = b7; p.b1 = b1; p.sa1 = sa1; p.sib1 = sib1; p.sa2 = sa2; p.sib2 = sib2;
  p.B = B; p.T = T; p.C = C; p.K = K; p.dil = dilation; p.pad_left = pad_left;
  p.slab_rows = (BM - 1) + (K - 1) * dilation + 1;
  p.n_pow2 = C <= 32 ? 32 : 64;
  p.tiles_per_item = (T + BM - 1) / BM;
  const long long total = (long long)p.tiles_per_item * B;
  if (total > 2147483647ll) return fail(BC_EINVAL, "resunit(group): too many tiles");
  p.total_tiles = (int)total;
  p.G = pl.G;
  p.tmem_cols = 32;
  while (p.tmem_cols < (pl.split + 1) * pl.G * p.n_pow2) p.tmem_cols <<= 1;   // power of two >= 32
  p.idesc = idesc_bf16_m128(C);
  p.group_bytes = pl.group_bytes;
  p.a_bytes = pl.a_bytes;
  if ((size_t)p.slab_rows * 16 * 2 >= (1u << 18)) return fail(BC_EUNSUPPORTED, "resunit(group): descriptor offset overflow");
  void (*kern)(const RgParams) = nullptr;
  const int gi = C == 32 ? 0 : 1;
  if (pl.split == 1) kern = gi == 0 ? ru_group_kernel<1, 2> : ru_group_kernel<1, 4>;
  else               kern = gi == 0 ? ru_group_kernel<2, 2> : ru_group_kernel<2, 4>;
  static bool configured[64][4] = {{false}};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int slot = (pl.split - 1) * 2 + gi;
  if (dev < 0 || dev >= 64 || !configured[dev][slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_check(e, "cudaFuncSetAttribute(ru_group)");
    if (dev >= 0 && dev < 64) configured[dev][slot] = true;
  }
  const long long want = (total + pl.G - 1) / pl.G;
  const int grid = (int)(want < sms ? want : sms);
  kern<<<grid, RG_THREADS, pl.smem, st>>>(p);
  BC_LAUNCH_CHECK("ru_group_kernel");
  return BC_OK;
}

}  // namespace bc
