// LSTM recurrence on the tensor cores (tcgen05), sm_100a -- BC_PREC_BF16 / BC_PREC_BF16X3.
//
//   G_t[B x 4H] = pre_t + h_{t-1}[B x H] * W_hh^T ,  gates i,f,g,o   (nn.LSTM inside ResLSTM, vq/module.py:143-167)
//
// Decomposition.  A CTA owns NS gate columns (n) = U = NS/4 hidden units with all four gates -- its W_hh slice
// [NS x H] (bf16 hi [, lo]) stays resident in shared memory for the whole sequence -- and TPC batch tiles of 128 rows.
// Every time step, for each of its batch tiles in turn:
//   TMA thread   waits until the n-slices that own a K chunk of the tile's h_{t-1} have published it, then streams
//                the bf16 h tile [128 x H] (stored in HBM/L2 directly in the UMMA K-major image) in K chunks
//                through a shared-memory ring (cp.async.bulk + mbarrier);
//   MMA thread   H/16 [x2] tcgen05.mma (M=128) into the tile's TMEM accumulator, chunk by chunk as they land.
//                Split precision: w_hi and w_lo sit side by side as ONE B operand of 2*NS rows, so a_hi meets both in
//                a single MMA of width 2*NS (columns [0,NS) = hi*hi, [NS,2NS) = hi*lo) and a_lo * w_hi is a second one
//                of width NS into columns [0,NS): two A fetches and 48 + 40 tensor cycles per 16 channels instead of
//                three and 3 x 40 (NS = 32);
//   16 gate warps  tcgen05.ld their (32 rows x UPW units x 4 gates) patch, add the pre-activation (prefetched
//                from HBM during the MMA phase), apply the gates with c kept in registers, write y (fp32,
//                + skip) and publish h_t as bf16 hi[/lo] into the exchange buffer, then bump the tile's
//                step counter (release); no grid-wide barrier -- only the CTAs that share batch rows wait
//                for each other.
// TPC = 2 (batches above 256 rows in split precision): the two tiles of a CTA are independent sequences, so while the
// h_t of one is being published, becoming visible and travelling through L2 (the latency chain that bounds a step:
// ~8 us, tensor pipe and copy engine mostly idle), the other tile's copies, MMAs and gates run: 11.2 us per step of
// 512 rows against 2 x 8.0 us for two launches of 256 (B200, H = 512).  Measured and rejected: separate TMA threads,
// gate-warp halves and a tagged, shared ring per tile (13.2 us: eight gate warps per tile double the gate phase that
// sits on each tile's chain, and the ring is still only two 64 KB slots deep); 64-channel chunks (12.5 us at 256 rows:
// the per-chunk barrier waits and commits of the MMA thread cost more than the finer pipelining gains); one 16-byte
// poll of four chunk counters (neutral).
// The kernel is launched cooperatively (all CTAs must be co-resident because they wait on each other).
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace bc::tc;

constexpr int LM = 128;         // batch rows per tile
constexpr int KC = 128;         // K elements per streamed chunk
constexpr int GATE_WARPS = 16;
constexpr int CNT_STRIDE = 16;  // counters per batch tile (>= H / KC)
constexpr int L_THREADS = (4 + GATE_WARPS) * 32;   // warp 0: TMA, warp 1: MMA, warps 2-3 idle, warps 4-19: gates
constexpr int MAX_TPC = 2;

struct LstmTcParams {
  const float* pre;        // [B][T][4H]
  const uint4* wimg;       // [n_slices][H/16][2][split*NS][8] bf16 (rows: hi slice, then lo slice)
  const float* skip;       // [B][T][H] or NULL
  float* y;                // [B][T][H]
  __nv_bfloat16* hx;       // [2][m_tiles][split][H/8][128][8]
  unsigned int* counters;  // [m_tiles][CNT_STRIDE]: one step counter per (batch tile, K chunk of h)
  int B, T, H, NS, nslot, n_slices, m_tiles;
  int pre_rows, y_rows;    // rows (time steps) between consecutive batch items of pre / of y and skip (>= T: chunked sequences)
  uint32_t split_stride, ring_bytes;   // shared-memory ring: hi -> lo stride inside a slot, bytes reserved for all slots
  int whole;               // 1: the compact tile image is one streamed piece per split (see the kernel)
  int rows;                // rows per 8-channel plane of the exchange image: 128, or round_up(B, 8) for one-tile launches
  int t_base;              // global index of this launch's first step (exchange-buffer parity; > 0: continue from c_state / hx)
  float* c_state;          // [m_tiles * 128][H] cell state carried between the chunks of a sequence (NULL: none)
  uint32_t idesc_wide, idesc_ns;   // N = split*NS (a_hi x [w_hi | w_lo]) and N = NS (a_lo x w_hi)
  long long* trace;   // debug: [step < 64][8] clock64 stamps of CTA (0,0), steps 100.. (NULL = off)
};

#ifdef BC_TRACE
#define LTRACE(ev) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && j == 0 && t >= 100 && t < 164) p.trace[(t - 100) * 8 + (ev)] = clock64(); } while (0)
#else
#define LTRACE(ev) do { } while (0)
#endif

// Polling a step counter: BC_LSTM_POLL 0 = ld.acquire.gpu every iteration; 1 = relaxed loads and one fence.acq_rel.gpu once
// the count is reached.  Measured: 1 is SLOWER (B = 512: 13.0 vs 11.3 us per step, B = 1: 6.5 vs 5.0) -- the fence is a
// full MEMBAR, while the ~1.6 k cycles a poll takes in the step trace are the L2 round trip under the exchange traffic,
// which the relaxed load pays as well.
#ifndef BC_LSTM_POLL
#define BC_LSTM_POLL 0
#endif
#if BC_LSTM_POLL == 0
#define POLL_LD(seen, ptr) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ptr) : "memory")
#define POLL_FENCE() do { } while (0)
#else
#define POLL_LD(seen, ptr) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ptr) : "memory")
#define POLL_FENCE() asm volatile("fence.acq_rel.gpu;" ::: "memory")
#endif

// Gate non-linearities of the tensor-core modes: SFU exp + approximate divide (|error| ~ 2e-7, far below the
// bf16x3 operand error; the fp32 mode runs lstm.cu with expf / tanhf).  Both saturate correctly for ANY finite
// input: the cell state is unbounded (c grows by up to 1 per step), so tanh must not produce inf/inf.
//   sigmoid: exp(-x) -> +inf for x < -88 and __fdividef(1, inf) = 0; exp(-x) -> 0 for x > 88 gives 1.
//   tanh:    evaluated on |x| (e = exp(-2|x|) in (0, 1], denominator in [1, 2]) and the sign restored.
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  const float e = __expf(-2.f * fabsf(x));
  return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

template <int SPLIT, int UPW, int TPC>
__global__ void __launch_bounds__(L_THREADS, 1) lstm_tc_kernel(const LstmTcParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, NS = p.NS, U = NS / 4;
  const int nchunks_cnt = H / KC;                   // step counters per tile (one per 128 channels of h)
  // `whole` (compact image that fits one ring slot): the tile's image is ONE piece per split -- one poll of all chunk counters,
  // one bulk copy per split, one barrier -- instead of H / KC chunks with their own poll, copies and barriers
  const bool whole = p.whole != 0;
  const int nchunks = whole ? 1 : nchunks_cnt;
  const int n = blockIdx.x;
  // batch tiles of this CTA: blockIdx.y, blockIdx.y + gridDim.y (the second one may not exist)
  int ntile = 0;
#pragma unroll
  for (int j = 0; j < TPC; ++j) ntile += ((int)blockIdx.y + j * (int)gridDim.y) < p.m_tiles ? 1 : 0;
  const uint32_t w_bytes = (uint32_t)SPLIT * NS * H * 2u;
  const uint32_t chunk_split = p.split_stride;                    // shared-memory bytes between the hi and lo images of a ring slot (32 KB; compact image: less)
  // Small batches (one tile, B < 128): the exchange image holds only R = round_up(B, 8) rows per 8-channel plane, planes
  // R * 16 bytes apart -- in HBM/L2 AND in shared memory (the operand descriptor's plane stride is R * 16).  The MMA still
  // reads 128 rows per plane; rows >= R alias the following planes (finite values, or stale shared memory past the last
  // plane) and only feed accumulator rows >= B, which nobody stores.  At B = 1 a step moves 16 KB of h per CTA instead of 256.
  const uint32_t R = (uint32_t)p.rows;
  const uint32_t chunk_bytes = R * KC * 2u;                       // bytes one bulk copy moves
  const uint32_t slot_bytes = chunk_split * SPLIT;
  const uint32_t acc_cols = (uint32_t)SPLIT * NS;                // TMEM columns of one tile's accumulator
  uint8_t* sW = smem_raw;
  uint8_t* sA = sW + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + p.ring_bytes);
  // bars: full[nslot] | empty[nslot] | acc_full[MAX_TPC] | acc_empty[MAX_TPC] | w_full
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8u * p.nslot, bar_accf = bar0 + 16u * p.nslot,
                 bar_acce = bar_accf + 8u * MAX_TPC, bar_w = bar_acce + 8u * MAX_TPC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.nslot + 2 * MAX_TPC + 1);
  uint32_t tmem_cols = 32;
  while (tmem_cols < acc_cols * TPC) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < p.nslot; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int j = 0; j < MAX_TPC; ++j) {
      mbar_init(bar_accf + 8u * j, 1);
      mbar_init(bar_acce + 8u * j, GATE_WARPS);
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_w, w_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) + (size_t)n * w_bytes;
    for (uint32_t off = 0; off < w_bytes; off += 32768u)
      bulk_g2s_notx(smem_u32(sW) + off, src + off, min(32768u, w_bytes - off), bar_w);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const size_t hx_split = (size_t)(H / 8) * R * 8;                // bf16 elements of one split of an m-tile image
  const size_t hx_tile = (size_t)SPLIT * hx_split;
  const size_t hx_parity = hx_tile * p.m_tiles;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t cc = 0;
      for (int t = 0; t < p.T; ++t) {
#pragma unroll
        for (int j = 0; j < TPC; ++j) {
          if (j >= ntile) break;
          const int m = (int)blockIdx.y + j * (int)gridDim.y;
          // K chunk c of h_{t-1} is written by the KC/U n-slices that own its hidden units: each chunk is fetched as soon
          // as ITS producers have published (per-chunk step counters), so the copies and MMAs of the early chunks overlap
          // the stragglers of the later ones instead of waiting for the slowest of all n-slices
          const __nv_bfloat16* src = p.hx + (size_t)((p.t_base + t + 1) & 1) * hx_parity + (size_t)m * hx_tile;
          for (int c = 0; c < nchunks; ++c, ++cc) {
            if (t > 0 && whole) {
              // ONE 16-byte acquire load covers the tile's chunk counters (polls one after the other cost an L2 round trip
              // each, ~1.5 k cycles, even when already satisfied)
              {
                const unsigned int target = (unsigned int)t * (unsigned int)(KC / U);
                unsigned int s0, s1, s2, s3, spins = 0;
                do {
                  asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3)
                               : "l"(p.counters + m * CNT_STRIDE) : "memory");
                  if (nchunks_cnt < 4) s3 = target;
                  if (nchunks_cnt < 3) s2 = target;
                  if (nchunks_cnt < 2) s1 = target;
                  if (++spins > (1u << 26)) __trap();
                } while (s0 < target || s1 < target || s2 < target || s3 < target);
                asm volatile("fence.proxy.async.global;" ::: "memory");
              }
            } else if (t > 0) {
              const unsigned int target = (unsigned int)t * (unsigned int)(KC / U);
              unsigned int seen;
              unsigned int spins = 0;
              do {
                POLL_LD(seen, p.counters + m * CNT_STRIDE + c);
                if (++spins > (1u << 26)) __trap();
              } while (seen < target);
              POLL_FENCE();
              asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes of other CTAs -> async-proxy reads
            }
            if (c == 0) LTRACE(0);
            const uint32_t slot = cc % p.nslot, use = cc / p.nslot;
            mbar_wait(bar_empty + 8u * slot, (use & 1u) ^ 1u);
            const uint32_t piece = whole ? (uint32_t)hx_split * 2u : chunk_bytes;
            mbar_expect_tx(bar_full + 8u * slot, piece * SPLIT);
#pragma unroll
            for (int sp = 0; sp < SPLIT; ++sp)
              bulk_g2s_notx(smem_u32(sA) + slot * slot_bytes + sp * chunk_split,
                            src + (size_t)sp * hx_split + (size_t)c * (KC / 8) * R * 8, piece, bar_full + 8u * slot);
          }
          LTRACE(1);
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (warp-uniform loop, elected lane issues) =======================
    {
      mbar_wait(bar_w, 0);
      const uint32_t a_plane = R * 16u;
      const uint32_t hi_d = desc_hi(128u);
      const uint32_t b_plane = acc_cols * 16u;                      // stride between the two k-planes of a 16-channel group
      const uint32_t w_lo0 = desc_lo(smem_u32(sW), b_plane);
      uint32_t cc = 0;
      for (int t = 0; t < p.T; ++t) {
#pragma unroll
        for (int j = 0; j < TPC; ++j) {
          if (j >= ntile) break;
          mbar_wait(bar_acce + 8u * j, ((uint32_t)t & 1u) ^ 1u);   // gate warps have drained this tile's accumulator of step t-1
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)j * acc_cols;
          for (int c = 0; c < nchunks_cnt; ++c) {
            const bool first_piece = !whole || c == 0, last_piece = !whole || c == nchunks_cnt - 1;
            const uint32_t slot = cc % p.nslot, use = cc / p.nslot;
            if (first_piece) {
              mbar_wait(bar_full + 8u * slot, use & 1u);
              tc_fence_after();
            }
            // whole image: the 128-channel blocks of the single piece follow one another (KC / 8 planes each)
            uint32_t a_lo = desc_lo(smem_u32(sA) + slot * slot_bytes + (whole ? (uint32_t)c * (KC / 8) * a_plane : 0u), a_plane);
            uint32_t b_lo = w_lo0 + (((uint32_t)c * (KC / 16) * 2u * b_plane) >> 4);
            const uint32_t a_g = (2u * a_plane) >> 4, b_g = (2u * b_plane) >> 4;
#pragma unroll
            for (int g = 0; g < KC / 16; ++g, a_lo += a_g, b_lo += b_g) {
              if (g == 0 && c == 0) mma_bf16_lohi<false>(d, a_lo, b_lo, hi_d, hi_d, p.idesc_wide);
              else                  mma_bf16_lohi<true>(d, a_lo, b_lo, hi_d, hi_d, p.idesc_wide);
              if (SPLIT == 2) mma_bf16_lohi<true>(d, a_lo + (chunk_split >> 4), b_lo, hi_d, hi_d, p.idesc_ns);   // a_lo * w_hi
            }
            if (last_piece) {
              if (elect_one()) umma_commit(bar_empty + 8u * slot);
              __syncwarp();
              ++cc;
            }
          }
          if (elect_one()) umma_commit(bar_accf + 8u * j);
          __syncwarp();
          if (lane == 0) LTRACE(2);
        }
      }
    }
  } else if (warp >= 4) {
    // ======================= gate warps =======================
    const int gw = warp - 4;
    const int q = warp & 3;                 // TMEM lane quarter
    const int ug = gw >> 2;                 // unit group within the slice
    const int row = q * 32 + lane;          // row within the m-tile
    const int u_loc = ug * UPW;             // first unit (within the slice) of this thread
    const int u_glb = n * U + u_loc;        // first hidden unit (global index)
    float c_state[TPC][UPW];
    float pg[TPC][4][UPW];
    bool row_ok[TPC];
    const float* pre_row[TPC];
    size_t out_row[TPC], hx_off[TPC];
    unsigned int* counter[TPC];
#pragma unroll
    for (int j = 0; j < TPC; ++j) {
      const int m = (int)blockIdx.y + j * (int)gridDim.y;
      const int b = m * LM + row;
      row_ok[j] = j < ntile && b < p.B;
      pre_row[j] = p.pre + (size_t)(row_ok[j] ? b : 0) * p.pre_rows * 4 * H + u_glb;
      out_row[j] = (size_t)(row_ok[j] ? b : 0) * p.y_rows * H + u_glb;
      // exchange-buffer position of this thread's units: plane = u_glb / 8, element = u_glb % 8
      hx_off[j] = (size_t)(j < ntile ? m : 0) * hx_tile + ((size_t)(u_glb >> 3) * R + row) * 8 + (u_glb & 7);
      counter[j] = p.counters + (j < ntile ? m : 0) * CNT_STRIDE + (n * U) / KC;
#pragma unroll
      for (int u = 0; u < UPW; ++u)
        c_state[j][u] = (p.c_state && p.t_base > 0 && row_ok[j]) ? p.c_state[(size_t)b * H + u_glb + u] : 0.f;
      // prefetch pre-activations of step 0
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int u = 0; u < UPW; ++u) pg[j][g][u] = row_ok[j] ? __ldcs(pre_row[j] + (size_t)g * H + u) : 0.f;
    }
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < p.T; ++t) {
#pragma unroll
      for (int j = 0; j < TPC; ++j) {
        if (j >= ntile) break;
        mbar_wait(bar_accf + 8u * j, (uint32_t)t & 1u);
        tc_fence_after();
        if (gw == 0 && lane == 0) LTRACE(3);
        const uint32_t taddr = taddr0 + (uint32_t)j * acc_cols;
        uint32_t acc[SPLIT][4][UPW];
#pragma unroll
        for (int sp = 0; sp < SPLIT; ++sp)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t col = taddr + (uint32_t)(sp * NS + g * U + u_loc);
            if constexpr (UPW == 4) {
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(acc[sp][g][0]), "=r"(acc[sp][g][1]), "=r"(acc[sp][g][2]), "=r"(acc[sp][g][3])
                           : "r"(col));
            } else {
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                           : "=r"(acc[sp][g][0]), "=r"(acc[sp][g][1])
                           : "r"(col));
            }
          }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce + 8u * j);
        float hv[UPW];
#pragma unroll
        for (int u = 0; u < UPW; ++u) {
          float a4[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float a = __uint_as_float(acc[0][g][u]);
            if (SPLIT == 2) a += __uint_as_float(acc[SPLIT - 1][g][u]);    // (hi*hi + lo*hi) + hi*lo
            a4[g] = a + pg[j][g][u];
          }
          const float c = fmaf(sigmoid_acc(a4[1]), c_state[j][u], sigmoid_acc(a4[0]) * tanh_acc(a4[2]));
          c_state[j][u] = c;
          hv[u] = sigmoid_acc(a4[3]) * tanh_acc(c);
        }
        // publish h_t (bf16 hi[/lo]) for the next step's MMA, in the UMMA K-major image -- FIRST: every other CTA of
        // this batch tile waits for it.  The fences below wait for all earlier memory operations of the thread, so
        // nothing else (output store, skip load, next step's pre-activation loads) may be in flight before them.
        if ((uint32_t)row < R) {
          __nv_bfloat16* dst = p.hx + (size_t)((p.t_base + t) & 1) * hx_parity + hx_off[j];
          __nv_bfloat16 hi[UPW], lo[UPW];
#pragma unroll
          for (int u = 0; u < UPW; ++u) {
            hi[u] = __float2bfloat16_rn(hv[u]);
            lo[u] = __float2bfloat16_rn(hv[u] - __bfloat162float(hi[u]));
          }
          if constexpr (UPW == 4) {
            *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(hi);
            if (SPLIT == 2) *reinterpret_cast<uint2*>(dst + hx_split) = *reinterpret_cast<uint2*>(lo);
          } else {
            *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<uint32_t*>(hi);
            if (SPLIT == 2) *reinterpret_cast<uint32_t*>(dst + hx_split) = *reinterpret_cast<uint32_t*>(lo);
          }
        }
        // make the h stores visible: generic -> async proxy per thread, then the CTA barrier orders every gate
        // thread's stores before ONE release-increment at gpu scope (release is cumulative over the barrier, the
        // pattern of a grid barrier) -- a per-thread __threadfence before the barrier cost one more L2 round trip
        // on the step's critical path (8.8 -> 8.0 us per step).  Polling the chunk counters from one lane per chunk
        // instead of one after the other was measured slower (9.1 us: four times the polling traffic on four hot words).
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (gw == 0 && lane == 0) LTRACE(4);
        asm volatile("bar.sync 1, %0;" ::"n"(GATE_WARPS * 32) : "memory");
        if (gw == 0 && lane == 0) {
          LTRACE(6);
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter[j]) : "memory");
          LTRACE(5);
        }
        // off the critical path (overlaps the other CTAs' publishes, the h copies and the next MMA phase):
        // next step's pre-activations and this step's output row
        if (t + 1 < p.T) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int u = 0; u < UPW; ++u)
              pg[j][g][u] = row_ok[j] ? __ldcs(pre_row[j] + (size_t)(t + 1) * 4 * H + (size_t)g * H + u) : 0.f;
        }
        if (row_ok[j]) {
          const size_t o = out_row[j] + (size_t)t * H;
          if constexpr (UPW == 4) {
            float4 v = make_float4(hv[0], hv[1], hv[2], hv[3]);
            if (p.skip) {
              const float4 s4 = __ldcs(reinterpret_cast<const float4*>(p.skip + o));
              v.x += s4.x; v.y += s4.y; v.z += s4.z; v.w += s4.w;
            }
            __stcs(reinterpret_cast<float4*>(p.y + o), v);
          } else {
            float2 v = make_float2(hv[0], hv[1]);
            if (p.skip) {
              const float2 s2 = __ldcs(reinterpret_cast<const float2*>(p.skip + o));
              v.x += s2.x; v.y += s2.y;
            }
            __stcs(reinterpret_cast<float2*>(p.y + o), v);
          }
        }
      }
    }
    if (p.c_state) {   // carry the cell state to the next chunk of the sequence
#pragma unroll
      for (int j = 0; j < TPC; ++j) {
        if (!row_ok[j]) continue;
        const int b = ((int)blockIdx.y + j * (int)gridDim.y) * LM + row;
#pragma unroll
        for (int u = 0; u < UPW; ++u) p.c_state[(size_t)b * H + u_glb + u] = c_state[j][u];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------
// CTA-pair form (tcgen05 cta_group::2; split precision, an even number of batch tiles; opt-in: BC_LSTM_PAIR=1).
// MEASURED: bit-identical to the two-tiles-per-CTA form and NOT faster (B = 512, H = 512: 11.3 vs 10.9-11.2 us per step)
// although it halves the h traffic -- the step is a chain of L2 round trips (h stores + release 3.5 k cycles, the
// publish becoming visible to a poll 2.5-3.2 k, ~1.6 k per poll, first 64 KB chunk 2.7 k after the counter), not L2 or
// shared-memory bandwidth; the ping-pong of two independent tiles per CTA hides more of that chain than the pair saves.  The two CTAs of a cluster own
// TWO batch tiles (M = 256: one each) and 64 gate columns = 16 hidden units: every SM still holds 96 rows of W_hh (rank 0:
// the w_hi rows of the 64 columns + the first half of w_hi again, rank 1: the w_lo rows + the second half of w_hi -- the
// two halves of the B operands  [w_hi | w_lo] (N = 128)  and  w_hi (N = 64)  of a pair MMA) and fetches ONE 256 KB h tile per
// step, but that tile now meets 64 gate columns instead of 32: 512 rows take 128 CTAs x 256 KB of h through L2 and
// shared memory per step where the two-tiles-per-CTA form takes twice that.  Same arithmetic per element (the same
// products accumulated in the same order), same exchange buffer, counters (8 arrivals per 128-channel chunk and step
// instead of 16) and weight image: the per-rank operand halves are gathered from the [slice][H/16][2][hi 32 | lo 32][8]
// image by 512-byte bulk copies at launch.  Hand-offs as in conv_stream.cu's pair form: bulk copies complete on a barrier
// of their own CTA and a relay lane forwards "slot full" to the leader, the leader's MMA lane frees slots and publishes
// accumulators in both CTAs with multicast commits, the gate warps of both CTAs return the accumulator to the leader.
constexpr int PNS = 64;     // gate columns of a pair slice (two neighbouring 32-column slices of the weight image)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(L_THREADS, 1) lstm_pair_kernel(const LstmTcParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_rank();
  constexpr int j = 0;   // (LTRACE)
  (void)j;
  const int H = p.H;
  constexpr int U = PNS / 4;                           // hidden units of the pair slice
  const int nchunks = H / KC;
  const int S = (int)blockIdx.x >> 1;                  // pair slice: image slices 2S, 2S + 1
  const int m = 2 * (int)blockIdx.y + rank;            // this CTA's batch tile
  const uint32_t x_bytes = (uint32_t)PNS * H * 2u;     // region X: 64 B-operand rows per 16-channel group and k-plane
  const uint32_t y_bytes = x_bytes / 2u;               // region Y: 32 rows
  const uint32_t w_bytes = x_bytes + y_bytes;
  const uint32_t chunk_split = (uint32_t)LM * KC * 2u;
  const uint32_t slot_bytes = chunk_split * 2u;
  constexpr uint32_t acc_cols = 2u * PNS;
  uint8_t* sW = smem_raw;
  uint8_t* sA = sW + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)slot_bytes * p.nslot);
  // bars: full[nslot] | ready[nslot] | empty[nslot] | acc_full | acc_empty | w_full | w_ready
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_full = bar0, bar_ready = bar0 + 8u * p.nslot, bar_empty = bar0 + 16u * p.nslot, bar_accf = bar0 + 24u * p.nslot,
                 bar_acce = bar_accf + 8u, bar_w = bar_acce + 8u, bar_wr = bar_w + 8u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.nslot + 4);

  if (tid == 0) {
    for (int s = 0; s < p.nslot; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_ready + 8u * s, 2);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 2 * GATE_WARPS);
    mbar_init(bar_w, 1);
    mbar_init(bar_wr, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(acc_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 3) {
    // this rank's halves of the B operands, gathered from the 32-column slices 2S and 2S + 1 of the weight image
    if (lane == 0) mbar_expect_tx(bar_w, w_bytes);
    __syncwarp();
    const uint8_t* img = reinterpret_cast<const uint8_t*>(p.wimg);
    const int groups = H / 16;
    for (int i = lane; i < groups * 2 * 3; i += 32) {
      const int part = i / (groups * 2), gp = i - part * groups * 2;       // part 0, 1: X from slice 2S + part; 2: Y
      const int slice = 2 * S + (part < 2 ? part : rank);
      const size_t src = (((size_t)slice * groups * 2 + gp) * 64 + (part < 2 ? rank * 32 : 0)) * 16;
      const uint32_t dst = part < 2 ? ((uint32_t)gp * 64u + (uint32_t)part * 32u) * 16u : x_bytes + (uint32_t)gp * 32u * 16u;
      bulk_g2s_notx(smem_u32(sW) + dst, img + src, 512u, bar_w);
    }
  }

  const size_t hx_tile = (size_t)2 * (H / 8) * LM * 8;
  const size_t hx_parity = hx_tile * p.m_tiles;

  if (warp == 0) {
    // ======================= TMA producer: this CTA's h tile =======================
    if (lane == 0) {
      uint32_t cc = 0;
      for (int t = 0; t < p.T; ++t) {
        const __nv_bfloat16* src = p.hx + (size_t)((p.t_base + t + 1) & 1) * hx_parity + (size_t)m * hx_tile;
        for (int c = 0; c < nchunks; ++c, ++cc) {
          if (t > 0) {
            const unsigned int target = (unsigned int)t * (unsigned int)(KC / U);
            unsigned int seen, spins = 0;
            bool first = true;
            do {
              POLL_LD(seen, p.counters + m * CNT_STRIDE + c);
              if (++spins > (1u << 26)) __trap();
#ifdef BC_TRACE
              if (c == 0 && first && seen > target - (unsigned int)(KC / U)) { first = false; LTRACE(7); }
#endif
            } while (seen < target);
            (void)first;
#ifdef BC_TRACE
            if (c == 0 && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && t >= 100 && t < 164) p.trace[(t - 100) * 8 + 6] = spins;
#endif
            POLL_FENCE();
            asm volatile("fence.proxy.async.global;" ::: "memory");
          }
          if (c == 0) LTRACE(0);
          const uint32_t slot = cc % p.nslot, use = cc / p.nslot;
          mbar_wait(bar_empty + 8u * slot, (use & 1u) ^ 1u);
          mbar_expect_tx(bar_full + 8u * slot, slot_bytes);
#pragma unroll
          for (int sp = 0; sp < 2; ++sp)
            bulk_g2s_notx(smem_u32(sA) + slot * slot_bytes + sp * chunk_split,
                          src + (size_t)sp * (H / 8) * LM * 8 + (size_t)c * (KC / 8) * LM * 8, chunk_split, bar_full + 8u * slot);
        }
        LTRACE(1);
      }
    }
  } else if (warp == 2) {
    // ======================= relay: "landed in this CTA" -> the leader's barriers =======================
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      mbar_arrive_leader(bar_wr);
      const uint32_t total = (uint32_t)p.T * (uint32_t)nchunks;
      for (uint32_t cc = 0; cc < total; ++cc) {
        const uint32_t slot = cc % p.nslot, use = cc / p.nslot;
        mbar_wait(bar_full + 8u * slot, use & 1u);
        mbar_arrive_leader(bar_ready + 8u * slot);

      }
    }
  } else if (warp == 1) {
    // ======================= MMA issue: the leader's elected lane, M = 256 over both tiles =======================
    if (rank == 0 && elect_one()) {
      mbar_wait_cluster(bar_wr, 0);
      const uint32_t a_plane = LM * 16u;
      const uint32_t hi_d = desc_hi(128u);
      const uint32_t idesc_x = idesc_bf16_m256(2 * PNS), idesc_y = idesc_bf16_m256(PNS);
      const uint32_t bx0 = desc_lo(smem_u32(sW), (uint32_t)PNS * 16u);                   // k-planes 64 rows apart
      const uint32_t by0 = desc_lo(smem_u32(sW) + x_bytes, (uint32_t)PNS * 8u);          // k-planes 32 rows apart
      const uint32_t a_g = (2u * a_plane) >> 4, bx_g = (2u * PNS * 16u) >> 4, by_g = (2u * PNS * 8u) >> 4;
      uint32_t cc = 0;
      for (int t = 0; t < p.T; ++t) {
        mbar_wait_cluster(bar_acce, ((uint32_t)t & 1u) ^ 1u);      // both CTAs' gate warps have drained step t-1
        tc_fence_after();
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const uint32_t slot = cc % p.nslot, use = cc / p.nslot;
          mbar_wait_cluster(bar_ready + 8u * slot, use & 1u);
          tc_fence_after();
          uint32_t a_lo = desc_lo(smem_u32(sA) + slot * slot_bytes, a_plane);
          uint32_t bx = bx0 + (uint32_t)c * (KC / 16) * bx_g, by = by0 + (uint32_t)c * (KC / 16) * by_g;
#pragma unroll
          for (int g = 0; g < KC / 16; ++g, a_lo += a_g, bx += bx_g, by += by_g) {
            mma2_bf16_rt(tmem_base, a_lo, bx, hi_d, hi_d, idesc_x, (g | c) ? 1u : 0u);   // a_hi x [w_hi | w_lo]
            mma2_bf16_raw<true>(tmem_base, a_lo + (chunk_split >> 4), by, hi_d, hi_d, idesc_y);   // a_lo x w_hi
          }
          umma_commit_pair(bar_empty + 8u * slot);
        }
        umma_commit_pair(bar_accf);
        LTRACE(2);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ======================= gate warps: 32 rows x 4 units x 4 gates per thread =======================
    constexpr int UPW = 4;
    const int gw = warp - 4;
    const int q = warp & 3;
    const int ug = gw >> 2;                              // units 4*ug .. 4*ug + 3 of the pair slice
    const int row = q * 32 + lane;
    const int u_glb = S * U + ug * UPW;
    const uint32_t col0 = (uint32_t)((ug >> 1) * 32 + (ug & 1) * UPW);   // column of (gate 0, first unit): 32-column slice, then gate-major
    const int b = m * LM + row;
    const bool row_ok = b < p.B;
    const float* pre_row = p.pre + (size_t)(row_ok ? b : 0) * p.pre_rows * 4 * H + u_glb;
    const size_t out_row = (size_t)(row_ok ? b : 0) * p.y_rows * H + u_glb;
    const size_t hx_off = (size_t)m * hx_tile + ((size_t)(u_glb >> 3) * LM + row) * 8 + (u_glb & 7);
    unsigned int* counter = p.counters + m * CNT_STRIDE + (S * U) / KC;
    float c_state[UPW], pg[4][UPW];
#pragma unroll
    for (int u = 0; u < UPW; ++u) c_state[u] = (p.c_state && p.t_base > 0 && row_ok) ? p.c_state[(size_t)b * H + u_glb + u] : 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 v = row_ok ? __ldcs(reinterpret_cast<const float4*>(pre_row + (size_t)g * H)) : make_float4(0.f, 0.f, 0.f, 0.f);
      pg[g][0] = v.x; pg[g][1] = v.y; pg[g][2] = v.z; pg[g][3] = v.w;
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + col0;
    for (int t = 0; t < p.T; ++t) {
      mbar_wait(bar_accf, (uint32_t)t & 1u);
      tc_fence_after();
      if (gw == 0 && lane == 0) LTRACE(3);
      uint32_t acc[2][4][UPW];
#pragma unroll
      for (int sp = 0; sp < 2; ++sp)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(acc[sp][g][0]), "=r"(acc[sp][g][1]), "=r"(acc[sp][g][2]), "=r"(acc[sp][g][3])
                       : "r"(taddr + (uint32_t)(sp * PNS + g * 8)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_acce);
      float hv[UPW];
#pragma unroll
      for (int u = 0; u < UPW; ++u) {
        float a4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) a4[g] = (__uint_as_float(acc[0][g][u]) + __uint_as_float(acc[1][g][u])) + pg[g][u];
        const float c = fmaf(sigmoid_acc(a4[1]), c_state[u], sigmoid_acc(a4[0]) * tanh_acc(a4[2]));
        c_state[u] = c;
        hv[u] = sigmoid_acc(a4[3]) * tanh_acc(c);
      }
      {
        __nv_bfloat16* dst = p.hx + (size_t)((p.t_base + t) & 1) * hx_parity + hx_off;
        __nv_bfloat16 hi[UPW], lo[UPW];
#pragma unroll
        for (int u = 0; u < UPW; ++u) {
          hi[u] = __float2bfloat16_rn(hv[u]);
          lo[u] = __float2bfloat16_rn(hv[u] - __bfloat162float(hi[u]));
        }
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(hi);
        *reinterpret_cast<uint2*>(dst + (size_t)(H / 8) * LM * 8) = *reinterpret_cast<uint2*>(lo);
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");
      if (gw == 0 && lane == 0) LTRACE(4);
      asm volatile("bar.sync 1, %0;" ::"n"(GATE_WARPS * 32) : "memory");
      if (gw == 0 && lane == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        LTRACE(5);
      }
      if (t + 1 < p.T) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 v = row_ok ? __ldcs(reinterpret_cast<const float4*>(pre_row + (size_t)(t + 1) * 4 * H + (size_t)g * H)) : make_float4(0.f, 0.f, 0.f, 0.f);
          pg[g][0] = v.x; pg[g][1] = v.y; pg[g][2] = v.z; pg[g][3] = v.w;
        }
      }
      if (row_ok) {
        const size_t o = out_row + (size_t)t * H;
        float4 v = make_float4(hv[0], hv[1], hv[2], hv[3]);
        if (p.skip) {
          const float4 s4 = __ldcs(reinterpret_cast<const float4*>(p.skip + o));
          v.x += s4.x; v.y += s4.y; v.z += s4.z; v.w += s4.w;
        }
        __stcs(reinterpret_cast<float4*>(p.y + o), v);
      }
    }
    if (p.c_state && row_ok) {
#pragma unroll
      for (int u = 0; u < UPW; ++u) p.c_state[(size_t)b * H + u_glb + u] = c_state[u];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // nobody frees TMEM / exits while the peer may still signal or read
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(acc_cols) : "memory");
}

long long* g_lstm_trace = nullptr;

struct LstmTcPlan {
  int NS, nslot, n_slices, m_tiles, split, tpc, grid_y;
  int pair, pair_nslot;          // CTA-pair form: 1 = use lstm_pair_kernel (grid 2 * n_slices / 2 x m_tiles / 2)
  size_t smem, hx_bytes, ws_bytes, pair_smem;
};

int device_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    sms = 148;
  }
  return sms;
}

bool lstm_tc_plan(int B, int H, int precision, LstmTcPlan* pl) {
  if (precision == BC_PREC_FP32) return false;
  if (H % KC != 0 || H < KC || H / KC > CNT_STRIDE) return false;
  pl->split = precision == BC_PREC_BF16X3 ? 2 : 1;
  pl->NS = pl->split == 2 ? 32 : 64;
  if ((4 * H) % pl->NS != 0) return false;
  pl->n_slices = 4 * H / pl->NS;
  pl->m_tiles = (B + LM - 1) / LM;
  const size_t w = (size_t)pl->split * pl->NS * H * 2;
  const size_t slot = (size_t)pl->split * LM * KC * 2;
  int nslot = H / KC;   // whole tile resident if it fits
  while (nslot > 2 && w + nslot * slot + 1024 > 225 * 1024) --nslot;
  if (w + nslot * slot + 1024 > 225 * 1024) return false;
  pl->nslot = nslot;
  pl->smem = w + nslot * slot + (2 * (H / KC > nslot ? H / KC : nslot) + 2 * MAX_TPC + 1) * 8 + 64;   // barriers for up to H / KC slots (compact image)
  pl->hx_bytes = (size_t)2 * pl->m_tiles * pl->split * (H / 8) * LM * 8 * 2;
  pl->ws_bytes = pl->hx_bytes + (size_t)pl->m_tiles * CNT_STRIDE * sizeof(unsigned int) + 256;
  // one batch tile per CTA while the tiles fit side by side (the latency-optimal shape); otherwise two independent
  // tiles per CTA, interleaved step by step
  const int side_by_side = device_sms() / pl->n_slices;
  pl->tpc = (pl->m_tiles > side_by_side && bc::policy().lstm_pingpong) ? 2 : 1;
  pl->grid_y = (pl->m_tiles + pl->tpc - 1) / pl->tpc;
  // CTA pairs where one tile per CTA no longer fits side by side: an even number of tiles, 64-column pair slices
  pl->pair = 0; pl->pair_nslot = 0; pl->pair_smem = 0;
  if (pl->split == 2 && bc::policy().lstm_pair && (pl->m_tiles > side_by_side || bc::policy().lstm_pair >= 2) && pl->m_tiles % 2 == 0 && pl->n_slices % 2 == 0 &&
      (pl->n_slices / 2) * pl->m_tiles <= device_sms()) {
    const size_t wp = (size_t)(PNS + PNS / 2) * H * 2;
    int ns = H / KC;
    while (ns > 2 && wp + ns * slot + 512 > 227 * 1024) --ns;
    if (wp + ns * slot + 512 <= 227 * 1024) {
      pl->pair = 1;
      pl->pair_nslot = ns;
      pl->pair_smem = wp + ns * slot + (3 * ns + 4) * 8 + 64;
    }
  }
  return true;
}

}  // namespace

extern "C" size_t bc_lstm_tc_workspace_bytes(int B, int H, int precision) {
  LstmTcPlan pl;
  if (B <= 0 || H <= 0 || !lstm_tc_plan(B, H, precision, &pl)) return 0;
  return pl.ws_bytes;
}

extern "C" int bc_lstm_tc_slice_cols(int precision) {
  return precision == BC_PREC_BF16 ? 64 : (precision == BC_PREC_BF16X3 ? 32 : 0);
}

extern "C" int bc_lstm_tc_max_batch(int H, int precision) {
  LstmTcPlan pl;
  if (!lstm_tc_plan(128, H, precision, &pl)) return 0;
  const int m_tiles = device_sms() / pl.n_slices;
  return m_tiles * LM * (bc::policy().lstm_pingpong ? MAX_TPC : 1);
}

static int lstm_tc_launch(const float* pre, const void* w_image, const float* skip, float* y, void* workspace,
                          float* c_state, int B, int T, int pre_rows, int y_rows, int t_base, int H, int precision,
                          bc_stream_t s) {
  BC_REQUIRE(pre && w_image && y && workspace, "lstm_tc: null pointer");
  BC_REQUIRE(B > 0 && T > 0 && H > 0 && pre_rows >= T && y_rows >= T && t_base >= 0, "lstm_tc: bad shape B=%d T=%d H=%d", B, T, H);
  BC_REQUIRE(t_base == 0 || c_state, "lstm_tc: continuing a sequence needs the carried cell state");
  LstmTcPlan pl;
  if (!lstm_tc_plan(B, H, precision, &pl))
    return bc::fail(BC_EUNSUPPORTED, "lstm_tc: H=%d precision=%d has no tensor-core plan", H, precision);
  BC_REQUIRE(bc::aligned16(w_image) && bc::aligned16(workspace) && bc::aligned16(y) && bc::aligned16(pre) && (!skip || bc::aligned16(skip)),
             "lstm_tc: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) return bc::fail(BC_ENODEVICE, "lstm_tc: device does not support cooperative launch");
  if (pl.n_slices * pl.grid_y > sms)
    return bc::fail(BC_EUNSUPPORTED, "lstm_tc: B=%d needs %d co-resident CTAs, device has %d SMs (split the batch)", B, pl.n_slices * pl.grid_y, sms);
  LstmTcParams p;
  p.trace = g_lstm_trace;
  p.pre = pre; p.wimg = reinterpret_cast<const uint4*>(w_image); p.skip = skip; p.y = y;
  p.hx = reinterpret_cast<__nv_bfloat16*>(workspace);
  const size_t hx_pad = (pl.hx_bytes + 127) & ~size_t(127);
  p.counters = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(workspace) + hx_pad);
  p.B = B; p.T = T; p.H = H; p.NS = pl.NS; p.nslot = pl.nslot; p.n_slices = pl.n_slices; p.m_tiles = pl.m_tiles;
  p.pre_rows = pre_rows; p.y_rows = y_rows; p.t_base = t_base; p.c_state = c_state;
  p.rows = (pl.m_tiles == 1 && !pl.pair && bc::policy().lstm_compact) ? ((B + 7) / 8) * 8 : LM;
  p.split_stride = (uint32_t)LM * KC * 2u;
  p.ring_bytes = (uint32_t)pl.nslot * p.split_stride * (uint32_t)pl.split;
  p.whole = 0;
  if (p.rows < LM && H / KC <= 4) {
    // compact image that fits the ring as ONE piece per split: a single slot, hi image then lo image.  The MMA reads 128 rows
    // per plane whatever R is: 4 KB of the ring stay free behind the images for that over-read.
    const uint32_t split_bytes = (uint32_t)(H / 8) * (uint32_t)p.rows * 16u;
    const uint32_t stride_c = (split_bytes + 1023u) & ~1023u;
    if ((size_t)stride_c * pl.split + 4096 <= p.ring_bytes) { p.split_stride = stride_c; p.nslot = 1; p.whole = 1; }
  }
  p.idesc_wide = bc::tc::idesc_bf16_m128(pl.split * pl.NS);
  p.idesc_ns = bc::tc::idesc_bf16_m128(pl.NS);
  // a new sequence starts from h = 0 (the exchange buffer) and fresh step counters; a continued one keeps h and only
  // restarts the counters (they count the steps of THIS launch)
  cudaError_t e = t_base == 0 ? cudaMemsetAsync(workspace, 0, pl.ws_bytes, st)
                              : cudaMemsetAsync(reinterpret_cast<uint8_t*>(workspace) + hx_pad, 0, pl.ws_bytes - hx_pad, st);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaMemsetAsync(lstm_tc)");
  if (pl.pair) {
    static bool configured[64] = {false};
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      e = cudaFuncSetAttribute(lstm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.pair_smem);
      if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(lstm_pair)");
      // every CTA waits for the others: all clusters must be co-resident
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(pl.n_slices, pl.m_tiles / 2); cfg.blockDim = dim3(L_THREADS); cfg.dynamicSmemBytes = pl.pair_smem;
      int clusters = 0;
      e = cudaOccupancyMaxActiveClusters(&clusters, lstm_pair_kernel, &cfg);
      if (e != cudaSuccess) return bc::cuda_check(e, "cudaOccupancyMaxActiveClusters(lstm_pair)");
      if (clusters * 2 < pl.n_slices * (pl.m_tiles / 2))
        return bc::fail(BC_EUNSUPPORTED, "lstm_tc(pair): %d CTA pairs do not fit the device side by side (%d)", pl.n_slices * pl.m_tiles / 4, clusters);
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    p.nslot = pl.pair_nslot;
    lstm_pair_kernel<<<dim3(pl.n_slices, pl.m_tiles / 2), L_THREADS, pl.pair_smem, st>>>(p);
    BC_LAUNCH_CHECK("lstm_pair_kernel");
    return BC_OK;
  }
  void* kern = nullptr;
  if (pl.split == 2) kern = pl.tpc == 2 ? (void*)lstm_tc_kernel<2, 2, 2> : (void*)lstm_tc_kernel<2, 2, 1>;
  else               kern = pl.tpc == 2 ? (void*)lstm_tc_kernel<1, 4, 2> : (void*)lstm_tc_kernel<1, 4, 1>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(lstm_tc)");
  void* args[] = {(void*)&p};
  e = cudaLaunchCooperativeKernel(kern, dim3(pl.n_slices, pl.grid_y), dim3(L_THREADS), args, pl.smem, st);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaLaunchCooperativeKernel(lstm_tc)");
  return BC_OK;
}

extern "C" int bc_lstm_tc_recurrent_fwd(const float* pre, const void* w_image, const float* skip, float* y,
                                        void* workspace, int B, int T, int H, int precision, bc_stream_t s) {
  return lstm_tc_launch(pre, w_image, skip, y, workspace, nullptr, B, T, T, T, 0, H, precision, s);
}

extern "C" int bc_lstm_tc_recurrent_chunk_fwd(const float* pre, const void* w_image, const float* skip, float* y,
                                              void* workspace, float* c_state, int B, int T_chunk, int pre_rows, int y_rows,
                                              int t_base, int H, int precision, bc_stream_t s) {
  BC_REQUIRE(c_state, "lstm_tc(chunk): null cell-state buffer");
  return lstm_tc_launch(pre, w_image, skip, y, workspace, c_state, B, T_chunk, pre_rows, y_rows, t_base, H, precision, s);
}

extern "C" int bc_lstm_tc_ctas(int B, int H, int precision) {
  LstmTcPlan pl;
  if (B <= 0 || !lstm_tc_plan(B, H, precision, &pl)) return 0;
  return pl.pair ? (pl.n_slices / 2) * pl.m_tiles : pl.n_slices * pl.grid_y;
}

// debug hook (not part of the product path): device buffer of 64*8 int64 receiving clock64 stamps of CTA (0,0)
extern "C" int bc_debug_set_lstm_trace(void* device_buffer) {
  g_lstm_trace = reinterpret_cast<long long*>(device_buffer);
  return BC_OK;
}
