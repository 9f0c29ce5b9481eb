// LSTM recurrence on the tensor cores (tcgen05), sm_100a -- BC_PREC_BF16 / BC_PREC_BF16X3.
//
//   G_t[B x 4H] = pre_t + h_{t-1}[B x H] * W_hh^T ,  gates i,f,g,o   (nn.LSTM inside ResLSTM, vq/module.py:143-167)
//
// Decomposition.  CTA (m, n) owns 128 batch rows (m) and NS gate columns (n) = U = NS/4 hidden units with
// all four gates; its W_hh slice [NS x H] (bf16 hi [, lo]) stays resident in shared memory for the whole
// sequence.  Every time step:
//   TMA thread   waits until all n-slices of its m-tile have published h_{t-1}, then streams the
//                bf16 h tile [128 x H] (stored in HBM/L2 directly in the UMMA K-major image) in K chunks
//                through a shared-memory ring (cp.async.bulk + mbarrier);
//   MMA thread   H/16 [x3] tcgen05.mma (M=128, N=NS) into one TMEM accumulator, chunk by chunk as they land;
//   16 gate warps  tcgen05.ld their (32 rows x UPW units x 4 gates) patch, add the pre-activation (prefetched
//                from HBM during the MMA phase), apply the gates with c kept in registers, write y (fp32,
//                + skip) and publish h_t as bf16 hi[/lo] into the exchange buffer, then bump the m-tile's
//                step counter (release); no grid-wide barrier -- only the CTAs that share batch rows wait
//                for each other.
// The kernel is launched cooperatively (all CTAs must be co-resident because they wait on each other).
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace bc::tc;

constexpr int LM = 128;         // batch rows per CTA
constexpr int KC = 128;         // K elements per streamed chunk
constexpr int GATE_WARPS = 16;
constexpr int CNT_STRIDE = 16;  // counters per batch tile (>= H / KC)
constexpr int L_THREADS = (4 + GATE_WARPS) * 32;   // warp 0: TMA, warp 1: MMA, warps 2-3 idle, warps 4-19: gates

struct LstmTcParams {
  const float* pre;        // [B][T][4H]
  const uint4* wimg;       // [n_slices][split][H/16][2][NS][8] bf16
  const float* skip;       // [B][T][H] or NULL
  float* y;                // [B][T][H]
  __nv_bfloat16* hx;       // [2][m_tiles][split][H/8][128][8]
  unsigned int* counters;  // [m_tiles][CNT_STRIDE]: one step counter per (batch tile, K chunk of h)
  int B, T, H, NS, nslot, n_slices;
  uint32_t idesc;
  long long* trace;   // debug: [step < 64][8] clock64 stamps of CTA (0,0), steps 100.. (NULL = off)
};

#ifdef BC_TRACE
#define LTRACE(ev) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && t >= 100 && t < 164) p.trace[(t - 100) * 8 + (ev)] = clock64(); } while (0)
#else
#define LTRACE(ev) do { } while (0)
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

template <int SPLIT, int UPW>
__global__ void __launch_bounds__(L_THREADS, 1) lstm_tc_kernel(const LstmTcParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, NS = p.NS, U = NS / 4;
  const int nchunks = H / KC;
  const int n = blockIdx.x, m = blockIdx.y;
  const uint32_t w_split = (uint32_t)NS * H * 2u;
  const uint32_t chunk_split = (uint32_t)LM * KC * 2u;           // one chunk of the h tile, one split: 32 KB
  const uint32_t slot_bytes = chunk_split * SPLIT;
  uint8_t* sW = smem_raw;
  uint8_t* sA = sW + w_split * SPLIT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)slot_bytes * p.nslot);
  // bars: full[nslot] | empty[nslot] | acc_full | acc_empty | w_full
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8u * p.nslot, bar_accf = bar0 + 16u * p.nslot,
                 bar_acce = bar_accf + 8u, bar_w = bar_accf + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.nslot + 3);

  if (tid == 0) {
    for (int s = 0; s < p.nslot; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, GATE_WARPS);
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t wbytes = w_split * SPLIT;
    mbar_expect_tx(bar_w, wbytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) + (size_t)n * wbytes;
    for (uint32_t off = 0; off < wbytes; off += 32768u)
      bulk_g2s_notx(smem_u32(sW) + off, src + off, min(32768u, wbytes - off), bar_w);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(NS < 32 ? 32 : NS)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const size_t hx_tile = (size_t)SPLIT * (H / 8) * LM * 8;        // bf16 elements of one m-tile image (all splits)
  const size_t hx_parity = hx_tile * gridDim.y;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t cc = 0;
      for (int t = 0; t < p.T; ++t) {
        // K chunk c of h_{t-1} is written by the KC/U n-slices that own its hidden units: each chunk is fetched as soon
        // as ITS producers have published (per-chunk step counters), so the copies and MMAs of the early chunks overlap
        // the stragglers of the later ones instead of waiting for the slowest of all n-slices
        const __nv_bfloat16* src = p.hx + (size_t)((t + 1) & 1) * hx_parity + (size_t)m * hx_tile;
        for (int c = 0; c < nchunks; ++c, ++cc) {
          if (t > 0) {
            const unsigned int target = (unsigned int)t * (unsigned int)(KC / U);
            unsigned int seen;
            unsigned int spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.counters + m * CNT_STRIDE + c) : "memory");
              if (++spins > (1u << 26)) __trap();
            } while (seen < target);
            asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes of other CTAs -> async-proxy reads
          }
          if (c == 0) LTRACE(0);
          const uint32_t slot = cc % p.nslot, use = cc / p.nslot;
          mbar_wait(bar_empty + 8u * slot, (use & 1u) ^ 1u);
          mbar_expect_tx(bar_full + 8u * slot, slot_bytes);
#pragma unroll
          for (int sp = 0; sp < SPLIT; ++sp)
            bulk_g2s_notx(smem_u32(sA) + slot * slot_bytes + sp * chunk_split,
                          src + (size_t)sp * (H / 8) * LM * 8 + (size_t)c * (KC / 8) * LM * 8, chunk_split, bar_full + 8u * slot);
        }
        LTRACE(1);
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (warp-uniform loop, elected lane issues) =======================
    {
      mbar_wait(bar_w, 0);
      const uint32_t a_plane = LM * 16u;
      const uint32_t hi_d = desc_hi(128u);
      const uint32_t w_lo0 = desc_lo(smem_u32(sW), (uint32_t)NS * 16u);
      uint32_t cc = 0;
      for (int t = 0; t < p.T; ++t) {
        mbar_wait(bar_acce, ((uint32_t)t & 1u) ^ 1u);   // gate warps have drained the accumulator of step t-1
        tc_fence_after();
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const uint32_t slot = cc % p.nslot, use = cc / p.nslot;
          mbar_wait(bar_full + 8u * slot, use & 1u);
          tc_fence_after();
          uint32_t a_lo = desc_lo(smem_u32(sA) + slot * slot_bytes, a_plane);
          uint32_t b_lo = w_lo0 + (((uint32_t)c * (KC / 16) * NS * 32u) >> 4);
          const uint32_t a_g = (2u * a_plane) >> 4, b_g = ((uint32_t)NS * 32u) >> 4;
#pragma unroll
          for (int g = 0; g < KC / 16; ++g, a_lo += a_g, b_lo += b_g) {
            if (g == 0 && c == 0) mma_bf16_lohi<false>(tmem_base, a_lo, b_lo, hi_d, hi_d, p.idesc);
            else                  mma_bf16_lohi<true>(tmem_base, a_lo, b_lo, hi_d, hi_d, p.idesc);
            if (SPLIT == 2) {
              mma_bf16_lohi<true>(tmem_base, a_lo, b_lo + (w_split >> 4), hi_d, hi_d, p.idesc);
              mma_bf16_lohi<true>(tmem_base, a_lo + (chunk_split >> 4), b_lo, hi_d, hi_d, p.idesc);
            }
          }
          if (elect_one()) umma_commit(bar_empty + 8u * slot);
          __syncwarp();
        }
        if (elect_one()) umma_commit(bar_accf);
        __syncwarp();
        if (lane == 0) LTRACE(2);
      }
    }
  } else if (warp >= 4) {
    // ======================= gate warps =======================
    const int gw = warp - 4;
    const int q = warp & 3;                 // TMEM lane quarter
    const int ug = gw >> 2;                 // unit group within the slice
    const int row = q * 32 + lane;          // row within the m-tile
    const int b = m * LM + row;
    const bool row_ok = b < p.B;
    const int u_loc = ug * UPW;             // first unit (within the slice) of this thread
    const int u_glb = n * U + u_loc;        // first hidden unit (global index)
    float c_state[UPW];
#pragma unroll
    for (int j = 0; j < UPW; ++j) c_state[j] = 0.f;
    const float* pre_row = p.pre + (size_t)(row_ok ? b : 0) * p.T * 4 * H + u_glb;
    const size_t out_row = (size_t)(row_ok ? b : 0) * p.T * H + u_glb;
    // exchange-buffer position of this thread's units: plane = u_glb / 8, element = u_glb % 8
    const size_t hx_off = (size_t)m * hx_tile + ((size_t)(u_glb >> 3) * LM + row) * 8 + (u_glb & 7);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float pg[4][UPW];
    // prefetch pre-activations of step 0
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int j = 0; j < UPW; ++j) pg[g][j] = row_ok ? __ldcs(pre_row + (size_t)g * H + j) : 0.f;
    for (int t = 0; t < p.T; ++t) {
      mbar_wait(bar_accf, (uint32_t)t & 1u);
      tc_fence_after();
      if (gw == 0 && lane == 0) LTRACE(3);
      uint32_t acc[4][UPW];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if constexpr (UPW == 4) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(acc[g][0]), "=r"(acc[g][1]), "=r"(acc[g][2]), "=r"(acc[g][3])
                       : "r"(taddr + (uint32_t)(g * U + u_loc)));
        } else {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                       : "=r"(acc[g][0]), "=r"(acc[g][1])
                       : "r"(taddr + (uint32_t)(g * U + u_loc)));
        }
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce);
      float hv[UPW];
#pragma unroll
      for (int j = 0; j < UPW; ++j) {
        const float gi = __uint_as_float(acc[0][j]) + pg[0][j];
        const float gf = __uint_as_float(acc[1][j]) + pg[1][j];
        const float gg = __uint_as_float(acc[2][j]) + pg[2][j];
        const float go = __uint_as_float(acc[3][j]) + pg[3][j];
        const float c = fmaf(sigmoid_acc(gf), c_state[j], sigmoid_acc(gi) * tanh_acc(gg));
        c_state[j] = c;
        hv[j] = sigmoid_acc(go) * tanh_acc(c);
      }
      // publish h_t (bf16 hi[/lo]) for the next step's MMA, in the UMMA K-major image -- FIRST: every other CTA of
      // this batch tile waits for it.  The fences below wait for all earlier memory operations of the thread, so
      // nothing else (output store, skip load, next step's pre-activation loads) may be in flight before them.
      {
        __nv_bfloat16* dst = p.hx + (size_t)(t & 1) * hx_parity + hx_off;
        __nv_bfloat16 hi[UPW], lo[UPW];
#pragma unroll
        for (int j = 0; j < UPW; ++j) {
          hi[j] = __float2bfloat16_rn(hv[j]);
          lo[j] = __float2bfloat16_rn(hv[j] - __bfloat162float(hi[j]));
        }
        if constexpr (UPW == 4) {
          *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(hi);
          if (SPLIT == 2) *reinterpret_cast<uint2*>(dst + (size_t)(H / 8) * LM * 8) = *reinterpret_cast<uint2*>(lo);
        } else {
          *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<uint32_t*>(hi);
          if (SPLIT == 2) *reinterpret_cast<uint32_t*>(dst + (size_t)(H / 8) * LM * 8) = *reinterpret_cast<uint32_t*>(lo);
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
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.counters + m * CNT_STRIDE + (n * U) / KC) : "memory");
        LTRACE(5);
      }
      // off the critical path (overlaps the other CTAs' publishes, the h copies and the next MMA phase):
      // next step's pre-activations and this step's output row
      if (t + 1 < p.T) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int j = 0; j < UPW; ++j) pg[g][j] = row_ok ? __ldcs(pre_row + (size_t)(t + 1) * 4 * H + (size_t)g * H + j) : 0.f;
      }
      if (row_ok) {
        const size_t o = out_row + (size_t)t * H;
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(NS < 32 ? 32 : NS)) : "memory");
  }
}

long long* g_lstm_trace = nullptr;

struct LstmTcPlan {
  int NS, nslot, n_slices, m_tiles, split;
  size_t smem, hx_bytes, ws_bytes;
};

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
  pl->smem = w + nslot * slot + (2 * nslot + 3) * 8 + 64;
  pl->hx_bytes = (size_t)2 * pl->m_tiles * pl->split * (H / 8) * LM * 8 * 2;
  pl->ws_bytes = pl->hx_bytes + (size_t)pl->m_tiles * CNT_STRIDE * sizeof(unsigned int) + 256;
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
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    sms = 148;
  }
  const int m_tiles = sms / pl.n_slices;
  return m_tiles * LM;
}

extern "C" int bc_lstm_tc_recurrent_fwd(const float* pre, const void* w_image, const float* skip, float* y,
                                        void* workspace, int B, int T, int H, int precision, bc_stream_t s) {
  BC_REQUIRE(pre && w_image && y && workspace, "lstm_tc: null pointer");
  BC_REQUIRE(B > 0 && T > 0 && H > 0, "lstm_tc: bad shape B=%d T=%d H=%d", B, T, H);
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
  if (pl.n_slices * pl.m_tiles > sms)
    return bc::fail(BC_EUNSUPPORTED, "lstm_tc: B=%d needs %d co-resident CTAs, device has %d SMs (split the batch)", B, pl.n_slices * pl.m_tiles, sms);
  LstmTcParams p;
  p.trace = g_lstm_trace;
  p.pre = pre; p.wimg = reinterpret_cast<const uint4*>(w_image); p.skip = skip; p.y = y;
  p.hx = reinterpret_cast<__nv_bfloat16*>(workspace);
  p.counters = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(workspace) + ((pl.hx_bytes + 127) & ~size_t(127)));
  p.B = B; p.T = T; p.H = H; p.NS = pl.NS; p.nslot = pl.nslot; p.n_slices = pl.n_slices;
  p.idesc = bc::tc::idesc_bf16_m128(pl.NS);
  cudaError_t e = cudaMemsetAsync(workspace, 0, pl.ws_bytes, st);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaMemsetAsync(lstm_tc)");
  void* kern = pl.split == 2 ? (void*)lstm_tc_kernel<2, 2> : (void*)lstm_tc_kernel<1, 4>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(lstm_tc)");
  void* args[] = {(void*)&p};
  e = cudaLaunchCooperativeKernel(kern, dim3(pl.n_slices, pl.m_tiles), dim3(L_THREADS), args, pl.smem, st);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaLaunchCooperativeKernel(lstm_tc)");
  return BC_OK;
}

// debug hook (not part of the product path): device buffer of 64*8 int64 receiving clock64 stamps of CTA (0,0)
extern "C" int bc_debug_set_lstm_trace(void* device_buffer) {
  g_lstm_trace = reinterpret_cast<long long*>(device_buffer);
  return BC_OK;
}
