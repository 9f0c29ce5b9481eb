// tcgen05 implicit-GEMM 1-D convolution for sm_100a (BC_PREC_BF16 / BC_PREC_BF16X3).
//
//   D[128 time steps x N_t channels] (fp32, TMEM) += A[128 x 16] (bf16, smem) * B[N_t x 16]^T (bf16, smem)
//
// per (tap k, 16-input-channel group g).  The conv taps are NOT materialised (no im2col):
// the activation slab of a CTA tile is staged once per input-channel chunk in the UMMA
// "K-major, no swizzle" canonical layout
//        [8-channel plane][stride phase][row][8 x bf16 = 16 B]
// in which consecutive rows are 16 bytes apart, so tap k of a dilated / strided conv is just
// the SAME slab read through a descriptor whose start address is advanced by
// (k*dil % stride) * rows_per_phase + (k*dil / stride) rows.  SnakeBeta, the fp32 -> bf16
// (hi [, lo]) split and zero padding happen while the slab is staged from HBM, so the
// activation makes no extra HBM round trip; bias / residual / tanh are applied when the
// accumulator is read back from TMEM (tcgen05.ld).
//
// BC_PREC_BF16X3: a = a_hi + a_lo, w = w_hi + w_lo (bf16 each); the product is accumulated as
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi in fp32 -- ~16 mantissa bits, fp32-class parity on the
// tensor cores at 3 MMAs per term.
//
// Weights arrive pre-packed (host side, once per load) as bf16 blocks in exactly the smem
// image the kernel needs: [n_tile][chunk][split][tap][group][2 k-planes][N_t][8].
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace {
using namespace bc::tc;

constexpr int BM = 128;
constexpr int TC_THREADS = 256;
#ifndef TC_MIN_CTAS
#define TC_MIN_CTAS 2
#endif

struct TcParams {
  const float* x;
  const uint4* wpk;
  const float* bias;
  const float* sa;
  const float* sib;
  const float* res;
  float* y;
  int B, T_in, C_in, T_out, C_out, K, stride, dil, pad_left;
  int y_rows, y_tstride, y_toffset, flags;
  int n_tile, gpc, nchunks, rpp, slab_rows, split, tmem_cols;
  uint32_t idesc;
  int variant;
  // fused ResidualUnit tail: y = res + W2 * snake2(conv(x) + bias) + bias2   (1x1 conv, C_in == C_out == n_tile)
  int fuse2;
  const uint4* w2pk;
  const float* bias2;
  const float* sa2;
  const float* sib2;
  uint32_t region1_bytes;  // [A slab | B image], re-used for the bf16 intermediate of the fused tail
  // persistent mode (single chunk, single n-tile): weights loaded once, CTA loops over (item, time-tile)
  int persist, tiles_per_item, total_tiles;
};

constexpr int STAGE_BATCH = 4;  // slab items whose global loads are issued back to back per thread

template <int SPLIT, bool FUSE2>
__global__ void __launch_bounds__(TC_THREADS, TC_MIN_CTAS) conv1d_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int planes = 2 * p.gpc;
  const uint32_t plane_bytes = (uint32_t)p.stride * p.rpp * 16u;  // one 8-channel plane of the slab
  const uint32_t a_split_bytes = planes * plane_bytes;
  const uint32_t a_bytes = a_split_bytes * SPLIT;
  const uint32_t b_split_bytes = (uint32_t)p.K * p.gpc * p.n_tile * 32u;
  const uint32_t b_bytes = b_split_bytes * SPLIT;
  uint8_t* sA = smem_raw;
  uint8_t* sB = smem_raw + ((a_bytes + 127u) & ~127u);
  uint8_t* sB2 = smem_raw + p.region1_bytes;
  const uint32_t b2_split_bytes = FUSE2 ? (uint32_t)p.n_tile * p.n_tile * 2u : 0u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + ((b2_split_bytes * SPLIT + 127u) & ~127u));
  const uint32_t mbar = smem_u32(bars), bbar = smem_u32(bars + 1), b2bar = smem_u32(bars + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int nt = blockIdx.y;
  const bool snake = (p.flags & BC_CONV_SNAKE_IN) != 0;

  // ---- one-time setup: mbarriers, TMEM allocation, first weight image in flight ----
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bbar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b2bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    bulk_g2s(smem_u32(sB), p.wpk + (size_t)nt * p.nchunks * (b_bytes / 16), b_bytes, bbar);
    if (FUSE2) bulk_g2s(smem_u32(sB2), p.w2pk, b2_split_bytes * SPLIT, b2bar);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  uint32_t phase = 0, bphase = 0;
  const int items = planes * p.slab_rows;  // 16-byte slab items per split
  const int pshift = 31 - __clz(planes);   // planes = 2*gpc is a power of two
  const int pl = tid & (planes - 1);        // ... and divides the block size: a thread always serves the same plane
  // persistent mode: grid-stride over (item, time-tile); otherwise exactly one tile per CTA
  const int tile_step = p.persist ? (int)gridDim.x : p.total_tiles;
  for (int tile = p.persist ? (int)blockIdx.x : (int)(blockIdx.z * p.tiles_per_item + blockIdx.x); tile < p.total_tiles;
       tile += tile_step) {
  const int b = tile / p.tiles_per_item;
  const int t0 = (tile - b * p.tiles_per_item) * BM;
  const int g0 = t0 * p.stride - p.pad_left;
  const float* xb = p.x + (size_t)b * p.T_in * p.C_in;
  for (int ch = 0; ch < p.nchunks; ++ch) {
    const int ci0 = ch * p.gpc * 16;
    // ---- stage A: x (fp32, HBM) -> snake -> bf16 hi[/lo] -> canonical K-major slab ----
    float4 a0, a1, b0, b1;
    if (snake) {
      a0 = __ldg(reinterpret_cast<const float4*>(p.sa + ci0 + pl * 8));
      a1 = __ldg(reinterpret_cast<const float4*>(p.sa + ci0 + pl * 8) + 1);
      b0 = __ldg(reinterpret_cast<const float4*>(p.sib + ci0 + pl * 8));
      b1 = __ldg(reinterpret_cast<const float4*>(p.sib + ci0 + pl * 8) + 1);
    }
    const float* xcol = xb + ci0 + pl * 8;
    for (int i0 = tid; i0 < items; i0 += TC_THREADS * STAGE_BATCH) {
      float4 lo4[STAGE_BATCH], hi4[STAGE_BATCH];
#pragma unroll
      for (int j = 0; j < STAGE_BATCH; ++j) {   // all loads first: STAGE_BATCH x 32 B in flight per thread
        const int i = i0 + j * TC_THREADS;
        const int g = g0 + (i >> pshift);
        if (i < items && g >= 0 && g < p.T_in) {
          const float4* src = reinterpret_cast<const float4*>(xcol + (size_t)g * p.C_in);
          lo4[j] = __ldg(src);
          hi4[j] = __ldg(src + 1);
        } else {
          lo4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          hi4[j] = lo4[j];
        }
      }
#pragma unroll
      for (int j = 0; j < STAGE_BATCH; ++j) {
        const int i = i0 + j * TC_THREADS;
        if (i < items) {
          const int r = i >> pshift;
          float v[8] = {lo4[j].x, lo4[j].y, lo4[j].z, lo4[j].w, hi4[j].x, hi4[j].y, hi4[j].z, hi4[j].w};
          if (snake) snake8<SPLIT>(v, a0, a1, b0, b1);   // snake(0) == 0: padding rows stay zero
          const int ph = r % p.stride, rr = r / p.stride;
          split_store<SPLIT>(v, sA + (size_t)pl * plane_bytes + ((size_t)ph * p.rpp + rr) * 16, a_split_bytes);
        }
      }
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- warp 0 runs the issue loop (warp-uniform), one elected lane issues each MMA and the commit ----
    if (warp == 0) {
      mbar_wait(bbar, bphase);   // this chunk's weight image has landed (bulk copy)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // descriptors are built from kernel parameters and loop counters only (uniform registers); one elected
      // lane issues the whole chunk back to back
      const uint32_t hi_d = desc_hi(128u);
      const uint32_t smem0 = smem_u32(smem_raw);
      const uint32_t a_lo0 = desc_lo(smem0, plane_bytes);
      const uint32_t b_lo0 = desc_lo(smem0 + ((a_bytes + 127u) & ~127u), (uint32_t)p.n_tile * 16u);
      const uint32_t a_g = (2u * plane_bytes) >> 4, b_g = ((uint32_t)p.n_tile * 32u) >> 4;
      const uint32_t a_sp = a_split_bytes >> 4, b_sp = b_split_bytes >> 4;
      if (elect_one()) {
        uint32_t b_lo = b_lo0;
        uint32_t acc = ch == 0 ? 0u : 1u;
        for (int k = 0; k < p.K; ++k) {
          const int sh = k * p.dil;
          uint32_t a_lo = a_lo0 + (uint32_t)(sh % p.stride) * p.rpp + (uint32_t)(sh / p.stride);
          for (int g = 0; g < p.gpc; ++g, a_lo += a_g, b_lo += b_g) {
            mma_bf16_raw_rt(tmem_base, a_lo, b_lo, hi_d, hi_d, p.idesc, acc);
            acc = 1u;
            if (SPLIT == 2) {
              mma_bf16_raw<true>(tmem_base, a_lo, b_lo + b_sp, hi_d, hi_d, p.idesc);   // a_hi * w_lo
              mma_bf16_raw<true>(tmem_base, a_lo + a_sp, b_lo, hi_d, hi_d, p.idesc);   // a_lo * w_hi
            }
          }
        }
        umma_commit(mbar);
      }
      __syncwarp();
    }
    if (!p.persist) bphase ^= 1u;   // persistent mode: the single weight image was loaded once (phase 0 stays complete)
    // everyone waits until the tensor core has consumed this chunk's smem
    mbar_wait(mbar, phase);
    phase ^= 1u;
    if (tid == 0 && ch + 1 < p.nchunks)   // next chunk's weights stream in while the slab is being staged
      bulk_g2s(smem_u32(sB), p.wpk + ((size_t)nt * p.nchunks + ch + 1) * (b_bytes / 16), b_bytes, bbar);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const int q = warp & 3;            // TMEM lane quarter this warp may access
  const int half = warp >> 2;        // column half
  const int ncols = p.n_tile / 2;
  const int col0 = half * ncols;
  const int t = t0 + q * 32 + lane;
  const bool row_ok = t < p.T_out;
  uint32_t acc_col = 0;

  if (FUSE2) {
    // ---- fused ResidualUnit tail: h = snake2(acc + bias) -> bf16 smem tile -> 1x1 conv on the tensor core ----
    const uint32_t a2_plane = BM * 16u;
    const uint32_t a2_split = (uint32_t)(p.n_tile / 8) * a2_plane;
    uint8_t* sA2 = smem_raw;  // aliases [A | B]: every MMA that read them has completed
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      const int n8 = min(4, (ncols - c0) / 8);
      uint32_t r[32];
      tmem_load(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + c0), n8, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < n8) {
          const int c = col0 + c0 + 8 * j;
          const float4 bi0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
          const float4 bi1 = __ldg(reinterpret_cast<const float4*>(p.bias + c) + 1);
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.sa2 + c));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.sa2 + c) + 1);
          const float4 i0 = __ldg(reinterpret_cast<const float4*>(p.sib2 + c));
          const float4 i1 = __ldg(reinterpret_cast<const float4*>(p.sib2 + c) + 1);
          float v[8];
            acc_bias8(r + 8 * j, bi0, bi1, v);
          snake8<SPLIT>(v, s0, s1, i0, i1);
          split_store<SPLIT>(v, sA2 + (size_t)(c / 8) * a2_plane + (size_t)(q * 32 + lane) * 16, a2_split);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    acc_col = (uint32_t)p.tmem_cols / 2;
    if (warp == 0) {
      mbar_wait(b2bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t hi_d = desc_hi(128u);
      const uint32_t smem0 = smem_u32(smem_raw);
      const uint32_t a_lo0 = desc_lo(smem0, a2_plane), b_lo0 = desc_lo(smem0 + p.region1_bytes, (uint32_t)p.n_tile * 16u);
      const uint32_t a_g = (2u * a2_plane) >> 4, b_g = ((uint32_t)p.n_tile * 32u) >> 4;
      if (elect_one()) {
        uint32_t a_lo = a_lo0, b_lo = b_lo0;
        for (int g = 0; g < p.n_tile / 16; ++g, a_lo += a_g, b_lo += b_g) {
          mma_bf16_raw_rt(tmem_base + acc_col, a_lo, b_lo, hi_d, hi_d, p.idesc, g ? 1u : 0u);
          if (SPLIT == 2) {
            mma_bf16_raw<true>(tmem_base + acc_col, a_lo, b_lo + (b2_split_bytes >> 4), hi_d, hi_d, p.idesc);
            mma_bf16_raw<true>(tmem_base + acc_col, a_lo + (a2_split >> 4), b_lo, hi_d, hi_d, p.idesc);
          }
        }
        umma_commit(mbar);
      }
      __syncwarp();
    }
  }

  // ---- epilogue: TMEM -> registers -> (+bias, +residual, tanh) -> HBM ----
  {
    const size_t row = (size_t)b * p.y_rows + (size_t)(row_ok ? t : 0) * p.y_tstride + p.y_toffset;
    const int co_base = nt * p.n_tile + col0;
    float* yp = p.y + row * p.C_out + co_base;
    const float* rp = p.res ? p.res + row * p.C_out + co_base : nullptr;
    const float* bias = FUSE2 ? p.bias2 : p.bias;
    const bool tanh_out = (p.flags & BC_CONV_TANH_OUT) != 0;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      const int n8 = min(4, (ncols - c0) / 8);
      // residual rows are fetched before the accumulator is touched: their latency overlaps the MMA tail
      float4 res4[8];
      if (rp && row_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < 2 * n8) res4[j] = __ldcs(reinterpret_cast<const float4*>(rp + c0) + j);
      }
      if (FUSE2 && c0 == 0) {
        mbar_wait(mbar, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      uint32_t r[32];
      tmem_load(tmem_base + acc_col + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + c0), n8, r);
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < 2 * n8) {
            float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                   __uint_as_float(r[4 * j + 3]));
            if (bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + co_base + c0) + j);
              v = add4(v, bb);
            }
            if (rp) { v.x += res4[j].x; v.y += res4[j].y; v.z += res4[j].z; v.w += res4[j].w; }
            if (tanh_out) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
            reinterpret_cast<float4*>(yp + c0)[j] = v;
          }
        }
      }
    }
  }
  if (FUSE2) phase ^= 1u;   // the tail's commit completed one more mbarrier phase
  // accumulator reads of this tile must be ordered before the next tile's MMAs (next __syncthreads)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }  // tile loop
  // ---- teardown ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int pick_n_tile(int C_out) {
  for (int n = 128; n >= 16; n -= 16)
    if (C_out % n == 0) return n;
  return 0;
}

}  // namespace

namespace bc {

// Geometry shared with the host-side weight packer (bc_tc_plan).
int tc_plan(int C_in, int C_out, int K, int stride, int dilation, int precision, int* n_tile, int* gpc, int* nchunks) {
  if (C_in % 16 != 0 || C_out % 16 != 0) return BC_EUNSUPPORTED;
  if (stride > 1 && dilation > 1) return BC_EUNSUPPORTED;
  const int nt = pick_n_tile(C_out);
  if (nt == 0) return BC_EUNSUPPORTED;
  const int split = precision == BC_PREC_BF16X3 ? 2 : 1;
  // 16-channel groups per staged chunk: largest power of two (<= 4) dividing C_in/16 whose weight image fits ~56 KB
  int g = 4;
  const int groups = C_in / 16;
  while (g > 1 && (groups % g != 0 || (size_t)split * K * g * nt * 32 > 56 * 1024)) g >>= 1;
  if ((size_t)split * K * g * nt * 32 > 100 * 1024) return BC_EUNSUPPORTED;
  *n_tile = nt;
  *gpc = g;
  *nchunks = groups / g;
  return BC_OK;
}

static int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

static int launch_tc(TcParams& p, int precision, cudaStream_t st) {
  p.split = precision == BC_PREC_BF16X3 ? 2 : 1;
  p.slab_rows = (BM - 1) * p.stride + (p.K - 1) * p.dil + 1;
  p.rpp = (p.slab_rows + p.stride - 1) / p.stride;
  p.tmem_cols = p.fuse2 ? 2 * pow2_cols(p.n_tile) : pow2_cols(p.n_tile);
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  p.variant = bc::policy().tc_variant;
  const size_t a_bytes = (size_t)p.split * 2 * p.gpc * p.stride * p.rpp * 16;
  const size_t b_bytes = (size_t)p.split * p.K * p.gpc * p.n_tile * 32;
  size_t region1 = ((a_bytes + 127) & ~size_t(127)) + ((b_bytes + 127) & ~size_t(127));
  size_t b2 = 0;
  const size_t a2 = p.fuse2 ? (size_t)p.split * p.n_tile * BM * 2 : 0;   // bf16 intermediate tile(s)
  if (p.fuse2) {
    if (a2 > region1) region1 = (a2 + 127) & ~size_t(127);
    b2 = ((size_t)p.split * p.n_tile * p.n_tile * 2 + 127) & ~size_t(127);
  }
  p.tiles_per_item = (p.T_out + BM - 1) / BM;
  const long long total_tiles = (long long)p.tiles_per_item * p.B;
  if (total_tiles > 2147483647ll) return fail(BC_EINVAL, "conv1d(tensor-core): too many tiles");
  p.total_tiles = (int)total_tiles;
  // persistent mode: one chunk, one n-tile, and (fused) the intermediate must fit inside the slab region alone
  p.persist = (p.nchunks == 1 && p.C_out == p.n_tile && (!p.fuse2 || a2 <= ((a_bytes + 127) & ~size_t(127))) &&
               bc::policy().tc_persist) ? 1 : 0;   // opt-in: measured slower than the one-tile-per-CTA schedule
  p.region1_bytes = (uint32_t)region1;
  const size_t smem = region1 + b2 + 64;
  if (smem > 227 * 1024) return fail(BC_EUNSUPPORTED, "conv1d(tensor-core): tile needs %zu B of shared memory", smem);
  if (p.tmem_cols > 512) return fail(BC_EUNSUPPORTED, "conv1d(tensor-core): %d TMEM columns", p.tmem_cols);
  if (2 * (size_t)p.gpc * p.stride * p.rpp * 16 >= (1u << 18) || (size_t)p.n_tile * 16 >= (1u << 18))
    return fail(BC_EUNSUPPORTED, "conv1d(tensor-core): descriptor offset overflow");
  void (*kern)(const TcParams) = nullptr;
  int slot = 0;
  if (p.split == 1 && !p.fuse2) { kern = conv1d_tc_kernel<1, false>; slot = 0; }
  if (p.split == 2 && !p.fuse2) { kern = conv1d_tc_kernel<2, false>; slot = 1; }
  if (p.split == 1 && p.fuse2) { kern = conv1d_tc_kernel<1, true>; slot = 2; }
  if (p.split == 2 && p.fuse2) { kern = conv1d_tc_kernel<2, true>; slot = 3; }
  static bool configured[64][4] = {{false}};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev][slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_check(e, "cudaFuncSetAttribute(conv1d_tc)");
    if (dev >= 0 && dev < 64) configured[dev][slot] = true;
  }
  dim3 grid(p.tiles_per_item, p.C_out / p.n_tile, p.B);
  if (p.persist) {
    int occ = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TC_THREADS, smem);
    if (e != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    // TMEM: 512 columns per SM shared by the co-resident CTAs
    if (occ * p.tmem_cols > 512) occ = 512 / p.tmem_cols;
    const long long want = (long long)occ * sms;
    grid = dim3((unsigned)(want < total_tiles ? want : total_tiles), 1, 1);
  }
  kern<<<grid, TC_THREADS, smem, st>>>(p);
  BC_LAUNCH_CHECK("conv1d_tc_kernel");
  return BC_OK;
}

int conv1d_tc_fwd(const float* x, const float* w, const float* bias, const float* snake_a, const float* snake_ib,
                  const float* res, float* y, int B, int T_in, int C_in, int T_out, int C_out, int K, int stride,
                  int dilation, int pad_left, int y_rows, int y_tstride, int y_toffset, int flags, int precision,
                  cudaStream_t st) {
  TcParams p;
  int rc = tc_plan(C_in, C_out, K, stride, dilation, precision, &p.n_tile, &p.gpc, &p.nchunks);
  if (rc != BC_OK)
    return fail(rc, "conv1d(tensor-core): unsupported geometry C_in=%d C_out=%d K=%d stride=%d dil=%d", C_in, C_out, K, stride, dilation);
  if (!aligned16(x) || !aligned16(w) || !aligned16(y) || (res && !aligned16(res)) || (bias && !aligned16(bias)) ||
      ((flags & BC_CONV_SNAKE_IN) && (!aligned16(snake_a) || !aligned16(snake_ib))))
    return fail(BC_EINVAL, "conv1d(tensor-core): pointers must be 16-byte aligned");
  p.x = x; p.wpk = reinterpret_cast<const uint4*>(w); p.bias = bias; p.sa = snake_a; p.sib = snake_ib; p.res = res; p.y = y;
  p.B = B; p.T_in = T_in; p.C_in = C_in; p.T_out = T_out; p.C_out = C_out; p.K = K; p.stride = stride; p.dil = dilation;
  p.pad_left = pad_left; p.y_rows = y_rows; p.y_tstride = y_tstride; p.y_toffset = y_toffset; p.flags = flags;
  p.fuse2 = 0; p.w2pk = nullptr; p.bias2 = nullptr; p.sa2 = nullptr; p.sib2 = nullptr;
  return launch_tc(p, precision, st);
}

int ru_persist_slots(int C, int K, int dilation, int precision);
int ru_group_groups(int C, int K, int dilation, int precision);
int resunit_group_fwd(const float* x, const float* w7, const float* b7, const float* sa1, const float* sib1,
                      const float* w1, const float* b1, const float* sa2, const float* sib2, float* y, int B, int T, int C,
                      int K, int dilation, int pad_left, int precision, cudaStream_t st);
int resunit_persist_fwd(const float* x, const float* w7, const float* b7, const float* sa1, const float* sib1,
                        const float* w1, const float* b1, const float* sa2, const float* sib2, float* y, int B, int T,
                        int C, int K, int dilation, int pad_left, int precision, cudaStream_t st);

int ru_pair_layout(int C, int K, int dilation, int precision);
int resunit_pair_fwd(const float* x, const void* w7_pair, const float* b7, const float* sa1, const float* sib1,
                     const void* w1_pair, const float* b1, const float* sa2, const float* sib2, float* y, int B, int T,
                     int C, int K, int dilation, int pad_left, int precision, cudaStream_t st);

int resunit_tc_fwd(const float* x, const float* w7, const float* b7, const float* sa1, const float* sib1,
                   const float* w1, const float* b1, const float* sa2, const float* sib2, float* y, int B, int T, int C,
                   int K, int dilation, int pad_left, int precision, cudaStream_t st) {
  if (ru_pair_layout(C, K, dilation, precision) > 0)      // CTA-pair kernel: w7 / w1 are the per-rank pair images
    return resunit_pair_fwd(x, w7, b7, sa1, sib1, w1, b1, sa2, sib2, y, B, T, C, K, dilation, pad_left, precision, st);
  if (ru_persist_slots(C, K, dilation, precision) > 0 && ru_group_groups(C, K, dilation, precision) > 0)
    return resunit_group_fwd(x, w7, b7, sa1, sib1, w1, b1, sa2, sib2, y, B, T, C, K, dilation, pad_left, precision, st);
  if (ru_persist_slots(C, K, dilation, precision) > 0)
    return resunit_persist_fwd(x, w7, b7, sa1, sib1, w1, b1, sa2, sib2, y, B, T, C, K, dilation, pad_left, precision, st);
  TcParams p;
  int rc = tc_plan(C, C, K, 1, dilation, precision, &p.n_tile, &p.gpc, &p.nchunks);
  if (rc != BC_OK || p.n_tile != C)
    return fail(BC_EUNSUPPORTED, "resunit(tensor-core): C=%d K=%d has no single-tile tensor-core plan", C, K);
  if (!aligned16(x) || !aligned16(w7) || !aligned16(w1) || !aligned16(y) || !aligned16(b7) || !aligned16(b1) ||
      !aligned16(sa1) || !aligned16(sib1) || !aligned16(sa2) || !aligned16(sib2))
    return fail(BC_EINVAL, "resunit(tensor-core): pointers must be 16-byte aligned");
  p.x = x; p.wpk = reinterpret_cast<const uint4*>(w7); p.bias = b7; p.sa = sa1; p.sib = sib1; p.res = x; p.y = y;
  p.B = B; p.T_in = T; p.C_in = C; p.T_out = T; p.C_out = C; p.K = K; p.stride = 1; p.dil = dilation;
  p.pad_left = pad_left; p.y_rows = T; p.y_tstride = 1; p.y_toffset = 0; p.flags = BC_CONV_SNAKE_IN;
  p.fuse2 = 1; p.w2pk = reinterpret_cast<const uint4*>(w1); p.bias2 = b1; p.sa2 = sa2; p.sib2 = sib2;
  return launch_tc(p, precision, st);
}

}  // namespace bc

extern "C" int bc_tc_plan(int C_in, int C_out, int K, int stride, int dilation, int precision, int* n_tile, int* gpc,
                          int* nchunks) {
  if (!n_tile || !gpc || !nchunks) return bc::fail(BC_EINVAL, "tc_plan: null output");
  int rc = bc::tc_plan(C_in, C_out, K, stride, dilation, precision, n_tile, gpc, nchunks);
  if (rc != BC_OK) bc::set_error("tc_plan: geometry C_in=%d C_out=%d K=%d stride=%d dil=%d has no tensor-core tiling", C_in, C_out, K, stride, dilation);
  return rc;
}

extern "C" int bc_resunit_fwd(const float* x, const float* w7, const float* b7, const float* sa1, const float* sib1,
                              const float* w1, const float* b1, const float* sa2, const float* sib2, float* y, int B,
                              int T, int C, int K, int dilation, int pad_left, int precision, bc_stream_t s) {
  BC_REQUIRE(x && w7 && b7 && sa1 && sib1 && w1 && b1 && sa2 && sib2 && y, "resunit: null pointer");
  BC_REQUIRE(B > 0 && T > 0 && C > 0 && K > 0 && dilation > 0 && B <= 65535, "resunit: bad shape B=%d T=%d C=%d K=%d", B, T, C, K);
  BC_REQUIRE(x != y, "resunit: cannot run in place (neighbouring tiles read the input halo)");
  if (precision == BC_PREC_FP32)
    return bc::fail(BC_EUNSUPPORTED, "resunit: the fused kernel exists for the tensor-core modes only; chain two bc_conv1d_fwd calls in fp32 mode");
  return bc::resunit_tc_fwd(x, w7, b7, sa1, sib1, w1, b1, sa2, sib2, y, B, T, C, K, dilation, pad_left, precision, (cudaStream_t)s);
}

extern "C" int bc_resunit_plan(int C, int K, int dilation, int precision, int* n_tile, int* gpc, int* nchunks,
                               int* persistent) {
  if (!n_tile || !gpc || !nchunks || !persistent) return bc::fail(BC_EINVAL, "resunit_plan: null output");
  if (precision == BC_PREC_FP32) return bc::fail(BC_EUNSUPPORTED, "resunit_plan: tensor-core modes only");
  if (bc::ru_pair_layout(C, K, dilation, precision) > 0) {
    *n_tile = C; *gpc = C / 16; *nchunks = 1;
    *persistent = 2 + bc::ru_pair_layout(C, K, dilation, precision);           // 3: CTA-pair kernel, plain image; 4: stacked image
    return BC_OK;
  }
  if (bc::ru_persist_slots(C, K, dilation, precision) > 0) {
    *n_tile = C; *gpc = C / 16; *nchunks = 1;
    *persistent = bc::ru_group_groups(C, K, dilation, precision) > 0 ? 2 : 1;   // 2: warpgroup-per-tile kernel, 1: role pipeline
    return BC_OK;
  }
  *persistent = 0;
  int rc = bc::tc_plan(C, C, K, 1, dilation, precision, n_tile, gpc, nchunks);
  if (rc != BC_OK || *n_tile != C) return bc::fail(BC_EUNSUPPORTED, "resunit_plan: C=%d K=%d has no fused tensor-core plan", C, K);
  // the per-tile fused kernel only pays off while two CTAs still fit on an SM (measured: C=128 in the
  // split-precision mode needs 133 KB and is faster as two separate launches)
  {
    const size_t split = precision == BC_PREC_BF16X3 ? 2 : 1;
    const size_t slab_rows = 127 + (size_t)(K - 1) * dilation + 1;
    const size_t a_bytes = split * 2 * (*gpc) * slab_rows * 16, b_bytes = split * K * (*gpc) * C * 32;
    size_t region1 = ((a_bytes + 127) & ~size_t(127)) + ((b_bytes + 127) & ~size_t(127));
    const size_t a2 = split * C * 128 * 2;
    if (a2 > region1) region1 = a2;
    const size_t smem = region1 + split * C * C * 2 + 64;
    if (smem > 110 * 1024) return bc::fail(BC_EUNSUPPORTED, "resunit_plan: fused tile needs %zu B of shared memory; chain two convs", smem);
  }
  return BC_OK;
}
