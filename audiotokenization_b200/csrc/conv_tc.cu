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
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int TC_THREADS = 256;

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
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int variant) {
  // cute::UMMA::SmemDescriptor: start [0,14) | LBO [16,30) | SBO [32,46) | version=1 [46,48) | layout_type=0 (no swizzle)
  if (variant & 1) { uint32_t t = lbo_bytes; lbo_bytes = sbo_bytes; sbo_bytes = t; }
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  if (!(variant & 2)) d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded: a wrong descriptor must surface as a launch failure, never as a hung GPU
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (it > (1u << 22)) __trap();
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void split_store(float v[8], uint4* hi_dst, uint4* lo_dst) {
  uint4 h;
  h.x = pack_bf16x2(v[0], v[1]); h.y = pack_bf16x2(v[2], v[3]);
  h.z = pack_bf16x2(v[4], v[5]); h.w = pack_bf16x2(v[6], v[7]);
  *hi_dst = h;
  if (lo_dst) {
    float r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = v[e] - __bfloat162float(__float2bfloat16_rn(v[e]));
    uint4 l;
    l.x = pack_bf16x2(r[0], r[1]); l.y = pack_bf16x2(r[2], r[3]);
    l.z = pack_bf16x2(r[4], r[5]); l.w = pack_bf16x2(r[6], r[7]);
    *lo_dst = l;
  }
}

__global__ void __launch_bounds__(TC_THREADS) conv1d_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int planes = 2 * p.gpc;
  const uint32_t plane_bytes = (uint32_t)p.stride * p.rpp * 16u;  // one 8-channel plane of the slab
  const uint32_t a_split_bytes = planes * plane_bytes;
  const uint32_t a_bytes = a_split_bytes * p.split;
  const uint32_t b_split_bytes = (uint32_t)p.K * p.gpc * p.n_tile * 32u;
  const uint32_t b_bytes = b_split_bytes * p.split;
  uint8_t* sA = smem_raw;
  uint8_t* sB = smem_raw + ((a_bytes + 127u) & ~127u);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sB + ((b_bytes + 127u) & ~127u));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int b = blockIdx.z;
  const int nt = blockIdx.y;
  const int t0 = blockIdx.x * BM;
  const int g0 = t0 * p.stride - p.pad_left;
  const float* xb = p.x + (size_t)b * p.T_in * p.C_in;
  const bool snake = (p.flags & BC_CONV_SNAKE_IN) != 0;

  // ---- one-time setup: mbarrier + TMEM allocation ----
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  uint32_t phase = 0;
  const int items = planes * p.slab_rows;  // 16-byte slab items per split
  for (int ch = 0; ch < p.nchunks; ++ch) {
    const int ci0 = ch * p.gpc * 16;
    // ---- stage A: x (fp32, HBM) -> snake -> bf16 hi[/lo] -> canonical K-major slab ----
    for (int i = tid; i < items; i += TC_THREADS) {
      const int pl = i % planes;
      const int r = i / planes;
      const int g = g0 + r;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (g >= 0 && g < p.T_in) {
        const float* src = xb + (size_t)g * p.C_in + ci0 + pl * 8;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w;
        v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
        if (snake) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.sa + ci0 + pl * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.sa + ci0 + pl * 8) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.sib + ci0 + pl * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.sib + ci0 + pl * 8) + 1);
          v[0] = bc::snake_ref(v[0], a0.x, b0.x); v[1] = bc::snake_ref(v[1], a0.y, b0.y);
          v[2] = bc::snake_ref(v[2], a0.z, b0.z); v[3] = bc::snake_ref(v[3], a0.w, b0.w);
          v[4] = bc::snake_ref(v[4], a1.x, b1.x); v[5] = bc::snake_ref(v[5], a1.y, b1.y);
          v[6] = bc::snake_ref(v[6], a1.z, b1.z); v[7] = bc::snake_ref(v[7], a1.w, b1.w);
        }
      }
      const int ph = r % p.stride, rr = r / p.stride;
      uint8_t* dst = sA + (size_t)pl * plane_bytes + ((size_t)ph * p.rpp + rr) * 16;
      split_store(v, reinterpret_cast<uint4*>(dst), p.split == 2 ? reinterpret_cast<uint4*>(dst + a_split_bytes) : nullptr);
    }
    // ---- stage B: pre-packed weight image for (n-tile, chunk): straight 16-byte copy ----
    {
      const uint4* src = p.wpk + ((size_t)nt * p.nchunks + ch) * (b_bytes / 16);
      uint4* dst = reinterpret_cast<uint4*>(sB);
      for (int i = tid; i < (int)(b_bytes / 16); i += TC_THREADS) dst[i] = __ldg(src + i);
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- one thread issues every MMA of this chunk, then commits to the mbarrier ----
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      const int nterms = p.split == 2 ? 3 : 1;
      for (int k = 0; k < p.K; ++k) {
        const int sh = k * p.dil;
        const uint32_t a_row = ((uint32_t)(sh % p.stride) * p.rpp + (uint32_t)(sh / p.stride)) * 16u;
        for (int g = 0; g < p.gpc; ++g) {
          const uint32_t a_off = (uint32_t)(2 * g) * plane_bytes + a_row;
          const uint32_t b_off = (uint32_t)(k * p.gpc + g) * p.n_tile * 32u;
          for (int term = 0; term < nterms; ++term) {
            // term 0: a_hi*w_hi, 1: a_hi*w_lo, 2: a_lo*w_hi
            const uint32_t aa = a_base + a_off + (term == 2 ? a_split_bytes : 0u);
            const uint32_t bb = b_base + b_off + (term == 1 ? b_split_bytes : 0u);
            const uint64_t ad = make_desc(aa, plane_bytes, 128u, p.variant);
            const uint64_t bd = make_desc(bb, (uint32_t)p.n_tile * 16u, 128u, p.variant);
            const uint32_t acc = (ch | k | g | term) ? 1u : 0u;
            mma_bf16(tmem_base, ad, bd, p.idesc, acc);
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    // everyone waits until the tensor core has consumed this chunk's smem
    mbar_wait(smem_u32(mbar), phase);
    phase ^= 1u;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: TMEM -> registers -> (+bias, +residual, tanh) -> HBM ----
  {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = warp >> 2;        // column half
    const int ncols = p.n_tile / 2;
    const int col0 = half * ncols;
    const int t = t0 + q * 32 + lane;
    const bool row_ok = t < p.T_out;
    const size_t row = (size_t)b * p.y_rows + (size_t)(row_ok ? t : 0) * p.y_tstride + p.y_toffset;
    const int co_base = nt * p.n_tile + col0;
    float* yp = p.y + row * p.C_out + co_base;
    const float* rp = p.res ? p.res + row * p.C_out + co_base : nullptr;
    const bool tanh_out = (p.flags & BC_CONV_TANH_OUT) != 0;
    for (int c = 0; c < ncols; c += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + c);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row_ok) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[e]);
        if (p.bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + co_base + c));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + co_base + c) + 1);
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (rp) {
          const float4 r0 = *reinterpret_cast<const float4*>(rp + c);
          const float4 r1 = *(reinterpret_cast<const float4*>(rp + c) + 1);
          v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
          v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if (tanh_out) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = tanhf(v[e]);
        }
        *reinterpret_cast<float4*>(yp + c) = make_float4(v[0], v[1], v[2], v[3]);
        *(reinterpret_cast<float4*>(yp + c) + 1) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
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
  p.split = precision == BC_PREC_BF16X3 ? 2 : 1;
  p.slab_rows = (BM - 1) * stride + (K - 1) * dilation + 1;
  p.rpp = (p.slab_rows + stride - 1) / stride;
  p.tmem_cols = p.n_tile <= 32 ? 32 : (p.n_tile <= 64 ? 64 : 128);
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const char* var = getenv("BC_TC_VARIANT");
  p.variant = var ? atoi(var) : 0;
  const size_t a_bytes = (size_t)p.split * 2 * p.gpc * stride * p.rpp * 16;
  const size_t b_bytes = (size_t)p.split * K * p.gpc * p.n_tile * 32;
  const size_t smem = ((a_bytes + 127) & ~size_t(127)) + ((b_bytes + 127) & ~size_t(127)) + 64;
  if (smem > 227 * 1024) return fail(BC_EUNSUPPORTED, "conv1d(tensor-core): tile needs %zu B of shared memory", smem);
  if (2 * (size_t)p.gpc * stride * p.rpp * 16 >= (1u << 18) || (size_t)p.n_tile * 16 >= (1u << 18))
    return fail(BC_EUNSUPPORTED, "conv1d(tensor-core): descriptor offset overflow");
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv1d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_check(e, "cudaFuncSetAttribute(conv1d_tc)");
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  dim3 grid((T_out + BM - 1) / BM, C_out / p.n_tile, B);
  conv1d_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p);
  BC_LAUNCH_CHECK("conv1d_tc_kernel");
  return BC_OK;
}

}  // namespace bc

extern "C" int bc_tc_plan(int C_in, int C_out, int K, int stride, int dilation, int precision, int* n_tile, int* gpc,
                          int* nchunks) {
  if (!n_tile || !gpc || !nchunks) return bc::fail(BC_EINVAL, "tc_plan: null output");
  int rc = bc::tc_plan(C_in, C_out, K, stride, dilation, precision, n_tile, gpc, nchunks);
  if (rc != BC_OK) bc::set_error("tc_plan: geometry C_in=%d C_out=%d K=%d stride=%d dil=%d has no tensor-core tiling", C_in, C_out, K, stride, dilation);
  return rc;
}
