// Factorized VQ: in_proj -> L2 normalise -> cosine argmax over the codebook -> int32,
// and the inverse (codebook gather + out_proj).  CUDA cores: the contraction depth is
// the codebook dimension D = 8.
//
// Encode, per CTA of 8 warps = 32 frames:
//   phase 1  warp w projects frames 4w..4w+3: lanes stride the C input channels
//            (coalesced), D partial sums per frame, butterfly-reduced; then the
//            projected vector is normalised exactly like F.normalize (x / max(|x|, 1e-12)).
//   phase 2  the same warp scans the whole (pre-normalised) codebook for its 4 frames:
//            lane l visits codes l, l+32, ... (coalesced 32-byte rows, L1/L2 resident),
//            keeps top-1 / top-2 in registers; warp-shuffle merge, ties -> lowest index
//            (torch.max semantics, factorized_vector_quantize.py:106).
#include "common.cuh"
#include <float.h>

namespace {

constexpr int VQ_WARPS = 8;
constexpr int FPW = 4;  // frames per warp
constexpr int MAXD = 16;

struct Top2 {
  float v1, v2;
  int i1;
};

__device__ __forceinline__ void top2_push(Top2& t, float v, int i) {
  // strict > keeps the lowest index among equal values when codes are visited in increasing order
  if (v > t.v1) {
    t.v2 = t.v1;
    t.v1 = v;
    t.i1 = i;
  } else if (v > t.v2) {
    t.v2 = v;
  }
}

__device__ __forceinline__ void top2_merge(Top2& a, float v1, float v2, int i1) {
  // merge another lane's (v1 >= v2, i1)
  if (v1 > a.v1 || (v1 == a.v1 && i1 < a.i1)) {
    a.v2 = fmaxf(a.v1, v2);
    a.v1 = v1;
    a.i1 = i1;
  } else {
    a.v2 = fmaxf(a.v2, v1);
  }
}

template <int D>
__global__ void __launch_bounds__(VQ_WARPS * 32) vq_encode_kernel(const float* __restrict__ z, const float* __restrict__ w_in,
                                                                  const float* __restrict__ b_in,
                                                                  const float* __restrict__ cbn, int32_t* __restrict__ idx,
                                                                  float* __restrict__ margin, float* __restrict__ z_e_out,
                                                                  int N, int C, int Kc) {
  extern __shared__ __align__(16) float smem[];  // w_in [D][C] (if projecting)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool proj = w_in != nullptr;
  if (proj) {
    for (int i = tid; i < D * C; i += VQ_WARPS * 32) smem[i] = __ldg(w_in + i);
  }
  __syncthreads();

  const int n0 = (blockIdx.x * VQ_WARPS + warp) * FPW;
  if (n0 >= N) return;

  // ---- phase 1: projection ----
  float e[FPW][D];
#pragma unroll
  for (int f = 0; f < FPW; ++f)
#pragma unroll
    for (int d = 0; d < D; ++d) e[f][d] = 0.f;
  if (proj) {
    for (int c = lane; c < C; c += 32) {
      float zv[FPW];
#pragma unroll
      for (int f = 0; f < FPW; ++f) zv[f] = (n0 + f < N) ? __ldcs(z + (size_t)(n0 + f) * C + c) : 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float w = smem[d * C + c];
#pragma unroll
        for (int f = 0; f < FPW; ++f) e[f][d] = fmaf(zv[f], w, e[f][d]);
      }
    }
#pragma unroll
    for (int f = 0; f < FPW; ++f)
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float v = e[f][d];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        e[f][d] = v + __ldg(b_in + d);
      }
  } else {
#pragma unroll
    for (int f = 0; f < FPW; ++f)
#pragma unroll
      for (int d = 0; d < D; ++d) e[f][d] = (n0 + f < N) ? __ldg(z + (size_t)(n0 + f) * C + d) : 0.f;
  }
  if (z_e_out && lane == 0) {
#pragma unroll
    for (int f = 0; f < FPW; ++f)
      if (n0 + f < N)
#pragma unroll
        for (int d = 0; d < D; ++d) z_e_out[(size_t)(n0 + f) * D + d] = e[f][d];
  }
  // F.normalize: x / max(||x||_2, eps)
#pragma unroll
  for (int f = 0; f < FPW; ++f) {
    float ss = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) ss = fmaf(e[f][d], e[f][d], ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int d = 0; d < D; ++d) e[f][d] *= inv;
  }

  // ---- phase 2: codebook scan ----
  Top2 best[FPW];
#pragma unroll
  for (int f = 0; f < FPW; ++f) { best[f].v1 = -FLT_MAX; best[f].v2 = -FLT_MAX; best[f].i1 = 0x7fffffff; }
  for (int k = lane; k < Kc; k += 32) {
    float cv[D];
    const float* cp = cbn + (size_t)k * D;
    if (D % 4 == 0) {
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(cp + d));
        cv[d] = t4.x; cv[d + 1] = t4.y; cv[d + 2] = t4.z; cv[d + 3] = t4.w;
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) cv[d] = __ldg(cp + d);
    }
#pragma unroll
    for (int f = 0; f < FPW; ++f) {
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) dot = fmaf(e[f][d], cv[d], dot);
      top2_push(best[f], dot, k);
    }
  }
#pragma unroll
  for (int f = 0; f < FPW; ++f) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov1 = __shfl_xor_sync(0xffffffffu, best[f].v1, o);
      const float ov2 = __shfl_xor_sync(0xffffffffu, best[f].v2, o);
      const int oi1 = __shfl_xor_sync(0xffffffffu, best[f].i1, o);
      top2_merge(best[f], ov1, ov2, oi1);
    }
    if (lane == 0 && n0 + f < N) {
      idx[n0 + f] = best[f].i1;
      if (margin) margin[n0 + f] = best[f].v1 - best[f].v2;
    }
  }
}

// q[n][c] = b_out[c] + sum_d w_out[c][d] * cb[idx[n]][d]
template <int D>
__global__ void __launch_bounds__(256) vq_dequant_kernel(const int32_t* __restrict__ idx, const float* __restrict__ cb,
                                                         const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                         float* __restrict__ z_q, float* __restrict__ residual,
                                                         int* __restrict__ bad_count, int N, int C, int Kc, int accumulate) {
  // CTA = 8 frames x all channels; thread strides channels
  __shared__ float code[8][D];
  const int n0 = blockIdx.x * 8;
  const int tid = threadIdx.x;
  if (tid < 8 * D) {
    const int f = tid / D, d = tid % D;
    float v = 0.f;
    if (n0 + f < N) {
      int k = idx[n0 + f];
      if (k < 0 || k >= Kc) {
        if (bad_count && d == 0) atomicAdd(bad_count, 1);
        k = k < 0 ? 0 : Kc - 1;
      }
      v = __ldg(cb + (size_t)k * D + d);
    }
    code[f][d] = v;
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float w[D];
    if (w_out) {
#pragma unroll
      for (int d = 0; d < D; ++d) w[d] = __ldg(w_out + (size_t)c * D + d);
    }
    const float bias = (w_out && b_out) ? __ldg(b_out + c) : 0.f;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      if (n0 + f >= N) break;
      float q;
      if (w_out) {
        q = bias;
#pragma unroll
        for (int d = 0; d < D; ++d) q = fmaf(w[d], code[f][d], q);
      } else {
        q = code[f][c];  // identity projection: C == D
      }
      const size_t o = (size_t)(n0 + f) * C + c;
      z_q[o] = accumulate ? z_q[o] + q : q;
      if (residual) residual[o] -= q;
    }
  }
}

}  // namespace

#define VQ_DISPATCH_D(D_, ...)                                          \
  switch (D_) {                                                         \
    case 4: { constexpr int DD = 4; __VA_ARGS__; } break;               \
    case 8: { constexpr int DD = 8; __VA_ARGS__; } break;               \
    case 16: { constexpr int DD = 16; __VA_ARGS__; } break;             \
    default: return bc::fail(BC_EUNSUPPORTED, "vq: codebook_dim=%d not in {4,8,16}", D_); \
  }

extern "C" int bc_vq_encode(const float* z, const float* w_in, const float* b_in, const float* cb_norm, int32_t* idx,
                            float* margin, float* z_e, int N, int C, int D, int Kc, bc_stream_t s) {
  BC_REQUIRE(z && cb_norm && idx, "vq_encode: null pointer");
  BC_REQUIRE(N > 0 && C > 0 && D > 0 && Kc >= 2, "vq_encode: bad shape N=%d C=%d D=%d K=%d", N, C, D, Kc);
  BC_REQUIRE((w_in == nullptr) == (b_in == nullptr), "vq_encode: w_in and b_in must both be given or both NULL");
  if (!w_in) BC_REQUIRE(C == D, "vq_encode: identity projection needs C == D (C=%d D=%d)", C, D);
  BC_REQUIRE(bc::aligned16(cb_norm), "vq_encode: codebook must be 16-byte aligned");
  const size_t smem = w_in ? (size_t)D * C * sizeof(float) : 0;
  if (smem > 200 * 1024) return bc::fail(BC_EUNSUPPORTED, "vq_encode: in_proj %dx%d does not fit shared memory", D, C);
  const int frames_per_cta = VQ_WARPS * FPW;
  const unsigned grid = (unsigned)((N + frames_per_cta - 1) / frames_per_cta);
  cudaStream_t st = (cudaStream_t)s;
  VQ_DISPATCH_D(D, {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(vq_encode_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(vq_encode)");
    }
    vq_encode_kernel<DD><<<grid, VQ_WARPS * 32, smem, st>>>(z, w_in, b_in, cb_norm, idx, margin, z_e, N, C, Kc);
  });
  BC_LAUNCH_CHECK("vq_encode_kernel");
  return BC_OK;
}

extern "C" int bc_vq_dequant(const int32_t* idx, const float* cb, const float* w_out, const float* b_out, float* z_q,
                             float* residual, int* bad_count, int N, int C, int D, int Kc, int accumulate,
                             bc_stream_t s) {
  BC_REQUIRE(idx && cb && z_q, "vq_dequant: null pointer");
  BC_REQUIRE(N > 0 && C > 0 && D > 0 && Kc > 0, "vq_dequant: bad shape N=%d C=%d D=%d K=%d", N, C, D, Kc);
  if (!w_out) BC_REQUIRE(C == D, "vq_dequant: identity projection needs C == D (C=%d D=%d)", C, D);
  const unsigned grid = (unsigned)((N + 7) / 8);
  cudaStream_t st = (cudaStream_t)s;
  VQ_DISPATCH_D(D, {
    vq_dequant_kernel<DD><<<grid, 256, 0, st>>>(idx, cb, w_out, b_out, z_q, residual, bad_count, N, C, Kc, accumulate);
  });
  BC_LAUNCH_CHECK("vq_dequant_kernel");
  return BC_OK;
}
