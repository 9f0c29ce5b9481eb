// Factorized VQ: in_proj -> L2 normalise -> cosine argmax over the codebook -> int32,
// and the inverse (codebook gather + out_proj).  CUDA cores: the contraction depth is
// the codebook dimension D = 8, so the search is bound by the fp32 FMA pipe
// (2*K*D FLOP against 2052 B per frame), not by HBM.
//
// Encode, per CTA of 256 threads = FR * 256 frames (D = 8: vq_scan_kernel; other D: the generic kernel below):
//   phase 1  warp w projects its share of the CTA's frames, 4 at a time: lanes stride the C input channels
//            (coalesced), D partial sums per frame, butterfly-reduced; the projected vector is normalised exactly
//            like F.normalize (x / max(|x|, 1e-12)) and parked in shared memory.
//   phase 2  every thread takes FR = 8 frames into registers (as 4 packed fp32 pairs) and scans the whole codebook,
//            which streams through shared memory in tiles of 512 codes with every component DUPLICATED (c, c): one
//            broadcast LDS.128 feeds two packed FMAs (fma.rn.f32x2: two frames per instruction, one issue slot),
//            so a code costs 4 LDS + 32 FFMA2 + the top-1 bookkeeping (3 instructions per frame) for 64 FMAs -- the
//            FMA pipe, not instruction issue, is the limit.  Codes are visited in increasing order by every thread
//            and a strict > keeps the lowest index among equal values (torch.max semantics,
//            factorized_vector_quantize.py:106): no cross-lane merge at all.  The top-2 bookkeeping (margin) is a
//            template variant and only runs when a margin is requested.
#include "common.cuh"
#include "tc_common.cuh"
#include <float.h>

namespace {
using bc::tc::f32x2;
using bc::tc::fma2;
using bc::tc::pack2;
using bc::tc::unpack2;

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

// ---------------------------------------------------------------------------------------------------------------
// D = 8: register-tiled scan over a shared-memory-resident codebook tile
// ---------------------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_FR_MAX = 8;                     // frames per thread: 8 (large inputs) or 2 (fewer than one wave of 2048-frame CTAs)
constexpr int SC_TK = 512;                       // codes per shared-memory tile
constexpr int SC_D = 8;
constexpr size_t SC_SMEM = (size_t)SC_THREADS * SC_FR_MAX * SC_D * sizeof(float);   // 64 KB: projected frames, then two duplicated code tiles

template <bool MARGIN, int SC_FR>
__global__ void __launch_bounds__(SC_THREADS, 2) vq_scan_kernel(const float* __restrict__ z, const float* __restrict__ w_in,
                                                                const float* __restrict__ b_in, const float* __restrict__ cbn,
                                                                int32_t* __restrict__ idx, float* __restrict__ margin,
                                                                float* __restrict__ z_e_out, int N, int C, int Kc) {
  extern __shared__ __align__(16) float smem[];
  constexpr int D = SC_D;
  constexpr int SC_FRAMES = SC_THREADS * SC_FR;    // frames per CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n_base = (long long)blockIdx.x * SC_FRAMES;
  const int n_here = (int)min((long long)SC_FRAMES, (long long)N - n_base);
  float* sE = smem;                              // [SC_FRAMES][D] normalised projections (phase 1 -> phase 2 hand-off)

  // ---- phase 1: projection + normalisation (w_in read through L1: 16 KB shared by all warps of the SM) ----
  const bool proj = w_in != nullptr;
  for (int f0 = warp * FPW; f0 < n_here; f0 += VQ_WARPS * FPW) {
    float e[FPW][D];
#pragma unroll
    for (int f = 0; f < FPW; ++f)
#pragma unroll
      for (int d = 0; d < D; ++d) e[f][d] = 0.f;
    if (proj) {
      for (int c = lane; c < C; c += 32) {
        float zv[FPW];
#pragma unroll
        for (int f = 0; f < FPW; ++f) zv[f] = (f0 + f < n_here) ? __ldcs(z + (size_t)(n_base + f0 + f) * C + c) : 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float w = __ldg(w_in + d * C + c);
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
        for (int d = 0; d < D; ++d) e[f][d] = (f0 + f < n_here) ? __ldg(z + (size_t)(n_base + f0 + f) * C + d) : 0.f;
    }
    if (lane < FPW && f0 + lane < n_here) {       // lane f finishes frame f0 + f (every lane holds all sums after the butterfly)
      float v[D];
#pragma unroll
      for (int f = 0; f < FPW; ++f)
        if (lane == f) {
#pragma unroll
          for (int d = 0; d < D; ++d) v[d] = e[f][d];
        }
      if (z_e_out) {
#pragma unroll
        for (int d = 0; d < D; ++d) z_e_out[(size_t)(n_base + f0 + lane) * D + d] = v[d];
      }
      float ss = 0.f;                             // F.normalize: x / max(||x||_2, eps)
#pragma unroll
      for (int d = 0; d < D; ++d) ss = fmaf(v[d], v[d], ss);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      float4* dst = reinterpret_cast<float4*>(sE + (size_t)(f0 + lane) * D);
      dst[0] = make_float4(v[0] * inv, v[1] * inv, v[2] * inv, v[3] * inv);
      dst[1] = make_float4(v[4] * inv, v[5] * inv, v[6] * inv, v[7] * inv);
    }
  }
  __syncthreads();

  // ---- phase 2: this thread's frames tid, tid + 256, ... as packed pairs (frame 2p, frame 2p+1) ----
  f32x2 e2[SC_FR / 2][D];
#pragma unroll
  for (int pp = 0; pp < SC_FR / 2; ++pp) {
    const int fa = tid + SC_THREADS * (2 * pp), fb = fa + SC_THREADS;
    float a[D], b[D];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 va = fa < n_here ? reinterpret_cast<const float4*>(sE + (size_t)fa * D)[h] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 vb = fb < n_here ? reinterpret_cast<const float4*>(sE + (size_t)fb * D)[h] : make_float4(0.f, 0.f, 0.f, 0.f);
      a[4 * h] = va.x; a[4 * h + 1] = va.y; a[4 * h + 2] = va.z; a[4 * h + 3] = va.w;
      b[4 * h] = vb.x; b[4 * h + 1] = vb.y; b[4 * h + 2] = vb.z; b[4 * h + 3] = vb.w;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) e2[pp][d] = pack2(a[d], b[d]);
  }
  __syncthreads();                                // sE is dead: the buffer becomes the two code tiles

  float best[SC_FR], second[SC_FR];
  int besti[SC_FR];
#pragma unroll
  for (int f = 0; f < SC_FR; ++f) { best[f] = -FLT_MAX; second[f] = -FLT_MAX; besti[f] = 0x7fffffff; }

  // code tiles: [SC_TK][D][2] floats (component duplicated), double-buffered; filled by all threads with coalesced loads
  float* sC0 = smem;
  float* sC1 = smem + (size_t)SC_TK * D * 2;
  auto fill = [&](float* dstT, int k0) {
    // 2 codes per thread: one float4 pair (32 B) per code from global, 64 B duplicated into shared memory
    for (int i = tid; i < SC_TK; i += SC_THREADS) {
      const int k = k0 + i;
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
      if (k < Kc) {
        lo = __ldg(reinterpret_cast<const float4*>(cbn + (size_t)k * D));
        hi = __ldg(reinterpret_cast<const float4*>(cbn + (size_t)k * D) + 1);
      }
      float4* d4 = reinterpret_cast<float4*>(dstT + (size_t)i * D * 2);
      d4[0] = make_float4(lo.x, lo.x, lo.y, lo.y);
      d4[1] = make_float4(lo.z, lo.z, lo.w, lo.w);
      d4[2] = make_float4(hi.x, hi.x, hi.y, hi.y);
      d4[3] = make_float4(hi.z, hi.z, hi.w, hi.w);
    }
  };
  fill(sC0, 0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < Kc; k0 += SC_TK, buf ^= 1) {
    const float* cur = buf ? sC1 : sC0;
    if (k0 + SC_TK < Kc) fill(buf ? sC0 : sC1, k0 + SC_TK);      // next tile: its global loads overlap this tile's scan
    const int kend = min(SC_TK, Kc - k0);
#pragma unroll 2
    for (int i = 0; i < kend; ++i) {
      const ulonglong2* cp = reinterpret_cast<const ulonglong2*>(cur + (size_t)i * D * 2);   // broadcast: every lane reads the same code
      const ulonglong2 c01 = cp[0], c23 = cp[1], c45 = cp[2], c67 = cp[3];
      const f32x2 cd[D] = {c01.x, c01.y, c23.x, c23.y, c45.x, c45.y, c67.x, c67.y};
      const int k = k0 + i;
#pragma unroll
      for (int pp = 0; pp < SC_FR / 2; ++pp) {
        f32x2 acc = fma2(e2[pp][0], cd[0], pack2(0.f, 0.f));
#pragma unroll
        for (int d = 1; d < D; ++d) acc = fma2(e2[pp][d], cd[d], acc);
        float da, db;
        unpack2(acc, da, db);
        if (MARGIN) {
          if (da > best[2 * pp]) { second[2 * pp] = best[2 * pp]; best[2 * pp] = da; besti[2 * pp] = k; }
          else if (da > second[2 * pp]) second[2 * pp] = da;
          if (db > best[2 * pp + 1]) { second[2 * pp + 1] = best[2 * pp + 1]; best[2 * pp + 1] = db; besti[2 * pp + 1] = k; }
          else if (db > second[2 * pp + 1]) second[2 * pp + 1] = db;
        } else {
          if (da > best[2 * pp]) { best[2 * pp] = da; besti[2 * pp] = k; }
          if (db > best[2 * pp + 1]) { best[2 * pp + 1] = db; besti[2 * pp + 1] = k; }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int f = 0; f < SC_FR; ++f) {
    const int fr = tid + SC_THREADS * f;
    if (fr < n_here) {
      idx[n_base + fr] = besti[f];
      if (MARGIN) margin[n_base + fr] = best[f] - second[f];
    }
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

// ---------------------------------------------------------------------------------------------------------------
// Finite scalar quantisation (the other quantizer BigCodecDecoder can select: vq/codec_decoder.py:41-47,87-89;
// FSQ.forward / bound / quantize / codes_to_indices, finite_scalar_quantization.py:111-148,170-175,203-259):
//   z_e = W_in z + b_in                      (nn.Linear dim -> d)
//   bounded_j = tanh(z_e_j + shift_j) * half_l_j - offset_j
//   q_j = round_half_even(bounded_j),  code_j = q_j / half_width_j
//   index = int32( sum_j (code_j * half_width_j + half_width_j) * basis_j )
// One warp per 4 frames: lanes stride the C input channels (coalesced), d partial sums per frame, butterfly-reduced;
// HBM-bound (2 KB read per frame).  `boundary` (optional) receives min_j | bounded_j - nearest rounding boundary |.
// ---------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(VQ_WARPS * 32) fsq_encode_kernel(const float* __restrict__ z, const float* __restrict__ w_in,
                                                                   const float* __restrict__ b_in, const float* __restrict__ prm,
                                                                   int32_t* __restrict__ idx, float* __restrict__ codes,
                                                                   float* __restrict__ boundary, int N, int C) {
  // prm: [5][D] = half_l | offset | shift | half_width | basis
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = (blockIdx.x * VQ_WARPS + warp) * FPW;
  if (n0 >= N) return;
  float e[FPW][D];
#pragma unroll
  for (int f = 0; f < FPW; ++f)
#pragma unroll
    for (int d = 0; d < D; ++d) e[f][d] = 0.f;
  if (w_in) {
    for (int c = lane; c < C; c += 32) {
      float zv[FPW];
#pragma unroll
      for (int f = 0; f < FPW; ++f) zv[f] = (n0 + f < N) ? __ldcs(z + (size_t)(n0 + f) * C + c) : 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float w = __ldg(w_in + d * C + c);
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
        e[f][d] = v + (b_in ? __ldg(b_in + d) : 0.f);
      }
  } else {
#pragma unroll
    for (int f = 0; f < FPW; ++f)
#pragma unroll
      for (int d = 0; d < D; ++d) e[f][d] = (n0 + f < N) ? __ldg(z + (size_t)(n0 + f) * C + d) : 0.f;
  }
  if (lane < FPW && n0 + lane < N) {
    float v[D];
#pragma unroll
    for (int f = 0; f < FPW; ++f)
      if (lane == f) {
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = e[f][d];
      }
    float sum = 0.f, bmin = FLT_MAX;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float half_l = __ldg(prm + d), offset = __ldg(prm + D + d), shift = __ldg(prm + 2 * D + d);
      const float hw = __ldg(prm + 3 * D + d), basis = __ldg(prm + 4 * D + d);
      const float bounded = __fsub_rn(__fmul_rn(tanhf(__fadd_rn(v[d], shift)), half_l), offset);
      const float q = rintf(bounded);                                  // torch.round: half to even
      const float code = __fdiv_rn(q, hw);
      const float zhat = __fadd_rn(__fmul_rn(code, hw), hw);           // _scale_and_shift, operation for operation
      sum = __fadd_rn(sum, __fmul_rn(zhat, basis));
      bmin = fminf(bmin, 0.5f - fabsf(bounded - q));
      if (codes) codes[(size_t)(n0 + lane) * D + d] = code;
    }
    idx[n0 + lane] = (int32_t)sum;
    if (boundary) boundary[n0 + lane] = bmin;
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
    case 1: { constexpr int DD = 1; __VA_ARGS__; } break;               \
    case 2: { constexpr int DD = 2; __VA_ARGS__; } break;               \
    case 3: { constexpr int DD = 3; __VA_ARGS__; } break;               \
    case 4: { constexpr int DD = 4; __VA_ARGS__; } break;               \
    case 5: { constexpr int DD = 5; __VA_ARGS__; } break;               \
    case 6: { constexpr int DD = 6; __VA_ARGS__; } break;               \
    case 7: { constexpr int DD = 7; __VA_ARGS__; } break;               \
    case 8: { constexpr int DD = 8; __VA_ARGS__; } break;               \
    case 16: { constexpr int DD = 16; __VA_ARGS__; } break;             \
    default: return bc::fail(BC_EUNSUPPORTED, "vq: codebook_dim=%d not in {1..8,16}", D_); \
  }

extern "C" int bc_vq_encode(const float* z, const float* w_in, const float* b_in, const float* cb_norm, int32_t* idx,
                            float* margin, float* z_e, int N, int C, int D, int Kc, bc_stream_t s) {
  BC_REQUIRE(z && cb_norm && idx, "vq_encode: null pointer");
  BC_REQUIRE(N > 0 && C > 0 && D > 0 && Kc >= 2, "vq_encode: bad shape N=%d C=%d D=%d K=%d", N, C, D, Kc);
  BC_REQUIRE((w_in == nullptr) == (b_in == nullptr), "vq_encode: w_in and b_in must both be given or both NULL");
  if (!w_in) BC_REQUIRE(C == D, "vq_encode: identity projection needs C == D (C=%d D=%d)", C, D);
  BC_REQUIRE(bc::aligned16(cb_norm), "vq_encode: codebook must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
  if (D == SC_D && N >= 8192 && bc::aligned16(z)) {
    // register-tiled scan over shared-memory code tiles: 8 frames per thread when that still fills the chip, 2 below
    // (four times the CTAs at a lower FMA density); tiny calls keep the per-warp kernel (more CTAs than SMs there)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    typedef void (*scan_fn)(const float*, const float*, const float*, const float*, int32_t*, float*, float*, int, int, int);
    const scan_fn k8 = margin ? vq_scan_kernel<true, 8> : vq_scan_kernel<false, 8>;
    const scan_fn k2 = margin ? vq_scan_kernel<true, 2> : vq_scan_kernel<false, 2>;
    static bool configured[64][2] = {{false}};
    if (dev < 0 || dev >= 64 || !configured[dev][margin ? 1 : 0]) {
      cudaError_t e = cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM);
      if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(vq_scan)");
      if (dev >= 0 && dev < 64) configured[dev][margin ? 1 : 0] = true;
    }
    // whole waves of 2048-frame CTAs (two per SM) first; what is left -- or an input smaller than one such wave -- goes to
    // 512-frame CTAs, so that neither a tiny grid nor a nearly empty last wave leaves most of the chip idle
    const long long wave8 = (long long)SC_THREADS * 8 * 2 * sms;
    const long long n8 = ((long long)N / wave8) * wave8;
    if (n8 > 0) {
      k8<<<(unsigned)(n8 / (SC_THREADS * 8)), SC_THREADS, SC_SMEM, st>>>(z, w_in, b_in, cb_norm, idx, margin, z_e, (int)n8, C, Kc);
      BC_LAUNCH_CHECK("vq_scan_kernel");
    }
    const long long rem = (long long)N - n8;
    if (rem > 0) {
      k2<<<(unsigned)((rem + SC_THREADS * 2 - 1) / (SC_THREADS * 2)), SC_THREADS, SC_SMEM, st>>>(
          z + (size_t)n8 * C, w_in, b_in, cb_norm, idx + n8, margin ? margin + n8 : nullptr, z_e ? z_e + (size_t)n8 * D : nullptr, (int)rem, C, Kc);
      BC_LAUNCH_CHECK("vq_scan_kernel");
    }
    return BC_OK;
  }
  const size_t smem = w_in ? (size_t)D * C * sizeof(float) : 0;
  if (smem > 200 * 1024) return bc::fail(BC_EUNSUPPORTED, "vq_encode: in_proj %dx%d does not fit shared memory", D, C);
  const int frames_per_cta = VQ_WARPS * FPW;
  const unsigned grid = (unsigned)((N + frames_per_cta - 1) / frames_per_cta);
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

extern "C" int bc_fsq_encode(const float* z, const float* w_in, const float* b_in, const float* params5xd, int32_t* idx,
                             float* codes, float* boundary, int N, int C, int D, bc_stream_t s) {
  BC_REQUIRE(z && params5xd && idx, "fsq_encode: null pointer");
  BC_REQUIRE(N > 0 && C > 0 && D > 0, "fsq_encode: bad shape N=%d C=%d D=%d", N, C, D);
  if (!w_in) BC_REQUIRE(C == D, "fsq_encode: identity projection needs C == D (C=%d D=%d)", C, D);
  const unsigned grid = (unsigned)((N + VQ_WARPS * FPW - 1) / (VQ_WARPS * FPW));
  cudaStream_t st = (cudaStream_t)s;
  switch (D) {
    case 1: fsq_encode_kernel<1><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    case 2: fsq_encode_kernel<2><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    case 3: fsq_encode_kernel<3><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    case 4: fsq_encode_kernel<4><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    case 5: fsq_encode_kernel<5><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    case 6: fsq_encode_kernel<6><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    case 7: fsq_encode_kernel<7><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    case 8: fsq_encode_kernel<8><<<grid, VQ_WARPS * 32, 0, st>>>(z, w_in, b_in, params5xd, idx, codes, boundary, N, C); break;
    default: return bc::fail(BC_EUNSUPPORTED, "fsq_encode: %d levels (supported: 1..8)", D);
  }
  BC_LAUNCH_CHECK("fsq_encode_kernel");
  return BC_OK;
}
