// Exact-float32 dense 1-D convolution on CUDA cores (the parity mode, BC_PREC_FP32).
//
// One kernel covers every dense contraction of the path (SURVEY.md section 8 rows a3-a5,
// a10-a11): dilated k7 "same" convs, 1x1 convs, strided down convs (k = 2s), the
// 2-tap phase convs a transposed conv decomposes into, the 1->C and C->1 edge convs and
// the LSTM input projection (k = 1).  SnakeBeta is applied while the input slab is
// staged (so the activation never makes an HBM round trip), bias / residual / tanh in
// the epilogue.
//
// Tiling: CTA = 128 output steps x BN output channels, 128 threads, each thread an
// 8 x (BN/8) register tile; input channels are consumed in chunks of 16 through a
// shared-memory slab holding every input row the tile's taps touch.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;     // output time steps per CTA
constexpr int CK = 16;      // input-channel chunk
constexpr int CKP = CK + 1; // padded slab row (bank-conflict-free column reads)
constexpr int NTHREADS = 128;

struct ConvParams {
  const float* x;
  const float* w;
  const float* bias;
  const float* sa;
  const float* sib;
  const float* res;
  float* y;
  int B, T_in, C_in, T_out, C_out, K, stride, dil, pad_left;
  int y_rows, y_tstride, y_toffset, flags;
  int slab_rows;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS) conv1d_f32_kernel(const ConvParams p) {
  constexpr int TN = BN / 8;  // output channels per thread
  extern __shared__ __align__(16) float smem[];
  float* slab = smem;                          // [slab_rows][CKP]
  float* ws = smem + (size_t)p.slab_rows * CKP; // [K][CK][BN]
  // keep ws 16-byte aligned
  ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;
  const float* xb = p.x + (size_t)b * p.T_in * p.C_in;
  const int g0 = t0 * p.stride - p.pad_left;  // global input row of slab row 0
  const bool snake = (p.flags & BC_CONV_SNAKE_IN) != 0;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int arow[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) arow[i] = (ty + 16 * i) * p.stride * CKP;

  const bool vec_x = (p.C_in % 4 == 0) && bc::aligned16(p.x);
  const bool vec_w = (p.C_out % 4 == 0) && bc::aligned16(p.w);

  for (int ci0 = 0; ci0 < p.C_in; ci0 += CK) {
    __syncthreads();  // previous chunk fully consumed
    // ---- stage the input slab (activation fused) ----
    if (vec_x) {
      const int c4 = (tid & 3) * 4;
      const int ci = ci0 + c4;
      const bool cok = ci < p.C_in;  // C_in % 4 == 0 => whole float4 in range
      float4 a4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (snake && cok) {
        a4 = __ldg(reinterpret_cast<const float4*>(p.sa + ci));
        b4 = __ldg(reinterpret_cast<const float4*>(p.sib + ci));
      }
      for (int r = tid >> 2; r < p.slab_rows; r += NTHREADS / 4) {
        const int g = g0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cok && g >= 0 && g < p.T_in) {
          v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)g * p.C_in + ci));
          if (snake) {
            v.x = bc::snake_ref(v.x, a4.x, b4.x);
            v.y = bc::snake_ref(v.y, a4.y, b4.y);
            v.z = bc::snake_ref(v.z, a4.z, b4.z);
            v.w = bc::snake_ref(v.w, a4.w, b4.w);
          }
        }
        float* d = slab + r * CKP + c4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      const int c = tid & 15;
      const int ci = ci0 + c;
      const bool cok = ci < p.C_in;
      float a = 1.f, ib = 0.f;
      if (snake && cok) { a = __ldg(p.sa + ci); ib = __ldg(p.sib + ci); }
      for (int r = tid >> 4; r < p.slab_rows; r += NTHREADS / 16) {
        const int g = g0 + r;
        float v = 0.f;
        if (cok && g >= 0 && g < p.T_in) {
          v = __ldg(xb + (size_t)g * p.C_in + ci);
          if (snake) v = bc::snake_ref(v, a, ib);
        }
        slab[r * CKP + c] = v;
      }
    }
    // ---- stage the weight tile w[k][ci0+c][co0+n] ----
    if (vec_w) {
      constexpr int N4 = BN / 4;
      const int total = p.K * CK * N4;
      for (int e = tid; e < total; e += NTHREADS) {
        const int n4 = e % N4;
        const int kc = e / N4;
        const int c = kc % CK, k = kc / CK;
        const int ci = ci0 + c, co = co0 + n4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ci < p.C_in && co < p.C_out)
          v = __ldg(reinterpret_cast<const float4*>(p.w + ((size_t)k * p.C_in + ci) * p.C_out + co));
        *reinterpret_cast<float4*>(ws + (size_t)kc * BN + n4 * 4) = v;
      }
    } else {
      const int total = p.K * CK * BN;
      for (int e = tid; e < total; e += NTHREADS) {
        const int n = e % BN;
        const int kc = e / BN;
        const int c = kc % CK, k = kc / CK;
        const int ci = ci0 + c, co = co0 + n;
        float v = 0.f;
        if (ci < p.C_in && co < p.C_out) v = __ldg(p.w + ((size_t)k * p.C_in + ci) * p.C_out + co);
        ws[(size_t)kc * BN + n] = v;
      }
    }
    __syncthreads();

    // ---- FFMA main loop ----
    for (int k = 0; k < p.K; ++k) {
      const float* sx = slab + k * p.dil * CKP;
      const float* sw = ws + (size_t)k * CK * BN + tx * 4;
#pragma unroll
      for (int c = 0; c < CK; ++c) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = sx[arow[i] + c];
        float bv[TN];
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(sw + c * BN + (j >> 2) * 32);
          bv[j] = t4.x; bv[j + 1] = t4.y; bv[j + 2] = t4.z; bv[j + 3] = t4.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue: bias, residual, tanh, store ----
  // thread's channel j lives at column (j/4)*32 + tx*4 + j%4 of the tile (conflict-free
  // float4 smem reads above, 128-byte contiguous row segments per 8 lanes here)
  const bool tanh_out = (p.flags & BC_CONV_TANH_OUT) != 0;
  const bool vec_ok = (p.C_out % 4 == 0) && bc::aligned16(p.y) && (!p.res || bc::aligned16(p.res));
#pragma unroll
  for (int jg = 0; jg < TN / 4; ++jg) {
    const int co = co0 + jg * 32 + tx * 4;
    if (co >= p.C_out) continue;
    float bias[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) bias[e] = (p.bias && co + e < p.C_out) ? __ldg(p.bias + co + e) : 0.f;
    const bool vec_y = vec_ok && (co + 4 <= p.C_out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 + ty + 16 * i;
      if (t >= p.T_out) continue;
      const size_t row = (size_t)b * p.y_rows + (size_t)t * p.y_tstride + p.y_toffset;
      float* yp = p.y + row * p.C_out + co;
      const float* rp = p.res ? p.res + row * p.C_out + co : nullptr;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = acc[i][jg * 4 + e] + bias[e];
      if (vec_y) {
        if (rp) {
          const float4 r4 = *reinterpret_cast<const float4*>(rp);
          v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
        }
        if (tanh_out) { v[0] = tanhf(v[0]); v[1] = tanhf(v[1]); v[2] = tanhf(v[2]); v[3] = tanhf(v[3]); }
        *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (co + e < p.C_out) {
            float o = v[e];
            if (rp) o += rp[e];
            if (tanh_out) o = tanhf(o);
            yp[e] = o;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Edge convs.  The 1 -> C stem (vq/codec_encoder.py:35) and the C -> 1 tail (vq/codec_decoder.py:77) are
// HBM-bound streaming stencils, not contractions: dedicated kernels, exact fp32.
// ---------------------------------------------------------------------------
// stem: y[b][t][co] = bias[co] + sum_k w[k][co] * x[b][t + k - pad_left]      (C_in == 1, stride 1)
// thread = (time step, 4 output channels): a warp writes 4 steps x 128 B... contiguous rows of the output.
template <int K>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int T_in,
                                                        int T_out, int C_out, int pad_left, long long total) {
  const int groups = C_out >> 2;                       // float4 groups per row
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int cg = (int)(e % groups);
    const long long row = e / groups;                  // b * T_out + t
    const int t = (int)(row % T_out);
    const long long b = row / T_out;
    const float* xb = x + b * T_in;
    float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int g = t + k - pad_left;
      const float xv = (g >= 0 && g < T_in) ? __ldg(xb + g) : 0.f;
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * C_out) + cg);
      acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
    }
    __stcs(reinterpret_cast<float4*>(y + row * C_out) + cg, acc);
  }
}

// stem, sliding-window form: thread = (run of STEM_RUN consecutive steps, 4 output channels).  The K weight quads and the
// bias stay in registers, the input window slides by one sample per step (one 4-byte load, shared by the lanes of
// the row), and the G = C_out/4 lanes of a row store one contiguous output row per step -- no per-element index
// arithmetic, no weight reloads.  G must be a power of two <= 32.
constexpr int STEM_RUN = 32;
template <int K>
__global__ void __launch_bounds__(256) stem_conv_run_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y, int T_in,
                                                            int T_out, int C_out, int pad_left, int runs_per_item,
                                                            long long total_runs) {
  const int G = C_out >> 2;
  const int cg = threadIdx.x & (G - 1);
  const int runs_per_block = 256 / G;
  float4 wv[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wv[k] = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * C_out) + cg);
  const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long run = (long long)blockIdx.x * runs_per_block + threadIdx.x / G; run < total_runs;
       run += (long long)gridDim.x * runs_per_block) {
    const long long b = run / runs_per_item;
    const int t0 = (int)(run - b * runs_per_item) * STEM_RUN;
    const float* xb = x + b * T_in;
    float4* yb = reinterpret_cast<float4*>(y + ((size_t)b * T_out + t0) * C_out) + cg;
    float win[K];
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
      const int g = t0 + k - pad_left;
      win[k] = (g >= 0 && g < T_in) ? __ldg(xb + g) : 0.f;
    }
    const int n = min(STEM_RUN, T_out - t0);
    for (int i = 0; i < n; ++i) {
      const int g = t0 + i + K - 1 - pad_left;
      win[K - 1] = (g >= 0 && g < T_in) ? __ldg(xb + g) : 0.f;
      float4 acc = bv;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        acc.x = fmaf(win[k], wv[k].x, acc.x); acc.y = fmaf(win[k], wv[k].y, acc.y);
        acc.z = fmaf(win[k], wv[k].z, acc.z); acc.w = fmaf(win[k], wv[k].w, acc.w);
      }
      __stcs(yb + (size_t)i * G, acc);
#pragma unroll
      for (int k = 0; k < K - 1; ++k) win[k] = win[k + 1];
    }
  }
}

// tail: y[b][t] = act(bias + sum_k sum_ci w[k][ci] * snake?(x[b][t + k - pad_left][ci]))   (C_out == 1, stride 1)
// CTA = 256 consecutive output steps of one item.  The activated input rows are staged ONCE in shared
// memory (row stride C_in + 1: conflict-free column walks), then thread t walks its K x C_in window.
constexpr int TAIL_TT = 256;
__global__ void __launch_bounds__(TAIL_TT) tail_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, const float* __restrict__ sa,
                                                            const float* __restrict__ sib, float* __restrict__ y, int T_in,
                                                            int T_out, int C_in, int K, int pad_left, int flags) {
  extern __shared__ float tail_smem[];
  const int ld = C_in + 1;
  const int rows = TAIL_TT + K - 1;
  float* sx = tail_smem;               // [rows][ld]
  float* sw = sx + (size_t)rows * ld;  // [K][C_in]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TAIL_TT;
  const float* xb = x + (size_t)b * T_in * C_in;
  const bool snake = (flags & BC_CONV_SNAKE_IN) != 0;
  for (int i = threadIdx.x; i < K * C_in; i += TAIL_TT) sw[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < rows * C_in; i += TAIL_TT) {
    const int r = i / C_in, c = i - r * C_in;
    const int g = t0 + r - pad_left;
    float v = 0.f;
    if (g >= 0 && g < T_in) {
      v = __ldg(xb + (size_t)g * C_in + c);
      if (snake) v = bc::snake_ref(v, __ldg(sa + c), __ldg(sib + c));
    }
    sx[r * ld + c] = v;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= T_out) return;
  float acc = bias ? __ldg(bias) : 0.f;
  for (int k = 0; k < K; ++k) {
    const float* row = sx + (size_t)(threadIdx.x + k) * ld;
    const float* wk = sw + k * C_in;
#pragma unroll 8
    for (int c = 0; c < C_in; ++c) acc = fmaf(row[c], wk[c], acc);
  }
  if (flags & BC_CONV_TANH_OUT) acc = tanhf(acc);
  y[(size_t)b * T_out + t] = acc;
}

// tail, warp-per-block form (C_in <= 32, the codec's C -> 1 waveform conv): a warp produces 32 consecutive outputs; lane c
// owns input channel c and its K weights, walks the 32 + K - 1 input rows of the block (one coalesced 128-byte row per
// load), applies SnakeBeta ONCE per element and adds w[k][c] * s to the (at most K) outputs the row feeds -- 32 partial
// sums in registers, all indices compile-time.  A 31-shuffle transposing reduction then leaves output j in lane j, so the
// block leaves as one coalesced 128-byte store.  ~22 warp-instructions per output (the staged kernel above: ~53).
// SnakeBeta uses the range-reduced SFU sine of the tensor-core kernels (|error| ~ 4e-7 absolute).
template <int K>
__global__ void __launch_bounds__(256) tail_conv_warp_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ sa,
                                                             const float* __restrict__ sib, float* __restrict__ y, int T_in,
                                                             int T_out, int C_in, int pad_left, int flags, int blocks_per_item,
                                                             long long total_blocks) {
  const int lane = threadIdx.x & 31;
  const long long wb = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wb >= total_blocks) return;
  const int b = (int)(wb / blocks_per_item);
  const int tb = (int)(wb - (long long)b * blocks_per_item) * 32;
  const bool ch_ok = lane < C_in;
  const bool snake = (flags & BC_CONV_SNAKE_IN) != 0;
  float wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wk[k] = ch_ok ? __ldg(w + k * C_in + lane) : 0.f;
  const float a = (snake && ch_ok) ? __ldg(sa + lane) : 0.f, ib = (snake && ch_ok) ? __ldg(sib + lane) : 0.f;
  const float* xb = x + (size_t)b * T_in * C_in + lane;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  // rows r = tb - pad_left + i, i in [0, 32 + K - 1): row i feeds output j = i - k through tap k
  constexpr int ROWS = 32 + K - 1;
  float xv[ROWS];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int g = tb - pad_left + i;
    xv[i] = (ch_ok && g >= 0 && g < T_in) ? __ldg(xb + (size_t)g * C_in) : 0.f;     // zero padding AFTER the activation: snake(0) = 0
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const float sv = snake ? bc::tc::snake_tc(xv[i], a, ib) : xv[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int j = i - k;
      if (j >= 0 && j < 32) acc[j] = fmaf(wk[k], sv, acc[j]);
    }
  }
  // transposing reduction: after the step with offset o, a lane keeps the partial sums of the outputs whose bit o equals its own
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? acc[i] : acc[i + o];
      const float keep = up ? acc[i + o] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  const int t = tb + lane;
  if (t < T_out) {
    float v = acc[0] + (bias ? __ldg(bias) : 0.f);
    if (flags & BC_CONV_TANH_OUT) v = tanhf(v);
    y[(size_t)b * T_out + t] = v;
  }
}

template <int BN>
int launch(const ConvParams& p, cudaStream_t st) {
  const size_t smem = ((size_t)p.slab_rows * CKP + (size_t)p.K * CK * BN) * sizeof(float) + 16;
  if (smem > 227 * 1024) return bc::fail(BC_EUNSUPPORTED, "conv1d(fp32): tile needs %zu B of shared memory (K=%d stride=%d dil=%d)", smem, p.K, p.stride, p.dil);
  if (smem > 48 * 1024) {
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(conv1d_f32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(conv1d_f32)");
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  dim3 grid((p.T_out + BM - 1) / BM, (p.C_out + BN - 1) / BN, p.B);
  conv1d_f32_kernel<BN><<<grid, NTHREADS, smem, st>>>(p);
  BC_LAUNCH_CHECK("conv1d_f32_kernel");
  return BC_OK;
}

}  // namespace

namespace bc {
int conv1d_tc_fwd(const float* x, const float* w, const float* bias, const float* snake_a, const float* snake_ib,
                  const float* res, float* y, int B, int T_in, int C_in, int T_out, int C_out, int K, int stride,
                  int dilation, int pad_left, int y_rows, int y_tstride, int y_toffset, int flags, int precision,
                  cudaStream_t st);
}

extern "C" int bc_conv1d_fwd(const float* x, const float* w, const float* bias, const float* snake_a,
                             const float* snake_ib, const float* res, float* y, int B, int T_in, int C_in, int T_out,
                             int C_out, int K, int stride, int dilation, int pad_left, int y_rows, int y_tstride,
                             int y_toffset, int flags, int precision, bc_stream_t s) {
  BC_REQUIRE(x && w && y, "conv1d: null pointer");
  BC_REQUIRE(B > 0 && T_in > 0 && C_in > 0 && T_out > 0 && C_out > 0, "conv1d: bad shape B=%d T_in=%d C_in=%d T_out=%d C_out=%d", B, T_in, C_in, T_out, C_out);
  BC_REQUIRE(K > 0 && K <= 64 && stride > 0 && dilation > 0, "conv1d: bad K=%d stride=%d dilation=%d", K, stride, dilation);
  BC_REQUIRE(y_tstride > 0 && y_toffset >= 0 && y_rows >= (T_out - 1) * y_tstride + y_toffset + 1, "conv1d: output geometry y_rows=%d tstride=%d toffset=%d T_out=%d", y_rows, y_tstride, y_toffset, T_out);
  BC_REQUIRE(B <= 65535, "conv1d: B=%d > 65535 (split the batch)", B);
  if (flags & BC_CONV_SNAKE_IN) BC_REQUIRE(snake_a && snake_ib, "conv1d: BC_CONV_SNAKE_IN without snake parameters");
  if (precision != BC_PREC_FP32)
    return bc::conv1d_tc_fwd(x, w, bias, snake_a, snake_ib, res, y, B, T_in, C_in, T_out, C_out, K, stride, dilation,
                             pad_left, y_rows, y_tstride, y_toffset, flags, precision, (cudaStream_t)s);
  cudaStream_t st_ = (cudaStream_t)s;
  // ---- edge convs: streaming stencils ----
  if (C_in == 1 && stride == 1 && dilation == 1 && C_out % 4 == 0 && !(flags & (BC_CONV_SNAKE_IN | BC_CONV_TANH_OUT)) && !res &&
      y_tstride == 1 && y_toffset == 0 && y_rows == T_out && (K == 7 || K == 3 || K == 1) && bc::aligned16(w) && bc::aligned16(y) &&
      (!bias || bc::aligned16(bias))) {
    const int G = C_out / 4;
    if (K == 7 && G <= 32 && (G & (G - 1)) == 0) {   // the codec's stem (k = 7, C_out = 16 / 32 / 64 / 128): sliding-window kernel
      const int runs_per_item = (T_out + STEM_RUN - 1) / STEM_RUN;
      const long long total_runs = (long long)B * runs_per_item;
      const long long want = (total_runs + (256 / G) - 1) / (256 / G);
      const unsigned blocks = (unsigned)(want < 148ll * 16 ? want : 148ll * 16);
      stem_conv_run_kernel<7><<<blocks, 256, 0, st_>>>(x, w, bias, y, T_in, T_out, C_out, pad_left, runs_per_item, total_runs);
      BC_LAUNCH_CHECK("stem_conv_run_kernel");
      return BC_OK;
    }
    const long long total = (long long)B * T_out * (C_out / 4);
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148ll * 32 ? (total + 255) / 256 : 148ll * 32);
    if (K == 7) stem_conv_kernel<7><<<blocks, 256, 0, st_>>>(x, w, bias, y, T_in, T_out, C_out, pad_left, total);
    else if (K == 3) stem_conv_kernel<3><<<blocks, 256, 0, st_>>>(x, w, bias, y, T_in, T_out, C_out, pad_left, total);
    else stem_conv_kernel<1><<<blocks, 256, 0, st_>>>(x, w, bias, y, T_in, T_out, C_out, pad_left, total);
    BC_LAUNCH_CHECK("stem_conv_kernel");
    return BC_OK;
  }
  if (C_out == 1 && stride == 1 && dilation == 1 && !res && y_tstride == 1 && y_toffset == 0 && y_rows == T_out && K == 7 &&
      C_in <= 32 && C_in >= 8) {
    const int blocks_per_item = (T_out + 31) / 32;
    const long long total_blocks = (long long)blocks_per_item * B;
    const unsigned grid = (unsigned)((total_blocks + 7) / 8);
    tail_conv_warp_kernel<7><<<grid, 256, 0, st_>>>(x, w, bias, snake_a, snake_ib, y, T_in, T_out, C_in, pad_left, flags,
                                                   blocks_per_item, total_blocks);
    BC_LAUNCH_CHECK("tail_conv_warp_kernel");
    return BC_OK;
  }
  if (C_out == 1 && stride == 1 && dilation == 1 && !res && y_tstride == 1 && y_toffset == 0 && y_rows == T_out &&
      (size_t)((TAIL_TT + K - 1) * (C_in + 1) + K * C_in) * sizeof(float) <= 160 * 1024) {
    const size_t smem = (size_t)((TAIL_TT + K - 1) * (C_in + 1) + K * C_in) * sizeof(float);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(tail_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(tail_conv)");
    }
    dim3 grid((T_out + TAIL_TT - 1) / TAIL_TT, B);
    tail_conv_kernel<<<grid, TAIL_TT, smem, st_>>>(x, w, bias, snake_a, snake_ib, y, T_in, T_out, C_in, K, pad_left, flags);
    BC_LAUNCH_CHECK("tail_conv_kernel");
    return BC_OK;
  }
  ConvParams p;
  p.x = x; p.w = w; p.bias = bias; p.sa = snake_a; p.sib = snake_ib; p.res = res; p.y = y;
  p.B = B; p.T_in = T_in; p.C_in = C_in; p.T_out = T_out; p.C_out = C_out; p.K = K; p.stride = stride; p.dil = dilation;
  p.pad_left = pad_left; p.y_rows = y_rows; p.y_tstride = y_tstride; p.y_toffset = y_toffset; p.flags = flags;
  p.slab_rows = (BM - 1) * stride + (K - 1) * dilation + 1;
  if (C_out > 32) return launch<64>(p, (cudaStream_t)s);
  return launch<32>(p, (cudaStream_t)s);
}

extern "C" int bc_convtr1d_fwd(const float* x, const float* w_phases, const float* bias, const float* snake_a,
                               const float* snake_ib, float* y, int B, int T_in, int C_in, int C_out, int stride,
                               int padding, int flags, int precision, bc_stream_t s) {
  BC_REQUIRE(x && w_phases && y, "convtr1d: null pointer");
  BC_REQUIRE(stride >= 2 && padding >= 0 && padding < stride, "convtr1d: stride=%d padding=%d (need stride >= 2, 0 <= padding < stride)", stride, padding);
  const int T_out = T_in * stride;
  for (int ph = 0; ph < stride; ++ph) {
    const int q = (ph + padding) / stride;  // 0 or 1
    // y[m*stride + ph] = W[j0+stride]^T x[m+q-1] + W[j0]^T x[m+q]
    // fp32: [phase][2][C_in][C_out] floats; tensor-core modes: one bc_tc_plan image per phase
    // (split * 2*C_in*C_out bf16 = split * C_in*C_out float-sized words)
    const size_t phase_words = precision == BC_PREC_FP32 ? (size_t)2 * C_in * C_out
                                                         : (size_t)(precision == BC_PREC_BF16X3 ? 2 : 1) * C_in * C_out;
    const float* wp = w_phases + (size_t)ph * phase_words;
    int rc = bc_conv1d_fwd(x, wp, bias, snake_a, snake_ib, nullptr, y, B, T_in, C_in, T_in, C_out, /*K=*/2,
                           /*stride=*/1, /*dil=*/1, /*pad_left=*/1 - q, /*y_rows=*/T_out, /*y_tstride=*/stride,
                           /*y_toffset=*/ph, flags, precision, s);
    if (rc != BC_OK) return rc;
  }
  return BC_OK;
}
