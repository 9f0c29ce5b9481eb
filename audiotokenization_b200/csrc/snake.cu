// SnakeBeta and the anti-aliased Activation1d (2x up FIR -> SnakeBeta -> 2x down FIR)
// as ONE pass over a channels-last [B][T][C] tensor: one read and one write per element.
//
// Math (SURVEY.md section 8 row a6'; verified against the reference in tests):
//   u[m] = 2 * sum_j x[clamp(j,0,T-1)] * f[m + 5 - 2j],   0 <= m+5-2j <= 11      (resample.py:25-33)
//   s[m] = u[m] + ib * sin(a*u[m])^2                                              (activations.py:107-119)
//   y[t] = sum_{k<12} s[clamp(2t + k - 5, 0, 2T-1)] * f[k]                        (filter.py:86-95)
#include "common.cuh"
#include "tc_common.cuh"
#include <type_traits>

namespace {

// ---------------------------------------------------------------------------
// plain SnakeBeta: pure streaming, float4 over channels when C % 4 == 0
// ---------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256) snake_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                    const float* __restrict__ sa, const float* __restrict__ sib,
                                                    size_t n_elems, int C) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (VEC4) {
    const size_t n4 = n_elems >> 2;
    const int C4 = C >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    const float4* a4 = reinterpret_cast<const float4*>(sa);
    const float4* b4 = reinterpret_cast<const float4*>(sib);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int c = (int)(i % (size_t)C4);
      const float4 v = __ldcs(x4 + i);
      const float4 a = __ldg(a4 + c), b = __ldg(b4 + c);
      float4 o;
      o.x = bc::snake_ref(v.x, a.x, b.x);
      o.y = bc::snake_ref(v.y, a.y, b.y);
      o.z = bc::snake_ref(v.z, a.z, b.z);
      o.w = bc::snake_ref(v.w, a.w, b.w);
      y4[i] = o;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
      const int c = (int)(i % (size_t)C);
      y[i] = bc::snake_ref(x[i], __ldg(sa + c), __ldg(sib + c));
    }
  }
}

// ---------------------------------------------------------------------------
// anti-aliased variant.  Thread = (batch b, run of RUN output steps, channel c);
// consecutive threads own consecutive channels, so every global access of a warp
// is one contiguous row segment.  Per output step a thread loads ONE new input
// sample and keeps a 6-wide x window and a 12-wide s window in registers.
// ---------------------------------------------------------------------------
constexpr int AA_RUN = 32;

struct Fir12 {
  float f[12];
};

__device__ __forceinline__ float aa_s_at(const float* __restrict__ xc, int C, int T, int m, const Fir12& F, float a,
                                         float ib) {
  // s[clamp(m)] from global memory (window warm-up only)
  m = max(0, min(m, 2 * T - 1));
  const int aa = m >> 1;
  float acc = 0.f;
  if (m & 1) {  // odd: j in [aa-2, aa+3], f index 10,8,...,0
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      int j = max(0, min(aa - 2 + i, T - 1));
      acc = fmaf(__ldg(xc + (size_t)j * C), F.f[10 - 2 * i], acc);
    }
  } else {  // even: j in [aa-3, aa+2], f index 11,9,...,1
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      int j = max(0, min(aa - 3 + i, T - 1));
      acc = fmaf(__ldg(xc + (size_t)j * C), F.f[11 - 2 * i], acc);
    }
  }
  return bc::snake_ref(2.f * acc, a, ib);
}

__global__ void __launch_bounds__(128) snake_aa_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       const float* __restrict__ sa, const float* __restrict__ sib,
                                                       const float* __restrict__ fir, int T, int C, int runs_per_item) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int b = blockIdx.x / runs_per_item;
  const int run = blockIdx.x - b * runs_per_item;
  const int t0 = run * AA_RUN;
  const int t1 = min(T, t0 + AA_RUN);
  Fir12 F;
#pragma unroll
  for (int i = 0; i < 12; ++i) F.f[i] = __ldg(fir + i);
  const float a = __ldg(sa + c), ib = __ldg(sib + c);
  const float* xc = x + (size_t)b * T * C + c;
  float* yc = y + (size_t)b * T * C + c;

  float sw[12];  // s[clamp(2t-5+i)]
#pragma unroll
  for (int i = 0; i < 10; ++i) sw[i] = aa_s_at(xc, C, T, 2 * t0 - 5 + i, F, a, ib);
  float xw[6];  // x[clamp(t+i)]
#pragma unroll
  for (int i = 0; i < 6; ++i) xw[i] = __ldg(xc + (size_t)min(t0 + i, T - 1) * C);

  const int m_last = 2 * T - 1;
  for (int t = t0; t < t1; ++t) {
    // prefetch the next window element early
    const float x_next = __ldg(xc + (size_t)min(t + 6, T - 1) * C);
    // m = 2t+5 (odd, a = t+2: j = t..t+5 -> f[10,8,6,4,2,0]); m = 2t+6 (even, a = t+3: j = t..t+5 -> f[11,9,...,1])
    float uo = 0.f, ue = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      uo = fmaf(xw[i], F.f[10 - 2 * i], uo);
      ue = fmaf(xw[i], F.f[11 - 2 * i], ue);
    }
    sw[10] = (2 * t + 5 <= m_last) ? bc::snake_ref(2.f * uo, a, ib) : sw[9];
    sw[11] = (2 * t + 6 <= m_last) ? bc::snake_ref(2.f * ue, a, ib) : sw[10];
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc = fmaf(sw[k], F.f[k], acc);
    yc[(size_t)t * C] = acc;
#pragma unroll
    for (int i = 0; i < 10; ++i) sw[i] = sw[i + 2];
#pragma unroll
    for (int i = 0; i < 5; ++i) xw[i] = xw[i + 1];
    xw[5] = x_next;
  }
}

// ---------------------------------------------------------------------------
// anti-aliased variant, packed form (C even).  Thread = (batch b, run of AA2_RUN output steps, channel PAIR):
// every quantity is an fp32 pair (fma.rn.f32x2: one issue slot for the two channels), the 6-wide x window and the
// 12-wide s window rotate through registers with compile-time indices (the time loop is unrolled by 6 = one full
// rotation of both, so no register moves), the next six input rows are requested one unrolled block ahead (six
// 8-byte loads in flight per thread), SnakeBeta uses the range-reduced SFU sine of the tensor-core kernels
// (|error| ~ 4e-7), and a run warms its s window up by STARTING FIVE STEPS EARLY with the output suppressed instead
// of re-deriving ten samples from global memory.  Lanes run along channel pairs, so every access of a warp is one or
// more whole contiguous row segments (128 B at C = 32).  The kernel is bound by instruction issue (~26 issue slots per
// element), not by HBM; see DESIGN.md section 4.3.
// ---------------------------------------------------------------------------
#ifndef BC_AA2_RUN
#define BC_AA2_RUN 126
#endif
#ifndef BC_AA2_MINB
#define BC_AA2_MINB 5
#endif
constexpr int AA2_RUN = BC_AA2_RUN;   // output steps per thread (a multiple of 6)
constexpr int AA2_THREADS = 128;

using bc::tc::f32x2;
using bc::tc::pack2;
using bc::tc::unpack2;
using bc::tc::fma2;
using bc::tc::mul2;

__global__ void __launch_bounds__(AA2_THREADS, BC_AA2_MINB) snake_aa2_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                const float* __restrict__ sa, const float* __restrict__ sib,
                                                                const float* __restrict__ fir, int T, int C, int pairs_pb,
                                                                int runs_pb, int runs_per_item, int B) {
  // block = pairs_pb channel pairs x runs_pb runs; grid.x walks (item, run group), grid.y the channel-pair groups
  const int pl = threadIdx.x % pairs_pb, rl = threadIdx.x / pairs_pb;
  const int cp = blockIdx.y * pairs_pb + pl;
  const int c0 = 2 * cp;
  const int groups_per_item = (runs_per_item + runs_pb - 1) / runs_pb;
  const int b = blockIdx.x / groups_per_item;
  const int run = (blockIdx.x - b * groups_per_item) * runs_pb + rl;
  if (c0 >= C || run >= runs_per_item || b >= B) return;
  const int t0 = run * AA2_RUN;
  const int t1 = min(T, t0 + AA2_RUN);
  f32x2 Fd[12], Fu[12];                     // down-filter taps and 2 x up-filter taps, each duplicated over the pair
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float f = __ldg(fir + i);
    Fd[i] = pack2(f, f);
    Fu[i] = pack2(2.f * f, 2.f * f);
  }
  const f32x2 a2 = pack2(__ldg(sa + c0), __ldg(sa + c0 + 1)), ib2 = pack2(__ldg(sib + c0), __ldg(sib + c0 + 1));
  const float* xc = x + (size_t)b * T * C + c0;
  float* yc = y + (size_t)b * T * C + c0;
  const int m_last = 2 * T - 1;
  auto ldx = [&](int t) -> f32x2 {
    t = max(0, min(t, T - 1));
    return __ldg(reinterpret_cast<const unsigned long long*>(xc + (size_t)t * C));
  };

  f32x2 xw[6], sw[12], xq[6];
  int tb = t0 - 6;                           // first unrolled block: the warm-up (its outputs are suppressed)
#pragma unroll
  for (int i = 0; i < 6; ++i) xw[i] = ldx(tb + i);
#pragma unroll
  for (int i = 0; i < 6; ++i) xq[i] = ldx(tb + 6 + i);
#pragma unroll
  for (int i = 0; i < 12; ++i) sw[i] = pack2(0.f, 0.f);

  // one unrolled block of six steps; SAFE = edge handling (index clamps, suppressed outputs, replicate padding of s),
  // compiled out for interior blocks, which are almost all of them
  auto block = [&](auto safe_tag) {
    constexpr bool SAFE = decltype(safe_tag)::value;
    f32x2 xn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i)             // the window tail of the block after the next one
      xn[i] = SAFE ? ldx(tb + 12 + i) : __ldg(reinterpret_cast<const unsigned long long*>(xc + (size_t)(tb + 12 + i) * C));
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int t = tb + s;
      // window invariants at step t (indices are compile-time after unrolling):
      //   x[clamp(t + i)] = xw[(s + i) % 6],   s[clamp(2t - 5 + i)] = sw[(2s + i) % 12] for i < 10
      f32x2 uo = pack2(0.f, 0.f), ue = uo;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        uo = fma2(xw[(s + i) % 6], Fu[10 - 2 * i], uo);         // m = 2t+5 (odd):  j = t..t+5 -> f[10,8,...,0]
        ue = fma2(xw[(s + i) % 6], Fu[11 - 2 * i], ue);         // m = 2t+6 (even): j = t..t+5 -> f[11,9,...,1]
      }
      const f32x2 so = bc::tc::snake_tc2(uo, a2, ib2), se = bc::tc::snake_tc2(ue, a2, ib2);
      const int i10 = (2 * s + 10) % 12, i11 = (2 * s + 11) % 12, i9 = (2 * s + 9) % 12;
      if (SAFE) {
        sw[i10] = (2 * t + 5 <= m_last) ? so : sw[i9];          // replicate padding at the doubled rate
        sw[i11] = (2 * t + 6 <= m_last) ? se : sw[i10];
        if (s == 5 && t == -1) {    // start of the item (t0 = 0): the window holds s[-7 .. 4]; replicate s[0] into m < 0
#pragma unroll
          for (int i = 0; i < 7; ++i) sw[(2 * s + i) % 12] = sw[(2 * s + 7) % 12];
        }
      } else {
        sw[i10] = so;
        sw[i11] = se;
      }
      if (!SAFE || (t >= t0 && t < t1)) {
        f32x2 acc = pack2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 12; ++k) acc = fma2(sw[(2 * s + k) % 12], Fd[k], acc);
        __stcs(reinterpret_cast<unsigned long long*>(yc + (size_t)t * C), acc);
      }
      xw[s % 6] = xq[s];                                        // slot of x[t] becomes x[t + 6]
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) xq[i] = xn[i];
  };
  for (; tb < t1; tb += 6) {
    // interior: every load index in range, every step an output of this run, no s beyond the end of the item
    const bool interior = tb >= t0 && tb + 6 <= t1 && tb + 18 <= T;
    if (interior) block(std::false_type{});
    else block(std::true_type{});
  }
}

}  // namespace

extern "C" int bc_snake_fwd(const float* x, float* y, const float* snake_a, const float* snake_ib,
                            const float* fir12, int B, int T, int C, int antialias, bc_stream_t s) {
  BC_REQUIRE(x && y && snake_a && snake_ib, "snake: null pointer");
  BC_REQUIRE(B > 0 && T > 0 && C > 0, "snake: bad shape B=%d T=%d C=%d", B, T, C);
  cudaStream_t st = (cudaStream_t)s;
  if (!antialias) {
    const size_t n = (size_t)B * T * C;
    const bool vec = (C % 4 == 0) && bc::aligned16(x) && bc::aligned16(y) && bc::aligned16(snake_a) && bc::aligned16(snake_ib);
    size_t work = vec ? n / 4 : n;
    unsigned blocks = (unsigned)((work + 255) / 256 < (size_t)148 * 16 ? (work + 255) / 256 : (size_t)148 * 16);
    if (blocks == 0) blocks = 1;
    if (vec)
      snake_kernel<true><<<blocks, 256, 0, st>>>(x, y, snake_a, snake_ib, n, C);
    else
      snake_kernel<false><<<blocks, 256, 0, st>>>(x, y, snake_a, snake_ib, n, C);
    BC_LAUNCH_CHECK("snake_kernel");
    return BC_OK;
  }
  BC_REQUIRE(fir12 != nullptr, "snake: antialias needs the 12 FIR taps");
  BC_REQUIRE(x != y, "snake: antialias cannot run in place");
  if (C % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 7u) == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0) {
    const int pairs = C / 2;
    int pairs_pb = 1;
    while (pairs_pb * 2 <= pairs && pairs_pb * 2 <= AA2_THREADS && pairs % (pairs_pb * 2) == 0) pairs_pb *= 2;
    if (pairs_pb >= 8) {                                         // whole 64-byte row segments per warp, at least
      const int runs_pb = AA2_THREADS / pairs_pb;
      const int runs_per_item = (T + AA2_RUN - 1) / AA2_RUN;
      const long long gx = (long long)((runs_per_item + runs_pb - 1) / runs_pb) * B;
      BC_REQUIRE(gx <= 2147483647ll && pairs / pairs_pb <= 65535, "snake: grid too large");
      dim3 grid2((unsigned)gx, (unsigned)(pairs / pairs_pb));
      snake_aa2_kernel<<<grid2, AA2_THREADS, 0, st>>>(x, y, snake_a, snake_ib, fir12, T, C, pairs_pb, runs_pb, runs_per_item, B);
      BC_LAUNCH_CHECK("snake_aa2_kernel");
      return BC_OK;
    }
  }
  const int runs = (T + AA_RUN - 1) / AA_RUN;
  const long long gy = (long long)runs * B;
  BC_REQUIRE(gy <= 2147483647ll, "snake: too many runs");
  const int threads = C >= 128 ? 128 : ((C + 31) / 32) * 32;
  dim3 grid((unsigned)gy, (C + threads - 1) / threads);
  BC_REQUIRE(grid.y <= 65535, "snake: too many channels");
  snake_aa_kernel<<<grid, threads, 0, st>>>(x, y, snake_a, snake_ib, fir12, T, C, runs);
  BC_LAUNCH_CHECK("snake_aa_kernel");
  return BC_OK;
}
