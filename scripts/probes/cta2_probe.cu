// Probe for round 2: tcgen05.mma.cta_group::2 (one MMA of M = 256 across a CTA pair, each SM holding its 128 rows of A and
// HALF of the B operand) with the K-major no-swizzle shared-memory layout of the conv kernels.  Checks D = A * B^T against
// the host and times back-to-back 2-CTA MMAs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cta2_probe cta2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../audiotokenization_b200/csrc/tc_common.cuh"
namespace bc { void set_error(const char*, ...) {} }
using namespace bc::tc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma2_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

// A [256][K], B [N][K] bf16 row-major (K contiguous); D [256][N] fp32.  One cluster = one CTA pair.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, long long* cycles, int N, int K, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int planes = K / 8, nh = N / 2;
  uint8_t* sA = smem;                                   // [planes][128][16 B]
  uint8_t* sB = smem + (size_t)planes * 128 * 16;       // [planes][N/2][16 B]
  for (int i = tid; i < planes * 128; i += 128) {
    const int pl = i / 128, r = i % 128;
    *reinterpret_cast<uint4*>(sA + (size_t)i * 16) = *reinterpret_cast<const uint4*>(A + (size_t)(rank * 128 + r) * K + pl * 8);
  }
  for (int i = tid; i < planes * nh; i += 128) {
    const int pl = i / nh, r = i % nh;
    *reinterpret_cast<uint4*>(sB + (size_t)i * 16) = *reinterpret_cast<const uint4*>(B + (size_t)(rank * nh + r) * K + pl * 8);
  }
  fence_async_smem();
  uint32_t cols = 32;
  while (cols < (uint32_t)N) cols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs: operands staged, barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  const uint32_t hi_d = desc_hi(128u);
  long long t0 = 0, t1 = 0;
  if (rank == 0 && warp == 0) {
    if (elect_one()) {
      const uint32_t a_lo0 = desc_lo(smem_u32(sA), 128u * 16u), b_lo0 = desc_lo(smem_u32(sB), (uint32_t)nh * 16u);
      const uint32_t a_g = (2u * 128u * 16u) >> 4, b_g = (2u * (uint32_t)nh * 16u) >> 4;
      t0 = clock64();
      for (int rep = 0; rep < reps; ++rep)
        for (int k = 0; k < K / 16; ++k)
          mma2_bf16(tmem, a_lo0 + (uint32_t)k * a_g, b_lo0 + (uint32_t)k * b_g, hi_d, hi_d, idesc, (rep | k) ? 1u : 0u);
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar_done), 0);
  tc_fence_after();
  if (rank == 0 && warp == 0) {
    const long long t2 = clock64();
    unsigned m = __ballot_sync(0xffffffffu, t0 != 0);
    const int src = __ffs(m) - 1;
    const long long a = __shfl_sync(0xffffffffu, t0, src), b = __shfl_sync(0xffffffffu, t1, src);
    if (lane == 0 && blockIdx.x == 0) { cycles[0] = b - a; cycles[1] = t2 - a; }
  }
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_load32(tmem + (uint32_t)c0 + ((uint32_t)(warp * 32) << 16), r);
    for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(r[j]) / (float)reps;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody frees TMEM / exits while the peer may still be reading or signalling
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}

int main() {
  for (int N : {64, 128, 256}) {
    const int K = 64, M = 256;
    std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    uint32_t s = 12345u + N;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)((s >> 9) & 0xff) - 128) / 64.0f; };
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD; long long* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dC, 16);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(K / 8) * 128 * 16 + (size_t)(K / 8) * (N / 2) * 16;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int reps : {1, 256}) {
      cudaMemset(dD, 0, (size_t)M * N * 4);
      probe<<<2, 128, smem>>>(dA, dB, dD, dC, N, K, reps);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d reps=%d: CUDA error %s\n", N, reps, cudaGetErrorString(e)); return 1; }
      std::vector<float> hD((size_t)M * N);
      long long cyc[2];
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
          maxerr = fmax(maxerr, fabs(ref - hD[(size_t)m * N + n]));
          maxref = fmax(maxref, fabs(ref));
        }
      const int n_mma = reps * K / 16;
      printf("cta_group::2 M=256 N=%3d K=%d reps=%3d: max |err| %.3g (max |ref| %.3g)  issue %lld cyc (%.1f/mma)  complete %lld cyc (%.1f/mma)\n",
             N, K, reps, maxerr, maxref, cyc[0], (double)cyc[0] / n_mma, cyc[1], (double)cyc[1] / n_mma);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  }
  return 0;
}
