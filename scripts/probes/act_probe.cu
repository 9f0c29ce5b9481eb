// Probe: CUDA-core throughput of the activation stage of the tensor-core kernels (SnakeBeta + bf16 hi/lo split + K-major
// stores), per SM sub-partition, as a function of the warps sharing it.  Build + run: scripts/probes/act_probe.sh
#include "../../audiotokenization_b200/csrc/tc_common.cuh"
#include <cstdio>
using namespace bc::tc;

__device__ __forceinline__ void sts64(uint8_t* p, uint2 v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint8_t* p, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(const void* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}
template <int SPLIT>
__device__ __forceinline__ void store_quad(const float4& v, uint8_t* d8, uint32_t lo_offset) {
  uint2 h;
  h.x = pack_bf16x2(v.x, v.y); h.y = pack_bf16x2(v.z, v.w);
  sts64(d8, h);
  if (SPLIT == 2) {
    uint2 l;
    float r0, r1, r2, r3;
    unpack2(sub2(pack2(v.x, v.y), bf16x2_as_f32x2(h.x)), r0, r1);
    unpack2(sub2(pack2(v.z, v.w), bf16x2_as_f32x2(h.y)), r2, r3);
    l.x = pack_bf16x2(r0, r1);
    l.y = pack_bf16x2(r2, r3);
    sts64(d8 + lo_offset, l);
  }
}
// hi by truncation (one LOP per element, PRMT pack), lo = exact remainder rounded to bf16
__device__ __forceinline__ void store_quad_trunc(const float4& v, uint8_t* d8, uint32_t lo_offset) {
  const uint32_t bx = __float_as_uint(v.x) & 0xffff0000u, by = __float_as_uint(v.y) & 0xffff0000u;
  const uint32_t bz = __float_as_uint(v.z) & 0xffff0000u, bw = __float_as_uint(v.w) & 0xffff0000u;
  uint2 h;
  h.x = __byte_perm(bx, by, 0x7632); h.y = __byte_perm(bz, bw, 0x7632);
  sts64(d8, h);
  uint2 l;
  float r0, r1, r2, r3;
  unpack2(sub2(pack2(v.x, v.y), pack2(__uint_as_float(bx), __uint_as_float(by))), r0, r1);
  unpack2(sub2(pack2(v.z, v.w), pack2(__uint_as_float(bz), __uint_as_float(bw))), r2, r3);
  l.x = pack_bf16x2(r0, r1); l.y = pack_bf16x2(r2, r3);
  sts64(d8 + lo_offset, l);
}

// MODE 0: snake (no reduction) + split   1: snake with range reduction + split   2: snake only (raw fp32 store)
//      3: split only   4: loads + stores only   5: snake + truncation split   6: snake scalar (not paired) + split
template <int MODE, int ILP>
__global__ void __launch_bounds__(512, 1) probe(const float* __restrict__ x, float* out, long long* cycles, int iters, int warps_used) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp >= warps_used) return;
  // per-thread source rows in shared memory (filled once), destination = K-major style 16-byte rows
  float4* src = reinterpret_cast<float4*>(smem) + tid;          // row i of the thread: src[i * 512] (lane-contiguous: conflict-free)
  for (int i = 0; i < 8; ++i) src[i * 512] = reinterpret_cast<const float4*>(x)[(tid * 8 + i) % 4096];
  uint8_t* dst = smem + 512 * 8 * 16 + (size_t)warp * 4096 + ((MODE == 2 || MODE == 4) ? lane * 16 : lane * 8);
  const float4 sa = make_float4(1.1f, 0.9f, 1.3f, 0.7f), sb = make_float4(0.8f, 1.2f, 1.0f, 0.9f);
  __syncwarp();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float4 v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = lds128(src + ((it * ILP + j) & 7) * 512);
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      float4 w = v[j];
      if (MODE == 0 || MODE == 2 || MODE == 5) snake4<1>(w, sa, sb);
      if (MODE == 1) {
        const f32x2 y0 = snake_tc2(pack2(w.x, w.y), pack2(sa.x, sa.y), pack2(sb.x, sb.y));
        const f32x2 y1 = snake_tc2(pack2(w.z, w.w), pack2(sa.z, sa.w), pack2(sb.z, sb.w));
        unpack2(y0, w.x, w.y); unpack2(y1, w.z, w.w);
      }
      if (MODE == 6) { w.x = snake_bf(w.x, sa.x, sb.x); w.y = snake_bf(w.y, sa.y, sb.y); w.z = snake_bf(w.z, sa.z, sb.z); w.w = snake_bf(w.w, sa.w, sb.w); }
      if (MODE == 2 || MODE == 4) sts128(dst + j * 1024, w);
      else if (MODE == 5) store_quad_trunc(w, dst + j * 1024, 256);
      else store_quad<2>(w, dst + j * 1024, 256);
    }
  }
  __syncwarp();
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  if (tid == 0 && iters < 0) out[0] = *reinterpret_cast<float*>(dst);
}

template <int MODE, int ILP>
void run(const char* name, const float* x, float* out, long long* cyc, int warps) {
  const int iters = 2000;
  cudaFuncSetAttribute(probe<MODE, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<MODE, ILP><<<148, 512, 200 * 1024>>>(x, out, cyc, iters, warps);
  probe<MODE, ILP><<<148, 512, 200 * 1024>>>(x, out, cyc, iters, warps);
  cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  // warps per sub-partition = warps / 4; float4-warps per sub-partition = warps / 4 * iters * ILP
  const double per = mx / ((double)(warps < 4 ? 1 : warps / 4) * iters * ILP);
  printf("%-44s ILP %d  warps %2d: %7.1f cycles per float4-warp and sub-partition  (%.2f elements/clk/SM)\n", name, ILP, warps, per, 4.0 * 128.0 / per);
}

int main() {
  float *x, *out; long long* cyc;
  cudaMalloc(&x, 4096 * 16); cudaMalloc(&out, 64); cudaMalloc(&cyc, 148 * 16 * 8);
  float h[16384];
  for (int i = 0; i < 16384; ++i) h[i] = (float)((i * 7919) % 2001 - 1000) * 0.004f;
  cudaMemcpy(x, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int warps : {4, 8, 12, 16}) {
    run<0, 1>("snake + split", x, out, cyc, warps);
    run<0, 2>("snake + split", x, out, cyc, warps);
    run<0, 3>("snake + split", x, out, cyc, warps);
    run<1, 3>("snake (range reduction) + split", x, out, cyc, warps);
    run<2, 3>("snake only (fp32 store)", x, out, cyc, warps);
    run<3, 3>("split only", x, out, cyc, warps);
    run<4, 3>("load + store only", x, out, cyc, warps);
    run<5, 3>("snake + truncation split", x, out, cyc, warps);
    run<6, 3>("scalar snake + split", x, out, cyc, warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
