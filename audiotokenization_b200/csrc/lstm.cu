// Uni-directional LSTM layer, recurrent part, as ONE persistent cooperative kernel.
//
// Replaces the time loop inside nn.LSTM (ResLSTM, vq/module.py:143-167).  The input
// projection W_ih x_t + b for all t is a dense K=1 contraction done beforehand by
// bc_conv1d_fwd; only W_hh h_{t-1} is sequential.
//
// Decomposition: CTA j owns U = 4 hidden units (16 gate rows of W_hh, kept resident in
// shared memory for the whole sequence: 16 x H floats).  Every step each CTA
//   1. copies h_{t-1} for a tile of 32 batch items from global (L2) into smem, [j][b];
//   2. 8 warps split the H-long reduction; lane = (unit u, batch quad) holds a
//      4 gates x 4 batch register tile -> 2 LDS.128 per 16 FFMA;
//   3. partial sums meet in smem; thread (u, b) adds the pre-activation, applies the
//      gates, keeps c in a register and publishes h_t (global, [parity][H][B]);
//   4. grid-wide barrier (cooperative groups).
// h is double-buffered by step parity, so a CTA that runs ahead never overwrites data a
// slower CTA still reads.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int U = 4;          // hidden units per CTA
constexpr int R = 4 * U;      // gate rows per CTA
constexpr int BT = 32;        // batch tile
constexpr int NWARPS = 8;
constexpr int NT = NWARPS * 32;
constexpr int MAX_BTILES = 8; // c state lives in registers: B <= 256 per launch

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// packed W_hh: [H/U CTAs][H (j)][R] with r = u*4 + gate
__global__ void __launch_bounds__(NT) lstm_rec_kernel(const float* __restrict__ pre, const float* __restrict__ wpk,
                                                      const float* __restrict__ skip, float* __restrict__ y,
                                                      float* __restrict__ hbuf, int B, int T, int H, int Bpad) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                       // [H][R]
  float* hs = ws + (size_t)H * R;         // [H][BT]
  float* red = hs + (size_t)H * BT;       // [NWARPS][R][BT]
  cg::grid_group grid = cg::this_grid();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = blockIdx.x * U;
  {  // resident weights
    const float4* src = reinterpret_cast<const float4*>(wpk + (size_t)blockIdx.x * H * R);
    float4* dst = reinterpret_cast<float4*>(ws);
    for (int i = tid; i < H * R / 4; i += NT) dst[i] = __ldg(src + i);
  }
  const int nbt = (B + BT - 1) / BT;
  const int ru = lane >> 3;   // unit handled in the dot-product phase
  const int bq = lane & 7;    // batch quad
  const int jchunk = H / NWARPS;
  const int j0 = warp * jchunk;

  // finalisation role: thread (fu, fb) for tid < U*BT
  const int fu = tid / BT, fb = tid % BT;
  float c_state[MAX_BTILES];
#pragma unroll
  for (int i = 0; i < MAX_BTILES; ++i) c_state[i] = 0.f;

  __syncthreads();

  for (int t = 0; t < T; ++t) {
    const float* hprev = hbuf + (size_t)((t + 1) & 1) * H * Bpad;  // written at step t-1 (zeros for t = 0)
    float* hcur = hbuf + (size_t)(t & 1) * H * Bpad;
#pragma unroll 1
    for (int bt = 0; bt < nbt; ++bt) {
      const int b0 = bt * BT;
      // early issue of the pre-activations this thread will need at the end
      float pg[4] = {0.f, 0.f, 0.f, 0.f};
      const bool fin = (tid < U * BT) && (b0 + fb < B);
      if (fin) {
        const float* pp = pre + ((size_t)(b0 + fb) * T + t) * 4 * H + u0 + fu;
#pragma unroll
        for (int g = 0; g < 4; ++g) pg[g] = __ldcs(pp + (size_t)g * H);
      }
      // 1. h_{t-1} tile -> smem (L2 reads, bypass L1: other SMs wrote it)
      {
        const float4* src = reinterpret_cast<const float4*>(hprev + b0);
        for (int i = tid; i < H * (BT / 4); i += NT) {
          const int j = i >> 3, q = i & 7;
          const float4 v = __ldcg(src + ((size_t)j * Bpad) / 4 + q);
          *reinterpret_cast<float4*>(hs + j * BT + q * 4) = v;
        }
      }
      __syncthreads();
      // 2. partial dot products over this warp's slice of j
      float acc[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[g][e] = 0.f;
      const float* wp = ws + (size_t)j0 * R + ru * 4;
      const float* hp = hs + (size_t)j0 * BT + bq * 4;
#pragma unroll 4
      for (int j = 0; j < jchunk; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(wp + j * R);
        const float4 h4 = *reinterpret_cast<const float4*>(hp + j * BT);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
        const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[g][e] = fmaf(wv[g], hv[e], acc[g][e]);
      }
      // 3. cross-warp reduction through smem: red[warp][r = ru*4+g][b = bq*4+e]
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<float4*>(red + ((size_t)warp * R + ru * 4 + g) * BT + bq * 4) =
            make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
      __syncthreads();
      if (tid < U * BT) {
        float gate[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float sum = 0.f;
#pragma unroll
          for (int w = 0; w < NWARPS; ++w) sum += red[((size_t)w * R + fu * 4 + g) * BT + fb];
          gate[g] = sum + pg[g];
        }
        if (fin) {
          const float ig = sigmoid_f(gate[0]), fg = sigmoid_f(gate[1]);
          const float gg = tanhf(gate[2]), og = sigmoid_f(gate[3]);
          float c = c_state[0];
#pragma unroll
          for (int i = 1; i < MAX_BTILES; ++i) c = (bt == i) ? c_state[i] : c;
          c = fmaf(fg, c, ig * gg);
#pragma unroll
          for (int i = 0; i < MAX_BTILES; ++i) c_state[i] = (bt == i) ? c : c_state[i];
          const float h = og * tanhf(c);
          __stcg(hcur + (size_t)(u0 + fu) * Bpad + b0 + fb, h);
          const size_t o = ((size_t)(b0 + fb) * T + t) * H + u0 + fu;
          y[o] = skip ? h + __ldcs(skip + o) : h;
        }
      }
      __syncthreads();  // hs / red reused by the next batch tile
    }
    grid.sync();
  }
}


// Wide layers (H = 1536 in the original BigCodec config: 37.7 MB of fp32 W_hh, more than the shared memory of
// the whole chip): same arithmetic and the same packed weight image, but the gate rows are streamed from L2
// every step, a CTA walks several unit groups, and the cell state lives in the workspace instead of registers.
__global__ void __launch_bounds__(NT) lstm_rec_stream_kernel(const float* __restrict__ pre, const float* __restrict__ wpk,
                                                             const float* __restrict__ skip, float* __restrict__ y,
                                                             float* __restrict__ hbuf, float* __restrict__ cbuf,
                                                             int B, int T, int H, int Bpad, int ngroups) {
  extern __shared__ __align__(16) float smem[];
  float* hs = smem;                       // [H][BT]
  float* red = hs + (size_t)H * BT;       // [NWARPS][R][BT]
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbt = (B + BT - 1) / BT;
  const int ru = lane >> 3, bq = lane & 7;
  const int jchunk = H / NWARPS;
  const int j0 = warp * jchunk;
  const int fu = tid / BT, fb = tid % BT;

  for (int t = 0; t < T; ++t) {
    const float* hprev = hbuf + (size_t)((t + 1) & 1) * H * Bpad;
    float* hcur = hbuf + (size_t)(t & 1) * H * Bpad;
#pragma unroll 1
    for (int bt = 0; bt < nbt; ++bt) {
      const int b0 = bt * BT;
      {
        const float4* src = reinterpret_cast<const float4*>(hprev + b0);
        for (int i = tid; i < H * (BT / 4); i += NT) {
          const int j = i >> 3, q = i & 7;
          *reinterpret_cast<float4*>(hs + j * BT + q * 4) = __ldcg(src + ((size_t)j * Bpad) / 4 + q);
        }
      }
      __syncthreads();
#pragma unroll 1
      for (int ug = blockIdx.x; ug < ngroups; ug += gridDim.x) {
        const int u0 = ug * U;
        float pg[4] = {0.f, 0.f, 0.f, 0.f};
        const bool fin = (tid < U * BT) && (b0 + fb < B);
        if (fin) {
          const float* pp = pre + ((size_t)(b0 + fb) * T + t) * 4 * H + u0 + fu;
#pragma unroll
          for (int g = 0; g < 4; ++g) pg[g] = __ldcs(pp + (size_t)g * H);
        }
        float acc[4][4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[g][e] = 0.f;
        const float* wp = wpk + ((size_t)ug * H + j0) * R + ru * 4;
        const float* hp = hs + (size_t)j0 * BT + bq * 4;
#pragma unroll 4
        for (int j = 0; j < jchunk; ++j) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * R));
          const float4 h4 = *reinterpret_cast<const float4*>(hp + j * BT);
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
          const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[g][e] = fmaf(wv[g], hv[e], acc[g][e]);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<float4*>(red + ((size_t)warp * R + ru * 4 + g) * BT + bq * 4) =
              make_float4(acc[g][0], acc[g][1], acc[g][2], acc[g][3]);
        __syncthreads();
        if (fin) {
          float gate[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) sum += red[((size_t)w * R + fu * 4 + g) * BT + fb];
            gate[g] = sum + pg[g];
          }
          const float ig = sigmoid_f(gate[0]), fg = sigmoid_f(gate[1]);
          const float gg = tanhf(gate[2]), og = sigmoid_f(gate[3]);
          float* cp = cbuf + (size_t)(u0 + fu) * Bpad + b0 + fb;   // only this thread ever touches it
          const float c = fmaf(fg, *cp, ig * gg);
          *cp = c;
          const float h = og * tanhf(c);
          __stcg(hcur + (size_t)(u0 + fu) * Bpad + b0 + fb, h);
          const size_t o = ((size_t)(b0 + fb) * T + t) * H + u0 + fu;
          y[o] = skip ? h + __ldcs(skip + o) : h;
        }
        __syncthreads();  // red reused by the next unit group, hs by the next batch tile
      }
    }
    grid.sync();
  }
}

}  // namespace

extern "C" size_t bc_lstm_workspace_bytes(int B, int H) {
  if (B <= 0 || H <= 0) return 0;
  const size_t bpad = ((size_t)B + BT - 1) / BT * BT;
  return 3 * (size_t)H * bpad * sizeof(float);   // h (two step parities) + c (streamed-weight variant)
}

extern "C" size_t bc_lstm_packed_whh_floats(int H) { return H > 0 ? (size_t)4 * H * H : 0; }

extern "C" int bc_lstm_pack_whh(const float* w_hh, float* packed, int H) {
  BC_REQUIRE(w_hh && packed && H > 0 && H % U == 0, "lstm_pack_whh: bad arguments (H=%d)", H);
  // packed[cta][j][u*4+g] = w_hh[g*H + cta*U + u][j]
  const int ncta = H / U;
  for (int cta = 0; cta < ncta; ++cta)
    for (int j = 0; j < H; ++j)
      for (int u = 0; u < U; ++u)
        for (int g = 0; g < 4; ++g)
          packed[((size_t)cta * H + j) * R + u * 4 + g] = w_hh[((size_t)g * H + cta * U + u) * H + j];
  return BC_OK;
}

extern "C" int bc_lstm_recurrent_fwd(const float* pre, const float* w_hh_packed, const float* skip, float* y,
                                     void* workspace, int B, int T, int H, bc_stream_t s) {
  BC_REQUIRE(pre && w_hh_packed && y && workspace, "lstm: null pointer");
  BC_REQUIRE(B > 0 && T > 0 && H > 0, "lstm: bad shape B=%d T=%d H=%d", B, T, H);
  BC_REQUIRE(H % (NWARPS * 4) == 0, "lstm: H=%d must be a multiple of %d", H, NWARPS * 4);
  if (B > BT * MAX_BTILES) return bc::fail(BC_EUNSUPPORTED, "lstm: B=%d > %d per launch (split the batch)", B, BT * MAX_BTILES);
  BC_REQUIRE(bc::aligned16(w_hh_packed) && bc::aligned16(workspace), "lstm: packed weights / workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
  const int ncta = H / U;
  const int bpad = (B + BT - 1) / BT * BT;
  int dev = 0, sms = 0, occ = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) return bc::fail(BC_ENODEVICE, "lstm: device does not support cooperative launch");
  cudaError_t e = cudaMemsetAsync(workspace, 0, bc_lstm_workspace_bytes(B, H), st);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaMemsetAsync(lstm)");
  float* hbuf = reinterpret_cast<float*>(workspace);
  float* cbuf = hbuf + 2 * (size_t)H * bpad;
  int Bpad = bpad;
  // resident-weight kernel when every unit group gets its own co-resident CTA
  const size_t smem = ((size_t)H * R + (size_t)H * BT + (size_t)NWARPS * R * BT) * sizeof(float);
  bool resident = smem <= 227 * 1024;
  if (resident) {
    e = cudaFuncSetAttribute(lstm_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(lstm)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lstm_rec_kernel, NT, smem);
    if (e != cudaSuccess) return bc::cuda_check(e, "occupancy(lstm)");
    resident = occ * sms >= ncta;
  }
  if (resident) {
    void* args[] = {(void*)&pre, (void*)&w_hh_packed, (void*)&skip, (void*)&y, (void*)&hbuf, (void*)&B, (void*)&T, (void*)&H, (void*)&Bpad};
    e = cudaLaunchCooperativeKernel((void*)lstm_rec_kernel, dim3(ncta), dim3(NT), args, smem, st);
    if (e != cudaSuccess) return bc::cuda_check(e, "cudaLaunchCooperativeKernel(lstm)");
    return BC_OK;
  }
  const size_t smem_s = ((size_t)H * BT + (size_t)NWARPS * R * BT) * sizeof(float);
  if (smem_s > 227 * 1024)
    return bc::fail(BC_EUNSUPPORTED, "lstm: H=%d needs %zu B of shared memory per CTA (max 232448)", H, smem_s);
  e = cudaFuncSetAttribute(lstm_rec_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(lstm, streamed weights)");
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lstm_rec_stream_kernel, NT, smem_s);
  if (e != cudaSuccess) return bc::cuda_check(e, "occupancy(lstm, streamed weights)");
  if (occ < 1) return bc::fail(BC_EUNSUPPORTED, "lstm: H=%d does not fit one CTA per SM", H);
  int ngroups = ncta;
  const int grid = ngroups < occ * sms ? ngroups : occ * sms;
  void* args[] = {(void*)&pre, (void*)&w_hh_packed, (void*)&skip, (void*)&y, (void*)&hbuf, (void*)&cbuf,
                  (void*)&B, (void*)&T, (void*)&H, (void*)&Bpad, (void*)&ngroups};
  e = cudaLaunchCooperativeKernel((void*)lstm_rec_stream_kernel, dim3(grid), dim3(NT), args, smem_s, st);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaLaunchCooperativeKernel(lstm, streamed weights)");
  return BC_OK;
}
