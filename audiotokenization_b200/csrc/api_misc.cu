// Error plumbing, device query, layout transforms.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace bc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static Policy read_policy() {
  Policy p;
  const char* e;
  e = getenv("BC_TC_VARIANT");   p.tc_variant = e ? atoi(e) & 3 : 0;
  e = getenv("BC_TC_PERSIST");   p.tc_persist = (e && e[0] == '1') ? 1 : 0;
  e = getenv("BC_RU_GROUP");     p.ru_group = (e && e[0] == '0') ? 0 : 1;
  e = getenv("BC_RU_PERSIST");   p.ru_persist = (e && e[0] == '0') ? 0 : 1;
  e = getenv("BC_RU_PAIR");      p.ru_pair = (e && e[0] == '0') ? 0 : 1;
  e = getenv("BC_STREAM_PAIR");  p.stream_pair = (e && e[0] == '0') ? 0 : 1;
  e = getenv("BC_STREAM_TMA");   p.stream_tma = e ? atoi(e) & 3 : 1;     // 0 off, 1 TMA stores, 2 TMA stores + loads
  e = getenv("BC_LSTM_PINGPONG"); p.lstm_pingpong = (e && e[0] == '0') ? 0 : 1;
  e = getenv("BC_LSTM_COMPACT"); p.lstm_compact = (e && e[0] == '0') ? 0 : 1;  // compact h exchange image for batches below 128 rows
  e = getenv("BC_LSTM_PAIR");    p.lstm_pair = e ? atoi(e) & 3 : 0;            // 1: where one tile per CTA does not fit (measured: no gain over the ping-pong form), 2: whenever the tile count is even
  return p;
}
const Policy& policy() {
  static const Policy p = read_policy();   // thread-safe one-time initialisation
  return p;
}
}  // namespace bc

extern "C" int bc_policy(char* buf, size_t n) {
  if (!buf || n == 0) return bc::fail(BC_EINVAL, "bc_policy: null buffer");
  const bc::Policy& p = bc::policy();
#ifdef BC_TRACE
  const int trace = 1;
#else
  const int trace = 0;
#endif
  snprintf(buf, n, "tc_variant=%d tc_persist=%d ru_group=%d ru_persist=%d ru_pair=%d stream_pair=%d stream_tma=%d lstm_pingpong=%d lstm_pair=%d lstm_compact=%d trace_build=%d",
           p.tc_variant, p.tc_persist, p.ru_group, p.ru_persist, p.ru_pair, p.stream_pair, p.stream_tma, p.lstm_pingpong, p.lstm_pair, p.lstm_compact, trace);
  return BC_OK;
}

extern "C" int bc_abi_version(void) { return BC_ABI_VERSION; }
extern "C" const char* bc_last_error(void) { return bc::g_err; }

extern "C" int bc_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || dev < 0 || dev >= n) {
    cudaGetLastError();
    return bc::fail(BC_ENODEVICE, "no CUDA device %d (count=%d, %s)", dev, n, cudaGetErrorString(e));
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaGetDeviceProperties");
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = p.totalGlobalMem;
  if (p.major != 10) return bc::fail(BC_ENODEVICE, "device %d is sm_%d%d, this library is sm_100a only", dev, p.major, p.minor);
  return BC_OK;
}

// ---------------------------------------------------------------------------
// [B][R][Cc] -> [B][Cc][R] tiled transpose (32x32 tile, +1 padding)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int Cc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* xb = x + (size_t)b * R * Cc;
  float* yb = y + (size_t)b * R * Cc;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int r = r0 + ty + i, c = c0 + tx;
    if (r < R && c < Cc) tile[ty + i][tx] = xb[(size_t)r * Cc + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int c = c0 + ty + i, r = r0 + tx;
    if (r < R && c < Cc) yb[(size_t)c * R + r] = tile[tx][ty + i];
  }
}

static int launch_transpose(const float* x, float* y, int B, int R, int Cc, bc_stream_t s) {
  BC_REQUIRE(x && y, "transpose: null pointer");
  BC_REQUIRE(B > 0 && R > 0 && Cc > 0, "transpose: bad shape B=%d R=%d C=%d", B, R, Cc);
  BC_REQUIRE(B <= 65535, "transpose: B=%d > 65535", B);
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, B);
  BC_REQUIRE(grid.y <= 65535, "transpose: too many row tiles (%u)", grid.y);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, y, R, Cc);
  BC_LAUNCH_CHECK("transpose_kernel");
  return BC_OK;
}

extern "C" int bc_transpose_bct_to_btc(const float* x, float* y, int B, int C, int T, bc_stream_t s) {
  // rows = C, cols = T ; grid.y = C tiles (small), grid.x = T tiles
  return launch_transpose(x, y, B, C, T, s);
}

extern "C" int bc_transpose_btc_to_bct(const float* x, float* y, int B, int T, int C, bc_stream_t s) {
  // rows = T, cols = C: swap roles so the large dimension rides grid.x
  BC_REQUIRE(x && y, "transpose: null pointer");
  BC_REQUIRE(B > 0 && T > 0 && C > 0 && B <= 65535, "transpose: bad shape B=%d T=%d C=%d", B, T, C);
  if ((T + 31) / 32 <= 65535) return launch_transpose(x, y, B, T, C, s);
  // very long sequences: split the time axis into chunks that fit grid.y
  // (output rows are strided by T, so chunking needs the generic kernel below)
  return bc::fail(BC_EUNSUPPORTED, "transpose: T=%d too long for one launch", T);
}

__global__ void idx_to_i16_kernel(const int32_t* __restrict__ idx, int16_t* __restrict__ out, int n_q, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_q * N) return;
  int n = i / n_q, q = i - n * n_q;
  out[i] = (int16_t)idx[(size_t)q * N + n];
}

extern "C" int bc_indices_to_int16(const int32_t* idx, int16_t* out, int n_q, int N, bc_stream_t s) {
  BC_REQUIRE(idx && out && n_q > 0 && N > 0, "indices_to_int16: bad arguments");
  long long tot = (long long)n_q * N;
  BC_REQUIRE(tot < (1ll << 31), "indices_to_int16: too many indices");
  idx_to_i16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)s>>>(idx, out, n_q, N);
  BC_LAUNCH_CHECK("idx_to_i16_kernel");
  return BC_OK;
}

// ---------------------------------------------------------------------------
// Peer memory for the one ordered hand-off of the long-form path (the conv front end of ONE long
// recording is dealt out over the GPUs chunk by chunk; each owner then stores its frame-rate features straight into the
// LSTM owner's buffer over NVLink -- plain peer stores, not a collective, no NCCL).  One process per GPU, so the
// buffer crosses processes as a CUDA IPC handle.
// ---------------------------------------------------------------------------
extern "C" int bc_ipc_alloc(void** dev_ptr, size_t bytes) {
  BC_REQUIRE(dev_ptr && bytes > 0, "ipc_alloc: bad arguments");
  return bc::cuda_check(cudaMalloc(dev_ptr, bytes), "cudaMalloc(ipc)");
}
extern "C" int bc_ipc_free(void* dev_ptr) { return bc::cuda_check(cudaFree(dev_ptr), "cudaFree(ipc)"); }
extern "C" int bc_ipc_export(const void* dev_ptr, unsigned char* handle64) {
  BC_REQUIRE(dev_ptr && handle64, "ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));
  if (e != cudaSuccess) return bc::cuda_check(e, "cudaIpcGetMemHandle");
  memcpy(handle64, &h, 64);
  return BC_OK;
}
extern "C" int bc_ipc_open(const unsigned char* handle64, void** peer_ptr) {
  BC_REQUIRE(handle64 && peer_ptr, "ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return bc::cuda_check(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
extern "C" int bc_ipc_close(void* peer_ptr) { return bc::cuda_check(cudaIpcCloseMemHandle(peer_ptr), "cudaIpcCloseMemHandle"); }
// dst / src: any two device pointers valid in this process (local, or a peer buffer opened with bc_ipc_open)
extern "C" int bc_peer_copy(void* dst, const void* src, size_t bytes, bc_stream_t s) {
  BC_REQUIRE(dst && src, "peer_copy: null pointer");
  if (bytes == 0) return BC_OK;
  return bc::cuda_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)s), "cudaMemcpyAsync(peer)");
}
