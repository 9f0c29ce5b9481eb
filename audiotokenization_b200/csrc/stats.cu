// Codebook usage statistics on the device: the step right after the indices in both evaluation flows of the
// reference (CodebookPerplexity / CodebookUtilization, lightning_module.py:26-73; calculate_perplexity,
// inference_full.py:570-604).  One pass over the indices builds the usage histogram; entropy, perplexity and
// utilisation are a K-element reduction of it.
#include "common.cuh"

namespace {

constexpr int H_THREADS = 256;
constexpr int H_SMEM_CODES = 8192;   // codebooks up to this size are counted in shared memory first

// counts[c] += #{n : idx[n] == c}; indices outside [0, Kc) are counted into *bad and otherwise ignored
__global__ void __launch_bounds__(H_THREADS) code_histogram_kernel(const int32_t* __restrict__ idx, long long N, int Kc,
                                                                   unsigned long long* __restrict__ counts,
                                                                   int* __restrict__ bad, int use_smem) {
  extern __shared__ unsigned int sh[];
  if (use_smem) {
    for (int i = threadIdx.x; i < Kc; i += H_THREADS) sh[i] = 0u;
    __syncthreads();
  }
  int nbad = 0;
  const long long stride = (long long)gridDim.x * H_THREADS;
  for (long long n = (long long)blockIdx.x * H_THREADS + threadIdx.x; n < N; n += stride) {
    const int c = __ldg(idx + n);
    if (c < 0 || c >= Kc) { ++nbad; continue; }
    if (use_smem) atomicAdd(&sh[c], 1u);
    else atomicAdd(&counts[c], 1ull);
  }
  if (nbad && bad) atomicAdd(bad, nbad);
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < Kc; i += H_THREADS) {
      const unsigned int v = sh[i];
      if (v) atomicAdd(&counts[i], (unsigned long long)v);
    }
  }
}

// out[0] = entropy (nats) of counts / total over the used codes, out[1] = number of used codes, out[2] = total
__global__ void __launch_bounds__(H_THREADS) code_entropy_kernel(const unsigned long long* __restrict__ counts, int Kc,
                                                                 double* __restrict__ out) {
  __shared__ double red[H_THREADS];
  __shared__ double s_total;
  double tot = 0.0;
  for (int i = threadIdx.x; i < Kc; i += H_THREADS) tot += (double)counts[i];
  red[threadIdx.x] = tot;
  __syncthreads();
  for (int o = H_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_total = red[0];
  __syncthreads();
  const double total = s_total;
  double ent = 0.0, used = 0.0;
  if (total > 0.0) {
    for (int i = threadIdx.x; i < Kc; i += H_THREADS) {
      const double c = (double)counts[i];
      if (c > 0.0) {
        const double pr = c / total;
        ent -= pr * log(pr);
        used += 1.0;
      }
    }
  }
  __syncthreads();
  red[threadIdx.x] = ent;
  __syncthreads();
  for (int o = H_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double e = red[0];
  __syncthreads();
  red[threadIdx.x] = used;
  __syncthreads();
  for (int o = H_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = e;
    out[1] = red[0];
    out[2] = total;
  }
}

}  // namespace

extern "C" int bc_code_histogram(const int32_t* idx, long long N, int Kc, unsigned long long* counts, int* bad_count,
                                 bc_stream_t s) {
  BC_REQUIRE(idx && counts, "code_histogram: null pointer");
  BC_REQUIRE(N >= 0 && Kc > 0, "code_histogram: bad shape N=%lld Kc=%d", N, Kc);
  if (N == 0) return BC_OK;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    return bc::fail(BC_ENODEVICE, "code_histogram: no CUDA device");
  }
  const int use_smem = Kc <= H_SMEM_CODES ? 1 : 0;
  long long want = (N + H_THREADS * 8 - 1) / (H_THREADS * 8);
  if (want > 4ll * sms) want = 4ll * sms;
  if (want < 1) want = 1;
  code_histogram_kernel<<<(unsigned)want, H_THREADS, use_smem ? (size_t)Kc * sizeof(unsigned int) : 0, (cudaStream_t)s>>>(
      idx, N, Kc, counts, bad_count, use_smem);
  BC_LAUNCH_CHECK("code_histogram_kernel");
  return BC_OK;
}

extern "C" int bc_code_entropy(const unsigned long long* counts, int Kc, double* out3, bc_stream_t s) {
  BC_REQUIRE(counts && out3, "code_entropy: null pointer");
  BC_REQUIRE(Kc > 0, "code_entropy: bad codebook size %d", Kc);
  code_entropy_kernel<<<1, H_THREADS, 0, (cudaStream_t)s>>>(counts, Kc, out3);
  BC_LAUNCH_CHECK("code_entropy_kernel");
  return BC_OK;
}
