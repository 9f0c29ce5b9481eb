// Shared helpers for the BigCodec sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/bigcodec_b200.h"

namespace bc {

void set_error(const char* fmt, ...);

// Kernel-selection knobs, read from the environment ONCE per process (first use) and reported by bc_policy():
//   BC_TC_VARIANT (descriptor variant of the per-tile kernel, bring-up only), BC_TC_PERSIST=1, BC_RU_GROUP=0,
//   BC_RU_PERSIST=0, BC_RU_PAIR=0, BC_STREAM_PAIR=0, BC_LSTM_PINGPONG=0.  None of them changes results; experiments that do are compiled in only with -DBC_TRACE.
struct Policy {
  int tc_variant, tc_persist, ru_group, ru_persist, ru_pair, lstm_pingpong, lstm_pair, lstm_compact, stream_pair, stream_tma;
};
const Policy& policy();

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return code;
}

inline int cuda_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return BC_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

#define BC_REQUIRE(cond, ...)                          \
  do {                                                 \
    if (!(cond)) return bc::fail(BC_EINVAL, __VA_ARGS__); \
  } while (0)

#define BC_LAUNCH_CHECK(name)                                      \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return bc::cuda_check(e__, name);      \
  } while (0)

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float snake_f32(float x, float a, float ib) {
  float s = sinf(x * a);
  return fmaf(ib, s * s, x);
}

// x + ib * sin(x*a)^2 evaluated as the reference does: x + (ib * (s*s)), no fma contraction
__device__ __forceinline__ float snake_ref(float x, float a, float ib) {
  float s = sinf(__fmul_rn(x, a));
  return __fadd_rn(x, __fmul_rn(ib, __fmul_rn(s, s)));
}

}  // namespace bc
