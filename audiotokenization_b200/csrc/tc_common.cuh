// Device-side building blocks shared by the tcgen05 kernels (sm_100a): shared-memory matrix descriptors,
// MMA issue, mbarrier waits, bulk async copies, bf16 hi/lo splitting, TMEM loads.
// Bit layouts follow cute/arch/mma_sm100_desc.hpp (CUTLASS 4.x, vendored in the image) -- third-party
// documentation of the hardware formats, not the reference repository.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace bc {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int variant = 0) {
  // cute::UMMA::SmemDescriptor: start [0,14) | LBO [16,30) | SBO [32,46) | version=1 [46,48) | layout_type=0 (no swizzle)
  if (variant & 1) { uint32_t t = lbo_bytes; lbo_bytes = sbo_bytes; sbo_bytes = t; }
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  if (!(variant & 2)) d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// one lane of a converged warp (the surrounding control flow stays warp-uniform, so descriptor arithmetic
// lives in uniform registers and feeds UTCHMMA without per-MMA R2UR round trips)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Descriptor handling for issue loops: the 64-bit descriptor is (hi << 32) | lo with every per-MMA change
// confined to the 14-bit start-address field of `lo` (16-byte units), so an issue loop only adds small
// integers to `lo`.  `ACC` selects overwrite (first MMA of a tile) or accumulate at compile time.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }

template <bool ACC>
__device__ __forceinline__ void mma_bf16_raw(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                             uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}

// runtime accumulate flag (0 = overwrite the accumulator, else accumulate)
__device__ __forceinline__ void mma_bf16_raw_rt(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool ACC>
__device__ __forceinline__ void mma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                              uint32_t idesc) {
  if (elect_one()) asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}

// (A back-off variant that sleeps between polls was measured: the polls of waiting warps only use issue slots nobody
// else wants, while the extra wake-up latency costs 5-7 % on the narrow layers -- so every role polls.)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~1 ms
  // passes) instead of burning issue slots in a spin loop.  Bounded: a wrong descriptor or a broken
  // pipeline must surface as a launch failure, never as a hung GPU.
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    if (ok) return;
    if (it > 4000u) __trap();
  }
}

// 1-D bulk async copy global -> shared (TMA engine, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- packed fp32 pairs (FFMA2 / FADD2 / FMUL2, sm_100): one issue slot for two IEEE-identical fp32 operations.
// The activation math of these kernels is bound by instruction issue, not by the FMA pipe, so pairing halves its cost.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// bf16 hi halves of two floats as the two fp32 values they represent
__device__ __forceinline__ f32x2 bf16x2_as_f32x2(uint32_t h) { return pack2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u)); }

// 8 fp32 values -> the bf16 hi (and lo = v - hi) halves of one 16-byte K-major row chunk, in registers
template <int SPLIT>
__device__ __forceinline__ void split8(const float v[8], uint4& h, uint4& l) {
  h.x = pack_bf16x2(v[0], v[1]); h.y = pack_bf16x2(v[2], v[3]);
  h.z = pack_bf16x2(v[4], v[5]); h.w = pack_bf16x2(v[6], v[7]);
  if (SPLIT == 2) {
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
    uint32_t ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r0, r1;
      unpack2(sub2(pack2(v[2 * e], v[2 * e + 1]), bf16x2_as_f32x2(hh[e])), r0, r1);   // exact: v - bf16(v)
      ll[e] = pack_bf16x2(r0, r1);
    }
    l = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

template <int SPLIT>
__device__ __forceinline__ void split_store(const float v[8], uint8_t* dst, uint32_t lo_offset) {
  uint4 h;
  h.x = pack_bf16x2(v[0], v[1]); h.y = pack_bf16x2(v[2], v[3]);
  h.z = pack_bf16x2(v[4], v[5]); h.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = h;
  if (SPLIT == 2) {
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
    uint32_t ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r0, r1;
      unpack2(sub2(pack2(v[2 * e], v[2 * e + 1]), bf16x2_as_f32x2(hh[e])), r0, r1);   // exact: v - bf16(v)
      ll[e] = pack_bf16x2(r0, r1);
    }
    *reinterpret_cast<uint4*>(dst + lo_offset) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

// SnakeBeta with an explicit range reduction to [-pi/2, pi/2] followed by the SFU sine: sin^2 is
// pi-periodic, so the quadrant never matters.  |error| ~ 4e-7 absolute, independent of |x*a|.
__device__ __forceinline__ float snake_tc(float x, float a, float ib) {
  const float t = x * a;
  // round-to-nearest via the 1.5*2^23 trick (FMA pipe) instead of FRND (SFU pipe); |t/pi| < 2^22 always here
  const float n = __fadd_rn(__fmaf_rn(t, 0.318309886183790672f, 12582912.f), -12582912.f);
  float r = fmaf(n, -3.14159274101257324f, t);
  r = fmaf(n, 8.74227765734758577e-8f, r);   // pi - float(pi) = -8.742e-8
  const float s = __sinf(r);
  return fmaf(ib, s * s, x);
}

// the same arithmetic, operation for operation, on two elements per instruction (bit-identical results)
__device__ __forceinline__ f32x2 snake_tc2(f32x2 x, f32x2 a, f32x2 ib) {
  const f32x2 t = mul2(x, a);
  f32x2 n = fma2(t, pack2(0.318309886183790672f, 0.318309886183790672f), pack2(12582912.f, 12582912.f));
  n = add2(n, pack2(-12582912.f, -12582912.f));
  f32x2 r = fma2(n, pack2(-3.14159274101257324f, -3.14159274101257324f), t);
  r = fma2(n, pack2(8.74227765734758577e-8f, 8.74227765734758577e-8f), r);
  float r0, r1;
  unpack2(r, r0, r1);
  const f32x2 s = pack2(__sinf(r0), __sinf(r1));
  return fma2(ib, mul2(s, s), x);
}

// single-pass bf16 mode: the operand is rounded to 8 mantissa bits right after, so the SFU sine on the raw
// product is plenty (its error grows like |x*a| * 6e-8, three orders below the bf16 rounding).
__device__ __forceinline__ float snake_bf(float x, float a, float ib) {
  const float s = __sinf(x * a);
  return fmaf(ib, s * s, x);
}
__device__ __forceinline__ f32x2 snake_bf2(f32x2 x, f32x2 a, f32x2 ib) {
  float t0, t1;
  unpack2(mul2(x, a), t0, t1);
  const f32x2 s = pack2(__sinf(t0), __sinf(t1));
  return fma2(ib, mul2(s, s), x);
}

// SnakeBeta of the tensor-core prologues / MID stages.  BC_SNAKE_REDUCE = 1 keeps the explicit range reduction in split
// precision (absolute error of sin independent of |x*a|); 0 (default) feeds the SFU the raw product in both modes.  The SFU's
// own reduction works on the product rounded to turns, so its error grows like |x*a| * 1.2e-7 rad -- but y = x + sin^2/b then
// carries a RELATIVE error of at most 1.2e-7 * (a/b), because |x*a| large means |x| large: two orders below the 2^-17
// relative error of the hi/lo split that follows, at half the FMA-pipe work per element (4 instead of 8 operations; the
// producers and MID stages are bound by exactly this arithmetic).  The exact-fp32 kernels and the anti-aliased
// stencil keep the reduced form.
#ifndef BC_SNAKE_REDUCE
#define BC_SNAKE_REDUCE 0
#endif
template <int SPLIT>
__device__ __forceinline__ void snake8(float v[8], const float4& a0, const float4& a1, const float4& b0, const float4& b1) {
  const f32x2 aa[4] = {pack2(a0.x, a0.y), pack2(a0.z, a0.w), pack2(a1.x, a1.y), pack2(a1.z, a1.w)};
  const f32x2 bb[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const f32x2 x = pack2(v[2 * e], v[2 * e + 1]);
    const f32x2 y = (SPLIT == 2 && BC_SNAKE_REDUCE) ? snake_tc2(x, aa[e], bb[e]) : snake_bf2(x, aa[e], bb[e]);
    unpack2(y, v[2 * e], v[2 * e + 1]);
  }
}

// accumulator words r[0..7] + bias -> v[0..7], two elements per instruction
__device__ __forceinline__ void acc_bias8(const uint32_t* r, const float4& b0, const float4& b1, float v[8]) {
  unpack2(add2(pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack2(b0.x, b0.y)), v[0], v[1]);
  unpack2(add2(pack2(__uint_as_float(r[2]), __uint_as_float(r[3])), pack2(b0.z, b0.w)), v[2], v[3]);
  unpack2(add2(pack2(__uint_as_float(r[4]), __uint_as_float(r[5])), pack2(b1.x, b1.y)), v[4], v[5]);
  unpack2(add2(pack2(__uint_as_float(r[6]), __uint_as_float(r[7])), pack2(b1.z, b1.w)), v[6], v[7]);
}
// (r[0..3] + b) as a float4
__device__ __forceinline__ float4 acc_bias4(const uint32_t* r, const float4& b) {
  float4 v;
  unpack2(add2(pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack2(b.x, b.y)), v.x, v.y);
  unpack2(add2(pack2(__uint_as_float(r[2]), __uint_as_float(r[3])), pack2(b.z, b.w)), v.z, v.w);
  return v;
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  float4 v;
  unpack2(add2(pack2(a.x, a.y), pack2(b.x, b.y)), v.x, v.y);
  unpack2(add2(pack2(a.z, a.w), pack2(b.z, b.w)), v.z, v.w);
  return v;
}

// four elements (one 16-byte load): the producers of the streamed-weight kernel
template <int SPLIT>
__device__ __forceinline__ void snake4(float4& v, const float4& a, const float4& b) {
  const f32x2 y0 = (SPLIT == 2 && BC_SNAKE_REDUCE) ? snake_tc2(pack2(v.x, v.y), pack2(a.x, a.y), pack2(b.x, b.y)) : snake_bf2(pack2(v.x, v.y), pack2(a.x, a.y), pack2(b.x, b.y));
  const f32x2 y1 = (SPLIT == 2 && BC_SNAKE_REDUCE) ? snake_tc2(pack2(v.z, v.w), pack2(a.z, a.w), pack2(b.z, b.w)) : snake_bf2(pack2(v.z, v.w), pack2(a.z, a.w), pack2(b.z, b.w));
  unpack2(y0, v.x, v.y);
  unpack2(y1, v.z, v.w);
}

// 32 accumulator columns of this thread's TMEM lane in one instruction
__device__ __forceinline__ void tmem_load32(uint32_t taddr, uint32_t r[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// accumulator columns [c0, c0 + 8*n8) of this thread's TMEM lane -> r[]   (n8 <= 4, warp-uniform)
__device__ __forceinline__ void tmem_load(uint32_t taddr, int n8, uint32_t r[32]) {
  if (n8 == 4) { tmem_load32(taddr, r); return; }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < n8) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[8 * j + 0]), "=r"(r[8 * j + 1]), "=r"(r[8 * j + 2]), "=r"(r[8 * j + 3]), "=r"(r[8 * j + 4]),
                     "=r"(r[8 * j + 5]), "=r"(r[8 * j + 6]), "=r"(r[8 * j + 7])
                   : "r"(taddr + 8u * j));
    }
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_notx(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG): 3-D tiles {channel, row, item} of a channels-last
// activation tensor.  Rows / items outside the tensor read as zeros and are not written, so tile edges need no code.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) { asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128
__host__ __device__ inline uint32_t idesc_bf16_m128(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA pairs (tcgen05 cta_group::2): one MMA of M = 256 across a cluster of two CTAs, each SM contributing the 128 rows of
// its own tile as the A operand and HALF of the B operand's rows.  Protocol shared by ru_pair.cu and the pair form of
// conv_stream.cu: producers of BOTH CTAs arrive on the LEADER's barrier (rank 0), the leader's MMA thread signals
// consumers in both CTAs with a multicast commit; every barrier sits at the same shared-memory offset in both CTAs.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Memory-model scope of the cross-CTA hand-offs.  What travels between the CTAs of a pair is SHARED-MEMORY data for the
// tensor core (activation slabs, re-quantised tiles: generic-proxy stores made visible to the async proxy with
// fence.proxy.async before the arrive) and "accumulator drained" notifications -- never global memory.  Cluster-scope
// release / acquire compiles to MEMBAR + CCTL.IVALL (an L1 invalidation per wait: `cuobjdump -sass` showed 20 CCTL in
// the first ru_pair build), which the loaders' L1-cached activation reads pay for.  The default semantics
// (release / acquire at CTA scope -- what CUTLASS' ClusterBarrier::arrive(cta_id) emits for its 2-SM kernels) order the
// shared-memory stores before the arrive, which is all these hand-offs need.  -DBC_PAIR_STRONG=1 restores cluster scope.
#ifndef BC_PAIR_STRONG
#define BC_PAIR_STRONG 0
#endif
// arrive on the LEADER CTA's copy of a barrier (rank 0 of the pair), from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t local_bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(0u));
#if BC_PAIR_STRONG
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#else
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#endif
}
// wait on a barrier whose arrivals come from both CTAs
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
#if BC_PAIR_STRONG
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    if (ok) return;
    if (it > 4000u) __trap();
  }
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void mma2_bf16_raw(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                              uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "n"(ACC ? 1 : 0)
      : "memory");
}
// runtime accumulate flag (0 = overwrite the accumulator, else accumulate)
__device__ __forceinline__ void mma2_bf16_rt(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 256 across the CTA pair
__host__ __device__ inline uint32_t idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}


}  // namespace tc
}  // namespace bc
