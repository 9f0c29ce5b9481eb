// Persistent, warp-specialised tcgen05 convolution / fused ResidualUnit with STREAMED weights, sm_100a.
//
// For the wide layers (C >= 128: ResidualUnits of the low-rate blocks, the strided down-sampling convs,
// the LSTM input projection, the final conv) the weight set does not fit in shared memory, so it flows
// through a ring of (tap, 16-channel-group) units fed by bulk async copies from L2 while the CTA walks its
// share of the (n-tile, item, 128-step) tiles.  Every stage runs on its own warps and overlaps the others:
//
//   warps 0-7    PRODUCE  x (fp32, HBM) -> SnakeBeta -> bf16 hi[/lo] -> K-major slab of ONE 16-channel group
//                         (two teams of 4 warps take alternate groups: one team's HBM latency hides behind
//                         the other team's activation math)                                      (a_full)
//   warp  20     WEIGHTS  cp.async.bulk of the next weight unit into the B ring                  (b_full)
//   warp  21     MMA      K-tap conv: groups x taps [x3] tcgen05.mma into acc1 (TMEM)            (acc1_full)
//                         fused: 1x1 conv on the re-quantised tile into acc2                     (acc2_full)
//   warps 8-15   MID      (fused) acc1 -> +b7 -> snake2 -> bf16 hi[/lo] -> 64-channel A2 chunks  (a2_full)
//   warps 16-19  STORE    acc -> +bias (+residual) -> y (fp32, HBM) through a per-warp transposing stage
//
// Hand-offs are mbarriers; smem slots and TMEM accumulators are released by tcgen05.commit.  With two
// accumulator stages (N <= 128) the MMA warp issues the K-tap conv of tile i+1 before the 1x1 conv of tile i,
// so the tensor core never waits for the MID stage.
//
// Weight image (host-packed, pack_stream_weight):  [n_tile][group][tap][split][2 k-planes][N][8] bf16.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>
#include <cuda.h>      // CUtensorMap + enums only; the encoder comes from cudaGetDriverEntryPoint (no libcuda link)

namespace {
using namespace bc::tc;

constexpr int BM = 128;
constexpr int PROD_WARPS = 8;
// Warp ids grow with how latency-critical the role is: the SM's warp arbiter favours the higher warp id among
// eligible warps, and a delayed MMA-issue or weight-copy instruction idles the tensor core, while the
// activation math of the producers has slack.
#ifndef BC_STREAM_MID_WARPS
#define BC_STREAM_MID_WARPS 8
#endif
#ifndef BC_STREAM_PLAIN_PROD
#define BC_STREAM_PLAIN_PROD 12
#endif
#ifndef BC_STREAM_REG_CTRL
#define BC_STREAM_REG_CTRL 56
#endif
#ifndef BC_STREAM_REG_PROD_F     // fused kernel: registers per producer / MID thread after the rebalance (launch: 80)
#define BC_STREAM_REG_PROD_F 80
#endif
#ifndef BC_STREAM_REG_MID_F
#define BC_STREAM_REG_MID_F 80
#endif
#ifndef BC_STREAM_REG_PROD_P     // plain conv: registers per producer thread (launch: 96)
#define BC_STREAM_REG_PROD_P 96
#endif
// Two issue-path experiments, both measured SLOWER in the kernel although they win in the stand-alone issue probe
// (scripts/probes/mma_probe.cu modes 136/138: the weight ring is only ~2 units deep in time, and releasing a slot
// one tap later / holding a half-issued unit while waiting for the next one costs more than the hidden latency):
#ifndef BC_STREAM_DEFER_COMMITS   // 1: a unit's commits are issued behind the first tap of the next unit
#define BC_STREAM_DEFER_COMMITS 0
#endif
#ifndef BC_STREAM_PREFETCH_WAITS  // 1: the next unit's barriers are waited for behind the first tap of the current one
#define BC_STREAM_PREFETCH_WAITS 0
#endif
#ifndef BC_STREAM_PLAIN_PB
#define BC_STREAM_PLAIN_PB 6
#endif
constexpr int MID_WARPS = BC_STREAM_MID_WARPS;   // 8: two per TMEM lane quarter (one 32-column half each); 4: one per quarter, both halves
constexpr int MID_HALVES = 8 / MID_WARPS;
constexpr int EPI_WARPS = 4;              // 8 (two per TMEM lane quarter, alternate 32-column blocks) measured slower: the
                                          // register cap of the larger CTA (72) costs every role more than the stores gain
// Role layout of one CTA.  The fused ResidualUnit needs the MID stage (22 warps, 80 registers); a plain conv drops
// it and spends part of the freed register file on more producers and on a larger register cap (18 warps, 96
// registers: the tile-loop state of every role stays out of local memory).
template <bool FUSE> struct Roles {
  static constexpr int PROD = FUSE ? PROD_WARPS : BC_STREAM_PLAIN_PROD;   // multiple of 4 (TMEM lane quarter = warp % 4)
  static constexpr int MID0 = PROD;
  static constexpr int MID = FUSE ? MID_WARPS : 0;
  static constexpr int EPI0 = MID0 + MID;
  static constexpr int LOAD = EPI0 + EPI_WARPS;
  static constexpr int MMA = LOAD + 1;
  static constexpr int WARPS = (MMA + 1 + 3) / 4 * 4;                    // whole warpgroups: setmaxnreg works on groups of 4 warps
  static constexpr int THREADS = WARPS * 32;
  // register budget moved between warpgroups after launch: the two control warps (+2 idle) hand most of theirs
  // to the store warps, whose 32-column accumulator block + residual prefetch do not fit the launch-time cap
  static constexpr int REG_LAUNCH = FUSE ? 80 : 96;                       // what ptxas picks for THREADS
  static constexpr int REG_CTRL = BC_STREAM_REG_CTRL;
  static constexpr int REG_PROD = FUSE ? BC_STREAM_REG_PROD_F : BC_STREAM_REG_PROD_P;
  static constexpr int REG_MID = BC_STREAM_REG_MID_F;
  static constexpr int REG_STORE = (WARPS / 4) * REG_LAUNCH - (PROD / 4) * REG_PROD - (MID / 4) * REG_MID - REG_CTRL;
  static_assert(REG_STORE % 8 == 0 && REG_STORE >= REG_LAUNCH && REG_STORE <= 256, "store warpgroup register budget");
  static_assert(PROD % 4 == 0 && MID % 4 == 0, "roles must fill whole warpgroups");
  static constexpr int PB = FUSE ? 6 : BC_STREAM_PLAIN_PB;              // 16-byte loads a producer thread keeps in flight
};
constexpr uint32_t A2_PLANE = BM * 16u;  // one 8-channel plane of the re-quantised tile
constexpr int A2_CH = 64;              // channels per A2 chunk
constexpr int EPI_LD = 36;               // staging row stride in floats (32 + 4: conflict-free 16-byte accesses both ways)
constexpr int MAX_TPU = 8;             // taps per weight unit (issue block is unrolled this far)
constexpr uint32_t UNIT_MAX_BYTES = 32768;

enum { B_A_FULL = 0, B_A_EMPTY = 4, B_B_FULL = 8, B_B_EMPTY = 16, B_ACC1_FULL = 24, B_ACC1_EMPTY = 26,
       B_A2_FULL = 28, B_A2_EMPTY = 30, B_ACC2_FULL = 32, B_ACC2_EMPTY = 34, B_B_READY = 36, B_X_FULL = 44, B_X_EMPTY = 52,
       N_BARS = 60 };
constexpr int MAX_NX = 8;
#ifndef BC_STREAM_XGPB
#define BC_STREAM_XGPB 2
#endif
#ifndef BC_STREAM_XWANT      // bytes of x the ring should hold in flight
#define BC_STREAM_XWANT 49152u
#endif
constexpr uint32_t XT_STAGE_BYTES = 8192;   // TMA form: two swizzled [32 rows][128 B] store tiles per store warp

struct SParams {
  const float* x;
  float* y;
  const float* res;
  const uint8_t* w7;
  const uint8_t* w1;
  const float* bias;
  const float* bias2;
  const float* sa1;
  const float* sib1;
  const float* sa2;
  const float* sib2;
  int B, T_in, T_out, C_in, C_out, K, stride, dil, pad_left, flags;
  int nt_shift;     // n-tiles >= nt_shift read their taps one row later (pad_left - 1): the q = 1 phases of a transposed conv
  int N, groups, tpu, upg, gpu1;
  int slab_rows, rpp, NA, NB, acc_stages, acc_stride, epi_warps;
  int NT;           // producer teams (NA % NT == 0): slot s belongs to team s % NT
  uint32_t a_stage, unit_bytes, tap_bytes, plane_bytes;
  int NX, x_rows, x_nbox, x_gpb;   // TMA form: ring of NX fp32 units [x_rows][16 * x_gpb channels], x_nbox units per (tile, x_gpb groups)
  uint32_t x_unit, x_bytes;        // unit stride in the ring / bytes one box delivers
  int tiles_per_item, tiles_per_nt, total_tiles, n_tiles;
  int pairs_per_nt, total_pairs;   // pair form: tile pairs (2q, 2q+1) inside one n-tile; an odd count leaves rank 1 a phantom tile
  uint32_t idesc;
  int tmem_cols;
  long long* trace;   // debug: [tile_it < 64][16] clock64 stamps / wait totals of CTA 0 (NULL = off)
#ifdef BC_TRACE
  int dbg_skip;     // timing experiment bits: 1 producers skip loads+math, 2 MID skips math, 4 STORE skips global traffic
  int dbg_bshift;   // timing experiment: copy only 1/2^n of every weight unit (results are wrong)
#endif
};
// The timing experiments above produce WRONG results by design; they exist only in -DBC_TRACE builds
// (BC_TRACE=1 python -m audiotokenization_b200.build), never in the product library.
#ifdef BC_TRACE
#define DBG_SKIP(bits) ((p.dbg_skip & (bits)) != 0)
#define DBG_BSHIFT (p.dbg_bshift)
#else
#define DBG_SKIP(bits) false
#define DBG_BSHIFT 0
#endif

// 4 fp32 values -> bf16 hi (and lo = v - hi) halves of a 16-byte K-major row chunk
template <int SPLIT>
__device__ __forceinline__ void store_quad(const float4& v, uint8_t* d8, uint32_t lo_offset) {
  uint2 h;
  h.x = pack_bf16x2(v.x, v.y); h.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(d8) = h;
  if (SPLIT == 2) {
    uint2 l;
    float r0, r1, r2, r3;
    unpack2(sub2(pack2(v.x, v.y), bf16x2_as_f32x2(h.x)), r0, r1);
    unpack2(sub2(pack2(v.z, v.w), bf16x2_as_f32x2(h.y)), r2, r3);
    l.x = pack_bf16x2(r0, r1);
    l.y = pack_bf16x2(r2, r3);
    *reinterpret_cast<uint2*>(d8 + lo_offset) = l;
  }
}

// the same in two steps (BC_STREAM_PROD_ILP): all the arithmetic of a batch first, branch-free, then the predicated stores
template <int SPLIT>
__device__ __forceinline__ void split_quad(const float4& v, uint2& h, uint2& l) {
  h.x = pack_bf16x2(v.x, v.y); h.y = pack_bf16x2(v.z, v.w);
  if (SPLIT == 2) {
    float r0, r1, r2, r3;
    unpack2(sub2(pack2(v.x, v.y), bf16x2_as_f32x2(h.x)), r0, r1);
    unpack2(sub2(pack2(v.z, v.w), bf16x2_as_f32x2(h.y)), r2, r3);
    l.x = pack_bf16x2(r0, r1);
    l.y = pack_bf16x2(r2, r3);
  }
}
// The producers' per-element chain (SnakeBeta -> split) is latency-bound when few producer warps are active at a time
// (ncu source view of the 32 -> 64 conv: 24 % of all warp samples in fixed-latency dependency stalls): with
// `if (row valid) { snake; split; store }` per item the compiler emits one branch-guarded block per item and the chains
// run one after the other; computing the whole batch unconditionally lets it interleave P_BATCH independent chains.
#ifndef BC_STREAM_PROD_ILP
#define BC_STREAM_PROD_ILP 1
#endif
#ifndef BC_STREAM_MCHUNK
#define BC_STREAM_MCHUNK 4
#endif

// (n-tile, item, tile-in-item) of a CTA's current tile, advanced incrementally: the tile loop has no division.
// Tile order: the n-tiles of one 128-step row tile are NEIGHBOURS (tile = row_tile * n_tiles + nt), so the CTAs that run
// side by side read the same x tile and it comes from HBM once (n-tile-major order streamed x once per n-tile: 16 times
// for the 512 -> 2048 LSTM input projection, which made that launch HBM-bound).
struct TilePos {
  int nt, b, tt, inc_nt, inc_rt;
  __device__ __forceinline__ void init(int tile, int inc, int tiles_per_item, int n_tiles) {
    const int rt = tile / n_tiles;
    nt = tile - rt * n_tiles;
    b = rt / tiles_per_item;
    tt = rt - b * tiles_per_item;
    inc_rt = inc / n_tiles;
    inc_nt = inc - inc_rt * n_tiles;
  }
  __device__ __forceinline__ void advance(int tiles_per_item, int n_tiles) {
    nt += inc_nt;
    tt += inc_rt;
    if (nt >= n_tiles) { nt -= n_tiles; ++tt; }
    while (tt >= tiles_per_item) { tt -= tiles_per_item; ++b; }
  }
};

// The tiles one CTA walks, in order.  Single-CTA form: tiles first, first + step, ... (TilePos, no division per tile).
// Pair form: the cluster takes the tile PAIRS Q = pi, pi + npairs, ...; pair Q covers row tiles 2q + {0, 1} (q = Q / n_tiles)
// of n-tile Q % n_tiles, and this CTA the one of its rank (`valid` false: the phantom tile behind an odd count -- staged
// as zeros, never stored).
template <bool PAIR>
struct Walk {
  int nt, b, tt;
  bool valid;
  TilePos tp;
  int Q;
  __device__ __forceinline__ void locate(const SParams& p, int rank) {
    const int q = Q / p.n_tiles;
    nt = Q - q * p.n_tiles;
    const int ti = 2 * q + rank;
    valid = ti < p.tiles_per_nt;
    const int tl = valid ? ti : p.tiles_per_nt - 1;
    b = tl / p.tiles_per_item;
    tt = tl - b * p.tiles_per_item;
  }
  __device__ __forceinline__ void init(const SParams& p, int first, int step, int rank) {
    if (PAIR) { Q = first; locate(p, rank); }
    else { tp.init(first, step, p.tiles_per_item, p.n_tiles); nt = tp.nt; b = tp.b; tt = tp.tt; valid = true; }
  }
  __device__ __forceinline__ void advance(const SParams& p, int step, int rank) {
    if (PAIR) { Q += step; locate(p, rank); }
    else { tp.advance(p.tiles_per_item, p.n_tiles); nt = tp.nt; b = tp.b; tt = tp.tt; }
  }
};

// stage stamps are compiled in only with -DBC_TRACE (BC_TRACE=1 python -m audiotokenization_b200.build)
#ifdef BC_TRACE
#define STRACE(ev) do { if (p.trace && blockIdx.x == 0 && it < 64 && lane == 0) p.trace[it * 16 + (ev)] = clock64(); } while (0)
#define STRACE_ON (p.trace != nullptr)
#else
#define STRACE(ev) do { } while (0)
#define STRACE_ON false
#endif

// Tensor-map TMA forms of the plain convs (XMODE 1, the default: BC_STREAM_TMA): the store warps write swizzled
// [32 rows][128 B] tiles that leave by cp.async.bulk.tensor stores (UTMASTG) instead of a transposing pass through a padded
// staging block + coalesced STG -- the STORE stage of a tile falls from 4.0 k to 1.9 k cycles and the output-heavy LSTM
// input projection (512 -> 2048, 10 GB of y per launch) runs 16 % faster; the other convs are within +-2 %.
// XMODE 2 (BC_STREAM_TMA=2) also moves x by TMA: a loader thread keeps a ring of fp32 boxes {16 channels, x_rows rows} in
// flight (UTMALDG -> X_FULL) and the producers read them from shared memory.  Measured NOT faster (32->64: 266 us both
// ways, 64->128 / 128->256 / 256->512 4-30 % slower): the producers are not waiting for HBM, they are bound by their own
// arithmetic (SnakeBeta with range reduction + the hi/lo split: ~45 instructions per 4 elements, one dependent chain per
// float4), and the extra shared-memory round trip of x costs more than the shorter load latency gains.
// XMODE: 0 = no tensor maps, 1 = y by TMA stores (x through the producers' own loads), 2 = x and y by TMA.
template <int SPLIT, bool FUSE, bool PAIR, int XMODE = 0>
__device__ __forceinline__ void conv_stream_body(const SParams& p, const CUtensorMap* tmx = nullptr, const CUtensorMap* tmy = nullptr) {
  constexpr bool XT = XMODE != 0, XL = XMODE == 2;
  static_assert(!(XT && FUSE), "the TMA form exists for the plain convs");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = PAIR ? (int)cluster_rank() : 0;
  // hand-offs towards the MMA thread come from both CTAs of a pair and land on the leader's barrier; hand-offs from it
  // are multicast commits that every CTA waits for on its local copy
#define ARRIVE_MMA(bar) do { if (PAIR) mbar_arrive_leader(bar); else mbar_arrive(bar); } while (0)
#define WAIT_MMA(bar, par) do { if (PAIR) mbar_wait_cluster(bar, par); else mbar_wait(bar, par); } while (0)
#define COMMIT(bar) do { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); } while (0)
  using R = Roles<FUSE>;
  constexpr int N_PROD = R::PROD, MID_WARP0 = R::MID0, EPI_WARP0 = R::EPI0, LOAD_WARP = R::LOAD, MMA_WARP = R::MMA;
  constexpr int S_THREADS = R::THREADS, P_BATCH = R::PB;
  constexpr int M_CHUNK = P_BATCH <= 6 ? P_BATCH : BC_STREAM_MCHUNK;   // chains interleaved at a time (bounds the registers of a long batch)
  const uint32_t plane_bytes = p.plane_bytes;
  const uint32_t a_split = 2u * plane_bytes;
  const uint32_t a2_split = (uint32_t)(A2_CH / 8) * A2_PLANE;
  const uint32_t a2_chunk = a2_split * SPLIT;
  // TMA form: the swizzled store tiles come first (1024-byte aligned), the x ring sits behind the weight ring
  uint8_t* sA = smem_raw + (XT ? (size_t)p.epi_warps * XT_STAGE_BYTES : 0);
  uint8_t* sB = sA + (size_t)p.a_stage * p.NA;
  uint8_t* sA2 = sB + (size_t)p.unit_bytes * p.NB;
  uint8_t* sX = sA2;                                       // XT: NX x x_unit
  uint8_t* sStage = XT ? smem_raw : sA2 + (FUSE ? 2u * a2_chunk : 0u);   // EPI_WARPS x [32][EPI_LD] fp32
  uint64_t* bars = XT ? reinterpret_cast<uint64_t*>(sX + (size_t)p.NX * p.x_unit)
                      : reinterpret_cast<uint64_t*>(sStage + p.epi_warps * 32 * EPI_LD * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  uint32_t* s_off = tmem_slot + 2;   // [32 + MAX_TPU] slab-row shift of tap k (16-byte units), padded for the unrolled issue block
  float* sPar = reinterpret_cast<float*>(s_off + 32 + MAX_TPU + 2);   // fused: b7 | snake2 a | snake2 1/b | b1, N floats each (16-byte aligned)
  const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

  if (tid == 0) {
    constexpr int NCTA = PAIR ? 2 : 1;
    for (int s = 0; s < 4; ++s) {
      mbar_init(BAR(B_A_FULL + s), NCTA * (N_PROD / p.NT));
      mbar_init(BAR(B_A_EMPTY + s), 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(BAR(B_B_FULL + s), 1);
      mbar_init(BAR(B_B_EMPTY + s), 1);
      mbar_init(BAR(B_B_READY + s), 2);
    }
    if (XL) {
      for (int s = 0; s < MAX_NX; ++s) {
        mbar_init(BAR(B_X_FULL + s), 1);
        mbar_init(BAR(B_X_EMPTY + s), p.x_gpb * (N_PROD / p.NT));
      }
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(B_ACC1_FULL + s), 1);
      mbar_init(BAR(B_ACC1_EMPTY + s), NCTA * (FUSE ? MID_WARPS : p.epi_warps));
      mbar_init(BAR(B_A2_FULL + s), NCTA * MID_WARPS);
      mbar_init(BAR(B_A2_EMPTY + s), 1);
      mbar_init(BAR(B_ACC2_FULL + s), 1);
      mbar_init(BAR(B_ACC2_EMPTY + s), NCTA * p.epi_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32 + MAX_TPU) {
    const int sh = min(tid, p.K - 1) * p.dil;
    s_off[tid] = (uint32_t)(sh % p.stride) * (uint32_t)p.rpp + (uint32_t)(sh / p.stride);
  }
  if (tid == 0) { s_off[32 + MAX_TPU] = p.idesc; s_off[32 + MAX_TPU + 1] = (uint32_t)p.dil; }
  if (XL && (p.flags & BC_CONV_SNAKE_IN)) {   // SnakeBeta parameters of the prologue: a | 1/b, C_in floats each
    for (int i = tid; i < p.C_in; i += S_THREADS) {
      sPar[i] = __ldg(p.sa1 + i);
      sPar[p.C_in + i] = __ldg(p.sib1 + i);
    }
  }
  if (FUSE) {
    for (int i = tid; i < p.N; i += S_THREADS) {
      sPar[i] = __ldg(p.bias + i);
      sPar[p.N + i] = __ldg(p.sa2 + i);
      sPar[2 * p.N + i] = __ldg(p.sib2 + i);
      sPar[3 * p.N + i] = __ldg(p.bias2 + i);
    }
  }
  if (warp == MMA_WARP) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                     // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // single CTA: tiles blockIdx.x, + gridDim.x, ...; pair: the cluster's tile pairs (blockIdx.x / 2), + gridDim.x / 2, ...
  const int first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_units = PAIR ? p.total_pairs : p.total_tiles;
  const bool defer = FUSE && p.acc_stages == 2;   // 1x1 conv of tile i issued after the K-tap conv of tile i+1
  int n_my = 0;
  for (int tile = first; tile < n_units; tile += step) ++n_my;

  const bool freerun = DBG_SKIP(8);      // timing experiment: MMA thread free-runs on whatever is in smem
  const bool free_b = DBG_SKIP(16);      // ... only the weight ring is ignored
  if (warp < N_PROD) {
    if (R::REG_PROD < R::REG_LAUNCH) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::REG_PROD));
    if (freerun) goto done;
    // ======================= PRODUCE: activation slabs, one 16-channel group per stage =======================
    // The teams take the groups round-robin, so up to four groups' HBM loads are in flight.  Within a warp 4 lanes cover the 64 contiguous bytes a row holds for this
    // group (one LDG.128 each) and 8 row-quads go side by side: a load instruction touches 8 half-lines
    // instead of 32 lines, which keeps the L1 wavefront queue out of the critical path.
    // One team per ring slot (NA teams of PROD_WARPS / NA warps): a slot's barriers then see one producer, which
    // is never more than one phase ahead of them.
    // TMA form: TWO teams that alternate between two slots each, so a team stages tile i+1 while the MMAs read tile i
    const int n_teams = XL ? p.NT : p.NA;            // (every other form: one team per slot)
    const int team_warps = N_PROD / n_teams;
    const int team = warp / team_warps, tw = warp - team * team_warps;
    if (team >= n_teams) goto done;
    const int rstep = 8 * team_warps;               // rows covered by one load instruction of the team
    const int c4 = lane & 3;                        // which 4 of the group's 16 channels
    const int r_first = tw * 8 + (lane >> 2);       // slab rows r_first + rstep*j
    const bool snake = (p.flags & BC_CONV_SNAKE_IN) != 0;
    const int ph_first = r_first % p.stride, rr_first = r_first / p.stride;   // once per thread
    const int rr_step = rstep / p.stride, ph_step = rstep - rr_step * p.stride;
    const uint32_t dst_off = (uint32_t)(c4 >> 1) * plane_bytes + (uint32_t)(c4 & 1) * 8u;
    int slot_c = 0, use_c = 0, sq = 0;   // running ring position over ALL stages (every team counts every stage)
    uint32_t xslot = 0, xph = 0;         // TMA form: ring position of the next x unit (same count in every team)
    Walk<PAIR> tp;
    tp.init(p, first, step, rank);
    for (int it = 0; it < n_my; ++it, tp.advance(p, step, rank)) {
      const int b = tp.b;
      const int t0 = tp.tt * BM;
      // a phantom tile (pair form, odd tile count) is staged as if it lay entirely beyond the item: all zeros
      const int g0row = tp.valid ? t0 * p.stride - p.pad_left + (tp.nt >= p.nt_shift ? 1 : 0) : p.T_in + p.slab_rows;
      const float* xb = p.x + (size_t)b * p.T_in * p.C_in + c4 * 4;
      long long wE = 0, wL = 0, wM = 0, wS = 0, wQ = 0;
      for (int g = 0; g < p.groups; ++g, ++sq) {
        const int slot = slot_c, use = use_c;
        if (++slot_c == p.NA) { slot_c = 0; ++use_c; }
        const bool mine = XL ? (slot % p.NT == team) : (slot == team);
        if (XL) {
          const int gsub = g % p.x_gpb;             // a unit carries x_gpb neighbouring groups (one team each)
          const bool last_sub = gsub == p.x_gpb - 1;
          if (!mine) {                  // another team's group: only the ring position moves on
            if (last_sub) {
              xslot += (uint32_t)p.x_nbox;
              while (xslot >= (uint32_t)p.NX) { xslot -= (uint32_t)p.NX; xph ^= 1u; }
            }
            continue;
          }
          if (g == 0 && tw == 0) STRACE(0);
          float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
          if (snake) {
            sa = *reinterpret_cast<const float4*>(sPar + g * 16 + c4 * 4);
            sb = *reinterpret_cast<const float4*>(sPar + p.C_in + g * 16 + c4 * 4);
          }
          uint8_t* dst = sA + (size_t)slot * p.a_stage + dst_off;
          const uint32_t xpitch = (uint32_t)p.x_gpb * 64u;
          uint32_t us = xslot, uph = xph;           // this group's units
          for (int c = 0; c < p.x_nbox; ++c) {
            long long tx_ = STRACE_ON ? clock64() : 0;
            mbar_wait(BAR(B_X_FULL + us), uph);
            if (STRACE_ON) wL += clock64() - tx_;
            if (c == 0) {
              long long tw_ = STRACE_ON ? clock64() : 0;
              mbar_wait(BAR(B_A_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
              if (STRACE_ON) wE += clock64() - tw_;
            }
            const uint8_t* src = sX + (size_t)us * p.x_unit + (size_t)gsub * 64 + (size_t)c4 * 16;
            // slab row r = c * x_rows + i lives at (phase r % stride, row r / stride)
            int ph = 0, rr = r_first;
            if (p.stride > 1) { const int r = c * p.x_rows + r_first; rr = r / p.stride; ph = r - rr * p.stride; }
            for (int i0 = r_first; i0 < p.x_rows; i0 += rstep * P_BATCH) {
              float4 v4[P_BATCH];
              if (DBG_SKIP(1)) break;                    // timing experiment: no loads, math or stores
              const long long tq0_ = STRACE_ON ? clock64() : 0;
#pragma unroll
              for (int j = 0; j < P_BATCH; ++j) {
                if (i0 + rstep * j < p.x_rows) v4[j] = *reinterpret_cast<const float4*>(src + (size_t)(i0 + rstep * j) * xpitch);
                else v4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              if (STRACE_ON) {
                float acc_ = 0.f;
#pragma unroll
                for (int j = 0; j < P_BATCH; ++j) acc_ += v4[j].x;
                if (acc_ == 1.2345e-30f) ++wE;
                wQ += clock64() - tq0_;                  // shared-memory loads landed
              }
              // branch-free arithmetic of the whole batch, then the predicated stores (see BC_STREAM_PROD_ILP)
              uint2 hq[P_BATCH], lq[P_BATCH];
              if (snake && !DBG_SKIP(32)) {              // snake(0) == 0: rows outside the item stay zero
#pragma unroll
                for (int j = 0; j < P_BATCH; ++j) snake4<SPLIT>(v4[j], sa, sb);
              }
#pragma unroll
              for (int j = 0; j < P_BATCH; ++j) split_quad<SPLIT>(v4[j], hq[j], lq[j]);
#pragma unroll
              for (int j = 0; j < P_BATCH; ++j) {
                if (i0 + rstep * j < p.x_rows) {
                  uint8_t* d_ = dst + ((uint32_t)ph * (uint32_t)p.rpp + (uint32_t)rr) * 16u;
                  *reinterpret_cast<uint2*>(d_) = hq[j];
                  if (SPLIT == 2) *reinterpret_cast<uint2*>(d_ + a_split) = lq[j];
                }
                ph += ph_step; rr += rr_step;
                if (ph >= p.stride) { ph -= p.stride; ++rr; }
              }
            }
            if (STRACE_ON) wM += clock64() - tx_;
            __syncwarp();
            if (STRACE_ON) wS += clock64() - tx_;
            if (lane == 0) mbar_arrive(BAR(B_X_EMPTY + us));        // this warp has read its rows of the unit
            if (++us == (uint32_t)p.NX) { us = 0; uph ^= 1u; }
          }
          if (last_sub) { xslot = us; xph = uph; }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) ARRIVE_MMA(BAR(B_A_FULL + slot));
          if (g == p.groups - 1 && tw == 0) STRACE(1);
          continue;
        }
        if (!mine) continue;
        if (g == 0 && tw == 0) STRACE(0);
        float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
        if (snake) {
          sa = __ldg(reinterpret_cast<const float4*>(p.sa1 + g * 16 + c4 * 4));
          sb = __ldg(reinterpret_cast<const float4*>(p.sib1 + g * 16 + c4 * 4));
        }
        const float* xcol = xb + g * 16;
        uint8_t* dst = sA + (size_t)slot * p.a_stage + dst_off;
        bool waited = false;
        if (p.stride == 1 && g0row >= 0 && g0row + p.slab_rows <= p.T_in && !DBG_SKIP(1)) {
          // interior tile of an un-strided conv (almost every tile): no bounds tests, pointers advance by constants
          const float* src = xcol + (size_t)(g0row + r_first) * p.C_in;
          const size_t rstride = (size_t)rstep * p.C_in;
          uint8_t* d8 = dst + (size_t)r_first * 16;
          const uint32_t dstep = (uint32_t)rstep * 16u;
          for (int r0 = r_first; r0 < p.slab_rows; r0 += rstep * P_BATCH, src += P_BATCH * rstride, d8 += P_BATCH * dstep) {
            float4 v4[P_BATCH];
#pragma unroll
            for (int j = 0; j < P_BATCH; ++j) {
              if (r0 + rstep * j < p.slab_rows) v4[j] = __ldg(reinterpret_cast<const float4*>(src + j * rstride));
              else if (BC_STREAM_PROD_ILP) v4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!waited) {
              long long tw_ = STRACE_ON ? clock64() : 0;
              mbar_wait(BAR(B_A_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
              if (STRACE_ON) wE += clock64() - tw_;
              waited = true;
            }
            if (BC_STREAM_PROD_ILP) {
#pragma unroll
              for (int j0 = 0; j0 < P_BATCH; j0 += M_CHUNK) {     // M_CHUNK independent chains at a time
                uint2 hq[M_CHUNK], lq[M_CHUNK];
                if (snake) {
#pragma unroll
                  for (int j = 0; j < M_CHUNK; ++j) if (j0 + j < P_BATCH) snake4<SPLIT>(v4[j0 + j], sa, sb);
                }
#pragma unroll
                for (int j = 0; j < M_CHUNK; ++j) if (j0 + j < P_BATCH) split_quad<SPLIT>(v4[j0 + j], hq[j], lq[j]);
#pragma unroll
                for (int j = 0; j < M_CHUNK; ++j) {
                  if (j0 + j < P_BATCH && r0 + rstep * (j0 + j) < p.slab_rows) {
                    *reinterpret_cast<uint2*>(d8 + (j0 + j) * dstep) = hq[j];
                    if (SPLIT == 2) *reinterpret_cast<uint2*>(d8 + (j0 + j) * dstep + a_split) = lq[j];
                  }
                }
              }
            } else {
#pragma unroll
            for (int j = 0; j < P_BATCH; ++j) {
              if (r0 + rstep * j < p.slab_rows) {
                float4 v = v4[j];
                if (snake) {
                  snake4<SPLIT>(v, sa, sb);
                }
                store_quad<SPLIT>(v, d8 + j * dstep, a_split);
              }
            }
            }
          }
        } else {
          // strided convs and edge tiles: slab row r lives at (phase r % stride, row r / stride); both advance
          // incrementally from the thread's constant first row
          const bool inside = g0row >= 0 && g0row + p.slab_rows <= p.T_in;
          int ph = ph_first, rr = rr_first;
          for (int r0 = r_first; r0 < p.slab_rows && !DBG_SKIP(1); r0 += rstep * P_BATCH) {
            float4 v4[P_BATCH];
#pragma unroll
            for (int j = 0; j < P_BATCH; ++j) {
              const int r = r0 + rstep * j;
              const int gr = g0row + r;
              if (r < p.slab_rows && (inside || (gr >= 0 && gr < p.T_in))) v4[j] = __ldg(reinterpret_cast<const float4*>(xcol + (size_t)gr * p.C_in));
              else v4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!waited) {   // the loads above are in flight while we wait for the slot
              long long tw_ = STRACE_ON ? clock64() : 0;
              mbar_wait(BAR(B_A_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
              if (STRACE_ON) wE += clock64() - tw_;
              waited = true;
            }
            long long tl_ = 0;
            if (STRACE_ON) {   // debug: time until the batch's loads have all landed, then the math + stores
              tl_ = clock64();
              float acc_ = 0.f;
#pragma unroll
              for (int j = 0; j < P_BATCH; ++j) acc_ += v4[j].x;
              if (acc_ == 1.2345e-30f) ++wE;
              const long long t2_ = clock64();
              wL += t2_ - tl_;
              tl_ = t2_;
            }
            if (BC_STREAM_PROD_ILP) {
#pragma unroll
              for (int j0 = 0; j0 < P_BATCH; j0 += M_CHUNK) {     // M_CHUNK independent chains at a time
                uint2 hq[M_CHUNK], lq[M_CHUNK];
                if (snake) {   // snake(0) == 0: padding rows stay zero
#pragma unroll
                  for (int j = 0; j < M_CHUNK; ++j) if (j0 + j < P_BATCH) snake4<SPLIT>(v4[j0 + j], sa, sb);
                }
#pragma unroll
                for (int j = 0; j < M_CHUNK; ++j) if (j0 + j < P_BATCH) split_quad<SPLIT>(v4[j0 + j], hq[j], lq[j]);
#pragma unroll
                for (int j = 0; j < M_CHUNK; ++j) {
                  if (j0 + j < P_BATCH) {
                    if (r0 + rstep * (j0 + j) < p.slab_rows) {
                      uint8_t* d_ = dst + ((uint32_t)ph * (uint32_t)p.rpp + (uint32_t)rr) * 16u;
                      *reinterpret_cast<uint2*>(d_) = hq[j];
                      if (SPLIT == 2) *reinterpret_cast<uint2*>(d_ + a_split) = lq[j];
                    }
                    ph += ph_step; rr += rr_step;
                    if (ph >= p.stride) { ph -= p.stride; ++rr; }
                  }
                }
              }
            } else {
#pragma unroll
            for (int j = 0; j < P_BATCH; ++j) {
              if (r0 + rstep * j < p.slab_rows) {
                float4 v = v4[j];
                if (snake) {   // snake(0) == 0: padding rows stay zero
                  snake4<SPLIT>(v, sa, sb);
                }
                store_quad<SPLIT>(v, dst + ((uint32_t)ph * (uint32_t)p.rpp + (uint32_t)rr) * 16u, a_split);
              }
              ph += ph_step; rr += rr_step;
              if (ph >= p.stride) { ph -= p.stride; ++rr; }
            }
            }
            if (STRACE_ON) wM += clock64() - tl_;
          }
        }
        if (!waited) mbar_wait(BAR(B_A_EMPTY + slot), (uint32_t)((use & 1) ^ 1));
        fence_async_smem();
        __syncwarp();
        if (lane == 0) ARRIVE_MMA(BAR(B_A_FULL + slot));
        if (g == p.groups - 1 && tw == 0) STRACE(1);
      }
      if (STRACE_ON && blockIdx.x == 0 && it < 64 && warp == 0 && lane == 0) { p.trace[it * 16 + 12] = wE; p.trace[it * 16 + 13] = wL; p.trace[it * 16 + 14] = wM; p.trace[it * 16 + 15] = wS; p.trace[it * 16 + 11] = wQ; }
    }
  } else if (warp >= LOAD_WARP) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::REG_CTRL));
   if (warp == LOAD_WARP) {
    // ======================= WEIGHTS: unit ring, same order as the MMA warp consumes =======================
    if (lane == 0 && !freerun && !free_b) {
      uint32_t slot = 0, phase = 1;   // ring position; `phase` = parity a free slot's empty barrier must have completed
      Walk<PAIR> ltp;
      ltp.init(p, first, step, rank);
      const uint32_t uB = smem_u32(sB);
      const int nchunk = p.N / A2_CH;
      const int last = defer ? n_my : n_my - 1;
      // pair form: the image holds, per n-tile, rank 0's half of every unit (rows [0, N/2) of each k-plane) and then
      // rank 1's; p.tap_bytes is already the per-rank size
      const size_t nt_bytes = (size_t)p.groups * p.K * p.tap_bytes;
      const uint8_t* w1r = PAIR ? p.w1 + (size_t)rank * (p.N / 16) * p.tap_bytes : p.w1;
      for (int it = 0; it <= last; ++it) {
        if (it < n_my) {
          const int nt = ltp.nt;
          ltp.advance(p, step, rank);
          const uint8_t* wnt = p.w7 + (PAIR ? (size_t)(2 * nt + rank) : (size_t)nt) * nt_bytes;
          for (int g = 0; g < p.groups; ++g)
            for (int u = 0; u < p.upg; ++u) {
              const int k0 = u * p.tpu, k1 = min(p.K, k0 + p.tpu);
              mbar_wait(BAR(B_B_EMPTY + slot), phase);
              bulk_g2s(uB + slot * p.unit_bytes, wnt + (size_t)(g * p.K + k0) * p.tap_bytes, ((uint32_t)(k1 - k0) * p.tap_bytes) >> DBG_BSHIFT,
                       BAR(B_B_FULL + slot));
              if (++slot == (uint32_t)p.NB) { slot = 0; phase ^= 1u; }
            }
        }
        if (FUSE) {
          const int j = defer ? it - 1 : it;
          if (j >= 0 && j < n_my) {
            for (int c = 0; c < nchunk; ++c)
              for (int gu = 0; gu < 4 / p.gpu1; ++gu) {
                mbar_wait(BAR(B_B_EMPTY + slot), phase);
                bulk_g2s(uB + slot * p.unit_bytes, w1r + (size_t)(c * 4 + gu * p.gpu1) * p.tap_bytes, ((uint32_t)p.gpu1 * p.tap_bytes) >> DBG_BSHIFT,
                         BAR(B_B_FULL + slot));
                if (++slot == (uint32_t)p.NB) { slot = 0; phase ^= 1u; }
              }
          }
        }
      }
    }
   } else if (PAIR && warp == LOAD_WARP + 2) {
    // ======================= RELAY (pair form): this CTA's half of a weight unit has landed -> tell the leader =======================
    // A bulk copy completes on a barrier of its own CTA; the leader's MMA thread waits on B_READY (one arrival per CTA).
    if (lane == 0 && !freerun && !free_b) {
      const int nchunk = p.N / A2_CH;
      long long units = (long long)n_my * p.groups * p.upg + (FUSE ? (long long)n_my * nchunk * (4 / p.gpu1) : 0);
      uint32_t slot = 0, phase = 0;
      for (; units > 0; --units) {
        mbar_wait(BAR(B_B_FULL + slot), phase);
        mbar_arrive_leader(BAR(B_B_READY + slot));
        if (++slot == (uint32_t)p.NB) { slot = 0; phase ^= 1u; }
      }
    }
   } else if (XL && warp == LOAD_WARP + 3) {
    // ======================= X LOADER (TMA form): fp32 boxes of x, in the order the producer teams consume them =======================
    if (lane == 0 && !freerun) {
      tma_prefetch_desc(tmx);
      Walk<PAIR> xtp;
      xtp.init(p, first, step, rank);
      const uint32_t uX = smem_u32(sX);
      uint32_t slot = 0, phase = 1;
      for (int it = 0; it < n_my; ++it, xtp.advance(p, step, rank)) {
        // a phantom tile (pair form, odd tile count) is read from beyond the item: all zeros
        const int g0row = xtp.valid ? xtp.tt * BM * p.stride - p.pad_left + (xtp.nt >= p.nt_shift ? 1 : 0) : p.T_in + p.slab_rows;
        for (int g = 0; g < p.groups; g += p.x_gpb)
          for (int c = 0; c < p.x_nbox; ++c) {
            mbar_wait(BAR(B_X_EMPTY + slot), phase);
            tma_load_3d(uX + slot * p.x_unit, tmx, g * 16, g0row + c * p.x_rows, xtp.b, BAR(B_X_FULL + slot), p.x_bytes);
            if (++slot == (uint32_t)p.NX) { slot = 0; phase ^= 1u; }
          }
      }
    }
   } else if (warp == MMA_WARP && (!PAIR || rank == 0)) {
    // ======================= MMA issue: one elected lane runs the whole role =======================
    // The tensor pipe queues only a couple of MMAs, so every instruction between two MMA bursts costs tensor time
    // (scripts/probes/mma_probe.cu, modes 100+: a satisfied mbarrier try_wait ~55 cycles, a tcgen05.commit ~35-60,
    // runtime tap guards ~10 per tap; a bare loop reaches the 64-cycle floor at N=128, this nest ~80).  One elected
    // lane runs the whole role, so nothing re-converges per unit and no descriptor is recomputed per lane.
    // Optional (off): commits of a unit issued behind the first tap of the next one, next unit's barriers waited for
    // behind the first tap of the current one (BC_STREAM_DEFER_COMMITS / BC_STREAM_PREFETCH_WAITS).
    const uint32_t hi_d = desc_hi(128u);
    const uint32_t uA = smem_u32(sA), uB = smem_u32(sB), uA2 = smem_u32(sA2);
    const uint32_t n_rows = PAIR ? (uint32_t)p.N / 2u : (uint32_t)p.N;   // B rows held by this SM
    const uint32_t b_kplane = n_rows * 16u;              // LBO of the B operand: stride between the two k-planes
    const uint32_t b_lo_off = (n_rows * 32u) >> 4;       // lo split of a tap, in 16-byte units
    const uint32_t tap16 = p.tap_bytes >> 4;
    const uint32_t a_sp = a_split >> 4;
    const int nchunk = p.N / A2_CH;
    // Kernel parameters used between two MMAs are read back from shared memory (written in the prologue), so that
    // they live in registers instead of being re-read from the constant bank inside the issue block
    const uint32_t idesc = *reinterpret_cast<volatile uint32_t*>(s_off + 32 + MAX_TPU);
    const uint32_t dil_r = *reinterpret_cast<volatile uint32_t*>(s_off + 32 + MAX_TPU + 1);
    const uint32_t units2 = 4u / (uint32_t)p.gpu1;
    const int last = defer ? n_my : n_my - 1;
    if (elect_one()) {
      uint32_t aslot = 0, aph = 0, bslot = 0, bph = 0, a2seq = 0;
      uint32_t pend_b = 0, pend_a = 0, pend_acc = 0;     // commits owed for the previous unit (barrier addresses, 0 = none)
      bool a_ready = false, b_ready = false;              // the next slab / weight unit has already been waited for
      long long units_left = (long long)n_my * p.groups * p.upg + (FUSE ? (long long)n_my * nchunk * units2 : 0);
#define FLUSH_COMMITS() do { if (pend_b) COMMIT(pend_b); if (pend_a) COMMIT(pend_a); if (pend_acc) COMMIT(pend_acc); \
                             pend_b = pend_a = pend_acc = 0; } while (0)
#define MMA_RT(d_, a_, b_, acc_) do { if (PAIR) mma2_bf16_rt(d_, a_, b_, hi_d, hi_d, idesc, acc_); else mma_bf16_raw_rt(d_, a_, b_, hi_d, hi_d, idesc, acc_); } while (0)
#define MMA_ACC(d_, a_, b_) do { if (PAIR) mma2_bf16_raw<true>(d_, a_, b_, hi_d, hi_d, idesc); else mma_bf16_raw<true>(d_, a_, b_, hi_d, hi_d, idesc); } while (0)
#define B_RDY (PAIR ? B_B_READY : B_B_FULL)
#define PREFETCH_B() do { if (BC_STREAM_PREFETCH_WAITS && --units_left > 0 && !freerun && !free_b) { \
                            uint32_t nb = bslot + 1, nph = bph; if (nb == (uint32_t)p.NB) { nb = 0; nph ^= 1u; } \
                            WAIT_MMA(BAR(B_RDY + nb), nph); b_ready = true; } } while (0)
      for (int it = 0; it <= last; ++it) {
        if (it < n_my) {
          const int as = p.acc_stages == 2 ? (it & 1) : 0, ause = p.acc_stages == 2 ? (it >> 1) : it;
          FLUSH_COMMITS();
          WAIT_MMA(BAR(B_ACC1_EMPTY + as), (uint32_t)((ause & 1) ^ 1));
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(as * p.acc_stride);
          STRACE(2);
          for (int g = 0; g < p.groups; ++g) {
            if (!a_ready && !freerun) WAIT_MMA(BAR(B_A_FULL + aslot), aph);
            a_ready = false;
            const uint32_t a_lo0 = desc_lo(uA + aslot * p.a_stage, plane_bytes);
            int k = 0;
            for (int u = 0; u < p.upg; ++u) {
              if (!b_ready && !freerun && !free_b) WAIT_MMA(BAR(B_RDY + bslot), bph);
              b_ready = false;
              const int nt = min(p.tpu, p.K - k);
              uint32_t off[MAX_TPU];
#pragma unroll
              for (int j = 0; j < MAX_TPU; ++j)   // un-strided (always, when fused): plain arithmetic, no table look-up
                off[j] = FUSE ? (uint32_t)(k + j) * dil_r : s_off[k + j];
              const uint32_t b_lo0 = desc_lo(uB + bslot * p.unit_bytes, b_kplane);
              const bool last_u = u == p.upg - 1;
#pragma unroll
              for (int j = 0; j < MAX_TPU; ++j) {
                if (j < nt) {
                  const uint32_t a_lo = a_lo0 + off[j], b_lo = b_lo0 + (uint32_t)j * tap16;
                  if (j == 0) MMA_RT(d, a_lo, b_lo, (g | k) ? 1u : 0u);
                  else        MMA_ACC(d, a_lo, b_lo);
                  if (SPLIT == 2) {
                    MMA_ACC(d, a_lo, b_lo + b_lo_off);   // a_hi * w_lo
                    MMA_ACC(d, a_lo + a_sp, b_lo);       // a_lo * w_hi
                  }
                }
                if (j == 0) {   // housekeeping behind the first tap
                  FLUSH_COMMITS();
                  PREFETCH_B();
                  if (BC_STREAM_PREFETCH_WAITS && last_u && g + 1 < p.groups && !freerun) {
                    uint32_t na = aslot + 1, nph = aph;
                    if (na == (uint32_t)p.NA) { na = 0; nph ^= 1u; }
                    WAIT_MMA(BAR(B_A_FULL + na), nph);
                    a_ready = true;
                  }
                }
              }
              if (!freerun && !free_b) pend_b = BAR(B_B_EMPTY + bslot);
              if (last_u) {
                if (!freerun) pend_a = BAR(B_A_EMPTY + aslot);
                if (g == p.groups - 1) pend_acc = BAR(B_ACC1_FULL + as);
              }
              if (!BC_STREAM_DEFER_COMMITS) FLUSH_COMMITS();
              k += nt;
              if (++bslot == (uint32_t)p.NB) { bslot = 0; bph ^= 1u; }
            }
            if (++aslot == (uint32_t)p.NA) { aslot = 0; aph ^= 1u; }
          }
          STRACE(3);
        }
        if (FUSE) {
          const int j = defer ? it - 1 : it;
          if (j >= 0 && j < n_my) {
            const int as = p.acc_stages == 2 ? (j & 1) : 0, ause = p.acc_stages == 2 ? (j >> 1) : j;
            FLUSH_COMMITS();            // MID(j) may be waiting for the accumulator commit that is still owed
            WAIT_MMA(BAR(B_ACC2_EMPTY + as), (uint32_t)((ause & 1) ^ 1));
            tc_fence_after();
            { const int it = j; STRACE(6); }
            const uint32_t d2 = tmem_base + (uint32_t)(as * p.acc_stride + p.N);
            for (int c = 0; c < nchunk; ++c, ++a2seq) {
              const uint32_t s2 = a2seq & 1u, u2 = a2seq >> 1;
              FLUSH_COMMITS();
              WAIT_MMA(BAR(B_A2_FULL + s2), u2 & 1u);
              tc_fence_after();
              const uint32_t a_lo0 = desc_lo(uA2 + s2 * a2_chunk, A2_PLANE);
              for (uint32_t gu = 0; gu < units2; ++gu) {
                if (!b_ready && !freerun && !free_b) WAIT_MMA(BAR(B_RDY + bslot), bph);
                b_ready = false;
                const uint32_t b_lo0 = desc_lo(uB + bslot * p.unit_bytes, b_kplane);
                const uint32_t gc0 = gu * (uint32_t)p.gpu1;
#pragma unroll
                for (int gg = 0; gg < 4; ++gg) {
                  if (gg < p.gpu1) {
                    const uint32_t gc = gc0 + (uint32_t)gg;
                    const uint32_t a_lo = a_lo0 + gc * ((2u * A2_PLANE) >> 4), b_lo = b_lo0 + (uint32_t)gg * tap16;
                    MMA_RT(d2, a_lo, b_lo, ((uint32_t)c | gc) ? 1u : 0u);
                    if (SPLIT == 2) {
                      MMA_ACC(d2, a_lo, b_lo + b_lo_off);
                      MMA_ACC(d2, a_lo + (a2_split >> 4), b_lo);
                    }
                  }
                  if (gg == 0) {
                    FLUSH_COMMITS();
                    PREFETCH_B();
                  }
                }
                if (!freerun && !free_b) pend_b = BAR(B_B_EMPTY + bslot);
                if (gu == units2 - 1) {
                  pend_a = BAR(B_A2_EMPTY + s2);
                  if (c == nchunk - 1) pend_acc = BAR(B_ACC2_FULL + as);
                }
                if (!BC_STREAM_DEFER_COMMITS) FLUSH_COMMITS();
                if (++bslot == (uint32_t)p.NB) { bslot = 0; bph ^= 1u; }
              }
            }
            { const int it = j; STRACE(7); }
          }
        }
      }
      FLUSH_COMMITS();
#undef FLUSH_COMMITS
#undef PREFETCH_B
#undef MMA_RT
#undef MMA_ACC
#undef B_RDY
    }
    __syncwarp();
   }
  } else if (warp < EPI_WARP0) {
    // ======================= MID (fused): acc1 -> +b7 -> snake2 -> bf16 A2 chunks =======================
    if (FUSE) {
      if (R::REG_MID > R::REG_LAUNCH) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R::REG_MID));
      if (R::REG_MID < R::REG_LAUNCH) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::REG_MID));
      const int q = warp & 3;
      const int half = (warp - MID_WARP0) >> 2;
      const int row = q * 32 + lane;
      const int nchunk = p.N / A2_CH;
      uint32_t a2seq = 0;
      for (int it = 0; it < n_my; ++it) {
        const int as = p.acc_stages == 2 ? (it & 1) : 0, ause = p.acc_stages == 2 ? (it >> 1) : it;
        mbar_wait(BAR(B_ACC1_FULL + as), (uint32_t)(ause & 1));
        tc_fence_after();
        if (warp == MID_WARP0) STRACE(4);
        const uint32_t taddr = tmem_base + (uint32_t)(as * p.acc_stride) + ((uint32_t)(q * 32) << 16);
        for (int c = 0; c < nchunk; ++c, ++a2seq) {
          const uint32_t s2 = a2seq & 1u, u2 = a2seq >> 1;
#pragma unroll
          for (int hh = 0; hh < MID_HALVES; ++hh) {
            const int hcol = half + hh * (MID_WARPS / 4);        // 32-column half of the 64-channel chunk
            const int cbase = c * A2_CH + hcol * 32;
            uint32_t r[32];
            tmem_load32(taddr + (uint32_t)cbase, r);
            if (c == nchunk - 1 && hh == MID_HALVES - 1) {   // this warp's share of acc1 is in registers: hand the accumulator back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) ARRIVE_MMA(BAR(B_ACC1_EMPTY + as));
            }
            if (hh == 0) mbar_wait(BAR(B_A2_EMPTY + s2), (u2 & 1u) ^ 1u);
            uint8_t* dst = sA2 + (size_t)s2 * a2_chunk + (size_t)(hcol * 4) * A2_PLANE + (size_t)row * 16;
#pragma unroll
            for (int j = 0; j < (DBG_SKIP(2) ? 0 : 4); ++j) {
              const int ch = cbase + 8 * j;
              const float4 bi0 = *reinterpret_cast<const float4*>(sPar + ch), bi1 = *reinterpret_cast<const float4*>(sPar + ch + 4);
              const float4 s0 = *reinterpret_cast<const float4*>(sPar + p.N + ch), s1 = *reinterpret_cast<const float4*>(sPar + p.N + ch + 4);
              const float4 i0 = *reinterpret_cast<const float4*>(sPar + 2 * p.N + ch), i1 = *reinterpret_cast<const float4*>(sPar + 2 * p.N + ch + 4);
              float v[8];
            acc_bias8(r + 8 * j, bi0, bi1, v);
              snake8<SPLIT>(v, s0, s1, i0, i1);
              split_store<SPLIT>(v, dst + (size_t)j * A2_PLANE, a2_split);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) ARRIVE_MMA(BAR(B_A2_FULL + s2));
        }
        if (warp == MID_WARP0) STRACE(5);
      }
    }
  } else {
    // ======================= STORE: acc -> +bias (+residual) -> y =======================
    // The accumulator arrives one row per lane; HBM wants whole lines.  Each warp owns a padded [32 rows][32 + 4]
    // fp32 staging block: the residual is fetched with 8 lanes per row (4 full lines per load instruction), lands in
    // the block, is combined in place by the lane that owns the row, and leaves the same coalesced way.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R::REG_STORE));
    const int ew = warp - EPI_WARP0;
    if (ew >= p.epi_warps) goto done;
    const int q = warp & 3;
    const int cb0 = (ew >> 2) * 32, cbstep = (p.epi_warps >> 2) * 32;   // this warp's 32-column blocks
    const bool tanh_out = (p.flags & BC_CONV_TANH_OUT) != 0;
    const float* bias = FUSE ? p.bias2 : p.bias;
    float* sT = reinterpret_cast<float*>(sStage) + (size_t)ew * (32 * EPI_LD);
    const int crow = lane >> 3, cchunk = (lane & 7) * 4;          // coalesced mapping: rows crow + 4*i, 4 floats at cchunk
    Walk<PAIR> tp;
    tp.init(p, first, step, rank);
    int xt_seq = 0;      // TMA form: store tiles written so far by this warp
    for (int it = 0; it < n_my; ++it, tp.advance(p, step, rank)) {
      const int as = p.acc_stages == 2 ? (it & 1) : 0, ause = p.acc_stages == 2 ? (it >> 1) : it;
      const int nt = tp.nt, b = tp.b;
      const int trow0 = tp.tt * BM + q * 32;     // first output row of this warp's block
      if (XT) {
        // TMA form: the lane's accumulator row goes into a [32 rows][128 B] tile with the 128-byte swizzle (16-byte chunk
        // index ^ row % 8: conflict-free although every lane writes its own row), which one TMA store un-swizzles on its
        // way to y; rows beyond T_out are clipped by the tensor map.  Two tiles per warp: the previous block's store
        // reads its tile while this block is written.
        const uint32_t fullbar = BAR(B_ACC1_FULL + as), emptybar = BAR(B_ACC1_EMPTY + as);
        const uint32_t taddr = tmem_base + (uint32_t)(as * p.acc_stride) + ((uint32_t)(q * 32) << 16);
        const float* bp = bias ? bias + (size_t)nt * p.N : nullptr;
        for (int c0 = cb0; c0 < p.N; c0 += cbstep, ++xt_seq) {
          if (c0 == cb0) {
            mbar_wait(fullbar, (uint32_t)(ause & 1));
            tc_fence_after();
            if (warp == EPI_WARP0) STRACE(8);
          }
          uint32_t r[32];
          tmem_load32(taddr + (uint32_t)c0, r);
          if (c0 + cbstep >= p.N) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) ARRIVE_MMA(emptybar);
          }
          if (xt_seq >= 2) {                     // the store that read this tile two blocks ago is done with it
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
          }
          uint8_t* tile = sStage + (size_t)ew * XT_STAGE_BYTES + (size_t)(xt_seq & 1) * 4096;
          const uint32_t own = smem_u32(tile) + (uint32_t)lane * 128u, sw = ((uint32_t)lane & 7u) << 4;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                   __uint_as_float(r[4 * j + 3]));
            if (bp) v = add4(v, __ldg(reinterpret_cast<const float4*>(bp + c0) + j));
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(own + (((uint32_t)j << 4) ^ sw)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (tp.valid && !DBG_SKIP(4)) tma_store_3d(tmy, nt * p.N + c0, trow0, b, smem_u32(tile));
            bulk_commit_group();
          }
        }
        if (warp == EPI_WARP0) STRACE(9);
        continue;
      }
      const size_t off0 = ((size_t)b * p.T_out + trow0 + crow) * p.C_out + (size_t)nt * p.N + cchunk;
      const float* rp = (p.res && !DBG_SKIP(4)) ? p.res + off0 : nullptr;
      float* yp = p.y + off0;
      const size_t istep = (size_t)4 * p.C_out;                         // 4 rows further per load/store instruction
      const int rows_ok = tp.valid ? p.T_out - trow0 - crow : 0;       // row 4*i of this lane is valid iff 4*i < rows_ok (phantom tile: none)
      const float* bp = bias ? bias + (size_t)nt * p.N : nullptr;
      const uint32_t taddr = tmem_base + (uint32_t)(as * p.acc_stride + (FUSE ? p.N : 0)) + ((uint32_t)(q * 32) << 16);
      const uint32_t fullbar = BAR((FUSE ? B_ACC2_FULL : B_ACC1_FULL) + as);
      const uint32_t emptybar = BAR((FUSE ? B_ACC2_EMPTY : B_ACC1_EMPTY) + as);
      float4 res4[8];
      if (rp) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          res4[i] = 4 * i < rows_ok ? __ldg(reinterpret_cast<const float4*>(rp + cb0 + i * istep)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int c0 = cb0; c0 < p.N; c0 += cbstep) {
        if (rp) {
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(sT + (4 * i + crow) * EPI_LD + cchunk) = res4[i];
          if (c0 + cbstep < p.N) {   // next block's residual: in flight while this block is combined and stored
#pragma unroll
            for (int i = 0; i < 8; ++i)
              res4[i] = 4 * i < rows_ok ? __ldg(reinterpret_cast<const float4*>(rp + c0 + cbstep + i * istep)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          __syncwarp();
        }
        if (c0 == cb0) {
          mbar_wait(fullbar, (uint32_t)(ause & 1));
          tc_fence_after();
          if (warp == EPI_WARP0) STRACE(8);
        }
        uint32_t r[32];
        tmem_load32(taddr + (uint32_t)c0, r);
        if (c0 + cbstep >= p.N) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) ARRIVE_MMA(emptybar);
        }
        float* own = sT + lane * EPI_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (FUSE) {
            const float4 bb = *reinterpret_cast<const float4*>(sPar + 3 * p.N + c0 + 4 * j);
            v = add4(v, bb);
          } else if (bp) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bp + c0) + j);
            v = add4(v, bb);
          }
          if (rp) {
            const float4 t4 = *reinterpret_cast<const float4*>(own + 4 * j);
            v = add4(v, t4);
          }
          if (tanh_out) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
          *reinterpret_cast<float4*>(own + 4 * j) = v;
        }
        __syncwarp();
        if (!DBG_SKIP(4)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(sT + (4 * i + crow) * EPI_LD + cchunk);
            if (4 * i < rows_ok) *reinterpret_cast<float4*>(yp + c0 + i * istep) = v;
          }
        }
        __syncwarp();
      }
      if (warp == EPI_WARP0) STRACE(9);
    }
    if (XT && lane == 0) bulk_wait_all();      // the last stores have left shared memory and are visible
  }
done:
#undef BAR
#undef ARRIVE_MMA
#undef WAIT_MMA
#undef COMMIT
  // ---- teardown (pair: nobody frees TMEM / exits while the peer may still be reading or signalling) ----
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == MMA_WARP) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    else      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

template <int SPLIT, bool FUSE>
__global__ void __launch_bounds__(Roles<FUSE>::THREADS, 1) conv_stream_kernel(const SParams p) {
  conv_stream_body<SPLIT, FUSE, false>(p);
}
// the same roles on a CTA pair: one tcgen05.mma.cta_group::2 of M = 256 per two tiles, half of every weight unit per SM
template <int SPLIT, bool FUSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Roles<FUSE>::THREADS, 1) conv_stream_pair_kernel(const SParams p) {
  conv_stream_body<SPLIT, FUSE, true>(p);
}

// TMA form of the plain conv (tensor maps of x and y as grid constants)
template <int SPLIT, int XMODE>
__global__ void __launch_bounds__(Roles<false>::THREADS, 1)
conv_stream_tma_kernel(const SParams p, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy) {
  conv_stream_body<SPLIT, false, false, XMODE>(p, &tmx, &tmy);
}
template <int SPLIT, int XMODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Roles<false>::THREADS, 1)
conv_stream_tma_pair_kernel(const SParams p, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy) {
  conv_stream_body<SPLIT, false, true, XMODE>(p, &tmx, &tmy);
}

long long* g_stream_trace = nullptr;

// cuTensorMapEncodeTiled through the runtime's driver entry point query
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// fp32 channels-last tensor [B][T][C] as a 3-D map {C, T, B} with box {box_c, box_t, 1}
int encode_cl_map(CUtensorMap* m, const float* base, int B, int T, int C, int box_c, int box_t, bool swizzle128) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return bc::fail(BC_ENODEVICE, "conv(stream): cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)C * 4u, (cuuint64_t)T * (cuuint64_t)C * 4u};
  const cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_t, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return bc::fail(BC_EINVAL, "conv(stream): cuTensorMapEncodeTiled failed (%d) for [%d][%d][%d] box %d x %d", (int)r, B, T, C, box_c, box_t);
  return BC_OK;
}

struct StreamPlan {
  int N, n_tiles, groups, tpu, upg, gpu1, slab_rows, rpp, NA, NB, acc_stages, acc_stride, tmem_cols, split, epi_warps;
  int NX, x_rows, x_nbox, x_gpb, NT;
  uint32_t a_stage, unit_bytes, tap_bytes, plane_bytes, x_unit, x_bytes;
  size_t smem;
};

// Shared-memory plan of the TMA form (N, split, tap_bytes already set): 8 KB swizzled store tiles per store warp, TWO
// activation stages (the x ring, not the stage count, now covers the HBM latency), a weight ring of >= 3 units and an x
// ring that holds about 48 KB of fp32 boxes in flight.
bool stream_plan_tma(int C_in, int K, int stride, int dilation, StreamPlan* pl, bool loads) {
  const int N = pl->N, split = pl->split;
  pl->slab_rows = (BM - 1) * stride + (K - 1) * dilation + 1;
  pl->rpp = (pl->slab_rows + stride - 1) / stride;
  // stride 2: a half-warp of a producer store covers rows r .. r+3 = two rows of each phase; the phases' blocks must sit
  // 32 bytes apart modulo 64 (rpp = 2 mod 4) or the two 32-byte pieces overlap in the banks (ncu: 308 store conflicts per
  // tile of the 32 -> 64 conv with rpp = 129).  Strides 4 and 5 want rpp = 1 mod 4, which 129 already is.
  if (stride == 2) while (pl->rpp % 4 != 2) ++pl->rpp;
  pl->x_rows = stride > 1 ? pl->rpp : pl->slab_rows;       // one box per phase-sized run of slab rows
  pl->x_nbox = stride > 1 ? stride : 1;
  if (pl->x_rows > 256) return false;                      // TMA box limit
  // a box carries BC_STREAM_XGPB (2) neighbouring 16-channel groups: TMA's cost is per box ROW, and 64-byte rows only
  // reached ~7 B/clk per SM
  pl->x_gpb = BC_STREAM_XGPB;
  while ((C_in / 16) % pl->x_gpb != 0) pl->x_gpb >>= 1;
  pl->x_bytes = (uint32_t)pl->x_rows * 64u * (uint32_t)pl->x_gpb;
  pl->x_unit = (pl->x_bytes + 127u) & ~127u;
  pl->plane_bytes = (uint32_t)stride * pl->rpp * 16u;
  while (pl->plane_bytes % 128u != 64u) pl->plane_bytes += 16u;
  pl->a_stage = (uint32_t)((size_t)split * 2 * pl->plane_bytes + 127) & ~127u;
  if ((size_t)pl->plane_bytes * 2 >= (1u << 18)) return false;
  pl->epi_warps = 4;
  const size_t misc = N_BARS * 8 + 64 + (32 + MAX_TPU + 2) * 4 + 16 + (size_t)pl->epi_warps * XT_STAGE_BYTES + (size_t)2 * C_in * 4;
  const size_t budget = 225 * 1024;
  int want = (int)((BC_STREAM_XWANT + pl->x_unit - 1) / pl->x_unit);
  if (want > MAX_NX) want = MAX_NX;
  if (want < 2) want = 2;
  if (!loads) {
    // stores by TMA, x through the producers' own loads: the old ring logic with the larger store tiles
    pl->NX = 0; pl->x_unit = 0; pl->x_bytes = 0; pl->x_gpb = 1;
    int tpu = (int)(UNIT_MAX_BYTES / pl->tap_bytes);
    if (tpu < 1) tpu = 1;
    if (tpu > MAX_TPU) tpu = MAX_TPU;
    if (tpu > K) tpu = K;
    pl->upg = (K + tpu - 1) / tpu;
    tpu = (K + pl->upg - 1) / pl->upg;
    pl->tpu = tpu;
    pl->unit_bytes = (uint32_t)tpu * pl->tap_bytes;
    int gpu1 = 1;
    while (gpu1 * 2 <= tpu && gpu1 * 2 <= 4) gpu1 *= 2;
    pl->gpu1 = gpu1;
    int NB = (int)(98304u / pl->unit_bytes), NA = 0;
    if (NB > 8) NB = 8;
    if (NB < 3) NB = 3;
    const int groups = C_in / 16;
    const int want_na = groups < 4 ? (groups < 2 ? 2 : groups) : 4;
    int best_nb = 0, best_na = 0;
    for (; NB >= 3; --NB) {
      const size_t used = (size_t)NB * pl->unit_bytes + misc;
      if (used >= budget) continue;
      NA = (int)((budget - used) / pl->a_stage);
      if (NA > 4) NA = 4;
      if (NA > best_na) { best_na = NA; best_nb = NB; }
      if (NA >= want_na) break;
    }
    if (best_na < 2) return false;
    NA = best_na;
    while (Roles<false>::PROD % NA != 0) --NA;
    pl->NA = NA; pl->NB = best_nb; pl->NT = NA;
    pl->acc_stride = N;
    pl->acc_stages = 2 * N <= 512 ? 2 : 1;
    int cols = pl->acc_stages * pl->acc_stride, pw = 32;
    while (pw < cols) pw <<= 1;
    pl->tmem_cols = pw;
    pl->smem = (size_t)best_nb * pl->unit_bytes + misc + (size_t)NA * pl->a_stage;
    return true;
  }
  // four activation stages (two per team: double buffering against the MMAs) when the rings still fit, else two
  int best_nx = 0, NA = 0;
  for (int na = 4; na >= 2 && best_nx < want; na -= 2) {
    best_nx = 0;
    for (uint32_t cap = UNIT_MAX_BYTES; cap >= 16384u && best_nx < want; cap >>= 1) {
      int tpu = (int)(cap / pl->tap_bytes);
      if (tpu < 1) tpu = 1;
      if (tpu > MAX_TPU) tpu = MAX_TPU;
      if (tpu > K) tpu = K;
      const int upg = (K + tpu - 1) / tpu;
      tpu = (K + upg - 1) / upg;
      const uint32_t unit_bytes = (uint32_t)tpu * pl->tap_bytes;
      int NB = (int)(98304u / unit_bytes);
      if (NB > 8) NB = 8;
      for (; NB >= 3; --NB) {
        const size_t used = (size_t)NB * unit_bytes + misc + (size_t)na * pl->a_stage;
        if (used >= budget) continue;
        int nx = (int)((budget - used) / pl->x_unit);
        if (nx > MAX_NX) nx = MAX_NX;
        if (nx > best_nx) {
          best_nx = nx; NA = na;
          pl->tpu = tpu; pl->upg = upg; pl->unit_bytes = unit_bytes; pl->NB = NB; pl->NX = nx;
        }
        if (best_nx >= want) break;
      }
    }
  }
  if (best_nx < 2 || NA < 2) return false;
  int gpu1 = 1;
  while (gpu1 * 2 <= pl->tpu && gpu1 * 2 <= 4) gpu1 *= 2;
  pl->gpu1 = gpu1;
  const size_t used = (size_t)pl->NB * pl->unit_bytes + misc + (size_t)pl->NX * pl->x_unit;
  pl->NA = NA;
  pl->NT = 2;                                            // two teams of six warps; with NA = 4 each alternates between two slots
  pl->acc_stride = N;
  pl->acc_stages = 2 * N <= 512 ? 2 : 1;
  int cols = pl->acc_stages * pl->acc_stride, pw = 32;
  while (pw < cols) pw <<= 1;
  pl->tmem_cols = pw;
  pl->smem = used + (size_t)NA * pl->a_stage;
  return true;
}

bool stream_plan(int C_in, int C_out, int K, int stride, int dilation, int precision, int fused, StreamPlan* pl, bool pair = false, bool tma = false) {
  if (precision != BC_PREC_BF16 && precision != BC_PREC_BF16X3) return false;
  if (C_in % 16 != 0 || C_in < 32 || K < 1 || K > 32 || stride < 1 || dilation < 1) return false;
  if (stride > 1 && dilation > 1) return false;
  int N = 0;
  if (C_out <= 256 && (C_out & (C_out - 1)) == 0) N = C_out;
  else if (C_out % 256 == 0) N = 256;
  else if (C_out % 128 == 0) N = 128;      // e.g. the 5 x 128 channel blocks of a 256 -> 128 stride-5 transposed conv
  else if (C_out % 64 == 0) N = 64;
  if (N < 64 || N % 32 != 0 || (N & (N - 1)) != 0) return false;   // 64, 128, 256
  if (fused && (C_in != C_out || N != C_out || stride != 1 || N % A2_CH != 0)) return false;
  const int split = precision == BC_PREC_BF16X3 ? 2 : 1;
  pl->split = split;
  pl->N = N;
  pl->n_tiles = C_out / N;
  pl->groups = C_in / 16;
  pl->tap_bytes = (uint32_t)(pair ? N / 2 : N) * 32u * split;   // pair form: every SM holds half of the B rows
  pl->NX = 0; pl->x_rows = 0; pl->x_nbox = 0; pl->x_gpb = 1; pl->x_unit = 0; pl->x_bytes = 0;
  if (tma) return !fused && stream_plan_tma(C_in, K, stride, dilation, pl, bc::policy().stream_tma >= 2);
  int tpu = (int)(UNIT_MAX_BYTES / pl->tap_bytes);
  if (tpu < 1) tpu = 1;
  if (tpu > MAX_TPU) tpu = MAX_TPU;
  if (tpu > K) tpu = K;
  pl->upg = (K + tpu - 1) / tpu;
  tpu = (K + pl->upg - 1) / pl->upg;   // balance the units of a group
  pl->tpu = tpu;
  int gpu1 = 1;
  while (gpu1 * 2 <= tpu && gpu1 * 2 <= 4) gpu1 *= 2;
  pl->gpu1 = gpu1;
  pl->unit_bytes = (uint32_t)tpu * pl->tap_bytes;
  pl->slab_rows = (BM - 1) * stride + (K - 1) * dilation + 1;
  pl->rpp = (pl->slab_rows + stride - 1) / stride;
  // stride 2: a half-warp of a producer store covers rows r .. r+3 = two rows of each phase; the phases' blocks must sit
  // 32 bytes apart modulo 64 (rpp = 2 mod 4) or the two 32-byte pieces overlap in the banks (ncu: 308 store conflicts per
  // tile of the 32 -> 64 conv with rpp = 129).  Strides 4 and 5 want rpp = 1 mod 4, which 129 already is.
  if (stride == 2) while (pl->rpp % 4 != 2) ++pl->rpp;
  // plane stride: rows * 16 B, padded so a group's two 8-channel planes sit 64 bytes apart modulo 128
  pl->plane_bytes = (uint32_t)stride * pl->rpp * 16u;
  while (pl->plane_bytes % 128u != 64u) pl->plane_bytes += 16u;
  pl->a_stage = (uint32_t)((size_t)split * 2 * pl->plane_bytes + 127) & ~127u;
  if ((size_t)pl->plane_bytes * 2 >= (1u << 18)) return false;
  const size_t a2 = fused ? (size_t)2 * (A2_CH / 8) * A2_PLANE * split : 0;
  // eight store warps when the tile has at least two 32-column blocks and shared memory allows, else four
  int epi_warps = 4;
  size_t misc = 0;
  const size_t misc0 = N_BARS * 8 + 64 + (32 + MAX_TPU + 2) * 4 + 16 + (fused ? (size_t)4 * N * 4 : 0);
  const size_t budget = 225 * 1024;
  int NB = (int)(98304u / pl->unit_bytes);
  if (NB > 8) NB = 8;
  if (NB < 3) NB = 3;
  int NA = 0;
  const int NB0 = NB;
  for (; epi_warps >= 4; epi_warps -= 4) {   // prefer a full activation ring over the second set of store warps
    misc = misc0 + (size_t)epi_warps * 32 * EPI_LD * 4;
    for (NB = NB0; NB >= 3; --NB) {
      const size_t used = (size_t)NB * pl->unit_bytes + a2 + misc;
      if (used >= budget) continue;
      NA = (int)((budget - used) / pl->a_stage);
      if (NA >= 2) break;
    }
    const int want = pl->groups < 4 ? pl->groups : 4;
    if (NB >= 3 && NA >= (want < 2 ? 2 : want)) break;
    if (epi_warps == 4 && NB >= 3 && NA >= 2) break;
  }
  if (epi_warps < 4) epi_warps = 4;
  if (NB < 3 || NA < 2) return false;
  if (NA > 4) NA = 4;
  pl->epi_warps = epi_warps;
  pl->NA = NA;
  pl->NT = NA;
  pl->NB = NB;
  pl->acc_stride = fused ? 2 * N : N;
  pl->acc_stages = 2 * pl->acc_stride <= 512 ? 2 : 1;
  int cols = pl->acc_stages * pl->acc_stride;
  int pw = 32;
  while (pw < cols) pw <<= 1;
  pl->tmem_cols = pw;
  if (pw > 512) return false;
  pl->smem = (size_t)NA * pl->a_stage + (size_t)NB * pl->unit_bytes + a2 + misc;
  return true;
}

int launch_stream(SParams& p, const StreamPlan& pl, int fused, cudaStream_t st, bool pair = false, bool tma = false) {
  p.N = pl.N; p.groups = pl.groups; p.tpu = pl.tpu; p.upg = pl.upg; p.gpu1 = pl.gpu1;
  p.slab_rows = pl.slab_rows; p.rpp = pl.rpp; p.NA = pl.NA; p.NB = pl.NB; p.NT = pl.NT;
  p.acc_stages = pl.acc_stages; p.acc_stride = pl.acc_stride; p.epi_warps = pl.epi_warps; p.a_stage = pl.a_stage; p.unit_bytes = pl.unit_bytes;
  p.tap_bytes = pl.tap_bytes; p.tmem_cols = pl.tmem_cols; p.plane_bytes = pl.plane_bytes;
  p.NX = pl.NX; p.x_rows = pl.x_rows; p.x_nbox = pl.x_nbox; p.x_gpb = pl.x_gpb; p.x_unit = pl.x_unit; p.x_bytes = pl.x_bytes;
  p.tiles_per_item = (p.T_out + BM - 1) / BM;
  const long long per_nt = (long long)p.tiles_per_item * p.B;
  const long long total = per_nt * pl.n_tiles;
  if (total > 2147483647ll) return bc::fail(BC_EINVAL, "conv(stream): too many tiles");
  p.tiles_per_nt = (int)per_nt;
  p.n_tiles = pl.n_tiles;
  p.total_tiles = (int)total;
  p.pairs_per_nt = (int)((per_nt + 1) / 2);
  p.total_pairs = p.pairs_per_nt * pl.n_tiles;
  p.idesc = pair ? idesc_bf16_m256(pl.N) : idesc_bf16_m128(pl.N);
  p.trace = g_stream_trace;
#ifdef BC_TRACE
  { const char* e = getenv("BC_STREAM_BSHIFT"); p.dbg_bshift = e ? (atoi(e) & 15) : 0; }
  { const char* e = getenv("BC_STREAM_SKIP"); p.dbg_skip = e ? atoi(e) : 0; }
#endif
  void (*kern)(const SParams) = nullptr;
  int slot = 0;
  if (!pair) {
    if (pl.split == 1 && !fused) { kern = conv_stream_kernel<1, false>; slot = 0; }
    if (pl.split == 2 && !fused) { kern = conv_stream_kernel<2, false>; slot = 1; }
    if (pl.split == 1 && fused) { kern = conv_stream_kernel<1, true>; slot = 2; }
    if (pl.split == 2 && fused) { kern = conv_stream_kernel<2, true>; slot = 3; }
  } else {
    if (pl.split == 1 && !fused) { kern = conv_stream_pair_kernel<1, false>; slot = 4; }
    if (pl.split == 2 && !fused) { kern = conv_stream_pair_kernel<2, false>; slot = 5; }
    if (pl.split == 1 && fused) { kern = conv_stream_pair_kernel<1, true>; slot = 6; }
    if (pl.split == 2 && fused) { kern = conv_stream_pair_kernel<2, true>; slot = 7; }
  }
  static bool configured[64][16] = {{false}};
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 0 || dev >= 64 || !configured[dev][slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(conv_stream)");
    if (dev >= 0 && dev < 64) configured[dev][slot] = true;
  }
  int grid;
  if (pair) {
    const int max_pairs = sms / 2;
    grid = 2 * (p.total_pairs < max_pairs ? p.total_pairs : max_pairs);
  } else {
    grid = p.total_tiles < sms ? p.total_tiles : sms;
  }
  if (tma) {
    // x boxes: 16 channels x x_rows rows (dense, 64-byte rows); y tiles: 32 channels x 32 rows with the 128-byte swizzle
    CUtensorMap tmx, tmy;
    int rc = encode_cl_map(&tmx, p.x, p.B, p.T_in, p.C_in, 16 * p.x_gpb, p.x_rows > 0 ? p.x_rows : 1, false);
    if (rc != BC_OK) return rc;
    rc = encode_cl_map(&tmy, p.y, p.B, p.T_out, p.C_out, 32, 32, true);
    if (rc != BC_OK) return rc;
    const bool xl = pl.NX > 0;
    void (*tk)(const SParams, const CUtensorMap, const CUtensorMap) =
        xl ? (pair ? (pl.split == 2 ? conv_stream_tma_pair_kernel<2, 2> : conv_stream_tma_pair_kernel<1, 2>)
                   : (pl.split == 2 ? conv_stream_tma_kernel<2, 2> : conv_stream_tma_kernel<1, 2>))
           : (pair ? (pl.split == 2 ? conv_stream_tma_pair_kernel<2, 1> : conv_stream_tma_pair_kernel<1, 1>)
                   : (pl.split == 2 ? conv_stream_tma_kernel<2, 1> : conv_stream_tma_kernel<1, 1>));
    const int tslot = 8 + (xl ? 4 : 0) + (pair ? 2 : 0) + (pl.split == 2 ? 1 : 0);
    if (dev < 0 || dev >= 64 || !configured[dev][tslot]) {
      cudaError_t e = cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return bc::cuda_check(e, "cudaFuncSetAttribute(conv_stream_tma)");
      if (dev >= 0 && dev < 64) configured[dev][tslot] = true;
    }
    tk<<<grid, Roles<false>::THREADS, pl.smem, st>>>(p, tmx, tmy);
    BC_LAUNCH_CHECK(pair ? "conv_stream_tma_pair_kernel" : "conv_stream_tma_kernel");
    return BC_OK;
  }
  kern<<<grid, fused ? Roles<true>::THREADS : Roles<false>::THREADS, pl.smem, st>>>(p);
  BC_LAUNCH_CHECK(pair ? "conv_stream_pair_kernel" : "conv_stream_kernel");
  return BC_OK;
}

}  // namespace

// Geometry of the streamed-weight kernel: BC_OK when (C_in, C_out, K, stride, dilation) has a plan; *n_tile is
// the N of the weight image [C_out/N][C_in/16][K][split][2][N][8].
extern "C" int bc_stream_plan(int C_in, int C_out, int K, int stride, int dilation, int precision, int fused, int* n_tile) {
  if (!n_tile) return bc::fail(BC_EINVAL, "stream_plan: null output");
  StreamPlan pl;
  if (!stream_plan(C_in, C_out, K, stride, dilation, precision, fused, &pl))
    return bc::fail(BC_EUNSUPPORTED, "stream_plan: C_in=%d C_out=%d K=%d stride=%d dil=%d fused=%d has no streamed-weight plan", C_in,
                    C_out, K, stride, dilation, fused);
  *n_tile = pl.N;
  return BC_OK;
}

static int conv1d_stream_impl(const float* x, const void* w_image, const float* bias, const float* snake_a,
                              const float* snake_ib, const float* res, float* y, int B, int T_in, int C_in, int T_out,
                              int C_out, int K, int stride, int dilation, int pad_left, int flags, int precision,
                              bc_stream_t s, bool pair) {
  BC_REQUIRE(x && w_image && y, "conv1d(stream): null pointer");
  BC_REQUIRE(B > 0 && T_in > 0 && T_out > 0, "conv1d(stream): bad shape B=%d T_in=%d T_out=%d", B, T_in, T_out);
  BC_REQUIRE(!(flags & BC_CONV_SNAKE_IN) || (snake_a && snake_ib), "conv1d(stream): BC_CONV_SNAKE_IN needs snake_a and snake_ib");
  StreamPlan pl;
  // TMA form (x boxes and y tiles by tensor map) unless the epilogue needs what only the store-through-registers form has
  const bool tma = bc::policy().stream_tma && !res && !(flags & BC_CONV_TANH_OUT) &&
                   stream_plan(C_in, C_out, K, stride, dilation, precision, 0, &pl, pair, true);
  if (!tma && !stream_plan(C_in, C_out, K, stride, dilation, precision, 0, &pl, pair))
    return bc::fail(BC_EUNSUPPORTED, "conv1d(stream): unsupported geometry C_in=%d C_out=%d K=%d stride=%d dil=%d", C_in, C_out, K, stride, dilation);
  BC_REQUIRE(bc::aligned16(x) && bc::aligned16(w_image) && bc::aligned16(y) && (!res || bc::aligned16(res)) &&
                 (!bias || bc::aligned16(bias)) && (!snake_a || bc::aligned16(snake_a)) && (!snake_ib || bc::aligned16(snake_ib)),
             "conv1d(stream): pointers must be 16-byte aligned");
  SParams p{};
  p.x = x; p.y = y; p.res = res; p.w7 = reinterpret_cast<const uint8_t*>(w_image); p.w1 = nullptr;
  p.bias = bias; p.bias2 = nullptr; p.sa1 = snake_a; p.sib1 = snake_ib; p.sa2 = nullptr; p.sib2 = nullptr;
  p.B = B; p.T_in = T_in; p.T_out = T_out; p.C_in = C_in; p.C_out = C_out; p.K = K; p.stride = stride; p.dil = dilation;
  p.pad_left = pad_left; p.flags = flags; p.nt_shift = 0x7fffffff;
  return launch_stream(p, pl, 0, (cudaStream_t)s, pair, tma);
}

// 1 when the CTA-pair form has a plan for this geometry (same n_tile as bc_stream_plan) and the policy enables it
extern "C" int bc_stream_pair_ok(int C_in, int C_out, int K, int stride, int dilation, int precision, int fused) {
  if (!bc::policy().stream_pair) return 0;
  // Where the pair form pays (B200, 8 x 30 s clips, split precision; single CTA in brackets): the kernels that are bound by
  // shared-memory bandwidth -- fused ResidualUnits C = 128: 300 us [327], C = 256: 220 [225]; the 512 -> 512 k = 3 conv:
  // 451 [513].  The strided down-sampling convs are bound by their producers (fp32 load + SnakeBeta + split per element,
  // few MMAs per tile) and only pay for the pair's lock-step: 32->64 302 [274], 64->128 296 [252], 128->256 224 [201],
  // 256->512 219 [197]; the K = 1 input projection is neutral (1072 [1068]).
  if (!fused && !(stride == 1 && K >= 3 && C_in >= 256)) return 0;
  StreamPlan a, b;
  if (!stream_plan(C_in, C_out, K, stride, dilation, precision, fused, &a, false)) return 0;
  if (!stream_plan(C_in, C_out, K, stride, dilation, precision, fused, &b, true)) return 0;
  return a.N == b.N ? 1 : 0;
}

extern "C" int bc_conv1d_stream_fwd(const float* x, const void* w_image, const float* bias, const float* snake_a,
                                    const float* snake_ib, const float* res, float* y, int B, int T_in, int C_in, int T_out,
                                    int C_out, int K, int stride, int dilation, int pad_left, int flags, int precision,
                                    bc_stream_t s) {
  return conv1d_stream_impl(x, w_image, bias, snake_a, snake_ib, res, y, B, T_in, C_in, T_out, C_out, K, stride, dilation, pad_left,
                            flags, precision, s, false);
}

// CTA-pair form (tcgen05 cta_group::2): same arithmetic, w_image = the PAIR image (bc_stream_pair_image_layout)
extern "C" int bc_conv1d_stream_pair_fwd(const float* x, const void* w_image, const float* bias, const float* snake_a,
                                         const float* snake_ib, const float* res, float* y, int B, int T_in, int C_in, int T_out,
                                         int C_out, int K, int stride, int dilation, int pad_left, int flags, int precision,
                                         bc_stream_t s) {
  return conv1d_stream_impl(x, w_image, bias, snake_a, snake_ib, res, y, B, T_in, C_in, T_out, C_out, K, stride, dilation, pad_left,
                            flags, precision, s, true);
}

// Transposed conv (k = 2*stride, T_out = T_in*stride; vq/module.py:67-72,119-136) as ONE launch of the streamed-weight
// kernel: output phase ph of row m is  W[j0+stride]^T x[m+q-1] + W[j0]^T x[m+q]  with j0 = (ph+padding) % stride,
// q = (ph+padding) / stride, so all phases together are a 2-tap conv with stride*C_out output channels whose output
// [B][T_in][stride*C_out] IS y [B][T_in*stride][C_out]; the n-tiles of the q = 1 phases read their taps one row later.
// w_image: pack_stream_weight of [2][C_in][stride*C_out] (column ph*C_out + co = phase filter ph), bias_tiled:
// [stride*C_out].  Needs C_out % n_tile == 0 (every n-tile inside one phase); other geometries use the K = 3 zero-padded
// form through bc_conv1d_stream_fwd or the per-phase path of bc_convtr1d_fwd.
static int convtr1d_stream_impl(const float* x, const void* w_image, const float* bias_tiled, const float* snake_a,
                                const float* snake_ib, float* y, int B, int T_in, int C_in, int C_out, int stride,
                                int padding, int flags, int precision, bc_stream_t s, bool pair) {
  BC_REQUIRE(x && w_image && y, "convtr1d(stream): null pointer");
  BC_REQUIRE(B > 0 && T_in > 0 && stride >= 2 && padding >= 0 && padding < stride, "convtr1d(stream): bad shape B=%d T_in=%d stride=%d padding=%d", B, T_in, stride, padding);
  BC_REQUIRE(!(flags & BC_CONV_SNAKE_IN) || (snake_a && snake_ib), "convtr1d(stream): BC_CONV_SNAKE_IN needs snake_a and snake_ib");
  StreamPlan pl;
  const bool tma = bc::policy().stream_tma && !(flags & BC_CONV_TANH_OUT) && stream_plan(C_in, stride * C_out, 2, 1, 1, precision, 0, &pl, pair, true);
  if ((!tma && !stream_plan(C_in, stride * C_out, 2, 1, 1, precision, 0, &pl, pair)) || C_out % pl.N != 0)
    return bc::fail(BC_EUNSUPPORTED, "convtr1d(stream): C_in=%d C_out=%d stride=%d has no single-launch plan", C_in, C_out, stride);
  BC_REQUIRE(bc::aligned16(x) && bc::aligned16(w_image) && bc::aligned16(y) && (!bias_tiled || bc::aligned16(bias_tiled)) &&
                 (!snake_a || bc::aligned16(snake_a)) && (!snake_ib || bc::aligned16(snake_ib)),
             "convtr1d(stream): pointers must be 16-byte aligned");
  SParams p{};
  p.x = x; p.y = y; p.res = nullptr; p.w7 = reinterpret_cast<const uint8_t*>(w_image); p.w1 = nullptr;
  p.bias = bias_tiled; p.bias2 = nullptr; p.sa1 = snake_a; p.sib1 = snake_ib; p.sa2 = nullptr; p.sib2 = nullptr;
  p.B = B; p.T_in = T_in; p.T_out = T_in; p.C_in = C_in; p.C_out = stride * C_out; p.K = 2; p.stride = 1; p.dil = 1;
  p.pad_left = 1; p.flags = flags;
  p.nt_shift = (stride - padding) * (C_out / pl.N);          // first n-tile of phase ph = stride - padding, the first with q = 1
  return launch_stream(p, pl, 0, (cudaStream_t)s, pair, tma);
}

extern "C" int bc_convtr1d_stream_fwd(const float* x, const void* w_image, const float* bias_tiled, const float* snake_a,
                                      const float* snake_ib, float* y, int B, int T_in, int C_in, int C_out, int stride,
                                      int padding, int flags, int precision, bc_stream_t s) {
  return convtr1d_stream_impl(x, w_image, bias_tiled, snake_a, snake_ib, y, B, T_in, C_in, C_out, stride, padding, flags, precision, s, false);
}
// CTA-pair form: w_image = the pair image of the same [2][C_in][stride*C_out] filter
extern "C" int bc_convtr1d_stream_pair_fwd(const float* x, const void* w_image, const float* bias_tiled, const float* snake_a,
                                           const float* snake_ib, float* y, int B, int T_in, int C_in, int C_out, int stride,
                                           int padding, int flags, int precision, bc_stream_t s) {
  return convtr1d_stream_impl(x, w_image, bias_tiled, snake_a, snake_ib, y, B, T_in, C_in, C_out, stride, padding, flags, precision, s, true);
}

static int resunit_stream_impl(const float* x, const void* w7_image, const float* b7, const float* snake1_a,
                               const float* snake1_ib, const void* w1_image, const float* b1, const float* snake2_a,
                               const float* snake2_ib, float* y, int B, int T, int C, int K, int dilation, int pad_left,
                               int precision, bc_stream_t s, bool pair) {
  BC_REQUIRE(x && w7_image && b7 && snake1_a && snake1_ib && w1_image && b1 && snake2_a && snake2_ib && y, "resunit(stream): null pointer");
  BC_REQUIRE(B > 0 && T > 0 && C > 0 && K > 0 && dilation > 0, "resunit(stream): bad shape B=%d T=%d C=%d K=%d", B, T, C, K);
  BC_REQUIRE(x != y, "resunit(stream): cannot run in place (neighbouring tiles read the input halo)");
  StreamPlan pl;
  if (!stream_plan(C, C, K, 1, dilation, precision, 1, &pl, pair))
    return bc::fail(BC_EUNSUPPORTED, "resunit(stream): C=%d K=%d dil=%d has no streamed-weight plan", C, K, dilation);
  BC_REQUIRE(bc::aligned16(x) && bc::aligned16(w7_image) && bc::aligned16(w1_image) && bc::aligned16(y) && bc::aligned16(b7) &&
                 bc::aligned16(b1) && bc::aligned16(snake1_a) && bc::aligned16(snake1_ib) && bc::aligned16(snake2_a) && bc::aligned16(snake2_ib),
             "resunit(stream): pointers must be 16-byte aligned");
  SParams p{};
  p.x = x; p.y = y; p.res = x; p.w7 = reinterpret_cast<const uint8_t*>(w7_image); p.w1 = reinterpret_cast<const uint8_t*>(w1_image);
  p.bias = b7; p.bias2 = b1; p.sa1 = snake1_a; p.sib1 = snake1_ib; p.sa2 = snake2_a; p.sib2 = snake2_ib;
  p.B = B; p.T_in = T; p.T_out = T; p.C_in = C; p.C_out = C; p.K = K; p.stride = 1; p.dil = dilation;
  p.pad_left = pad_left; p.flags = BC_CONV_SNAKE_IN; p.nt_shift = 0x7fffffff;
  return launch_stream(p, pl, 1, (cudaStream_t)s, pair);
}

extern "C" int bc_resunit_stream_fwd(const float* x, const void* w7_image, const float* b7, const float* snake1_a,
                                     const float* snake1_ib, const void* w1_image, const float* b1, const float* snake2_a,
                                     const float* snake2_ib, float* y, int B, int T, int C, int K, int dilation, int pad_left,
                                     int precision, bc_stream_t s) {
  return resunit_stream_impl(x, w7_image, b7, snake1_a, snake1_ib, w1_image, b1, snake2_a, snake2_ib, y, B, T, C, K, dilation,
                             pad_left, precision, s, false);
}

extern "C" int bc_resunit_stream_pair_fwd(const float* x, const void* w7_image, const float* b7, const float* snake1_a,
                                          const float* snake1_ib, const void* w1_image, const float* b1, const float* snake2_a,
                                          const float* snake2_ib, float* y, int B, int T, int C, int K, int dilation, int pad_left,
                                          int precision, bc_stream_t s) {
  return resunit_stream_impl(x, w7_image, b7, snake1_a, snake1_ib, w1_image, b1, snake2_a, snake2_ib, y, B, T, C, K, dilation,
                             pad_left, precision, s, true);
}

// debug hook (not part of the product path): device buffer of 64*16 int64 (zeroed by the caller) that receives
// clock64 stamps / wait totals of CTA 0 of the streamed-weight kernel
extern "C" int bc_debug_set_stream_trace(void* device_buffer) {
  g_stream_trace = reinterpret_cast<long long*>(device_buffer);
  return BC_OK;
}
