/*
 * bigcodec_b200.h -- C ABI of the B200-native BigCodec audio-tokenizer hot path.
 *
 * The reference (hoyso48/AudioTokenization, BigCodec_SSL/) has no FFI layer: its
 * boundary is the Python nn.Module API of the `vq/` package, and every numeric
 * step below is a PyTorch library call there.  Each entry point names the
 * reference call site(s) it replaces (paths relative to BigCodec_SSL/).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, ints, an opaque stream handle (cudaStream_t).
 *     No torch / C++ types cross this boundary.
 *   - activations are float32, CHANNELS-LAST: x[b][t][c]  (the reference uses
 *     [b][c][t]; bc_transpose_* converts at the module boundary).
 *   - weights are pre-folded (weight_norm g*v/||v|| is done once on the host,
 *     SURVEY.md App. C) and packed tap-major: w[k][c_in][c_out].
 *   - every function returns 0 on success, a negative BC_E* code on bad
 *     arguments / unsupported configuration, or a positive cudaError_t value.
 *     Nothing aborts; bc_last_error() returns a thread-local message.
 *   - re-entrant, asynchronous on `stream`, no hidden synchronisation, no
 *     allocation (callers pass workspaces), no global mutable state.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns an error.
 */
#ifndef BIGCODEC_B200_H
#define BIGCODEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BC_ABI_VERSION 4

typedef void* bc_stream_t; /* cudaStream_t */

enum {
  BC_OK = 0,
  BC_EINVAL = -1,       /* bad shape / null pointer / misaligned pointer */
  BC_EUNSUPPORTED = -2, /* valid in the reference, not implemented here (e.g. bidirectional LSTM) */
  BC_ENODEVICE = -3     /* no CUDA device / wrong architecture */
};

/* flags for bc_conv1d_fwd */
enum {
  BC_CONV_SNAKE_IN = 1, /* apply SnakeBeta to the input while staging it (needs snake_a, snake_ib) */
  BC_CONV_TANH_OUT = 2, /* y = tanh(y) after bias/residual (decoder tail, vq/codec_decoder.py:78) */
  BC_CONV_ACCUM_OUT = 4 /* y += result (internal use) */
};

/* arithmetic mode of the dense contractions */
enum {
  BC_PREC_FP32 = 0,   /* CUDA-core FFMA, exact float32 (parity mode) */
  BC_PREC_BF16 = 1,   /* tcgen05 bf16 x bf16 -> fp32, single pass (fast mode) */
  BC_PREC_BF16X3 = 2  /* tcgen05, hi/lo split: hi*hi + hi*lo + lo*hi (fp32-class accuracy on tensor cores) */
};

int bc_abi_version(void);
const char* bc_last_error(void);
/* Kernel-selection knobs this process runs with (read from the environment once, at first use), as a
 * "key=value ..." string: recorded by bench.py next to every number.  None of them changes results
 * (same arithmetic per element; the forms they select are pinned against each other by the tests):
 *   BC_RU_GROUP / BC_RU_PERSIST / BC_RU_PAIR = 0   switch off the warpgroup-per-tile / role-pipeline / CTA-pair fused units
 *   BC_STREAM_PAIR = 0     streamed-weight kernel without the CTA-pair (cta_group::2) form
 *   BC_STREAM_TMA = 0|1|2  plain streamed convs: no tensor maps | y by TMA stores (default) | x boxes by TMA as well
 *   BC_LSTM_PINGPONG = 0   one batch tile per CTA only (max batch 256);  BC_LSTM_PAIR = 1  CTA-pair recurrence kernel
 *   BC_LSTM_COMPACT = 0    full 128-row h exchange image also for batches below 128 rows
 *   BC_TC_VARIANT / BC_TC_PERSIST   per-tile tcgen05 conv variants (development) */
int bc_policy(char* buf, size_t n);
/* 0 if device `dev` exists and is sm_100; fills optional outputs. */
int bc_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---- layout ------------------------------------------------------------- */
/* [B,C,T] -> [B,T,C] and back (module boundary: the reference's tensors are [B,C,T],
 * vq/codec_encoder.py:62-64, vq/codec_decoder.py:85-94). */
int bc_transpose_bct_to_btc(const float* x, float* y, int B, int C, int T, bc_stream_t s);
int bc_transpose_btc_to_bct(const float* x, float* y, int B, int T, int C, bc_stream_t s);

/* ---- SnakeBeta / anti-aliased Activation1d ------------------------------ */
/* Replaces Activation1d.forward (vq/alias_free_torch/act.py:25-32) wrapping
 * SnakeBeta(alpha_logscale=True).forward (vq/activations.py:107-119):
 *   y = x + snake_ib[c] * sin(x * snake_a[c])^2,  snake_a = exp(alpha), snake_ib = 1/(exp(beta)+1e-9).
 * antialias != 0 additionally applies UpSample1d (resample.py:25-33) before and
 * DownSample1d/LowPassFilter1d (resample.py:47-49, filter.py:86-95) after, fused in
 * one pass; fir12 = the 12 Kaiser-sinc taps (filter.py:28-57), a DEVICE pointer.
 * x, y: [B,T,C] channels-last; y may not alias x when antialias != 0. */
int bc_snake_fwd(const float* x, float* y, const float* snake_a, const float* snake_ib,
                 const float* fir12, int B, int T, int C, int antialias, bc_stream_t s);

/* ---- dense convolutions -------------------------------------------------- */
/* Replaces nn.Conv1d under WNConv1d / CausalConv1d (vq/module.py:11-48,59-65) and, with
 * the phase decomposition of bc_convtr1d_fwd, nn.ConvTranspose1d (vq/module.py:50-57,67-72).
 *
 *   y[b][t*y_tstride + y_toffset][co] = bias[co]
 *        + sum_{k<K} sum_{ci<C_in} w[k][ci][co] * act(x[b][t*stride + k*dilation - pad_left][ci])
 *        (+ res[b][same row][co])                       for t in [0, T_out)
 * rows outside [0, T_in) read as zero AFTER the activation (the reference zero-pads the
 * activation's output).  act = SnakeBeta if BC_CONV_SNAKE_IN else identity.
 * y has y_rows rows per batch item (y_rows >= (T_out-1)*y_tstride + y_toffset + 1).
 * res (optional) has the same geometry as y.  bias may be NULL. */
int bc_conv1d_fwd(const float* x, const float* w, const float* bias,
                  const float* snake_a, const float* snake_ib, const float* res, float* y,
                  int B, int T_in, int C_in, int T_out, int C_out, int K, int stride, int dilation,
                  int pad_left, int y_rows, int y_tstride, int y_toffset, int flags, int precision,
                  bc_stream_t s);

/* Tensor-core tiling of a conv geometry: n_tile output channels per CTA, gpc 16-channel groups
 * per staged chunk, nchunks = C_in / (16*gpc).  For precision BC_PREC_BF16 / BC_PREC_BF16X3 the
 * `w` argument of bc_conv1d_fwd is NOT the fp32 [K][C_in][C_out] array but its bf16 image
 *   [C_out/n_tile][nchunks][split][K][gpc][2][n_tile][8],  split = 1 (bf16) or 2 (hi, lo = w - hi),
 * input channel = chunk*16*gpc + g*16 + h*8 + e.  Returns BC_EUNSUPPORTED when the geometry has no
 * tensor-core tiling (C_in or C_out not a multiple of 16: the 1->C and C->1 edge convs, which stay
 * on the fp32 kernel). */
int bc_tc_plan(int C_in, int C_out, int K, int stride, int dilation, int precision,
               int* n_tile, int* gpc, int* nchunks);

/* Fused ResidualUnit (vq/module.py:74-89) for the tensor-core modes, one kernel, one HBM read and one
 * write of the activation:
 *   y = x + W1 * snake2( W7 (*) snake1(x) + b7 ) + b1
 * W7: dilated K-tap "same" conv C->C (image per bc_tc_plan(C, C, K, 1, dilation)), W1: 1x1 conv C->C
 * (image per bc_tc_plan with n_tile = C, gpc = C/16, nchunks = 1).  The K-tap accumulator stays in TMEM,
 * is activated and re-quantised to bf16 in shared memory and feeds the second MMA chain directly.
 * Needs C % 16 == 0 and C <= 128 (one accumulator tile); otherwise BC_EUNSUPPORTED (callers chain two
 * bc_conv1d_fwd calls).  pad_left = dilation*(K-1)/2 (or dilation*(K-1) for the causal variant). */
/* Weight-image geometry bc_resunit_fwd expects for W7 (same meaning as bc_tc_plan; W1 is always
 * n_tile = C, gpc = C/16, nchunks = 1).  *persistent = 1 when the persistent warp-specialised kernel
 * (weights resident in shared memory, C in {16,32,64}) will run, 2 when its warpgroup-per-tile form does
 * (same weight images; csrc/ru_group.cu), 0 for the per-tile kernel.
 * *persistent = 3 / 4: the CTA-pair kernel (tcgen05 cta_group::2, csrc/ru_pair.cu; C = 64, split precision) runs and
 * w7 / w1 must be its PER-RANK images, two consecutive blocks (rank 0, rank 1) of equal size.  A block of rows R of a
 * bf16 matrix m[k][c_in][n] is laid out [k][c_in/16][2 k-planes][|R|][8]  (channel = g*16 + h*8 + e).  With hi = bf16(w),
 * lo = bf16(w - hi) and half(r) = rows [r*C/2, (r+1)*C/2):
 *   3 (plain)    w7 rank r = block(hi, half(r)) | block(lo, half(r))
 *   4 (stacked)  w7 rank r = block(r == 0 ? hi : lo, all rows) | block(hi, half(r))
 *   w1 (both)    rank r = block(hi, half(r)) | block(lo, half(r)) */
int bc_resunit_plan(int C, int K, int dilation, int precision, int* n_tile, int* gpc, int* nchunks, int* persistent);
int bc_resunit_fwd(const float* x, const float* w7, const float* b7, const float* snake1_a, const float* snake1_ib,
                   const float* w1, const float* b1, const float* snake2_a, const float* snake2_ib, float* y,
                   int B, int T, int C, int K, int dilation, int pad_left, int precision, bc_stream_t s);

/* Streamed-weight persistent kernels for the wide layers (tensor-core modes): one CTA per SM walks the
 * (n-tile, item, 128-step) tiles while the weights flow from L2 through a shared-memory ring in
 * (16-channel group, tap) units; activation staging, MMA issue, the ResidualUnit's middle activation and
 * the output stores run on separate warps (csrc/conv_stream.cu).  Same arithmetic and arguments as
 * bc_conv1d_fwd (y_rows = T_out, y_tstride = 1, y_toffset = 0) / bc_resunit_fwd; the weight image is
 *   [C_out/n_tile][C_in/16][K][split][2][n_tile][8] bf16,  input channel = g*16 + h*8 + e,
 * n_tile from bc_stream_plan (BC_EUNSUPPORTED: no plan, use bc_conv1d_fwd / bc_resunit_fwd).
 * `fused` = 1 asks for the ResidualUnit plan (W7 image as above, W1 image with K = 1). */
int bc_stream_plan(int C_in, int C_out, int K, int stride, int dilation, int precision, int fused, int* n_tile);
int bc_conv1d_stream_fwd(const float* x, const void* w_image, const float* bias, const float* snake_a,
                         const float* snake_ib, const float* res, float* y, int B, int T_in, int C_in, int T_out,
                         int C_out, int K, int stride, int dilation, int pad_left, int flags, int precision,
                         bc_stream_t s);
int bc_resunit_stream_fwd(const float* x, const void* w7_image, const float* b7, const float* snake1_a,
                          const float* snake1_ib, const void* w1_image, const float* b1, const float* snake2_a,
                          const float* snake2_ib, float* y, int B, int T, int C, int K, int dilation, int pad_left,
                          int precision, bc_stream_t s);

/* CTA-pair form of the two entry points above (tcgen05 cta_group::2: one MMA of M = 256 over two 128-step tiles of a
 * cluster of two CTAs, each SM holding HALF of every weight unit -- half the B-operand fetch and half the weight stream
 * per SM, the two things that bound the wide layers).  Same arguments and arithmetic; the weight
 * image is the PAIR image [C_out/n_tile][2 ranks][C_in/16][K][split][2][n_tile/2][8]: rank r holds rows
 * [r*n_tile/2, (r+1)*n_tile/2) of every k-plane.  bc_stream_pair_ok = 1 when the geometry has a pair plan (and
 * BC_STREAM_PAIR != 0). */
int bc_stream_pair_ok(int C_in, int C_out, int K, int stride, int dilation, int precision, int fused);
int bc_conv1d_stream_pair_fwd(const float* x, const void* w_image, const float* bias, const float* snake_a,
                              const float* snake_ib, const float* res, float* y, int B, int T_in, int C_in, int T_out,
                              int C_out, int K, int stride, int dilation, int pad_left, int flags, int precision,
                              bc_stream_t s);
int bc_resunit_stream_pair_fwd(const float* x, const void* w7_image, const float* b7, const float* snake1_a,
                               const float* snake1_ib, const void* w1_image, const float* b1, const float* snake2_a,
                               const float* snake2_ib, float* y, int B, int T, int C, int K, int dilation, int pad_left,
                               int precision, bc_stream_t s);

/* Transposed conv as `stride` output phases of 2-tap convs:
 * w_phases[phase][2][C_in][C_out] (host-packed from the folded [C_in,C_out,2*stride]
 * weight: tap0 = W[:,:,j0+stride], tap1 = W[:,:,j0], j0 = (phase+padding) % stride).
 * T_out = T_in*stride for both the padded (padding = ceil(stride/2), output_padding =
 * stride%2; vq/module.py:119-136) and the causal (padding 0, last `stride` samples
 * dropped; vq/module.py:50-57) variants. */
int bc_convtr1d_fwd(const float* x, const float* w_phases, const float* bias,
                    const float* snake_a, const float* snake_ib, float* y,
                    int B, int T_in, int C_in, int C_out, int stride, int padding, int flags,
                    int precision, bc_stream_t s);

/* The same transposed conv as ONE launch of the streamed-weight kernel (tensor-core modes): all phases together are a
 * 2-tap conv with stride*C_out output channels whose output [B][T_in][stride*C_out] is y [B][T_in*stride][C_out]; n-tiles
 * of the phases with (ph+padding)/stride = 1 read their taps one row later.  w_image = the bc_conv1d_stream_fwd image
 * of [2][C_in][stride*C_out] (column ph*C_out + co = phase filter ph, taps as in w_phases), n_tile from
 * bc_stream_plan(C_in, stride*C_out, 2, 1, 1, ...); bias_tiled [stride*C_out] = bias repeated per phase.  Needs
 * C_out % n_tile == 0, else BC_EUNSUPPORTED (narrower layers: the 3-tap zero-padded form through
 * bc_conv1d_stream_fwd, or bc_convtr1d_fwd). */
int bc_convtr1d_stream_fwd(const float* x, const void* w_image, const float* bias_tiled, const float* snake_a,
                           const float* snake_ib, float* y, int B, int T_in, int C_in, int C_out, int stride,
                           int padding, int flags, int precision, bc_stream_t s);

/* ... and its CTA-pair form (pair image, bc_stream_pair_ok(C_in, stride*C_out, 2, 1, 1, ...)). */
int bc_convtr1d_stream_pair_fwd(const float* x, const void* w_image, const float* bias_tiled, const float* snake_a,
                                const float* snake_ib, float* y, int B, int T_in, int C_in, int C_out, int stride,
                                int padding, int flags, int precision, bc_stream_t s);

/* ---- LSTM ---------------------------------------------------------------- */
/* One uni-directional LSTM layer, recurrent part (nn.LSTM inside ResLSTM,
 * vq/module.py:143-167; gate order i,f,g,o; h0 = c0 = 0).
 *   pre   [B][T][4H]  = W_ih x_t + b_ih + b_hh for all t (computed by bc_conv1d_fwd, K=1)
 *   w_hh_packed       = bc_lstm_pack_whh layout of W_hh [4H][H]
 *   skip  [B][T][H]   optional, added to the output (ResLSTM's `y + x`, last layer only)
 *   y     [B][T][H]
 *   workspace: bc_lstm_workspace_bytes(B,H) bytes, device, need not be zeroed.
 * Launched as ONE cooperative persistent kernel (grid-wide sync per time step).  W_hh rows stay resident
 * in shared memory when every group of 4 hidden units gets its own co-resident CTA (H <= 592 on 148 SMs);
 * wider layers (H = 1536 of the original BigCodec config) stream them from L2 every step. */
size_t bc_lstm_workspace_bytes(int B, int H);
size_t bc_lstm_packed_whh_floats(int H);
/* host-side helper: w_hh [4H][H] row-major (host ptr) -> packed (host ptr) */
int bc_lstm_pack_whh(const float* w_hh, float* packed, int H);
int bc_lstm_recurrent_fwd(const float* pre, const float* w_hh_packed, const float* skip, float* y,
                          void* workspace, int B, int T, int H, bc_stream_t s);

/* Tensor-core recurrence (BC_PREC_BF16 / BC_PREC_BF16X3): W_hh slices resident in shared memory, h
 * exchanged between CTAs as bf16 (hi[, lo]) UMMA images through HBM/L2, per-batch-tile step counters
 * instead of a grid barrier.  w_image = [4H/NS slices][H/16][2][split*NS][8] bf16: per 16-channel group and
 * k-plane the NS rows of the hi slice, then (split = 2) the NS rows of the lo slice (one B operand of 2*NS rows);
 * slice row g*(NS/4)+u = W_hh[g*H + slice*(NS/4) + u], NS = bc_lstm_tc_slice_cols(precision).
 * Batches that do not fit one 128-row tile per CTA run two independent tiles per CTA, interleaved step by step.
 * bc_lstm_tc_max_batch: largest B one launch accepts on the current device (0 = no tensor-core plan for
 * this H; use bc_lstm_recurrent_fwd).  Workspace need not be zeroed.
 * Gate non-linearities in these modes use the SFU exponential and the approximate divide (|error| ~2e-7,
 * below the split-operand error); they saturate to 0 / +-1 for any finite pre-activation or cell state. */
int bc_lstm_tc_slice_cols(int precision);
int bc_lstm_tc_max_batch(int H, int precision);
size_t bc_lstm_tc_workspace_bytes(int B, int H, int precision);
int bc_lstm_tc_recurrent_fwd(const float* pre, const void* w_image, const float* skip, float* y,
                             void* workspace, int B, int T, int H, int precision, bc_stream_t s);
/* The same recurrence over ONE CHUNK of a longer sequence: steps [t_base, t_base + T_chunk).  pre / y / skip point at the
 * chunk's first step; consecutive batch items are pre_rows (resp. y_rows) steps apart.  The hidden state travels in the
 * workspace (same workspace for every chunk of a sequence, chunks in order on one stream), the cell state in
 * c_state [ceil(B/128)*128][H] floats (written by every chunk, read when t_base > 0).  With this the two layers of a ResLSTM
 * run as a wave front on two streams -- layer 1 on chunk c while layer 0 is on chunk c+1 -- which halves the number of
 * sequential steps of a small batch (host side: vq/module.py ResLSTM).  bc_lstm_tc_ctas = CTAs one launch occupies
 * (two launches are co-resident when twice that fits the SM count). */
int bc_lstm_tc_recurrent_chunk_fwd(const float* pre, const void* w_image, const float* skip, float* y,
                                   void* workspace, float* c_state, int B, int T_chunk, int pre_rows, int y_rows,
                                   int t_base, int H, int precision, bc_stream_t s);
int bc_lstm_tc_ctas(int B, int H, int precision);

/* ---- factorized VQ -------------------------------------------------------- */
/* Replaces FactorizedVectorQuantize.forward / decode_latents in eval mode
 * (vq/factorized_vector_quantize.py:29-76,93-109): in_proj -> L2 normalise -> nearest
 * code by cosine (== argmax(-dist), ties -> lowest index) -> int32 index.
 *   z        [N][C] channels-last latents (N = B*T')
 *   w_in     [D][C] folded in_proj weight, b_in [D]   (NULL,NULL => identity, C == D)
 *   cb_norm  [Kc][D] codebook rows L2-normalised with F.normalize's eps (1e-12)
 *   idx      [N] int32 out
 *   margin   [N] optional: top-1 minus top-2 cosine
 *   z_e      [N][D] optional: projected (un-normalised) latents
 * D must be <= 16; Kc >= 2. */
int bc_vq_encode(const float* z, const float* w_in, const float* b_in, const float* cb_norm,
                 int32_t* idx, float* margin, float* z_e, int N, int C, int D, int Kc,
                 bc_stream_t s);

/* Replaces embed_code + out_proj (factorized_vector_quantize.py:72-74,78-91) and one
 * iteration of the ResidualVQ loop (vq/residual_vq.py:27-33):
 *   q[n][c] = b_out[c] + sum_d w_out[c][d] * cb[idx[n]][d]      (cb = RAW codebook rows)
 *   z_q[n][c]       = (accumulate ? z_q[n][c] : 0) + q[n][c]
 *   residual[n][c] -= q[n][c]                                   (if residual != NULL)
 * w_out NULL => identity (C == D).  Out-of-range indices return BC_EINVAL lazily:
 * they are clamped on the device and counted into *bad_count (optional, device int). */
int bc_vq_dequant(const int32_t* idx, const float* cb, const float* w_out, const float* b_out,
                  float* z_q, float* residual, int* bad_count, int N, int C, int D, int Kc,
                  int accumulate, bc_stream_t s);

/* Finite scalar quantisation, the quantizer BigCodecDecoder selects with fsq=True (vq/codec_decoder.py:41-47,87-89;
 * FSQ.forward, vq/vector_quantize_pytorch_lucidrains/finite_scalar_quantization.py:111-148,170-175,203-259), eval mode,
 * one codebook:
 *   z_e = w_in z + b_in ; bounded = tanh(z_e + shift) * half_l - offset ; q = round-half-even(bounded)
 *   codes = q / half_width ; idx = int32(sum_j (codes_j * half_width_j + half_width_j) * basis_j)
 *   z [N][C], w_in [D][C] (+ b_in [D]; NULL, NULL => identity, C == D), params5xd = [5][D] floats:
 *   half_l | offset | shift | half_width | basis (computed by the host module with the reference's own expressions),
 *   idx [N] int32, codes [N][D] optional, boundary [N] optional = distance of the closest `bounded` component to a
 *   rounding boundary (parity tests gate on it).  D = number of levels <= 8.  Dequantisation is bc_vq_dequant with the
 *   implicit codebook (FSQ._indices_to_codes) as `cb`. */
int bc_fsq_encode(const float* z, const float* w_in, const float* b_in, const float* params5xd, int32_t* idx,
                  float* codes, float* boundary, int N, int C, int D, bc_stream_t s);

/* int32 [n_q][N] indices -> int16 [N][n_q], the on-disk layout of extract_indices.py:520-532. */
int bc_indices_to_int16(const int32_t* idx, int16_t* out, int n_q, int N, bc_stream_t s);

/* ---- codebook statistics ---------------------- */
/* Replaces the index bookkeeping of CodebookPerplexity.update / CodebookUtilization.update
 * (lightning_module.py:33-36,62-64: one-hot sum / used-code mask) and the Counter of inference_full.py:570-604:
 *   counts[c] += #{n < N : idx[n] == c}      (uint64, DEVICE, accumulating: zero it before the first call)
 * Indices outside [0, Kc) are counted into *bad_count (optional device int) and ignored. */
int bc_code_histogram(const int32_t* idx, long long N, int Kc, unsigned long long* counts, int* bad_count, bc_stream_t s);
/* out3[0] = entropy in nats of counts/sum(counts) over the used codes (CodebookPerplexity.compute,
 * lightning_module.py:38-51, before the exp), out3[1] = number of used codes (CodebookUtilization.compute,
 * :66-69, before the division), out3[2] = sum(counts).  out3: 3 doubles, DEVICE. */
int bc_code_entropy(const unsigned long long* counts, int Kc, double* out3, bc_stream_t s);

/* ---- peer memory for the long-form hand-off ----------------------------------------------
 * BASELINE configs[3] on several GPUs: the conv front end of ONE recording is sharded by chunk, then the frame-rate
 * features (2 KB per frame) make one ordered hand-off to the GPU that runs the sequential LSTM + VQ.  The reference has
 * no multi-GPU inference path at all; SURVEY prescribes peer copies over NVLink, not a collective.  One process per GPU:
 * the owner allocates the receive buffer (bc_ipc_alloc: plain cudaMalloc, exportable), publishes its 64-byte IPC handle
 * (bc_ipc_export; the handle travels over the host-side process group), peers map it (bc_ipc_open, peer access enabled
 * lazily) and store their rows with bc_peer_copy (cudaMemcpyAsync on the caller's stream).  Completion is the callers'
 * business (stream synchronise + host barrier). */
int bc_ipc_alloc(void** dev_ptr, size_t bytes);
int bc_ipc_free(void* dev_ptr);
int bc_ipc_export(const void* dev_ptr, unsigned char* handle64);
int bc_ipc_open(const unsigned char* handle64, void** peer_ptr);
int bc_ipc_close(void* peer_ptr);
int bc_peer_copy(void* dst, const void* src, size_t bytes, bc_stream_t s);

/* debug only: per-stage clock64 stamps of the persistent ResidualUnit kernel (CTA 0, first 64 tiles) */
int bc_debug_set_ru_trace(void* device_buffer);
/* same for the streamed-weight kernel (events: producer, MMA, MID, store stamps and MMA-warp wait totals) */
int bc_debug_set_stream_trace(void* device_buffer);
/* tensor-core LSTM: [64 steps][8] stamps (counter seen, h copies issued, MMAs issued, gates start, h stored, published) */
int bc_debug_set_lstm_trace(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* BIGCODEC_B200_H */
