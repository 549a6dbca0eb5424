/*
 * icm_b200.h -- C ABI of the B200-native hot path of stm233/image-compression-for-machine.
 *
 * This is the drop-in boundary: plain pointers, sizes and a cudaStream_t (passed as void*); no torch,
 * pybind11 or C++ types.  Each entry point names the reference interface it replaces (paths relative
 * to the reference tree, "ans.so@0x...." = symbol offset inside the reference's shipped
 * compressai/ans.cpython-38-x86_64-linux-gnu.so whose sources are absent).
 *
 * Conventions
 *   - every function returns ICM_OK (0) or a negative ICM_ERR_* code and never throws;
 *     icm_last_error() returns a thread-local message for the last failure;
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers;
 *   - all device work is enqueued on `stream` and is asynchronous unless stated otherwise;
 *   - "stream order" of an image's symbols = the order the reference feeds its coder:
 *       y: slice-major, then (c, h, w) inside a slice   (compressai/models/stf.py:718-719)
 *       z: (c, h, w)                                      (entropy_models.py:227-235)
 */
#ifndef ICM_B200_H
#define ICM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICM_OK 0
#define ICM_ERR_INVALID_ARG (-1)
#define ICM_ERR_CUDA (-2)
#define ICM_ERR_NO_DEVICE (-3)
#define ICM_ERR_CAPACITY (-4)
#define ICM_ERR_BAD_INDEX (-5)
#define ICM_ERR_UNSUPPORTED (-6)

const char *icm_last_error(void);
/* ABI version of this library (bumped on incompatible changes). */
int icm_abi_version(void);
/* Number of CUDA kernels this library has launched in the calling process (bench.py: gpu_launches). */
int64_t icm_launch_count(void);
/* A caller that replays this library's kernels from a CUDA graph it captured (compressai/utils/pipeline.py) reports the
 * number of kernel nodes of each replay here, so that icm_launch_count keeps counting kernels that ran, not API calls. */
int64_t icm_note_graph_launches(int64_t n);

/* ------------------------------------------------------------------------------------------------
 * R5  compressai._CXX.pmf_to_quantized_cdf(list[float] pmf, int precision) -> list[int]
 *     (_CXX.so@0x68c0; called from entropy_models.py:60-63).  HOST function, no GPU needed.
 *     out must hold n+1 entries.  Returns n+1 on success. */
int icm_pmf_to_quantized_cdf(const float *h_pmf, int n, int precision, uint32_t *h_out);

/* ------------------------------------------------------------------------------------------------
 * Quantised-CDF tables resident on the device.  Replaces the (cdfs, cdfs_sizes, offsets) list
 * arguments that every compressai.ans call receives and that pybind11 re-copies per call
 * (entropy_models.py:231-233, stf.py:694-696,727,766). h_cdfs is row-major int32 [n_cdf][stride]. */
typedef struct icm_tables icm_tables;
int icm_tables_create(const int32_t *h_cdfs, int n_cdf, int stride, const int32_t *h_sizes,
                      const int32_t *h_offsets, icm_tables **out);
void icm_tables_destroy(icm_tables *t);

/* ------------------------------------------------------------------------------------------------
 * R1+R2+R3  ans.RansEncoder.encode_with_indexes / BufferedRansEncoder.encode_with_indexes + flush
 *     (ans.so@0x8d70, @0x8a10, @0x8730; third_party/ryg_rans/rans64.h:65-103).
 *     Encodes n_streams independent streams (one 64-bit rANS state each, exactly the reference's
 *     coder).  Stream s covers d_symbols/d_indexes[s*n_per_stream .. +n_per_stream) (stream order).
 *     d_work:   scratch, icm_rans_encode_workspace_bytes(n_streams, n_per_stream) bytes.
 *     d_packed: output, the streams' bytes back to back; capacity `packed_capacity` bytes.
 *     d_sizes:  int32[n_streams+1]: byte length of every stream, then the total; a stream whose
 *               symbols referenced an invalid table index reports ICM_ERR_BAD_INDEX there. */
int64_t icm_rans_encode_workspace_bytes(int n_streams, int64_t n_per_stream);
int icm_rans_encode_batch(const icm_tables *t, const int32_t *d_symbols, const int32_t *d_indexes,
                          int n_streams, int64_t n_per_stream, void *d_work, uint8_t *d_packed,
                          int64_t packed_capacity, int32_t *d_sizes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * R4  ans.RansDecoder.set_stream / decode_stream / decode_with_indexes
 *     (ans.so@0x7c40, @0x7ce0, @0x8060; rans64.h:107-142).
 *     A decoder object holds n_streams {state, read position} pairs on the device, which persist
 *     across decode_step calls like the reference object does across decode_stream calls
 *     (stf.py:751-766: one set_stream, twelve decode_stream calls per image). */
typedef struct icm_rans_decoder icm_rans_decoder;
int icm_rans_decoder_create(int n_streams, icm_rans_decoder **out);
void icm_rans_decoder_destroy(icm_rans_decoder *d);
/* d_bytes: all streams back to back (each length a multiple of 4, 4-byte aligned start);
 * h_offsets/h_sizes: byte offset and length of every stream inside d_bytes. */
int icm_rans_decoder_set_streams(icm_rans_decoder *d, const uint8_t *d_bytes, const int64_t *h_offsets,
                                 const int64_t *h_sizes, void *stream);
/* Same, for streams that are already on the device as icm_rans_encode_batch left them: d_bytes = its
 * d_packed, d_sizes = its int32 byte sizes (streams back to back in order).  No host synchronisation. */
int icm_rans_decoder_set_streams_device(icm_rans_decoder *d, const uint8_t *d_bytes, const int32_t *d_sizes, void *stream);
/* Streams (= warps) per decoder CTA: 0 = automatic, else 1, 2, 4, 8 or 16.  They share one copy of the tables in
 * shared memory: one stream per CTA is fastest per stream, many streams per CTA leave the other SMs to the
 * convolutions.  kernel = 1 forces the round-1 warp-search kernel (the fallback for table sets that do not fit the
 * bucket image; A/B tests).  Thread-local.  See csrc/rans.cu. */
int icm_set_decoder_layout(int streams_per_cta, int kernel);
int icm_set_decoder_streams_per_cta(int n); /* = icm_set_decoder_layout(n, current kernel) */
/* Decodes the next n_per_stream symbols of every stream.  d_indexes / d_out: [n_streams][n_per_stream]. */
int icm_rans_decoder_step(icm_rans_decoder *d, const icm_tables *t, const int32_t *d_indexes,
                          int64_t n_per_stream, int32_t *d_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Strided view of a [B, C, P] tensor (P = H*W pixels).  NCHW: sb=C*P, sc=P, sp=1;
 * NHWC with row pitch `pitch`: sb=P*pitch, sc=1, sp=pitch.  Strides are in elements. */
typedef struct {
    void *ptr;
    int64_t sb, sc, sp;
} icm_view;

/* E1+E3 (+E2)  GaussianConditional.build_indexes + quantize(..., "symbols", means) and ŷ = q + means
 *     (entropy_models.py:661-666, :126-150; stf.py:714-716).  fp32 in; int32 symbols / indexes written
 *     in stream order at d_symbols[b*stream_stride + stream_offset + c*P + p].  Optional outputs
 *     (ptr may be NULL): y_hat fp32 view and up to two bf16 views (the conv stacks' support buffers).
 *     mu may be absent (means=None); scale and d_indexes may both be absent (plain quantize). */
int icm_gc_quantize_index(icm_view y, icm_view mu, icm_view scale, int B, int C, int64_t P,
                          const float *d_scale_table, int n_levels, float scale_bound,
                          int32_t *d_symbols, int32_t *d_indexes, int64_t stream_stride, int64_t stream_offset,
                          icm_view y_hat_f32, icm_view y_hat_bf16_a, icm_view y_hat_bf16_b, void *stream);

/* E3 alone, for the decoder: indexes in stream order from a scale view (stf.py:764). */
int icm_gc_build_indexes(icm_view scale, int B, int C, int64_t P, const float *d_scale_table, int n_levels,
                         float scale_bound, int32_t *d_indexes, int64_t stream_stride, int64_t stream_offset,
                         void *stream);

/* E2  EntropyModel.dequantize(symbols, means) (entropy_models.py:159-165; stf.py:767-768): symbols in
 *     stream order -> ŷ = float(q) + mu, same optional outputs as above. */
int icm_gc_dequantize(const int32_t *d_symbols, int64_t stream_stride, int64_t stream_offset, icm_view mu,
                      int B, int C, int64_t P, icm_view y_hat_f32, icm_view y_hat_bf16_a,
                      icm_view y_hat_bf16_b, void *stream);

/* E4  GaussianConditional.forward in eval mode (entropy_models.py:626-659): ŷ = round(y-mu)+mu and
 *     likelihood = max(Phi((.5-|ŷ-mu|)/s) - Phi((-.5-|ŷ-mu|)/s), bound), s = max(scale, scale_bound). */
int icm_gc_likelihood(icm_view y, icm_view mu, icm_view scale, int B, int C, int64_t P, float scale_bound,
                      float likelihood_bound, icm_view y_hat_f32, icm_view likelihood, icm_view y_hat_bf16_a,
                      icm_view y_hat_bf16_b, void *stream);

/* y_hat += lrp (fp32), refreshing the bf16 copies (stf.py:628-631, 723-726, 773-776). */
int icm_add_lrp(icm_view y_hat_f32, icm_view lrp, int B, int C, int64_t P, icm_view y_hat_bf16_a,
                icm_view y_hat_bf16_b, void *stream);

/* E5/E6  EntropyBottleneck (entropy_models.py:400-433, 446-489, 492-522).
 *     d_params: the module's parameters packed per channel, 59 floats each:
 *       softplus-free raw _matrix0[3] _bias0[3] _factor0[3] _matrix1[9] _bias1[3] _factor1[3]
 *       _matrix2[9] _bias2[3] _factor2[3] _matrix3[9] _bias3[3] _factor3[3] _matrix4[3] _bias4[1]
 *       median[1]   (filters (3,3,3,3) only, the reference's default and the only one its models use).
 *     mode 0: symbols = round(z - median) in stream order (c,h,w) + ẑ          (compress, :508-515)
 *     mode 1: ẑ and likelihood (eval forward, :446-489)
 *     mode 2: ẑ = float(symbols) + median                                       (decompress, :517-522) */
#define ICM_EB_PARAMS_PER_CHANNEL 59
int icm_eb_process(int mode, icm_view z, int B, int C, int64_t P, const float *d_params,
                   float likelihood_bound, int32_t *d_symbols, int32_t *d_indexes, icm_view z_hat_f32,
                   icm_view z_hat_bf16, icm_view likelihood, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Transforms (rows T1-T11).  Activations are channels-last; GEMM operands bf16, accumulation fp32.
 *
 * icm_conv2d: implicit-GEMM convolution / linear layer on the tcgen05 tensor cores.
 *     in:   bf16 NHWC [B, H, W, in_pitch], the first Cin channels are read;
 *     w:    bf16 [Cout_pad][KH*KW*Cin_pad] (tap-major, then channel; Cin_pad = Cin rounded up to 64 with
 *           zero weights), packed by icm_pack_conv_weight;
 *     out:  row pitch out_pitch, channel offset applied by the caller through the pointer.
 *     Replaces torch.conv2d / F.linear at stf.py:34-40,97,119,233,256,464-546 and cnn.py:31-52. */
#define ICM_ACT_NONE 0
#define ICM_ACT_GELU 1       /* exact erf form (nn.GELU default) */
#define ICM_ACT_HALF_TANH 2  /* 0.5*tanh(x)  (stf.py:628) */
#define ICM_ACT_SIGMOID 3
#define ICM_ACT_RSQRT 4      /* GDN:  x * rsqrt(beta + gamma . x^2)  with res_mode = multiply (gdn.py:62-75) */
#define ICM_ACT_SQRT 5       /* IGDN: x * sqrt(...) */
#define ICM_OUT_BF16 0
#define ICM_OUT_F32 1
typedef struct {
    const void *in;        /* bf16 */
    const void *weight;    /* bf16 packed */
    const float *bias;     /* fp32 [Cout] or NULL */
    void *out;
    const void *residual;  /* second epilogue operand, fp32 or bf16 (res_dtype), same indexing as out, or NULL */
    int B, H, W;           /* input spatial size */
    int Cin, in_pitch;     /* channels read per pixel (any value; zero-filled up to a multiple of 64) and row pitch */
    int Cout, out_pitch;
    int KH, KW, stride, pad;
    int act, out_dtype;
    int pixel_shuffle;     /* 0 or r: output written as PixelShuffle(r) of the conv result */
    int res_pitch;
    int res_dtype;         /* ICM_OUT_F32 / ICM_OUT_BF16 */
    int res_mode;          /* 0: out = act(acc) + res   1: out = act(acc + res)   2: out = act(acc) * res */
} icm_conv_args;
int icm_conv2d(const icm_conv_args *a, void *stream);
/* G convolutions of ONE geometry (a) in one launch: the channel-conditional stacks of the context model are
 * independent of each other wherever the reference's data flow allows it -- cc_mean_transforms[i] and
 * cc_scale_transforms[i] of a slice (stf.py:613-620), and every stack of the slices i >= max_support_slices,
 * whose support is the same first max_support_slices decoded slices (stf.py:612) -- so they share a launch
 * instead of running as G small ones.  Group g uses rows [g*weight_group_rows, +Cout) of the stacked packed
 * weight, bias + g*bias_group_stride, images [in_image_offset[g], +B) of `in` (a tensor of in_images images: 0 for
 * a shared input, g*B for stacked ones) and writes at out + g*out_group_stride elements.  tail_channel[g] >= 0
 * (then for every g): the last 64-channel chunk of each tap is read at that channel of `in` instead of at
 * 64*(chunks-1) -- the LRP stacks read [support | y_hat_i] with a per-slice position of y_hat_i.  Per output element
 * the reduction order equals icm_conv2d's, so the results are bit-identical to G separate calls. */
#define ICM_MAX_CONV_GROUPS 16
typedef struct {
    int groups;
    int in_images;
    int64_t weight_group_rows;
    int64_t bias_group_stride;   /* floats */
    int64_t out_group_stride;    /* elements of the output type */
    int in_image_offset[ICM_MAX_CONV_GROUPS];
    int tail_channel[ICM_MAX_CONV_GROUPS];   /* -1: none */
} icm_conv_groups;
int icm_conv2d_grouped(const icm_conv_args *a, const icm_conv_groups *g, void *stream);
/* Cap on the SMs icm_conv2d / icm_swin_mlp occupy (0 = all), for overlap with the rANS coders on another stream. */
int icm_set_conv_sm_limit(int n_sms);
/* Fused Swin MLP (stf.py:34-40 inside :194-198): x[row] += fc2(GELU(fc1(h[row]))), h = LayerNorm2(x) in bf16 [rows, C],
 * x the fp32 residual stream [rows, C] updated in place; w1 / w2 packed by icm_pack_conv_weight ([4C][ceil64(C)] and
 * [C][4C]).  C in {48, 96, 192}; the 4C-wide hidden activation stays in shared memory / TMEM.  Bit-identical to two
 * icm_conv2d launches (ICM_ACT_GELU, then residual). */
int icm_swin_mlp(const void *d_h_bf16, const void *d_w1_packed, const float *d_b1, const void *d_w2_packed, const float *d_b2,
                 float *d_x, int64_t rows, int C, void *stream);
/* Whole Swin block in one pass over the residual stream, for the narrow stages C in {48, 96} (window 4, head_dim 16;
 * stf.py:149-199): parts & 1: x += proj(attention(qkv(LayerNorm1(x)))); parts & 2: x += fc2(GELU(fc1(LayerNorm2(x)))).
 * x fp32 [B*H*W, C] in place; one warp per 4x4 window, every intermediate in registers (csrc/swin_fused.cu).
 * Weights: bf16 rows as icm_pack_conv_weight leaves a 1x1 layer ([Cout][Cin padded to 64]); biases / LayerNorm
 * parameters / relative-position table [49][heads] fp32.  Replaces icm_layernorm + icm_conv2d (qkv) +
 * icm_window_attention + icm_conv2d (proj) + icm_layernorm + icm_swin_mlp at these widths. */
int icm_swin_block(float *d_x, int B, int H, int W, int C, int heads, int window, int shift, int parts,
                   const void *d_w_qkv, const float *d_b_qkv, const void *d_w_proj, const float *d_b_proj,
                   const float *d_rel_table, const float *d_ln1_g, const float *d_ln1_b,
                   const void *d_w_fc1, const float *d_b_fc1, const void *d_w_fc2, const float *d_b_fc2,
                   const float *d_ln2_g, const float *d_ln2_b, void *stream);
int icm_pack_conv_weight(const float *d_w_oihw, int Cout, int Cin, int KH, int KW, int Cin_pad, int Cout_pad,
                         int pixel_shuffle, void *d_out_bf16, void *stream);

/* LayerNorm over the last dim (eps 1e-5), fp32 in -> bf16 or fp32 out.  gather = 1 applies the
 * PatchMerging 2x2 gather first (stf.py:225-232: channel blocks x(0,0),x(1,0),x(0,1),x(1,1)). */
int icm_layernorm(const float *d_in, const float *d_gamma, const float *d_beta, void *d_out, int out_dtype,
                  int64_t rows, int C, int gather, int B, int H, int W, void *stream);

/* fp32 -> bf16 copy of a [rows, C] block with row pitches (elements); C % 4 == 0. */
int icm_cast_bf16(const float *d_in, int64_t rows, int C, int64_t in_pitch, void *d_out, int64_t out_pitch, void *stream);

/* Per-stream status of a decoder (0, ICM_ERR_BAD_INDEX, or the error its stream's encoder reported when the streams
 * were handed over on the device); synchronises the stream. */
int icm_rans_decoder_status(icm_rans_decoder *d, int32_t *h_status, void *stream);
/* Asynchronous form for pipelines: *d_flag = min(*d_flag, statuses) / min(*d_flag, d_values[0..n)), so that one
 * device word collects the encoder sizes (negative = ICM_ERR_*) and decoder statuses of a whole round trip and the
 * host reads it once.  The reference has nothing to report here (compiled-out asserts, SURVEY.md 8b "Errors"). */
int icm_rans_decoder_status_min(icm_rans_decoder *d, int32_t *d_flag, void *stream);
int icm_min_i32(const int32_t *d_values, int64_t n, int32_t *d_flag, void *stream);

/* T4 core: softmax(q k^T * scale + bias[rel_idx] + mask) v per (window, head) on a qkv tensor
 *     bf16 [B, H, W, 3C] (feature order s*C + h*hd + d, stf.py:97), writing bf16 [B, H, W, C] in token
 *     order; `shift` > 0 applies the cyclic shift and the region mask (stf.py:166-171,316-334). */
int icm_window_attention(const void *d_qkv, void *d_out, const float *d_bias_table, int B, int H, int W,
                         int C, int heads, int window, int shift, void *stream);

/* T1  PatchEmbed: Conv2d(3->C, k2, s2) + LayerNorm(C) from an NCHW fp32 image to fp32 tokens
 *     (stf.py:365-381). */
int icm_patch_embed(const float *d_img, const float *d_w, const float *d_b, const float *d_gamma,
                    const float *d_beta, float *d_tokens, int B, int H, int W, int C, void *stream);

/* T10 tail: Conv2d(C->3, k3, p1) from bf16 NHWC to an fp32 NCHW image, optional clamp to [0,1]
 *     (stf.py:466,784). */
int icm_final_conv(const void *d_in_bf16, const float *d_w, const float *d_b, float *d_img, int B, int H,
                   int W, int C, int clamp01, void *stream);

/* ------------------------------------------------------------------------------------------------
 * T11  WACNN ("cnn" / "cnn2" codec) pieces; the convolutions themselves go through icm_conv2d.
 *
 * NCHW fp32 image <-> channels-last bf16 (`pitch` channels, the extra ones zero / ignored); the output
 * conversion optionally clamps to [0,1] (cnn.py:331). */
int icm_image_to_nhwc(const float *d_img, void *d_out_bf16, int B, int C, int H, int W, int pitch, void *stream);
int icm_nhwc_to_image(const void *d_in_bf16, float *d_img, int B, int C, int H, int W, int pitch, int clamp01, void *stream);
/* mode 0: out = x*x (the GDN operand, gdn.py:68); mode 1: out = a*s + x (gated attention output, layers.py:87-89);
 * mode 2: the same written as fp32 (the latent y);
 * bf16 [rows, C] blocks with row pitches in elements. */
int icm_eltwise_bf16(int mode, const void *d_a, int64_t pitch_a, const void *d_s, int64_t pitch_s, const void *d_x,
                     int64_t pitch_x, void *d_out, int64_t pitch_out, int64_t rows, int C, void *stream);
/* WindowAttention core of WinBasedAttention (win_attention.py:90-207): window 8 / head_dim 24 and window 4 /
 * head_dim 40, optional cyclic shift + region mask, no padding; qkv bf16 [B,H,W,3C] -> bf16 [B,H,W,C]. */
int icm_window_attention_wacnn(const void *d_qkv, void *d_out, const float *d_bias_table, int B, int H, int W, int C,
                               int heads, int window, int shift, void *stream);
/* ConvTranspose2d(k5, s2, p2, output_padding 1) weight [Cin][Cout][5][5] -> the packed weight of an equivalent
 * 3x3 convolution with 4*Cq phase-major output channels, to be run by icm_conv2d with pixel_shuffle = 2
 * (models/utils.py:124-132). */
int icm_pack_deconv_weight(const float *d_w_iohw, int Cin, int Cout, int Cin_pad, int Cq, void *d_out_bf16, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Training step (BASELINE.json configs[4]; reference train.py:172-214).  The transforms' backward
 * runs through the host framework's autograd; these are the fused memory-bound pieces around it.
 *
 * icm_rows: a [rows, n] fp32 tensor whose rows are `stride` elements apart (a 32-channel slice of an NCHW latent:
 * rows = B, n = 32*H*W, stride = C*H*W).
 *
 * icm_gc_train_forward: GaussianConditional.forward in training mode fused with the straight-through y_hat:
 *     likelihood = max(Phi((.5 - v)/s) - Phi((-.5 - v)/s), likelihood_bound), v = |y + noise - mu|, s = max(scale, scale_bound)
 *                  (entropy_models.py:126-135 "noise", :626-659; LowerBound bound_ops.py:21-62)
 *     y_hat      = round(y - mu) + mu                                        (stf.py:622, ops.py:20-34)
 * icm_gc_train_backward: gradients towards y, mu and scale given those of likelihood and y_hat, with LowerBound's
 *     pass-through rule (x >= bound or grad < 0) for both bounds and the identity gradient of ste_round. */
typedef struct {
    void *ptr;
    int64_t stride;
} icm_rows;
int icm_gc_train_forward(icm_rows y, icm_rows noise, icm_rows mu, icm_rows scale, int64_t rows, int64_t n, float scale_bound,
                         float likelihood_bound, icm_rows likelihood, icm_rows y_hat, void *stream);
int icm_gc_train_backward(icm_rows y, icm_rows noise, icm_rows mu, icm_rows scale, icm_rows g_likelihood, icm_rows g_y_hat,
                          int64_t rows, int64_t n, float scale_bound, float likelihood_bound, icm_rows g_y, icm_rows g_mu,
                          icm_rows g_scale, void *stream);
/* torch.nn.utils.clip_grad_norm_ (train.py:208-209) on ONE flat fp32 gradient buffer, nothing read back by the host:
 * *d_sumsq += sum(x^2) (zero it first; call once per buffer), then
 * *d_coef = pre_scale * min(1, max_norm / (pre_scale * sqrt(*d_sumsq) + 1e-6))   (max_norm <= 0: no clipping),
 * pre_scale = 1 / world when the buffer holds gradient SUMS over the data-parallel ranks; *d_norm (optional) = the norm. */
int icm_grad_sumsq(const float *d_x, int64_t n, float *d_sumsq, void *stream);
int icm_clip_coef(const float *d_sumsq, float max_norm, float pre_scale, float *d_coef, float *d_norm, void *stream);
/* torch.optim.Adam.step (train.py:161-168, 210, 214) over flat buffers: g' = grad * grad_scale * (*d_grad_scale if given);
 * m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  t = step >= 1, or, with
 * d_step_state (3 floats {t, 1-b1^t, 1/sqrt(1-b2^t)}, zero-initialised), a counter kept and advanced on the device, so that a
 * step captured in a CUDA graph advances on every replay. */
int icm_adam_step(float *d_param, const float *d_grad, float *d_exp_avg, float *d_exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float *d_step_state, const float *d_grad_scale, float grad_scale, void *stream);

/* nn.LayerNorm (eps 1e-5) of the Swin blocks in the training step (stf.py:155,197,232,256,379): forward keeps the per-row mean and
 * 1/std; y fp32 or bf16 (ICM_OUT_*).  Backward: grad_x, and grad_gamma / grad_beta summed over the rows (zeroed here first); grad_y fp32
 * or bf16.  x is [rows, C] fp32 contiguous, C a multiple of 4, <= 768. */
int icm_layernorm_train_forward(const float *d_x, const float *d_gamma, const float *d_beta, void *d_y, int y_dtype, float *d_mean,
                                float *d_rstd, int64_t rows, int C, void *stream);
int icm_layernorm_train_backward(const float *d_x, const void *d_grad_y, int grad_dtype, const float *d_gamma, const float *d_mean,
                                 const float *d_rstd, float *d_grad_x, float *d_grad_gamma, float *d_grad_beta, int64_t rows, int C,
                                 void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ICM_B200_H */
