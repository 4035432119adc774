/* mmdx.h - C ABI of the B200-native batched multimodal inference forward.
 *
 * The reference (PravCoder/Multi-Modal-Medical-Imaging-and-Report-ML-Diagnosis-System) has no
 * FFI/plugin interface: its boundary is the Python function
 *     inference(model_bundle, image_pil, patient_details, device=None, gen_kwargs=None)
 *     backend/ml/pipelines/inference_pipeline.py:150-206   (called from backend/api/views.py:85)
 * and the three nn.Modules it drives (backend/ml/pipelines/training_pipeline.py:157-311,
 * :348-508, :516-618).  This header is what a ctypes binding behind that function binds
 * (see INTEGRATION.md); each entry point names the reference step it replaces.
 *
 * Conventions: plain pointers and sizes only; `d_` = device pointer, `h_` = host pointer;
 * every call returns 0 on success, non-zero on failure with the message available from
 * mmdx_last_error() (thread-local).  `stream` is a cudaStream_t passed as void*.
 * There is NO CPU fallback: without a CUDA device mmdx_create fails.
 */
#ifndef MMDX_H
#define MMDX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmdx_engine mmdx_engine;

typedef struct mmdx_config {
  int32_t device;         /* CUDA device ordinal */
  int32_t resize_short;   /* T.Resize(256)      training_pipeline.py:113 ; 0 = no resize */
  int32_t crop;           /* T.CenterCrop(224)  training_pipeline.py:114 ; 0 = no crop (needs resize_short==0) */
  int32_t n_heads;        /* BERT attention heads (12) */
  float   mean[3];        /* T.Normalize mean   training_pipeline.py:117 */
  float   std[3];         /* T.Normalize std */
  int32_t keep_fp32;      /* != 0: mmdx_finalize_weights also keeps an fp32 copy of every weight (+452 MB) for mmdx_forward_f32 */
} mmdx_config;

const char* mmdx_last_error(void);
const char* mmdx_version(void);

/* Host-only helper (no GPU needed): Pillow ImagingResample bilinear coefficients for output
 * indices [out_first, out_first+n) of an axis resized in_size -> out_size.  `weights` is
 * [n][ksize] int32 in 22-bit fixed point.  Returns ksize (>0) or -1. */
int mmdx_resample_coeffs(int in_size, int out_size, int out_first, int n, int32_t* first, int32_t* count,
                         int32_t* weights, int weights_capacity);
/* torchvision Resize(int) output size + CenterCrop offsets (functional.py:353-384, :592-594). */
int mmdx_resize_geometry(int h, int w, int resize_short, int crop, int* out_h, int* out_w, int* top, int* left);

/* ---- lifetime / weights: replaces api/views.py:216-234 (build modules + load_state_dict) ---- */
int mmdx_create(const mmdx_config* cfg, mmdx_engine** out);
void mmdx_destroy(mmdx_engine* e);
/* One call per state_dict tensor, fp32 host data, names prefixed "image." / "text." / "fusion."
 * followed by the reference's state_dict key. */
int mmdx_load_tensor(mmdx_engine* e, const char* name, const float* h_data, int ndim, const int64_t* shape);
/* Fold BatchNorm into conv weights+bias, fuse Q/K/V, cast to bf16, lay out K-major for TMA, upload. */
int mmdx_finalize_weights(mmdx_engine* e);
/* Packed weight file (replaces the per-process bundle rebuild of backend/api/views.py:188-258):
 * mmdx_save_packed writes the finalized arena (BN folded, QKV fused, bf16, kernel-ready layouts) with its table of
 * dimensions and offsets; mmdx_load_packed on a freshly created engine reads it back with one host-to-device copy
 * instead of mmdx_load_tensor x N + mmdx_finalize_weights.  Checksummed; a file of another layout version is rejected. */
int mmdx_save_packed(mmdx_engine* e, const char* path);
int mmdx_load_packed(mmdx_engine* e, const char* path);
int mmdx_num_sms(mmdx_engine* e);
/* dims read from the loaded weights: d_img, d_txt, d_fuse_hidden, n_disease, hidden, n_layers, width of cond_proj
 * (n_cond * h_dec; 0 if the bundle has none), rows of the position table (longest sequence) */
int mmdx_dims(mmdx_engine* e, int32_t out[8]);
/* rows of the word / position / token-type embedding tables.  nn.Embedding raises IndexError on an index outside its
 * table (training_pipeline.py:470-473 -> HF BertEmbeddings): the HOST entry points (mmdx_forward_host*) check every id
 * against these sizes and fail; the device-pointer entry points cannot see the ids, so the kernel clamps them (memory
 * safety) and the Python wrappers validate at token-packing time. */
int mmdx_table_sizes(mmdx_engine* e, int32_t out[3]);

/* ---- the hot path ------------------------------------------------------------------------ */
/* image_transfom_into_tensor (training_pipeline.py:112-119) + ImageEncoderCNN.forward (:306-311).
 * d_images: uint8 [B,H,W,C] (C = 1 or 3).  Outputs nullable: d_feats fp32 [B,2048], d_z_img fp32 [B,d_img]. */
int mmdx_image_encode(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, float* d_feats,
                      float* d_z_img, void* stream);
/* TextEncoderTransformer.forward (:503-508) over packed (unpadded) tokens: d_ids/d_pos/d_tt int32 [T],
 * d_cu_seqlens int32 [B+1].  Outputs nullable: d_pooled fp32 [B,768], d_z_txt fp32 [B,d_txt]. */
int mmdx_text_encode(mmdx_engine* e, const int32_t* d_ids, const int32_t* d_pos, const int32_t* d_tt,
                     const int32_t* d_cu_seqlens, int B, int T, int max_len, float* d_pooled, float* d_z_txt,
                     void* stream);
/* FusionTransformerModel.forward with report_labels=None (:584-592) + sigmoid/threshold
 * (inference_pipeline.py:185-186) on the embeddings left in the engine by the two calls above. */
int mmdx_head(mmdx_engine* e, int B, const float* d_thresholds, float* d_z_fuse, float* d_logits, float* d_probs,
              uint8_t* d_vector, void* stream);
/* All three, device buffers. */
int mmdx_forward(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, const int32_t* d_ids,
                 const int32_t* d_pos, const int32_t* d_tt, const int32_t* d_cu_seqlens, int T, int max_len,
                 const float* d_thresholds, float* d_logits, float* d_probs, uint8_t* d_vector, void* stream);
/* The same path in fp32 (the reference's own precision, inference_pipeline.py:156-157): weights, activations and
 * arithmetic in fp32 on the CUDA cores, held to 1e-5 on the probabilities and to EXACT labels against the reference (BASELINE
 * north_star "1e-5 if run in fp32").  A parity mode - an order of magnitude slower than mmdx_forward.  Needs an engine
 * created with keep_fp32 != 0 and weights loaded by mmdx_load_tensor.  Every fp32 intermediate output is nullable:
 * d_feats [B,2048], d_z_img [B,d_img], d_pooled [B,hidden], d_z_txt [B,d_txt], d_z_fuse [B,d_fuse]. */
int mmdx_forward_f32(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, const int32_t* d_ids,
                     const int32_t* d_pos, const int32_t* d_tt, const int32_t* d_cu_seqlens, int T, int max_len,
                     const float* d_thresholds, float* d_feats, float* d_z_img, float* d_pooled, float* d_z_txt,
                     float* d_z_fuse, float* d_logits, float* d_probs, uint8_t* d_vector, void* stream);
/* Same with HOST buffers (pinned recommended): H2D copies, forward, D2H copies, stream sync. */
int mmdx_forward_host(mmdx_engine* e, const uint8_t* h_images, int B, int H, int W, int C, const int32_t* h_ids,
                      const int32_t* h_pos, const int32_t* h_tt, const int32_t* h_cu_seqlens, int T, int max_len,
                      const float* h_thresholds, float* h_logits, float* h_probs, uint8_t* h_vector, void* stream);
/* Report-generation conditioning (first step): cond = GELU(z_fuse * Wc^T + bc), fp32
 * [B, n_cond * h_dec], for the batch mmdx_head (or mmdx_forward*) has just processed - FusionTransformerModel.
 * _make_encoder_outputs (training_pipeline.py:574-578), i.e. the "encoder output" the T5 decoder cross-attends to.
 * Fails if the bundle carried no fusion.cond_proj.0 weights. */
int mmdx_cond_tokens(mmdx_engine* e, int B, float* d_cond, void* stream);
/* The same call as a two-deep pipeline for throughput serving: _submit enqueues request `slot` (0 or 1: H2D of its inputs,
 * forward, D2H of its results) and returns; _wait blocks until that slot's results are in the host buffers.  With two
 * requests in flight the image batch of request k+1 crosses PCIe under the kernels of request k.  Host buffers must stay
 * valid (pinned, for real overlap) until _wait; submitting to a busy slot waits for it first. */
int mmdx_forward_host_submit(mmdx_engine* e, int slot, const uint8_t* h_images_u8, int B, int H, int W, int C,
                             const int32_t* h_ids, const int32_t* h_pos, const int32_t* h_tt, const int32_t* h_cu_seqlens,
                             int T, int max_len, const float* h_thresholds, float* h_logits, float* h_probs,
                             uint8_t* h_vector, void* stream);
int mmdx_forward_host_wait(mmdx_engine* e, int slot);
/* Optional GPU JPEG decode (api/views.py:70 decodes with Pillow on the host, which sustains 6.8 k
 * images/s on 16 cores against 29 k studies/s of the GPU path): n baseline JPEGs of one size H x W (host pointers) ->
 * uint8 HWC RGB [n,H,W,3] on the device = the image input of mmdx_forward.  nvJPEG is loaded with dlopen on first use
 * (no link dependency); NOT bit-identical to libjpeg-turbo (~2 % of the bytes differ by one), so the drop-in inference()
 * keeps Pillow and this path has its own tolerance test.  mmdx_jpeg_backend: nvjpegBackend_t in use (-1 before first use). */
int mmdx_decode_jpeg_batch(mmdx_engine* e, const uint8_t* const* h_blobs, const size_t* h_sizes, int n, int H, int W,
                           uint8_t* d_out_rgb, void* stream);
int mmdx_jpeg_backend(mmdx_engine* e);
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t mmdx_launch_count(mmdx_engine* e);
/* Per-kernel-class device time: between begin and end every launch is bracketed by CUDA events on its
 * stream.  Classes: 0 preprocess, 1 stem conv, 2 pooling, 3 bottleneck convs, 4 text GEMMs, 5 attention,
 * 6 layernorm/embedding, 7 head, 8 other.  mmdx_profile_end returns the number of classes (9) on success, 1 on error. */
int mmdx_profile_begin(mmdx_engine* e);
int mmdx_profile_end(mmdx_engine* e, float* ms_by_class, int64_t* launches_by_class, int n_classes);
/* same, per launch: device time (ms) and class of every launch since mmdx_profile_begin, in launch order; returns the
 * number of launches written (<= cap) or -1 */
int mmdx_profile_end_list(mmdx_engine* e, float* ms, int32_t* cls, int cap);

/* ---- native WordPiece tokenizer (host only, no GPU needed) ---------------------------------
 * Replaces the per-call HF tokenizer of tokenize_patient_details (training_pipeline.py:323,335-342) for 7-bit ASCII text
 * with bit-identical ids: BertNormalizer (clean text, lower-case) -> BertPreTokenizer (whitespace + punctuation) ->
 * WordPiece ("##" continuation, 100-character words, [UNK]) -> [CLS] pieces[:max_len-2] [SEP] -> [PAD] to max_len.
 * vocab_txt: the bytes of a BERT vocab.txt (one token per line, id = line number).
 * mmdx_tokenize_batch: string i is text[offsets[i] : offsets[i+1]] (UTF-8 bytes, no terminator).  Writes ids [n, max_len]
 * and lens [n] (valid tokens including [CLS]/[SEP]; attention_mask = position < len, token_type_ids = 0).
 * fallback[i] = 1 marks a string this code does not handle (a byte >= 0x80, or the literal text of a special token such as
 * "[SEP]"): its ids row is undefined and the caller must tokenise exactly that string with the bundle's HF tokenizer.
 * n_threads <= 0: one per hardware thread (never more than one per 64 strings). */
typedef struct mmdx_tokenizer mmdx_tokenizer;
int mmdx_tokenizer_create(const char* vocab_txt, size_t vocab_bytes, int lower_case, mmdx_tokenizer** out);
void mmdx_tokenizer_destroy(mmdx_tokenizer* t);
int mmdx_tokenize_batch(mmdx_tokenizer* t, const char* text, const int64_t* offsets, int n, int max_len, int n_threads,
                        int32_t* ids, int32_t* lens, uint8_t* fallback);
const char* mmdx_tokenizer_last_error(void);

/* ---- KV-cached T5 decoder step (csrc/t5_decoder.cu) -------------------------------------------
 * The model call inside the reference's report generation - FusionTransformerModel.generate -> HF
 * T5ForConditionalGeneration.generate (training_pipeline.py:613-618, inference_pipeline.py:190-196) - as CUDA kernels: one
 * new token for each of R = batch x beams rows over a private KV cache, fp32 weights and arithmetic, ONE cooperative
 * launch per token.  Two drivers: HF's own beam search with only the forward replaced (mmdx_b200/t5_fast.FastT5Generator:
 * begin / step / reorder), or the whole search here (mmdx_t5_generate, the default of inference()); both reproduce HF's
 * tokens in the tests.
 * Weights: one mmdx_t5_load_tensor per key of T5ForConditionalGeneration.state_dict() that the decoder uses
 * ("shared.weight", "decoder.block.N.layer.*", "decoder.final_layer_norm.weight", "lm_head.weight" when untied).
 * tied_embeddings: 1 = LM head tied to the embedding and decoder output scaled by d_model^-0.5 (the T5 default), 0 = separate
 * lm_head.weight, no scaling (transformers 4.x with tie_word_embeddings=False), 2 = tied but unscaled (transformers 5.x with
 * tie_word_embeddings=False, where only the scaling is switched off).
 * mmdx_t5_begin: d_enc [R, n_enc, d_model] conditioning tokens per row (mmdx_cond_tokens, repeated over the beams);
 * h_bias [max_steps][n_heads] = relative-position bias by distance (host; the caller evaluates HF's bucket formula).
 * mmdx_t5_step: d_tokens int32 [R] -> d_logits fp32 [R, vocab].  mmdx_t5_reorder: row r continues row d_beam_idx[r]. */
typedef struct mmdx_t5 mmdx_t5;
int mmdx_t5_create(int device, int d_model, int n_heads, int d_kv, int d_ff, int n_layers, int vocab, float eps,
                   int tied_embeddings, mmdx_t5** out);
void mmdx_t5_destroy(mmdx_t5* t);
int mmdx_t5_load_tensor(mmdx_t5* t, const char* name, const float* h_data, int64_t n_elems);
int mmdx_t5_finalize(mmdx_t5* t);
int mmdx_t5_begin(mmdx_t5* t, const float* d_enc, int R, int n_enc, int max_steps, const float* h_bias, void* stream);
int mmdx_t5_reorder(mmdx_t5* t, const int32_t* d_beam_idx, void* stream);
int mmdx_t5_step(mmdx_t5* t, const int32_t* d_tokens, float* d_logits, void* stream);
/* Beam-search scoring of the logits of the last step (native search, t5_fast.NativeBeamSearch): per study the k <= 8 best
 * (row, token) continuations of log_softmax(logits) + beam_score, EOS (ban_eos) and the per-row banned tokens (int32
 * [R, max_ban], -1 padded) masked to -inf (written into d_logits); outputs [R / num_beams, k], descending, idx = row_in_study * vocab + token. */
int mmdx_t5_score_topk(mmdx_t5* t, float* d_logits, const float* d_beam_scores, const int32_t* d_banned, int max_ban,
                       int ban_eos, int eos_id, int num_beams, int k, float* d_out_scores, int32_t* d_out_idx, void* stream);
/* The whole report generation in one call: HF's beam search (GenerationMixin._beam_search; do_sample = False, one EOS id,
 * min_new_tokens and no_repeat_ngram_size processors, length_penalty, early_stopping 0 = False / 1 = True / 2 = "never") with
 * the per-token bookkeeping on the host in C++.  d_cond [B, n_enc, d_model] (mmdx_cond_tokens), num_beams <= 4.
 * h_out int32 [B, max_new_tokens + 1] starts with the decoder start token and is padded the way HF pads (pad id, or EOS when
 * the pad id is 0); *h_out_len = the length of the tensor HF's generate would return. */
int mmdx_t5_generate(mmdx_t5* t, const float* d_cond, int B, int n_enc, int num_beams, int max_new_tokens, int min_new_tokens,
                     int no_repeat_ngram, float length_penalty, int early_stopping, int eos_id, int pad_id, int start_id,
                     const float* h_bias, int32_t* h_out, int32_t* h_out_len, void* stream);
int64_t mmdx_t5_launch_count(mmdx_t5* t);
/* Debug aid: globaltimer (ns) at the phase boundaries of the last decoder step (needs MMDX_T5_PROF=1 at the first step). */
int mmdx_t5_step_profile(mmdx_t5* t, uint64_t* h_out, int cap, int* n_out);
const char* mmdx_t5_last_error(void);

/* ---- single-kernel entry points (parity tests call the hot kernels one at a time) ---------- */
/* out[M,N] = act(A[M,K] * Wt[N,K]^T + bias (+ residual)); bf16 A/Wt/residual, bf16 or fp32 out. */
int mmdx_op_gemm(mmdx_engine* e, const void* d_a, int64_t lda, const void* d_w, const float* d_bias,
                 const void* d_residual, int64_t ldr, void* d_out, int64_t ldc, int M, int N, int K, int act,
                 int out_f32, int bn, void* stream);
/* BertSelfOutput / BertOutput in one launch (HF modeling_bert.py:287-298, 345-356): pre = A * Wt^T + bias + residual
 * (bf16, [M, ldc]), out = LayerNorm(pre) * gamma + beta ([M, ldo]; may alias pre).  N = 768, M >= 256: each CTA pair computes
 * the three 256-wide tiles of its rows back to back and then normalises the rows it has just stored. */
int mmdx_op_gemm_ln(mmdx_engine* e, const void* d_a, int64_t lda, const void* d_w, const float* d_bias,
                    const void* d_residual, int64_t ldr, void* d_pre, int64_t ldc, const float* d_gamma,
                    const float* d_beta, float eps, void* d_out, int64_t ldo, int M, int N, int K, void* stream);
/* NHWC conv k in {1,3}, stride in {1,2}, pad k/2; weights packed [Cout][k*k][Cin] bf16 (BN folded). */
int mmdx_op_conv(mmdx_engine* e, const void* d_in, int NB, int H, int W, int Cin, const void* d_w,
                 const float* d_bias, const void* d_residual, void* d_out, int Cout, int k, int stride, int act,
                 void* stream);
/* One layer-1 bottleneck of ResNet-50 shifted by one conv (torchvision Bottleneck.forward, training_pipeline.py:178-183):
 * t2 = relu(conv3x3(t1, w2) + b2) [64 ch, stays on chip]; y = relu(conv1x1(t2, w3) + b3 + shortcut) [256 ch];
 * t1n = relu(conv1x1(y, w1n) + b1n) [c1n = 64 or 128 ch, the next block's conv1; c1n = 0: not computed].
 * shortcut = d_res [NB,H,W,256], or - when d_x, d_wd, d_bd are given (layer1.0; needs c1n = 64) - the block's downsample
 * conv1x1(x, wd) + bd of the 64-channel block input, computed in the kernel.  NHWC bf16; weights packed
 * [Cout][k*k][Cin] bf16 with BN folded, fp32 biases. */
int mmdx_op_bneck64(mmdx_engine* e, const void* d_t1, const void* d_res, const void* d_w2, const float* d_b2,
                    const void* d_w3, const float* d_b3, const void* d_w1n, const float* d_b1n, int c1n, const void* d_x,
                    const void* d_wd, const float* d_bd, void* d_y, void* d_t1n, int NB, int H, int W, void* stream);
/* Last conv of a bottleneck with a downsample shortcut (layer2-4.0): out = relu(conv1x1(t2, w3) + conv1x1_stride_s(x, wd) + b)
 * as ONE implicit GEMM over [t2 | x]; d_wcat = [Cout][Cmid + Cin] bf16 (w3 | wd along K, BN folded), d_bias = b3 + bd.
 * t2 [NB,OH,OW,Cmid], x [NB,H,W,Cin], out [NB,OH,OW,Cout], OH = (H-1)/stride + 1. */
int mmdx_op_conv3_ds(mmdx_engine* e, const void* d_t2, const void* d_x, const void* d_wcat, const float* d_bias, void* d_out,
                     int NB, int H, int W, int Cin, int Cmid, int Cout, int stride, void* stream);
/* conv3 of a bottleneck and conv1 of the NEXT block as one two-GEMM launch (layers 2-3; csrc/gemm2_tcgen05.cuh):
 * y [M,N1] = relu(t2 [M,K1] * w3^T + b3 + res [M,N1]); t1n [M,N2] = relu(y * w1n^T + b1n).  M = NB*H*W pixel rows (both convs
 * are 1x1 over NHWC), bf16, BN folded; K1 % 64 == 0, N1 and N2 multiples of 128 (of 256 for the 256-wide variant). */
int mmdx_op_conv3_conv1(mmdx_engine* e, const void* d_t2, const void* d_w3, const float* d_b3, const void* d_res, void* d_y,
                        const void* d_w1n, const float* d_b1n, void* d_t1n, int64_t M, int K1, int N1, int N2, void* stream);
/* host-only: the static job schedule of the two-GEMM kernel for one CTA pair, (type 0|1, item, n_tile) triples into out[3*cap];
 * returns the number of jobs */
int mmdx_gemm2_schedule(int num_items, int nt1, int nt2, int reverse, int group, int num_groups, int32_t* out, int cap);
int mmdx_padded_dims(int H, int W, int* hp, int* wp);
/* Fused stem: conv 7x7/2 + bias + ReLU (+ MaxPool 3x3/2 pad 1 when pool != 0) over the same padded 4-channel image.
 * d_w_packed: 14336 bf16 (7 x 64 x 32) from mmdx_pack_stem_weights (host helper: fp32 [64,3,7,7] x optional per-channel scale).
 * d_out: bf16 [NB, OH, OW, 64] (pool == 0) or [NB, PH, PW, 64]; OH = (H-1)/2+1, PH = (OH-1)/2+1. */
int mmdx_pack_stem_weights(const float* w_oihw, const float* scale, uint16_t* out_bf16);
int mmdx_op_stem_pool(mmdx_engine* e, const void* d_in_padded, int NB, int H, int W, const void* d_w_packed,
                      const float* d_bias, void* d_out, int pool, void* stream);
int mmdx_op_preprocess(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, void* d_out_padded,
                       int* out_h, int* out_w, void* stream);
int mmdx_op_resample_u8(mmdx_engine* e, const uint8_t* d_images, int B, int H, int W, int C, uint8_t* d_out,
                        void* stream);
int mmdx_op_avgpool(mmdx_engine* e, const void* d_in, int B, int HW, int C, void* d_out_bf16, float* d_out_f32,
                    void* stream);
int mmdx_op_layernorm(mmdx_engine* e, const void* d_x, int rows, int N, const float* d_gamma, const float* d_beta,
                      float eps, void* d_y, void* stream);
int mmdx_op_embed_ln(mmdx_engine* e, const int32_t* d_ids, const int32_t* d_pos, const int32_t* d_tt, int rows,
                     const void* d_word, const void* d_ptab, const void* d_ttab, const float* d_gamma,
                     const float* d_beta, float eps, void* d_y, void* stream);
/* softmax(Q K^T / 8) V over packed tokens: d_qkv bf16 [T, 3*hidden] (16-byte aligned), d_ctx bf16 [T, hidden]. */
int mmdx_op_attention(mmdx_engine* e, const void* d_qkv, const int32_t* d_cu_seqlens, int n_seq, int T, int max_len,
                      int n_heads, int hidden, void* d_ctx, void* stream);
int mmdx_op_seq_mean_pool(mmdx_engine* e, const void* d_h, const int32_t* d_cu_seqlens, int n_seq, int hidden,
                          void* d_out_bf16, float* d_out_f32, void* stream);
int mmdx_op_head_tail(mmdx_engine* e, const float* d_hidden, int B, int D, const float* d_ln_g, const float* d_ln_b,
                      float eps, const float* d_w, const float* d_b, int n_cls, const float* d_thresholds,
                      float* d_z_fuse, float* d_logits, float* d_probs, uint8_t* d_vector, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDX_H */
