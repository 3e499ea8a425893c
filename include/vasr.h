/* vasr.h — C ABI of libvasr.so, the B200 (sm_100a) inference path for VELOCITY-ASR v2.
 *
 * The reference (shaderko/velocity-asr) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b).  Its two seams are the package API (velocity_asr/__init__.py:95-145)
 * and the scan_mode switch (velocity_asr/ssm.py:119-126).  Each entry point below replaces
 * one reference callable; the Python package in velocity-asr_b200/velocity_asr binds them
 * with ctypes and keeps the reference's names and argument meaning.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  `*_dev` pointers are device memory on
 *     the handle's GPU, `*_host` pointers are host memory.  All tensors are contiguous fp32
 *     in the reference's own layouts (channels-last: (B, T, C)).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *     asynchronous with respect to the host unless the name ends in `_host`.
 *   - every function returns VASR_OK or an error code; vasr_last_error() gives the text for
 *     the calling thread.  There is no CPU fallback: without a CUDA device every compute
 *     entry point fails with VASR_ERR_CUDA.
 *   - the caller owns all buffers it passes; the handle owns weights and workspace.  The
 *     workspace grows on first use of a larger (B, S) and is then reused: no allocation on
 *     the hot path after warm-up.  One handle per GPU per process; calls on one handle must
 *     not overlap.
 */
#ifndef VASR_H_
#define VASR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vasr_handle vasr_handle;

enum {
  VASR_OK = 0,
  VASR_ERR_INVALID = 1,     /* bad argument / unknown name -> ValueError (ssm.py:126)            */
  VASR_ERR_SHAPE = 2,       /* shape the model cannot take -> RuntimeError (model.py:125)        */
  VASR_ERR_CUDA = 3,        /* CUDA runtime / no device                                          */
  VASR_ERR_STATE = 4,       /* weights missing or not committed                                  */
  VASR_ERR_UNSUPPORTED = 5  /* configuration outside what the kernels are built for              */
};

enum { VASR_SCAN_SEQUENTIAL = 0, VASR_SCAN_PARALLEL = 1, VASR_SCAN_MAMBA = 2 };

/* Mirrors VelocityASRConfig (velocity_asr/model.py:23-68); dropout / checkpointing /
 * use_compile have no meaning for inference and are not carried. */
typedef struct vasr_config {
  int32_t mel_bins;              /* 80  */
  int32_t d_model;               /* 192 */
  int32_t ssm_layers;            /* 8   */
  int32_t ssm_state_dim;         /* 64  */
  int32_t ssm_expand_ratio;      /* 2   */
  int32_t ssm_kernel_size;       /* 4   */
  int32_t global_ssm_layers;     /* 2   */
  int32_t global_ssm_state_dim;  /* 32  */
  int32_t attention_heads;       /* 4   */
  int32_t attention_dim;         /* 48  */
  int32_t vocab_size;            /* 1000 */
  int32_t scan_mode;             /* VASR_SCAN_*; the reference default is PARALLEL (model.py:60) */
} vasr_config;

const char* vasr_last_error(void);
const char* vasr_version(void);

/* ---- handle and weights: VELOCITYASR.__init__ / load_state_dict (model.py:255-299, 416-431) */
int vasr_create(const vasr_config* cfg, int device, vasr_handle** out);
void vasr_destroy(vasr_handle* h);
/* One tensor of the reference state_dict, by its reference key (the 208 names of SURVEY.md
 * section 8b), fp32, host memory, numel elements.  Two extra optional keys override the
 * built-in front-end tables: "frontend.mel_filterbank" (n_mels x 201) and "frontend.window"
 * (400).  Unknown keys -> VASR_ERR_INVALID. */
int vasr_set_weight(vasr_handle* h, const char* name, const float* host, int64_t numel);
/* Packs the tensors into kernel layouts; fails with VASR_ERR_STATE naming the first missing key. */
int vasr_commit_weights(vasr_handle* h);

/* ---- shape helpers: audio.py:104-112 (center=False after reflect pad), model.py:370-383 */
int64_t vasr_num_frames(int64_t samples);   /* 1 + samples / 160                 */
int64_t vasr_num_tokens(int64_t frames);    /* (frames + 1) / 2                  */

/* ---- compute_mel_spectrogram (velocity_asr/audio.py:65-143)
 * pcm (B, S) -> mel (B, T, n_mels), T = vasr_num_frames(S).  S must be > 200 (reflect pad). */
int vasr_log_mel(vasr_handle* h, const float* pcm_dev, int64_t B, int64_t S, int normalize,
                 float* mel_dev, void* stream);

/* ---- VELOCITYASR.forward (velocity_asr/model.py:333-368)
 * mel (B, T, mel_bins) -> logits (B, L, vocab), L = vasr_num_tokens(T).  The three feature
 * pointers are the return_features dict ('temporal_binding', 'local_features',
 * 'fused_features'), each (B, L, d_model); pass NULL to skip. */
int vasr_forward(vasr_handle* h, const float* mel_dev, int64_t B, int64_t T, float* logits_dev,
                 float* feat_tb_dev, float* feat_local_dev, float* feat_fused_dev, void* stream);

/* ---- module-level seams used by the parity tests (reference classes in __all__) ---------- */
/* SSMBlock.forward (ssm.py:404-427).  stack 0 = local_ssm.layers[layer], 1 =
 * global_context.global_ssm.layers[layer]; scan_mode < 0 = the mode the reference would use
 * (config scan_mode for local, PARALLEL for global, ssm.py:529-538). */
int vasr_ssm_block(vasr_handle* h, int stack, int layer, int scan_mode, const float* x_dev,
                   int64_t B, int64_t L, float* out_dev, void* stream);
/* HierarchicalGlobalContext.forward (attention.py:283-319): (B, L, d_model) -> same. */
int vasr_global_context(vasr_handle* h, const float* local_dev, int64_t B, int64_t L,
                        float* out_dev, void* stream);
/* CTCOutputHead.forward (model.py:229-239): (B, L, d_model) -> (B, L, vocab). */
int vasr_ctc_head(vasr_handle* h, const float* x_dev, int64_t B, int64_t L, float* logits_dev,
                  void* stream);

/* ---- the scan operator: SelectiveSSM._sequential_scan / _parallel_scan / _mamba_scan
 * (ssm.py:134-337).  x, dt: (B, L, Di) with row strides ldx, lddt; Bm, Cm: (B, L, N) with row
 * strides ldb, ldc; A: (N), D: (Di) or NULL (no skip term); z: (B, L, Di) stride ldz or NULL
 * (when given, y is multiplied by silu(z), ssm.py:129); y: (B, L, Di) stride ldy.
 * N in {16, 32, 64}.  All device pointers.  No handle: the operator is stateless. */
int vasr_selective_scan(const float* x, int64_t ldx, const float* dt, int64_t lddt, const float* A,
                        const float* Bm, int64_t ldb, const float* Cm, int64_t ldc, const float* D,
                        const float* z, int64_t ldz, float* y, int64_t ldy, int64_t B, int64_t L,
                        int64_t Di, int64_t N, int scan_mode, void* stream);

/* ---- ctc_greedy_decode (velocity_asr/decode.py:27-71)
 * logits (B, L, V) -> tokens (B, L) int32 left-packed, lens (B) int32.  argmax ties -> lowest
 * index; blanks dropped; repeats collapsed when collapse != 0; a blank resets the repeat state. */
int vasr_ctc_greedy(const float* logits_dev, int64_t B, int64_t L, int64_t V, int blank, int collapse,
                    int32_t* tokens_dev, int32_t* lens_dev, void* stream);

/* ---- ctc_greedy_decode_with_timestamps (velocity_asr/decode.py:74-125)
 * As above with repeats collapsed, plus per token the [start, end) frame range of its run of equal
 * predictions: tokens / starts / ends (B, L) int32 left-packed, lens (B). */
int vasr_ctc_greedy_timestamps(const float* logits_dev, int64_t B, int64_t L, int64_t V, int blank,
                               int32_t* tokens_dev, int32_t* starts_dev, int32_t* ends_dev, int32_t* lens_dev,
                               void* stream);

/* ---- ragged batches (the reference's collator pads with zeros and never masks, data.py:145-203, so an
 * utterance's transcript depends on its batch mates; these entry points take the true lengths instead)
 * pcm (B, S) holds utterance b in its first sample_lens[b] samples (200 < len <= S; the rest is ignored).
 * Each utterance is processed exactly as if it were alone: reflect padding, frame count and mel statistics from
 * its own length, zero frames after its end, pooling windows and attention keys from its own token count, decode
 * over its own tokens.  tokens (B, L) / lens (B) as vasr_transcribe, L = vasr_num_tokens(vasr_num_frames(S)).
 * sample_lens / frame_lens are HOST arrays.  vasr_forward_ragged: mel (B, T, n_mels) with frame_lens[b] valid
 * frames (1 <= len <= T); logits rows at or past an utterance's own token count are padding. */
int vasr_transcribe_ragged(vasr_handle* h, const float* pcm_dev, const int32_t* sample_lens_host, int64_t B,
                           int64_t S, int32_t* tokens_dev, int32_t* lens_dev, void* stream);
int vasr_transcribe_ragged_host(vasr_handle* h, const float* pcm_host, const int32_t* sample_lens_host, int64_t B,
                                int64_t S, int32_t* tokens_host, int32_t* lens_host);
int vasr_forward_ragged(vasr_handle* h, const float* mel_dev, const int32_t* frame_lens_host, int64_t B, int64_t T,
                        float* logits_dev, void* stream);

/* ---- ctc_beam_search (velocity_asr/decode.py:128-217), lm_scorer = None
 * Prefix beam search with the reference's max-merge rule, fp64 scores and insertion-order tie-breaks.
 * tokens (B, beam_width, L) int32 (-1 padded), lens (B, beam_width) int32 (-1 where the utterance has fewer
 * beams), scores (B, beam_width) fp64, best beam first.  1 <= beam_width <= 32. */
int vasr_ctc_beam_search(const float* logits_dev, int64_t B, int64_t L, int64_t V, int beam_width, int blank,
                         int32_t* tokens_dev, int32_t* lens_dev, double* scores_dev, void* stream);

/* ---- transcribe: load -> mel -> model -> greedy (scripts/transcribe.py:69-82), batched.
 * tokens (B, L) int32, lens (B); L = vasr_num_tokens(vasr_num_frames(S)). */
int vasr_transcribe(vasr_handle* h, const float* pcm_dev, int64_t B, int64_t S,
                    int32_t* tokens_dev, int32_t* lens_dev, void* stream);
/* Host buffers in, host buffers out; H2D and D2H copies inside; synchronous. */
int vasr_transcribe_host(vasr_handle* h, const float* pcm_host, int64_t B, int64_t S,
                         int32_t* tokens_host, int32_t* lens_host);

/* ---- config 5: FakeQuantize of the 12 non-SSM modules (velocity_asr/quantize.py)
 * prepare_model_for_qat (quantize.py:269-322, ssm_state_fp32 = True) replaces temporal_binding.conv,
 * pool1/pool2.pool_proj, cross_attention.{q,k,v,out}_proj, fusion.{gate_proj.0,local_proj,global_proj,
 * out_proj} and ctc_head.proj.2 by QuantizedLinear / QuantizedConv1d: per-output-channel symmetric int8
 * FakeQuantize of the weight, fp32 matmul, per-tensor asymmetric uint8 FakeQuantize of the output
 * (quantize.py:180-191, 248-266).  vasr_set_quantization(h, 1) + vasr_commit_weights applies the weight side
 * (the module's own weights are kept, unlike the reference, which re-initialises them: SURVEY.md 5.8);
 * output nodes pass through until calibrated (quantize.py:82-84).
 * vasr_calibrate runs one forward on mel (B, T, mel_bins) with the output nodes in training mode: each takes
 * min / max of the tensor it is about to quantise, sets scale / zero point (quantize.py:99-121) and
 * quantises with them; the values stay for later forwards (calibrate_model, quantize.py:325-371).
 * module = reference module name, e.g. "ctc_head.proj.2". */
int vasr_set_quantization(vasr_handle* h, int enabled);
int vasr_calibrate(vasr_handle* h, const float* mel_dev, int64_t B, int64_t T, void* stream);
int vasr_get_quant_params(vasr_handle* h, const char* module, float* scale, float* zero_point);
int vasr_set_quant_params(vasr_handle* h, const char* module, float scale, float zero_point);
/* One quantised projection on its own (parity seam for QuantizedLinear.forward / QuantizedConv1d.forward,
 * quantize.py:180-191, 248-266): the projection that hosts `module` — fused neighbours included: k_proj | v_proj
 * share one, gate_proj.0 | local_proj | global_proj another (input [local | ctx]) — on x (B, R, K) with its
 * fake-quantised weights, bias and output FakeQuantize, no activation.  For temporal_binding.conv x is the mel
 * (B, R = T, mel_bins) and out has (R + 1) / 2 rows per utterance.  out: (B, R_out, N) with N the projection's
 * full width; *n_out, *col0, *ncol (may be NULL) receive N and the column range of `module` inside it. */
int vasr_quant_site(vasr_handle* h, const char* module, const float* x_dev, int64_t B, int64_t R, float* out_dev,
                    int32_t* n_out, int32_t* col0, int32_t* ncol, void* stream);

/* ---- a plain linear layer (F.linear), exposed so the GEMM kernel can be tested alone.
 * act: 0 none, 1 gelu(erf), 2 softplus, 3 sigmoid.  x (M, K) stride ldx, w (N, K), bias (N) or
 * NULL, out (M, N) stride ldo.  K % 16 == 0. */
int vasr_linear(const float* x_dev, int64_t ldx, const float* w_dev, const float* bias_dev,
                float* out_dev, int64_t ldo, int64_t M, int64_t K, int64_t N, int act, void* stream);

/* Same contract through the tensor-core kernel (tcgen05 + TMEM + TMA, 3xTF32 split): the
 * kernel every token-sized projection of the model runs on.  K % 4 == 0, N % 4 == 0.
 * The kernel reads the weight as two TF32 matrices [hi | lo] (2*N*K floats) made by
 * vasr_split_tf32; pass w_split_dev = NULL to have the call split w_dev on the fly. */
int vasr_split_tf32(const float* w_dev, float* split_dev, int64_t numel, void* stream);
/* resid_dev (M, N) with row stride ldr, or NULL: added after the activation (the residual adds of
 * ssm.py:418-425 are fused this way). */
int vasr_linear_tc(const float* x_dev, int64_t ldx, const float* w_dev, const float* w_split_dev,
                   const float* bias_dev, const float* resid_dev, int64_t ldr, float* out_dev, int64_t ldo,
                   int64_t M, int64_t K, int64_t N, int act, void* stream);

/* ---- bookkeeping for the bench: kernels launched by this handle since creation, and the
 * share of the last vasr_transcribe spent in the scan (device ms, CUDA events on `stream`)
 * when timing is enabled with vasr_set_timing(h, 1). */
int64_t vasr_kernel_launches(const vasr_handle* h);
int64_t vasr_tc_launches(const vasr_handle* h);   /* of which: tensor-core projection launches */
int vasr_set_timing(vasr_handle* h, int enabled);
int vasr_last_timing(const vasr_handle* h, float* scan_ms, int32_t* scan_launches, float* total_ms);
int64_t vasr_workspace_bytes(const vasr_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* VASR_H_ */
