/* syncfusion_b200.h - C ABI of libsyncfusion_b200.so
 *
 * B200-native (sm_100a) replacement for the ONE hot path of mcomunita/syncfusion: the conditioned 1-D U-Net
 * v-diffusion sampling loop behind
 *
 *     model.model.sample(x_noisy=, num_steps=, channels=, embedding=, embedding_scale=)
 *         reference: main/generation.py:77-83, main/module_diffusion.py:200-206
 *     net(x, time, embedding=, embedding_scale=, channels=)          (inner boundary used by the sampler)
 *         reference: exp/model/diffusion.yaml:11-33 -> audio_diffusion_pytorch.UNetV0 / VSampler (un-vendored)
 *
 * The reference boundary is a Python callable, not an FFI; these entry points are what a ctypes binding of that
 * callable needs (INTEGRATION.md shows the stub).  Conventions: plain pointers and sizes, no C++/torch types; every
 * call returns 0 on success or a negative sfb_status, with text in sfb_last_error(); one handle = one device, not
 * thread-safe per handle; all work is enqueued on the caller's cudaStream_t (passed as void*) with no implicit
 * synchronisation, so the calls are CUDA-graph capturable; the caller owns every activation / output / workspace
 * buffer (device memory), the library owns only its re-packed weights.  There is NO CPU fallback.
 */
#ifndef SYNCFUSION_B200_H
#define SYNCFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_MAX_DEPTH 16

typedef enum {
  SFB_OK = 0,
  SFB_ERR_INVALID = -1,     /* bad argument / shape (mirrors the AssertionErrors of the upstream items) */
  SFB_ERR_MISSING = -2,     /* parameter missing at finalize */
  SFB_ERR_CUDA = -3,        /* CUDA runtime / driver error */
  SFB_ERR_UNSUPPORTED = -4, /* configuration outside what the kernels implement */
  SFB_ERR_STATE = -5        /* call order (e.g. forward before finalize) */
} sfb_status;

typedef enum { SFB_PRECISION_FP32 = 0, SFB_PRECISION_BF16 = 1 } sfb_precision;  /* fp32: TF32 MMA, fp32 storage */
typedef enum { SFB_UPSAMPLE_NEAREST = 0, SFB_UPSAMPLE_TRANSPOSE = 1 } sfb_upsample_mode;
typedef enum { SFB_DTYPE_F32 = 0 } sfb_dtype;

/* Field-for-field mirror of exp/model/diffusion.yaml:15-33 (+ the upstream UNetV0 defaults it relies on). */
typedef struct sfb_unet_config {
  int32_t depth;                              /* len(channels) = 8 */
  int32_t in_channels;                        /* :15  1 */
  int32_t channels[SFB_MAX_DEPTH];            /* :16  [8,32,64,128,256,512,1024,1024] */
  int32_t factors[SFB_MAX_DEPTH];             /* :17  [1,4,4,4,2,2,2,2] */
  int32_t items[SFB_MAX_DEPTH];               /* :18  [1,2,2,2,2,2,2,4] */
  int32_t attentions[SFB_MAX_DEPTH];          /* :19  [0,0,0,0,1,1,1,1] */
  int32_t cross_attentions[SFB_MAX_DEPTH];    /* :33  [1]*8 */
  int32_t context_channels[SFB_MAX_DEPTH];    /* :22  [2,8,16,32,64,128,256,256] */
  int32_t attention_heads;                    /* :20  8 */
  int32_t attention_features;                 /* :21  64 */
  int32_t embedding_features;                 /* :32  512 */
  int32_t embedding_max_length;               /* :31  1 */
  int32_t resnet_groups;                      /* upstream default 8 */
  int32_t modulation_features;                /* upstream default 1024 */
  int32_t upsample_mode;                      /* sfb_upsample_mode */
  int32_t precision;                          /* sfb_precision */
} sfb_unet_config;

typedef struct sfb_handle sfb_handle;

/* Replaces hydra instantiate of audio_diffusion_pytorch.DiffusionModel (exp/model/diffusion.yaml:11-33). */
int sfb_create(const sfb_unet_config* cfg, int device, sfb_handle** out);
void sfb_destroy(sfb_handle* h);
const char* sfb_last_error(const sfb_handle* h);

/* Replaces model.load_state_dict (main/generation.py:40-43).  `name` is the flat parameter name documented in
 * DESIGN.md ("d3.items_down.0.resnet.conv1.weight", "time.mlp.weight", ...); `data` may be a host or device
 * pointer (fp32, contiguous).  The tensor is copied; finalize re-packs everything for the kernels. */
int sfb_set_param(sfb_handle* h, const char* name, const void* data, int dtype, const int64_t* shape, int ndim);
int sfb_finalize(sfb_handle* h);

/* Bytes of caller-owned device workspace needed for batch B, length L, CFG on/off and up to `rows` conditioning
 * rows (num_steps + 1 for sample, B for a free-standing net call). */
int sfb_workspace_bytes(sfb_handle* h, int64_t B, int64_t L, int cfg_on, int64_t rows, size_t* out);
/* Same for an embedding of M context tokens (1 <= M <= embedding_max_length, M <= 128).  M = 1 (the shipped
 * embedding_max_length: 1, main/module_diffusion.py:67,71) needs no extra space: the cross-attention items collapse to
 * per-clip bias vectors.  M > 1 runs the general CrossAttentionItem (LayerNorm -> q projection -> attention over the M
 * keys -> output projection) and adds the q / attention-output / k|v buffers. */
int sfb_workspace_bytes_m(sfb_handle* h, int64_t B, int64_t L, int cfg_on, int64_t rows, int64_t M, size_t* out);

/* net(x, time, embedding=, embedding_scale=, channels=) -> v      (inner boundary)
 *   x [B,1,L] f32, sigma [B] f32, channels[d] [B, ctx_d, L_d] f32 (NCL as the reference passes them),
 *   embedding [B, M, emb] f32, v_out [B,1,L] f32; all device pointers. */
int sfb_unet_forward(sfb_handle* h, const float* x, const float* sigma, const float* const* channels, int n_channels,
                     const float* embedding, int64_t M, float embedding_scale, float* v_out, int64_t B, int64_t L,
                     void* workspace, size_t workspace_bytes, void* stream);

/* model.sample(x_noisy, num_steps, channels, embedding, embedding_scale) -> x    (main/generation.py:77-83)
 *   x_out [B,1,L] f32; traj_x / traj_v (nullable) receive every step's x_{i+1} / v_i as [num_steps, B, 1, L];
 *   teacher_x (nullable) [num_steps, B, 1, L]: if given, step i evaluates the net on teacher_x[i] instead of the
 *   running state (teacher-forced parity protocol). */
int sfb_sample(sfb_handle* h, const float* x_noisy, int num_steps, const float* const* channels, int n_channels,
               const float* embedding, int64_t M, float embedding_scale, float* x_out, float* traj_x, float* traj_v,
               const float* teacher_x, int64_t B, int64_t L, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels the last sfb_sample / sfb_unet_forward call enqueued (bench.py's gpu_launches). */
int64_t sfb_last_launch_count(const sfb_handle* h);

/* ---- onset encoder: the producer of `channels` -------------------------------------------------
 * Replaces audio_encoders_pytorch.Encoder1d as configured at exp/model/diffusion.yaml:35-43 and called at
 * main/generation.py:71 / main/module_diffusion.py:196 (`_, y_latent = model.onsets_encoder(y, with_info=True)`).
 * Field-for-field mirror of the yaml block: */
typedef struct sfb_encoder_config {
  int32_t in_channels;                        /* :37  1 */
  int32_t channels;                           /* :38  2 */
  int32_t n_levels;                           /* len(factors) = 8 */
  int32_t multipliers[SFB_MAX_DEPTH + 1];     /* :39  [1,1,4,8,16,32,64,128,128] */
  int32_t factors[SFB_MAX_DEPTH];             /* :40  [1,4,4,4,2,2,2,2] */
  int32_t num_blocks[SFB_MAX_DEPTH];          /* :41  [2]*8 */
  int32_t resnet_groups;                      /* :42  2 */
  int32_t patch_size;                         /* :43  1 */
} sfb_encoder_config;
typedef struct sfb_encoder sfb_encoder;
int sfb_encoder_create(const sfb_encoder_config* cfg, int device, sfb_encoder** out);
void sfb_encoder_destroy(sfb_encoder* h);
const char* sfb_encoder_last_error(const sfb_encoder* h);
/* Parameters under the upstream module names ("to_in.block1.groupnorm.weight", "downsamples.3.downsample.weight",
 * "downsamples.3.blocks.1.block2.project.bias", ...); f32, host or device pointer; copied. */
int sfb_encoder_set_param(sfb_encoder* h, const char* name, const void* data, const int64_t* shape, int ndim);
int sfb_encoder_finalize(sfb_encoder* h);
/* Output length of pyramid level `level` (0 .. n_levels-1; -1: the to_in output = L) for an input of L samples. */
int64_t sfb_encoder_level_length(sfb_encoder* h, int64_t L, int level);
int sfb_encoder_workspace_bytes(sfb_encoder* h, int64_t B, int64_t L, size_t* out);
/* y [B, in_channels, L] f32 -> xs_out[0] = to_in(y) [B, channels*m0, L], xs_out[1 + i] = level i [B, channels*m_{i+1}, L_i]
 * (NCL f32 device buffers, n_out = n_levels + 1): info['xs'][1:-1] of the reference; the last one is also z. */
int sfb_encoder_forward(sfb_encoder* h, const float* y, int64_t B, int64_t L, float* const* xs_out, int n_out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Generation post-processing for a whole batch (main/generation.py:85-98), one pass on the GPU:
 *   cut_prefix (onsets != NULL): gen[i, :, :first_onset_i] = 0 with first_onset_i = first non-zero sample of onsets[i]
 *   (:86-88; first_onset [B] int32 device buffer receives it, L for a track with no onset - the reference raises there);
 *   crop to cut_length (:90,:96); new_freq != 0: torchaudio.functional.resample(orig_freq -> new_freq) (:90-92).
 * gen / onsets [B, 1, L] f32, out [B, 1, out_len] f32 with out_len = sfb_postprocess_out_len(...); device pointers.
 * No handle: the resampling tap table is cached per (device, rate pair). */
int64_t sfb_postprocess_out_len(int64_t cut_length, int orig_freq, int new_freq);
int sfb_postprocess(int device, const float* gen, const float* onsets, int64_t B, int64_t L, int64_t cut_length, int orig_freq,
                    int new_freq, float* out, int64_t out_len, int* first_onset, void* stream);

/* ---- test / profiling hooks (stable, but not part of the reference surface) -------------------------------- */
/* Stop every U-Net evaluation after `n_ops` plan operations (<0: run everything). */
int sfb_dbg_set_op_limit(sfb_handle* h, int n_ops);
/* Number of plan ops for (B, L, cfg_on) and a one-line description of op i:
 * "kind depth stack item out_offset out_bytes rows cols dtype". */
int sfb_dbg_plan_size(sfb_handle* h, int64_t B, int64_t L, int cfg_on, void* workspace, size_t workspace_bytes);
int sfb_dbg_plan_size_m(sfb_handle* h, int64_t B, int64_t L, int cfg_on, int64_t M, void* workspace, size_t workspace_bytes);
int sfb_dbg_op_info(sfb_handle* h, int i, char* buf, int buf_len);
/* Launch the persistent kernels (streaming-K / resident-weight / generic GEMM) with at most `max_ctas` CTAs (0: one per
 * SM).  Process-wide.  Fewer CTAs = more tiles per CTA: the parity tests use it to drive the shared-memory rings, the
 * TMEM double buffer and the residual-slot recycling through many wrap-arounds on shapes the oracle still handles. */
int sfb_dbg_set_grid_limit(sfb_handle* h, int max_ctas);
/* Post-mortem of a device barrier-wait timeout.  Every mbarrier wait of the tcgen05 / TMA pipelines is bounded
 * (about 8 s when nothing arrives); a waiter that hits the bound records {kernel source line, plan op, CTA, thread,
 * barrier, parity, raw barrier word} in a host-mapped log and traps, so the launch ends in a CUDA error instead of a
 * hung GPU.  Copies the decoded log into buf (NUL terminated, truncated to buf_len) and returns its full length,
 * 0 if no wait ever timed out.  Readable after the CUDA context is lost.  sfb_last_error() appends the same text to
 * any SFB_ERR_CUDA message. */
int sfb_dbg_wait_log(sfb_handle* h, char* buf, int buf_len);
/* Test hook for the above: launches one warp that waits on a barrier nobody arrives on and synchronises the stream;
 * returns SFB_ERR_CUDA with the decoded log in sfb_last_error().  The CUDA context is lost afterwards. */
int sfb_dbg_fault_inject(sfb_handle* h, void* stream);
/* Per-op device timing: when enabled, every launch of a U-Net evaluation is bracketed by CUDA events on the caller's
 * stream; after the caller synchronises, the report holds one line per plan op of the most recent evaluation:
 * "index kind depth stack item ms flops bytes" (algorithmic flops / bytes).  bench.py's roofline comes from this. */
int sfb_dbg_profile(sfb_handle* h, int enable);
int sfb_dbg_profile_report(sfb_handle* h, char* buf, int buf_len);
/* Device-clock timeline of one streaming-K plan op (CTA 0): attach with op_index >= 0, run an evaluation, then read
 * back with host_buf != NULL (8 roles x 256 clock64 stamps; tools/sk_timeline.py decodes them). */
int sfb_dbg_sk_timeline(sfb_handle* h, int op_index, long long* host_buf, int n);
/* Stand-alone kernels on raw device buffers (unit tests):
 *   gemm: out = resid + (A1 (*) W + A2 W2 + bias), A* [B, L, K*] in operand precision (bf16 or f32 per `bf16`),
 *         W [taps*N, K1+K2] same precision, out_r f32 / out_t operand precision (nullable), stats f64 [B,8,2]. */
int sfb_dbg_gemm(int bf16, const void* a1, const void* a2, const void* w, const float* bias, const float* resid,
                 float* out_r, void* out_t, double* stats, int B, int L, int K1, int K2, int N, int taps, int a2_bmod,
                 int bias_mod, int gs, void* stream);
/*   attention: qkv [B, N, 1536] -> out [B, N, 512] in operand precision. */
int sfb_dbg_attention(int bf16, const void* qkv, void* out, int B, int N, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNCFUSION_B200_H */
