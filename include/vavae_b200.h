/*
 * vavae_b200.h -- C ABI of the B200-native LightningDiT sampling engine.
 *
 * The reference (heimaoqqq/VA-VAE) is pure Python/PyTorch and has no FFI of its own for this
 * path; these entry points are what a ctypes binding behind the reference's Python objects
 * binds (INTEGRATION.md shows the stub).  Each function names the reference interface it
 * replaces (paths relative to /root/reference/LightningDiT/).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All tensors are dense, row-major.
 *   - "device" pointers are CUDA device pointers valid on the engine's GPU for the duration of
 *     the call (borrowed; the caller -- e.g. PyTorch -- owns them).  Outputs are pre-allocated
 *     by the caller.  The engine owns packed weights and workspace.
 *   - every call enqueues on the caller-supplied CUDA stream (pass the raw cudaStream_t as
 *     void*, 0 = legacy default stream) and does not synchronise, except the *_host calls,
 *     which return host data and therefore synchronise the stream before returning.
 *   - return value: 0 on success, non-zero on error; vavae_last_error() gives the message of
 *     the last failure on the calling thread.  One engine per (process, GPU); not re-entrant.
 */
#ifndef VAVAE_B200_H
#define VAVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAVAE_ABI_VERSION 3

/* precision modes */
#define VAVAE_PREC_BF16 0 /* tcgen05 bf16 tensor-core path, fp32 accumulate / residual / statistics */
#define VAVAE_PREC_FP32 1 /* fp32 SIMT validation path (parity <= 1e-5 vs the fp32 reference)       */

/* ODE methods (fixed grid), transport/integrators.py:118-125 -> torchdiffeq.odeint(method=...) */
#define VAVAE_ODE_EULER 0
#define VAVAE_ODE_HEUN 1

typedef struct vavae_engine vavae_engine_t;

/* Constructor arguments of LightningDiT, models/lightningdit.py:279-297 (same names, same meaning). */
typedef struct vavae_config {
  int32_t abi_version; /* must be VAVAE_ABI_VERSION */
  int32_t input_size;
  int32_t patch_size;
  int32_t in_channels;
  int32_t hidden_size;
  int32_t depth;
  int32_t num_heads;
  int32_t mlp_hidden;  /* int(2/3*int(D*mlp_ratio)) if use_swiglu else int(D*mlp_ratio); lightningdit.py:213-224 */
  int32_t num_classes;
  int32_t label_rows;  /* rows of y_embedder.embedding_table: num_classes + (class_dropout_prob > 0) */
  int32_t learn_sigma;
  int32_t use_qknorm;
  int32_t use_swiglu;
  int32_t use_rope;
  int32_t use_rmsnorm;
  int32_t wo_shift;
  int32_t precision;      /* VAVAE_PREC_* */
  int32_t max_chunk_rows; /* model rows processed per pass of the layer stack (0 = engine default) */
} vavae_config_t;

/* Replaces LightningDiT.__init__ (models/lightningdit.py:279-338): allocates the engine on CUDA `device`. */
int vavae_create(const vavae_config_t* cfg, int device, vavae_engine_t** out);
int vavae_destroy(vavae_engine_t* e);

/* Replaces nn.Module.load_state_dict (inference.py:88-93): copy one fp32 tensor of the reference
 * state_dict (key names / shapes: SURVEY.md section 8b; nn.Linear (out,in) layout) into the engine.
 * `data` is a host pointer if is_device == 0, else a device pointer.  The extra key
 * "t_embedder.freqs" (128 floats, lightningdit.py:121-123) carries the sinusoid frequencies. */
int vavae_load_weight(vavae_engine_t* e, const char* name, const float* data, const int64_t* shape, int ndim,
                      int is_device);
/* Packs weights for the selected precision (bf16 cast, w12 row interleave, K padding, batched
 * adaLN matrix, TMA descriptors) and allocates workspace.  Must follow the last load_weight. */
int vavae_finalize(vavae_engine_t* e, void* stream);

/* Replaces LightningDiT.forward(x, t, y), models/lightningdit.py:391-418.
 * x: (batch, C, H, W) f32 device; t: (batch,) f32 device; y: (batch,) int64 device; out: (batch, C, H, W) f32 device. */
int vavae_forward(vavae_engine_t* e, const float* x, const float* t, const int64_t* y, float* out, int batch,
                  void* stream);

/* Replaces LightningDiT.forward_with_cfg, models/lightningdit.py:420-442.  batch = 2n rows; only x[:n] is read.
 * cfg_interval: 0 = None/False, 1 = True (guidance only when t[0] >= cfg_interval_start).
 * cfg_channels: number of leading channels that receive guidance (reference: 3, lightningdit.py:432). */
int vavae_forward_with_cfg(vavae_engine_t* e, const float* x, const float* t, const int64_t* y, float cfg_scale,
                           int cfg_interval, float cfg_interval_start, int cfg_channels, float* out, int batch,
                           void* stream);

/* Replaces sample_fn = Sampler.sample_ode(...) -> ode.sample (transport/transport.py:398-443,
 * transport/integrators.py:107-126) driving model.forward_with_cfg (use_cfg=1) or model.forward (use_cfg=0)
 * over the fixed grid `t_grid` (HOST pointer, n_points fp32 values as built by integrators.py:94-101).
 * z, out_final: (batch, C, H, W) f32 device; y: (batch,) int64 device.  With use_cfg=1 batch = 2n as in
 * inference.py:205-211.  out_final receives the last grid point ([-1] of the reference trajectory).
 * out_trace: optional (NULL to skip) device buffer of (n_evals, batch, C, H, W) f32 receiving every
 * evaluation's velocity in call order (n_evals = n_points-1 for Euler, 2*(n_points-1) for Heun). */
/* flags (bit set):
 *   VAVAE_SAMPLE_SKIP_DISCARDED: with use_cfg=1 and cfg_interval=1, an evaluation at t < cfg_interval_start keeps only the
 *   conditional output - the reference still computes the unconditional forward and throws it away
 *   (models/lightningdit.py:436-439).  With this flag the engine runs the n conditional rows only for those
 *   evaluations.  Rows [0, n) of out_final (the half every caller keeps, inference.py:216-218) are bit-identical to the
 *   literal mode; rows [n, 2n) then follow the conditional velocity on every channel instead of the discarded
 *   unconditional one, so leave the flag clear when the literal (2n, C, H, W) state of the reference is needed. */
#define VAVAE_SAMPLE_SKIP_DISCARDED 1
int vavae_sample_ode(vavae_engine_t* e, const float* z, const int64_t* y, const float* t_grid, int n_points,
                     int method, int use_cfg, float cfg_scale, int cfg_interval, float cfg_interval_start,
                     int cfg_channels, float* out_final, float* out_trace, int batch, int flags, void* stream);

/* Same as vavae_sample_ode with HOST buffers for z, y and out_final (H2D / D2H copies inside the call;
 * synchronises the stream before returning).  This is the end-to-end entry bench.py times as `e2e`. */
int vavae_sample_ode_host(vavae_engine_t* e, const float* z_host, const int64_t* y_host, const float* t_grid,
                          int n_points, int method, int use_cfg, float cfg_scale, int cfg_interval,
                          float cfg_interval_start, int cfg_channels, float* out_final_host, int batch, int flags,
                          void* stream);

/* Replaces sample_fn built with sampling_method="dopri5" (the default of Sampler.sample_ode, transport/transport.py:401,
 * and what the fork's generation scripts use): adaptive Dormand-Prince 5(4) from t0 to t1 with torchdiffeq's step
 * controller (RMS error norm over the whole state, atol/rtol, safety 0.9, ifactor 10, dfactor 0.2, Hairer initial
 * step, FSAL, dense output at t1).  torchdiffeq is not vendored by the reference: parity for this entry point is
 * pinned only against this repo's restatement of that algorithm.  Synchronises the stream once per step
 * (the controller needs the error norm on the host).  stats (optional, host): {accepted, rejected, evaluations}. */
int vavae_sample_dopri5(vavae_engine_t* e, const float* z, const int64_t* y, double t0, double t1, double atol, double rtol,
                        int use_cfg, float cfg_scale, int cfg_interval, float cfg_interval_start, int cfg_channels,
                        float* out_final, int batch, int64_t* stats, void* stream);

/* nn.Embedding raises on a label outside the table (models/lightningdit.py:168); a kernel cannot.  The engine's label
 * gathers clamp such an index to row 0 (never an out-of-bounds read) and raise a device flag; this call synchronises
 * `stream`, returns non-zero ("index out of range in self") if the flag was raised since the last check, and clears it.
 * The *_host entry points perform the same check before returning. */
int vavae_device_status(vavae_engine_t* e, void* stream);

/* Number of kernels this engine has launched since creation (bench.py's gpu_launches). */
uint64_t vavae_launch_count(const vavae_engine_t* e);
/* Model rows pushed through the layer stack since creation (= forwards actually executed; bench.py reports
 * utilisation against this, so a skipped unconditional forward is never claimed as work). */
uint64_t vavae_forward_rows(const vavae_engine_t* e);
/* Bytes of device memory the engine owns (weights + workspace). */
uint64_t vavae_device_bytes(const vavae_engine_t* e);
const char* vavae_last_error(void);
int vavae_abi_version(void);

/* Live per-kernel measurement for bench.py's roofline: between profile_begin(kernel) and profile_end every
 * launch of that kernel class ("gemm_qkv", "attention", "gemm_proj", "gemm_w12", "gemm_w3", "norm_modulate")
 * is bracketed by a CUDA event pair on the launching stream.  profile_end synchronises those events and
 * returns their summed duration, the launch count and the summed algorithmic work (FLOPs, or bytes for
 * norm_modulate). */
int vavae_profile_begin(vavae_engine_t* e, const char* kernel);
int vavae_profile_end(vavae_engine_t* e, double* total_ms, int64_t* launches, double* total_work);

/* ---- kernel-level test / measurement entry points (tests/ and bench.py; not needed by a drop-in) ---- */

/* C[M,N] (f32) = A[M,K] (bf16) * W[N,K]^T (bf16) + bias[N] (f32, may be NULL) on the tcgen05 GEMM. */
int vavae_test_gemm_bf16(const void* A, const void* W, const float* bias, float* C, int M, int N, int K,
                         void* stream);
/* O[B*T, H*hd] (bf16) = softmax(q k^T / sqrt(hd)) v on the tcgen05 attention kernel.
 * q, k: (B,H,T,hdp) bf16, hdp = head_dim rounded up to 16, padding zero; vt: (B,H,hdv,T) bf16 with
 * hdv = (head_dim + 1) rounded up to 16, rows [0,hd) = v^T, row hd = ones, remaining rows zero
 * (see DESIGN.md "attention layouts"). */
int vavae_test_attention_bf16(const void* q, const void* k, const void* vt, void* out, int B, int H, int T, int hd,
                              void* stream);
/* One launch of the production tcgen05 GEMM with one of its fused epilogues, on caller-provided device operands, so
 * each epilogue can be checked against an fp32 reference of the same op (tests/test_gpu_epilogues.py):
 *   VAVAE_EPI_QKV            attn.qkv: + bias, [qk RMSNorm], RoPE (models/pos_embed.py:38-42,135), q/k scattered head-major
 *                            (B,H,T,hdp) and V stored transposed (B,H,hdv,T)   (models/lightningdit.py:66-76)
 *   VAVAE_EPI_SWIGLU         mlp.w12: + bias, silu(x1) * x2 on row-interleaved weights (models/swiglu_ffn.py:31-36)
 *   VAVAE_EPI_GATED_RESIDUAL attn.proj / mlp.w3: X += gate[b] * (acc + bias) (models/lightningdit.py:248-249), optionally
 *                            emitting the next RMSNorm's un-normalised bf16 operand and per-row partial sums of squares
 * Optional consumer-side fused RMSNorm (kinds QKV, SWIGLU): accumulator rows are scaled by
 * rsqrt(total / in_dim + 1e-6) and sp[row / T, :] replaces the bias.  Partial sums come in two layouts (rows part_ld
 * floats apart): COARSE, one float per 128-column group, summed in ascending order (in_nparts of them); or FINE
 * (in_fine = 1), eight floats per 32-column chunk (sums of squares of 4 columns each; in_nparts chunks), which the
 * consumer folds into exactly the coarse values first.  A producer (GATED_RESIDUAL with nx_w) reports which layout it
 * wrote in fine_out (tiles of whole 128-column groups: coarse; tiles that split a group: fine) and how many slots /
 * chunks in nparts_out. */
#define VAVAE_EPI_QKV 0
#define VAVAE_EPI_SWIGLU 1
#define VAVAE_EPI_GATED_RESIDUAL 2
typedef struct vavae_epilogue_test {
  int32_t kind;
  int32_t M, N, K, T;       /* token rows, output columns, reduction length, tokens per sample (T % 32 == 0 for the fast path) */
  const void* A;            /* bf16 [M, K] */
  const void* W;            /* bf16 [N, K]; SWIGLU: rows interleaved (x1_j, x2_j) */
  const float* bias;        /* [N] (SWIGLU: interleaved like W) */
  const float* in_part;     /* optional f32 [M, part_ld] partial sums of squares */
  int32_t in_nparts, in_dim;
  const float* sp;          /* f32 [M / T, sp_ld] per-sample shift.W^T + bias vectors (with in_part) */
  int32_t sp_ld;
  int32_t H, hd;            /* QKV */
  const float* cos;         /* f32 (T, hd) or NULL */
  const float* sin;
  const float* qn_w;        /* f32 (hd) q_norm / k_norm weights or NULL */
  const float* kn_w;
  void* q;                  /* bf16 (M/T, H, T, hdp), hdp = hd rounded up to 16 */
  void* k;
  void* vt;                 /* bf16 (M/T, H, hdv, T), hdv = hd + 1 rounded up to 16 */
  void* out_bf16;           /* SWIGLU: bf16 [M, ldo], N / 2 columns written */
  int32_t ldo;
  float* X;                 /* GATED_RESIDUAL: f32 [M, N] residual stream, updated in place */
  const float* mod;         /* f32 [M / T, mod_ld] adaLN output */
  int32_t mod_ld, gate_off;
  const float* nx_w;        /* optional f32 [N]: weight of the RMSNorm that follows */
  int32_t nx_scale_off;     /* column of its adaLN scale in mod */
  void* nx_A;               /* bf16 [M, N]: x_new * nx_w * (1 + scale) */
  float* nx_part;           /* f32 [M, part_ld]: partial sums of x_new^2 (part_ld >= 16 coarse, >= N / 4 fine) */
  int32_t nparts_out;       /* written by the call */
  int32_t in_fine;          /* consumer: in_part holds fine (32-column) slots */
  int32_t fine_out;         /* written by the call: the producer wrote fine slots */
  int32_t part_ld;          /* row pitch (floats) of in_part / nx_part */
} vavae_epilogue_test_t;
int vavae_test_gemm_epilogue(vavae_epilogue_test_t* t, void* stream);

/* Host-only (no GPU needed): how the library would tile and run one layer GEMM of M token rows - the rules the parity
 * of small and large batches rests on (tests/test_host_logic.py).  kind: VAVAE_EPI_*; T: tokens per sample; H, hd: QKV
 * only; fused: the gated-residual launch also emits the next RMSNorm's operand and partial sums.
 * out[6] = {column-tile width, width of the last column tile, 1 if the TMA epilogue runs, its epilogue warps (4 / 8),
 *           its ring slots per warp, 1 if the launch writes coarse (128-column) partial-sum slots}. */
int vavae_gemm_plan(int kind, int M, int N, int K, int T, int H, int hd, int fused, int32_t* out);

/* 1: GEMMs use the CTA-pair kernel (cluster of 2, tcgen05 cta_group::2, M = 256 per pair) where the problem
 * allows it; 0: always the single-CTA kernel.  Process-wide; default set by the engine. */
int vavae_set_gemm_two_cta(int on);
/* Per-kernel timing: runs the named kernel of the engine's own layer 0 `iters` times on the given
 * number of model rows with CUDA events on `stream`; returns mean milliseconds per launch in *ms. */
int vavae_time_kernel(vavae_engine_t* e, const char* kernel, int rows, int iters, float* ms, double* work,
                      void* stream);

/* ===== SDE sampler update =====================================================
 * Replaces the elementwise arithmetic of one step of Sampler.sample_sde (transport/integrators.py:30-49 Euler-Maruyama /
 * Heun, transport/transport.py:306-334 drift and last step) for path_type "Linear" with a velocity model: the drift
 * v + diffusion(t) * score(v, x, t), the score from the velocity (transport/path.py:70-84) and the state update in ONE
 * pass over the latents instead of ~10 elementwise tensor ops.  The batch shares one time, so the coefficients are host
 * scalars (computed by the caller exactly as the reference computes them, in fp32).  The noise `w` is drawn by the
 * caller, which keeps the reference's RNG stream.  Modes: 0 Euler-Maruyama step; 1 Heun x-hat; 2 Heun stage 1 (out =
 * predictor state, out2 = k1); 3 Heun stage 2; 4 last step "Mean"; 5 "Tweedie"; 6 "Euler".  Stateless: needs no engine. */
typedef struct vavae_sde_update {
  int32_t mode;
  const float* x;    /* state the drift is evaluated at (n elements) */
  const float* v;    /* model output at x (modes 0, 2-6) */
  const float* w;    /* standard normal noise (modes 0, 1) */
  const float* k1;   /* mode 3: stage-1 drift */
  const float* xh;   /* mode 3: x-hat */
  float* out;
  float* out2;       /* mode 2: k1 */
  int64_t n;
  float t, dt, diffusion, var, sqrt_2diffusion, sqrt_dt, c0, c1;
} vavae_sde_update_t;
int vavae_sde_update(const vavae_sde_update_t* u, void* stream);

/* ===== VA-VAE decode path =====================================================
 * Replaces, for inference, AutoencoderKL.decode (LightningDiT/tokenizer/autoencoder.py:490-493: post_quant_conv ->
 * Decoder, :277-399) and the caller-side post-processing of VA_VAE.decode_to_images (LightningDiT/tokenizer/
 * vavae.py:72-82: clamp(127.5 x + 128, 0, 255) -> uint8 NHWC).  Weights are loaded under the reference's
 * state_dict names ("post_quant_conv.weight", "decoder.conv_in.weight", "decoder.up.3.block.0.nin_shortcut.bias",
 * ...), fp32, PyTorch layouts. */
typedef struct vavae_decoder vavae_decoder_t;

typedef struct vavae_decoder_config {
  int32_t abi_version;     /* must be VAVAE_ABI_VERSION */
  int32_t z_channels;      /* embed_dim = 32, tokenizer/configs/vavae_f16d32.yaml:8,28 */
  int32_t ch;              /* 128 */
  int32_t num_levels;      /* len(ch_mult), 5 */
  int32_t ch_mult[8];      /* (1, 1, 2, 2, 4) */
  int32_t num_res_blocks;  /* 2 (the decoder runs num_res_blocks + 1 blocks per level, autoencoder.py:343) */
  int32_t num_attn_res;    /* number of entries of attn_res */
  int32_t attn_res[4];     /* resolutions with attention blocks, (16) */
  int32_t resolution;      /* output height = width, 256 */
  int32_t out_ch;          /* 3 */
  int32_t precision;       /* 0: bf16 tensor-core path (fp32 accumulate, fp32 residual stream); 1: fp32 SIMT validation */
  int32_t max_batch;       /* images decoded per pass (workspace size); 0 = default (8) */
} vavae_decoder_config_t;

/* Same conventions as the sampler engine: 0 / non-zero return, vavae_last_error() for the message, no CPU fallback. */
int vavae_decoder_create(const vavae_decoder_config_t* cfg, int device, vavae_decoder_t** out);
int vavae_decoder_destroy(vavae_decoder_t* d);
/* One call per state_dict entry the decode path reads (oracle/vavae_decoder_oracle.py:state_dict_spec lists them). */
int vavae_decoder_load_weight(vavae_decoder_t* d, const char* name, const float* data, const int64_t* shape, int ndim,
                              int is_device);
/* Validates shapes, packs tap-major bf16 convolution weights, allocates the workspace. */
int vavae_decoder_finalize(vavae_decoder_t* d);
/* z: device fp32 (n, z_channels, r, r) NCHW latents, already de-normalised (inference.py:220).  Writes either or both
 * of out_nchw: device fp32 (n, out_ch, R, R) = AutoencoderKL.decode(z), and out_u8_nhwc: device uint8 (n, R, R, out_ch)
 * = VA_VAE.decode_to_images(z).  Enqueued on `stream`, no synchronisation. */
int vavae_decoder_decode(vavae_decoder_t* d, const float* z, int n, float* out_nchw, uint8_t* out_u8_nhwc, void* stream);
uint64_t vavae_decoder_launch_count(vavae_decoder_t* d);
uint64_t vavae_decoder_device_bytes(vavae_decoder_t* d);

#ifdef __cplusplus
}
#endif
#endif /* VAVAE_B200_H */
