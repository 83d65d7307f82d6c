// Shared declarations for the B200 LightningDiT engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <stdexcept>

typedef __nv_bfloat16 bf16;

namespace vv {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define VV_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw vv::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                      std::to_string(__LINE__));                                                   \
  } while (0)

#define VV_CHECK(cond, msg)                                          \
  do {                                                               \
    if (!(cond)) throw vv::Error(std::string("check failed: ") + msg); \
  } while (0)

// Launch bookkeeping: every kernel launch goes through this so bench.py can report gpu_launches.
struct LaunchCounter {
  uint64_t n = 0;
};
extern thread_local LaunchCounter* g_counter;
inline void count_launch() {
  if (g_counter) g_counter->n++;
}
#define VV_LAUNCH_CHECK()        \
  do {                           \
    vv::count_launch();          \
    VV_CUDA(cudaGetLastError()); \
  } while (0)

// Maps a model row of the current chunk to rows of the caller's tensors.  A chunk has up to two
// segments of `seg_rows` rows: under CFG segment 0 = conditional rows, segment 1 = unconditional
// rows of the SAME samples (both read the same latent rows, lightningdit.py:425-426).
struct RowMap {
  int seg_rows;
  int x_off[2];    // row offset into the latent input
  int y_off[2];    // offset into labels
  int t_off[2];    // offset into t (multiplied by t_stride)
  int out_off[2];  // row offset into the velocity output
  int t_stride;    // 0: every row uses t[0]
};

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- epilogue selectors for the SIMT fp32 GEMM
enum F32Epi { F32_EPI_NONE = 0, F32_EPI_SILU = 1, F32_EPI_GELU_TANH = 2, F32_EPI_ADD_LABEL_SILU = 3 };

struct F32GemmArgs {
  const float* A;  // [M, K]
  int lda;
  const float* W;  // [N, K]
  int ldw;
  const float* bias;  // [N] or null
  float* C;           // [M, ldc]
  int ldc;
  int M, N, K;
  int epi;
  // F32_EPI_ADD_LABEL_SILU: C = bias + acc + table[y[row]] ; also writes silu(C) to C2 (fp32) / C2h (bf16)
  const float* table;
  const int64_t* labels;
  RowMap rm;
  float* C2;
  bf16* C2h;
  // label range check (nn.Embedding raises on an out-of-range index; a kernel cannot): rows of `table`, and a device
  // flag that is set when a label falls outside [0, label_rows) - the read is clamped to row 0, never out of bounds
  int label_rows;
  int* err_flag;
};
void launch_gemm_f32(const F32GemmArgs& a, cudaStream_t s);

// ---- bandwidth kernels (kernels_bw.cu)
struct EmbedArgs {
  const float* x;    // (rows, C, HW, HW)
  const float* wT;   // [C*p*p, D] transposed patch-embed weight
  const float* bias; // [D]
  const float* pos;  // [T, D]
  float* X;          // [rows*T, D]
  int rows, C, img, p, D, T;
  RowMap rm;
};
void launch_patch_embed(const EmbedArgs& a, cudaStream_t s);

void launch_cond_combine(const float* temb, const float* table, const int64_t* labels, RowMap rm, float* sc_f, bf16* sc_h,
                         int rows, int D, int label_rows, int* err_flag, cudaStream_t s);
void launch_patchify_split3(const float* x, bf16* out, int rows, int C, int img, int p, int T, RowMap rm, cudaStream_t s);
void launch_tfreq(const float* t, const float* freqs, float* out, int rows, RowMap rm, cudaStream_t s);

struct NormModArgs {
  const float* X;      // [M, D] fp32 residual stream
  const float* w;      // [D] RMSNorm weight, or null => affine-free LayerNorm(eps 1e-6)
  const float* mod;    // [rows, mod_ld] adaLN output
  int mod_ld;
  int shift_off;       // column offset of shift in mod, or -1
  int scale_off;       // column offset of scale
  void* out;           // [M, ldo] float or bf16
  int ldo;
  int M, D, T;
  int out_bf16;
  int split3;  // bf16 only: write [hi | lo | hi] (3*D columns) so a bf16 GEMM against [Whi | Whi | Wlo]
               // reproduces the fp32 product to ~2^-17 (used for the final layer, whose rounding error
               // would otherwise be amplified by the CFG scale)
};
void launch_norm_modulate(const NormModArgs& a, cudaStream_t s);

struct QkvPostArgs {
  const float* qkv;  // [M, 3D] fp32 (bias already added)
  float *q, *k, *v;  // [B, H, T, hd] fp32
  const float *cos, *sin;      // [T, hd] or null
  const float *qn_w, *qn_b, *kn_w, *kn_b;  // qk-norm params or null
  int rms;           // 1: RMSNorm(eps 1e-6), 0: LayerNorm(eps 1e-5, affine)
  int M, T, H, hd;
};
void launch_qkv_post_f32(const QkvPostArgs& a, cudaStream_t s);
void launch_attention_f32(const float* q, const float* k, const float* v, float* out, int B, int H, int T, int hd,
                          cudaStream_t s);
void launch_gated_residual_f32(float* X, const float* Y, const float* mod, int mod_ld, int gate_off, int M, int D,
                               int T, cudaStream_t s);
void launch_swiglu_f32(const float* U, float* Hh, int M, int S, int ldu, int ldh, cudaStream_t s);

struct UnpatchArgs {
  const float* Y;  // [rows*T, p*p*Cout] token-major final-linear output
  float* out;      // (rows, Ckeep, img, img)
  int rows, T, p, Cout, Ckeep, img;
  RowMap rm;
};
void launch_unpatchify(const UnpatchArgs& a, cudaStream_t s);

struct CfgArgs {
  const float* v;    // (2n, C, HW) model output
  float* out;        // (2n, C, HW)
  const float* t;    // device pointer to t[0]
  int n, C, HW;
  float scale;
  int interval;      // 1: guidance only when t[0] >= start
  float start;
  int channels;
};
void launch_cfg_combine(const CfgArgs& a, cudaStream_t s);

// Fused CFG-combine + ODE update (Appendix A items 6-7 of SURVEY.md).
//   stage 0 (Euler)      : x += dt * vhat
//   stage 1 (Heun pred)  : k1 = vhat ; xp = x + dt * k1
//   stage 2 (Heun corr)  : x += dt/2 * (k1 + vhat)
struct OdeArgs {
  const float* v;  // (rows, C, HW) model output of this evaluation
  float* x;        // state (rows, C, HW)
  float* xp;       // Heun predictor state (stage 1 output, model input of stage 2)
  float* k1;       // Heun first-stage guided velocity
  float* trace;    // optional: guided velocity of this evaluation, (rows, C, HW)
  const float* t;  // device scalar: time of this evaluation
  int rows;        // 2n under CFG, batch otherwise
  int C, HW;
  int use_cfg;
  float scale;
  int interval;
  float start;
  int channels;
  float dt;
  int stage;
  int skip_uncond;  // 1: the unconditional forward of this (gated-off) evaluation was not run - rows >= n of `v` are stale
                    // and rows >= n of the state follow the conditional velocity (they never feed back, lightningdit.py:425)
};
void launch_ode_update(const OdeArgs& a, cudaStream_t s);

// Elementwise update of the SDE sampler (transport/integrators.py:30-49, transport/transport.py:306-334) for the linear
// path with a velocity model: the batch shares one time, so every coefficient is a host scalar.
//   drift(v, x) = v + diff * ((t * v - x) / var)                       (path.py:45-84 with alpha = t, sigma = 1 - t)
// mode 0  Euler-Maruyama : out = (x + drift(v,x) * dt) + s2d * (w * sqdt)
// mode 1  Heun, x-hat    : out = x + s2d * (w * sqdt)
// mode 2  Heun, stage 1  : out2 = k1 = drift(v, x) ; out = x + dt * k1           (x = x-hat)
// mode 3  Heun, stage 2  : out = xh + (0.5 * dt) * (k1 + drift(v, x))            (x = stage-1 output, xh = x-hat)
// mode 4  last step Mean : out = x + drift(v, x) * dt                            (dt = last_step_size)
// mode 5  Tweedie        : out = x / t + ((1-t)^2 / t) * ((t * v - x) / var)
// mode 6  last step Euler: out = x + v * dt
// Every operation is a separately rounded fp32 op in the reference's order (no FMA contraction).
struct SdeArgs {
  int mode;
  const float *x, *v, *w, *k1, *xh;
  float *out, *out2;
  size_t n;
  float t, dt, diff, var, s2d, sqdt, c0, c1;   // c0, c1: Tweedie 1/.. coefficients (alpha, sigma^2/alpha)
};
void launch_sde_update(const SdeArgs& a, cudaStream_t s);

// ---- adaptive RK (dopri5) helpers
struct RkCombineArgs {
  const float* y0;
  const float* k[7];
  float coef[7];   // already multiplied by dt
  int nk;
  float* out;
  size_t n4;       // number of float4 elements
};
void launch_rk_combine(const RkCombineArgs& a, cudaStream_t s);
struct RkNormArgs {
  const float* k[7];
  float coef[7];
  int nk;
  const float* sub;   // optional: subtracted from the combination
  const float* y0;
  const float* y1;    // optional
  float atol, rtol;
  double* partial;    // [blocks]
  int blocks;
  size_t n;
};
void launch_rk_sqnorm(const RkNormArgs& a, double* out, cudaStream_t s);
void launch_rk_interp(const float* y0, const float* y1, const float* ymid, const float* f0, const float* f1, float dt,
                      float x, float* out, size_t n, cudaStream_t s);

void launch_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t s);

}  // namespace vv
