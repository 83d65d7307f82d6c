// tcgen05 / TMEM / TMA GEMM with fused epilogues (bf16 path).
#pragma once
#include "common.cuh"

namespace vv {

enum TcEpi {
  TC_EPI_BIAS_F32 = 0,       // out f32 [M, ldo] = acc + bias                       (adaLN, tests)
  TC_EPI_BIAS_BF16 = 1,      // out bf16 = acc + bias
  TC_EPI_GATED_RESIDUAL = 2, // X f32 [M, ldo] += gate[b] * (acc + bias)            (attn.proj, mlp.w3 / fc2)
  TC_EPI_SWIGLU = 3,         // out bf16 [M, ldo][n/2] = silu(v[2j]) * v[2j+1]      (mlp.w12, rows interleaved)
  TC_EPI_GELU_BF16 = 4,      // out bf16 = gelu_tanh(acc + bias)                    (legacy mlp.fc1)
  TC_EPI_QKV = 5,            // bias, [qk-norm], RoPE, scatter q,k (B,H,T,hdp) / v^T (B,H,hdv,T)
  TC_EPI_FINAL = 6,          // acc + bias -> unpatchify -> NCHW rows of the caller  (final_layer.linear)
  TC_EPI_EMBED = 7,          // out f32 [M, ldo] = acc + bias + pos[row % T]          (x_embedder + pos_embed)
  // TC_EPI_GATED_RESIDUAL with the residual tile moved by TMA (loads into a shared-memory ring, in-place update,
  // TMA stores of X and of the next operand).  Chosen by launch_gemm_tc when every tile is interior; never requested
  // by callers.
  TC_EPI_GATED_RESIDUAL_TMA = 8,
};

struct TcGemmArgs {
  const bf16* A;   // [M, K] row-major, row pitch lda elements (lda % 8 == 0)
  int lda;
  const bf16* W;   // [N, K] row-major (nn.Linear layout), row pitch ldw elements
  int ldw;
  int M, N, K;
  int epi;
  const float* bias;  // [N] (already interleaved for SWIGLU) or null
  void* out;          // see TcEpi
  int ldo;
  // gated residual
  const float* mod;   // [rows, mod_ld]
  int mod_ld, gate_off, T;
  // qkv
  bf16 *q, *k, *vt;
  const float *cos, *sin;  // [T, hd] or null
  const float* rope_cs;    // optional [T, hd/2, 2] interleaved (cos, sin) per rotation pair: half the table traffic;
                           // only valid when the (T, hd) tables repeat every frequency twice (the reference's do)
  // optional compact form of the same tables when they are separable over the token grid (the reference's are:
  // columns [0, hd/2) depend on the token's grid row only, [hd/2, hd) on its grid column, pos_embed.py:121-128):
  // rope_tab[0][pos][q] = (cos, sin) of rotation pair q < hd/4 at grid row pos, rope_tab[1][pos][q] = pair hd/4 + q at
  // grid column pos; 2 * rope_gw * (hd/4) * 2 floats, kept in shared memory by the QKV epilogue (no L2 traffic)
  const float* rope_tab;
  int rope_gw;
  const float *qn_w, *qn_b, *kn_w, *kn_b;
  int qk_rms;
  int H, hd, hdp, hdv;
  // embed
  const float* pos;  // [T, N]
  // final
  int p, Cout, Ckeep, img;
  RowMap rm;
  // ---- RMSNorm + modulate fused across GEMMs (requires T % 32 == 0, N % 32 == 0 on the producer) ---------------
  // producer side (TC_EPI_GATED_RESIDUAL): besides x_new, write the NEXT GEMM's operand
  //   nx_A[row, col] = bf16(x_new * nx_w[col] * (1 + mod[b, nx_scale_off + col]))        (un-normalised)
  // and per-row partial sums of x_new^2 in nx_part[row * nx_part_ld + ...].  With s(c, j) = the sum of squares of
  // columns 32c + 4j .. 32c + 4j + 3, a COARSE slot (one float per 128-column group g) is
  //   fold8_j( ((s(4g, j) + s(4g+1, j)) + s(4g+2, j)) + s(4g+3, j) )
  // and the FINE slots of chunk c are the eight floats s(c, 0..7) (row pitch N / 4).  Which of the two a launch writes
  // is gemm_tc_part_coarse(args): coarse needs one warp to own whole groups, i.e. 128 / 256-wide tiles (what every
  // registry model gets at any M); tiles that split a group - hidden sizes that are not a multiple of 128 - write fine
  // slots.  Both are pure functions of the row's values and a consumer folds fine slots into exactly the coarse
  // values, so the tile shape - which depends on M - never changes a result.
  const float* nx_w;
  int nx_scale_off;
  bf16* nx_A;
  float* nx_part;
  int nx_part_ld;
  // consumer side (QKV / SWIGLU / GELU): the accumulator row is scaled by rstd = rsqrt(total / in_dim + 1e-6), total =
  // the row's group sums added in ascending order (in_nslots coarse slots; or, in_fine = 1, in_nslots chunks of fine
  // slots folded into their groups first), and `sp[b, n]` (= shift[b,:] . W[n,:] + bias[n], fp32, row pitch sp_ld)
  // replaces the bias
  const float* in_part;
  int in_part_ld, in_nslots, in_fine, in_dim;
  const float* sp;
  int sp_ld;
  // ---- implicit-GEMM convolution (VA-VAE decoder): A is an NHWC bf16 activation tensor (N_img, H, W, lda channels,
  // lda % 64 == 0) read through a 4-D tensor map; K = conv_taps * lda with tap-major weights [N][tap][lda]; tap t
  // of a 3x3 kernel reads pixel (h + t/3 - 1, w + t%3 - 1), out-of-image pixels are zero-filled by TMA.
  // M = N_img*H*W (H*W % 128 == 0).  conv_taps: 0 = plain GEMM, 1 = 1x1, 9 = 3x3 (stride 1, padding 1).
  int conv_taps, conv_H, conv_W;
  // TC_EPI_BIAS_F32 only: optional fp32 residual [M, ldo] added to acc + bias (may alias `out`)
  const float* res;
  // TC_EPI_BIAS_F32 only, optional: GroupNorm partial statistics of the OUTPUT for the normalisation that follows -
  // per (image, 32-row slab, group) sum and sum of squares: gn_part[((img * (gn_hw/32) + slab) * 32 + group) * 2 + {0,1}]
  // with group = column >> gn_cpg_shift (channels per group a power of two >= 4).  Requires every tile to be interior
  // (M % 128 == 0, N % tile width == 0): launch_gemm_tc checks it.
  float* gn_part;
  int gn_hw, gn_cpg_shift;
  // developer timing experiments on the register-prefetch gated-residual epilogue (VAVAE_GEMM_EXPERIMENT; any value also
  // disables the TMA epilogue).  Results are WRONG with bits 0-2: bit 0: no residual loads, bit 1: no next-operand
  // stores, bit 2: no residual stores; bit 3: residual loads with L1::no_allocate (results unchanged).  0 in production.
  int experiment;
  // Row-tile order: 0 = ascending, 1 = descending.  The engine alternates it from one kernel of the layer stack to the
  // next so that a consumer starts on the rows its producer wrote LAST - the part of a larger-than-L2 tensor that is
  // still cached.  Results do not depend on it (no atomics, fixed partial-sum slots).
  int reverse_m;
  // TC_EPI_GATED_RESIDUAL_TMA only (set by launch_gemm_tc): residual tiles per epilogue warp in the shared-memory ring
  int xt_slots;
  int xt_warps;   // 4 or 8 epilogue warps
  int nx_coarse;  // 1: coarse partial-sum slots (see nx_part)
};

void launch_gemm_tc(const TcGemmArgs& a, cudaStream_t s);
// 1 if this gated-residual launch writes coarse (128-column) partial-sum slots, 0 if fine (32-column) ones
int gemm_tc_part_coarse(const TcGemmArgs& a);
// number of slots (coarse) / chunks (fine) a consumer sums, and the row pitch (floats) either layout needs
inline int gemm_tc_part_slots(int n_cols, int coarse) { return coarse ? (n_cols + 127) / 128 : (n_cols + 31) / 32; }
int gemm_tc_part_ld(int n_cols, int coarse);
// {tile width, last tile width, TMA epilogue?, epilogue warps, ring slots, coarse slots?} of a launch (host only)
void gemm_tc_plan_query(const TcGemmArgs& a, int out[6]);
int gemm_tc_pick_bn(const TcGemmArgs& a);
// 1: use the CTA-pair (cta_group::2) kernel where the problem allows it; 0: always the 1-CTA kernel
void set_gemm_two_cta(int on);
int get_gemm_two_cta();

struct TcAttnArgs {
  const bf16 *q, *k;  // (B,H,T,hdp)
  const bf16* vt;     // (B,H,hdv,T)
  bf16* out;          // [B*T, H*hd]
  int B, H, T, hd, hdp, hdv;
  // developer A/B knobs (VAVAE_ATTN_EXPERIMENT), 0 in production: 1 = no row-max pass (timing only, results wrong),
  // 4 = round-to-nearest P pack (F2FP) instead of truncation, 8 = no early P V issue
  int experiment = 0;
  int reverse = 0;    // 1: work items in descending order (see TcGemmArgs::reverse_m)
};
void launch_attention_tc(const TcAttnArgs& a, cudaStream_t s);
bool pdl_enabled(cudaStream_t s);
// per-device helpers (function attributes and SM counts are per device, not per process)
int current_device();
int num_sms();
// rows of the transposed V tensor: head_dim + one row of ones (row-sum column of P V), rounded up to 16
int attention_hdv(int hd);

}  // namespace vv
