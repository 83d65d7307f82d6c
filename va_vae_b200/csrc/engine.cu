// Engine: weights, workspace, forward / forward_with_cfg / fixed-grid ODE sampler, C ABI (include/vavae_b200.h).
// Host logic only; every numeric op is one of the hand-written kernels in this directory.
#include "../../include/vavae_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>

using namespace vv;

thread_local std::string vv_last_error_buf;   // shared with decoder.cu
#define g_err vv_last_error_buf

struct Weight {
  float* d = nullptr;
  std::vector<int64_t> shape;
  size_t numel = 0;
};

struct LayerBf16 {
  bf16 *qkv = nullptr, *proj = nullptr, *w1 = nullptr, *w2 = nullptr;  // w1: w12 (interleaved) or fc1; w2: w3 or fc2
  float* b1 = nullptr;                                                 // interleaved w12 bias (or fc1 bias)
};

struct vavae_engine {
  vavae_config_t cfg;
  int device = 0;
  // derived sizes
  int D, T, H, hd, S, L, C, p, img, Cout, Ckeep, nmod, NM, Spad, hdp, hdv, grid_hw, Pout;
  int NMx = 0;      // row pitch of `mod`: NM modulation columns + (fused norm) per-layer shift.W^T+bias columns
  int N1 = 0;       // output columns of the first MLP GEMM (2S rounded to 4 for SwiGLU, S for GELU)
  bool fuse = false;  // RMSNorm+modulate folded into the producing / consuming GEMM epilogues
  float *part1 = nullptr, *part2 = nullptr;  // per-row partial sums of squares (TcGemmArgs::nx_part), part_floats each
  size_t part_floats = 0;
  bool fp32;
  bool finalized = false;
  int chunk_rows = 0;
  std::map<std::string, Weight> w;
  std::vector<void*> owned;  // everything cudaMalloc'ed by the engine
  uint64_t bytes = 0;
  LaunchCounter counter;
  int alt_order = 1;         // alternate the row-tile order of consecutive layer kernels (VAVAE_ALT_ORDER=0 disables)
  int order = 0;             // direction of the next layer kernel
  uint64_t fwd_rows = 0;     // model rows pushed through the layer stack since creation (executed forwards)
  int* err_flag = nullptr;   // device flag: a label outside [0, label_rows) was seen (vavae_device_status)
  // live per-kernel profiling (bench.py roofline): CUDA event pairs around every launch of one kernel class
  int prof_kernel = -1;
  std::vector<cudaEvent_t> prof_ev;  // pairs: [2i] start, [2i+1] stop
  size_t prof_used = 0;
  double prof_work = 0.0;

  // packed (both modes)
  float* embed_wT = nullptr;  // [C*p*p, D]
  float* ada_w = nullptr;     // fp32 mode: [NM, D]
  float* ada_b = nullptr;     // [NM]
  // packed (bf16 mode)
  float* rope_cs = nullptr;   // [T, hd/2, 2] interleaved (cos, sin) when the tables repeat each frequency twice
  float* rope_tab = nullptr;  // [2, gw, hd/4, 2] compact form when the tables are also separable over the token grid
  int rope_gw = 0;
  bf16* ada_w_h = nullptr;    // [NM, D]
  bf16* final_w_h = nullptr;  // [Pout, D]
  std::vector<LayerBf16> lw;

  // workspace
  float *X = nullptr, *tf = nullptr, *th = nullptr, *cbuf = nullptr, *mod = nullptr, *sc_f = nullptr;
  bf16* sc_h = nullptr;
  // fp32 path
  float *A_f = nullptr, *QKV_f = nullptr, *q_f = nullptr, *k_f = nullptr, *v_f = nullptr, *O_f = nullptr, *Y_f = nullptr,
        *U_f = nullptr, *Hh_f = nullptr, *Yfin_f = nullptr;
  // bf16 path
  bf16* A3_h = nullptr;  // [M, 3D] split-bf16 operand of the final linear
  bf16* Aemb_h = nullptr;   // [M, 3*C*p*p] split-bf16 patches (A operand of the patch-embed GEMM)
  bf16* embed_w_h = nullptr;  // [D, 3*C*p*p] = [W_hi | W_hi | W_lo]
  bf16 *A_h = nullptr, *q_h = nullptr, *k_h = nullptr, *vt_h = nullptr, *O_h = nullptr, *Hh_h = nullptr;

  // sampler state (lazily sized)
  int state_rows = 0;
  float *sx = nullptr, *sxp = nullptr, *sk1 = nullptr, *sv = nullptr, *st = nullptr;
  int64_t* sy = nullptr;
  float* tgrid_dev = nullptr;
  int tgrid_cap = 0;
  // adaptive RK (dopri5) state: y0, y1, ymid, ytmp + 7 stage derivatives, sized like the sampler state
  int rk_rows = 0;
  float* rk[11] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double* rk_partial = nullptr;
  double* rk_sum = nullptr;
  float* tvec = nullptr;  // [state_rows] per-row t staging for graph-replayed vavae_forward

  // CUDA graphs for the launch-bound small-batch regime: one instantiated graph per (entry point, batch, input buffer)
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    size_t nodes = 0;
    uint64_t rows = 0;   // model rows one replay pushes through the layer stack
  };
  std::map<std::string, GraphEntry> graphs;
  std::set<std::string> graph_warm;
  cudaStream_t gstream = nullptr;
  int graph_max_rows = 128;  // model rows up to which evaluations are replayed from a graph (VAVAE_GRAPH_MAX_ROWS)

  template <typename Tp>
  Tp* alloc(size_t n, bool zero = false) {
    void* ptr = nullptr;
    size_t b = n * sizeof(Tp);
    if (b == 0) b = sizeof(Tp);
    VV_CUDA(cudaMalloc(&ptr, b));
    if (zero) VV_CUDA(cudaMemset(ptr, 0, b));
    owned.push_back(ptr);
    bytes += b;
    return reinterpret_cast<Tp*>(ptr);
  }
  void release(void* ptr) {
    for (auto& o : owned)
      if (o == ptr && ptr) {
        cudaFree(ptr);
        o = nullptr;
      }
  }
  const Weight& W(const std::string& name) const {
    auto it = w.find(name);
    if (it == w.end()) throw Error("missing weight: " + name);
    return it->second;
  }
  const Weight& W(const std::string& name, std::initializer_list<int64_t> shape) const {
    const Weight& x = W(name);
    std::vector<int64_t> s(shape);
    if (x.shape != s) {
      std::string got, want;
      for (auto v : x.shape) got += std::to_string(v) + ",";
      for (auto v : s) want += std::to_string(v) + ",";
      throw Error("weight " + name + " has shape (" + got + ") expected (" + want + ")");
    }
    return x;
  }
  bool has(const std::string& name) const { return w.count(name) != 0; }
};

// --------------------------------------------------------------------------------------------- helpers
enum ProfKernel { PK_QKV = 0, PK_ATTN = 1, PK_PROJ = 2, PK_W12 = 3, PK_W3 = 4, PK_NORM = 5, PK_ODE = 6, PK_COUNT = 7 };
static const char* kProfNames[PK_COUNT] = {"gemm_qkv", "attention", "gemm_proj", "gemm_w12", "gemm_w3", "norm_modulate",
                                           "ode_update"};
constexpr size_t kProfMaxPairs = 1 << 16;

// RAII: records an event pair around one launch when its kernel class is being profiled
struct ProfScope {
  vavae_engine* e;
  cudaStream_t s;
  bool on;
  ProfScope(vavae_engine* e_, int kernel, double work, cudaStream_t s_) : e(e_), s(s_), on(false) {
    if (e->prof_kernel != kernel || e->prof_used >= kProfMaxPairs) return;
    if (e->prof_ev.size() < 2 * (e->prof_used + 1)) {
      cudaEvent_t a, b;
      VV_CUDA(cudaEventCreate(&a));
      VV_CUDA(cudaEventCreate(&b));
      e->prof_ev.push_back(a);
      e->prof_ev.push_back(b);
    }
    VV_CUDA(cudaEventRecord(e->prof_ev[2 * e->prof_used], s));
    e->prof_work += work;
    on = true;
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(e->prof_ev[2 * e->prof_used + 1], s);
    e->prof_used++;
  }
};
static std::string blk(int i, const char* s) { return "blocks." + std::to_string(i) + "." + s; }

__global__ void k_transpose(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int r = idx / cols, c = idx % cols;
  out[(size_t)c * rows + r] = in[idx];
}
// w12 rows interleaved: out row 2j = in row j (x1, gated), 2j+1 = in row S+j (x2); K pitch ldo
__global__ void k_pack_rows_bf16(const float* __restrict__ in, bf16* __restrict__ out, int rows, int K, int ldo,
                                 int interleave_S) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * K) return;
  int r = (int)(idx / K), c = (int)(idx % K);
  int src = r;
  if (interleave_S > 0) src = (r & 1) ? interleave_S + (r >> 1) : (r >> 1);
  out[(size_t)r * ldo + c] = __float2bfloat16_rn(in[(size_t)src * K + c]);
}
// final linear weight as [hi | hi | lo] along K (pairs with the [hi | lo | hi] activation layout)
__global__ void k_pack_split3(const float* __restrict__ in, bf16* __restrict__ out, int rows, int K) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * K) return;
  int r = (int)(idx / K), c = (int)(idx % K);
  float w = in[idx];
  bf16 hi = __float2bfloat16_rn(w);
  bf16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  bf16* o = out + (size_t)r * 3 * K;
  o[c] = hi;
  o[K + c] = hi;
  o[2 * K + c] = lo;
}
__global__ void k_interleave_bias(const float* __restrict__ in, float* __restrict__ out, int S) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 2 * S) return;
  out[r] = in[(r & 1) ? S + (r >> 1) : (r >> 1)];
}
__global__ void k_fill_ones_row(bf16* vt, int BH, int hdv, int row, int T) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BH * T) return;
  int bh = idx / T, t = idx % T;
  vt[((size_t)bh * hdv + row) * T + t] = __float2bfloat16_rn(1.0f);
}

// RoPE tables of the state_dict, (T, hd) cos / sin on the host -> the two packed forms the QKV epilogue reads:
//   cs  [T][hd/2][2]  one (cos, sin) per rotation pair, valid when every frequency is repeated twice (pos_embed.py:123-128)
//   tab [2][gw][hd/4][2]  additionally separable over the gw x gw token grid: first-half pairs depend on the grid row,
//                         second-half pairs on the grid column (pos_embed.py:121-128).  Both are verified bit for bit
//                         against the full tables; an empty vector means "not of that form".
static void pack_rope_tables(const std::vector<float>& hc, const std::vector<float>& hs, int T, int hd, std::vector<float>& cs,
                             std::vector<float>& tab, int& gw) {
  cs.clear(); tab.clear(); gw = 0;
  if (hd % 2) return;
  for (size_t i = 0; i < hc.size(); i += 2)
    if (hc[i] != hc[i + 1] || hs[i] != hs[i + 1]) return;
  cs.resize((size_t)T * hd);
  for (size_t i = 0; i < hc.size(); i += 2) { cs[i] = hc[i]; cs[i + 1] = hs[i]; }
  int g = (int)std::lround(std::sqrt((double)T));
  if (g * g != T || g < 4 || hd % 4) return;
  const int hq = hd / 4, hp = hd / 2;
  std::vector<float> t2((size_t)2 * g * hq * 2);
  for (int pos = 0; pos < g; ++pos)
    for (int q = 0; q < hq; ++q)
      for (int e = 0; e < 2; ++e) {
        t2[((size_t)(0 * g + pos) * hq + q) * 2 + e] = cs[((size_t)(pos * g) * hp + q) * 2 + e];        // token (pos, 0)
        t2[((size_t)(1 * g + pos) * hq + q) * 2 + e] = cs[((size_t)pos * hp + hq + q) * 2 + e];          // token (0, pos)
      }
  for (int t = 0; t < T; ++t)
    for (int r = 0; r < hp; ++r)
      for (int e = 0; e < 2; ++e) {
        const int pos = r < hq ? t / g : t % g, q = r < hq ? r : r - hq;
        if (cs[((size_t)t * hp + r) * 2 + e] != t2[((size_t)((r < hq ? 0 : 1) * g + pos) * hq + q) * 2 + e]) return;
      }
  tab.swap(t2);
  gw = g;
}

static void pack_bf16(vavae_engine* e, const float* src, bf16* dst, int rows, int K, int ldo, int interleave_S,
                      cudaStream_t s) {
  size_t total = (size_t)rows * K;
  k_pack_rows_bf16<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, dst, rows, K, ldo, interleave_S);
  VV_LAUNCH_CHECK();
}

static void derive(vavae_engine* e) {
  const vavae_config_t& c = e->cfg;
  VV_CHECK(c.abi_version == VAVAE_ABI_VERSION, "vavae_config_t.abi_version mismatch");
  VV_CHECK(c.patch_size >= 1 && c.input_size % c.patch_size == 0, "input_size must be divisible by patch_size");
  VV_CHECK(c.hidden_size % c.num_heads == 0, "dim should be divisible by num_heads");  // lightningdit.py:49
  e->D = c.hidden_size;
  e->H = c.num_heads;
  e->hd = e->D / e->H;
  e->L = c.depth;
  e->C = c.in_channels;
  e->p = c.patch_size;
  e->img = c.input_size;
  e->grid_hw = c.input_size / c.patch_size;
  e->T = e->grid_hw * e->grid_hw;
  e->S = c.mlp_hidden;
  e->Cout = c.learn_sigma ? 2 * c.in_channels : c.in_channels;
  e->Ckeep = c.in_channels;
  e->Pout = e->p * e->p * e->Cout;
  e->nmod = c.wo_shift ? 4 : 6;
  e->NM = e->L * e->nmod * e->D + 2 * e->D;
  e->Spad = ((e->S + 7) / 8) * 8;
  e->hdp = ((e->hd + 15) / 16) * 16;
  e->hdv = attention_hdv(e->hd);
  e->fp32 = c.precision == VAVAE_PREC_FP32;
  e->N1 = c.use_swiglu ? ((2 * e->S + 3) / 4) * 4 : e->S;
  e->fuse = !e->fp32 && c.use_rmsnorm && (e->T % 32 == 0);
  e->NMx = e->NM + ((e->fuse && !c.wo_shift) ? e->L * (3 * e->D + e->N1) : 0);
  VV_CHECK(c.precision == VAVAE_PREC_FP32 || c.precision == VAVAE_PREC_BF16, "unknown precision mode");
  VV_CHECK(e->D % 4 == 0 && e->D <= 2048, "hidden_size must be a multiple of 4 and <= 2048");
  VV_CHECK(e->hd % 2 == 0, "head_dim must be even (RoPE pairs)");
  VV_CHECK((e->img * e->img) % 4 == 0, "latent H*W must be a multiple of 4");
  if (!e->fp32) {
    VV_CHECK(e->D % 16 == 0 && e->hd % 8 == 0 && e->hd <= 112, "bf16 path: hidden % 16, head_dim % 8, head_dim <= 112");
    VV_CHECK(e->T % 8 == 0, "bf16 path: token count must be a multiple of 8");
  }
}

// --------------------------------------------------------------------------------------------- finalize
static void gemm32(const float* A, int lda, const float* Wt, const float* bias, float* Cc, int ldc, int M, int N, int K,
                   int epi, cudaStream_t s);
static void finalize(vavae_engine* e, cudaStream_t s) {
  const int D = e->D, L = e->L, S = e->S, C = e->C, p = e->p, T = e->T, hd = e->hd;
  const vavae_config_t& c = e->cfg;
  // ---- presence / shape validation of the reference state_dict
  e->W("pos_embed", {1, T, D});
  e->W("x_embedder.proj.weight", {D, C, p, p});
  e->W("x_embedder.proj.bias", {D});
  e->W("t_embedder.mlp.0.weight", {D, 256});
  e->W("t_embedder.mlp.0.bias", {D});
  e->W("t_embedder.mlp.2.weight", {D, D});
  e->W("t_embedder.mlp.2.bias", {D});
  e->W("t_embedder.freqs", {128});
  e->W("y_embedder.embedding_table.weight", {c.label_rows, D});
  if (c.use_rope) {
    e->W("feat_rope.freqs_cos", {T, hd});
    e->W("feat_rope.freqs_sin", {T, hd});
  }
  for (int i = 0; i < L; ++i) {
    if (c.use_rmsnorm) {
      e->W(blk(i, "norm1.weight"), {D});
      e->W(blk(i, "norm2.weight"), {D});
    }
    e->W(blk(i, "attn.qkv.weight"), {3 * D, D});
    e->W(blk(i, "attn.qkv.bias"), {3 * D});
    if (c.use_qknorm) {
      e->W(blk(i, "attn.q_norm.weight"), {hd});
      e->W(blk(i, "attn.k_norm.weight"), {hd});
      if (!c.use_rmsnorm) {
        e->W(blk(i, "attn.q_norm.bias"), {hd});
        e->W(blk(i, "attn.k_norm.bias"), {hd});
      }
    }
    e->W(blk(i, "attn.proj.weight"), {D, D});
    e->W(blk(i, "attn.proj.bias"), {D});
    if (c.use_swiglu) {
      e->W(blk(i, "mlp.w12.weight"), {2 * S, D});
      e->W(blk(i, "mlp.w12.bias"), {2 * S});
      e->W(blk(i, "mlp.w3.weight"), {D, S});
      e->W(blk(i, "mlp.w3.bias"), {D});
    } else {
      e->W(blk(i, "mlp.fc1.weight"), {S, D});
      e->W(blk(i, "mlp.fc1.bias"), {S});
      e->W(blk(i, "mlp.fc2.weight"), {D, S});
      e->W(blk(i, "mlp.fc2.bias"), {D});
    }
    e->W(blk(i, "adaLN_modulation.1.weight"), {e->nmod * D, D});
    e->W(blk(i, "adaLN_modulation.1.bias"), {e->nmod * D});
  }
  if (c.use_rmsnorm) e->W("final_layer.norm_final.weight", {D});
  e->W("final_layer.linear.weight", {e->Pout, D});
  e->W("final_layer.linear.bias", {e->Pout});
  e->W("final_layer.adaLN_modulation.1.weight", {2 * D, D});
  e->W("final_layer.adaLN_modulation.1.bias", {2 * D});

  // ---- packing shared by both modes
  const int Cpp = C * p * p;
  e->embed_wT = e->alloc<float>((size_t)Cpp * D);
  k_transpose<<<cdiv(D * Cpp, 256), 256, 0, s>>>(e->W("x_embedder.proj.weight").d, e->embed_wT, D, Cpp);
  VV_LAUNCH_CHECK();
  // batched adaLN: one [NM, D] matrix = every block's modulation Linear followed by the final layer's
  e->ada_b = e->alloc<float>(e->NMx, true);
  if (e->fp32) e->ada_w = e->alloc<float>((size_t)e->NMx * D);
  else e->ada_w_h = e->alloc<bf16>((size_t)e->NMx * D, true);
  for (int i = 0; i <= L; ++i) {
    const bool fin = i == L;
    const Weight& aw = e->W(fin ? std::string("final_layer.adaLN_modulation.1.weight") : blk(i, "adaLN_modulation.1.weight"));
    const Weight& ab = e->W(fin ? std::string("final_layer.adaLN_modulation.1.bias") : blk(i, "adaLN_modulation.1.bias"));
    const size_t row0 = (size_t)i * e->nmod * D;
    const int rows = fin ? 2 * D : e->nmod * D;
    VV_CUDA(cudaMemcpyAsync(e->ada_b + row0, ab.d, (size_t)rows * 4, cudaMemcpyDeviceToDevice, s));
    if (e->fp32) VV_CUDA(cudaMemcpyAsync(e->ada_w + row0 * D, aw.d, (size_t)rows * D * 4, cudaMemcpyDeviceToDevice, s));
    else pack_bf16(e, aw.d, e->ada_w_h + row0 * D, rows, D, D, 0, s);
  }
  if (e->fuse && !c.wo_shift) {
    // Fused-norm consumers need sp[b,n] = shift[b,:].W[n,:] + bias[n].  shift = silu(c).Wa_shift^T + ba_shift, so
    // sp = silu(c).(W.Wa_shift)^T + (W.ba_shift + bias): extra output columns of the single adaLN GEMM.
    const int N1 = e->N1;
    const int maxN = std::max(3 * D, 2 * S);
    float* waT = e->alloc<float>((size_t)D * D);
    float* wc = e->alloc<float>((size_t)maxN * D);
    float* bc = e->alloc<float>(maxN);
    for (int i = 0; i < L; ++i) {
      const float* Wa = e->W(blk(i, "adaLN_modulation.1.weight")).d;   // rows: shift_msa | scale_msa | gate_msa | shift_mlp | ...
      const float* ba = e->W(blk(i, "adaLN_modulation.1.bias")).d;
      const size_t row0 = (size_t)e->NM + (size_t)i * (3 * D + N1);
      for (int which = 0; which < 2; ++which) {
        const int chunk = which == 0 ? 0 : 3;                           // shift_msa, shift_mlp (lightningdit.py:246)
        const float* Wl = which == 0 ? e->W(blk(i, "attn.qkv.weight")).d
                                     : e->W(blk(i, c.use_swiglu ? "mlp.w12.weight" : "mlp.fc1.weight")).d;
        const float* bl = which == 0 ? e->W(blk(i, "attn.qkv.bias")).d
                                     : e->W(blk(i, c.use_swiglu ? "mlp.w12.bias" : "mlp.fc1.bias")).d;
        const int Nl = which == 0 ? 3 * D : (c.use_swiglu ? 2 * S : S);
        k_transpose<<<cdiv(D * D, 256), 256, 0, s>>>(Wa + (size_t)chunk * D * D, waT, D, D);
        VV_LAUNCH_CHECK();
        gemm32(Wl, D, waT, nullptr, wc, D, Nl, D, D, F32_EPI_NONE, s);                    // wc[n,j] = sum_k W[n,k] Wa_shift[k,j]
        gemm32(ba + (size_t)chunk * D, D, Wl, bl, bc, Nl, 1, Nl, D, F32_EPI_NONE, s);    // bc[n] = W[n,:].ba_shift + bias[n]
        const size_t r0 = row0 + (which == 0 ? 0 : 3 * D);
        const bool il = which == 1 && c.use_swiglu;
        pack_bf16(e, wc, e->ada_w_h + r0 * D, Nl, D, D, il ? S : 0, s);
        if (il) {
          k_interleave_bias<<<cdiv(2 * S, 256), 256, 0, s>>>(bc, e->ada_b + r0, S);
          VV_LAUNCH_CHECK();
        } else {
          VV_CUDA(cudaMemcpyAsync(e->ada_b + r0, bc, (size_t)Nl * 4, cudaMemcpyDeviceToDevice, s));
        }
      }
    }
    VV_CUDA(cudaStreamSynchronize(s));
    e->release(waT); e->release(wc); e->release(bc);
  }
  if (!e->fp32 && c.use_rope && hd % 2 == 0) {
    // RoPE tables of the reference repeat every frequency twice (pos_embed.py:123-128): if this state_dict's do too,
    // keep one (cos, sin) pair per rotation pair -> half the table traffic in the QKV epilogue
    std::vector<float> hc((size_t)T * hd), hs((size_t)T * hd), cs, tab;
    VV_CUDA(cudaMemcpy(hc.data(), e->W("feat_rope.freqs_cos").d, hc.size() * 4, cudaMemcpyDeviceToHost));
    VV_CUDA(cudaMemcpy(hs.data(), e->W("feat_rope.freqs_sin").d, hs.size() * 4, cudaMemcpyDeviceToHost));
    pack_rope_tables(hc, hs, T, hd, cs, tab, e->rope_gw);
    if (!cs.empty()) {
      e->rope_cs = e->alloc<float>(cs.size());
      VV_CUDA(cudaMemcpy(e->rope_cs, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
    }
    if (!tab.empty()) {
      e->rope_tab = e->alloc<float>(tab.size());
      VV_CUDA(cudaMemcpy(e->rope_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  if (!e->fp32) {
    e->lw.resize(L);
    for (int i = 0; i < L; ++i) {
      LayerBf16& lw = e->lw[i];
      lw.qkv = e->alloc<bf16>((size_t)3 * D * D);
      pack_bf16(e, e->W(blk(i, "attn.qkv.weight")).d, lw.qkv, 3 * D, D, D, 0, s);
      lw.proj = e->alloc<bf16>((size_t)D * D);
      pack_bf16(e, e->W(blk(i, "attn.proj.weight")).d, lw.proj, D, D, D, 0, s);
      if (c.use_swiglu) {
        // N is rounded up to a multiple of 4 with zero rows / zero bias so the epilogue can work on float4 groups
        const int N12 = ((2 * S + 3) / 4) * 4;
        lw.w1 = e->alloc<bf16>((size_t)N12 * D, true);
        pack_bf16(e, e->W(blk(i, "mlp.w12.weight")).d, lw.w1, 2 * S, D, D, S, s);
        lw.b1 = e->alloc<float>(N12, true);
        k_interleave_bias<<<cdiv(2 * S, 256), 256, 0, s>>>(e->W(blk(i, "mlp.w12.bias")).d, lw.b1, S);
        VV_LAUNCH_CHECK();
        lw.w2 = e->alloc<bf16>((size_t)D * e->Spad, true);
        pack_bf16(e, e->W(blk(i, "mlp.w3.weight")).d, lw.w2, D, S, e->Spad, 0, s);
      } else {
        lw.w1 = e->alloc<bf16>((size_t)S * D);
        pack_bf16(e, e->W(blk(i, "mlp.fc1.weight")).d, lw.w1, S, D, D, 0, s);
        lw.b1 = e->W(blk(i, "mlp.fc1.bias")).d;
        lw.w2 = e->alloc<bf16>((size_t)D * e->Spad, true);
        pack_bf16(e, e->W(blk(i, "mlp.fc2.weight")).d, lw.w2, D, S, e->Spad, 0, s);
      }
    }
    e->embed_w_h = e->alloc<bf16>((size_t)D * 3 * Cpp);
    k_pack_split3<<<cdiv(D * Cpp, 256), 256, 0, s>>>(e->W("x_embedder.proj.weight").d, e->embed_w_h, D, Cpp);
    VV_LAUNCH_CHECK();
    e->final_w_h = e->alloc<bf16>((size_t)e->Pout * 3 * D);
    k_pack_split3<<<cdiv(e->Pout * D, 256), 256, 0, s>>>(e->W("final_layer.linear.weight").d, e->final_w_h, e->Pout, D);
    VV_LAUNCH_CHECK();
    VV_CUDA(cudaStreamSynchronize(s));
    // the fp32 originals of the big matrices are no longer needed on the bf16 path
    for (int i = 0; i < L; ++i) {
      std::vector<std::string> big = {"attn.qkv.weight", "attn.proj.weight", "adaLN_modulation.1.weight"};
      if (c.use_swiglu) { big.push_back("mlp.w12.weight"); big.push_back("mlp.w3.weight"); }
      else { big.push_back("mlp.fc1.weight"); big.push_back("mlp.fc2.weight"); }
      for (auto& nme : big) {
        Weight& x = e->w[blk(i, nme.c_str())];
        e->release(x.d);
        e->bytes -= x.numel * 4;
        x.d = nullptr;
      }
    }
  }

  // ---- workspace
  int R = c.max_chunk_rows > 0 ? c.max_chunk_rows : (e->fp32 ? 16 : 512);
  if (R % 2) R += 1;
  e->chunk_rows = R;
  const size_t M = (size_t)R * T;
  e->X = e->alloc<float>(M * D);
  e->tf = e->alloc<float>((size_t)R * 256);
  e->th = e->alloc<float>((size_t)R * D);
  e->cbuf = e->alloc<float>((size_t)R * D);
  e->mod = e->alloc<float>((size_t)R * e->NMx);
  e->st = e->alloc<float>(4);
  e->err_flag = e->alloc<int>(1, true);
  if (e->fp32) {
    e->sc_f = e->alloc<float>((size_t)R * D);
    e->A_f = e->alloc<float>(M * D);
    e->QKV_f = e->alloc<float>(M * 3 * D);
    e->q_f = e->alloc<float>(M * D);
    e->k_f = e->alloc<float>(M * D);
    e->v_f = e->alloc<float>(M * D);
    e->O_f = e->alloc<float>(M * D);
    e->Y_f = e->alloc<float>(M * D);
    if (c.use_swiglu) e->U_f = e->alloc<float>(M * 2 * S);
    e->Hh_f = e->alloc<float>(M * S);
    e->Yfin_f = e->alloc<float>(M * e->Pout);
  } else {
    e->sc_h = e->alloc<bf16>((size_t)R * D);
    e->A_h = e->alloc<bf16>(M * D);
    e->A3_h = e->alloc<bf16>(M * 3 * D);
    e->Aemb_h = e->alloc<bf16>(M * 3 * e->C * e->p * e->p);
    if (e->fuse) {
      // coarse slots: 16 floats per row at any M; fine slots (D / 4 floats per row) only arise on the narrow tiles of
      // small-M launches - or at every M when D is not a multiple of 128 (no registry model)
      const size_t fine_rows = (D % 128) ? (size_t)M : std::min<size_t>((size_t)M, 4096);
      e->part_floats = std::max<size_t>((size_t)16 * M, fine_rows * (size_t)(D / 4));
      e->part1 = e->alloc<float>(e->part_floats, true);
      e->part2 = e->alloc<float>(e->part_floats, true);
    }
    // padding columns d in [hd, hdp) of q/k and rows [hd, hdv) of v^T stay zero forever (never written)
    e->q_h = e->alloc<bf16>((size_t)R * e->H * T * e->hdp, true);
    e->k_h = e->alloc<bf16>((size_t)R * e->H * T * e->hdp, true);
    e->vt_h = e->alloc<bf16>((size_t)R * e->H * e->hdv * T, true);
    // row hd of every (row, head) block of v^T is all ones: column hd of P V then is the softmax row sum
    k_fill_ones_row<<<cdiv(R * e->H * T, 256), 256, 0, s>>>(e->vt_h, R * e->H, e->hdv, e->hd, T);
    VV_LAUNCH_CHECK();
    e->O_h = e->alloc<bf16>(M * D);
    e->Hh_h = e->alloc<bf16>(M * e->Spad, true);
  }
  VV_CUDA(cudaStreamSynchronize(s));
  e->finalized = true;
}

// --------------------------------------------------------------------------------------------- forward (one chunk)
struct ModOff {
  int shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp;
};
static ModOff mod_offsets(const vavae_engine* e, int layer) {
  const int D = e->D, base = layer * e->nmod * D;
  ModOff o;
  if (e->cfg.wo_shift) {  // lightningdit.py:242-244
    o.shift_msa = -1; o.scale_msa = base; o.gate_msa = base + D;
    o.shift_mlp = -1; o.scale_mlp = base + 2 * D; o.gate_mlp = base + 3 * D;
  } else {  // lightningdit.py:246
    o.shift_msa = base; o.scale_msa = base + D; o.gate_msa = base + 2 * D;
    o.shift_mlp = base + 3 * D; o.scale_mlp = base + 4 * D; o.gate_mlp = base + 5 * D;
  }
  return o;
}

static void norm_mod(vavae_engine* e, const float* w, int shift_off, int scale_off, int M, cudaStream_t s,
                     bool split3 = false) {
  NormModArgs n;
  n.X = e->X; n.w = w; n.mod = e->mod; n.mod_ld = e->NMx; n.shift_off = shift_off; n.scale_off = scale_off;
  n.out = e->fp32 ? (void*)e->A_f : (split3 ? (void*)e->A3_h : (void*)e->A_h);
  n.ldo = split3 ? 3 * e->D : e->D; n.M = M; n.D = e->D; n.T = e->T;
  n.out_bf16 = e->fp32 ? 0 : 1;
  n.split3 = split3 ? 1 : 0;
  launch_norm_modulate(n, s);
}

static void gemm32(const float* A, int lda, const float* Wt, const float* bias, float* Cc, int ldc, int M, int N, int K,
                   int epi, cudaStream_t s) {
  F32GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.W = Wt; g.ldw = K; g.bias = bias; g.C = Cc; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.epi = epi;
  launch_gemm_f32(g, s);
}

static TcGemmArgs tc_args(const bf16* A, int lda, const bf16* Wt, int ldw, int M, int N, int K, int epi,
                          const float* bias) {
  TcGemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.W = Wt; g.ldw = ldw; g.M = M; g.N = N; g.K = K; g.epi = epi; g.bias = bias;
  return g;
}

static void conditioning(vavae_engine* e, int rows, const RowMap& rm, const float* t, const int64_t* y, cudaStream_t s) {
  const int D = e->D;
  if (rm.t_stride == 0) {
    // the sampler's case: one timestep for the whole batch -> run the timestep MLP on a single row
    RowMap r1 = rm;
    r1.seg_rows = 1;
    r1.t_off[0] = r1.t_off[1] = 0;
    launch_tfreq(t, e->W("t_embedder.freqs").d, e->tf, 1, r1, s);
    gemm32(e->tf, 256, e->W("t_embedder.mlp.0.weight").d, e->W("t_embedder.mlp.0.bias").d, e->th, D, 1, D, 256,
           F32_EPI_SILU, s);
    gemm32(e->th, D, e->W("t_embedder.mlp.2.weight").d, e->W("t_embedder.mlp.2.bias").d, e->cbuf, D, 1, D, D,
           F32_EPI_NONE, s);
    launch_cond_combine(e->cbuf, e->W("y_embedder.embedding_table.weight").d, y, rm, e->fp32 ? e->sc_f : nullptr,
                        e->fp32 ? nullptr : e->sc_h, rows, D, e->cfg.label_rows, e->err_flag, s);
  } else {
  launch_tfreq(t, e->W("t_embedder.freqs").d, e->tf, rows, rm, s);
  gemm32(e->tf, 256, e->W("t_embedder.mlp.0.weight").d, e->W("t_embedder.mlp.0.bias").d, e->th, D, rows, D, 256,
         F32_EPI_SILU, s);
  F32GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = e->th; g.lda = D; g.W = e->W("t_embedder.mlp.2.weight").d; g.ldw = D; g.bias = e->W("t_embedder.mlp.2.bias").d;
  g.C = e->cbuf; g.ldc = D; g.M = rows; g.N = D; g.K = D; g.epi = F32_EPI_ADD_LABEL_SILU;
  g.table = e->W("y_embedder.embedding_table.weight").d; g.labels = y; g.rm = rm;
  g.C2 = e->fp32 ? e->sc_f : nullptr; g.C2h = e->fp32 ? nullptr : e->sc_h;
  g.label_rows = e->cfg.label_rows; g.err_flag = e->err_flag;
  launch_gemm_f32(g, s);
  }
  // every adaLN modulation of the network in one GEMM: mod[rows, NM]
  if (e->fp32) {
    gemm32(e->sc_f, D, e->ada_w, e->ada_b, e->mod, e->NMx, rows, e->NMx, D, F32_EPI_NONE, s);
  } else {
    TcGemmArgs a = tc_args(e->sc_h, D, e->ada_w_h, D, rows, e->NMx, D, TC_EPI_BIAS_F32, e->ada_b);
    a.out = e->mod; a.ldo = e->NMx;
    launch_gemm_tc(a, s);
  }
}

// Direction of the next layer kernel: consecutive kernels walk the rows in opposite directions, so each one starts on
// the rows its predecessor produced last (the tail of a larger-than-L2 tensor that is still cached).
static int next_order(vavae_engine* e) {
  if (!e->alt_order) return 0;
  e->order ^= 1;
  return e->order;
}

static void forward_chunk(vavae_engine* e, int rows, const RowMap& rm, const float* x, const float* t, const int64_t* y,
                          float* out, cudaStream_t s) {
  const int D = e->D, T = e->T, H = e->H, hd = e->hd, S = e->S, L = e->L;
  const int M = rows * T;
  const vavae_config_t& c = e->cfg;
  VV_CHECK(rows <= e->chunk_rows, "forward_chunk: chunk larger than workspace");
  e->fwd_rows += (uint64_t)rows;
  e->order = 0;   // every forward starts from the same state: captured graphs and eager launches behave identically
  const int Cpp = e->C * e->p * e->p;
  if (!e->fp32 && (3 * Cpp) % 8 == 0) {
    // patch embed on the tensor cores: split-bf16 patches x [W_hi | W_hi | W_lo], fp32-grade; epilogue adds bias + pos
    launch_patchify_split3(x, e->Aemb_h, rows, e->C, e->img, e->p, T, rm, s);
    TcGemmArgs a = tc_args(e->Aemb_h, 3 * Cpp, e->embed_w_h, 3 * Cpp, M, D, 3 * Cpp, TC_EPI_EMBED, e->W("x_embedder.proj.bias").d);
    a.out = e->X; a.ldo = D; a.T = T; a.pos = e->W("pos_embed").d;
    launch_gemm_tc(a, s);
  } else {
    EmbedArgs ea;
    ea.x = x; ea.wT = e->embed_wT; ea.bias = e->W("x_embedder.proj.bias").d; ea.pos = e->W("pos_embed").d; ea.X = e->X;
    ea.rows = rows; ea.C = e->C; ea.img = e->img; ea.p = e->p; ea.D = D; ea.T = T; ea.rm = rm;
    launch_patch_embed(ea, s);
  }
  conditioning(e, rows, rm, t, y, s);
  const float* rcos = c.use_rope ? e->W("feat_rope.freqs_cos").d : nullptr;
  const float* rsin = c.use_rope ? e->W("feat_rope.freqs_sin").d : nullptr;

  int pc_proj = 0, pc_w3 = 0;  // kind of partial-sum slots (1 coarse / 0 fine) the last proj / w3 epilogue wrote
  for (int i = 0; i < L; ++i) {
    const ModOff mo = mod_offsets(e, i);
    const float* n1 = c.use_rmsnorm ? e->W(blk(i, "norm1.weight")).d : nullptr;
    const float* n2 = c.use_rmsnorm ? e->W(blk(i, "norm2.weight")).d : nullptr;
    const float *qnw = nullptr, *qnb = nullptr, *knw = nullptr, *knb = nullptr;
    if (c.use_qknorm) {
      qnw = e->W(blk(i, "attn.q_norm.weight")).d;
      knw = e->W(blk(i, "attn.k_norm.weight")).d;
      if (!c.use_rmsnorm) {
        qnb = e->W(blk(i, "attn.q_norm.bias")).d;
        knb = e->W(blk(i, "attn.k_norm.bias")).d;
      }
    }
    // ---- attention branch
    // With the fused norm (bf16 + RMSNorm + T % 32 == 0) only the first norm of the network runs as a kernel: every
    // gated-residual epilogue emits the next GEMM's un-normalised bf16 operand plus per-row partial sums of
    // squares, and the consuming GEMM applies 1/rms per row and the per-sample shift.W^T + bias vector.
    const bool fuse = e->fuse;
    const int sp_stride = 3 * D + e->N1;
    const float* sp_qkv = (fuse && !c.wo_shift) ? e->mod + e->NM + (size_t)i * sp_stride : nullptr;
    const float* sp_mlp = (fuse && !c.wo_shift) ? sp_qkv + 3 * D : nullptr;
    if (!fuse || i == 0) {
      ProfScope ps(e, PK_NORM, (double)M * D * (e->fp32 ? 8.0 : 6.0), s);
      norm_mod(e, n1, mo.shift_msa, mo.scale_msa, M, s);
    }
    if (e->fp32) {
      gemm32(e->A_f, D, e->W(blk(i, "attn.qkv.weight")).d, e->W(blk(i, "attn.qkv.bias")).d, e->QKV_f, 3 * D, M, 3 * D, D,
             F32_EPI_NONE, s);
      QkvPostArgs qa;
      qa.qkv = e->QKV_f; qa.q = e->q_f; qa.k = e->k_f; qa.v = e->v_f; qa.cos = rcos; qa.sin = rsin;
      qa.qn_w = qnw; qa.qn_b = qnb; qa.kn_w = knw; qa.kn_b = knb; qa.rms = c.use_rmsnorm; qa.M = M; qa.T = T; qa.H = H;
      qa.hd = hd;
      launch_qkv_post_f32(qa, s);
      launch_attention_f32(e->q_f, e->k_f, e->v_f, e->O_f, rows, H, T, hd, s);
      gemm32(e->O_f, D, e->W(blk(i, "attn.proj.weight")).d, e->W(blk(i, "attn.proj.bias")).d, e->Y_f, D, M, D, D,
             F32_EPI_NONE, s);
      launch_gated_residual_f32(e->X, e->Y_f, e->mod, e->NMx, mo.gate_msa, M, D, T, s);
    } else {
      const float* bqkv = e->W(blk(i, "attn.qkv.bias")).d;
      TcGemmArgs a = tc_args(e->A_h, D, e->lw[i].qkv, D, M, 3 * D, D, TC_EPI_QKV, bqkv);
      a.q = e->q_h; a.k = e->k_h; a.vt = e->vt_h; a.cos = rcos; a.sin = rsin; a.rope_cs = e->rope_cs; a.rope_tab = e->rope_tab; a.rope_gw = e->rope_gw; a.qn_w = qnw; a.qn_b = qnb; a.kn_w = knw;
      a.kn_b = knb; a.qk_rms = c.use_rmsnorm; a.H = H; a.hd = hd; a.hdp = e->hdp; a.hdv = e->hdv; a.T = T;
      if (fuse && i > 0) {
        a.in_part = e->part2; a.in_fine = !pc_w3; a.in_nslots = gemm_tc_part_slots(D, pc_w3); a.in_part_ld = gemm_tc_part_ld(D, pc_w3); a.in_dim = D;
        a.sp = sp_qkv ? sp_qkv : bqkv; a.sp_ld = sp_qkv ? e->NMx : 0;
      }
      {
        ProfScope ps(e, PK_QKV, 2.0 * M * 3.0 * D * D, s);
        a.reverse_m = next_order(e);
        launch_gemm_tc(a, s);
      }
      TcAttnArgs at;
      at.q = e->q_h; at.k = e->k_h; at.vt = e->vt_h; at.out = e->O_h; at.B = rows; at.H = H; at.T = T; at.hd = hd;
      at.hdp = e->hdp; at.hdv = e->hdv;
      {
        ProfScope ps(e, PK_ATTN, 4.0 * rows * (double)H * T * T * hd, s);
        at.reverse = next_order(e);
        launch_attention_tc(at, s);
      }
      TcGemmArgs pr = tc_args(e->O_h, D, e->lw[i].proj, D, M, D, D, TC_EPI_GATED_RESIDUAL, e->W(blk(i, "attn.proj.bias")).d);
      pr.out = e->X; pr.ldo = D; pr.mod = e->mod; pr.mod_ld = e->NMx; pr.gate_off = mo.gate_msa; pr.T = T;
      if (fuse) {  // emit the operand of this block's MLP norm (norm2)
        pr.nx_w = n2; pr.nx_scale_off = mo.scale_mlp; pr.nx_A = e->A_h; pr.nx_part = e->part1;
        pc_proj = gemm_tc_part_coarse(pr);
        pr.nx_part_ld = gemm_tc_part_ld(D, pc_proj);
        VV_CHECK((size_t)M * pr.nx_part_ld <= e->part_floats, "partial-sum buffer too small for this batch");
      }
      {
        ProfScope ps(e, PK_PROJ, 2.0 * M * (double)D * D, s);
        pr.reverse_m = next_order(e);
        launch_gemm_tc(pr, s);
      }
    }
    // ---- MLP branch
    if (!fuse) {
      ProfScope ps(e, PK_NORM, (double)M * D * (e->fp32 ? 8.0 : 6.0), s);
      norm_mod(e, n2, mo.shift_mlp, mo.scale_mlp, M, s);
    }
    if (e->fp32) {
      if (c.use_swiglu) {
        gemm32(e->A_f, D, e->W(blk(i, "mlp.w12.weight")).d, e->W(blk(i, "mlp.w12.bias")).d, e->U_f, 2 * S, M, 2 * S, D,
               F32_EPI_NONE, s);
        launch_swiglu_f32(e->U_f, e->Hh_f, M, S, 2 * S, S, s);
        gemm32(e->Hh_f, S, e->W(blk(i, "mlp.w3.weight")).d, e->W(blk(i, "mlp.w3.bias")).d, e->Y_f, D, M, D, S,
               F32_EPI_NONE, s);
      } else {
        gemm32(e->A_f, D, e->W(blk(i, "mlp.fc1.weight")).d, e->W(blk(i, "mlp.fc1.bias")).d, e->Hh_f, S, M, S, D,
               F32_EPI_GELU_TANH, s);
        gemm32(e->Hh_f, S, e->W(blk(i, "mlp.fc2.weight")).d, e->W(blk(i, "mlp.fc2.bias")).d, e->Y_f, D, M, D, S,
               F32_EPI_NONE, s);
      }
      launch_gated_residual_f32(e->X, e->Y_f, e->mod, e->NMx, mo.gate_mlp, M, D, T, s);
    } else {
      {
        TcGemmArgs a = tc_args(e->A_h, D, e->lw[i].w1, D, M, e->N1, D, c.use_swiglu ? TC_EPI_SWIGLU : TC_EPI_GELU_BF16,
                               e->lw[i].b1);
        a.out = e->Hh_h; a.ldo = e->Spad; a.T = T;
        if (fuse) {
          a.in_part = e->part1; a.in_fine = !pc_proj; a.in_nslots = gemm_tc_part_slots(D, pc_proj); a.in_part_ld = gemm_tc_part_ld(D, pc_proj); a.in_dim = D;
          a.sp = sp_mlp ? sp_mlp : e->lw[i].b1; a.sp_ld = sp_mlp ? e->NMx : 0;
        }
        ProfScope ps(e, PK_W12, 2.0 * M * (c.use_swiglu ? 2.0 : 1.0) * S * D, s);
        a.reverse_m = next_order(e);
        launch_gemm_tc(a, s);
      }
      const float* b2 = e->W(blk(i, c.use_swiglu ? "mlp.w3.bias" : "mlp.fc2.bias")).d;
      TcGemmArgs a = tc_args(e->Hh_h, e->Spad, e->lw[i].w2, e->Spad, M, D, S, TC_EPI_GATED_RESIDUAL, b2);
      a.out = e->X; a.ldo = D; a.mod = e->mod; a.mod_ld = e->NMx; a.gate_off = mo.gate_mlp; a.T = T;
      if (fuse && i + 1 < L) {  // emit the operand of the next block's attention norm (norm1)
        const ModOff nmo = mod_offsets(e, i + 1);
        a.nx_w = e->W(blk(i + 1, "norm1.weight")).d; a.nx_scale_off = nmo.scale_msa; a.nx_A = e->A_h; a.nx_part = e->part2;
        pc_w3 = gemm_tc_part_coarse(a);
        a.nx_part_ld = gemm_tc_part_ld(D, pc_w3);
        VV_CHECK((size_t)M * a.nx_part_ld <= e->part_floats, "partial-sum buffer too small for this batch");
      }
      {
        ProfScope ps(e, PK_W3, 2.0 * M * (double)S * D, s);
        a.reverse_m = next_order(e);
        launch_gemm_tc(a, s);
      }
    }
  }
  // ---- final layer (lightningdit.py:267-272): shift first, then scale
  const int fbase = L * e->nmod * D;
  norm_mod(e, c.use_rmsnorm ? e->W("final_layer.norm_final.weight").d : nullptr, fbase, fbase + D, M, s, !e->fp32);
  if (e->fp32) {
    gemm32(e->A_f, D, e->W("final_layer.linear.weight").d, e->W("final_layer.linear.bias").d, e->Yfin_f, e->Pout, M,
           e->Pout, D, F32_EPI_NONE, s);
    UnpatchArgs u;
    u.Y = e->Yfin_f; u.out = out; u.rows = rows; u.T = T; u.p = e->p; u.Cout = e->Cout; u.Ckeep = e->Ckeep; u.img = e->img;
    u.rm = rm;
    launch_unpatchify(u, s);
  } else {
    // split-bf16 product: [hi|lo|hi] x [Whi|Whi|Wlo] = hi*Whi + lo*Whi + hi*Wlo  (K = 3D), fp32-grade output
    TcGemmArgs a = tc_args(e->A3_h, 3 * D, e->final_w_h, 3 * D, M, e->Pout, 3 * D, TC_EPI_FINAL,
                           e->W("final_layer.linear.bias").d);
    a.out = out; a.T = T; a.p = e->p; a.Cout = e->Cout; a.Ckeep = e->Ckeep; a.img = e->img; a.rm = rm;
    launch_gemm_tc(a, s);
  }
}

// forward over `batch` independent rows (no CFG pairing)
static void forward_plain(vavae_engine* e, const float* x, const float* t, int t_stride, const int64_t* y, float* out,
                          int batch, cudaStream_t s) {
  for (int r0 = 0; r0 < batch; r0 += e->chunk_rows) {
    const int rows = std::min(e->chunk_rows, batch - r0);
    RowMap rm;
    rm.seg_rows = rows;
    rm.x_off[0] = rm.x_off[1] = r0;
    rm.y_off[0] = rm.y_off[1] = r0;
    rm.t_off[0] = rm.t_off[1] = r0;
    rm.out_off[0] = rm.out_off[1] = r0;
    rm.t_stride = t_stride;
    forward_chunk(e, rows, rm, x, t, y, out, s);
  }
}
// forward of [half, half] (lightningdit.py:425-427): model row b<n and b+n both read latent row b
static void forward_pairs(vavae_engine* e, const float* x, const float* t, int t_stride, const int64_t* y, float* out,
                          int n, cudaStream_t s) {
  const int per = e->chunk_rows / 2;
  for (int s0 = 0; s0 < n; s0 += per) {
    const int m = std::min(per, n - s0);
    RowMap rm;
    rm.seg_rows = m;
    rm.x_off[0] = s0; rm.x_off[1] = s0;
    rm.y_off[0] = s0; rm.y_off[1] = n + s0;
    rm.t_off[0] = s0; rm.t_off[1] = n + s0;
    rm.out_off[0] = s0; rm.out_off[1] = n + s0;
    rm.t_stride = t_stride;
    forward_chunk(e, 2 * m, rm, x, t, y, out, s);
  }
}

// forward of the n conditional rows only (model row b reads latent row b, label b, writes velocity row b): what is left
// of forward_pairs when the CFG-interval gate discards the unconditional half (lightningdit.py:436-439)
static void forward_cond_only(vavae_engine* e, const float* x, const float* t, int t_stride, const int64_t* y, float* out,
                              int n, cudaStream_t s) {
  forward_plain(e, x, t, t_stride, y, out, n, s);
}

static void ensure_state(vavae_engine* e, int rows) {
  if (rows <= e->state_rows) return;
  const size_t per = (size_t)e->C * e->img * e->img;
  for (float** pp : {&e->sx, &e->sxp, &e->sk1, &e->sv}) {
    if (*pp) e->release(*pp);
    *pp = e->alloc<float>((size_t)rows * per);
  }
  if (e->sy) e->release(e->sy);
  e->sy = e->alloc<int64_t>(rows);
  if (e->tvec) e->release(e->tvec);
  e->tvec = e->alloc<float>(rows);
  // buffers moved: graphs that baked the old addresses are stale
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.second.exec);
  e->graphs.clear();
  e->graph_warm.clear();
  e->state_rows = rows;
}

// Runs `launch(stream)` -- a fixed kernel sequence on engine-owned buffers -- through a CUDA graph: the first call
// with a given key runs eagerly (sets function attributes, warms caches), the second captures it on the engine's
// private stream (the caller's may be the legacy default stream, which cannot be captured) and every call from
// then on is a single cudaGraphLaunch on the caller's stream.
static void run_graphed(vavae_engine* e, const std::string& key, const std::function<void(cudaStream_t)>& launch,
                        cudaStream_t s) {
  auto it = e->graphs.find(key);
  if (it == e->graphs.end()) {
    if (!e->graph_warm.count(key)) {
      e->graph_warm.insert(key);
      launch(s);
      return;
    }
    if (e->graphs.size() >= 16) {
      for (auto& g : e->graphs) cudaGraphExecDestroy(g.second.exec);
      e->graphs.clear();
    }
    if (!e->gstream) VV_CUDA(cudaStreamCreateWithFlags(&e->gstream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    const uint64_t before = e->counter.n, rows_before = e->fwd_rows;
    VV_CUDA(cudaStreamBeginCapture(e->gstream, cudaStreamCaptureModeThreadLocal));
    try {
      launch(e->gstream);
    } catch (...) {
      cudaStreamEndCapture(e->gstream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    VV_CUDA(cudaStreamEndCapture(e->gstream, &graph));
    vavae_engine::GraphEntry ge;
    ge.nodes = (size_t)(e->counter.n - before);
    ge.rows = e->fwd_rows - rows_before;
    e->counter.n = before;  // captured launches did not execute; replays are counted below
    e->fwd_rows = rows_before;
    VV_CUDA(cudaGraphInstantiate(&ge.exec, graph, 0));
    VV_CUDA(cudaGraphDestroy(graph));
    it = e->graphs.emplace(key, ge).first;
  }
  VV_CUDA(cudaGraphLaunch(it->second.exec, s));
  e->counter.n += it->second.nodes;
  e->fwd_rows += it->second.rows;
}

static void sample_ode(vavae_engine* e, const float* z, const int64_t* y, const float* t_grid, int n_points, int method,
                       int use_cfg, float cfg_scale, int cfg_interval, float cfg_interval_start, int cfg_channels,
                       float* out_final, float* out_trace, int batch, int flags, cudaStream_t s) {
  // VAVAE_SAMPLE_SKIP_DISCARDED: an evaluation at t < cfg_interval_start keeps only the conditional output
  // (lightningdit.py:436-439 computes the unconditional forward and throws it away) - run the n conditional rows only
  const bool skip_discarded = (flags & VAVAE_SAMPLE_SKIP_DISCARDED) && use_cfg && cfg_interval;
  VV_CHECK(n_points >= 2, "sample_ode: need at least two grid points");
  VV_CHECK(method == VAVAE_ODE_EULER || method == VAVAE_ODE_HEUN, "sample_ode: method must be euler or heun (fixed grid)");
  VV_CHECK(!use_cfg || batch % 2 == 0, "sample_ode: CFG batch must be even (inference.py:205-211)");
  for (int j = 0; j + 1 < n_points; ++j) VV_CHECK(t_grid[j] < t_grid[j + 1], "ODE sampler has to be in forward time");
  ensure_state(e, batch);
  if (n_points > e->tgrid_cap) {
    if (e->tgrid_dev) e->release(e->tgrid_dev);
    e->tgrid_dev = e->alloc<float>(n_points);
    e->tgrid_cap = n_points;
  }
  const size_t per = (size_t)e->C * e->img * e->img;
  const size_t state_bytes = (size_t)batch * per * sizeof(float);
  VV_CUDA(cudaMemcpyAsync(e->tgrid_dev, t_grid, (size_t)n_points * sizeof(float), cudaMemcpyHostToDevice, s));
  // Small batches are launch-bound (an XL forward is ~230 launches): integrate on engine-owned buffers so one
  // evaluation is a fixed kernel sequence and replay it from a CUDA graph.
  const bool graphed = batch <= e->graph_max_rows && e->prof_kernel < 0;
  float* xs = out_final;  // large batches: integrate in place in the caller's output buffer
  const int64_t* yy = y;
  if (graphed) {
    xs = e->sx;
    if (y != e->sy) VV_CUDA(cudaMemcpyAsync(e->sy, y, (size_t)batch * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    yy = e->sy;
  }
  if (xs != z) VV_CUDA(cudaMemcpyAsync(xs, z, state_bytes, cudaMemcpyDeviceToDevice, s));
  int eval = 0;
  // returns true when only the conditional rows were evaluated (same fp32 compare as the device-side gate)
  auto eval_model = [&](const float* xin, int tidx) -> bool {
    const bool cond_only = skip_discarded && t_grid[tidx] < cfg_interval_start;
    if (graphed) {
      // the evaluation time goes through a fixed device scalar so the captured kernels never change
      VV_CUDA(cudaMemcpyAsync(e->st, e->tgrid_dev + tidx, sizeof(float), cudaMemcpyDeviceToDevice, s));
      const std::string key = std::string("S") + (use_cfg ? (cond_only ? "h" : "c") : "p") + (xin == e->sxp ? "1" : "0") + "_" +
                              std::to_string(batch);
      run_graphed(e, key, [&](cudaStream_t gs) {
        if (cond_only) forward_cond_only(e, xin, e->st, 0, yy, e->sv, batch / 2, gs);
        else if (use_cfg) forward_pairs(e, xin, e->st, 0, yy, e->sv, batch / 2, gs);
        else forward_plain(e, xin, e->st, 0, yy, e->sv, batch, gs);
      }, s);
      return cond_only;
    }
    const float* tp = e->tgrid_dev + tidx;  // every row uses the same time: t_stride = 0
    if (cond_only) forward_cond_only(e, xin, tp, 0, yy, e->sv, batch / 2, s);
    else if (use_cfg) forward_pairs(e, xin, tp, 0, yy, e->sv, batch / 2, s);
    else forward_plain(e, xin, tp, 0, yy, e->sv, batch, s);
    return cond_only;
  };
  OdeArgs o;
  o.v = e->sv; o.x = xs; o.xp = e->sxp; o.k1 = e->sk1; o.rows = batch; o.C = e->C; o.HW = e->img * e->img;
  o.use_cfg = use_cfg; o.scale = cfg_scale; o.interval = cfg_interval; o.start = cfg_interval_start;
  o.channels = cfg_channels; o.skip_uncond = 0;
  for (int j = 0; j + 1 < n_points; ++j) {
    const float dt = t_grid[j + 1] - t_grid[j];  // fp32, like torchdiffeq's t1 - t0 on fp32 tensors
    o.dt = dt;
    o.skip_uncond = eval_model(xs, j) ? 1 : 0;
    o.t = e->tgrid_dev + j;
    o.trace = out_trace ? out_trace + (size_t)(eval++) * batch * per : nullptr;
    o.stage = method == VAVAE_ODE_EULER ? 0 : 1;
    const double ode_bytes = (double)batch * per * 4.0 * 3.0;  // read v, read x, write x (Heun stages move a little more)
    {
      ProfScope ps(e, PK_ODE, ode_bytes, s);
      launch_ode_update(o, s);
    }
    if (method == VAVAE_ODE_HEUN) {
      o.skip_uncond = eval_model(e->sxp, j + 1) ? 1 : 0;
      o.t = e->tgrid_dev + j + 1;
      o.trace = out_trace ? out_trace + (size_t)(eval++) * batch * per : nullptr;
      o.stage = 2;
      ProfScope ps(e, PK_ODE, ode_bytes, s);
      launch_ode_update(o, s);
    }
  }
  if (xs != out_final) VV_CUDA(cudaMemcpyAsync(out_final, xs, state_bytes, cudaMemcpyDeviceToDevice, s));
}

// --------------------------------------------------------------------------------------------- adaptive dopri5
// Dormand-Prince 5(4) with the step controller the reference reaches through torchdiffeq.odeint(method="dopri5")
// (transport/integrators.py:118-125; the package is not vendored - semantics restated from its published algorithm:
// Hairer initial step, RMS error norm over the whole state, safety 0.9, ifactor 10, dfactor 0.2, order 5, FSAL,
// 4th-order dense output).  Step-size control runs on the host in double precision with one 8-byte D2H read per
// step (torchdiffeq synchronises per step as well); every tensor operation is a kernel of this library.
static const double kDpAlpha[6] = {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
static const double kDpBeta[6][6] = {
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
static const double kDpErr[7] = {35.0 / 384 - 1951.0 / 21600, 0, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                 -2187.0 / 6784 - -12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60};
static const double kDpMid[7] = {6025192743.0 / 30085553152.0 / 2, 0, 51252292925.0 / 65400821598.0 / 2,
                                 -2691868925.0 / 45128329728.0 / 2, 187940372067.0 / 1594534317056.0 / 2,
                                 -1776094331.0 / 19743644256.0 / 2, 11237099.0 / 235043384.0 / 2};

static void sample_dopri5(vavae_engine* e, const float* z, const int64_t* y, double t0, double t1, double atol,
                          double rtol, int use_cfg, float cfg_scale, int cfg_interval, float cfg_interval_start,
                          int cfg_channels, float* out_final, int batch, int64_t* stats, cudaStream_t s) {
  VV_CHECK(t0 < t1, "ODE sampler has to be in forward time");
  VV_CHECK(!use_cfg || batch % 2 == 0, "sample_dopri5: CFG batch must be even");
  ensure_state(e, batch);
  const size_t per = (size_t)e->C * e->img * e->img;
  const size_t n = (size_t)batch * per;
  if (batch > e->rk_rows) {
    for (int i = 0; i < 11; ++i) {
      if (e->rk[i]) e->release(e->rk[i]);
      e->rk[i] = e->alloc<float>(n);
    }
    if (!e->rk_partial) {
      e->rk_partial = e->alloc<double>(1024);
      e->rk_sum = e->alloc<double>(1);
    }
    e->rk_rows = batch;
  }
  float *y0 = e->rk[0], *y1 = e->rk[1], *ymid = e->rk[2], *ytmp = e->rk[3];
  float* k[7];
  for (int i = 0; i < 7; ++i) k[i] = e->rk[4 + i];
  VV_CUDA(cudaMemcpyAsync(y0, z, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  int64_t n_evals = 0, n_accept = 0, n_reject = 0;

  // f(t, yin) -> kout : the drift the reference integrates (forward_with_cfg / forward), evaluated at float32 t
  auto f = [&](float tval, const float* yin, float* kout) {
    VV_CUDA(cudaMemcpyAsync(e->st, &tval, sizeof(float), cudaMemcpyHostToDevice, s));
    if (use_cfg) {
      forward_pairs(e, yin, e->st, 0, y, e->sv, batch / 2, s);
      CfgArgs c;
      c.v = e->sv; c.out = kout; c.t = e->st; c.n = batch / 2; c.C = e->C; c.HW = e->img * e->img; c.scale = cfg_scale;
      c.interval = cfg_interval; c.start = cfg_interval_start; c.channels = cfg_channels;
      launch_cfg_combine(c, s);
    } else {
      forward_plain(e, yin, e->st, 0, y, kout, batch, s);
    }
    ++n_evals;
  };
  auto norm = [&](const float* const* kk, const double* coef, int nk, const float* sub, const float* ya, const float* yb) {
    RkNormArgs a;
    memset(&a, 0, sizeof(a));
    for (int j = 0; j < nk; ++j) { a.k[j] = kk[j]; a.coef[j] = (float)coef[j]; }
    a.nk = nk; a.sub = sub; a.y0 = ya; a.y1 = yb; a.atol = (float)atol; a.rtol = (float)rtol;
    a.partial = e->rk_partial; a.blocks = (int)std::min<size_t>(1024, (n + 255) / 256); a.n = n;
    launch_rk_sqnorm(a, e->rk_sum, s);
    double sum = 0.0;
    VV_CUDA(cudaMemcpyAsync(&sum, e->rk_sum, sizeof(double), cudaMemcpyDeviceToHost, s));
    VV_CUDA(cudaStreamSynchronize(s));
    return std::sqrt(sum / (double)n);   // RMS norm over the whole state
  };
  auto combine = [&](const float* base, const double* coef, int nk, double dt, float* out) {
    RkCombineArgs a;
    memset(&a, 0, sizeof(a));
    a.y0 = base; a.nk = nk; a.out = out; a.n4 = n / 4;
    for (int j = 0; j < nk; ++j) { a.k[j] = k[j]; a.coef[j] = (float)((float)coef[j] * (float)dt); }
    launch_rk_combine(a, s);
  };

  // ---- initial step (Hairer, Norsett & Wanner II.4, as used by torchdiffeq), order = 4
  f((float)t0, y0, k[0]);
  const double one = 1.0;
  const float* y0p[1] = {y0};
  const float* f0p[1] = {k[0]};
  const double d0 = norm(y0p, &one, 1, nullptr, y0, nullptr);
  const double d1 = norm(f0p, &one, 1, nullptr, y0, nullptr);
  double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
  {
    const double c1 = 1.0;
    combine(y0, &c1, 1, h0, ytmp);                       // y0 + h0 * f0
    f((float)(t0 + h0), ytmp, k[1]);
  }
  const float* f1p[1] = {k[1]};
  const double d2 = norm(f1p, &one, 1, k[0], y0, nullptr) / h0;
  double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? std::max(1e-6, h0 * 1e-3) : std::pow(0.01 / std::max(d1, d2), 1.0 / 5.0);
  double dt = std::min(100.0 * h0, h1);

  // ---- adaptive steps until the accepted interval covers t1, then dense output at t1
  double t = t0;
  const int max_steps = 100000;
  for (int it = 0;; ++it) {
    VV_CHECK(it < max_steps, "dopri5: max_num_steps exceeded");
    const float dtf = (float)dt, tf = (float)t;
    for (int i = 0; i < 6; ++i) {
      float* yi = (i == 5) ? y1 : ytmp;                  // the last stage input is the 5th-order solution (FSAL)
      combine(y0, kDpBeta[i], i + 1, dt, yi);
      const float ti = (kDpAlpha[i] == 1.0) ? (float)(t + dt) : tf + (float)kDpAlpha[i] * dtf;
      f(ti, yi, k[i + 1]);
    }
    double cerr[7];
    for (int j = 0; j < 7; ++j) cerr[j] = (double)((float)kDpErr[j] * dtf);
    const float* kc[7] = {k[0], k[1], k[2], k[3], k[4], k[5], k[6]};
    const double ratio = norm(kc, cerr, 7, nullptr, y0, y1);
    const bool accept = ratio <= 1.0;
    if (accept) {
      ++n_accept;
      if (t + dt >= t1) {
        combine(y0, kDpMid, 7, dt, ymid);
        const float x = (float)((t1 - t) / dt);
        launch_rk_interp(y0, y1, ymid, k[0], k[6], dtf, x, out_final, n, s);
        break;
      }
      std::swap(y0, y1);
      std::swap(k[0], k[6]);
      t += dt;
    } else {
      ++n_reject;
    }
    // torchdiffeq _optimal_step_size(dt, ratio, safety=0.9, ifactor=10, dfactor=0.2, order=5)
    double factor;
    if (ratio == 0.0) factor = 10.0;
    else factor = std::min(10.0, std::max(0.9 / std::pow(ratio, 1.0 / 5.0), ratio < 1.0 ? 1.0 : 0.2));
    dt *= factor;
  }
  if (stats) {
    stats[0] = n_accept;
    stats[1] = n_reject;
    stats[2] = n_evals;
  }
}

// --------------------------------------------------------------------------------------------- C ABI
#define VV_API_BEGIN(e)                 \
  try {                                 \
    vv::g_counter = (e) ? &(e)->counter : nullptr; \
    if (e) VV_CUDA(cudaSetDevice((e)->device));
#define VV_API_END()                    \
    return 0;                           \
  } catch (const std::exception& ex) {  \
    g_err = ex.what();                  \
    return 1;                           \
  }

extern "C" {

int vavae_abi_version(void) { return VAVAE_ABI_VERSION; }
const char* vavae_last_error(void) { return g_err.c_str(); }

int vavae_create(const vavae_config_t* cfg, int device, vavae_engine_t** out) {
  vavae_engine* e = nullptr;
  try {
    VV_CHECK(cfg && out, "null argument");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
      throw Error("no CUDA device available: this engine has no CPU fallback (" + std::string(cudaGetErrorString(ce)) + ")");
    VV_CHECK(device >= 0 && device < ndev, "device index out of range");
    VV_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw Error("this engine is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor));
    e = new vavae_engine();
    e->cfg = *cfg;
    e->device = device;
    derive(e);
    if (const char* gm = std::getenv("VAVAE_GRAPH_MAX_ROWS")) e->graph_max_rows = std::atoi(gm);
    if (const char* ao = std::getenv("VAVAE_ALT_ORDER")) e->alt_order = std::atoi(ao);
    *out = e;
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    delete e;
    return 1;
  }
}

int vavae_destroy(vavae_engine_t* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.second.exec);
  if (e->gstream) cudaStreamDestroy(e->gstream);
  for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
  for (void* p : e->owned)
    if (p) cudaFree(p);
  delete e;
  return 0;
}

int vavae_load_weight(vavae_engine_t* e, const char* name, const float* data, const int64_t* shape, int ndim,
                      int is_device) {
  VV_API_BEGIN(e)
  VV_CHECK(e && name && data && shape && ndim >= 1, "null argument");
  VV_CHECK(!e->finalized, "load_weight after finalize");
  Weight wt;
  wt.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    wt.shape.push_back(shape[i]);
    wt.numel *= (size_t)shape[i];
  }
  auto it = e->w.find(name);
  if (it != e->w.end()) {
    VV_CHECK(it->second.numel == wt.numel, std::string("reloaded weight changes size: ") + name);
    wt.d = it->second.d;
  } else {
    wt.d = e->alloc<float>(wt.numel);
  }
  VV_CUDA(cudaMemcpy(wt.d, data, wt.numel * sizeof(float), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  e->w[name] = wt;
  VV_API_END()
}

int vavae_finalize(vavae_engine_t* e, void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && !e->finalized, "finalize called twice");
  finalize(e, (cudaStream_t)stream);
  VV_API_END()
}

int vavae_forward(vavae_engine_t* e, const float* x, const float* t, const int64_t* y, float* out, int batch,
                  void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && e->finalized, "engine not finalized");
  VV_CHECK(batch > 0, "empty batch");
  cudaStream_t s = (cudaStream_t)stream;
  if (batch <= e->graph_max_rows && e->prof_kernel < 0) {
    // launch-bound regime: stage the inputs in engine-owned buffers and replay the forward from a CUDA graph
    ensure_state(e, batch);
    const size_t per = (size_t)e->C * e->img * e->img;
    VV_CUDA(cudaMemcpyAsync(e->sx, x, (size_t)batch * per * sizeof(float), cudaMemcpyDeviceToDevice, s));
    VV_CUDA(cudaMemcpyAsync(e->tvec, t, (size_t)batch * sizeof(float), cudaMemcpyDeviceToDevice, s));
    VV_CUDA(cudaMemcpyAsync(e->sy, y, (size_t)batch * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    run_graphed(e, "F_" + std::to_string(batch),
                [&](cudaStream_t gs) { forward_plain(e, e->sx, e->tvec, 1, e->sy, e->sv, batch, gs); }, s);
    VV_CUDA(cudaMemcpyAsync(out, e->sv, (size_t)batch * per * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    forward_plain(e, x, t, 1, y, out, batch, s);
  }
  VV_API_END()
}

int vavae_forward_with_cfg(vavae_engine_t* e, const float* x, const float* t, const int64_t* y, float cfg_scale,
                           int cfg_interval, float cfg_interval_start, int cfg_channels, float* out, int batch,
                           void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && e->finalized, "engine not finalized");
  VV_CHECK(batch > 0 && batch % 2 == 0, "forward_with_cfg needs an even batch (cond rows then null-label rows)");
  cudaStream_t s = (cudaStream_t)stream;
  ensure_state(e, batch);
  forward_pairs(e, x, t, 1, y, e->sv, batch / 2, s);
  CfgArgs c;
  c.v = e->sv; c.out = out; c.t = t; c.n = batch / 2; c.C = e->C; c.HW = e->img * e->img; c.scale = cfg_scale;
  c.interval = cfg_interval; c.start = cfg_interval_start; c.channels = cfg_channels;
  launch_cfg_combine(c, s);
  VV_API_END()
}

int vavae_sample_ode(vavae_engine_t* e, const float* z, const int64_t* y, const float* t_grid, int n_points, int method,
                     int use_cfg, float cfg_scale, int cfg_interval, float cfg_interval_start, int cfg_channels,
                     float* out_final, float* out_trace, int batch, int flags, void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && e->finalized, "engine not finalized");
  VV_CHECK(batch > 0 && z && y && t_grid && out_final, "null / empty argument");
  sample_ode(e, z, y, t_grid, n_points, method, use_cfg, cfg_scale, cfg_interval, cfg_interval_start, cfg_channels,
             out_final, out_trace, batch, flags, (cudaStream_t)stream);
  VV_API_END()
}

int vavae_sample_ode_host(vavae_engine_t* e, const float* z_host, const int64_t* y_host, const float* t_grid,
                          int n_points, int method, int use_cfg, float cfg_scale, int cfg_interval,
                          float cfg_interval_start, int cfg_channels, float* out_final_host, int batch, int flags,
                          void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && e->finalized, "engine not finalized");
  VV_CHECK(batch > 0 && z_host && y_host && t_grid && out_final_host, "null / empty argument");
  cudaStream_t s = (cudaStream_t)stream;
  ensure_state(e, batch);
  const size_t bytes = (size_t)batch * e->C * e->img * e->img * sizeof(float);
  VV_CUDA(cudaMemcpyAsync(e->sx, z_host, bytes, cudaMemcpyHostToDevice, s));
  VV_CUDA(cudaMemcpyAsync(e->sy, y_host, (size_t)batch * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  sample_ode(e, e->sx, e->sy, t_grid, n_points, method, use_cfg, cfg_scale, cfg_interval, cfg_interval_start,
             cfg_channels, e->sx, nullptr, batch, flags, s);
  VV_CUDA(cudaMemcpyAsync(out_final_host, e->sx, bytes, cudaMemcpyDeviceToHost, s));
  int bad = 0;
  VV_CUDA(cudaMemcpyAsync(&bad, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  VV_CUDA(cudaStreamSynchronize(s));
  if (bad) {
    VV_CUDA(cudaMemsetAsync(e->err_flag, 0, sizeof(int), s));
    throw Error("index out of range in self (label outside [0, label_rows) of y_embedder.embedding_table)");
  }
  VV_API_END()
}

int vavae_sample_dopri5(vavae_engine_t* e, const float* z, const int64_t* y, double t0, double t1, double atol, double rtol,
                        int use_cfg, float cfg_scale, int cfg_interval, float cfg_interval_start, int cfg_channels,
                        float* out_final, int batch, int64_t* stats, void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && e->finalized, "engine not finalized");
  VV_CHECK(batch > 0 && z && y && out_final, "null / empty argument");
  VV_CHECK(atol > 0 && rtol > 0, "dopri5 tolerances must be positive");
  sample_dopri5(e, z, y, t0, t1, atol, rtol, use_cfg, cfg_scale, cfg_interval, cfg_interval_start, cfg_channels, out_final,
                batch, stats, (cudaStream_t)stream);
  VV_API_END()
}

int vavae_device_status(vavae_engine_t* e, void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && e->finalized, "engine not finalized");
  cudaStream_t s = (cudaStream_t)stream;
  int bad = 0;
  VV_CUDA(cudaMemcpyAsync(&bad, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  VV_CUDA(cudaStreamSynchronize(s));
  if (bad) {
    VV_CUDA(cudaMemsetAsync(e->err_flag, 0, sizeof(int), s));
    throw Error("index out of range in self (label outside [0, label_rows) of y_embedder.embedding_table)");
  }
  VV_API_END()
}

uint64_t vavae_launch_count(const vavae_engine_t* e) { return e ? e->counter.n : 0; }
uint64_t vavae_forward_rows(const vavae_engine_t* e) { return e ? e->fwd_rows : 0; }
uint64_t vavae_device_bytes(const vavae_engine_t* e) { return e ? e->bytes : 0; }

int vavae_profile_begin(vavae_engine_t* e, const char* kernel) {
  VV_API_BEGIN(e)
  VV_CHECK(e, "null engine");
  e->prof_kernel = -1;
  e->prof_used = 0;
  e->prof_work = 0.0;
  if (kernel) {
    for (int i = 0; i < PK_COUNT; ++i)
      if (std::string(kernel) == kProfNames[i]) e->prof_kernel = i;
    VV_CHECK(e->prof_kernel >= 0, std::string("profile_begin: unknown kernel ") + kernel);
  }
  VV_API_END()
}

int vavae_profile_end(vavae_engine_t* e, double* total_ms, int64_t* launches, double* total_work) {
  VV_API_BEGIN(e)
  VV_CHECK(e && total_ms && launches && total_work, "null argument");
  double ms = 0.0;
  for (size_t i = 0; i < e->prof_used; ++i) {
    float t = 0.f;
    VV_CUDA(cudaEventSynchronize(e->prof_ev[2 * i + 1]));
    VV_CUDA(cudaEventElapsedTime(&t, e->prof_ev[2 * i], e->prof_ev[2 * i + 1]));
    ms += t;
  }
  *total_ms = ms;
  *launches = (int64_t)e->prof_used;
  *total_work = e->prof_work;
  e->prof_kernel = -1;
  VV_API_END()
}

// ---- kernel-level test entry points -----------------------------------------------------------
int vavae_test_gemm_bf16(const void* A, const void* Wt, const float* bias, float* Cc, int M, int N, int K, void* stream) {
  try {
    vv::g_counter = nullptr;
    TcGemmArgs a = tc_args((const bf16*)A, K, (const bf16*)Wt, K, M, N, K, TC_EPI_BIAS_F32, bias);
    a.out = Cc;
    a.ldo = N;
    launch_gemm_tc(a, (cudaStream_t)stream);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return 1;
  }
}

int vavae_test_attention_bf16(const void* q, const void* k, const void* vt, void* out, int B, int H, int T, int hd,
                              void* stream) {
  try {
    vv::g_counter = nullptr;
    TcAttnArgs a;
    a.q = (const bf16*)q; a.k = (const bf16*)k; a.vt = (const bf16*)vt; a.out = (bf16*)out;
    a.B = B; a.H = H; a.T = T; a.hd = hd; a.hdp = ((hd + 15) / 16) * 16; a.hdv = attention_hdv(hd);
    launch_attention_tc(a, (cudaStream_t)stream);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return 1;
  }
}

// One launch of the production tcgen05 GEMM with one of its fused epilogues on caller-provided operands (tests only).
int vavae_test_gemm_epilogue(vavae_epilogue_test_t* t, void* stream) {
  float* rope_cs = nullptr;
  try {
    vv::g_counter = nullptr;
    VV_CHECK(t && t->A && t->W && t->M > 0 && t->N > 0 && t->K > 0 && t->T > 0, "test_gemm_epilogue: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int epi = t->kind == VAVAE_EPI_QKV ? TC_EPI_QKV : (t->kind == VAVAE_EPI_SWIGLU ? TC_EPI_SWIGLU : TC_EPI_GATED_RESIDUAL);
    VV_CHECK(t->kind >= VAVAE_EPI_QKV && t->kind <= VAVAE_EPI_GATED_RESIDUAL, "test_gemm_epilogue: unknown kind");
    TcGemmArgs a = tc_args((const bf16*)t->A, t->K, (const bf16*)t->W, t->K, t->M, t->N, t->K, epi, t->bias);
    a.T = t->T;
    if (t->in_part) {
      VV_CHECK(t->kind != VAVAE_EPI_GATED_RESIDUAL && t->sp, "test_gemm_epilogue: fused-norm consumer needs sp and kind 0/1");
      a.in_part = t->in_part; a.in_nslots = t->in_nparts; a.in_fine = t->in_fine; a.in_part_ld = t->part_ld; a.in_dim = t->in_dim; a.sp = t->sp; a.sp_ld = t->sp_ld;
    }
    if (t->kind == VAVAE_EPI_QKV) {
      VV_CHECK(t->q && t->k && t->vt && t->H > 0 && t->hd > 0 && t->N == 3 * t->H * t->hd, "test_gemm_epilogue: qkv outputs");
      a.q = (bf16*)t->q; a.k = (bf16*)t->k; a.vt = (bf16*)t->vt; a.cos = t->cos; a.sin = t->sin; a.qn_w = t->qn_w; a.kn_w = t->kn_w;
      a.qk_rms = 1; a.H = t->H; a.hd = t->hd; a.hdp = ((t->hd + 15) / 16) * 16; a.hdv = attention_hdv(t->hd);
      if (t->cos && t->sin && t->hd % 2 == 0) {
        // same packing as vavae_finalize
        const size_t nel = (size_t)t->T * t->hd;
        std::vector<float> hc(nel), hs(nel), cs, tab;
        VV_CUDA(cudaStreamSynchronize(s));
        VV_CUDA(cudaMemcpy(hc.data(), t->cos, nel * 4, cudaMemcpyDeviceToHost));
        VV_CUDA(cudaMemcpy(hs.data(), t->sin, nel * 4, cudaMemcpyDeviceToHost));
        int gw = 0;
        pack_rope_tables(hc, hs, t->T, t->hd, cs, tab, gw);
        if (!cs.empty()) {
          VV_CUDA(cudaMalloc(&rope_cs, (nel + tab.size()) * 4));
          VV_CUDA(cudaMemcpy(rope_cs, cs.data(), nel * 4, cudaMemcpyHostToDevice));
          a.rope_cs = rope_cs;
          if (!tab.empty()) {
            VV_CUDA(cudaMemcpy(rope_cs + nel, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
            a.rope_tab = rope_cs + nel;
            a.rope_gw = gw;
          }
        }
      }
    } else if (t->kind == VAVAE_EPI_SWIGLU) {
      VV_CHECK(t->out_bf16 && t->ldo >= t->N / 2, "test_gemm_epilogue: swiglu output");
      a.out = t->out_bf16; a.ldo = t->ldo;
    } else {
      VV_CHECK(t->X && t->mod, "test_gemm_epilogue: gated residual needs X and mod");
      a.out = t->X; a.ldo = t->N; a.mod = t->mod; a.mod_ld = t->mod_ld; a.gate_off = t->gate_off;
      if (t->nx_w) {
        VV_CHECK(t->nx_A && t->nx_part, "test_gemm_epilogue: next-norm outputs");
        a.nx_w = t->nx_w; a.nx_scale_off = t->nx_scale_off; a.nx_A = (bf16*)t->nx_A; a.nx_part = t->nx_part;
        t->fine_out = !gemm_tc_part_coarse(a);
        t->nparts_out = gemm_tc_part_slots(t->N, !t->fine_out);
        a.nx_part_ld = t->part_ld;
        VV_CHECK(t->part_ld >= gemm_tc_part_ld(t->N, !t->fine_out), "test_gemm_epilogue: part_ld too small");
      }
    }
    launch_gemm_tc(a, s);
    if (rope_cs) {
      VV_CUDA(cudaStreamSynchronize(s));
      cudaFree(rope_cs);
    }
    return 0;
  } catch (const std::exception& ex) {
    if (rope_cs) cudaFree(rope_cs);
    g_err = ex.what();
    return 1;
  }
}

int vavae_sde_update(const vavae_sde_update_t* u, void* stream) {
  try {
    vv::g_counter = nullptr;
    VV_CHECK(u, "null argument");
    SdeArgs a;
    a.mode = u->mode; a.x = u->x; a.v = u->v; a.w = u->w; a.k1 = u->k1; a.xh = u->xh; a.out = u->out; a.out2 = u->out2;
    a.n = (size_t)u->n; a.t = u->t; a.dt = u->dt; a.diff = u->diffusion; a.var = u->var; a.s2d = u->sqrt_2diffusion;
    a.sqdt = u->sqrt_dt; a.c0 = u->c0; a.c1 = u->c1;
    launch_sde_update(a, (cudaStream_t)stream);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return 1;
  }
}

// Host-only: how launch_gemm_tc would tile and run one layer GEMM (tests of the tile-selection rules; no GPU needed).
int vavae_gemm_plan(int kind, int M, int N, int K, int T, int H, int hd, int fused, int32_t* out) {
  try {
    VV_CHECK(out && M > 0 && N > 0 && K > 0 && kind >= VAVAE_EPI_QKV && kind <= VAVAE_EPI_GATED_RESIDUAL, "gemm_plan: bad arguments");
    const int epi = kind == VAVAE_EPI_QKV ? TC_EPI_QKV : (kind == VAVAE_EPI_SWIGLU ? TC_EPI_SWIGLU : TC_EPI_GATED_RESIDUAL);
    bf16* fake = reinterpret_cast<bf16*>(uintptr_t(4096));   // aligned, never dereferenced
    TcGemmArgs a = tc_args(fake, K, fake, K, M, N, K, epi, nullptr);
    a.T = T; a.H = H; a.hd = hd;
    if (epi == TC_EPI_GATED_RESIDUAL) {
      a.out = fake; a.ldo = N; a.mod = reinterpret_cast<const float*>(fake);
      if (fused) { a.nx_w = reinterpret_cast<const float*>(fake); a.nx_A = fake; }
    }
    int o[6];
    gemm_tc_plan_query(a, o);
    for (int i = 0; i < 6; ++i) out[i] = o[i];
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return 1;
  }
}

int vavae_set_gemm_two_cta(int on) {
  vv::set_gemm_two_cta(on);
  return 0;
}

int vavae_time_kernel(vavae_engine_t* e, const char* kernel, int rows, int iters, float* ms, double* work, void* stream) {
  VV_API_BEGIN(e)
  VV_CHECK(e && e->finalized && kernel && ms && work, "bad argument");
  VV_CHECK(rows > 0 && rows <= e->chunk_rows, "time_kernel: rows must fit the engine's chunk");
  VV_CHECK(!e->fp32, "time_kernel measures the bf16 production kernels");
  cudaStream_t s = (cudaStream_t)stream;
  const int D = e->D, T = e->T, S = e->S, H = e->H, hd = e->hd, M = rows * T;
  const ModOff mo = mod_offsets(e, 0);
  const vavae_config_t& c = e->cfg;
  std::string k = kernel;
  std::function<void()> fn;
  const float* rcos = c.use_rope ? e->W("feat_rope.freqs_cos").d : nullptr;
  const float* rsin = c.use_rope ? e->W("feat_rope.freqs_sin").d : nullptr;
  // the kernels are timed as the sampler launches them: with the fused RMSNorm producer / consumer sides when active
  const bool fused = e->fuse;
  const float* sp0 = (fused && !c.wo_shift) ? e->mod + e->NM : nullptr;
  if (k == "gemm_qkv") {
    *work = 2.0 * M * 3.0 * D * D;
    fn = [&] {
      TcGemmArgs a = tc_args(e->A_h, D, e->lw[0].qkv, D, M, 3 * D, D, TC_EPI_QKV, e->W(blk(0, "attn.qkv.bias")).d);
      a.q = e->q_h; a.k = e->k_h; a.vt = e->vt_h; a.cos = rcos; a.sin = rsin; a.rope_cs = e->rope_cs; a.rope_tab = e->rope_tab; a.rope_gw = e->rope_gw; a.qk_rms = c.use_rmsnorm; a.H = H;
      a.hd = hd; a.hdp = e->hdp; a.hdv = e->hdv; a.T = T;
      if (c.use_qknorm) { a.qn_w = e->W(blk(0, "attn.q_norm.weight")).d; a.kn_w = e->W(blk(0, "attn.k_norm.weight")).d; }
      if (fused) {   // as launched for layers >= 1: 1/rms from the previous w3's partial sums, per-sample sp vector
        TcGemmArgs w3 = tc_args(e->Hh_h, e->Spad, e->lw[0].w2, e->Spad, M, D, S, TC_EPI_GATED_RESIDUAL, nullptr);
        w3.out = e->X; w3.ldo = D; w3.mod = e->mod; w3.T = T; w3.nx_w = e->W(blk(0, "norm1.weight")).d; w3.nx_A = e->A_h;
        const int pc = gemm_tc_part_coarse(w3);
        a.in_part = e->part2; a.in_fine = !pc; a.in_nslots = gemm_tc_part_slots(D, pc); a.in_part_ld = gemm_tc_part_ld(D, pc); a.in_dim = D;
        a.sp = sp0 ? sp0 : a.bias; a.sp_ld = sp0 ? e->NMx : 0;
      }
      launch_gemm_tc(a, s);
    };
  } else if (k == "gemm_proj") {
    *work = 2.0 * M * (double)D * D;
    fn = [&] {
      TcGemmArgs pr = tc_args(e->O_h, D, e->lw[0].proj, D, M, D, D, TC_EPI_GATED_RESIDUAL, e->W(blk(0, "attn.proj.bias")).d);
      pr.out = e->X; pr.ldo = D; pr.mod = e->mod; pr.mod_ld = e->NMx; pr.gate_off = mo.gate_msa; pr.T = T;
      if (fused) {
        pr.nx_w = e->W(blk(0, "norm2.weight")).d; pr.nx_scale_off = mo.scale_mlp; pr.nx_A = e->A_h; pr.nx_part = e->part1;
        pr.nx_part_ld = gemm_tc_part_ld(D, gemm_tc_part_coarse(pr));
      }
      launch_gemm_tc(pr, s);
    };
  } else if (k == "gemm_w12") {
    VV_CHECK(c.use_swiglu, "gemm_w12 needs use_swiglu");
    *work = 2.0 * M * 2.0 * S * D;
    fn = [&] {
      TcGemmArgs a = tc_args(e->A_h, D, e->lw[0].w1, D, M, ((2 * S + 3) / 4) * 4, D, TC_EPI_SWIGLU, e->lw[0].b1);
      a.out = e->Hh_h; a.ldo = e->Spad; a.T = T;
      if (fused) {
        TcGemmArgs pr = tc_args(e->O_h, D, e->lw[0].proj, D, M, D, D, TC_EPI_GATED_RESIDUAL, nullptr);
        pr.out = e->X; pr.ldo = D; pr.mod = e->mod; pr.T = T; pr.nx_w = e->W(blk(0, "norm2.weight")).d; pr.nx_A = e->A_h;
        const int pc = gemm_tc_part_coarse(pr);
        a.in_part = e->part1; a.in_fine = !pc; a.in_nslots = gemm_tc_part_slots(D, pc); a.in_part_ld = gemm_tc_part_ld(D, pc); a.in_dim = D;
        a.sp = sp0 ? sp0 + 3 * D : a.bias; a.sp_ld = sp0 ? e->NMx : 0;
      }
      launch_gemm_tc(a, s);
    };
  } else if (k == "gemm_w3") {
    *work = 2.0 * M * (double)S * D;
    fn = [&] {
      TcGemmArgs a = tc_args(e->Hh_h, e->Spad, e->lw[0].w2, e->Spad, M, D, S, TC_EPI_GATED_RESIDUAL,
                             e->W(blk(0, c.use_swiglu ? "mlp.w3.bias" : "mlp.fc2.bias")).d);
      a.out = e->X; a.ldo = D; a.mod = e->mod; a.mod_ld = e->NMx; a.gate_off = mo.gate_mlp; a.T = T;
      if (fused) {
        a.nx_w = e->W(blk(0, "norm1.weight")).d; a.nx_scale_off = mo.scale_msa; a.nx_A = e->A_h; a.nx_part = e->part2;
        a.nx_part_ld = gemm_tc_part_ld(D, gemm_tc_part_coarse(a));
      }
      launch_gemm_tc(a, s);
    };
  } else if (k == "attention") {
    *work = 4.0 * rows * (double)H * T * T * hd;
    fn = [&] {
      TcAttnArgs at;
      at.q = e->q_h; at.k = e->k_h; at.vt = e->vt_h; at.out = e->O_h; at.B = rows; at.H = H; at.T = T; at.hd = hd;
      at.hdp = e->hdp; at.hdv = e->hdv;
      launch_attention_tc(at, s);
    };
  } else if (k == "norm_modulate") {
    *work = (double)M * D * (4.0 + 2.0);  // bytes: read fp32 residual, write bf16 operand
    fn = [&] { norm_mod(e, c.use_rmsnorm ? e->W(blk(0, "norm1.weight")).d : nullptr, mo.shift_msa, mo.scale_msa, M, s); };
  } else if (k == "ode_update") {
    ensure_state(e, rows);
    const double el = (double)rows * e->C * e->img * e->img;
    *work = el * 4.0 * 3.0;  // bytes: read v, read x, write x
    VV_CUDA(cudaMemsetAsync(e->st, 0, 16, s));
    fn = [&] {
      OdeArgs o;
      o.v = e->sv; o.x = e->sx; o.xp = e->sxp; o.k1 = e->sk1; o.trace = nullptr; o.t = e->st; o.rows = rows; o.C = e->C;
      o.HW = e->img * e->img; o.use_cfg = rows % 2 == 0; o.scale = 1.0f; o.interval = 0; o.start = 0.f; o.channels = 3;
      o.dt = 0.f; o.stage = 0;
      launch_ode_update(o, s);
    };
  } else if (k == "patchify") {
    ensure_state(e, rows);
    const int Cpp = e->C * e->p * e->p;
    *work = (double)rows * e->C * e->img * e->img * 4.0 + (double)M * 3.0 * Cpp * 2.0;  // bytes: read latents, write split-bf16 patches
    RowMap rm;
    rm.seg_rows = rows; rm.x_off[0] = rm.x_off[1] = 0; rm.y_off[0] = rm.y_off[1] = 0; rm.t_off[0] = rm.t_off[1] = 0;
    rm.out_off[0] = rm.out_off[1] = 0; rm.t_stride = 0;
    fn = [&, rm] { launch_patchify_split3(e->sx, e->Aemb_h, rows, e->C, e->img, e->p, T, rm, s); };
  } else if (k == "cfg_combine") {
    ensure_state(e, rows);
    VV_CHECK(rows % 2 == 0, "cfg_combine needs an even number of rows");
    *work = (double)rows * e->C * e->img * e->img * 4.0 * 2.0;  // bytes: read v, write out
    VV_CUDA(cudaMemsetAsync(e->st, 0, 16, s));
    fn = [&] {
      CfgArgs cc;
      cc.v = e->sv; cc.out = e->sxp; cc.t = e->st; cc.n = rows / 2; cc.C = e->C; cc.HW = e->img * e->img; cc.scale = 4.0f;
      cc.interval = 0; cc.start = 0.f; cc.channels = 3;
      launch_cfg_combine(cc, s);
    };
  } else {
    throw Error("time_kernel: unknown kernel '" + k + "'");
  }
  cudaEvent_t e0, e1;
  VV_CUDA(cudaEventCreate(&e0));
  VV_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) fn();
  VV_CUDA(cudaEventRecord(e0, s));
  for (int i = 0; i < iters; ++i) fn();
  VV_CUDA(cudaEventRecord(e1, s));
  VV_CUDA(cudaEventSynchronize(e1));
  float total = 0.f;
  VV_CUDA(cudaEventElapsedTime(&total, e0, e1));
  *ms = total / (float)iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  VV_API_END()
}

}  // extern "C"
