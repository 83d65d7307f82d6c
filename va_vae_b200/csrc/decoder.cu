// VA-VAE (LDM AutoencoderKL) decode path on B200: post_quant_conv -> conv-ResNet decoder with 16x16 attention ->
// clamp(127.5 x + 128) uint8.   Reference: LightningDiT/tokenizer/autoencoder.py:10-160,277-399,490-493 and
// LightningDiT/tokenizer/vavae.py:72-82.
//
// Data layout: activations are NHWC.  The residual stream is fp32 [n*H*W, C]; every convolution operand is a bf16
// NHWC tensor with channels padded to a multiple of 64 (zero pad), produced by the GroupNorm(+swish) kernel, and the
// 3x3 / 1x1 convolutions run as implicit GEMMs on the tcgen05 kernel of gemm_tc.cu: the A tile is fetched by a 4-D
// TMA box over (C, W, H, N) whose (dy, dx)-shifted coordinates fall outside the image at the borders and are
// zero-filled by the hardware (= the padding), K = 9 * C with tap-major packed weights, bias and the ResNet skip
// connection are added in the epilogue.  The single-head 16x16 attention (head_dim = C = 512, above the flash
// kernel's limit) is three small GEMMs per image (QK^T, V^T, PV) around a row-softmax kernel.
// precision 1 (fp32) is a validation mode on SIMT kernels (direct convolution), used by the parity tests to separate
// indexing errors from bf16 rounding.
#include "../../include/vavae_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace vv;

extern thread_local std::string vv_last_error_buf;

namespace {

struct DWeight {
  float* d = nullptr;
  std::vector<int64_t> shape;
  size_t numel = 0;
};

struct ConvPack {
  bf16* w_h = nullptr;    // bf16 mode: [cout_pad][taps][cin_pad]
  float* w_f = nullptr;   // fp32 mode: same layout, fp32
  float* b = nullptr;     // [cout_pad]
  int cin = 0, cin_pad = 0, cout = 0, cout_pad = 0, taps = 0;
};

}  // namespace

struct vavae_decoder {
  vavae_decoder_config_t cfg;
  int device = 0;
  bool fp32 = false, finalized = false;
  std::map<std::string, DWeight> w;
  std::map<std::string, ConvPack> convs;
  std::vector<void*> owned;
  uint64_t bytes = 0;
  LaunchCounter counter;
  int max_batch = 0;
  // workspace (sized for max_batch images)
  float *X = nullptr, *Y = nullptr, *Hc = nullptr;   // fp32 NHWC activations
  void* A = nullptr;                                 // conv operand: bf16 (or fp32 in validation mode), channels padded
  void *QK = nullptr, *Vt = nullptr, *P = nullptr, *O = nullptr;   // attention temporaries
  float* Sc = nullptr;                               // attention scores [T, T] fp32 (one image at a time)
  float* gn_part = nullptr;                          // [n][slabs][32][2]
  float* gn_stat = nullptr;                          // [n][32][2] (mean, rstd)
  float* gn_wide = nullptr;                          // [n][hw/32][32][2] partial sums emitted by GEMM epilogues
  int max_slabs = 0;

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
  const DWeight& W(const std::string& n) const {
    auto it = w.find(n);
    if (it == w.end()) throw Error("decoder: missing weight '" + n + "'");
    return it->second;
  }
};

namespace {

constexpr int GN_GROUPS = 32;          // autoencoder.py:15-18
constexpr int GN_SLAB = 512;           // pixels per partial-sum block
inline int pad64(int c) { return (c + 63) / 64 * 64; }

// ------------------------------------------------------------------------------------------------ kernels
// eight consecutive operand channels in one store (16 bytes of bf16, or two float4 in the fp32 validation mode)
__device__ __forceinline__ void store8(bf16* o, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
  pk.z = *reinterpret_cast<uint32_t*>(&c); pk.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(o) = pk;
}
__device__ __forceinline__ void store8(float* o, const float* v) {
  *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// z (n, zc, r, r) NCHW fp32 -> post_quant_conv (1x1, autoencoder.py:469,491) -> NHWC operand with padded channels
template <typename OutT>
__global__ void k_post_quant(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ b,
                             OutT* __restrict__ out, int n, int zc, int hw, int cpad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n * hw * cpad;
  if (idx >= total) return;
  const int co = (int)(idx % cpad);
  const size_t pix = idx / cpad;
  const int img = (int)(pix / hw), p = (int)(pix % hw);
  float acc = 0.f;
  if (co < zc) {
    acc = b[co];
    for (int ci = 0; ci < zc; ++ci) acc += w[co * zc + ci] * z[((size_t)img * zc + ci) * hw + p];
  }
  out[idx] = (OutT)acc;
}

// GroupNorm statistics, stage 1: per (image, slab of GN_SLAB pixels, group) sum and sum of squares.  Each thread owns
// one float4 channel column for all its pixels; the block reduction runs in a fixed order (no atomics).
__global__ void __launch_bounds__(256) k_gn_partial(const float* __restrict__ x, float* __restrict__ part, int hw, int C,
                                                    int slabs) {
  extern __shared__ float sm[];   // [lanes][C] sums, then [lanes][C] squares
  const int img = blockIdx.y, slab = blockIdx.x;
  const int c4n = C >> 2, lanes = 256 / c4n;
  const int c4 = threadIdx.x % c4n, lane = threadIdx.x / c4n;
  const int p0 = slab * GN_SLAB, p1 = min(hw, p0 + GN_SLAB);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  const float4* base = reinterpret_cast<const float4*>(x + (size_t)img * hw * C);
  auto acc = [&](const float4 v) {
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
  };
  int p = p0 + lane;
  for (; p + 3 * lanes < p1; p += 4 * lanes) {   // four independent 16-byte loads in flight, fixed accumulation order
    const float4 v0 = base[(size_t)p * c4n + c4], v1 = base[(size_t)(p + lanes) * c4n + c4];
    const float4 v2 = base[(size_t)(p + 2 * lanes) * c4n + c4], v3 = base[(size_t)(p + 3 * lanes) * c4n + c4];
    acc(v0); acc(v1); acc(v2); acc(v3);
  }
  for (; p < p1; p += lanes) acc(base[(size_t)p * c4n + c4]);
  float* ss = sm;
  float* sq = sm + lanes * C;
  reinterpret_cast<float4*>(ss + lane * C)[c4] = s;
  reinterpret_cast<float4*>(sq + lane * C)[c4] = q;
  __syncthreads();
  const int cpg = C / GN_GROUPS;
  if (threadIdx.x < GN_GROUPS) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l)
      for (int c = 0; c < cpg; ++c) {
        a += ss[l * C + threadIdx.x * cpg + c];
        b += sq[l * C + threadIdx.x * cpg + c];
      }
    float* o = part + (((size_t)img * slabs + slab) * GN_GROUPS + threadIdx.x) * 2;
    o[0] = a;
    o[1] = b;
  }
}
// stage 2: (mean, rstd) per (image, group), fixed-order reduction in double; eps 1e-6, biased variance (torch GroupNorm)
__global__ void k_gn_final(const float* __restrict__ part, float* __restrict__ stat, int slabs, double count) {
  const int img = blockIdx.x, g = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < slabs; ++s) {
    const float* o = part + (((size_t)img * slabs + s) * GN_GROUPS + g) * 2;
    a += (double)o[0];
    b += (double)o[1];
  }
  const double mean = a / count;
  const double var = fmax(b / count - mean * mean, 0.0);
  stat[((size_t)img * GN_GROUPS + g) * 2] = (float)mean;
  stat[((size_t)img * GN_GROUPS + g) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-6));
}
// stage 2 for statistics emitted by the producing GEMM's epilogue (32-row slabs, many more of them): 8 threads per
// group walk interleaved slabs in double precision and are combined in a fixed order
__global__ void __launch_bounds__(256) k_gn_final_wide(const float* __restrict__ part, float* __restrict__ stat, int slabs,
                                                       double count) {
  __shared__ double sa[256], sb[256];
  const int img = blockIdx.x, g = threadIdx.x >> 3, sub = threadIdx.x & 7;
  double a = 0.0, b = 0.0;
  for (int s = sub; s < slabs; s += 8) {
    const float2 v = reinterpret_cast<const float2*>(part)[((size_t)img * slabs + s) * GN_GROUPS + g];
    a += (double)v.x;
    b += (double)v.y;
  }
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  if (sub == 0) {
    for (int i = 1; i < 8; ++i) { a += sa[threadIdx.x + i]; b += sb[threadIdx.x + i]; }
    const double mean = a / count;
    const double var = fmax(b / count - mean * mean, 0.0);
    stat[((size_t)img * GN_GROUPS + g) * 2] = (float)mean;
    stat[((size_t)img * GN_GROUPS + g) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-6));
  }
}
// stage 3: y = (x - mean) * rstd * gamma + beta, optional swish (autoencoder.py:10-12), written as the next
// convolution's operand (channels padded with zeros up to cpad)
template <typename OutT>
__global__ void k_gn_apply(const float* __restrict__ x, const float* __restrict__ stat, const float* __restrict__ gamma,
                           const float* __restrict__ beta, OutT* __restrict__ out, uint32_t npix, uint32_t hw, int C, int cpad,
                           int cpg_shift, int swish) {
  // one thread per 8 output channels; 32-bit index arithmetic (npix * cpad / 8 < 2^32 is checked by the launcher) and a
  // shift for the channel -> group map (channels per group is a power of two)
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8p = (uint32_t)cpad >> 3;
  const uint32_t pix = idx / c8p;
  if (pix >= npix) return;
  const int c = (int)(idx - pix * c8p) * 8;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    const uint32_t img = pix / hw;
    const float* xp = x + (size_t)pix * C + c;
    const float4 x0 = *reinterpret_cast<const float4*>(xp), x1 = *reinterpret_cast<const float4*>(xp + 4);
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const float gs[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float2* st = reinterpret_cast<const float2*>(stat) + (size_t)img * GN_GROUPS;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 mr = st[(c + j) >> cpg_shift];   // (mean, rstd)
      float y = (xs[j] - mr.x) * mr.y * gs[j] + bs[j];
      if (swish) y = y / (1.f + __expf(-y));
      v[j] = y;
    }
  }
  store8(out + (size_t)pix * cpad + c, v);
}
// fp32 [M, C] -> operand [M, cpad] (input of the 1x1 skip convolution)
template <typename OutT>
__global__ void k_cvt_pad(const float* __restrict__ x, OutT* __restrict__ out, size_t npix, int C, int cpad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * cpad) return;
  const int c = (int)(idx % cpad);
  const size_t pix = idx / cpad;
  out[idx] = (OutT)(c < C ? x[pix * C + c] : 0.f);
}
// nearest-neighbour x2 upsampling (autoencoder.py:31) straight into the operand of the following 3x3 convolution
template <typename OutT>
__global__ void k_upsample2x(const float* __restrict__ x, OutT* __restrict__ out, int n, int H, int W, int C, int cpad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 channels of an output pixel
  const int c8p = cpad >> 3;
  const size_t total = (size_t)n * 4 * H * W * c8p;
  if (idx >= total) return;
  const int c = (int)(idx % c8p) * 8;
  size_t pix = idx / c8p;
  const size_t opix = pix;
  const int ox = (int)(pix % (2 * W));
  pix /= 2 * W;
  const int oy = (int)(pix % (2 * H)), img = (int)(pix / (2 * H));
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    const float* src = x + (((size_t)img * H + (oy >> 1)) * W + (ox >> 1)) * C + c;
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  store8(out + opix * cpad + c, v);
}
// fp32 validation mode: direct convolution, one thread per output element
__global__ void k_conv_f32(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                           const float* __restrict__ res, float* __restrict__ out, int n, int H, int W, int cpad, int cout,
                           int ldo, int taps) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n * H * W * cout;
  if (idx >= total) return;
  const int co = (int)(idx % cout);
  const size_t pix = idx / cout;
  const int x0 = (int)(pix % W), y0 = (int)((pix / W) % H), img = (int)(pix / ((size_t)W * H));
  float acc = bias ? bias[co] : 0.f;
  for (int t = 0; t < taps; ++t) {
    const int yy = y0 + (taps == 9 ? t / 3 - 1 : 0), xx = x0 + (taps == 9 ? t % 3 - 1 : 0);
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float* ap = a + (((size_t)img * H + yy) * W + xx) * cpad;
    const float* wp = w + ((size_t)co * taps + t) * cpad;
    for (int c = 0; c < cpad; ++c) acc += ap[c] * wp[c];
  }
  if (res) acc += res[pix * ldo + co];
  out[pix * ldo + co] = acc;
}
// fp32 validation mode: C[M, N] = A[M, K] * B[N, K]^T (+ bias[n]) (+ res), arbitrary leading dimensions
__global__ void k_gemm_nt_f32(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                              const float* __restrict__ bias, const float* __restrict__ res, float* __restrict__ C, int ldc,
                              int M, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * N) return;
  const int nn = (int)(idx % N), m = (int)(idx / N);
  float acc = bias ? bias[nn] : 0.f;
  for (int k = 0; k < K; ++k) acc += A[(size_t)m * lda + k] * B[(size_t)nn * ldb + k];
  if (res) acc += res[(size_t)m * ldc + nn];
  C[(size_t)m * ldc + nn] = acc;
}
// softmax(scale * s) over the rows of a [T, T] score matrix (autoencoder.py:147-149), one warp per row
template <typename OutT>
__global__ void k_softmax_rows(const float* __restrict__ s, OutT* __restrict__ p, int T, float scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  const float* r = s + (size_t)row * T;
  float m = -INFINITY;
  for (int j = lane; j < T; j += 32) m = fmaxf(m, r[j] * scale);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) sum += __expf(r[j] * scale - m);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int j = lane; j < T; j += 32) p[(size_t)row * T + j] = (OutT)(__expf(r[j] * scale - m) * inv);
}
// [npix, ld] fp32 NHWC (first C columns) -> NCHW fp32 and / or uint8 NHWC clamp(127.5 x + 128, 0, 255) (vavae.py:80-81)
__global__ void k_to_image(const float* __restrict__ x, int ld, float* __restrict__ nchw, uint8_t* __restrict__ u8, int n,
                           int hw, int C) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * hw * C) return;
  const int c = (int)(idx % C);
  const size_t pix = idx / C;
  const float v = x[pix * ld + c];
  if (nchw) {
    const int img = (int)(pix / hw), p = (int)(pix % hw);
    nchw[((size_t)img * C + c) * hw + p] = v;
  }
  if (u8) u8[idx] = (uint8_t)fminf(fmaxf(127.5f * v + 128.0f, 0.f), 255.f);   // float -> uint8 truncates, as torch does
}

template <typename F>
void launch1d(size_t total, cudaStream_t s, F f) {
  if (total == 0) return;
  const int th = 256;
  f(dim3((unsigned)((total + th - 1) / th)), dim3(th), s);
  VV_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------ building blocks
struct Dec {
  vavae_decoder* d;
  cudaStream_t s;
  int n;   // images in this chunk
  const float* stats_of = nullptr;   // tensor whose GroupNorm partial sums the last GEMM epilogue left in d->gn_wide

  // can the GEMM writing [n*hw, C] fp32 emit the statistics of its output?  (tensor-core path, interior tiles only)
  bool can_emit(int hw, int C) const {
    const int cpg = C / GN_GROUPS;
    return !d->fp32 && d->gn_wide && cpg >= 4 && cpg <= 32 && (cpg & (cpg - 1)) == 0 && hw % 32 == 0 &&
           ((size_t)n * hw) % 128 == 0 && C % 128 == 0;
  }

  void group_norm(const float* x, const std::string& name, int hw, int C, bool swish) {
    if (stats_of == x) {
      // the producing GEMM already left per-slab sums: finish them (no pass over the activations)
      k_gn_final_wide<<<n, 256, 0, s>>>(d->gn_wide, d->gn_stat, hw / 32, (double)hw * (C / GN_GROUPS));
      VV_LAUNCH_CHECK();
    } else {
      const int slabs = cdiv(hw, GN_SLAB), c4n = C / 4, lanes = 256 / c4n;
      VV_CHECK(slabs <= d->max_slabs, "decoder: GroupNorm workspace too small");
      k_gn_partial<<<dim3(slabs, n), 256, (size_t)2 * lanes * C * sizeof(float), s>>>(x, d->gn_part, hw, C, slabs);
      VV_LAUNCH_CHECK();
      k_gn_final<<<n, GN_GROUPS, 0, s>>>(d->gn_part, d->gn_stat, slabs, (double)hw * (C / GN_GROUPS));
      VV_LAUNCH_CHECK();
    }
    stats_of = nullptr;
    const int cpad = pad64(C);
    const size_t npix = (size_t)n * hw;
    const float* g = d->W(name + ".weight").d;
    const float* b = d->W(name + ".bias").d;
    int cpg_shift = 0;
    while ((1 << cpg_shift) < C / GN_GROUPS) ++cpg_shift;
    VV_CHECK((1 << cpg_shift) == C / GN_GROUPS && npix * (size_t)(cpad / 8) < 0xFFFFFFFFull, "decoder: GroupNorm index range");
    launch1d(npix * (cpad / 8), s, [&](dim3 gr, dim3 bl, cudaStream_t st) {
      if (d->fp32) k_gn_apply<float><<<gr, bl, 0, st>>>(x, d->gn_stat, g, b, (float*)d->A, (uint32_t)npix, (uint32_t)hw, C, cpad, cpg_shift, swish);
      else k_gn_apply<bf16><<<gr, bl, 0, st>>>(x, d->gn_stat, g, b, (bf16*)d->A, (uint32_t)npix, (uint32_t)hw, C, cpad, cpg_shift, swish);
    });
  }
  void cvt_pad(const float* x, int hw, int C) {
    const int cpad = pad64(C);
    const size_t npix = (size_t)n * hw;
    launch1d(npix * cpad, s, [&](dim3 gr, dim3 bl, cudaStream_t st) {
      if (d->fp32) k_cvt_pad<float><<<gr, bl, 0, st>>>(x, (float*)d->A, npix, C, cpad);
      else k_cvt_pad<bf16><<<gr, bl, 0, st>>>(x, (bf16*)d->A, npix, C, cpad);
    });
  }
  // out[n*H*W, ldo] = conv(A operand) + bias (+ res); 3x3 pad 1 or 1x1
  void conv(const std::string& name, int H, int W, float* out, const float* res, bool emit_stats = false) {
    auto it = d->convs.find(name);
    VV_CHECK(it != d->convs.end(), "decoder: convolution '" + name + "' was not packed");
    const ConvPack& c = it->second;
    const int M = n * H * W;
    if (d->fp32) {
      launch1d((size_t)M * c.cout_pad, s, [&](dim3 gr, dim3 bl, cudaStream_t st) {
        k_conv_f32<<<gr, bl, 0, st>>>((const float*)d->A, c.w_f, c.b, res, out, n, H, W, c.cin_pad, c.cout_pad,
                                      c.cout_pad, c.taps);
      });
      return;
    }
    TcGemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = (const bf16*)d->A; g.lda = c.cin_pad;
    g.W = c.w_h; g.ldw = c.taps * c.cin_pad;
    g.M = M; g.N = c.cout_pad; g.K = c.taps * c.cin_pad;
    g.epi = TC_EPI_BIAS_F32; g.bias = c.b; g.out = out; g.ldo = c.cout_pad; g.res = res;
    g.T = 32;   // unused by this epilogue (rows are not grouped into samples)
    g.conv_taps = c.taps; g.conv_H = H; g.conv_W = W;
    if (c.taps == 1) g.conv_taps = 0;   // 1x1: rows of the NHWC operand are the GEMM rows already
    if (out == stats_of) stats_of = nullptr;
    if (emit_stats && can_emit(H * W, c.cout) && c.cout == c.cout_pad) {
      int sh = 0;
      while ((1 << sh) < c.cout / GN_GROUPS) ++sh;
      g.gn_part = d->gn_wide; g.gn_hw = H * W; g.gn_cpg_shift = sh;
      stats_of = out;
    }
    launch_gemm_tc(g, s);
  }
  // C[M, N] (+bias) = A[M, K] B[N, K]^T with operand-typed inputs; out_f32 selects the fp32 or the operand-typed output
  void gemm_nt(const void* A_, int lda, const void* B_, int ldb, const float* bias, void* C_, int ldc, int M, int N, int K,
               bool out_f32, const float* res = nullptr, int emit_hw = 0) {
    if (d->fp32) {
      launch1d((size_t)M * N, s, [&](dim3 gr, dim3 bl, cudaStream_t st) {
        k_gemm_nt_f32<<<gr, bl, 0, st>>>((const float*)A_, lda, (const float*)B_, ldb, bias, res, (float*)C_, ldc, M, N, K);
      });
      return;
    }
    TcGemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = (const bf16*)A_; g.lda = lda; g.W = (const bf16*)B_; g.ldw = ldb;
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = C_; g.ldo = ldc; g.T = 32;
    g.epi = out_f32 ? TC_EPI_BIAS_F32 : TC_EPI_BIAS_BF16;
    g.res = res;
    if (C_ == (const void*)stats_of) stats_of = nullptr;
    if (emit_hw > 0 && out_f32 && ldc == N && can_emit(emit_hw, N)) {
      int sh = 0;
      while ((1 << sh) < N / GN_GROUPS) ++sh;
      g.gn_part = d->gn_wide; g.gn_hw = emit_hw; g.gn_cpg_shift = sh;
      stats_of = (const float*)C_;
    }
    launch_gemm_tc(g, s);
  }

  // autoencoder.py:93-114 (temb None, dropout 0).  x: fp32 [n*hw, cin]; returns the buffer holding the result
  float* resnet(const std::string& name, float* x, float* spare, int H, int W, int cin, int cout) {
    const int hw = H * W;
    float* skip = x;
    if (cin != cout) {
      cvt_pad(x, hw, cin);
      conv(name + ".nin_shortcut", H, W, spare, nullptr);
      skip = spare;
    }
    group_norm(x, name + ".norm1", hw, cin, true);
    conv(name + ".conv1", H, W, d->Hc, nullptr, true);
    group_norm(d->Hc, name + ".norm2", hw, cout, true);
    conv(name + ".conv2", H, W, skip, skip, true);   // skip + conv2(...), in place
    return skip;
  }
  // autoencoder.py:135-160
  void attention(const std::string& name, float* x, int H, int W, int C) {
    const int T = H * W, cp = pad64(C), esz = d->fp32 ? 4 : 2;
    VV_CHECK(C % 64 == 0 && T % 64 == 0, "decoder: attention block needs channels % 64 and H*W % 64");
    group_norm(x, name + ".norm", T, C, false);
    const ConvPack& qk = d->convs.at(name + ".qk");
    const ConvPack& v = d->convs.at(name + ".v");
    const void* wqk = d->fp32 ? (const void*)qk.w_f : (const void*)qk.w_h;
    const void* wv = d->fp32 ? (const void*)v.w_f : (const void*)v.w_h;
    gemm_nt(d->A, cp, wqk, cp, qk.b, d->QK, 2 * C, n * T, 2 * C, C, d->fp32);   // [n*T, 2C] = (q | k)
    const float scale = 1.0f / sqrtf((float)C);
    for (int i = 0; i < n; ++i) {
      const char* qi = (const char*)d->QK + (size_t)i * T * 2 * C * esz;
      const char* ai = (const char*)d->A + (size_t)i * T * cp * esz;
      char* oi = (char*)d->O + (size_t)i * T * C * esz;
      gemm_nt(qi, 2 * C, qi + (size_t)C * esz, 2 * C, nullptr, d->Sc, T, T, T, C, true);            // S = q k^T
      if (d->fp32) k_softmax_rows<float><<<cdiv(T, 8), 256, 0, s>>>(d->Sc, (float*)d->P, T, scale);
      else k_softmax_rows<bf16><<<cdiv(T, 8), 256, 0, s>>>(d->Sc, (bf16*)d->P, T, scale);
      VV_LAUNCH_CHECK();
      gemm_nt(wv, cp, ai, cp, nullptr, d->Vt, T, C, T, C, d->fp32);                                 // V^T (no bias) [C, T]
      // O = P V + b_v  (rows of P sum to one, so the value bias can be added after the product)
      gemm_nt(d->P, T, d->Vt, T, v.b, oi, C, T, C, T, d->fp32);
    }
    // proj_out (1x1) + residual, in place on x
    const ConvPack& po = d->convs.at(name + ".proj_out");
    const void* wp = d->fp32 ? (const void*)po.w_f : (const void*)po.w_h;
    gemm_nt(d->O, C, wp, cp, po.b, x, C, n * T, C, C, true, x, T);
  }
};

// pack one convolution: PyTorch weight [cout, cin, kh, kw] -> [cout_pad][tap = kh*kw_index][cin_pad]
void pack_conv(vavae_decoder* d, const std::string& name, const std::vector<std::string>& srcs, int cin, int taps) {
  ConvPack c;
  c.cin = cin; c.cin_pad = pad64(cin); c.taps = taps;
  int cout = 0;
  for (auto& sname : srcs) cout += (int)d->W(sname + ".weight").shape[0];
  c.cout = cout; c.cout_pad = cout < 16 ? 16 : (cout + 3) / 4 * 4;   // narrow outputs (conv_out) use one 16-column tile
  std::vector<float> hw((size_t)c.cout_pad * taps * c.cin_pad, 0.f), hb(c.cout_pad, 0.f);
  int row0 = 0;
  for (auto& sname : srcs) {
    const DWeight& wt = d->W(sname + ".weight");
    const DWeight& bs = d->W(sname + ".bias");
    const int co_n = (int)wt.shape[0];
    VV_CHECK(wt.shape.size() == 4 && wt.shape[1] == cin && wt.shape[2] * wt.shape[3] == taps,
             "decoder: unexpected shape of " + sname + ".weight");
    VV_CHECK(bs.numel == (size_t)co_n, "decoder: unexpected shape of " + sname + ".bias");
    std::vector<float> src(wt.numel), bsrc(bs.numel);
    VV_CUDA(cudaMemcpy(src.data(), wt.d, wt.numel * sizeof(float), cudaMemcpyDeviceToHost));
    VV_CUDA(cudaMemcpy(bsrc.data(), bs.d, bs.numel * sizeof(float), cudaMemcpyDeviceToHost));
    for (int co = 0; co < co_n; ++co) {
      hb[row0 + co] = bsrc[co];
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < taps; ++t)
          hw[((size_t)(row0 + co) * taps + t) * c.cin_pad + ci] = src[((size_t)co * cin + ci) * taps + t];
    }
    row0 += co_n;
  }
  c.b = d->alloc<float>(hb.size());
  VV_CUDA(cudaMemcpy(c.b, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (d->fp32) {
    c.w_f = d->alloc<float>(hw.size());
    VV_CUDA(cudaMemcpy(c.w_f, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    std::vector<bf16> hh(hw.size());
    for (size_t i = 0; i < hw.size(); ++i) hh[i] = __float2bfloat16(hw[i]);
    c.w_h = d->alloc<bf16>(hh.size());
    VV_CUDA(cudaMemcpy(c.w_h, hh.data(), hh.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  }
  d->convs[name] = c;
}

inline bool has_attn(const vavae_decoder_config_t& c, int res) {
  for (int i = 0; i < c.num_attn_res; ++i)
    if (c.attn_res[i] == res) return true;
  return false;
}

void finalize(vavae_decoder* d) {
  const vavae_decoder_config_t& c = d->cfg;
  VV_CHECK(c.num_levels >= 1 && c.num_levels <= 8 && c.num_res_blocks >= 1, "decoder: bad level / block count");
  VV_CHECK(c.num_attn_res >= 0 && c.num_attn_res <= 4, "decoder: at most 4 attention resolutions");
  VV_CHECK(c.resolution % (1 << (c.num_levels - 1)) == 0, "decoder: resolution not divisible by the upsampling factor");
  auto chan_ok = [](int ch) { return ch % GN_GROUPS == 0 && 256 % (ch / 4) == 0; };
  int block_in = c.ch * c.ch_mult[c.num_levels - 1];
  int curr = c.resolution >> (c.num_levels - 1);
  size_t max_f = 0, max_a = 0;   // per-image elements of the fp32 activations / of the padded operand
  int max_attn_c = 0, max_attn_t = 0;
  auto track = [&](int res, int ch) {
    VV_CHECK(chan_ok(ch), "decoder: channel counts must be 32 * 2^k (GroupNorm(32) and the reduction layout)");
    max_f = std::max(max_f, (size_t)res * res * ch);
    max_a = std::max(max_a, (size_t)res * res * pad64(ch));
    d->max_slabs = std::max(d->max_slabs, cdiv(res * res, GN_SLAB));
  };
  auto pack_res = [&](const std::string& name, int cin, int cout) {
    pack_conv(d, name + ".conv1", {name + ".conv1"}, cin, 9);
    pack_conv(d, name + ".conv2", {name + ".conv2"}, cout, 9);
    if (cin != cout) pack_conv(d, name + ".nin_shortcut", {name + ".nin_shortcut"}, cin, 1);
    VV_CHECK(d->W(name + ".norm1.weight").numel == (size_t)cin && d->W(name + ".norm2.weight").numel == (size_t)cout,
             "decoder: norm size mismatch in " + name);
  };
  auto pack_attn = [&](const std::string& name, int ch, int res) {
    pack_conv(d, name + ".qk", {name + ".q", name + ".k"}, ch, 1);
    pack_conv(d, name + ".v", {name + ".v"}, ch, 1);
    pack_conv(d, name + ".proj_out", {name + ".proj_out"}, ch, 1);
    max_attn_c = std::max(max_attn_c, ch);
    max_attn_t = std::max(max_attn_t, res * res);
  };
  max_a = std::max(max_a, (size_t)curr * curr * pad64(c.z_channels));
  VV_CHECK(d->W("post_quant_conv.weight").numel == (size_t)c.z_channels * c.z_channels, "decoder: post_quant_conv shape");
  pack_conv(d, "decoder.conv_in", {"decoder.conv_in"}, c.z_channels, 9);
  track(curr, block_in);
  pack_res("decoder.mid.block_1", block_in, block_in);
  pack_attn("decoder.mid.attn_1", block_in, curr);
  pack_res("decoder.mid.block_2", block_in, block_in);
  for (int lvl = c.num_levels - 1; lvl >= 0; --lvl) {
    const int block_out = c.ch * c.ch_mult[lvl];
    for (int ib = 0; ib <= c.num_res_blocks; ++ib) {
      const std::string nm = "decoder.up." + std::to_string(lvl) + ".block." + std::to_string(ib);
      pack_res(nm, block_in, block_out);
      block_in = block_out;
      track(curr, block_in);
      if (has_attn(c, curr)) pack_attn("decoder.up." + std::to_string(lvl) + ".attn." + std::to_string(ib), block_in, curr);
    }
    if (lvl != 0) {
      pack_conv(d, "decoder.up." + std::to_string(lvl) + ".upsample.conv",
                {"decoder.up." + std::to_string(lvl) + ".upsample.conv"}, block_in, 9);
      curr *= 2;
      track(curr, block_in);
    }
  }
  pack_conv(d, "decoder.conv_out", {"decoder.conv_out"}, block_in, 9);
  max_f = std::max(max_f, (size_t)curr * curr * 16);
  d->max_batch = c.max_batch > 0 ? c.max_batch : 8;
  const size_t nb = d->max_batch;
  const size_t esz = d->fp32 ? 4 : 2;
  d->X = d->alloc<float>(nb * max_f);
  d->Y = d->alloc<float>(nb * max_f);
  d->Hc = d->alloc<float>(nb * max_f);
  d->A = d->alloc<char>(nb * max_a * esz);
  if (max_attn_c) {
    d->QK = d->alloc<char>(nb * max_attn_t * 2 * max_attn_c * esz);
    d->O = d->alloc<char>(nb * max_attn_t * max_attn_c * esz);
    d->Vt = d->alloc<char>((size_t)max_attn_t * max_attn_c * esz);
    d->P = d->alloc<char>((size_t)max_attn_t * max_attn_t * esz);
    d->Sc = d->alloc<float>((size_t)max_attn_t * max_attn_t);
  }
  d->gn_part = d->alloc<float>(nb * d->max_slabs * GN_GROUPS * 2);
  d->gn_stat = d->alloc<float>(nb * GN_GROUPS * 2);
  if (!d->fp32) d->gn_wide = d->alloc<float>(nb * (size_t)(c.resolution * c.resolution / 32 + 1) * GN_GROUPS * 2);
  d->finalized = true;
}

void decode_chunk(vavae_decoder* d, const float* z, int n, float* out_nchw, uint8_t* out_u8, cudaStream_t s) {
  const vavae_decoder_config_t& c = d->cfg;
  Dec k{d, s, n};
  int block_in = c.ch * c.ch_mult[c.num_levels - 1];
  int curr = c.resolution >> (c.num_levels - 1);
  const int zpad = pad64(c.z_channels);
  // z -> post_quant_conv -> operand of conv_in
  launch1d((size_t)n * curr * curr * zpad, s, [&](dim3 gr, dim3 bl, cudaStream_t st) {
    const float* w = d->W("post_quant_conv.weight").d;
    const float* b = d->W("post_quant_conv.bias").d;
    if (d->fp32) k_post_quant<float><<<gr, bl, 0, st>>>(z, w, b, (float*)d->A, n, c.z_channels, curr * curr, zpad);
    else k_post_quant<bf16><<<gr, bl, 0, st>>>(z, w, b, (bf16*)d->A, n, c.z_channels, curr * curr, zpad);
  });
  float* x = d->X;
  float* spare = d->Y;
  k.conv("decoder.conv_in", curr, curr, x, nullptr, true);
  auto run_res = [&](const std::string& name, int cin, int cout) {
    float* r = k.resnet(name, x, spare, curr, curr, cin, cout);
    if (r != x) std::swap(x, spare);
  };
  run_res("decoder.mid.block_1", block_in, block_in);
  k.attention("decoder.mid.attn_1", x, curr, curr, block_in);
  run_res("decoder.mid.block_2", block_in, block_in);
  for (int lvl = c.num_levels - 1; lvl >= 0; --lvl) {
    const int block_out = c.ch * c.ch_mult[lvl];
    for (int ib = 0; ib <= c.num_res_blocks; ++ib) {
      run_res("decoder.up." + std::to_string(lvl) + ".block." + std::to_string(ib), block_in, block_out);
      block_in = block_out;
      if (has_attn(c, curr))
        k.attention("decoder.up." + std::to_string(lvl) + ".attn." + std::to_string(ib), x, curr, curr, block_in);
    }
    if (lvl != 0) {
      const int cpad = pad64(block_in);
      launch1d((size_t)n * 4 * curr * curr * (cpad / 8), s, [&](dim3 gr, dim3 bl, cudaStream_t st) {
        if (d->fp32) k_upsample2x<float><<<gr, bl, 0, st>>>(x, (float*)d->A, n, curr, curr, block_in, cpad);
        else k_upsample2x<bf16><<<gr, bl, 0, st>>>(x, (bf16*)d->A, n, curr, curr, block_in, cpad);
      });
      curr *= 2;
      k.conv("decoder.up." + std::to_string(lvl) + ".upsample.conv", curr, curr, spare, nullptr, true);
      std::swap(x, spare);
    }
  }
  k.group_norm(x, "decoder.norm_out", curr * curr, block_in, true);
  const ConvPack& co = d->convs.at("decoder.conv_out");
  k.conv("decoder.conv_out", curr, curr, d->Hc, nullptr);
  launch1d((size_t)n * curr * curr * c.out_ch, s, [&](dim3 gr, dim3 bl, cudaStream_t st) {
    k_to_image<<<gr, bl, 0, st>>>(d->Hc, co.cout_pad, out_nchw, out_u8, n, curr * curr, c.out_ch);
  });
}

}  // namespace

// --------------------------------------------------------------------------------------------- C ABI
#define DEC_API_BEGIN(d)                                     \
  try {                                                      \
    vv::g_counter = (d) ? &(d)->counter : nullptr;           \
    if (d) VV_CUDA(cudaSetDevice((d)->device));
#define DEC_API_END()                    \
    return 0;                            \
  } catch (const std::exception& ex) {   \
    vv_last_error_buf = ex.what();       \
    return 1;                            \
  }

extern "C" {

int vavae_decoder_create(const vavae_decoder_config_t* cfg, int device, vavae_decoder_t** out) {
  vavae_decoder* d = nullptr;
  try {
    VV_CHECK(cfg && out, "null argument");
    VV_CHECK(cfg->abi_version == VAVAE_ABI_VERSION, "decoder config: ABI version mismatch");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
      throw Error("no CUDA device available: the decoder has no CPU fallback (" + std::string(cudaGetErrorString(ce)) + ")");
    VV_CHECK(device >= 0 && device < ndev, "device index out of range");
    VV_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw Error("this decoder is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor));
    d = new vavae_decoder();
    d->cfg = *cfg;
    d->device = device;
    d->fp32 = cfg->precision == 1;
    *out = d;
    return 0;
  } catch (const std::exception& ex) {
    vv_last_error_buf = ex.what();
    delete d;
    return 1;
  }
}

int vavae_decoder_destroy(vavae_decoder_t* d) {
  if (!d) return 0;
  cudaSetDevice(d->device);
  cudaDeviceSynchronize();
  for (void* p : d->owned)
    if (p) cudaFree(p);
  delete d;
  return 0;
}

int vavae_decoder_load_weight(vavae_decoder_t* d, const char* name, const float* data, const int64_t* shape, int ndim,
                              int is_device) {
  DEC_API_BEGIN(d)
  VV_CHECK(d && name && data && shape && ndim >= 1, "null argument");
  VV_CHECK(!d->finalized, "load_weight after finalize");
  DWeight wt;
  wt.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    wt.shape.push_back(shape[i]);
    wt.numel *= (size_t)shape[i];
  }
  VV_CHECK(d->w.find(name) == d->w.end(), std::string("decoder: weight loaded twice: ") + name);
  wt.d = d->alloc<float>(wt.numel);
  VV_CUDA(cudaMemcpy(wt.d, data, wt.numel * sizeof(float), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  d->w[name] = wt;
  DEC_API_END()
}

int vavae_decoder_finalize(vavae_decoder_t* d) {
  DEC_API_BEGIN(d)
  VV_CHECK(d && !d->finalized, "finalize called twice");
  finalize(d);
  DEC_API_END()
}

int vavae_decoder_decode(vavae_decoder_t* d, const float* z, int n, float* out_nchw, uint8_t* out_u8_nhwc, void* stream) {
  DEC_API_BEGIN(d)
  VV_CHECK(d && d->finalized, "decoder not finalized");
  VV_CHECK(z && n > 0 && (out_nchw || out_u8_nhwc), "decode: null argument / empty batch");
  const vavae_decoder_config_t& c = d->cfg;
  const int zr = c.resolution >> (c.num_levels - 1);
  const size_t zper = (size_t)c.z_channels * zr * zr, oper = (size_t)c.out_ch * c.resolution * c.resolution;
  for (int i0 = 0; i0 < n; i0 += d->max_batch) {
    const int nb = std::min(d->max_batch, n - i0);
    decode_chunk(d, z + (size_t)i0 * zper, nb, out_nchw ? out_nchw + (size_t)i0 * oper : nullptr,
                 out_u8_nhwc ? out_u8_nhwc + (size_t)i0 * oper : nullptr, (cudaStream_t)stream);
  }
  DEC_API_END()
}

uint64_t vavae_decoder_launch_count(vavae_decoder_t* d) { return d ? d->counter.n : 0; }
uint64_t vavae_decoder_device_bytes(vavae_decoder_t* d) { return d ? d->bytes : 0; }

}  // extern "C"
