// fp32 SIMT GEMM for the validation mode (true fp32 FMA; tcgen05 has no fp32 MMA and kind::tf32
// would not meet the 1e-5 parity bar) and for the tiny conditioning GEMMs of both modes.
//   C[M,N] = A[M,K] * W[N,K]^T + bias[N]   (nn.Linear layout)
#include "common.cuh"

namespace vv {

constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_tanh_f(float x) {
  // nn.GELU(approximate="tanh"), lightningdit.py:214
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_gemm_f32(F32GemmArgs a) {
  __shared__ __align__(16) float As[GB_K][GB_M + 4];
  __shared__ __align__(16) float Ws[GB_K][GB_N + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const int tx = tid & 15, ty = tid >> 4;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lr = tid >> 2;         // 0..63 : tile row loaded by this thread
  const int lk = (tid & 3) * 4;    // 0,4,8,12 : k offset
  for (int k0 = 0; k0 < a.K; k0 += GB_K) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
    int gm = m0 + lr, gn = n0 + lr, gk = k0 + lk;
    if (VEC) {
      if (gm < a.M && gk < a.K) {
        float4 t = *reinterpret_cast<const float4*>(a.A + (size_t)gm * a.lda + gk);
        av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
      }
      if (gn < a.N && gk < a.K) {
        float4 t = *reinterpret_cast<const float4*>(a.W + (size_t)gn * a.ldw + gk);
        wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (gm < a.M && gk + i < a.K) av[i] = a.A[(size_t)gm * a.lda + gk + i];
        if (gn < a.N && gk + i < a.K) wv[i] = a.W[(size_t)gn * a.ldw + gk + i];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = av[i];
      Ws[lk + i][lr] = wv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float ar[4] = {a4.x, a4.y, a4.z, a4.w}, wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
    int lab_row = 0;
    if (a.epi == F32_EPI_ADD_LABEL_SILU) {
      int seg = gm / a.rm.seg_rows, ii = gm - seg * a.rm.seg_rows;
      const int64_t lab = a.labels[a.rm.y_off[seg] + ii];
      if (lab < 0 || lab >= (int64_t)a.label_rows) {
        if (a.err_flag) *a.err_flag = 1;   // surfaced by vavae_device_status(); the read is clamped
      } else {
        lab_row = (int)lab;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float v = acc[i][j] + (a.bias ? a.bias[gn] : 0.f);
      if (a.epi == F32_EPI_SILU) v = silu_f(v);
      else if (a.epi == F32_EPI_GELU_TANH) v = gelu_tanh_f(v);
      else if (a.epi == F32_EPI_ADD_LABEL_SILU) {
        v += a.table[(size_t)lab_row * a.N + gn];   // c = t_emb + y_emb, lightningdit.py:405
        float sv = silu_f(v);                        // SiLU at the head of every adaLN_modulation
        if (a.C2) a.C2[(size_t)gm * a.ldc + gn] = sv;
        if (a.C2h) a.C2h[(size_t)gm * a.ldc + gn] = __float2bfloat16_rn(sv);
      }
      a.C[(size_t)gm * a.ldc + gn] = v;
    }
  }
}

void launch_gemm_f32(const F32GemmArgs& a, cudaStream_t s) {
  dim3 grid(cdiv(a.N, GB_N), cdiv(a.M, GB_M));
  bool vec = (a.K % 4 == 0) && (a.lda % 4 == 0) && (a.ldw % 4 == 0) &&
             ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.W) & 15) == 0);
  if (vec)
    k_gemm_f32<true><<<grid, 256, 0, s>>>(a);
  else
    k_gemm_f32<false><<<grid, 256, 0, s>>>(a);
  VV_LAUNCH_CHECK();
}

}  // namespace vv
