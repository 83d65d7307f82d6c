// fp32 SIMT GEMM for the validation mode (true fp32 FMA; tcgen05 has no fp32 MMA and kind::tf32
// would not meet the 1e-5 parity bar) and for the tiny conditioning GEMMs of both modes.
//   C[M,N] = A[M,K] * W[N,K]^T + bias[N]   (nn.Linear layout)
#include "common.cuh"

namespace vv {

constexpr int GB_M = 64, GB_N = 64;

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_tanh_f(float x) {
  // nn.GELU(approximate="tanh"), lightningdit.py:214
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}

// GB_K: K extent of one shared-memory tile.  Every output element accumulates its products in ascending k with one
// fmaf chain whatever GB_K is, so the tile depth changes the schedule, never the result: 16 for large problems, 64 for
// the single-row-tile conditioning GEMMs (M <= 64), whose duration is K / GB_K exposed global-load latencies.
template <bool VEC, int GB_K>
__global__ void __launch_bounds__(256) k_gemm_f32(F32GemmArgs a) {
  __shared__ __align__(16) float As[GB_K][GB_M + 4];
  __shared__ __align__(16) float Ws[GB_K][GB_N + 4];
  constexpr int NV = GB_K / 16;    // float4 loads per thread, per operand, per tile
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const int tx = tid & 15, ty = tid >> 4;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lr = tid >> 2;         // 0..63 : tile row loaded by this thread
  const int lk = (tid & 3) * 4;    // 0,4,8,12 : k offset inside a 16-wide slice of the tile
  for (int k0 = 0; k0 < a.K; k0 += GB_K) {
    float av[NV][4], wv[NV][4];
    const int gm = m0 + lr, gn = n0 + lr;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int gk = k0 + v * 16 + lk;
#pragma unroll
      for (int i = 0; i < 4; ++i) av[v][i] = wv[v][i] = 0.f;
      if (VEC) {
        if (gm < a.M && gk < a.K) {
          float4 t = *reinterpret_cast<const float4*>(a.A + (size_t)gm * a.lda + gk);
          av[v][0] = t.x; av[v][1] = t.y; av[v][2] = t.z; av[v][3] = t.w;
        }
        if (gn < a.N && gk < a.K) {
          float4 t = *reinterpret_cast<const float4*>(a.W + (size_t)gn * a.ldw + gk);
          wv[v][0] = t.x; wv[v][1] = t.y; wv[v][2] = t.z; wv[v][3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (gm < a.M && gk + i < a.K) av[v][i] = a.A[(size_t)gm * a.lda + gk + i];
          if (gn < a.N && gk + i < a.K) wv[v][i] = a.W[(size_t)gn * a.ldw + gk + i];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        As[v * 16 + lk + i][lr] = av[v][i];
        Ws[v * 16 + lk + i][lr] = wv[v][i];
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
  const bool deep = grid.y == 1;   // one row tile: latency-bound on the k loop
  if (vec && deep) k_gemm_f32<true, 64><<<grid, 256, 0, s>>>(a);
  else if (vec) k_gemm_f32<true, 16><<<grid, 256, 0, s>>>(a);
  else if (deep) k_gemm_f32<false, 64><<<grid, 256, 0, s>>>(a);
  else k_gemm_f32<false, 16><<<grid, 256, 0, s>>>(a);
  VV_LAUNCH_CHECK();
}

}  // namespace vv
