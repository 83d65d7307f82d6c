// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//   C[M,N] = A[M,K] (bf16, K-major) * W[N,K]^T (bf16, K-major), fp32 accumulation in TMEM.
//   warp 0      : TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16 per instruction)
//   warps 2..9  : epilogue - tcgen05.ld from TMEM (lane quarter = warp%4, two warps per quarter taking alternate
//                 column chunks), fused epilogue, smem-staged coalesced stores, residual prefetched a chunk ahead
// Accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1; shared-memory operand tiles run through an N-stage full/empty mbarrier ring.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace vv {

// ------------------------------------------------------------------------------------------------ host tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw Error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// Descriptors are cached per (pointer, shape, box, swizzle): a sampler step re-launches the same ~10 GEMM shapes on the
// same workspace pointers tens of thousands of times, and cuTensorMapEncodeTiled is a driver call.
struct TmapKey {
  uint64_t v[12];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 12; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
template <typename F>
static CUtensorMap cached_tmap(const TmapKey& key, F&& encode) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  const CUtensorMap m = encode();
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, m);
  return m;
}

CUtensorMap make_tmap_2d_bf16(const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes,
                              uint32_t box_inner, uint32_t box_outer) {
  const TmapKey key = {{2, (uint64_t)(uintptr_t)base, inner, outer, row_pitch_bytes, box_inner, box_outer, 128, 0, 0, 0, 0}};
  return cached_tmap(key, [&] {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  VV_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && row_pitch_bytes % 16 == 0, "tensor map alignment");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(2d) failed with code " + std::to_string((int)r));
  return m;
  });
}
CUtensorMap make_tmap_3d_bf16(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes,
                              uint64_t pitch2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  const TmapKey key = {{3, (uint64_t)(uintptr_t)base, d0, d1, d2, pitch1_bytes, pitch2_bytes, b0, b1, b2, 128, 0}};
  return cached_tmap(key, [&] {
  CUtensorMap m;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  VV_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && pitch1_bytes % 16 == 0 && pitch2_bytes % 16 == 0,
           "tensor map alignment");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(3d) failed with code " + std::to_string((int)r));
  return m;
  });
}

CUtensorMap make_tmap_nhwc_bf16(const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint32_t bw,
                                uint32_t bh, uint32_t bn) {
  const TmapKey key = {{4, (uint64_t)(uintptr_t)base, C, W, H, N, bw, bh, bn, 0, 0, 0}};
  return cached_tmap(key, [&] {
  CUtensorMap m;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, C * W * 2, C * W * H * 2};
  cuuint32_t box[4] = {64, bw, bh, bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  VV_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 64 == 0, "NHWC tensor map: base alignment / channels % 64");
  VV_CHECK(bw <= 256 && bh <= 256 && bn <= 256 && bw * bh * bn == 128, "NHWC tensor map: box must cover 128 pixels");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);   // out-of-bounds elements read as zero = the conv padding
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(nhwc) failed with code " + std::to_string((int)r));
  return m;
  });
}

CUtensorMap make_tmap_3d_bf16_sw(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes,
                                 uint64_t pitch2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes) {
  const TmapKey key = {{5, (uint64_t)(uintptr_t)base, d0, d1, d2, pitch1_bytes, pitch2_bytes, b0, b1, b2,
                        (uint64_t)swizzle_bytes, 0}};
  return cached_tmap(key, [&] {
  CUtensorMap m;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  VV_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && pitch1_bytes % 16 == 0 && pitch2_bytes % 16 == 0,
           "tensor map alignment");
  VV_CHECK((int)b0 * 2 <= swizzle_bytes, "tensor map: inner box wider than the swizzle span");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(3d, swizzle) failed with code " + std::to_string((int)r));
  return m;
  });
}

// General 4-D tiled map (bf16 or fp32 elements; swizzle_bytes 0 = none / 128).  Used for TMA STORES of output tiles
// (attention output rows of one head, residual-stream tiles); out-of-bounds parts of a box are not written.
CUtensorMap make_tmap_4d(const void* base, int elem_bytes, const uint64_t dims_[4], const uint64_t strides_bytes[3],
                         const uint32_t box_[4], int swizzle_bytes) {
  const TmapKey key = {{6, (uint64_t)(uintptr_t)base, dims_[0] | (dims_[1] << 32), dims_[2] | (dims_[3] << 32),
                        strides_bytes[0], strides_bytes[1], strides_bytes[2],
                        (uint64_t)box_[0] | ((uint64_t)box_[1] << 16) | ((uint64_t)box_[2] << 32) | ((uint64_t)box_[3] << 48),
                        (uint64_t)swizzle_bytes, (uint64_t)elem_bytes, 0, 0}};
  return cached_tmap(key, [&] {
  CUtensorMap m;
  cuuint64_t dims[4] = {dims_[0], dims_[1], dims_[2], dims_[3]};
  cuuint64_t strides[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t box[4] = {box_[0], box_[1], box_[2], box_[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  VV_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && strides[0] % 16 == 0 && strides[1] % 16 == 0 &&
               strides[2] % 16 == 0 && ((uint64_t)box[0] * elem_bytes) % 16 == 0,
           "tensor map (4d) alignment");
  VV_CHECK(swizzle_bytes == 0 || (int)box[0] * elem_bytes <= swizzle_bytes, "tensor map (4d): inner box wider than the swizzle span");
  CUresult r = get_encode()(&m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                            const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(4d) failed with code " + std::to_string((int)r));
  return m;
  });
}

// ------------------------------------------------------------------------------------------------ kernel
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KiB
constexpr int TC_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int TC_MAX_STAGES = 8;

__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + k1 * x * x * x)));
}
// x.x^2 + x.y^2 + x.z^2 + x.w^2 with a pinned contraction: every path that emits RMSNorm partial sums (generic, fast
// and TMA epilogues) must round identically, or results would depend on which path a row's tile took
__device__ __forceinline__ float sumsq4(const float4 x) {
  return fmaf(x.w, x.w, fmaf(x.z, x.z, fmaf(x.y, x.y, x.x * x.x)));
}
// ---- RMSNorm partial sums (see TcGemmArgs::nx_part).  s(c, j) = sumsq4 of columns 32c + 4j .. 32c + 4j + 3 of a row.
//   coarse slot g (one float):   fold8_j( ((s(4g, j) + s(4g+1, j)) + s(4g+2, j)) + s(4g+3, j) )
//   fine slots of chunk c (eight floats):  s(c, 0..7)
// fold8 is the order of an 8-lane xor-butterfly (4, 2, 1) as seen by lane 0.
__device__ __forceinline__ float fold8(const float* s) {
  return ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
}
// consumer: total of a row's slots in ascending group order (nslots = coarse slots, or fine chunks).  The fine form is
// only met on small or odd-shaped problems and is kept out of line so that it does not weigh on the register allocation
// of the consumer epilogues (the QKV epilogue sits at the 168-register cap).
__device__ __noinline__ float part_total_fine(const float* __restrict__ slots, int nslots) {
  float total = 0.f;
  for (int c0 = 0; c0 < nslots; c0 += 4) {
    float t[8];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      if (c0 + cc < nslots) {
        const float4 lo = *reinterpret_cast<const float4*>(slots + (c0 + cc) * 8);
        const float4 hi = *reinterpret_cast<const float4*>(slots + (c0 + cc) * 8 + 4);
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = cc == 0 ? v[j] : t[j] + v[j];
      }
    }
    total += fold8(t);
  }
  return total;
}
__device__ __forceinline__ float part_total(const float* __restrict__ slots, int nslots, int fine) {
  if (fine) return part_total_fine(slots, nslots);
  // coarse: at most 16 slots (2048 columns); independent 16-byte loads, then the adds in ascending order
  const float4* pp = reinterpret_cast<const float4*>(slots);
  const int n4 = (nslots + 3) >> 2;
  float4 pv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) pv[i] = i < n4 ? pp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = nslots - 4 * i;   // valid slots in this quad (stale values may follow the last one)
    if (r > 0) total += pv[i].x;
    if (r > 1) total += pv[i].y;
    if (r > 2) total += pv[i].z;
    if (r > 3) total += pv[i].w;
  }
  return total;
}
// gated-residual chunk ownership: of the two epilogue warps of a TMEM lane quarter, warp `half` owns the 128-column
// groups half, half + 2, ... of a tile and walks each group's 32-column chunks in ascending order (so that a coarse
// slot is accumulated by one lane in a fixed order).  n_chunks = "none left".
__device__ __forceinline__ int gr_first(int half, int n_chunks) { return half * 4 < n_chunks ? half * 4 : n_chunks; }
__device__ __forceinline__ int gr_next(int c, int n_chunks) {
  ++c;
  if ((c & 3) == 0) c += 4;
  return c < n_chunks ? c : n_chunks;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Per-row statistics over one head's hd accumulator columns (re-read from TMEM, 8 columns at a time).
// Called with the warp converged (head boundaries depend on the column only).
__device__ __noinline__ void qk_head_stats(uint32_t tbase, int hc0, int n0, int hd, const float* __restrict__ bias,
                                           float in_scale, int rms, float& mean, float& rstd) {
  float s1 = 0.f, s2 = 0.f;
  for (int dd = 0; dd < hd; dd += 8) {
    uint32_t t8[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(t8[0]), "=r"(t8[1]), "=r"(t8[2]), "=r"(t8[3]), "=r"(t8[4]), "=r"(t8[5]), "=r"(t8[6]),
                   "=r"(t8[7])
                 : "r"(tbase + hc0 + dd));
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float x = __uint_as_float(t8[e]) * in_scale + bias[n0 + hc0 + dd + e];
      s1 += x;
      s2 += x * x;
    }
  }
  if (rms) {
    mean = 0.f;
    rstd = rsqrtf(s2 / (float)hd + 1e-6f);
  } else {
    mean = s1 / (float)hd;
    rstd = rsqrtf(fmaxf(s2 / (float)hd - mean * mean, 0.f) + 1e-5f);
  }
}

// ------------------------------------------------------------------------------------------------ fast epilogue
// Guard-free epilogue for the production cases: interior tile (all 128 rows < M, all BN columns < N), 32-column
// chunks, the warp's 32 rows are consecutive tokens of one sample (T % 32 == 0), QKV without qk-norm.  Same math
// as the generic path below; every bound check, per-pass division and generic-address store is gone (the generic
// path measured ~900 instructions per 32x32 chunk, mostly integer address/guard code).
__device__ __forceinline__ void sts128(uint32_t addr, float a0, float a1, float a2, float a3) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <int EPI>
__device__ __forceinline__ bool epi_fast_supported(const TcGemmArgs& a) {
  if (EPI == TC_EPI_GATED_RESIDUAL || EPI == TC_EPI_SWIGLU || EPI == TC_EPI_EMBED || EPI == TC_EPI_BIAS_F32) return true;
  if (EPI == TC_EPI_QKV) return a.qn_w == nullptr && a.kn_w == nullptr;
  return false;
}

template <int EPI>
__device__ __forceinline__ void epi_tile_fast(const TcGemmArgs& a, const uint32_t st_s, const uint32_t tbase,
                                              uint64_t* tfull, const uint32_t aphase, const int m0, const int n0,
                                              const int n_blk, const int BN, const int q, const int half,
                                              const int lane, const uint32_t rope_s) {
  constexpr int CW = 32, LPR = 8, RPP = 4, SST = (CW + 4) * 4;  // staging row stride in bytes
  const int prow = lane >> 3, pcc = (lane & 7) * 4;
  const int wrow0 = m0 + q * 32;
  const int Tdiv = a.T > 0 ? a.T : 1;   // plain bias epilogues do not group rows into samples (T may be 0)
  const int wb = wrow0 / Tdiv, wt0 = wrow0 - wb * Tdiv;
  const int n_chunks = BN / CW;
  const uint32_t st_w = st_s + (uint32_t)lane * SST;            // thread == row staging write base
  const uint32_t st_r = st_s + (uint32_t)prow * SST + pcc * 4;  // row-pass read base (+ ps * RPP * SST)
  const bool fusedc = (EPI == TC_EPI_QKV || EPI == TC_EPI_SWIGLU) && a.in_part != nullptr;

  // ---- tile prologue (before the accumulator is ready)
  float in_rstd = 1.f;
  const float* bvec = a.bias;
  if (fusedc) {
    const float ssum = part_total(a.in_part + (size_t)(wrow0 + lane) * a.in_part_ld, a.in_nslots, a.in_fine);
    in_rstd = rsqrtf(ssum / (float)a.in_dim + 1e-6f);
    bvec = a.sp + (size_t)wb * a.sp_ld;
  }
  const float* modrow = a.mod + (size_t)wb * a.mod_ld;
  float* xptr = nullptr;   // this lane's first residual element of chunk 0, pass 0
  bf16* nxptr = nullptr;
  if (EPI == TC_EPI_GATED_RESIDUAL) {
    xptr = reinterpret_cast<float*>(a.out) + (size_t)(wrow0 + prow) * a.ldo + n0 + pcc;
    if (a.nx_w) nxptr = a.nx_A + (size_t)(wrow0 + prow) * a.ldo + n0 + pcc;
  }
  const size_t xpass = (size_t)RPP * a.ldo;  // elements between passes
  float4 xpre[LPR];
  // prefetched one chunk ahead: bias, gate, and the RAW weight / scale vectors of the norm that follows (they are
  // combined at the point of use - combining them here would stall the warp on its own loads)
  float4 bb_n = make_float4(0.f, 0.f, 0.f, 0.f), g_n = bb_n, nw_n = bb_n, ns_n = bb_n;
  auto prefetch = [&](int c) {
    if (c >= n_chunks) return;
    const int colv = n0 + c * CW + pcc;
    if (bvec) bb_n = *reinterpret_cast<const float4*>(bvec + colv);
    if (EPI == TC_EPI_GATED_RESIDUAL) {
      if (!(a.experiment & 1)) {
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) {
        if (a.experiment & 8) {   // experiment: residual loads that do not allocate L1 lines
          const float* p = xptr + ps * xpass + c * CW;
          asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(xpre[ps].x), "=f"(xpre[ps].y), "=f"(xpre[ps].z), "=f"(xpre[ps].w) : "l"(p) : "memory");
        } else {
          xpre[ps] = *reinterpret_cast<const float4*>(xptr + ps * xpass + c * CW);
        }
      }
      }
      g_n = *reinterpret_cast<const float4*>(modrow + a.gate_off + colv);
      if (a.nx_w) {
        nw_n = *reinterpret_cast<const float4*>(a.nx_w + colv);
        ns_n = *reinterpret_cast<const float4*>(modrow + a.nx_scale_off + colv);
      }
    }
  };
  prefetch(EPI == TC_EPI_GATED_RESIDUAL ? gr_first(half, n_chunks) : half);
  // consumer epilogues: the per-sample shift.W^T+bias vector lives in the big `mod` matrix (DRAM latency): request
  // the vectors of all (<= 4) chunks this warp will process before waiting for the accumulator
  constexpr bool kBiasAhead = EPI == TC_EPI_SWIGLU || EPI == TC_EPI_EMBED || EPI == TC_EPI_BIAS_F32;  // (QKV: register budget, stays one-ahead)
  float4 bball[4];
  if (kBiasAhead) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const int c = half + 2 * ci;
      bball[ci] = (bvec && c < n_chunks) ? *reinterpret_cast<const float4*>(bvec + n0 + c * CW + pcc)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float ssacc[LPR];   // gated residual, coarse slots: this lane's sums of squares over the current 128-column group
#pragma unroll
  for (int ps = 0; ps < LPR; ++ps) ssacc[ps] = 0.f;
  // QKV: head index / in-head offset by multiply-high (exact for col * hd < 2^32) instead of integer division
  const int qD = EPI == TC_EPI_QKV ? a.H * a.hd : 1;
  const uint32_t hd_magic = EPI == TC_EPI_QKV ? 0xFFFFFFFFu / (uint32_t)a.hd + 1u : 0u;
  // compact RoPE tables in shared memory: grid row / column of this lane's token in pass 0 (tokens advance by RPP per pass)
  const int gw = EPI == TC_EPI_QKV ? a.rope_gw : 1, hq = EPI == TC_EPI_QKV ? (a.hd >> 2) : 1;
  int rp_i0 = 0, rp_j0 = 0;
  if (EPI == TC_EPI_QKV && rope_s) { rp_i0 = (wt0 + prow) / gw; rp_j0 = (wt0 + prow) - rp_i0 * gw; }

  // SwiGLU gates in the thread == row domain (the gated pair sits in adjacent accumulator columns of one thread), so
  // only the bf16 result goes through the shared-memory transposition: a quarter of the staging traffic of the fp32
  // block - and shared-memory bandwidth (operand fill + tensor-core reads + staging) is what bounds these kernels.
  // The per-column bias / per-sample vectors of this warp's <= 4 chunks are parked behind the staging rows.
  constexpr int HST = 48;                           // bf16 staging row stride in bytes (32 B of data, conflict-free 16-byte stores)
  const uint32_t hb_s = st_s + 32 * HST;            // [4 chunks][32 columns] fp32
  if (EPI == TC_EPI_SWIGLU) {
    __syncwarp();                                   // the previous tile's row pass has finished reading this warp's staging area
    if (prow == 0) {
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) sts128(hb_s + ci * 128 + pcc * 4, bball[ci].x, bball[ci].y, bball[ci].z, bball[ci].w);
    }
    __syncwarp();
  }

  ptx::mbar_wait(tfull, aphase);
  ptx::tc_fence_after();

  auto chunk_body = [&](const int c, const float4 bb_ahead) {
    const int c0 = c * CW;
    uint32_t r[CW];
    ptx::tmem_ld_x32(tbase + c0, r);
    if (EPI == TC_EPI_SWIGLU) {
      const uint32_t bsrc = hb_s + (uint32_t)((c - half) >> 1) * 128;
      float4 b4[CW / 4];
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) b4[j] = lds128(bsrc + j * 16);   // same address in every lane: broadcast
      ptx::tmem_ld_wait();
      uint32_t h[CW / 4];
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const float v0 = __uint_as_float(r[4 * j]) * in_rstd + b4[j].x, v1 = __uint_as_float(r[4 * j + 1]) * in_rstd + b4[j].y;
        const float v2 = __uint_as_float(r[4 * j + 2]) * in_rstd + b4[j].z, v3 = __uint_as_float(r[4 * j + 3]) * in_rstd + b4[j].w;
        h[j] = pack_bf16(silu_fast(v0) * v1, silu_fast(v2) * v3);        // swiglu_ffn.py:33-35 on interleaved rows
      }
      __syncwarp();
      const uint32_t hw = st_s + (uint32_t)lane * HST;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hw), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hw + 16), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
      __syncwarp();
      // row pass: two lanes per row (16 bytes = 8 outputs each), 16 rows per store instruction
      bf16* ho = reinterpret_cast<bf16*>(a.out) + (size_t)(wrow0 + (lane >> 1)) * a.ldo + ((n0 + c0) >> 1) + (lane & 1) * 8;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(st_s + (uint32_t)((it * 16 + (lane >> 1)) * HST + (lane & 1) * 16)) : "memory");
        *reinterpret_cast<uint4*>(ho + (size_t)it * 16 * a.ldo) = v;
      }
      return;
    }
    float4 xcur[LPR];
    if (EPI == TC_EPI_GATED_RESIDUAL) {
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) xcur[ps] = xpre[ps];
    }
    const float4 bb = kBiasAhead ? bb_ahead : bb_n, g1 = g_n;
    const float4 cw = make_float4(nw_n.x * (1.f + ns_n.x), nw_n.y * (1.f + ns_n.y), nw_n.z * (1.f + ns_n.z),
                                  nw_n.w * (1.f + ns_n.w));
    if (!kBiasAhead) prefetch(EPI == TC_EPI_GATED_RESIDUAL ? gr_next(c, n_chunks) : c + 2);
    ptx::tmem_ld_wait();

    // stage the 32 x 32 fp32 block (thread == row), scaled by 1/rms when the norm is fused
    __syncwarp();
#pragma unroll
    for (int j4 = 0; j4 < CW / 4; ++j4)
      sts128(st_w + j4 * 16, __uint_as_float(r[j4 * 4]) * in_rstd, __uint_as_float(r[j4 * 4 + 1]) * in_rstd,
             __uint_as_float(r[j4 * 4 + 2]) * in_rstd, __uint_as_float(r[j4 * 4 + 3]) * in_rstd);
    __syncwarp();

    if (EPI == TC_EPI_GATED_RESIDUAL) {
      float* xo = xptr + c0;
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) {
        const float4 v = lds128(st_r + ps * RPP * SST);
        float4 x = xcur[ps];
        x.x = fmaf(g1.x, v.x + bb.x, x.x); x.y = fmaf(g1.y, v.y + bb.y, x.y);
        x.z = fmaf(g1.z, v.z + bb.z, x.z); x.w = fmaf(g1.w, v.w + bb.w, x.w);
        if (!(a.experiment & 4)) *reinterpret_cast<float4*>(xo + ps * xpass) = x;
        if (a.nx_w) {
          const float sq = sumsq4(x);   // s(chunk, lane & 7) of row ps * 4 + prow
          if (a.nx_coarse) ssacc[ps] = (c & 3) == 0 ? sq : ssacc[ps] + sq;
          else a.nx_part[(size_t)(wrow0 + ps * RPP + prow) * a.nx_part_ld + ((n0 + c0) >> 5) * 8 + (lane & 7)] = sq;
          if (!(a.experiment & 2))
          *reinterpret_cast<uint2*>(nxptr + c0 + ps * xpass) =
              make_uint2(pack_bf16(x.x * cw.x, x.y * cw.y), pack_bf16(x.z * cw.z, x.w * cw.w));
        }
      }
      if (a.nx_w && a.nx_coarse && (((c + 1) & 3) == 0 || c + 1 >= n_chunks)) {
        // the 128-column group is complete: butterfly over the row's eight lanes, one coarse slot per row
#pragma unroll
        for (int ps = 0; ps < LPR; ++ps) {
          float v = ssacc[ps];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          if ((lane & 7) == 0) a.nx_part[(size_t)(wrow0 + ps * RPP + prow) * a.nx_part_ld + ((n0 + c0) >> 7)] = v;
        }
      }
    } else if (EPI == TC_EPI_EMBED) {
      // x_embedder(x) + pos_embed (lightningdit.py:402): fp32 residual stream
      float* xo = reinterpret_cast<float*>(a.out) + (size_t)(wrow0 + prow) * a.ldo + n0 + c0 + pcc;
      const float* po = a.pos + (size_t)(wt0 + prow) * a.N + n0 + c0 + pcc;
      float4 p4[LPR];
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) p4[ps] = *reinterpret_cast<const float4*>(po + (size_t)ps * RPP * a.N);
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) {
        float4 v = lds128(st_r + ps * RPP * SST);
        v.x += bb.x + p4[ps].x; v.y += bb.y + p4[ps].y; v.z += bb.z + p4[ps].z; v.w += bb.w + p4[ps].w;
        *reinterpret_cast<float4*>(xo + (size_t)ps * RPP * a.ldo) = v;
      }
    } else if (EPI == TC_EPI_BIAS_F32) {
      // acc + bias (+ fp32 residual, may alias the output): adaLN rows, decoder convolutions with their skip connection
      const size_t off = (size_t)(wrow0 + prow) * a.ldo + n0 + c0 + pcc;
      float* xo = reinterpret_cast<float*>(a.out) + off;
      const size_t opass = (size_t)RPP * a.ldo;
      float4 r4[LPR];
      if (a.res) {
#pragma unroll
        for (int ps = 0; ps < LPR; ++ps) r4[ps] = *reinterpret_cast<const float4*>(a.res + off + ps * opass);
      }
      float4 gs = make_float4(0.f, 0.f, 0.f, 0.f), gq = gs;   // per-channel sum / sum of squares over this lane's 8 rows
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) {
        float4 v = lds128(st_r + ps * RPP * SST);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        if (a.res) { v.x += r4[ps].x; v.y += r4[ps].y; v.z += r4[ps].z; v.w += r4[ps].w; }
        *reinterpret_cast<float4*>(xo + ps * opass) = v;
        gs.x += v.x; gs.y += v.y; gs.z += v.z; gs.w += v.w;
        gq.x += v.x * v.x; gq.y += v.y * v.y; gq.z += v.z * v.z; gq.w += v.w * v.w;
      }
      if (a.gn_part) {
        // GroupNorm statistics of the tensor just written, for the normalisation that follows: this warp's 32 rows x
        // this chunk's 32 columns -> one (sum, sum of squares) per group, fixed reduction order (no atomics)
        float sg = (gs.x + gs.y) + (gs.z + gs.w), qg = (gq.x + gq.y) + (gq.z + gq.w);   // 4 channels of one group
        sg += __shfl_xor_sync(0xffffffffu, sg, 8);  qg += __shfl_xor_sync(0xffffffffu, qg, 8);    // rows: lane bits 3, 4
        sg += __shfl_xor_sync(0xffffffffu, sg, 16); qg += __shfl_xor_sync(0xffffffffu, qg, 16);
        for (int o = 1; o < (1 << (a.gn_cpg_shift - 2)); o <<= 1) {   // groups wider than 4 channels: neighbouring lanes
          sg += __shfl_xor_sync(0xffffffffu, sg, o);
          qg += __shfl_xor_sync(0xffffffffu, qg, o);
        }
        if (prow == 0 && ((lane & 7) & ((1 << (a.gn_cpg_shift - 2)) - 1)) == 0) {
          const int img = wrow0 / a.gn_hw, slab = (wrow0 - img * a.gn_hw) >> 5, slabs = a.gn_hw >> 5;
          const int g = (n0 + c0 + pcc) >> a.gn_cpg_shift;
          float2* o2 = reinterpret_cast<float2*>(a.gn_part) + ((size_t)img * slabs + slab) * 32 + g;
          *o2 = make_float2(sg, qg);
        }
      }
    } else if (EPI == TC_EPI_SWIGLU) {
      bf16* ho = reinterpret_cast<bf16*>(a.out) + (size_t)(wrow0 + prow) * a.ldo + ((n0 + c0 + pcc) >> 1);
      const size_t hpass = (size_t)RPP * a.ldo;
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) {
        float4 v = lds128(st_r + ps * RPP * SST);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        *reinterpret_cast<uint32_t*>(ho + ps * hpass) = pack_bf16(silu_fast(v.x) * v.y, silu_fast(v.z) * v.w);
      }
    } else if (EPI == TC_EPI_QKV) {
      const int D = qD;
      const int colc = n0 + c0;
      const int which = (colc >= D) + (colc >= 2 * D);  // chunk-uniform: D % 32 == 0
      if (which == 2) {
        // V^T: lane owns head-dim column j and 8 consecutive tokens -> one 16-byte store of (B,H,hdv,T).
        // Column reads of the staging block: rows 8 apart share banks, so the four token groups of a warp walk
        // their 8 tokens in rotated order (group tg starts at token 2*tg) and the packed pairs are rotated back.
        const int tg = lane & 3;
        uint32_t ro[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ro[e] = (uint32_t)(((e + 2 * tg) & 7) * SST);
        float bv4[4];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) bv4[ps] = bvec[colc + ps * 8 + (lane >> 2)];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int j = ps * 8 + (lane >> 2);
          const int rem = colc + j - 2 * D, h = (int)__umulhi((uint32_t)rem, hd_magic), d = rem - h * a.hd;
          const uint32_t sa = st_s + (uint32_t)(tg * 8) * SST + j * 4;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = lds32(sa + ro[e]) + bv4[ps];
          // pr[i] holds tokens 2*((i + tg) & 3), +1 ; wanted: pk[k] = tokens 2k, 2k+1 = pr[(k - tg) & 3]
          const uint32_t pr0 = pack_bf16(f[0], f[1]), pr1 = pack_bf16(f[2], f[3]), pr2 = pack_bf16(f[4], f[5]),
                         pr3 = pack_bf16(f[6], f[7]);
          const bool r1 = tg & 1, r2 = tg & 2;
          const uint32_t q0 = r1 ? pr3 : pr0, q1 = r1 ? pr0 : pr1, q2 = r1 ? pr1 : pr2, q3 = r1 ? pr2 : pr3;
          *reinterpret_cast<uint4*>(a.vt + (((size_t)wb * a.H + h) * a.hdv + d) * a.T + wt0 + tg * 8) =
              make_uint4(r2 ? q2 : q0, r2 ? q3 : q1, r2 ? q0 : q2, r2 ? q1 : q3);
        }
      } else {
        // q / k: bias, RoPE (pos_embed.py:38-42,135), head-major store
        const int rem = colc + pcc - which * D, h = (int)__umulhi((uint32_t)rem, hd_magic), d = rem - h * a.hd;
        bf16* dst = (which == 0 ? a.q : a.k) + (((size_t)wb * a.H + h) * a.T + wt0 + prow) * a.hdp + d;
        const size_t dpass = (size_t)RPP * a.hdp;
        if (rope_s) {
          // (cos, sin) of this lane's two rotation pairs from the shared-memory tables: pair r < hd/4 rotates by the
          // token's grid row, pair r >= hd/4 by its grid column
          const int r0 = d >> 1, r1 = r0 + 1;
          const uint32_t b0 = rope_s + (uint32_t)((r0 < hq ? r0 : gw * hq + r0 - hq) * 8);
          const uint32_t b1 = rope_s + (uint32_t)((r1 < hq ? r1 : gw * hq + r1 - hq) * 8);
          const bool c0 = r0 >= hq, c1 = r1 >= hq;
          int ti = rp_i0, tj = rp_j0;
#pragma unroll
          for (int ps = 0; ps < LPR; ++ps) {
            float2 cs0, cs1;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cs0.x), "=f"(cs0.y) : "r"(b0 + (uint32_t)((c0 ? tj : ti) * hq * 8)));
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cs1.x), "=f"(cs1.y) : "r"(b1 + (uint32_t)((c1 ? tj : ti) * hq * 8)));
            float4 v = lds128(st_r + ps * RPP * SST);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            float4 o;
            o.x = v.x * cs0.x - v.y * cs0.y;
            o.y = v.y * cs0.x + v.x * cs0.y;
            o.z = v.z * cs1.x - v.w * cs1.y;
            o.w = v.w * cs1.x + v.z * cs1.y;
            *reinterpret_cast<uint2*>(dst + ps * dpass) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            tj += RPP;
            if (tj >= gw) { tj -= gw; ++ti; }
          }
        } else if (a.rope_cs) {
          // tables that are not separable over the token grid: interleaved (cos, sin) per rotation pair from global memory
          const float* cs = a.rope_cs + ((size_t)(wt0 + prow) * (a.hd >> 1) + (d >> 1)) * 2;
          const size_t tpass = (size_t)RPP * a.hd;  // (hd/2 pairs) * 2 floats per token
          float4 cs_n[LPR];
#pragma unroll
          for (int ps = 0; ps < LPR; ++ps) cs_n[ps] = *reinterpret_cast<const float4*>(cs + ps * tpass);
#pragma unroll
          for (int ps = 0; ps < LPR; ++ps) {
            float4 v = lds128(st_r + ps * RPP * SST);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            float4 o;
            o.x = v.x * cs_n[ps].x - v.y * cs_n[ps].y;
            o.y = v.y * cs_n[ps].x + v.x * cs_n[ps].y;
            o.z = v.z * cs_n[ps].z - v.w * cs_n[ps].w;
            o.w = v.w * cs_n[ps].z + v.z * cs_n[ps].w;
            *reinterpret_cast<uint2*>(dst + ps * dpass) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
          }
        } else {
        const float* cp = a.cos ? a.cos + (size_t)(wt0 + prow) * a.hd + d : nullptr;
        const float* sp_ = a.cos ? a.sin + (size_t)(wt0 + prow) * a.hd + d : nullptr;
        const size_t tpass = (size_t)RPP * a.hd;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {  // two batches of four passes keep the register count in check
          float4 c4[4], s4[4];
          if (a.cos) {
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              c4[pp] = *reinterpret_cast<const float4*>(cp + (hf * 4 + pp) * tpass);
              s4[pp] = *reinterpret_cast<const float4*>(sp_ + (hf * 4 + pp) * tpass);
            }
          }
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const int ps = hf * 4 + pp;
            float4 v = lds128(st_r + ps * RPP * SST);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            if (a.cos) {
              float4 o;
              o.x = v.x * c4[pp].x - v.y * s4[pp].x;
              o.y = v.y * c4[pp].y + v.x * s4[pp].y;
              o.z = v.z * c4[pp].z - v.w * s4[pp].z;
              o.w = v.w * c4[pp].w + v.z * s4[pp].w;
              v = o;
            }
            *reinterpret_cast<uint2*>(dst + ps * dpass) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
          }
        }
        }
      }
    }
  };
  if (kBiasAhead) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const int c = half + 2 * ci;
      if (c < n_chunks) chunk_body(c, bball[ci]);
    }
  } else {
    if (EPI == TC_EPI_GATED_RESIDUAL) {
      for (int c = gr_first(half, n_chunks); c < n_chunks; c = gr_next(c, n_chunks)) chunk_body(c, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
      for (int c = half; c < n_chunks; c += 2) chunk_body(c, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

// ------------------------------------------------------------------------------------------------ TMA epilogue
// Gated residual with the residual stream moved by TMA (TC_EPI_GATED_RESIDUAL_TMA).  `xt_warps` epilogue warps - eight
// (two per TMEM lane quarter, alternate column chunks) where the epilogue dominates, four (one per quarter, the others
// retire at once; the shared memory saved keeps the fifth operand stage) where the main loop does - each with a ring of
// `xt_slots` 32 x 32 fp32 tiles (4 KB, 128-byte swizzle).  The tiles of the next chunks are in flight while the current one is processed; the update happens in
// place in the thread == row domain (lane = accumulator row = tile row; the swizzle makes 16-byte accesses of 32
// different rows conflict-free), then one elected lane stores the tile - and the bf16 operand tile of the next GEMM -
// with cp.async.bulk.tensor.  Nothing of the residual stream goes through the LSU / L1 miss path, whose limited
// outstanding-miss capacity (what the shared-memory carve-out leaves of L1) capped the register-prefetch epilogue.
// Same arithmetic as the other two paths: x' = fma(gate, acc + bias, x); operand = bf16(x' * w * (1 + scale));
// partial sums of x'^2 in the same grouping and order (sumsq4 per 4 columns, 8 column groups combined as a butterfly).
constexpr int XT_BYTES = 32 * 128;         // residual tile
constexpr int NXT_BYTES = 32 * 64;         // bf16 operand tile
constexpr int VEC_BYTES = 3 * 128;         // per chunk: [bias | gate | w*(1+scale)][32 columns]
constexpr int XT_MAX_SLOTS = 4;
// `warps` epilogue warps (4: one per TMEM lane quarter, every chunk; 8: two per quarter, alternate chunks)
__host__ __device__ constexpr int tma_epi_bytes(int slots, int warps) {
  return warps * (slots * XT_BYTES + NXT_BYTES + (32 / warps) * VEC_BYTES);
}

// (the residual / operand tensor maps are encoded with rank 4 - make_tmap_4d - so the 4-D instruction forms are used)
__device__ __forceinline__ void tma_load_tile(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(0)
      : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(0)
               : "memory");
}

struct TcSmemBars {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t xfull[8][XT_MAX_SLOTS];   // TMA epilogue: residual-tile ring, [epilogue warp][slot]
  uint32_t tmem_base;
};

// ------------------------------------------------------------------------------------------------ 2-CTA (cta_group::2)
namespace ptx2 {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const uint32_t bar_leader = ptx::smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
  const uint32_t bar_leader = ptx::smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows per CTA) * B^T (N/2 rows per CTA); issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the leader's prior MMAs retire) on the same-offset mbarrier of every CTA in `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(ptx::smem_u32(bar)),
      "r"(cta)
      : "memory");
}
}  // namespace ptx2

// TWO = true: CTA pairs (cluster of 2, cta_group::2).  A pair computes a 256 x BN tile: each CTA loads its own 128
// rows of A and HALF of the W tile (BN/2 rows); the leader's single thread issues M=256 MMAs that read both CTAs'
// shared memory and write each CTA's 128 x BN accumulator into its own TMEM.  Per flop this halves the W-tile
// shared-memory fill and L2 traffic, which is what limits (and powers) the 1-CTA kernel.
// Column tiles: n_tiles - 1 tiles of BN columns plus a last tile of BNlast <= BN columns (its own W tensor map, tmBl,
// with a shorter box): N = 1152 runs as 4 x 256 + 128 and N = 3456 as 13 x 256 + 128 instead of 6 / 18 tiles of 192 -
// the tensor core is fed 256-wide instructions wherever the problem allows it.
template <int EPI, int CW, bool TWO>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmX,
          const __grid_constant__ CUtensorMap tmNX, const TcGemmArgs a, const int BN, const int BNlast,
          const int nstages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_rows = TWO ? BN / 2 : BN;                 // W-tile rows this CTA loads
  const int stage_bytes = TC_A_BYTES + b_rows * 128;
  const uint32_t rank = TWO ? ptx2::cluster_ctarank() : 0u;   // 0 = leader
  const int tile_m = TWO ? 2 * TC_BM : TC_BM;           // rows per scheduled tile
  const int sched_id = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_n = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  float* staging = reinterpret_cast<float*>(smem + (size_t)nstages * stage_bytes);
  TcSmemBars* bars = reinterpret_cast<TcSmemBars*>(
      reinterpret_cast<uint8_t*>(staging) + (EPI == TC_EPI_GATED_RESIDUAL_TMA ? tma_epi_bytes(a.xt_slots, a.xt_warps) : 8 * 32 * (CW + 4) * 4));
  // QKV: compact RoPE tables (a few KB of constants) live in shared memory for the whole kernel
  float* rope_tab_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bars + 1) + 15) & ~(uintptr_t)15);
  const bool rope_smem = EPI == TC_EPI_QKV && a.rope_tab != nullptr;

  if (rope_smem) {
    const int n4 = a.rope_gw * (a.hd >> 2);   // float4 count: 2 tables * gw * hd/4 pairs * 2 floats / 4
    for (int i = threadIdx.x; i < n4; i += TC_THREADS)
      reinterpret_cast<float4*>(rope_tab_s)[i] = reinterpret_cast<const float4*>(a.rope_tab)[i];
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (a.M + tile_m - 1) / tile_m;
  const int n_tiles = (a.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (a.K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmBl);
    for (int i = 0; i < nstages; ++i) {
      ptx::mbar_init(&bars->full[i], 1);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->tfull[i], 1);
      // 2-CTA: both CTAs' epilogue threads release the leader; the TMA epilogue may run on four of the eight warps
      ptx::mbar_init(&bars->tempty[i], (TWO ? 2 : 1) * 32 * (EPI == TC_EPI_GATED_RESIDUAL_TMA ? a.xt_warps : 8));
    }
    if (EPI == TC_EPI_GATED_RESIDUAL_TMA) {
      ptx::prefetch_tmap(&tmX);
      ptx::prefetch_tmap(&tmNX);
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < XT_MAX_SLOTS; ++j) ptx::mbar_init(&bars->xfull[i][j], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (TWO) {
      ptx2::tmem_alloc2(&bars->tmem_base, 512);
      ptx2::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(&bars->tmem_base, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (TWO) ptx2::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // everything above overlapped the predecessor's tail (PDL); from here on its outputs are read
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
        const int m_blk = a.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles, n_blk = tile % n_tiles;
        const bool last_n = n_blk == n_tiles - 1 && BNlast != BN;   // narrower last column tile
        const int bn_t = last_n ? BNlast : BN;
        const int b_rows_t = TWO ? bn_t / 2 : bn_t;
        const CUtensorMap* tmb = last_n ? &tmBl : &tmB;
        const int a_row0 = m_blk * tile_m + (int)rank * TC_BM;   // first A row (= output pixel) of this CTA's tile
        // implicit-GEMM convolution: the tile is 128 consecutive pixels of one image (a row segment when W >= 128,
        // 128 / W whole rows otherwise); tap (dy, dx) shifts the box, TMA zero-fills what falls outside the image
        int cv_n = 0, cv_h = 0, cv_w = 0, cv_cpb = 1;
        if (a.conv_taps) {
          const int hw = a.conv_H * a.conv_W;
          cv_n = a_row0 / hw;
          const int rem = a_row0 - cv_n * hw;
          cv_h = rem / a.conv_W;
          cv_w = rem - cv_h * a.conv_W;
          cv_cpb = a.lda / TC_BK;
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + TC_A_BYTES;
          int tap = 0, cb = kb;
          if (a.conv_taps) { tap = kb / cv_cpb; cb = kb - tap * cv_cpb; }
          const int dy = a.conv_taps == 9 ? tap / 3 - 1 : 0, dx = a.conv_taps == 9 ? tap % 3 - 1 : 0;
          if (TWO) {
            // both CTAs' bytes are credited to the leader's barrier; only the leader arms it (with the pair's total)
            if (rank == 0) ptx::mbar_expect_tx(&bars->full[stage], (uint32_t)(2 * (TC_A_BYTES + b_rows_t * 128)));
            if (a.conv_taps)
              ptx2::tma_load_4d_2sm(sa, &tmA, &bars->full[stage], cb * TC_BK, cv_w + dx, cv_h + dy, cv_n);
            else
              ptx2::tma_load_2d_2sm(sa, &tmA, &bars->full[stage], kb * TC_BK, a_row0);
            ptx2::tma_load_2d_2sm(sb, tmb, &bars->full[stage], kb * TC_BK, n_blk * BN + (int)rank * b_rows_t);
          } else {
            ptx::mbar_expect_tx(&bars->full[stage], (uint32_t)(TC_A_BYTES + b_rows_t * 128));
            if (a.conv_taps)
              ptx::tma_load_4d(sa, &tmA, &bars->full[stage], cb * TC_BK, cv_w + dx, cv_h + dy, cv_n);
            else
              ptx::tma_load_2d(sa, &tmA, &bars->full[stage], kb * TC_BK, a_row0);
            ptx::tma_load_2d(sb, tmb, &bars->full[stage], kb * TC_BK, n_blk * BN);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc_main = ptx::make_idesc_bf16(tile_m, BN), idesc_last = ptx::make_idesc_bf16(tile_m, BNlast);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
        const uint32_t idesc = (tile % n_tiles == n_tiles - 1) ? idesc_last : idesc_main;
        ptx::mbar_wait(&bars->tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * 256;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = ptx::make_desc_k_sw128(sa);
          const uint64_t bdesc = ptx::make_desc_k_sw128(sa + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr>>4) field
            if (TWO)
              ptx2::umma_bf16_ss_2sm(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                     (uint32_t)((kb | k) != 0));
            else
              ptx::umma_bf16_ss(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                (uint32_t)((kb | k) != 0));
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if (TWO) ptx2::umma_commit_mc(&bars->empty[stage], 3); else ptx::umma_commit(&bars->empty[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (TWO) ptx2::umma_commit_mc(&bars->tfull[as], 3); else ptx::umma_commit(&bars->tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (8: two per TMEM lane quarter)
    const int q = warp & 3;            // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
    const int half = (warp - 2) >> 2;  // the two warps of a quarter take alternate column chunks
    float* st = staging + (size_t)(warp - 2) * 32 * (CW + 4);
    constexpr int LPR = CW / 4;    // lanes per row in the row-wise pass
    constexpr int RPP = 32 / LPR;  // rows per pass
    const int prow = lane / LPR;   // row (within a pass) and column group this lane owns in the row-wise pass
    const int pcc = (lane % LPR) * 4;
    // the warp's 32 rows are consecutive tokens of one sample (plain bias epilogues have no per-sample operands)
    const bool rows_one_sample = EPI == TC_EPI_BIAS_F32 ? true : (a.T > 0 && (a.T % 32) == 0);
    const bool fast_ok = rows_one_sample && (BN % 32 == 0) && epi_fast_supported<EPI>(a) &&
                         (EPI != TC_EPI_QKV || ((a.H * a.hd) % 32 == 0 && a.hd % 4 == 0));
    int as = 0;
    uint32_t aphase = 0;
    // releases an accumulator stage to the MMA-issuing thread (which lives in the leader CTA of a pair)
    auto release_acc = [&](int stage_idx) {
      ptx::tc_fence_before();
      if (TWO && rank != 0) ptx2::mbar_arrive_remote(&bars->tempty[stage_idx], 0);
      else ptx::mbar_arrive(&bars->tempty[stage_idx]);
    };
    if (EPI == TC_EPI_GATED_RESIDUAL_TMA) {
      // every tile is interior and the warp's 32 rows belong to one sample (launch_gemm_tc checked it)
      // Column chunks are owned in 128-column GROUPS (four 32-column chunks): with NH warps per TMEM lane quarter
      // (xt_warps / 4 = 1 or 2) warp `half` takes groups half, half + NH, ... of a tile and walks each group's chunks
      // in ascending order - so a coarse partial-sum slot, ((f0+f1)+f2)+f3, is built by one thread in a fixed order.
      const int NH = a.xt_warps >> 2;
      if (half < NH) {
      const int NS = a.xt_slots, ew = half * 4 + q, NW = a.xt_warps;
      const uint32_t epi_s = ptx::smem_u32(staging);
      const uint32_t xs = epi_s + (uint32_t)(ew * NS) * XT_BYTES;
      const uint32_t nxs = epi_s + (uint32_t)(NW * NS) * XT_BYTES + (uint32_t)ew * NXT_BYTES;
      const uint32_t vs = epi_s + (uint32_t)(NW * NS) * XT_BYTES + (uint32_t)NW * NXT_BYTES + (uint32_t)(ew * (32 / NW)) * VEC_BYTES;
      uint64_t* xbar = bars->xfull[ew];
      const bool has_nx = a.nx_w != nullptr;
      const bool coarse = a.nx_coarse != 0;
      auto tile_chunks = [&](int tile) { return ((tile % n_tiles == n_tiles - 1) ? BNlast : BN) >> 5; };
      // next chunk of this warp after chunk c of a tile with nch chunks; -1: the tile is finished
      auto next_chunk = [&](int c, int nch) {
        ++c;
        if ((c & 3) == 0 || c >= nch) {               // group finished
          c = (((c - 1) >> 2) + NH) << 2;
          if (c >= nch) c = -1;
        }
        return c;
      };
      auto first_chunk = [&](int nch) { return half * 4 < nch ? half * 4 : -1; };
      // this warp's (tile, chunk) stream; `ld_*` runs NS positions ahead of the chunk being processed
      int ld_tile = sched_id, ld_c = -1;
      auto ld_seek = [&]() {   // first tile from ld_tile on in which this warp owns a chunk
        while (ld_tile < num_tiles && (ld_c = first_chunk(tile_chunks(ld_tile))) < 0) ld_tile += sched_n;
      };
      ld_seek();
      auto ld_issue = [&](int slot) {   // lane 0 only
        if (ld_tile >= num_tiles) return;
        const int mb = a.reverse_m ? m_tiles - 1 - ld_tile / n_tiles : ld_tile / n_tiles, nb = ld_tile % n_tiles;
        const int row0 = mb * tile_m + (int)rank * TC_BM + q * 32, col0 = nb * BN + ld_c * 32;
        ptx::mbar_expect_tx(&xbar[slot], XT_BYTES);
        tma_load_tile(xs + (uint32_t)slot * XT_BYTES, &tmX, ptx::smem_u32(&xbar[slot]), col0, row0);
        ld_c = next_chunk(ld_c, tile_chunks(ld_tile));
        if (ld_c < 0) { ld_tile += sched_n; ld_seek(); }
      };
      if (lane == 0)
        for (int i = 0; i < NS; ++i) ld_issue(i);
      // chunk number ci of this warp in a tile (its valid chunks are a prefix of ci = 0 .. 7)
      auto chunk_of = [&](int ci) { return 4 * (half + (ci >> 2) * NH) + (ci & 3); };
      float pv_b[8], pv_g[8], pv_w[8], pv_s[8];   // next tile's bias / gate / norm weight / adaLN scale, lane = column
      auto vec_load = [&](int tile) {
        const int mb = a.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles, nb = tile % n_tiles;
        const int nch = tile_chunks(tile);
        const float* mr = a.mod + (size_t)((mb * tile_m + (int)rank * TC_BM + q * 32) / a.T) * a.mod_ld;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c = chunk_of(ci);
          if (c < nch) {
            const int col = nb * BN + c * 32 + lane;
            pv_b[ci] = a.bias ? a.bias[col] : 0.f;
            pv_g[ci] = mr[a.gate_off + col];
            if (has_nx) { pv_w[ci] = a.nx_w[col]; pv_s[ci] = mr[a.nx_scale_off + col]; }
          }
        }
      };
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) pv_b[ci] = pv_g[ci] = pv_w[ci] = pv_s[ci] = 0.f;
      if (sched_id < num_tiles) vec_load(sched_id);
      uint32_t slot = 0, sphase = 0;   // ring position of the chunk being processed
      for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
        const int m_blk = a.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles, n_blk = tile % n_tiles;
        const int m0 = m_blk * tile_m + (int)rank * TC_BM, n0 = n_blk * BN;
        const int n_chunks = tile_chunks(tile);
        const int wrow0 = m0 + q * 32;
        // per-column vectors of this warp's chunks (lane = column): bias, gate, and the next norm's w * (1 + scale).
        // They were requested one tile ago (vec_load below), so their L2 / DRAM latency is off the tile's critical path.
        __syncwarp();
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          if (chunk_of(ci) < n_chunks) {
            const uint32_t vb = vs + (uint32_t)ci * VEC_BYTES + (uint32_t)lane * 4;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(vb), "f"(pv_b[ci]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(vb + 128), "f"(pv_g[ci]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(vb + 256), "f"(has_nx ? pv_w[ci] * (1.f + pv_s[ci]) : 0.f) : "memory");
          }
        }
        __syncwarp();
        if (tile + sched_n < num_tiles) vec_load(tile + sched_n);
        ptx::mbar_wait(&bars->tfull[as], aphase);
        ptx::tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256;
        int c = first_chunk(n_chunks);
        if (c < 0) release_acc(as);      // a narrow tile none of whose groups is this warp's
        float gs[8];                     // this row's sums of squares over the current 128-column group (coarse slots)
#pragma unroll
        for (int j = 0; j < 8; ++j) gs[j] = 0.f;
        for (int ci = 0; c >= 0; ++ci) {
          const int cn = next_chunk(c, n_chunks);
          const uint32_t xt = xs + slot * XT_BYTES + (uint32_t)lane * 128;
          uint32_t r[32];
          ptx::tmem_ld_x32(tbase + (uint32_t)(c * 32), r);
          ptx::mbar_wait(&xbar[slot], sphase);          // the residual tile of this chunk has landed
          ptx::tmem_ld_wait();
          if (cn < 0) release_acc(as);                  // last read of this accumulator stage by this warp
          uint32_t nxw[16];
          float ss[8];
          const uint32_t vb = vs + (uint32_t)ci * VEC_BYTES;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const uint32_t xa = xt + (uint32_t)((j4 ^ (lane & 7)) << 4);
            float4 x = lds128(xa);
            const float4 b4 = lds128(vb + j4 * 16), g4 = lds128(vb + 128 + j4 * 16);   // broadcast reads
            x.x = fmaf(g4.x, __uint_as_float(r[j4 * 4]) + b4.x, x.x);
            x.y = fmaf(g4.y, __uint_as_float(r[j4 * 4 + 1]) + b4.y, x.y);
            x.z = fmaf(g4.z, __uint_as_float(r[j4 * 4 + 2]) + b4.z, x.z);
            x.w = fmaf(g4.w, __uint_as_float(r[j4 * 4 + 3]) + b4.w, x.w);
            sts128(xa, x.x, x.y, x.z, x.w);
            if (has_nx) {
              const float4 c4 = lds128(vb + 256 + j4 * 16);
              ss[j4] = sumsq4(x);
              nxw[j4 * 2] = pack_bf16(x.x * c4.x, x.y * c4.y);
              nxw[j4 * 2 + 1] = pack_bf16(x.z * c4.z, x.w * c4.w);
            }
          }
          if (has_nx) {
            const uint32_t na = nxs + (uint32_t)lane * 64;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(na + jj * 16), "r"(nxw[jj * 4]),
                           "r"(nxw[jj * 4 + 1]), "r"(nxw[jj * 4 + 2]), "r"(nxw[jj * 4 + 3])
                           : "memory");
            // partial sums of x'^2: eight fine slots per chunk, or one coarse slot per 128-column group
            float* prow_ = a.nx_part + (size_t)(wrow0 + lane) * a.nx_part_ld;
            const int gc = (n0 >> 5) + c;               // global chunk index
            if (!coarse) {
              *reinterpret_cast<float4*>(prow_ + gc * 8) = make_float4(ss[0], ss[1], ss[2], ss[3]);
              *reinterpret_cast<float4*>(prow_ + gc * 8 + 4) = make_float4(ss[4], ss[5], ss[6], ss[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) gs[j] = (c & 3) == 0 ? ss[j] : gs[j] + ss[j];
              if ((c & 3) == 3 || c + 1 >= n_chunks) prow_[gc >> 2] = fold8(gs);
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int col0 = n0 + c * 32;
            tma_store_tile(&tmX, xs + slot * XT_BYTES, col0, wrow0);
            if (has_nx) tma_store_tile(&tmNX, nxs, col0, wrow0);
            ptx::tma_store_commit();
            ptx::tma_store_wait_read0();   // both source tiles have been read: the slot takes the load NS chunks ahead
            ld_issue((int)slot);
          }
          __syncwarp();                    // no lane rewrites the operand tile before the store has read it
          if (++slot == (uint32_t)NS) { slot = 0; sphase ^= 1u; }
          c = cn;
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
      if (lane == 0) ptx::tma_store_wait_all();
      }
    } else
    for (int tile = sched_id; tile < num_tiles; tile += sched_n) {
      const int m_blk = a.reverse_m ? m_tiles - 1 - tile / n_tiles : tile / n_tiles, n_blk = tile % n_tiles;
      const int m0 = m_blk * tile_m + (int)rank * TC_BM, n0 = n_blk * BN;
      const int bn_t = (n_blk == n_tiles - 1) ? BNlast : BN;   // columns of this tile
      if (CW == 32 && fast_ok && (bn_t % 32 == 0) && m0 + TC_BM <= a.M && n0 + bn_t <= a.N) {
        // interior tile of a production epilogue: guard-free path
        epi_tile_fast<EPI>(a, ptx::smem_u32(st), tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256,
                           &bars->tfull[as], aphase, m0, n0, n_blk, bn_t, q, half, lane,
                           rope_smem ? ptx::smem_u32(rope_tab_s) : 0u);
        release_acc(as);
        as ^= 1;
        if (as == 0) aphase ^= 1;
        continue;
      }
      const int wrow0 = m0 + q * 32;     // first row of this warp
      const int my_row = wrow0 + lane;   // thread == row domain
      const int wb = wrow0 / a.T, wt0 = wrow0 - wb * a.T;  // sample / first token of the warp (valid if rows_one_sample)
      int n_chunks = (min(bn_t, a.N - n0) + CW - 1) / CW;

      // ---- prefetch for the gated-residual epilogue: the fp32 residual tile is loaded one chunk ahead so the
      //      HBM/L2 latency overlaps the TMEM drain of the previous chunk (and the wait for the MMA)
      float4 xpre[LPR];
      auto prefetch_x = [&](int c) {
        if (EPI == TC_EPI_GATED_RESIDUAL) {
#pragma unroll
          for (int ps = 0; ps < LPR; ++ps) {
            const int row = wrow0 + ps * RPP + prow, col = n0 + c * CW + pcc;
            xpre[ps] = (c < n_chunks && row < a.M && col < a.N)
                           ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.out) + (size_t)row * a.ldo + col)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      const bool grouped = EPI == TC_EPI_GATED_RESIDUAL && CW == 32;   // chunk ownership by 128-column groups
      prefetch_x(grouped ? gr_first(half, n_chunks) : half);
      // fused-norm consumer: this row's 1/rms from the producer's partial sums (fixed summation order)
      float in_rstd = 1.f;
      if ((EPI == TC_EPI_QKV || EPI == TC_EPI_SWIGLU || EPI == TC_EPI_GELU_BF16) && a.in_part) {
        float ssum = 0.f;
        if (my_row < a.M) ssum = part_total(a.in_part + (size_t)my_row * a.in_part_ld, a.in_nslots, a.in_fine);
        in_rstd = rsqrtf(ssum / (float)a.in_dim + 1e-6f);
      }
      // per-row bias source: per-sample vector sp[b,:] when the norm is fused, else the plain bias
      const float* brow = a.bias;
      if ((EPI == TC_EPI_QKV || EPI == TC_EPI_SWIGLU || EPI == TC_EPI_GELU_BF16) && a.in_part)
        brow = a.sp + (size_t)(min(my_row, a.M - 1) / a.T) * a.sp_ld;
      const float* bwarp = a.bias;  // same, for the warp's (single) sample in the row-wise pass
      if ((EPI == TC_EPI_QKV || EPI == TC_EPI_SWIGLU || EPI == TC_EPI_GELU_BF16) && a.in_part)
        bwarp = a.sp + (size_t)min(wb, (a.M - 1) / a.T) * a.sp_ld;

      // Per-column vectors of a chunk in the row-wise pass: bias (or the fused per-sample sp vector), and for the
      // gated residual the gate and the next norm's w*(1+scale).  sp / gate / scale live in the big `mod` matrix
      // (DRAM/L2 latency), so they are software-pipelined one chunk ahead like the residual tile.
      auto load_vecs = [&](int c, float4& bbv, float4& gv, float4& cwv) {
        bbv = gv = cwv = make_float4(0.f, 0.f, 0.f, 0.f);
        const int colv = n0 + c * CW + pcc;
        if (c >= n_chunks || colv >= a.N || EPI == TC_EPI_FINAL) return;
        if (a.bias && (EPI != TC_EPI_QKV || !a.qn_w)) bbv = *reinterpret_cast<const float4*>(bwarp + colv);
        if (EPI == TC_EPI_GATED_RESIDUAL && rows_one_sample && wrow0 < a.M) {
          gv = *reinterpret_cast<const float4*>(a.mod + (size_t)wb * a.mod_ld + a.gate_off + colv);
          if (a.nx_w) {
            const float4 w4 = *reinterpret_cast<const float4*>(a.nx_w + colv);
            const float4 s4 = *reinterpret_cast<const float4*>(a.mod + (size_t)wb * a.mod_ld + a.nx_scale_off + colv);
            cwv = make_float4(w4.x * (1.f + s4.x), w4.y * (1.f + s4.y), w4.z * (1.f + s4.z), w4.w * (1.f + s4.w));
          }
        }
      };
      float4 bb_n, g1_n, cw_n;
      load_vecs(grouped ? gr_first(half, n_chunks) : half, bb_n, g1_n, cw_n);
      float ssacc[LPR];
#pragma unroll
      for (int ps = 0; ps < LPR; ++ps) ssacc[ps] = 0.f;

      ptx::mbar_wait(&bars->tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 256;

      // qk-norm statistics are computed on demand per head (thread == row domain)
      float qk_mean = 0.f, qk_rstd = 1.f;
      int qk_head = -1;

      for (int c = grouped ? gr_first(half, n_chunks) : half; c < n_chunks; c = grouped ? gr_next(c, n_chunks) : c + 2) {
        const int c0 = c * CW;
        const int c_next = grouped ? gr_next(c, n_chunks) : c + 2;
        uint32_t r[CW];
        if (CW == 32) ptx::tmem_ld_x32(tbase + c0, r);
        else ptx::tmem_ld_x16(tbase + c0, r);
        float4 xcur[LPR];
        if (EPI == TC_EPI_GATED_RESIDUAL) {
#pragma unroll
          for (int ps = 0; ps < LPR; ++ps) xcur[ps] = xpre[ps];
          prefetch_x(c_next);
        }
        // per-column vectors of this chunk were requested one chunk ago; request the next chunk's now
        const float4 bb = bb_n, g1 = g1_n, cw = cw_n;
        load_vecs(c_next, bb_n, g1_n, cw_n);
        ptx::tmem_ld_wait();
        if ((EPI == TC_EPI_QKV || EPI == TC_EPI_SWIGLU || EPI == TC_EPI_GELU_BF16) && a.in_part) {
#pragma unroll
          for (int j = 0; j < CW; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * in_rstd);
        }

        if (EPI == TC_EPI_FINAL) {
          // unpatchify (lightningdit.py:376-389): token-major column (pa*p+pb)*Cout + c -> NCHW; lanes are
          // consecutive tokens so for p == 1 every store instruction writes 128 contiguous bytes.
          if (my_row < a.M) {
            const int rr = my_row / a.T, tok = my_row % a.T;
            const int seg = rr / a.rm.seg_rows, ii = rr - seg * a.rm.seg_rows;
            const int gw = a.img / a.p, ti = tok / gw, tj = tok % gw;
            float* ob = reinterpret_cast<float*>(a.out) + (size_t)(a.rm.out_off[seg] + ii) * a.Ckeep * a.img * a.img;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              const int col = n0 + c0 + j;
              if (col < a.N) {
                const int ch = col % a.Cout, ab = col / a.Cout, pa = ab / a.p, pb = ab % a.p;
                if (ch < a.Ckeep)
                  ob[((size_t)ch * a.img + ti * a.p + pa) * a.img + tj * a.p + pb] = __uint_as_float(r[j]) + a.bias[col];
              }
            }
          }
          continue;
        }

        bool v_chunk = false;
        if (EPI == TC_EPI_QKV) {
          const int colc = n0 + c0;
          const int D = a.H * a.hd;
          const int which = colc / D;
          v_chunk = which == 2;
          if (v_chunk && !(rows_one_sample && CW == 32)) {
            // generic V path: transposed scalar stores (B,H,hdv,T); lanes = consecutive tokens
            if (my_row < a.M) {
              const int b = my_row / a.T, t = my_row % a.T;
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                const int col = colc + j;
                if (col < a.N) {
                  const int rem = col - 2 * D, h = rem / a.hd, d = rem % a.hd;
                  a.vt[(((size_t)b * a.H + h) * a.hdv + d) * a.T + t] =
                      __float2bfloat16_rn(__uint_as_float(r[j]) + brow[col]);
                }
              }
            }
            continue;
          }
          // q / k with qk-norm: bias + normalisation in the thread == row domain (needs whole-head statistics).
          // Without qk-norm (production flags) the bias is added later in the coalesced row-wise pass.
          const float* nw = which == 0 ? a.qn_w : (which == 1 ? a.kn_w : nullptr);
          const float* nb = which == 0 ? a.qn_b : (which == 1 ? a.kn_b : nullptr);
          if (nw) {
            float bj[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) bj[j] = brow[colc + j];  // independent loads, issued back to back
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              const int col = colc + j;
              float v = __uint_as_float(r[j]) + bj[j];
              const int rem = col - which * D, h = rem / a.hd, d = rem % a.hd;
              if (h != qk_head) {
                qk_head = h;
                qk_head_stats(tbase, which * D + h * a.hd - n0, n0, a.hd, brow, in_rstd, a.qk_rms, qk_mean, qk_rstd);
              }
              v = (v - qk_mean) * qk_rstd * nw[d] + (nb ? nb[d] : 0.f);
              r[j] = __float_as_uint(v);
            }
          }
        }

        // ---- stage the 32 x CW fp32 block in shared memory (thread == row), then work row-wise (coalesced)
        __syncwarp();
#pragma unroll
        for (int j4 = 0; j4 < CW / 4; ++j4)
          *reinterpret_cast<float4*>(st + lane * (CW + 4) + j4 * 4) =
              make_float4(__uint_as_float(r[j4 * 4]), __uint_as_float(r[j4 * 4 + 1]), __uint_as_float(r[j4 * 4 + 2]),
                          __uint_as_float(r[j4 * 4 + 3]));
        __syncwarp();

        if (EPI == TC_EPI_QKV && v_chunk) {
          // V^T fast path: lane owns head-dim column d and 8 consecutive tokens -> one 16-byte store;
          // a warp instruction writes 8 columns x 64 contiguous bytes of (B,H,hdv,T)
          const int D = a.H * a.hd;
          const int b = wb, t0 = wt0;
          const int tg = lane & 3;
          float bv4[4];
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const int colb = n0 + c0 + ps * 8 + (lane >> 2);
            bv4[ps] = colb < a.N ? bwarp[colb] : 0.f;  // all four issued before the first store
          }
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) {
            const int j = ps * 8 + (lane >> 2);
            const int col = n0 + c0 + j;
            if (col < a.N && wrow0 + tg * 8 < a.M) {
              const int rem = col - 2 * D, h = rem / a.hd, d = rem % a.hd;
              const float bv = bv4[ps];
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = st[(tg * 8 + e) * (CW + 4) + j] + bv;
              uint4 pk = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
              *reinterpret_cast<uint4*>(a.vt + (((size_t)b * a.H + h) * a.hdv + d) * a.T + t0 + tg * 8) = pk;
            }
          }
          continue;
        }

        const int col = n0 + c0 + pcc;  // this lane's 4 columns in every pass
        const bool col_ok = col < a.N;

        if (EPI == TC_EPI_GATED_RESIDUAL) {
          // x = x + gate.unsqueeze(1) * f(...)   lightningdit.py:248-249 ; x was prefetched one chunk ahead
          // g1 = gate, cw = w_norm * (1 + scale) of the norm that follows this residual update (both prefetched)
#pragma unroll
          for (int ps = 0; ps < LPR; ++ps) {
            const int rr = ps * RPP + prow;
            const int row = wrow0 + rr;
            const bool ok = row < a.M && col_ok;
            float sq = 0.f;
            if (ok) {
              float4 v = *reinterpret_cast<const float4*>(st + rr * (CW + 4) + pcc);
              float4 g = g1;
              if (!rows_one_sample)
                g = *reinterpret_cast<const float4*>(a.mod + (size_t)(row / a.T) * a.mod_ld + a.gate_off + col);
              float4 x = xcur[ps];
              x.x = fmaf(g.x, v.x + bb.x, x.x); x.y = fmaf(g.y, v.y + bb.y, x.y);
              x.z = fmaf(g.z, v.z + bb.z, x.z); x.w = fmaf(g.w, v.w + bb.w, x.w);
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + (size_t)row * a.ldo + col) = x;
              if (a.nx_w) {
                sq = sumsq4(x);
                *reinterpret_cast<uint2*>(a.nx_A + (size_t)row * a.ldo + col) =
                    make_uint2(pack_bf16(x.x * cw.x, x.y * cw.y), pack_bf16(x.z * cw.z, x.w * cw.w));
              }
            }
            if (a.nx_w && CW == 32) {   // (launch_gemm_tc: N % 32 == 0, so a chunk's columns are all valid or all invalid)
              if (a.nx_coarse) ssacc[ps] = (c & 3) == 0 ? sq : ssacc[ps] + sq;
              else if (ok) a.nx_part[(size_t)row * a.nx_part_ld + ((n0 + c0) >> 5) * 8 + (lane & 7)] = sq;
            }
          }
          if (a.nx_w && CW == 32 && a.nx_coarse && (((c + 1) & 3) == 0 || c + 1 >= n_chunks)) {   // 128-column group complete
#pragma unroll
            for (int ps = 0; ps < LPR; ++ps) {
              float v = ssacc[ps];
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              const int row = wrow0 + ps * RPP + prow;
              if ((lane & 7) == 0 && row < a.M) a.nx_part[(size_t)row * a.nx_part_ld + ((n0 + c0) >> 7)] = v;
            }
          }
          continue;
        }

        if (EPI == TC_EPI_QKV) {
          // RoPE: t*cos + rotate_half(t)*sin with (a,b) -> (-b,a) on adjacent pairs; pos_embed.py:38-42,135.
          // All table loads of the chunk are issued before the first store (no load -> store -> load chains).
          const int D = a.H * a.hd;
          const int which = col / D, rem = col - which * D, h = rem / a.hd, d = rem % a.hd;
          float4 c4[LPR], s4[LPR];
          if (a.cos) {
#pragma unroll
            for (int ps = 0; ps < LPR; ++ps) {
              const int row = wrow0 + ps * RPP + prow;
              const int t = rows_one_sample ? wt0 + ps * RPP + prow : row % a.T;
              const bool ok = row < a.M && col_ok;
              c4[ps] = ok ? *reinterpret_cast<const float4*>(a.cos + (size_t)t * a.hd + d) : make_float4(1.f, 1.f, 1.f, 1.f);
              s4[ps] = ok ? *reinterpret_cast<const float4*>(a.sin + (size_t)t * a.hd + d) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int ps = 0; ps < LPR; ++ps) {
            const int rr = ps * RPP + prow;
            const int row = wrow0 + rr;
            if (row >= a.M || !col_ok) continue;
            float4 v = *reinterpret_cast<const float4*>(st + rr * (CW + 4) + pcc);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;  // zero when qk-norm already consumed the bias
            if (a.cos) {
              float4 o;
              o.x = v.x * c4[ps].x - v.y * s4[ps].x;
              o.y = v.y * c4[ps].y + v.x * s4[ps].y;
              o.z = v.z * c4[ps].z - v.w * s4[ps].z;
              o.w = v.w * c4[ps].w + v.z * s4[ps].w;
              v = o;
            }
            const int b = rows_one_sample ? wb : row / a.T;
            const int t = rows_one_sample ? wt0 + rr : row % a.T;
            bf16* dst = (which == 0 ? a.q : a.k) + (((size_t)b * a.H + h) * a.T + t) * a.hdp + d;
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
          }
          continue;
        }

#pragma unroll
        for (int ps = 0; ps < LPR; ++ps) {
          const int rr = ps * RPP + prow;
          const int row = wrow0 + rr;
          if (row >= a.M || !col_ok) continue;
          float4 v = *reinterpret_cast<const float4*>(st + rr * (CW + 4) + pcc);
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          if (EPI == TC_EPI_EMBED) {
            const float4 p4 = *reinterpret_cast<const float4*>(a.pos + (size_t)(row % a.T) * a.N + col);
            v.x += p4.x; v.y += p4.y; v.z += p4.z; v.w += p4.w;
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + (size_t)row * a.ldo + col) = v;
          } else if (EPI == TC_EPI_BIAS_F32) {
            if (a.res) {
              const float4 r4 = *reinterpret_cast<const float4*>(a.res + (size_t)row * a.ldo + col);
              v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
            }
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + (size_t)row * a.ldo + col) = v;
          } else if (EPI == TC_EPI_BIAS_BF16 || EPI == TC_EPI_GELU_BF16) {
            if (EPI == TC_EPI_GELU_BF16) {
              v.x = gelu_tanh_fast(v.x); v.y = gelu_tanh_fast(v.y); v.z = gelu_tanh_fast(v.z); v.w = gelu_tanh_fast(v.w);
            }
            uint2 pk = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(a.out) + (size_t)row * a.ldo + col) = pk;
          } else if (EPI == TC_EPI_SWIGLU) {
            // rows of w12 are interleaved at pack time: column 2j = x1[j] (gated), 2j+1 = x2[j]; swiglu_ffn.py:33-35
            uint32_t pk = pack_bf16(silu_fast(v.x) * v.y, silu_fast(v.z) * v.w);
            *reinterpret_cast<uint32_t*>(reinterpret_cast<bf16*>(a.out) + (size_t)row * a.ldo + (col >> 1)) = pk;
          }
        }
      }
      // release this accumulator stage to the MMA warp
      release_acc(as);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  ptx::tc_fence_before();
  if (TWO) ptx2::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (TWO) ptx2::tmem_dealloc2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host launch
// Programmatic dependent launch for the tcgen05 kernels (VAVAE_PDL: bit 0 = eager launches, bit 1 = launches captured
// into the small-batch CUDA graphs; default 2).  Eager launches at the benchmark batch sizes gain nothing (858-861 TFLOP/s
// either way: the persistent kernels' tails are short); inside the small-batch graphs, where a forward is ~230 kernels of
// ~10 us each, overlapping each prologue (barrier init, TMEM allocation, tensor-map prefetch) with the predecessor's
// tail is worth 1 - 4 % (XL/1 batch 1: 2.59 -> 2.48 ms, profiles/r2_small_batch.md).
bool pdl_enabled(cudaStream_t s) {
  static int on = -1;
  if (on < 0) {
    const char* e = std::getenv("VAVAE_PDL");
    on = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 2;   // bit 0: eager launches, bit 1: captured launches (default)
  }
  if (!on) return false;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st == cudaStreamCaptureStatusNone ? (on & 1) != 0 : (on & 2) != 0;
}

// Per-device state: function attributes and the SM count belong to a device (context), not to the process - a second
// engine on another GPU of the same process needs its own opt-in to > 48 KB of dynamic shared memory.
int current_device() {
  int dev;
  VV_CUDA(cudaGetDevice(&dev));
  VV_CHECK(dev >= 0 && dev < 64, "device index >= 64 not supported");
  return dev;
}
int num_sms() {
  static std::atomic<int> sms[64];
  const int dev = current_device();
  int n = sms[dev].load(std::memory_order_relaxed);
  if (!n) {
    VV_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
// One bit per device in a per-kernel mask.  The bit is set AFTER the attribute call, so a thread that sees it may
// launch; two threads racing on the first use both set the attribute, which is harmless.
bool device_marked(const std::atomic<uint64_t>& mask) {
  return (mask.load(std::memory_order_acquire) >> current_device()) & 1ull;
}
void mark_device(std::atomic<uint64_t>& mask) { mask.fetch_or(1ull << current_device(), std::memory_order_acq_rel); }

// 0: always the 1-CTA kernel; 1: CTA pairs only for the GEMMs whose main loop dominates (w12, w3, large-K QKV);
// 2 (default): CTA pairs wherever the shape allows it.  Measured on B200, whole 250-step sampler, power-capped
// (profiles/r1_two_cta_ab.md): B/1 741 / 771 / 785 TFLOP/s and XL/1 851 / 919 / 928 TFLOP/s for modes 0 / 1 / 2 --
// in isolation proj and the small-K QKV lose a few % with pairs, but under the 1 kW cap the halved W-tile
// traffic buys clock.
static int g_two_cta = 2;
static bool want_two_cta(const TcGemmArgs& a) {
  if (g_two_cta == 0 || a.M < 2 * TC_BM) return false;
  if (g_two_cta >= 2) return true;
  switch (a.epi == TC_EPI_GATED_RESIDUAL_TMA ? (int)TC_EPI_GATED_RESIDUAL : a.epi) {
    case TC_EPI_SWIGLU:
    case TC_EPI_GELU_BF16: return true;
    case TC_EPI_GATED_RESIDUAL: return a.K >= 2 * a.N;
    case TC_EPI_QKV: return a.K >= 1024;
    default: return false;
  }
}
// VAVAE_GEMM_BN (developer knob for same-box A/B runs): unset = 256-wide tiles plus one narrower last tile where N is
// not a multiple of 256 ("mixed"); "legacy" = one width per problem (256 / 192 / 128 by tile efficiency); a number
// forces that width for the layer GEMMs.
static int bn_mode() {
  static int mode = -2;
  if (mode == -2) {
    const char* e = std::getenv("VAVAE_GEMM_BN");
    if (!e || !e[0]) mode = 0;
    else if (e[0] == 'l') mode = -1;
    else mode = std::atoi(e);
  }
  return mode;
}
static bool layer_gemm(const TcGemmArgs& a) {
  if (a.conv_taps || a.gn_part) return false;
  const bool gated = a.epi == TC_EPI_GATED_RESIDUAL || a.epi == TC_EPI_GATED_RESIDUAL_TMA;
  // (attn.proj used to prefer 192-wide tiles with the register-prefetch epilogue; with the TMA epilogue the 256-wide
  //  tiling measured +8 % isolated and neutral in the sampler, and whole 128-column groups per tile are what coarse
  //  partial-sum slots need)
  return gated || a.epi == TC_EPI_SWIGLU || a.epi == TC_EPI_GELU_BF16 ||
         (a.epi == TC_EPI_QKV && !a.qn_w && (a.H * a.hd) % 32 == 0);
}
// Width of the last column tile: the remainder when the problem is tiled 256-wide and the remainder is a whole number
// of 32-column epilogue chunks; otherwise the main width (a partial last tile then relies on TMA zero fill).
int gemm_tc_last_bn(const TcGemmArgs& a, int BN) {
  const int rem = a.N % BN;
  if (bn_mode() == 0 && layer_gemm(a) && BN == 256 && rem != 0 && rem % 32 == 0) return rem;
  return BN;
}
// Small-M launches (a batch-1 forward has M = 256 token rows): with 256-wide tiles a layer GEMM would occupy 10 - 48 of
// the 148 SMs and run at the latency of one long main loop.  Narrow tiles spread the columns over (almost) one full
// wave of CTAs instead: the widest multiple of 32 that divides N and keeps cdiv(M, 128) * N / bn <= 148.  Safe because
// no result depends on the tile width (per-element K order is fixed; RMSNorm partial sums: see TcGemmArgs::nx_part).
// The gated-residual producers stay on multiples of 128 columns (coarse partial-sum slots): with narrower tiles they
// would write fine slots, and what their consumers then re-read per tile costs what the extra CTAs gain
// (XL/1 batch 1 / 2 / 4 / 8: 2.56 / 2.63 / 2.80 / 3.27 ms -> 2.20 / 2.20 / 2.52 / 2.97; every GEMM narrow: 2.23 / 2.27 /
// 2.86 / 2.97; consumers only: 2.50 / 2.53 / 2.80 / 3.27 - gpurun_out/r2_narrow2.log).
// VAVAE_GEMM_NARROW (developer knob, same-box A/B): 0 = off, 1 = every layer GEMM, 2 = consumers only, 3 = default.
static int narrow_bn(const TcGemmArgs& a, int regular_bn) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = std::getenv("VAVAE_GEMM_NARROW"); enabled = e ? std::atoi(e) : 3; }
  if (!enabled || a.N % 32 != 0) return regular_bn;
  const bool gated = a.epi == TC_EPI_GATED_RESIDUAL || a.epi == TC_EPI_GATED_RESIDUAL_TMA;
  if (gated && enabled == 2) return regular_bn;
  const int row_ctas = cdiv(a.M, TC_BM);
  if (row_ctas * cdiv(a.N, regular_bn) >= 111) return regular_bn;   // three quarters of the SMs are busy anyway
  int best = regular_bn, best_ctas = row_ctas * cdiv(a.N, regular_bn);
  for (int bn = 32; bn <= 256; bn += 32) {
    if (a.N % bn) continue;
    if (gated && enabled == 3 && bn % 128 != 0) continue;
    const int ctas = row_ctas * (a.N / bn);
    if (ctas <= 148 && ctas >= best_ctas) { best = bn; best_ctas = ctas; }   // ties: the wider tile
  }
  return best;
}

int gemm_tc_pick_bn(const TcGemmArgs& a) {
  if (layer_gemm(a) && bn_mode() == 0) {
    const int regular = (a.N > 256 && a.N % 256 != 0 && (a.N % 256) % 32 == 0) ? 256 : (a.N % 256 == 0 ? 256 : 0);
    if (regular) {
      const int nb = narrow_bn(a, regular);
      if (nb != regular) return nb;
    }
  }
  if (layer_gemm(a)) {
    const int mode = bn_mode();
    if (mode > 0 && mode % 32 == 0 && mode <= 256) return mode;
    if (mode == 0 && a.N > 256 && a.N % 256 != 0 && (a.N % 256) % 32 == 0) return 256;
  }
  if (a.epi == TC_EPI_FINAL) {
    int bn = ((a.N + 15) / 16) * 16;
    VV_CHECK(bn <= 256, "final linear wider than 256 columns is not supported");
    return bn;
  }
  if (a.epi == TC_EPI_QKV && a.qn_w) {
    // whole heads per tile so the per-head statistics live in one accumulator tile
    int best = 0;
    for (int bn = 16; bn <= 256; bn += 16)
      if (bn % a.hd == 0 && (a.H * a.hd) % bn == 0) best = bn;
    VV_CHECK(best > 0, "qk-norm epilogue: no tile width <= 256 holds whole heads (head_dim must divide a multiple of 16)");
    return best;
  }
  if (a.N % 256 == 0 && !(a.epi == TC_EPI_QKV && a.qn_w)) return 256;
  // The choice depends on N only - never on M: with the fused norm the per-row partial sums are grouped by column
  // tile, so a batch-size-dependent tile width would make results depend on how samples are sharded over GPUs.
  const int cands[3] = {256, 192, 128};
  int best = 128;
  double best_eff = -1.0;
  for (int i = 0; i < 3; ++i) {
    int bn = cands[i];
    if (a.epi == TC_EPI_QKV && (a.H * a.hd) % 16 != 0) continue;
    double eff = (double)a.N / (double)(cdiv(a.N, bn) * bn);
    // wider tiles read less shared memory per flop (B200: 128 B/clk/SM smem vs 8192 flop/clk/SM tensor)
    eff *= (bn >= 192 ? 1.0 : 0.93);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = bn; }
  }
  if (a.N <= 128) best = ((a.N + 15) / 16) * 16;
  return best;
}

void set_gemm_two_cta(int on) { g_two_cta = on; }
int get_gemm_two_cta() { return g_two_cta; }

template <int EPI, int CW, bool TWO>
static void launch_inst2(const TcGemmArgs& a, int BN, cudaStream_t s) {
  const int BNlast = gemm_tc_last_bn(a, BN);
  const int b_rows = TWO ? BN / 2 : BN;
  const int stage_bytes = TC_A_BYTES + b_rows * 128;
  const int rope_bytes = (EPI == TC_EPI_QKV && a.rope_tab) ? 2 * a.rope_gw * (a.hd / 4) * 8 + 16 : 0;
  const int epi_bytes = EPI == TC_EPI_GATED_RESIDUAL_TMA ? tma_epi_bytes(a.xt_slots, a.xt_warps) : 8 * 32 * (CW + 4) * 4;
  const int fixed = epi_bytes + (int)sizeof(TcSmemBars) + 1024 + rope_bytes;
  int nstages = (227 * 1024 - fixed) / stage_bytes;
  if (nstages > TC_MAX_STAGES) nstages = TC_MAX_STAGES;
  {
    static int cap = -1;   // VAVAE_GEMM_MAX_STAGES: developer experiment (how much does a shallower operand pipeline cost?)
    if (cap < 0) { const char* e = std::getenv("VAVAE_GEMM_MAX_STAGES"); cap = e ? std::atoi(e) : 0; }
    if (cap >= 2 && nstages > cap) nstages = cap;
    static int pcap = -1;  // VAVAE_GEMM_PROJ_STAGES: same, attn.proj-shaped gated-residual launches only (K <= N)
    if (pcap < 0) { const char* e = std::getenv("VAVAE_GEMM_PROJ_STAGES"); pcap = e ? std::atoi(e) : 0; }
    if (pcap >= 2 && nstages > pcap && a.K <= a.N &&
        (EPI == TC_EPI_GATED_RESIDUAL || EPI == TC_EPI_GATED_RESIDUAL_TMA)) nstages = pcap;
  }
  VV_CHECK(nstages >= 2, "gemm_tc: tile too large for shared memory");
  const size_t smem = (size_t)nstages * stage_bytes + fixed;
  static std::atomic<uint64_t> attr_devices{0};   // one bit per device, per kernel instantiation
  if (!device_marked(attr_devices)) {
    VV_CUDA(cudaFuncSetAttribute(k_gemm_tc<EPI, CW, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    mark_device(attr_devices);
  }
  CUtensorMap tmA;
  if (a.conv_taps) {
    // 128 consecutive pixels: a row segment (W >= 128), whole rows (W < 128 <= H*W) or whole images (H*W < 128)
    const uint32_t bw = a.conv_W >= 128 ? 128u : (uint32_t)a.conv_W;
    const uint32_t bh = std::min<uint32_t>((uint32_t)a.conv_H, 128u / bw);
    tmA = make_tmap_nhwc_bf16(a.A, (uint64_t)a.lda, (uint64_t)a.conv_W, (uint64_t)a.conv_H,
                              (uint64_t)(a.M / (a.conv_H * a.conv_W)), bw, bh, 128u / (bw * bh));
  } else {
    tmA = make_tmap_2d_bf16(a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda * 2, TC_BK, TC_BM);
  }
  CUtensorMap tmB = make_tmap_2d_bf16(a.W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldw * 2, TC_BK, (uint32_t)b_rows);
  CUtensorMap tmBl = BNlast == BN ? tmB
                                  : make_tmap_2d_bf16(a.W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldw * 2, TC_BK,
                                                      (uint32_t)(TWO ? BNlast / 2 : BNlast));
  // TMA epilogue: 32 x 32 tiles of the fp32 residual stream (128-byte swizzle) and of the next GEMM's bf16 operand
  CUtensorMap tmX = tmA, tmNX = tmA;
  if (EPI == TC_EPI_GATED_RESIDUAL_TMA) {
    const uint64_t xd[4] = {(uint64_t)a.N, (uint64_t)a.M, 1, 1};
    const uint64_t xstr[3] = {(uint64_t)a.ldo * 4, (uint64_t)a.ldo * 4 * a.M, (uint64_t)a.ldo * 4 * a.M};
    const uint32_t xbox[4] = {32, 32, 1, 1};
    tmX = make_tmap_4d(a.out, 4, xd, xstr, xbox, 128);
    if (a.nx_w) {
      const uint64_t nstr[3] = {(uint64_t)a.ldo * 2, (uint64_t)a.ldo * 2 * a.M, (uint64_t)a.ldo * 2 * a.M};
      tmNX = make_tmap_4d(a.nx_A, 2, xd, nstr, xbox, 0);
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled(s)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (!TWO) {
    const int tiles = cdiv(a.M, TC_BM) * cdiv(a.N, BN);
    cfg.gridDim = dim3(tiles < num_sms() ? tiles : num_sms());
    cfg.attrs = attr;
    cfg.numAttrs = na;
    VV_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_tc<EPI, CW, false>, tmA, tmB, tmBl, tmX, tmNX, a, BN, BNlast, nstages));
  } else {
    const int tiles = cdiv(a.M, 2 * TC_BM) * cdiv(a.N, BN);
    int pairs = std::min(tiles, num_sms() / 2);
    {
      static int max_pairs = -1;   // VAVAE_GEMM_MAX_PAIRS: developer experiment (is the operand feed limited per SM or chip-wide?)
      if (max_pairs < 0) { const char* e = std::getenv("VAVAE_GEMM_MAX_PAIRS"); max_pairs = e ? std::atoi(e) : 0; }
      if (max_pairs > 0) pairs = std::min(pairs, max_pairs);
    }
    cfg.gridDim = dim3(2 * pairs);
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    VV_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_tc<EPI, CW, true>, tmA, tmB, tmBl, tmX, tmNX, a, BN, BNlast, nstages));
  }
  VV_LAUNCH_CHECK();
}

template <int EPI, int CW>
static void launch_inst(const TcGemmArgs& a, int BN, cudaStream_t s) {
  // CTA pairs need two full 128-row halves worth of work and a W tile that splits into 8-row-aligned halves
  const bool two = want_two_cta(a) && (BN % 16 == 0) && BN >= 32;
  if (two) launch_inst2<EPI, CW, true>(a, BN, s);
  else launch_inst2<EPI, CW, false>(a, BN, s);
}

template <int EPI>
static void launch_epi(const TcGemmArgs& a, int BN, int cw, cudaStream_t s) {
  if (cw == 32) launch_inst<EPI, 32>(a, BN, s);
  else launch_inst<EPI, 16>(a, BN, s);
}

// How a gated-residual launch runs: register-prefetch epilogue or TMA epilogue (warps, ring depth), and which kind
// of partial-sum slots it writes.  One function, so that the engine can ask what its consumer will find.
struct GatedPlan { bool tma; int warps, slots, coarse; };
static GatedPlan gated_plan(const TcGemmArgs& a, int BN) {
  GatedPlan p = {false, 8, 2, 0};
  // VAVAE_GEMM_TMA_EPI (developer knob, same-box A/B): 0 = never; 1, 2 (default) = every eligible launch - eight epilogue
  // warps for attn.proj (K <= N: epilogue-bound), four for mlp.w3 (K > N: main-loop-bound, the four idle warps' shared
  // memory keeps all five operand stages): XL/1 w3 1230 -> 1320, B/1 w3 1240 -> 1280 TFLOP/s isolated; 3 = XL-only rule
  static int tma_epi = -1;
  if (tma_epi < 0) { const char* e = std::getenv("VAVAE_GEMM_TMA_EPI"); tma_epi = e ? std::atoi(e) : 2; }
  static int exp_mode = -1;
  if (exp_mode < 0) { const char* e = std::getenv("VAVAE_GEMM_EXPERIMENT"); exp_mode = e ? std::atoi(e) : 0; }
  const int BNl = gemm_tc_last_bn(a, BN);
  // guard-free geometry: whole 256-row tile pairs, whole 32-column chunks, 32-row groups inside one sample, 16-byte pitches
  const bool eligible = BN % 32 == 0 && BN <= 256 && BNl % 32 == 0 && a.T > 0 && a.T % 32 == 0 && a.M % (2 * TC_BM) == 0 &&
                        (a.N % BN == 0 || a.N % BN == BNl) && a.ldo % 8 == 0 && a.N <= a.ldo &&
                        (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                        (!a.nx_w || (reinterpret_cast<uintptr_t>(a.nx_A) & 15) == 0) && a.mod != nullptr;
  // 3: the rule before the tile-ahead vectors (K <= N or K >= 3072 only), kept for A/B
  const bool wanted = tma_epi == 1 || tma_epi == 2 || BN <= 128 || (tma_epi == 3 && (a.K <= a.N || a.K >= 3072));
  p.tma = tma_epi != 0 && exp_mode == 0 && eligible && wanted;
  if (p.tma) {
    static int xt_env = -1;   // VAVAE_GEMM_XT_SLOTS: developer knob, ring depth per epilogue warp
    if (xt_env < 0) { const char* e = std::getenv("VAVAE_GEMM_XT_SLOTS"); xt_env = e ? std::atoi(e) : 0; }
    p.slots = (xt_env >= 2 && xt_env <= XT_MAX_SLOTS) ? xt_env : 2;
    p.warps = (a.K <= a.N && BN > 128) ? 8 : 4;
    static int xw_env = -1;   // VAVAE_GEMM_XT_WARPS: developer knob (4 / 8), wide tiles only
    if (xw_env < 0) { const char* e = std::getenv("VAVAE_GEMM_XT_WARPS"); xw_env = e ? std::atoi(e) : 0; }
    if ((xw_env == 4 || xw_env == 8) && BN > 128) p.warps = xw_env;
  }
  // a CTA (and within it one warp) must own whole 128-column groups
  p.coarse = (BN % 128 == 0 && BNl % 128 == 0) ? 1 : 0;
  return p;
}
int gemm_tc_part_coarse(const TcGemmArgs& a) { return gated_plan(a, gemm_tc_pick_bn(a)).coarse; }
// host-only view of the decisions above, for tests of the tile-selection rules (no CUDA call)
void gemm_tc_plan_query(const TcGemmArgs& a, int out[6]) {
  const int BN = gemm_tc_pick_bn(a);
  out[0] = BN;
  out[1] = gemm_tc_last_bn(a, BN);
  out[2] = out[3] = out[4] = out[5] = 0;
  if (a.epi == TC_EPI_GATED_RESIDUAL) {
    const GatedPlan p = gated_plan(a, BN);
    out[2] = p.tma; out[3] = p.tma ? p.warps : 8; out[4] = p.tma ? p.slots : 0; out[5] = p.coarse;
  }
}
int gemm_tc_part_ld(int n_cols, int coarse) { return coarse ? 16 : n_cols / 4; }

void launch_gemm_tc(const TcGemmArgs& a, cudaStream_t s) {
  VV_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "gemm_tc: empty problem");
  VV_CHECK(a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm_tc: row pitch must be a multiple of 8 bf16 (16 bytes)");
  if (a.conv_taps) {
    VV_CHECK(a.conv_taps == 1 || a.conv_taps == 9, "gemm_tc: convolution taps must be 1 or 9");
    VV_CHECK(a.lda % TC_BK == 0 && a.K == a.conv_taps * a.lda, "gemm_tc: conv channels % 64 and K = taps * channels");
    const int hw = a.conv_H * a.conv_W;
    VV_CHECK((hw % TC_BM == 0 || TC_BM % hw == 0) && a.M % hw == 0, "gemm_tc: conv image size must divide or be a multiple of 128 pixels");
    VV_CHECK(a.conv_W >= 128 ? a.conv_W % 128 == 0 : 128 % a.conv_W == 0, "gemm_tc: conv width must divide or be a multiple of 128");
  }
  VV_CHECK(a.res == nullptr || a.epi == TC_EPI_BIAS_F32, "gemm_tc: residual input only with the fp32 bias epilogue");
  if (a.gn_part) {
    const int bn = gemm_tc_pick_bn(a);
    VV_CHECK(a.epi == TC_EPI_BIAS_F32 && a.M % TC_BM == 0 && a.N % bn == 0 && bn % 32 == 0 && a.gn_hw % 32 == 0 &&
                 a.M % a.gn_hw == 0 && a.gn_cpg_shift >= 2 && a.gn_cpg_shift <= 5,
             "gemm_tc: GroupNorm statistics need interior tiles, 32-row slabs inside one image and 4..32 channels per group");
  }
  const int BN = gemm_tc_pick_bn(a);
  int cw = (BN % 32 == 0) ? 32 : 16;
  if (a.epi == TC_EPI_QKV && (a.H * a.hd) % 32 != 0) cw = 16;
  VV_CHECK(BN % cw == 0, "gemm_tc: tile width / chunk width mismatch");
  if (a.epi != TC_EPI_FINAL) VV_CHECK(a.N % 4 == 0, "gemm_tc: N must be a multiple of 4");
  switch (a.epi) {
    case TC_EPI_BIAS_F32: launch_epi<TC_EPI_BIAS_F32>(a, BN, cw, s); break;
    case TC_EPI_BIAS_BF16: launch_epi<TC_EPI_BIAS_BF16>(a, BN, cw, s); break;
    case TC_EPI_GATED_RESIDUAL: {
      static int exp_mode = -1;
      if (exp_mode < 0) { const char* e = std::getenv("VAVAE_GEMM_EXPERIMENT"); exp_mode = e ? std::atoi(e) : 0; }
      TcGemmArgs b = a;
      b.experiment = exp_mode;
      VV_CHECK(!a.nx_w || (a.N % 32 == 0 && cw == 32), "gemm_tc: the fused next-norm producer needs N % 32 == 0");
      const GatedPlan plan = gated_plan(a, BN);
      b.nx_coarse = plan.coarse;
      VV_CHECK(!a.nx_w || a.nx_part_ld >= gemm_tc_part_ld(a.N, plan.coarse), "gemm_tc: partial-sum row pitch too small");
      VV_CHECK(!a.nx_w || !plan.coarse || a.N <= 2048, "gemm_tc: coarse partial sums cover at most 2048 columns");
      if (plan.tma) {
        b.epi = TC_EPI_GATED_RESIDUAL_TMA;
        b.xt_slots = plan.slots;
        b.xt_warps = plan.warps;
        launch_epi<TC_EPI_GATED_RESIDUAL_TMA>(b, BN, cw, s);
      } else {
        launch_epi<TC_EPI_GATED_RESIDUAL>(b, BN, cw, s);
      }
      break;
    }
    case TC_EPI_SWIGLU: launch_epi<TC_EPI_SWIGLU>(a, BN, cw, s); break;
    case TC_EPI_GELU_BF16: launch_epi<TC_EPI_GELU_BF16>(a, BN, cw, s); break;
    case TC_EPI_QKV:
      VV_CHECK(a.hd % 8 == 0 && (a.H * a.hd) % 16 == 0, "qkv epilogue: head_dim % 8 and hidden % 16 required");
      VV_CHECK(!a.rope_tab || (a.rope_gw >= 4 && a.rope_gw * a.rope_gw == a.T && a.rope_gw * (a.hd / 4) * 16 <= 32 * 1024),
               "qkv epilogue: compact RoPE tables need a square token grid of width >= 4 and at most 32 KB");
      {
        static int rope_smem = -1;   // VAVAE_ROPE_SMEM=0: developer knob, RoPE rows from global memory (same-box A/B)
        if (rope_smem < 0) { const char* e = std::getenv("VAVAE_ROPE_SMEM"); rope_smem = (e && e[0] == '0') ? 0 : 1; }
        if (!rope_smem && a.rope_tab) {
          TcGemmArgs b = a;
          b.rope_tab = nullptr;
          launch_epi<TC_EPI_QKV>(b, BN, cw, s);
          break;
        }
      }
      launch_epi<TC_EPI_QKV>(a, BN, cw, s);
      break;
    case TC_EPI_FINAL: launch_epi<TC_EPI_FINAL>(a, BN, cw, s); break;
    case TC_EPI_EMBED: launch_epi<TC_EPI_EMBED>(a, BN, cw, s); break;
    default: throw Error("gemm_tc: unknown epilogue");
  }
}

}  // namespace vv
