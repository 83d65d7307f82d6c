// tcgen05 flash attention for sm_100a:  O = softmax(Q K^T / sqrt(hd)) V   (lightningdit.py:77-88)
// One CTA per (batch*head, 128-query tile), keys processed in blocks of 256 with an online softmax.
//   S = Q K^T   : tcgen05.mma  M=128, N=keys (<=256), K=hd padded to 16        -> TMEM columns [0,256)
//   softmax     : 128 threads, thread == query row, S read with tcgen05.ld, exp2 on fp32
//   P           : bf16 probabilities written back into TMEM (packed two per column, aliasing S columns
//                 [0,128)) and consumed as the TMEM A-operand of the second MMA
//   O = P V     : tcgen05.mma  M=128, N=hd padded to 16, K=keys                 -> TMEM columns [128,256)
// Layouts (all K-major, 128B-swizzled TMA tiles): q,k (B,H,T,hdp) ; v stored transposed (B,H,hdv,T) by the
// QKV GEMM epilogue so that P V needs no MN-major operand.  256 TMEM columns and <=113 KB of shared memory
// per CTA so two CTAs share an SM and overlap each other's TMA / MMA / softmax phases.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace vv {

constexpr int AT_BQ = 128;   // queries per CTA
constexpr int AT_BKV = 256;  // keys per block
constexpr int AT_THREADS = 160;

struct AtBars {
  uint64_t q_full, kv_full, s_full, p_full, o_full, o_empty;
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T  (A operand from tensor memory)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// P_IN_TMEM = true : P goes to TMEM (TS MMA), 2 CTAs / SM.
// P_IN_TMEM = false: P goes to 128B-swizzled shared memory (SS MMA) - conservative variant kept for bring-up.
template <bool P_IN_TMEM, bool MULTI>
__global__ void __launch_bounds__(AT_THREADS, (P_IN_TMEM && !MULTI) ? 2 : 1)
k_attention_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const TcAttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nka = (a.hdp + 63) / 64;               // 64-column atoms along head_dim for Q / K
  const int q_bytes = nka * AT_BQ * 128;
  const int k_bytes = nka * AT_BKV * 128;
  const int v_atom = a.hdv * 128;                  // one 64-key atom of V^T
  const int v_bytes = 4 * v_atom;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_bytes;
  uint8_t* sV = sK + k_bytes;
  uint8_t* sP = sV + v_bytes;                      // only used when !P_IN_TMEM (4 atoms x 16 KiB)
  AtBars* bars = reinterpret_cast<AtBars*>(sP + (P_IN_TMEM ? 0 : 4 * AT_BQ * 128));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * AT_BQ;
  const int nblk = (a.T + AT_BKV - 1) / AT_BKV;
  constexpr uint32_t TMEM_COLS = P_IN_TMEM ? 256 : 512;
  constexpr uint32_t O_COL = P_IN_TMEM ? 128 : 256;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(&bars->q_full, 1);
    ptx::mbar_init(&bars->kv_full, 1);
    ptx::mbar_init(&bars->s_full, 1);
    ptx::mbar_init(&bars->p_full, 128);
    ptx::mbar_init(&bars->o_full, 1);
    ptx::mbar_init(&bars->o_empty, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 4) {
    // ------------------------------------------------------------- control thread: TMA + MMA issue
    if (lane == 0) {
      ptx::mbar_expect_tx(&bars->q_full, (uint32_t)q_bytes);
      for (int ka = 0; ka < nka; ++ka)
        ptx::tma_load_3d(sQ + ka * AT_BQ * 128, &tmQ, &bars->q_full, ka * 64, q0, bh);
      for (int blk = 0; blk < nblk; ++blk) {
        const uint32_t ph = blk & 1;
        const int kv0 = blk * AT_BKV;
        const int kvlen = min(AT_BKV, a.T - kv0);
        const int n_s = ((kvlen + 15) / 16) * 16;  // MMA N for S / K extent for P V
        if (blk > 0) ptx::mbar_wait(&bars->o_full, ph ^ 1);  // previous P V done: K / V tiles reusable
        ptx::mbar_expect_tx(&bars->kv_full, (uint32_t)(k_bytes + v_bytes));
        for (int ka = 0; ka < nka; ++ka)
          ptx::tma_load_3d(sK + ka * AT_BKV * 128, &tmK, &bars->kv_full, ka * 64, kv0, bh);
        for (int kk = 0; kk < 4; ++kk)
          ptx::tma_load_3d(sV + kk * v_atom, &tmV, &bars->kv_full, kv0 + kk * 64, 0, bh);
        if (blk == 0) ptx::mbar_wait(&bars->q_full, 0);
        ptx::mbar_wait(&bars->kv_full, ph);
        if (blk > 0) ptx::mbar_wait(&bars->o_empty, ph ^ 1);  // softmax threads have drained the previous O block
        ptx::tc_fence_after();
        // S = Q K^T
        {
          const uint32_t idesc = ptx::make_idesc_bf16(AT_BQ, n_s);
          const int ksteps = a.hdp / 16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int atom = ks >> 2, kin = ks & 3;
            const uint64_t ad = ptx::make_desc_k_sw128(ptx::smem_u32(sQ + atom * AT_BQ * 128)) + (uint64_t)(kin * 2);
            const uint64_t bd = ptx::make_desc_k_sw128(ptx::smem_u32(sK + atom * AT_BKV * 128)) + (uint64_t)(kin * 2);
            ptx::umma_bf16_ss(tmem_base, ad, bd, idesc, (uint32_t)(ks != 0));
          }
          ptx::umma_commit(&bars->s_full);
        }
        // O_blk = P V
        ptx::mbar_wait(&bars->p_full, ph);
        ptx::tc_fence_after();
        {
          const uint32_t idesc = ptx::make_idesc_bf16(AT_BQ, a.hdv);
          const int ksteps = n_s / 16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int atom = ks >> 2, kin = ks & 3;
            const uint64_t bd = ptx::make_desc_k_sw128(ptx::smem_u32(sV + atom * v_atom)) + (uint64_t)(kin * 2);
            if (P_IN_TMEM) {
              umma_bf16_ts(tmem_base + O_COL, tmem_base + (uint32_t)(ks * 8), bd, idesc, (uint32_t)(ks != 0));
            } else {
              const uint64_t ad = ptx::make_desc_k_sw128(ptx::smem_u32(sP + atom * AT_BQ * 128)) + (uint64_t)(kin * 2);
              ptx::umma_bf16_ss(tmem_base + O_COL, ad, bd, idesc, (uint32_t)(ks != 0));
            }
          }
          ptx::umma_commit(&bars->o_full);
        }
      }
    }
  } else {
    // ------------------------------------------------------------- softmax / output threads (thread == query row)
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sl2 = rsqrtf((float)a.hd) * 1.4426950408889634f;  // scale * log2(e)
    float m_run = -INFINITY, l_run = 0.f;
    constexpr bool multi = MULTI;
    float oacc[MULTI ? 128 : 1];
#pragma unroll
    for (int i = 0; i < (MULTI ? 128 : 1); ++i) oacc[i] = 0.f;
    for (int blk = 0; blk < nblk; ++blk) {
      const uint32_t ph = blk & 1;
      const int kvlen = min(AT_BKV, a.T - blk * AT_BKV);
      const int n_s = ((kvlen + 15) / 16) * 16;
      ptx::mbar_wait(&bars->s_full, ph);
      ptx::tc_fence_after();
      // pass 1: row maximum
      float mx = -INFINITY;
      for (int c0 = 0; c0 < n_s; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld_x16(trow + c0, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < kvlen) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * sl2);  // 0 on the first block (m_run = -inf)
      const float moff = m_new * sl2;
      // pass 2: p = exp2(s*scale*log2e - m), row sum, bf16 P
      float rs = 0.f;
      for (int c0 = 0; c0 < n_s; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld_x16(trow + c0, r);
        ptx::tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float p0 = (c0 + j < kvlen) ? exp2f(fmaf(__uint_as_float(r[j]), sl2, -moff)) : 0.f;
          float p1 = (c0 + j + 1 < kvlen) ? exp2f(fmaf(__uint_as_float(r[j + 1]), sl2, -moff)) : 0.f;
          __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
          // the row sum uses the rounded probabilities so numerator and denominator agree
          rs += __low2float(b2) + __high2float(b2);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&b2);
        }
        if (P_IN_TMEM) {
          // 16 keys -> 8 packed columns; written 16 keys behind the read front so S is never clobbered early
          // (tcgen05.st x16 needs 16 registers: stage two chunks)
          // simple form: write 8 columns with an x8 store
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                           trow + (uint32_t)(c0 >> 1)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                       : "memory");
        } else {
          // K-major SW128 tile: atom = 64 keys, row = 128 B, 16-byte chunk index XOR (row & 7)
          const int atom = c0 >> 6, ch = (c0 & 63) >> 3;
          uint8_t* base = sP + atom * AT_BQ * 128 + row * 128;
          *reinterpret_cast<uint4*>(base + (((ch) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(base + (((ch + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      if (P_IN_TMEM) tmem_st_wait();
      else ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars->p_full);
      l_run = l_run * alpha + rs;
      m_run = m_new;
      // O
      ptx::mbar_wait(&bars->o_full, ph);
      ptx::tc_fence_after();
      if (multi) {
        for (int c0 = 0; c0 < a.hdv; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld_x16(trow + O_COL + c0, r);
          ptx::tmem_ld_wait();
          // oacc is only ever indexed with compile-time constants (fully unrolled, guarded by c0)
#pragma unroll
          for (int cc = 0; cc < (MULTI ? 8 : 0); ++cc) {
            if (c0 == cc * 16) {
#pragma unroll
              for (int j = 0; j < 16; ++j) oacc[cc * 16 + j] = oacc[cc * 16 + j] * alpha + __uint_as_float(r[j]);
            }
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars->o_empty);
      }
    }
    // ---- write O / l   (token-major [B*T, H*hd], head h occupies columns h*hd..)
    const int b = bh / a.H, h = bh % a.H;
    const float inv = 1.0f / l_run;
    const bool valid = (q0 + row) < a.T;
    bf16* orow = a.out + ((size_t)b * a.T + q0 + row) * ((size_t)a.H * a.hd) + (size_t)h * a.hd;
    for (int c0 = 0; c0 < a.hdv; c0 += 16) {
      uint32_t r[16];
      if (!multi) {
        ptx::tmem_ld_x16(trow + O_COL + c0, r);
        ptx::tmem_ld_wait();
      }
      float v[16];
      if (MULTI) {
#pragma unroll
        for (int cc = 0; cc < (MULTI ? 8 : 0); ++cc) {
          if (c0 == cc * 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = oacc[cc * 16 + j] * inv;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * inv;
      }
      if (valid) {
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8) {
          if (c0 + j8 * 8 < a.hd) {  // hd % 8 == 0
            uint4 pk;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j8 * 8 + 0], v[j8 * 8 + 1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j8 * 8 + 2], v[j8 * 8 + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j8 * 8 + 4], v[j8 * 8 + 5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j8 * 8 + 6], v[j8 * 8 + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&t0);
            pk.y = *reinterpret_cast<uint32_t*>(&t1);
            pk.z = *reinterpret_cast<uint32_t*>(&t2);
            pk.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(orow + c0 + j8 * 8) = pk;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static bool g_attn_p_in_tmem = true;
void set_attention_variant(int p_in_tmem) { g_attn_p_in_tmem = p_in_tmem != 0; }

void launch_attention_tc(const TcAttnArgs& a, cudaStream_t s) {
  VV_CHECK(a.hd % 8 == 0 && a.hdp % 16 == 0 && a.hdv % 16 == 0 && a.hdp >= a.hd && a.hdv >= a.hd,
           "attention: head_dim must be a multiple of 8; padded dims multiples of 16");
  VV_CHECK(a.hdp <= 128 && a.hdv <= 128, "attention: head_dim > 128 not supported");
  VV_CHECK(a.T % 8 == 0, "attention: token count must be a multiple of 8");
  VV_CHECK((a.H * a.hd) % 8 == 0, "attention: hidden size must be a multiple of 8");
  const int nka = (a.hdp + 63) / 64;
  CUtensorMap tmQ = make_tmap_3d_bf16(a.q, (uint64_t)a.hdp, (uint64_t)a.T, (uint64_t)a.B * a.H, (uint64_t)a.hdp * 2,
                                      (uint64_t)a.T * a.hdp * 2, 64, AT_BQ, 1);
  CUtensorMap tmK = make_tmap_3d_bf16(a.k, (uint64_t)a.hdp, (uint64_t)a.T, (uint64_t)a.B * a.H, (uint64_t)a.hdp * 2,
                                      (uint64_t)a.T * a.hdp * 2, 64, AT_BKV, 1);
  CUtensorMap tmV = make_tmap_3d_bf16(a.vt, (uint64_t)a.T, (uint64_t)a.hdv, (uint64_t)a.B * a.H, (uint64_t)a.T * 2,
                                      (uint64_t)a.hdv * a.T * 2, 64, (uint32_t)a.hdv, 1);
  const bool ptm = g_attn_p_in_tmem;
  size_t smem = (size_t)nka * AT_BQ * 128 + (size_t)nka * AT_BKV * 128 + (size_t)4 * a.hdv * 128 +
                (ptm ? 0 : 4 * AT_BQ * 128) + sizeof(AtBars) + 1024;
  VV_CHECK(smem <= 227 * 1024, "attention: shared memory budget exceeded");
  dim3 grid(cdiv(a.T, AT_BQ), a.B * a.H);
  const bool multi = a.T > AT_BKV;
  auto go = [&](auto kern) {
    VV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<grid, AT_THREADS, smem, s>>>(tmQ, tmK, tmV, a);
  };
  if (ptm && !multi) go(k_attention_tc<true, false>);
  else if (ptm && multi) go(k_attention_tc<true, true>);
  else if (!ptm && !multi) go(k_attention_tc<false, false>);
  else go(k_attention_tc<false, true>);
  VV_LAUNCH_CHECK();
}

}  // namespace vv
