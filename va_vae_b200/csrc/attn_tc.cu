// tcgen05 flash attention for sm_100a:  O = softmax(Q K^T / sqrt(hd)) V   (lightningdit.py:77-88)
//
// Persistent kernel, two CTAs per SM; a work item is one (batch*head, 128-query tile) and keys are processed in
// blocks of 256 with an online softmax.  Per block ("step"):
//   S = Q K^T   : tcgen05.mma  M=128, N=keys (<=256), K=hd padded to 16         -> TMEM columns [0,256)
//   softmax     : 128 threads, thread == query row; S read with tcgen05.ld; p = ex2.approx.ftz.f32 (one MUFU op
//                 per probability), TRUNCATED to bf16 pairs by one PRMT (F2FP would share the XU pipe with MUFU; the
//                 truncation bias cancels in the normalisation, see ex2_trunc_bf16x2) and written straight back into TMEM
//   P           : bf16, two per 32-bit TMEM column, aliasing dead S columns; TMEM A-operand of the second MMA
//   O = P V     : tcgen05.mma  M=128, N=hdv, K=keys.  For a full 256-key block the MMA over the first 192 keys is
//                 issued while the last 64 keys are still being exponentiated (barrier p_part; O then sits at
//                 columns [96, 96+hdv) and the last two P chunks stay in place at columns 192 / 224)
//   row sums    : V^T carries a row of ones at index hd, so column hd of O *is* sum_j P_ij (exactly the bf16
//                 probabilities the MMA consumed) - no summation instructions in the softmax loop.
//   output      : normalised rows are packed to bf16 into a [128][hd] tile in the idle V^T buffer and written by ONE
//                 TMA store per item (cp.async.bulk.tensor, full lines) - row-strided 16-byte st.global from the
//                 thread == row domain cost 12 % of the kernel in LSU time.
// TMA loads are software-pipelined across steps with single buffers: Q/K of step s+1 are requested as soon as
// the S MMA of step s retires and S of step s+1 is issued as soon as the softmax threads have read O of step s;
// V of step s+1 is requested once the P V MMA of step s has retired (and, on an item's last block, once the output
// store has read the staging tile), so every load hides behind a softmax.  The second CTA on the SM fills the
// tensor / MUFU pipes meanwhile.
//
// Layouts (all K-major TMA tiles, 128-byte swizzle): q,k (B,H,T,hdp); v transposed (B,H,hdv,T), written by the
// QKV GEMM epilogue.  A 16-column tail of hdp (head_dim 72 -> 80) is loaded as a compact 32-byte-swizzled tile.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <atomic>
#include <cstdlib>

namespace vv {

bool device_marked(const std::atomic<uint64_t>& mask);
void mark_device(std::atomic<uint64_t>& mask);

constexpr int AT_BQ = 128;   // queries per work item
constexpr int AT_BKV = 256;  // keys per step
constexpr int AT_THREADS = 160;

struct AtBars {
  uint64_t qk_full, v_full, s_full, p_full, p_part, o_full, o_empty, o_staged, qk_empty, v_empty;
  uint32_t tmem_base;
};

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
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.wait::ld that also names the destination registers as in/out operands, so the compiler cannot schedule
// arithmetic on registers of an in-flight tcgen05.ld above the wait (needed once loads are software-pipelined).
__device__ __forceinline__ void tmem_ld_wait_dep32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// K-major operand tile whose rows are 32 bytes (16 bf16) with the 32-byte swizzle (CU_TENSOR_MAP_SWIZZLE_32B):
// 8-row groups are 256 bytes apart; layout type 6.
__device__ __forceinline__ uint64_t make_desc_k_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
// 2^x in fp32 on both values, packed to bf16x2 afterwards.  (ex2.approx.ftz.bf16x2 does not halve the MUFU work on
// sm_100a - it compiles to two scalar MUFU.EX2.BF16 plus a PRMT - so the fp32 form is one instruction per pair
// cheaper: measured +2.8 % on the isolated kernel.)
__device__ __forceinline__ uint32_t ex2_pack_bf16x2(float x0, float x1) {
  float e0, e1;
  uint32_t out;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(out) : "f"(e1), "f"(e0));
  return out;
}

// Same, but the pair is packed by TRUNCATION (one PRMT on the ALU pipe instead of F2FP, which shares the 16-lane XU
// pipe with MUFU).  The row sum is taken from the very probabilities the MMA consumes (ones row of V^T), so the mean
// relative bias of truncation (-2^-8) cancels in O = sum(p v) / sum(p); what is left has the variance of round-to-nearest.
__device__ __forceinline__ uint32_t ex2_trunc_bf16x2(float x0, float x1) {
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
  return __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

struct AtGeom {
  int n_full;     // number of full 64-column atoms of Q/K
  int tail16;     // 1: a compact 16-column tail tile follows the full atoms
  int q_bytes, k_bytes, v_atom, v_bytes;
};
__host__ __device__ inline AtGeom at_geom(int hdp, int hdv) {
  AtGeom g;
  g.tail16 = (hdp % 64 == 16) ? 1 : 0;
  g.n_full = g.tail16 ? hdp / 64 : (hdp + 63) / 64;
  g.q_bytes = g.n_full * AT_BQ * 128 + g.tail16 * AT_BQ * 32;
  g.k_bytes = g.n_full * AT_BKV * 128 + g.tail16 * AT_BKV * 32;
  g.v_atom = hdv * 128;
  g.v_bytes = 4 * g.v_atom;
  return g;
}

// OACC: fp32 output columns a thread accumulates in registers across key blocks (0: single-block items, O goes
// straight from TMEM to global; 80: head_dim <= 79 - every registry model; 128: up to head_dim 112)
template <int OACC>
__global__ void __launch_bounds__(AT_THREADS, 2)
k_attention_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQt,
               const __grid_constant__ CUtensorMap tmKt, const __grid_constant__ CUtensorMap tmO,
               const TcAttnArgs a) {
  constexpr bool MULTI = OACC > 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const AtGeom g = at_geom(a.hdp, a.hdv);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ((g.q_bytes + 1023) & ~1023);
  uint8_t* sV = sK + ((g.k_bytes + 1023) & ~1023);
  AtBars* bars = reinterpret_cast<AtBars*>(sV + g.v_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (a.T + AT_BQ - 1) / AT_BQ;
  const int nblk = (a.T + AT_BKV - 1) / AT_BKV;
  const int n_items = a.B * a.H * nq;
  constexpr uint32_t TMEM_COLS = 256;
  // O accumulator columns.  Generic blocks: P compact in [0,128), O at 128 (aliasing dead S columns).  Full 256-key
  // blocks in early-PV mode: P of key chunks 0..5 compact in [0,96), O at 96, P of chunks 6 / 7 in place at 192 / 224,
  // so the P V MMA over the first 192 keys runs while the last two chunks are still being exponentiated.
  constexpr uint32_t O_COL_GENERIC = 128, O_COL_EARLY = 96;
  const bool early_ok = !(a.experiment & 8) && a.hdv <= 96;
  const bool trunc_pack = !(a.experiment & 4);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::prefetch_tmap(&tmO);
    ptx::mbar_init(&bars->qk_full, 1);
    ptx::mbar_init(&bars->v_full, 1);
    ptx::mbar_init(&bars->s_full, 1);
    ptx::mbar_init(&bars->p_full, 128);
    ptx::mbar_init(&bars->p_part, 128);
    ptx::mbar_init(&bars->o_full, 1);
    ptx::mbar_init(&bars->o_empty, 128);
    ptx::mbar_init(&bars->o_staged, 128);
    ptx::mbar_init(&bars->qk_empty, 1);
    ptx::mbar_init(&bars->v_empty, 1);
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();   // q / k / v^T of the preceding QKV GEMM are read from here on

  if (warp == 4) {
    // ------------------------------------------------------------- control thread: TMA + MMA issue
    if (lane == 0) {
      // step = (item, blk) flattened; loads for a step: K (+Q on the first block of an item), and V
      auto load_qk = [&](int item, int blk) {
        const int it_ = a.reverse ? n_items - 1 - item : item;
        const int bh = it_ / nq, q0 = (it_ % nq) * AT_BQ, kv0 = blk * AT_BKV;
        ptx::mbar_expect_tx(&bars->qk_full, (uint32_t)(g.k_bytes + (blk == 0 ? g.q_bytes : 0)));
        for (int ka = 0; ka < g.n_full; ++ka) {
          ptx::tma_load_3d(sK + ka * AT_BKV * 128, &tmK, &bars->qk_full, ka * 64, kv0, bh);
          if (blk == 0) ptx::tma_load_3d(sQ + ka * AT_BQ * 128, &tmQ, &bars->qk_full, ka * 64, q0, bh);
        }
        if (g.tail16) {
          ptx::tma_load_3d(sK + g.n_full * AT_BKV * 128, &tmKt, &bars->qk_full, g.n_full * 64, kv0, bh);
          if (blk == 0) ptx::tma_load_3d(sQ + g.n_full * AT_BQ * 128, &tmQt, &bars->qk_full, g.n_full * 64, q0, bh);
        }
      };
      auto load_v = [&](int item, int blk) {
        const int it_ = a.reverse ? n_items - 1 - item : item;
        const int bh = it_ / nq, kv0 = blk * AT_BKV;
        ptx::mbar_expect_tx(&bars->v_full, (uint32_t)g.v_bytes);
        for (int kk = 0; kk < 4; ++kk) ptx::tma_load_3d(sV + kk * g.v_atom, &tmV, &bars->v_full, kv0 + kk * 64, 0, bh);
      };
      // S = Q K^T of one step (keys [blk*256, ...) of an item)
      auto issue_s = [&](int blk) {
        const int kvlen = min(AT_BKV, a.T - blk * AT_BKV);
        const uint32_t idesc = ptx::make_idesc_bf16(AT_BQ, ((kvlen + 15) / 16) * 16);
        int ks = 0;
        for (int ka = 0; ka < g.n_full; ++ka) {
          const int kmax = min(4, (a.hdp - ka * 64) / 16);
          for (int kin = 0; kin < kmax; ++kin, ++ks) {
            const uint64_t ad = ptx::make_desc_k_sw128(ptx::smem_u32(sQ + ka * AT_BQ * 128)) + (uint64_t)(kin * 2);
            const uint64_t bd = ptx::make_desc_k_sw128(ptx::smem_u32(sK + ka * AT_BKV * 128)) + (uint64_t)(kin * 2);
            ptx::umma_bf16_ss(tmem_base, ad, bd, idesc, (uint32_t)(ks != 0));
          }
        }
        if (g.tail16) {
          const uint64_t ad = make_desc_k_sw32(ptx::smem_u32(sQ + g.n_full * AT_BQ * 128));
          const uint64_t bd = make_desc_k_sw32(ptx::smem_u32(sK + g.n_full * AT_BKV * 128));
          ptx::umma_bf16_ss(tmem_base, ad, bd, idesc, (uint32_t)(ks != 0));
        }
        ptx::umma_commit(&bars->s_full);
        ptx::umma_commit(&bars->qk_empty);
      };
      uint32_t step = 0, items_done = 0;
      if ((int)blockIdx.x < n_items) {
        load_qk(blockIdx.x, 0);
        load_v(blockIdx.x, 0);
        ptx::mbar_wait(&bars->qk_full, 0);
        ptx::tc_fence_after();
        issue_s(0);
      }
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (int blk = 0; blk < nblk; ++blk, ++step) {
          // the S MMA of this step is already in flight (issued at the end of the previous step)
          const uint32_t ph = step & 1;
          const int kvlen = min(AT_BKV, a.T - blk * AT_BKV);
          const int n_s = ((kvlen + 15) / 16) * 16;  // K extent of P V
          int nitem = item, nblkk = blk + 1;          // next step (for prefetch)
          if (nblkk == nblk) { nitem = item + gridDim.x; nblkk = 0; }
          const bool has_next = nitem < n_items;

          // Q/K tiles are free once S retires: request the next step's Q/K now (hidden behind this softmax)
          ptx::mbar_wait(&bars->qk_empty, ph);
          if (has_next) load_qk(nitem, nblkk);

          // O_blk = P V
          ptx::mbar_wait(&bars->v_full, ph);
          {
            const uint32_t idesc = ptx::make_idesc_bf16(AT_BQ, a.hdv);
            const int ksteps = n_s / 16;
            const bool early = early_ok && kvlen == AT_BKV;
            const uint32_t o_col = early ? O_COL_EARLY : O_COL_GENERIC;
            auto pv = [&](int ks, uint32_t pcol) {
              const int atom = ks >> 2, kin = ks & 3;
              const uint64_t bd = ptx::make_desc_k_sw128(ptx::smem_u32(sV + atom * g.v_atom)) + (uint64_t)(kin * 2);
              umma_bf16_ts(tmem_base + o_col, tmem_base + pcol, bd, idesc, (uint32_t)(ks != 0));
            };
            if (early) {
              ptx::mbar_wait(&bars->p_part, ph);   // probabilities of keys [0,192) are in TMEM
              ptx::tc_fence_after();
              for (int ks = 0; ks < 12; ++ks) pv(ks, (uint32_t)(ks * 8));
              ptx::mbar_wait(&bars->p_full, ph);
              ptx::tc_fence_after();
              pv(12, 192); pv(13, 200); pv(14, 224); pv(15, 232);
            } else {
              if (early_ok) ptx::mbar_wait(&bars->p_part, ph);   // keeps the barrier's phase in step
              ptx::mbar_wait(&bars->p_full, ph);
              ptx::tc_fence_after();
              for (int ks = 0; ks < ksteps; ++ks) pv(ks, (uint32_t)(ks * 8));
            }
            ptx::umma_commit(&bars->o_full);
            ptx::umma_commit(&bars->v_empty);
          }
          ptx::mbar_wait(&bars->v_empty, ph);
          // next step's S as soon as its Q/K have landed and the softmax threads have read this step's O
          if (has_next) {
            ptx::mbar_wait(&bars->qk_full, ph ^ 1);
            ptx::mbar_wait(&bars->o_empty, ph);
            ptx::tc_fence_after();
            issue_s(nblkk);
          }
          if (blk == nblk - 1) {
            // the item's output rows, normalised and packed to bf16, are staged in the (now idle) V^T buffer by the
            // softmax threads: one TMA store writes them as full lines; V^T of the next step is requested once the
            // store has read the buffer (it is not needed before that step's P V, a whole softmax away)
            const int it_ = a.reverse ? n_items - 1 - item : item;
            const int bh = it_ / nq, q0 = (it_ % nq) * AT_BQ;
            ptx::mbar_wait(&bars->o_staged, items_done & 1);
            ++items_done;
            ptx::tma_store_4d(&tmO, sV, 0, bh % a.H, q0, bh / a.H);
            ptx::tma_store_commit();
            ptx::tma_store_wait_read0();
          }
          if (has_next) load_v(nitem, nblkk);
        }
      }
      ptx::tma_store_wait_all();
    }
  } else {
    // ------------------------------------------------------------- softmax / output threads (thread == query row)
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sl2 = rsqrtf((float)a.hd) * 1.4426950408889634f;  // scale * log2(e)
    const int ncol_o = a.hd + 1;                                   // head_dim columns + the row-sum column
    // output staging: row-major [128][hd] bf16 in the V^T buffer (idle between the last P V of an item and the next V^T
    // load); 16-byte chunks are XOR-swizzled like the tensor map when a row is exactly 128 bytes (conflict-free stores;
    // a 144-byte row pitch is conflict-free as it is)
    const uint32_t stage_row = ptx::smem_u32(sV) + (uint32_t)row * (uint32_t)(a.hd * 2);
    const uint32_t stage_xor = (a.hd == 64) ? (uint32_t)(row & 7) : 0u;
    auto stage16 = [&](int chunk, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (((uint32_t)chunk ^ stage_xor) << 4)),
                   "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                   : "memory");
    };
    uint32_t step = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int it_ = a.reverse ? n_items - 1 - item : item;
      const int bh = it_ / nq, q0 = (it_ % nq) * AT_BQ;
      float m_run = -INFINITY;
      float oacc[MULTI ? OACC : 1];
#pragma unroll
      for (int i = 0; i < (MULTI ? OACC : 1); ++i) oacc[i] = 0.f;

      for (int blk = 0; blk < nblk; ++blk, ++step) {
        const uint32_t ph = step & 1;
        const int kvlen = min(AT_BKV, a.T - blk * AT_BKV);
        const int n_s = ((kvlen + 15) / 16) * 16;
        const bool ragged = (kvlen & 15) != 0;  // only the last 16-column group can hold invalid keys
        ptx::mbar_wait(&bars->s_full, ph);
        ptx::tc_fence_after();
        float m_new, alpha, moff;
        const bool full_blk = n_s == AT_BKV && !ragged;
        const uint32_t o_col = (early_ok && full_blk) ? O_COL_EARLY : O_COL_GENERIC;
        if (full_blk) {
          // ---- full 256-key block: fully unrolled, TMEM loads software-pipelined (load chunk i+1 while reducing /
          //      exponentiating chunk i; two 32-register buffers)
          uint32_t ra[32], rb[32];
          ptx::tmem_ld_x32(trow, ra);
          tmem_ld_wait_dep32(ra);
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (a.experiment == 1) m4[0] = 0.f;   // EXPERIMENT (timing only): no row-max pass, offset 0
          if (a.experiment != 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint32_t* cur = (i & 1) ? rb : ra;
            uint32_t* nxt = (i & 1) ? ra : rb;
            ptx::tmem_ld_x32(trow + (uint32_t)(((i + 1) & 7) * 32), nxt);  // i == 7 prefetches chunk 0 for pass 2
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              m4[0] = fmaxf(m4[0], __uint_as_float(cur[j]));
              m4[1] = fmaxf(m4[1], __uint_as_float(cur[j + 1]));
              m4[2] = fmaxf(m4[2], __uint_as_float(cur[j + 2]));
              m4[3] = fmaxf(m4[3], __uint_as_float(cur[j + 3]));
            }
            tmem_ld_wait_dep32(nxt);
          }
          }
          const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          m_new = fmaxf(m_run, mx);
          alpha = exp2f((m_run - m_new) * sl2);  // 0 on the first block (m_run = -inf)
          moff = m_new * sl2;
          // pass 2: chunk 0 is already in `ra`.  P (16 packed columns per 32 keys) is written behind the read
          // front: chunk i+1 (columns 32(i+1)..) has been read into registers before P of chunk i lands on
          // columns 16i..16i+15.
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint32_t* cur = (i & 1) ? rb : ra;
            uint32_t* nxt = (i & 1) ? ra : rb;
            if (i < 7) ptx::tmem_ld_x32(trow + (uint32_t)((i + 1) * 32), nxt);
            uint32_t pk[16];
            if (trunc_pack) {
#pragma unroll
              for (int j = 0; j < 32; j += 2)
                pk[j >> 1] = ex2_trunc_bf16x2(fmaf(__uint_as_float(cur[j]), sl2, -moff),
                                              fmaf(__uint_as_float(cur[j + 1]), sl2, -moff));
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2)
                pk[j >> 1] = ex2_pack_bf16x2(fmaf(__uint_as_float(cur[j]), sl2, -moff),
                                             fmaf(__uint_as_float(cur[j + 1]), sl2, -moff));
            }
            if (i < 7) tmem_ld_wait_dep32(nxt);   // the next chunk is in registers before P overwrites its S columns
            // early-PV mode: chunks 6 / 7 stay in place (columns 192 / 224) - O occupies [96, 96 + hdv)
            tmem_st_x16(trow + (uint32_t)((early_ok && i >= 6) ? i * 32 : i * 16), pk);
            if (early_ok && i == 5) {
              tmem_st_wait();
              ptx::tc_fence_before();
              ptx::mbar_arrive(&bars->p_part);
            }
          }
        } else {
        // ---- generic block (fewer than 256 keys and / or a ragged tail)
        // pass 1: row maximum
        float mx = -INFINITY;
        for (int c0 = 0; c0 < n_s; c0 += 32) {
          uint32_t r[32];
          if (c0 + 32 <= n_s) {
            ptx::tmem_ld_x32(trow + c0, r);
          } else {
            ptx::tmem_ld_x16(trow + c0, r);
#pragma unroll
            for (int j = 16; j < 32; ++j) r[j] = 0xff800000u;  // -inf
          }
          ptx::tmem_ld_wait();
          if (ragged && c0 + 32 >= kvlen) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j >= kvlen) r[j] = 0xff800000u;
          }
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m4[0] = fmaxf(m4[0], __uint_as_float(r[j]));
            m4[1] = fmaxf(m4[1], __uint_as_float(r[j + 1]));
            m4[2] = fmaxf(m4[2], __uint_as_float(r[j + 2]));
            m4[3] = fmaxf(m4[3], __uint_as_float(r[j + 3]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        }
        m_new = fmaxf(m_run, mx);
        alpha = exp2f((m_run - m_new) * sl2);  // 0 on the first block (m_run = -inf)
        moff = m_new * sl2;
        // pass 2: p = 2^(s*scale*log2e - m) as packed bf16 pairs, written back to TMEM (16 keys -> 8 columns;
        // the write front at c0/2 never passes the read front at c0, so S is not clobbered before it is read)
        for (int c0 = 0; c0 < n_s; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld_x16(trow + c0, r);
          ptx::tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float x0 = fmaf(__uint_as_float(r[j]), sl2, -moff);
            float x1 = fmaf(__uint_as_float(r[j + 1]), sl2, -moff);
            if (ragged) {
              if (c0 + j >= kvlen) x0 = -INFINITY;
              if (c0 + j + 1 >= kvlen) x1 = -INFINITY;
            }
            pk[j >> 1] = ex2_pack_bf16x2(x0, x1);
          }
          tmem_st_x8(trow + (uint32_t)(c0 >> 1), pk);
        }
        }
        m_run = m_new;
        tmem_st_wait();
        ptx::tc_fence_before();
        if (early_ok && !full_blk) ptx::mbar_arrive(&bars->p_part);
        ptx::mbar_arrive(&bars->p_full);

        // O block (and, in column hd, the row sum of the probabilities the MMA consumed)
        ptx::mbar_wait(&bars->o_full, ph);
        ptx::tc_fence_after();
        const bool last = blk == nblk - 1;
        if (!MULTI) {
          // single block: normalise and store straight from TMEM.  All O columns (and the row-sum column hd) are
          // requested first and waited for once; TMEM is released before the global stores.
          const int ncol = ((a.hd >> 4) + 1) << 4;   // 16-column groups up to and including the one holding column hd
          auto store8 = [&](const uint32_t* r8, int c, float inv) {   // 8 consecutive columns -> one 16-byte chunk
            if (c >= a.hd) return;
            stage16(c >> 3, pack_bf16x2(__uint_as_float(r8[0]) * inv, __uint_as_float(r8[1]) * inv),
                    pack_bf16x2(__uint_as_float(r8[2]) * inv, __uint_as_float(r8[3]) * inv),
                    pack_bf16x2(__uint_as_float(r8[4]) * inv, __uint_as_float(r8[5]) * inv),
                    pack_bf16x2(__uint_as_float(r8[6]) * inv, __uint_as_float(r8[7]) * inv));
          };
          if (ncol == 80) {
            // head_dim 64 / 72 (every registry model): 80 columns in three loads, one wait, TMEM released early
            uint32_t o0[32], o1[32], o2[16];
            ptx::tmem_ld_x32(trow + o_col, o0);
            ptx::tmem_ld_x32(trow + o_col + 32, o1);
            ptx::tmem_ld_x16(trow + o_col + 64, o2);
            tmem_ld_wait_dep32(o0);
            tmem_ld_wait_dep32(o1);
            tmem_ld_wait_dep16(o2);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars->o_empty);   // O / P columns may be overwritten by the next S MMA from here on
            const float inv = 1.0f / __uint_as_float(a.hd == 64 ? o2[0] : o2[8]);   // column hd holds the row sum
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              store8(o0 + j * 8, j * 8, inv);
              store8(o1 + j * 8, 32 + j * 8, inv);
            }
            store8(o2, 64, inv);
          } else {
            uint32_t rl[16];
            ptx::tmem_ld_x16(trow + o_col + (uint32_t)((a.hd >> 4) << 4), rl);  // the 16-column group holding column hd
            tmem_ld_wait_dep16(rl);
            const float inv = 1.0f / __uint_as_float((a.hd & 15) ? rl[8] : rl[0]);
            for (int c0 = 0; c0 < a.hd; c0 += 16) {
              uint32_t r[16];
              ptx::tmem_ld_x16(trow + o_col + c0, r);
              tmem_ld_wait_dep16(r);
              store8(r, c0, inv);
              store8(r + 8, c0 + 8, inv);
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars->o_empty);
          }
        } else {
          // online softmax: rescale the running O (and running row sum in column hd) and add this block
          if (OACC == 80) {
            // head_dim 64 / 72: all 80 columns requested at once, one wait, TMEM released before the arithmetic
            uint32_t o0[32], o1[32], o2[16];
            ptx::tmem_ld_x32(trow + o_col, o0);
            ptx::tmem_ld_x32(trow + o_col + 32, o1);
            ptx::tmem_ld_x16(trow + o_col + 64, o2);
            tmem_ld_wait_dep32(o0);
            tmem_ld_wait_dep32(o1);
            tmem_ld_wait_dep16(o2);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&bars->o_empty);
#pragma unroll
            for (int j = 0; j < (OACC == 80 ? 32 : 0); ++j) {
              oacc[j] = oacc[j] * alpha + __uint_as_float(o0[j]);
              oacc[(OACC == 80 ? 32 : 0) + j] = oacc[(OACC == 80 ? 32 : 0) + j] * alpha + __uint_as_float(o1[j]);
            }
#pragma unroll
            for (int j = 0; j < (OACC == 80 ? 16 : 0); ++j)
              oacc[(OACC == 80 ? 64 : 0) + j] = oacc[(OACC == 80 ? 64 : 0) + j] * alpha + __uint_as_float(o2[j]);
          } else {
          for (int c0 = 0; c0 < ncol_o; c0 += 16) {
            uint32_t r[16];
            ptx::tmem_ld_x16(trow + o_col + c0, r);
            ptx::tmem_ld_wait();
            // oacc is only ever indexed with compile-time constants (fully unrolled, guarded by c0)
#pragma unroll
            for (int cc = 0; cc < OACC / 16; ++cc) {
              if (c0 == cc * 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) oacc[cc * 16 + j] = oacc[cc * 16 + j] * alpha + __uint_as_float(r[j]);
              }
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bars->o_empty);
          }
          if (last) {
            float l = 0.f;
#pragma unroll
            for (int i = 0; i < OACC; ++i)
              if (i == a.hd) l = oacc[i];
            const float inv = 1.0f / l;
#pragma unroll
            for (int c8 = 0; c8 < OACC / 8; ++c8) {
              if (c8 * 8 < a.hd)
                stage16(c8, pack_bf16x2(oacc[c8 * 8 + 0] * inv, oacc[c8 * 8 + 1] * inv),
                        pack_bf16x2(oacc[c8 * 8 + 2] * inv, oacc[c8 * 8 + 3] * inv),
                        pack_bf16x2(oacc[c8 * 8 + 4] * inv, oacc[c8 * 8 + 5] * inv),
                        pack_bf16x2(oacc[c8 * 8 + 6] * inv, oacc[c8 * 8 + 7] * inv));
            }
          }
        }
        if (last) {
          // the 128 output rows are staged: make them visible to the async proxy and hand them to the control thread
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&bars->o_staged);
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

// 3-D bf16 tensor map with a caller-chosen swizzle (32-byte swizzle for the compact 16-column tail tiles)
CUtensorMap make_tmap_3d_bf16_sw(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes,
                                 uint64_t pitch2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes);
CUtensorMap make_tmap_4d(const void* base, int elem_bytes, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4], int swizzle_bytes);

int attention_hdv(int hd) { return ((hd + 1 + 15) / 16) * 16; }

void launch_attention_tc(const TcAttnArgs& a, cudaStream_t s) {
  VV_CHECK(a.hd % 8 == 0 && a.hdp % 16 == 0 && a.hdv % 16 == 0 && a.hdp >= a.hd && a.hdv >= a.hd + 1,
           "attention: head_dim % 8; padded dims % 16; hdv must leave room for the row-sum column");
  VV_CHECK(a.hdp <= 128 && a.hdv <= 128, "attention: head_dim > 112 not supported");
  VV_CHECK(a.T % 8 == 0, "attention: token count must be a multiple of 8");
  VV_CHECK((a.H * a.hd) % 8 == 0, "attention: hidden size must be a multiple of 8");
  const AtGeom g = at_geom(a.hdp, a.hdv);
  const uint64_t BH = (uint64_t)a.B * a.H;
  CUtensorMap tmQ = make_tmap_3d_bf16(a.q, (uint64_t)a.hdp, (uint64_t)a.T, BH, (uint64_t)a.hdp * 2,
                                      (uint64_t)a.T * a.hdp * 2, 64, AT_BQ, 1);
  CUtensorMap tmK = make_tmap_3d_bf16(a.k, (uint64_t)a.hdp, (uint64_t)a.T, BH, (uint64_t)a.hdp * 2,
                                      (uint64_t)a.T * a.hdp * 2, 64, AT_BKV, 1);
  CUtensorMap tmV = make_tmap_3d_bf16(a.vt, (uint64_t)a.T, (uint64_t)a.hdv, BH, (uint64_t)a.T * 2,
                                      (uint64_t)a.hdv * a.T * 2, 64, (uint32_t)a.hdv, 1);
  CUtensorMap tmQt = tmQ, tmKt = tmK;
  if (g.tail16) {
    tmQt = make_tmap_3d_bf16_sw(a.q, (uint64_t)a.hdp, (uint64_t)a.T, BH, (uint64_t)a.hdp * 2, (uint64_t)a.T * a.hdp * 2,
                                16, AT_BQ, 1, 32);
    tmKt = make_tmap_3d_bf16_sw(a.k, (uint64_t)a.hdp, (uint64_t)a.T, BH, (uint64_t)a.hdp * 2, (uint64_t)a.T * a.hdp * 2,
                                16, AT_BKV, 1, 32);
  }
  // output rows of one (batch, head, 128-query tile): box over (hd, H, T, B), written by TMA from the staging tile
  const uint64_t odims[4] = {(uint64_t)a.hd, (uint64_t)a.H, (uint64_t)a.T, (uint64_t)a.B};
  const uint64_t ostr[3] = {(uint64_t)a.hd * 2, (uint64_t)a.H * a.hd * 2, (uint64_t)a.T * a.H * a.hd * 2};
  const uint32_t obox[4] = {(uint32_t)a.hd, 1u, (uint32_t)std::min(AT_BQ, a.T), 1u};
  CUtensorMap tmO = make_tmap_4d(a.out, 2, odims, ostr, obox, a.hd == 64 ? 128 : 0);
  const size_t smem = (size_t)((g.q_bytes + 1023) & ~1023) + ((g.k_bytes + 1023) & ~1023) + g.v_bytes + sizeof(AtBars) + 1024;
  VV_CHECK(smem <= 227 * 1024, "attention: shared memory budget exceeded");
  const int sms = num_sms();
  const int n_items = a.B * a.H * cdiv(a.T, AT_BQ);
  int per_sm = smem <= 113 * 1024 ? 2 : 1;
  {
    static int cap = -1;   // VAVAE_ATTN_PER_SM=1: developer experiment (do the two co-resident CTAs overlap?)
    if (cap < 0) { const char* e = std::getenv("VAVAE_ATTN_PER_SM"); cap = e ? std::atoi(e) : 0; }
    if (cap > 0) per_sm = std::min(per_sm, cap);
  }
  const int grid = std::min(n_items, sms * per_sm);
  const bool multi = a.T > AT_BKV;
  TcAttnArgs aa = a;
  {
    static int exp_mode = -1;
    if (exp_mode < 0) { const char* e = std::getenv("VAVAE_ATTN_EXPERIMENT"); exp_mode = e ? std::atoi(e) : 0; }
    aa.experiment = exp_mode;
  }
  static std::atomic<uint64_t> attr_devices{0};   // per device; not a stream operation, safe during graph capture
  if (!device_marked(attr_devices)) {
    VV_CUDA(cudaFuncSetAttribute(k_attention_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VV_CUDA(cudaFuncSetAttribute(k_attention_tc<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VV_CUDA(cudaFuncSetAttribute(k_attention_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    mark_device(attr_devices);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(AT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled(s)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (multi && a.hd + 1 <= 80) VV_CUDA(cudaLaunchKernelEx(&cfg, k_attention_tc<80>, tmQ, tmK, tmV, tmQt, tmKt, tmO, aa));
  else if (multi) VV_CUDA(cudaLaunchKernelEx(&cfg, k_attention_tc<128>, tmQ, tmK, tmV, tmQt, tmKt, tmO, aa));
  else VV_CUDA(cudaLaunchKernelEx(&cfg, k_attention_tc<0>, tmQ, tmK, tmV, tmQt, tmKt, tmO, aa));
  VV_LAUNCH_CHECK();
}

}  // namespace vv
