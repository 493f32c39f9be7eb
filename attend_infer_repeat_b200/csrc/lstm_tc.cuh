// The whole snt.LSTM recurrence of an unrolled AIR pass in ONE launch (sm_100a): gx = e @ W[:n_enc] + b once, then for
// t = 1..T: gates = gx + h_{t-1} @ W[n_enc:], (c, h_t) gate math -- mnist_model.py:35 [upstream snt.LSTM], cell.py:126-127.
//
// A 128-canvas row tile is owned by a CLUSTER of 4 CTAs.  CTA r computes hidden units [64r, 64r + 64): its weight slab is
// the 4 x 64 gate columns (i, j, f, o) of those units, so a step is one 128 x 256 x 256 tcgen05 GEMM per CTA (TS form, the
// A operand h_{t-1} resident in TMEM exactly as in chain_tc.cuh) followed by the gate math of 64 units x 128 rows in the
// epilogue warps -- i, j, f, o of a unit land in the same thread, c stays in registers for all T steps.  The four h slabs
// are exchanged through an L2-resident buffer in the chains' slice-major layout (every CTA needs all 256 units of h_t as
// the next step's operand) and a cluster-scope mbarrier (remote arrive through DSMEM); weights stream from L2 by TMA.
// 32 row tiles x 4 = 128 CTAs busy instead of the 32 a row-tile-only split would give.
#pragma once
#include "chain_tc.cuh"

namespace air {
namespace lstm {

using namespace air::chain;

constexpr int CLUSTER = 4;
constexpr int UPC = 64;               // hidden units per CTA
constexpr int NH = CLUSTER * UPC;     // this kernel is specialised for snt.LSTM(256) (mnist_model.py:35)

struct Params {
  CUtensorMap tm_x;       // prepared W[:n_enc]^T, rows permuted to [cta][gate][unit] (prep_weights_kernel lstm mode)
  CUtensorMap tm_h;       // prepared W[n_enc:]^T, same permutation
  const float* bias;      // lstm.b permuted the same way (bias arena), [4 * NH]
  const float* e;         // [B, n_enc] fp32: input-encoder output
  const __half* e_hl;     // the same as row-major hl planes [2][rows_alloc][e_ld] (preferred: 16-byte loads, no staging), or null
  size_t e_plane;
  int e_ld;
  int e_nsl;              // > 0: e_hl is slice-major tiled, [row tile][e_nsl][128 rows][16] per plane (coalesced 1 KB runs)
  int hs_last_only;       // inference: only the last step's fp32 h rows are needed (final_h); the heads read the hl copy
  int n_enc;
  const float* h_init;    // [B, NH] fp32 (row pitch h_init_ld; 0 = one [NH] vector broadcast to every canvas)
  int h_init_ld;
  // inference, input Encoder with >= 2 layers: the cluster computes the LAST encoder layer itself (e = ELU(e1 @ W2 + b2), every
  // CTA all n_enc columns of its 128 rows) as an extra GEMM group in front of gx; `e_hl` then holds e1 (K = n_e1) and the
  // activation never leaves tensor memory.  Saves a launch and the round trip of e through L2.
  int fuse_e2;
  CUtensorMap tm_e2;      // prepared W2^T [2][N_alloc][Kpad] (hi rows at 0, lo rows at e2_lo_row), box 256 x 64
  int e2_lo_row;
  int n_e1;               // K of the fused layer
  const float* bias_e2;   // [n_enc]
  const float* hw0;       // broadcast initial state only: h0 @ W[n_enc:] as fp32, permuted like `bias` (lstm_h0w_kernel), or null.
                          // Step 1 then needs no recurrent GEMM -- its gates are gx + hw0 -- and the gx epilogue IS step 1.
  const float* c_in;      // initial cell state, row pitch c_in_ld (0 = broadcast vector); the final state goes to `c`
  int c_in_ld;
  float* c;               // [B, NH] fp32: initial cell state in, final cell state out
  float* hs;              // [T, B, NH] fp32 out
  HlOut hs_hlt;           // slice-major tiled hl copy of hs (operand of the heads chain), row = t * B + b
  float* gx_scr;          // [tiles][4][4][16][128][4] fp32 scratch: gx in the layout of the thread that re-reads it
  __half* hx;             // [2][2 planes][tiles][16 slices][128][16] fp16: h exchange buffer, double buffered by step parity
  size_t hx_plane;        // halves per plane = tiles * 16 * 128 * 16
  float* gates_save;      // training mode: [T, B, 4 NH] pre-activation gates (order i, j, f, o) of every step, or null
  float* c_save;          // training mode: [T, B, NH] cell state after every step, or null
  int B, T;
  float forget_bias;
  int* range_flag;
  long long* trace;       // debug (AIR_LSTM_TRACE): [CTA][64] SM-clock stamps, or null
};
#define LSTM_TRACE(slot)                                                                                              \
  do {                                                                                                                \
    if (p.trace) p.trace[((size_t)(blockIdx.y * CLUSTER + blockIdx.x)) * 64 + (slot)] = clock64();                   \
  } while (0)

__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, float (&v)[4]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait4(float (&a)[4], float (&b)[4], float (&c)[4], float (&d)[4]) {
  uint32_t* r0 = reinterpret_cast<uint32_t*>(a);
  uint32_t* r1 = reinterpret_cast<uint32_t*>(b);
  uint32_t* r2 = reinterpret_cast<uint32_t*>(c);
  uint32_t* r3 = reinterpret_cast<uint32_t*>(d);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r0[0]), "+r"(r0[1]), "+r"(r0[2]), "+r"(r0[3]), "+r"(r1[0]), "+r"(r1[1]), "+r"(r1[2]), "+r"(r1[3]),
                 "+r"(r2[0]), "+r"(r2[1]), "+r"(r2[2]), "+r"(r2[3]), "+r"(r3[0]), "+r"(r3[1]), "+r"(r3[2]), "+r"(r3[3])
               :
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// arrive (release, cluster scope) on the mbarrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 ld_cg_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ex2_fast(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// c' = sig(f + fb) * c + sig(i) * tanh(j);  h = tanh(c') * sig(o).  Five MUFU.EX2 + three MUFU.RCP per unit: each product
// of a sigmoid and a tanh shares one reciprocal, sig(a) * tanh(b) = (E - 1) / ((1 + e^-a) (1 + E)), E = e^{2b}.  Arguments
// are clamped where the functions have saturated below fp32 resolution, so no intermediate overflows.  Absolute error of
// each factor <= 3e-7 (ex2.approx / rcp.approx are good to 2^-22 relative).
__device__ __forceinline__ float sig_tanh(float a, float b) {
  const float ea = ex2_fast(-1.4426950408889634f * fminf(fmaxf(a, -30.f), 30.f));
  const float eb = ex2_fast(2.8853900817779268f * fminf(fmaxf(b, -15.f), 15.f));
  return (eb - 1.0f) * rcp_fast((1.0f + ea) * (1.0f + eb));
}
__device__ __forceinline__ float sig_fast(float a) {
  return rcp_fast(1.0f + ex2_fast(-1.4426950408889634f * fminf(fmaxf(a, -30.f), 30.f)));
}

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NUM_THREADS, 1)
lstm_cluster_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SLOTS * TILE_BYTES);
  uint64_t* empty_bar = full_bar + SLOTS;
  uint64_t* a_ready = empty_bar + SLOTS;
  uint64_t* d_full = a_ready + 1;
  uint64_t* x_bar = d_full + 1;           // h slabs of all four CTAs are in the exchange buffer
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(x_bar + 1);
  float* s_stage = reinterpret_cast<float*>(smem + STAGE_OFFSET);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.y;
  const int m0 = tile * BM;
  const bool fold0 = p.hw0 != nullptr;        // step 1's recurrent product is the constant hw0
  const int gbase = p.fuse_e2 ? 1 : 0;        // index of the gx group
  const int n_groups = gbase + 1 + p.T - (fold0 ? 1 : 0);   // [last encoder layer,] gx, then one GEMM per step (per step after the first when folded)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_x);
    prefetch_tmap(&p.tm_h);
    if (p.fuse_e2) prefetch_tmap(&p.tm_e2);
    for (int s = 0; s < SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(a_ready, NUM_EPI_WARPS * 32);
    mbar_init(d_full, 1);
    mbar_init(x_bar, CLUSTER * NUM_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive can reach them
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ===== TMA producer: this CTA's weight slab (rows rank * NH .. + NH of the permuted W^T), hi and lo tiles =====
    if (elect_one()) {
      uint32_t par = 0;
      for (int g = 0; g < n_groups; ++g) {
        const bool enc = p.fuse_e2 && g == 0;
        const CUtensorMap* tm = enc ? &p.tm_e2 : (g == gbase ? &p.tm_x : &p.tm_h);
        const int K = enc ? p.n_e1 : (g == gbase ? p.n_enc : NH);
        const int row_hi = enc ? 0 : (int)rank * NH, row_lo = enc ? p.e2_lo_row : CLUSTER * NH + (int)rank * NH;
        const int nkb = ((K + 15) / 16 + 3) / 4;
        for (int kb = 0; kb < nkb; ++kb) {
          const int hs = kb & 3, ls = 4 + (kb & 1);
          mbar_wait(&empty_bar[hs], ((par >> hs) & 1) ^ 1);
          par ^= 1u << hs;
          mbar_expect_tx(&full_bar[hs], TILE_BYTES);
          tma_load_2d(smem + hs * TILE_BYTES, tm, kb * BK, row_hi, &full_bar[hs]);
          mbar_wait(&empty_bar[ls], ((par >> ls) & 1) ^ 1);
          par ^= 1u << ls;
          mbar_expect_tx(&full_bar[ls], TILE_BYTES);
          tma_load_2d(smem + ls * TILE_BYTES, tm, kb * BK, row_lo, &full_bar[ls]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: cross terms first, then main terms (see chain_tc.cuh) =====
    if (elect_one()) {
      uint32_t par = 0;
      constexpr uint32_t idesc = make_idesc_f16(BM, NH);
      for (int g = 0; g < n_groups; ++g) {
        const int K = (p.fuse_e2 && g == 0) ? p.n_e1 : (g == gbase ? p.n_enc : NH);
        const int nsl = (K + 15) / 16, nkb = (nsl + 3) / 4;
        mbar_wait(a_ready, g & 1);
        tc_fence_after();
        LSTM_TRACE(32 + 2 * g);
        for (int kb = 0; kb < nkb; ++kb) {
          const int hs = kb & 3, ls = 4 + (kb & 1);
          mbar_wait(&full_bar[hs], (par >> hs) & 1);
          par ^= 1u << hs;
          mbar_wait(&full_bar[ls], (par >> ls) & 1);
          par ^= 1u << ls;
          tc_fence_after();
          const uint64_t db_hi = make_smem_desc_sw128(smem + hs * TILE_BYTES);
          const uint64_t db_lo = make_smem_desc_sw128(smem + ls * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int sl = kb * 4 + k;
            if (sl < nsl) {
              const uint32_t a_col = (uint32_t)sl * 8u;
              const uint64_t adv = (uint64_t)(k * 2);
              umma_f16_ts(tmem_base + D_COL, tmem_base + A_LO_COL + a_col, db_hi + adv, idesc, sl != 0);
              umma_f16_ts(tmem_base + D_COL, tmem_base + A_HI_COL + a_col, db_lo + adv, idesc, 1);
            }
          }
          umma_commit(&empty_bar[ls]);
        }
        for (int kb = 0; kb < nkb; ++kb) {
          const int hs = kb & 3;
          const uint64_t db_hi = make_smem_desc_sw128(smem + hs * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int sl = kb * 4 + k;
            if (sl < nsl)
              umma_f16_ts(tmem_base + D_COL, tmem_base + A_HI_COL + (uint32_t)sl * 8u, db_hi + (uint64_t)(k * 2), idesc, 1);
          }
          umma_commit(&empty_bar[hs]);
        }
        umma_commit(d_full);
        LSTM_TRACE(33 + 2 * g);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps: thread <-> (row, 16 of this CTA's 64 units) =====
    const int q = warp & 3, cq = (warp - 2) >> 2;
    const int row_w = m0 + q * 32;
    const int rit = q * 32 + lane;          // row inside the tile
    const int row = m0 + rit;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stage = s_stage + (warp - 2) * 512;
    const int u0 = (int)rank * UPC + cq * 16;   // first hidden unit of this thread
    const int my_slice = (int)rank * 4 + cq;    // == u0 / 16
    uint32_t ovf = 0;
    const bool tr = threadIdx.x == 64;
    if (tr) LSTM_TRACE(0);
    // With the 227 KB shared-memory carve-out the L1 is a few KB and every first touch is an L2 round trip; the bias and
    // initial-state lines this thread needs in the gx epilogue are pulled in now, behind the operand load and the gx GEMM
    // (measured: the gx epilogue was a chain of eight exposed L2 latencies, 13 k of the kernel's 96 k clocks).
#pragma unroll
    for (int gate = 0; gate < 4; ++gate)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + (size_t)rank * NH + gate * UPC + cq * 16));
    if (fold0) {
#pragma unroll
      for (int gate = 0; gate < 4; ++gate)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p.hw0 + (size_t)rank * NH + gate * UPC + cq * 16));
    } else if (p.h_init_ld == 0) {
#pragma unroll
      for (int s = 0; s < 4; ++s) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.h_init + (cq + 4 * s) * 16));
    }

    // --- operand of the gx GEMM: e rows, fp32 -> hi/lo -> TMEM (each thread: slices cq, cq + 4, ...) ---
    {
      const int nsl = ((p.fuse_e2 ? p.n_e1 : p.n_enc) + 15) / 16;
      if (p.e_hl) {
        // the encoder's last layer wrote e as hl planes: the packed words are read as they are (four 16-byte loads per
        // slice, all in flight together; the staged fp32 path below cost 11 k clocks of a 96 k kernel)
        const __half* src = p.e_nsl ? p.e_hl + ((size_t)tile * p.e_nsl * BM + rit) * 16 : p.e_hl + (size_t)row * p.e_ld;
        const size_t sstride = p.e_nsl ? (size_t)BM * 16 : 16;   // halves between consecutive slices of this row
#pragma unroll 4
        for (int s = cq; s < nsl; s += 4) {
          const uint4* sp = reinterpret_cast<const uint4*>(src + s * sstride);
          const uint4* lp = reinterpret_cast<const uint4*>(src + p.e_plane + s * sstride);
          const uint4 h0 = __ldg(sp), h1 = __ldg(sp + 1);
          const uint4 l0 = __ldg(lp);
          const uint4 l1 = __ldg(lp + 1);
          const uint32_t hi[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          const uint32_t lo[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
          tmem_st_32x8(t_lane + A_HI_COL + s * 8, hi);
          tmem_st_32x8(t_lane + A_LO_COL + s * 8, lo);
        }
      } else {
        for (int s = cq; s < nsl; s += 4) {
          float v[16];
          tile_load(stage, lane, v, p.e, p.n_enc, row_w, s * 16, p.n_enc, p.B);
          uint32_t hi[8], lo[8];
          split_pack16(v, hi, lo, ovf);
          tmem_st_32x8(t_lane + A_HI_COL + s * 8, hi);
          tmem_st_32x8(t_lane + A_LO_COL + s * 8, lo);
        }
      }
    }
    // cell state of this thread's 16 units, in registers for the whole recurrence
    float c_reg[16];
    if (p.c_in_ld == 0) load16(p.c_in + u0, c_reg);   // one trainable vector for every canvas (cell.py:103)
    else tile_load(stage, lane, c_reg, p.c_in, p.c_in_ld, row_w, u0, NH, p.B);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(a_ready);
    if (tr) LSTM_TRACE(1);

    // --- fused last encoder layer: e = ELU(D + b2) -> fp16 hi/lo -> the A operand of the gx GEMM (slices cq, cq + 4, ...) ---
    if (p.fuse_e2) {
      mbar_wait(d_full, 0);
      tc_fence_after();
      const int nsl_e = (p.n_enc + 15) / 16;
      for (int s = cq; s < nsl_e; s += 4) {
        float v[16], b[16];
        tmem_ld_32x16(t_lane + D_COL + s * 16, v);
        load16(p.bias_e2 + s * 16, b);
        tmem_ld_wait(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (s * 16 + j < p.n_enc) ? elu_fast(fmaf(v[j], W_UNSCALE, b[j])) : 0.f;
        uint32_t hi[8], lo[8];
        split_pack16(v, hi, lo, ovf);
        tmem_st_32x8(t_lane + A_HI_COL + s * 8, hi);
        tmem_st_32x8(t_lane + A_LO_COL + s * 8, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(a_ready);
    }

    // --- gx: D + bias -> scratch, in this thread's own read-back order: [gate * 4 + k4][row][4 floats] ---
    float* gx_mine = p.gx_scr + ((((size_t)tile * CLUSTER + rank) * 4 + cq) * 16 * BM + rit) * 4;
    if (!fold0) {
    mbar_wait(d_full, gbase & 1);
    tc_fence_after();
    if (tr) LSTM_TRACE(2);
#pragma unroll
    for (int gate = 0; gate < 4; ++gate) {
      float v[16], b[16];
      tmem_ld_32x16(t_lane + D_COL + gate * UPC + cq * 16, v);
      load16(p.bias + (size_t)rank * NH + gate * UPC + cq * 16, b);
      tmem_ld_wait(v);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        *reinterpret_cast<float4*>(gx_mine + (size_t)(gate * 4 + k4) * BM * 4) =
            make_float4(fmaf(v[4 * k4], W_UNSCALE, b[4 * k4]), fmaf(v[4 * k4 + 1], W_UNSCALE, b[4 * k4 + 1]),
                        fmaf(v[4 * k4 + 2], W_UNSCALE, b[4 * k4 + 2]), fmaf(v[4 * k4 + 3], W_UNSCALE, b[4 * k4 + 3]));
    }
    // --- operand of step 1: h_init rows ---
    for (int s = cq; s < NH / 16; s += 4) {
      float v[16];
      if (p.h_init_ld == 0) load16(p.h_init + s * 16, v);   // the broadcast trainable initial state: same 64 bytes in every lane
      else tile_load(stage, lane, v, p.h_init, p.h_init_ld, row_w, s * 16, NH, p.B);
      uint32_t hi[8], lo[8];
      split_pack16(v, hi, lo, ovf);
      tmem_st_32x8(t_lane + A_HI_COL + s * 8, hi);
      tmem_st_32x8(t_lane + A_LO_COL + s * 8, lo);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(a_ready);
    if (tr) LSTM_TRACE(3);
    }

    for (int t = 0; t < p.T; ++t) {
      const bool first_folded = fold0 && t == 0;   // the accumulator holds gx: form gx + bias (kept for the later steps) and add hw0
      mbar_wait(d_full, (gbase + (fold0 ? t : t + 1)) & 1);
      tc_fence_after();
      if (tr) LSTM_TRACE(4 + 5 * t);
      // ---- gate math of 16 units, four at a time ----
      float h_new[16];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        float gi[4], gj[4], gf[4], go[4];
        tmem_ld_32x4(t_lane + D_COL + 0 * UPC + cq * 16 + k4 * 4, gi);
        tmem_ld_32x4(t_lane + D_COL + 1 * UPC + cq * 16 + k4 * 4, gj);
        tmem_ld_32x4(t_lane + D_COL + 2 * UPC + cq * 16 + k4 * 4, gf);
        tmem_ld_32x4(t_lane + D_COL + 3 * UPC + cq * 16 + k4 * 4, go);
        float4 xi, xj, xf, xo;
        if (first_folded) {
          const size_t bo = (size_t)rank * NH + cq * 16 + k4 * 4;
          xi = *reinterpret_cast<const float4*>(p.bias + bo + 0 * UPC);
          xj = *reinterpret_cast<const float4*>(p.bias + bo + 1 * UPC);
          xf = *reinterpret_cast<const float4*>(p.bias + bo + 2 * UPC);
          xo = *reinterpret_cast<const float4*>(p.bias + bo + 3 * UPC);
        } else {
          xi = *reinterpret_cast<const float4*>(gx_mine + (size_t)(0 * 4 + k4) * BM * 4);
          xj = *reinterpret_cast<const float4*>(gx_mine + (size_t)(1 * 4 + k4) * BM * 4);
          xf = *reinterpret_cast<const float4*>(gx_mine + (size_t)(2 * 4 + k4) * BM * 4);
          xo = *reinterpret_cast<const float4*>(gx_mine + (size_t)(3 * 4 + k4) * BM * 4);
        }
        tmem_ld_wait4(gi, gj, gf, go);
        if (first_folded) {
          // gx + bias of this thread's 4 x 4 values goes to the scratch the later steps re-read; the addend of THIS step
          // becomes bias + hw0
          if (p.T > 1) {
            *reinterpret_cast<float4*>(gx_mine + (size_t)(0 * 4 + k4) * BM * 4) = make_float4(
                fmaf(gi[0], W_UNSCALE, xi.x), fmaf(gi[1], W_UNSCALE, xi.y), fmaf(gi[2], W_UNSCALE, xi.z), fmaf(gi[3], W_UNSCALE, xi.w));
            *reinterpret_cast<float4*>(gx_mine + (size_t)(1 * 4 + k4) * BM * 4) = make_float4(
                fmaf(gj[0], W_UNSCALE, xj.x), fmaf(gj[1], W_UNSCALE, xj.y), fmaf(gj[2], W_UNSCALE, xj.z), fmaf(gj[3], W_UNSCALE, xj.w));
            *reinterpret_cast<float4*>(gx_mine + (size_t)(2 * 4 + k4) * BM * 4) = make_float4(
                fmaf(gf[0], W_UNSCALE, xf.x), fmaf(gf[1], W_UNSCALE, xf.y), fmaf(gf[2], W_UNSCALE, xf.z), fmaf(gf[3], W_UNSCALE, xf.w));
            *reinterpret_cast<float4*>(gx_mine + (size_t)(3 * 4 + k4) * BM * 4) = make_float4(
                fmaf(go[0], W_UNSCALE, xo.x), fmaf(go[1], W_UNSCALE, xo.y), fmaf(go[2], W_UNSCALE, xo.z), fmaf(go[3], W_UNSCALE, xo.w));
          }
          const size_t bo = (size_t)rank * NH + cq * 16 + k4 * 4;
          const float4 hi4 = *reinterpret_cast<const float4*>(p.hw0 + bo + 0 * UPC);
          const float4 hj4 = *reinterpret_cast<const float4*>(p.hw0 + bo + 1 * UPC);
          const float4 hf4 = *reinterpret_cast<const float4*>(p.hw0 + bo + 2 * UPC);
          const float4 ho4 = *reinterpret_cast<const float4*>(p.hw0 + bo + 3 * UPC);
          xi = make_float4(xi.x + hi4.x, xi.y + hi4.y, xi.z + hi4.z, xi.w + hi4.w);
          xj = make_float4(xj.x + hj4.x, xj.y + hj4.y, xj.z + hj4.z, xj.w + hj4.w);
          xf = make_float4(xf.x + hf4.x, xf.y + hf4.y, xf.z + hf4.z, xf.w + hf4.w);
          xo = make_float4(xo.x + ho4.x, xo.y + ho4.y, xo.z + ho4.z, xo.w + ho4.w);
        }
        const float ai[4] = {xi.x, xi.y, xi.z, xi.w}, aj[4] = {xj.x, xj.y, xj.z, xj.w};
        const float af[4] = {xf.x, xf.y, xf.z, xf.w}, ao[4] = {xo.x, xo.y, xo.z, xo.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float pi = fmaf(gi[j], W_UNSCALE, ai[j]), pj = fmaf(gj[j], W_UNSCALE, aj[j]);
          const float pf = fmaf(gf[j], W_UNSCALE, af[j]), po = fmaf(go[j], W_UNSCALE, ao[j]);
          const float cn = fmaf(sig_fast(pf + p.forget_bias), c_reg[4 * k4 + j], sig_tanh(pi, pj));
          c_reg[4 * k4 + j] = cn;
          h_new[4 * k4 + j] = sig_tanh(po, cn);
        }
        if (p.gates_save && row < p.B) {   // kept for the backward pass: 16-byte pieces of this row's four gate vectors
          float* gs = p.gates_save + ((size_t)t * p.B + row) * (4 * NH) + u0 + 4 * k4;
          *reinterpret_cast<float4*>(gs) = make_float4(fmaf(gi[0], W_UNSCALE, ai[0]), fmaf(gi[1], W_UNSCALE, ai[1]),
                                                       fmaf(gi[2], W_UNSCALE, ai[2]), fmaf(gi[3], W_UNSCALE, ai[3]));
          *reinterpret_cast<float4*>(gs + NH) = make_float4(fmaf(gj[0], W_UNSCALE, aj[0]), fmaf(gj[1], W_UNSCALE, aj[1]),
                                                            fmaf(gj[2], W_UNSCALE, aj[2]), fmaf(gj[3], W_UNSCALE, aj[3]));
          *reinterpret_cast<float4*>(gs + 2 * NH) = make_float4(fmaf(gf[0], W_UNSCALE, af[0]), fmaf(gf[1], W_UNSCALE, af[1]),
                                                                fmaf(gf[2], W_UNSCALE, af[2]), fmaf(gf[3], W_UNSCALE, af[3]));
          *reinterpret_cast<float4*>(gs + 3 * NH) = make_float4(fmaf(go[0], W_UNSCALE, ao[0]), fmaf(go[1], W_UNSCALE, ao[1]),
                                                                fmaf(go[2], W_UNSCALE, ao[2]), fmaf(go[3], W_UNSCALE, ao[3]));
        }
      }
      if (tr) LSTM_TRACE(5 + 5 * t);
      if (p.c_save) tile_store(stage, lane, c_reg, p.c_save + (size_t)t * p.B * NH, NH, row_w, u0, NH, p.B);
      // ---- h_t: fp32 rows (output + final state), hl copy for the heads chain, own slice straight into TMEM, and the
      //      exchange buffer for the other three CTAs ----
      if (!p.hs_last_only || t + 1 == p.T) tile_store(stage, lane, h_new, p.hs + (size_t)t * p.B * NH, NH, row_w, u0, NH, p.B);
      uint32_t hi[8], lo[8];
      split_pack16(h_new, hi, lo, ovf);
      if (p.hs_hlt.p && row < p.B) {
        __half* d = p.hs_hlt.p + hl_index(p.hs_hlt, (size_t)t * p.B + row, u0);
        *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(d + 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(d + p.hs_hlt.plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(d + p.hs_hlt.plane + 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      if (t + 1 == p.T) {
        tile_store(stage, lane, c_reg, p.c, NH, row_w, u0, NH, p.B);
        break;
      }
      tmem_st_32x8(t_lane + A_HI_COL + my_slice * 8, hi);
      tmem_st_32x8(t_lane + A_LO_COL + my_slice * 8, lo);
      // The h slabs reach the other three CTAs through L2.  When the canvas tile is also a tile of the heads' operand
      // (B a multiple of 128: row t * B + m0 starts a 128-row tile) the slice-major hl copy just written IS the exchange
      // buffer -- same [slice][128 rows][16] layout -- and nothing is stored twice; otherwise a private double buffer.
      const bool via_hlt = p.hs_hlt.p && (p.B % BM) == 0;
      __half* xb = via_hlt ? p.hs_hlt.p + ((((size_t)t * p.B + m0) >> 7) * (size_t)p.hs_hlt.nsl) * BM * 16
                           : p.hx + (size_t)(t & 1) * 2 * p.hx_plane + ((size_t)tile * 16 * BM) * 16;
      const size_t xplane = via_hlt ? p.hs_hlt.plane : p.hx_plane;
      if (!via_hlt) {
        __half* d = xb + ((size_t)my_slice * BM + rit) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(d + 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(d + p.hx_plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(d + p.hx_plane + 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      __syncwarp();
      if (tr) LSTM_TRACE(6 + 5 * t);
      if (lane == 0) {
#pragma unroll
        for (uint32_t r = 0; r < CLUSTER; ++r) mbar_arrive_remote(x_bar, r);
      }
      mbar_wait_cluster(x_bar, t & 1);
      if (tr) LSTM_TRACE(7 + 5 * t);
      // ---- the other CTAs' slabs: slices 4 r' + cq, r' != rank ----
#pragma unroll
      for (int rr = 1; rr < CLUSTER; ++rr) {
        const int s = ((int)((rank + rr) & 3)) * 4 + cq;
        const __half* src = xb + ((size_t)s * BM + rit) * 16;
        const uint4 h0 = ld_cg_u4(src), h1 = ld_cg_u4(src + 8);
        const uint4 l0 = ld_cg_u4(src + xplane), l1 = ld_cg_u4(src + xplane + 8);
        const uint32_t whi[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const uint32_t wlo[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        tmem_st_32x8(t_lane + A_HI_COL + s * 8, whi);
        tmem_st_32x8(t_lane + A_LO_COL + s * 8, wlo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(a_ready);
      if (tr) LSTM_TRACE(8 + 5 * t);
    }
    if (tr) LSTM_TRACE(30);
    if ((ovf & 0x80008000u) && p.range_flag) atomicOr(p.range_flag, 1);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while a peer may still arrive on its x_bar
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// hw0[perm(n)] = sum_k h0[k] * W[n_enc + k][n] in fp32: the recurrent product of the broadcast trainable initial state
// (cell.py:103), the same for every canvas; perm = the [cta][gate][unit] order of `bias` (linear_tc.cuh: prep_dst_row)
__global__ void __launch_bounds__(256)
lstm_h0w_kernel(const float* __restrict__ w_h, const float* __restrict__ h0, float* __restrict__ hw0, int nh) {
  griddep_launch();
  griddep_wait();
  // block = 32 outputs x 8 slices of k: eight independent partial sums per output, loads of a warp contiguous in n
  __shared__ float part[8][33];
  const int nl = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + nl;
  float acc = 0.f;
  if (n < 4 * nh) {
#pragma unroll 8
    for (int k = ks; k < nh; k += 8) acc = fmaf(h0[k], w_h[(size_t)k * 4 * nh + n], acc);
  }
  part[ks][nl] = acc;
  __syncthreads();
  if (ks == 0 && n < 4 * nh) {
    float sum = part[0][nl];
#pragma unroll
    for (int j = 1; j < 8; ++j) sum += part[j][nl];
    const int upc = nh >> 2, gate = n / nh, unit = n % nh;
    hw0[(unit / upc) * nh + gate * upc + unit % upc] = sum;
  }
}

inline cudaError_t launch_lstm(const Params& p, cudaStream_t st) {
  cudaError_t e = ensure_dynamic_smem(lstm_cluster_kernel, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_k(lstm_cluster_kernel, dim3(CLUSTER, (p.B + BM - 1) / BM), dim3(NUM_THREADS), SMEM_BYTES, st, p);
}

}  // namespace lstm
}  // namespace air
