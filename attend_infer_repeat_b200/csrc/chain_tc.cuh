// Fused dense-layer CHAINS on the tensor cores (sm_100a): a whole MLP (or several, back to back) per launch, one
// 128-row tile per CTA, activations never leaving the SM between layers.
//
//   * the A operand (activations, fp16 hi/lo split like linear_tc.cuh) lives in TENSOR MEMORY: tcgen05.mma's "TS" form
//     (A from TMEM, B from a shared-memory descriptor).  Row r of the tile is TMEM lane r; K elements are packed two per
//     32-bit column: hi plane in columns [256, 384), lo plane in [384, 512) -- up to K = 256 resident;
//   * the fp32 accumulator D occupies columns [0, 256), N <= 256 per pass.  ONE accumulator takes all three split
//     products, but in two sweeps over K: first every cross term (A_lo*W_hi + A_hi*W_lo, ~2^-11 of the result), then
//     every main term A_hi*W_hi.  The tensor core truncates the fp32 accumulator once per MMA; ordered like this the
//     truncations that matter (relative to the full-size sum) are one per K=16 slice, exactly as with the two separate
//     accumulators of linear_tc.cuh, instead of three;
//   * shared memory holds nothing but weight tiles of [<=256 rows(N)][64 K] fp16 (128-byte swizzled), streamed from the
//     L2-resident prepared-weight arena by TMA into six 32 KB slots: four "hi" slots (k-block kb -> slot kb & 3, held
//     for both sweeps) and two rotating "lo" slots (freed during the cross sweep);
//   * a layer's epilogue (16 warps, one TMEM lane = one row per thread) reads D with tcgen05.ld, applies bias + ELU,
//     splits to fp16 hi/lo and writes the NEXT layer's A operand straight back into TMEM with tcgen05.st -- no shared
//     memory, no HBM round trip, no swizzle arithmetic.  Output layers write fp32 rows to HBM; the what head
//     (ParametrisedGaussian + reparameterised sample, modules.py:11-24, cell.py:154-156) is an epilogue variant.
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2..17 = epilogue
// (TMEM lane quadrant = warp & 3, column quarter = (warp - 2) >> 2).
// The unit of synchronisation is a GROUP = (layer, N-pass, K-chunk of <= 256): the epilogue warps announce "A is in
// TMEM and D is free" (a_ready), the MMA warp issues the group's MMAs and commits to d_full.
#pragma once
#include "linear_tc.cuh"

namespace air {
namespace chain {

using namespace air::tc;

constexpr int MAXL = 12;
constexpr int D_COL = 0, A_HI_COL = 256, A_LO_COL = 384;
constexpr int SLOTS = 6;                       // 0..3: W_hi of k-block kb & 3; 4..5: W_lo of k-block kb & 1
constexpr int TILE_BYTES = 256 * 128;          // one slot: up to 256 rows x 64 fp16
constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int STAGE_OFFSET = SLOTS * TILE_BYTES + 256;   // per-warp 2 KB staging tiles
constexpr int SMEM_BYTES = STAGE_OFFSET + NUM_EPI_WARPS * 2048 + 1024;

enum { EPI_ELU_A = 0, EPI_F32 = 1, EPI_WHAT = 2 };
enum { A_KEEP = 0, A_LOAD = 1 };

struct Layer {
  int K;         // contraction length (A_KEEP: <= 256)
  int N;         // true output width
  int n_box;     // MMA N = rows of one weight tile (multiple of 16, <= 256)
  int n_pass;    // passes over N (only output layers may have more than one)
  int lo_row;    // row of the lo plane inside the weight tensor map (= N_alloc of the prepared weight)
  int epi;       // EPI_*
  int a_src;     // A_KEEP: operand left in TMEM by the previous layer's epilogue; A_LOAD: read from in[a_buf]
  int a_buf;
  const float* bias;   // zero-padded to n_pass * n_box entries (prepared with the weights)
  float* out;    // EPI_F32: fp32 [rows, ldo]
  int ldo;
  float* save;   // EPI_ELU_A, training mode: the activation is also kept as fp32 rows [rows, N] for the backward pass
};
struct HlIn {
  const __half* p;   // slice-major tiled hl operand: [row tile of 128][K slice][128 rows][16 fp16]; lo plane at + plane
  size_t plane;
  int nsl;           // K slices per row (round_up(K, 16) / 16)
};
struct Params {
  CUtensorMap tm[MAXL];
  Layer layer[MAXL];
  int n_layers;
  int M;
  HlIn in[2];
  // what head (EPI_WHAT): D columns [0, na) = loc, [na_off, na_off + na) = raw scale
  const float* eps_what;
  float* what;
  float* what_loc;
  float* what_scale;
  int na, na_off;
  float what_offset;
  int* range_flag;
  long long* trace;   // debug (AIR_CHAIN_TRACE): [CTA][TRACE_GROUPS][8] SM-clock stamps per group, or null
};
constexpr int TRACE_GROUPS = 32;
#define CHAIN_TRACE(grp, slot)                                                                           \
  do {                                                                                                   \
    if (p.trace && (grp) < TRACE_GROUPS) p.trace[((size_t)blockIdx.x * TRACE_GROUPS + (grp)) * 8 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// 16 fp32 values -> the 8 + 8 TMEM words of 16 consecutive K elements (hi plane, lo plane), two elements per
// conversion instruction.  `ovf` accumulates a flag in bits 15 / 31 when a hi half is Inf or NaN (exponent all ones).
__device__ __forceinline__ void split_pack16(const float (&x)[16], uint32_t (&hi)[8], uint32_t (&lo)[8], uint32_t& ovf) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __half2 h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x[2 * j] - hf.x, x[2 * j + 1] - hf.y);
    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    ovf |= (hi[j] & 0x7C007C00u) + 0x04000400u;
  }
}

// ELU with one MUFU.EX2 and no range fix-up: for y <= 0, ex2.approx.ftz(y * log2 e) - 1 is within 2.4e-7 of
// expf(y) - 1 (flushing results below 2^-126 to 0 changes nothing after the "- 1")
__device__ __forceinline__ float elu_fast(float y) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * 1.4426950408889634f));
  return y > 0.f ? y : e - 1.0f;
}

// 16 consecutive floats of a 64-byte aligned array, same address in every lane (one L1 wavefront per 16 bytes)
__device__ __forceinline__ void load16(const float* __restrict__ src, float (&b)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src) + j);
    b[4 * j] = t.x; b[4 * j + 1] = t.y; b[4 * j + 2] = t.z; b[4 * j + 3] = t.w;
  }
}

// Row-major fp32 HBM <-> "one row per lane" registers through a per-warp 32 x 16 staging tile in shared memory, so that
// the global accesses of a warp cover whole row segments instead of 32 different lines.  The tile is XOR-swizzled by
// 16-byte chunk (chunk ^ ((row >> 1) & 3)): both the row-per-lane and the row-segment-per-lane access are conflict-free.
__device__ __forceinline__ int stage_idx(int r, int chunk) { return r * 16 + ((chunk ^ ((r >> 1) & 3)) << 2); }

// v[j] of lane r = element (row_w + r, col0 + j); columns >= n_cols and rows >= n_rows are not written
__device__ __forceinline__ void tile_store(float* stage, int lane, const float (&v)[16], float* __restrict__ out, int ld,
                                           int row_w, int col0, int n_cols, int n_rows) {
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4)
    *reinterpret_cast<float4*>(stage + stage_idx(lane, c4)) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
  __syncwarp();
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && col0 + 16 <= n_cols;
  if (vec) {
    const int c4 = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (lane >> 2) + 8 * i;
      const float4 t = *reinterpret_cast<const float4*>(stage + stage_idx(r, c4));
      if (row_w + r < n_rows) *reinterpret_cast<float4*>(out + (size_t)(row_w + r) * ld + col0 + 4 * c4) = t;
    }
  } else {
    const int cc = lane & 15;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = (lane >> 4) + 2 * i;
      const float t = stage[stage_idx(r, cc >> 2) + (cc & 3)];
      if (row_w + r < n_rows && col0 + cc < n_cols) out[(size_t)(row_w + r) * ld + col0 + cc] = t;
    }
  }
  __syncwarp();
}
// the reverse: v[j] of lane r = in[row_w + r, col0 + j] (0 outside)
__device__ __forceinline__ void tile_load(float* stage, int lane, float (&v)[16], const float* __restrict__ in, int ld,
                                          int row_w, int col0, int n_cols, int n_rows) {
  const int cc = lane & 15;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = (lane >> 4) + 2 * i;
    float t = 0.f;
    if (row_w + r < n_rows && col0 + cc < n_cols) t = __ldg(in + (size_t)(row_w + r) * ld + col0 + cc);
    stage[stage_idx(r, cc >> 2) + (cc & 3)] = t;
  }
  __syncwarp();
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 t = *reinterpret_cast<const float4*>(stage + stage_idx(lane, c4));
    v[4 * c4] = t.x; v[4 * c4 + 1] = t.y; v[4 * c4 + 2] = t.z; v[4 * c4 + 3] = t.w;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(NUM_THREADS, 1) chain_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SLOTS * TILE_BYTES);
  uint64_t* empty_bar = full_bar + SLOTS;
  uint64_t* a_ready = empty_bar + SLOTS;
  uint64_t* d_full = a_ready + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(d_full + 1);
  float* s_stage = reinterpret_cast<float*>(smem + STAGE_OFFSET);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < p.n_layers; ++l) prefetch_tmap(&p.tm[l]);
    for (int s = 0; s < SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(a_ready, NUM_EPI_WARPS * 32);
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ===== TMA producer: weight tiles in the order the MMA warp consumes them.  `par` bit s = parity of the number of
    //       times slot s has been filled so far =====
    if (elect_one()) {
      uint32_t par = 0;
      for (int l = 0; l < p.n_layers; ++l) {
        const Layer& L = p.layer[l];
        const int nkb = ((L.K + 15) / 16 + 3) / 4;
        const uint32_t tile_bytes = (uint32_t)L.n_box * 128u;
        for (int pass = 0; pass < L.n_pass; ++pass)
          for (int kb = 0; kb < nkb; ++kb) {
            const int hs = kb & 3, ls = 4 + (kb & 1);
            mbar_wait(&empty_bar[hs], ((par >> hs) & 1) ^ 1);
            par ^= 1u << hs;
            mbar_expect_tx(&full_bar[hs], tile_bytes);
            tma_load_2d(smem + hs * TILE_BYTES, &p.tm[l], kb * BK, pass * L.n_box, &full_bar[hs]);
            mbar_wait(&empty_bar[ls], ((par >> ls) & 1) ^ 1);
            par ^= 1u << ls;
            mbar_expect_tx(&full_bar[ls], tile_bytes);
            tma_load_2d(smem + ls * TILE_BYTES, &p.tm[l], kb * BK, L.lo_row + pass * L.n_box, &full_bar[ls]);
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      uint32_t par = 0;
      int g = 0;
      for (int l = 0; l < p.n_layers; ++l) {
        const Layer& L = p.layer[l];
        const int nsl = (L.K + 15) / 16, nkb = (nsl + 3) / 4, nchunks = (nsl + 15) / 16;
        const uint32_t idesc = make_idesc_f16(BM, L.n_box);
        for (int pass = 0; pass < L.n_pass; ++pass)
          for (int c = 0; c < nchunks; ++c, ++g) {
            mbar_wait(a_ready, g & 1);
            tc_fence_after();
            CHAIN_TRACE(g, 0);
            const int kb_end = min(nkb, 4 * c + 4);
            // sweep 1: cross terms
            for (int kb = 4 * c; kb < kb_end; ++kb) {
              const int hs = kb & 3, ls = 4 + (kb & 1);
              mbar_wait(&full_bar[hs], (par >> hs) & 1);
              par ^= 1u << hs;
              mbar_wait(&full_bar[ls], (par >> ls) & 1);
              par ^= 1u << ls;
              tc_fence_after();
              if (kb == 4 * c) CHAIN_TRACE(g, 1);
              if (kb == kb_end - 1) CHAIN_TRACE(g, 2);
              const uint64_t db_hi = make_smem_desc_sw128(smem + hs * TILE_BYTES);
              const uint64_t db_lo = make_smem_desc_sw128(smem + ls * TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int sl = kb * 4 + k;
                if (sl < nsl) {
                  const uint32_t a_col = (uint32_t)(sl - 16 * c) * 8u;
                  const uint64_t adv = (uint64_t)(k * 2);   // 16 fp16 = 32 bytes along K, >> 4
                  umma_f16_ts(tmem_base + D_COL, tmem_base + A_LO_COL + a_col, db_hi + adv, idesc, sl != 0);
                  umma_f16_ts(tmem_base + D_COL, tmem_base + A_HI_COL + a_col, db_lo + adv, idesc, 1);
                }
              }
              umma_commit(&empty_bar[ls]);
            }
            // sweep 2: main terms
            for (int kb = 4 * c; kb < kb_end; ++kb) {
              const int hs = kb & 3;
              const uint64_t db_hi = make_smem_desc_sw128(smem + hs * TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int sl = kb * 4 + k;
                if (sl < nsl) {
                  const uint32_t a_col = (uint32_t)(sl - 16 * c) * 8u;
                  umma_f16_ts(tmem_base + D_COL, tmem_base + A_HI_COL + a_col, db_hi + (uint64_t)(k * 2), idesc, 1);
                }
              }
              umma_commit(&empty_bar[hs]);
            }
            umma_commit(d_full);
            CHAIN_TRACE(g, 3);
          }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps: thread <-> (row = TMEM lane, quarter of the columns) =====
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int cq = (warp - 2) >> 2;         // which quarter of the columns / K slices
    const int row_w = m0 + q * 32;          // first row of this warp
    const int row = row_w + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stage = s_stage + (warp - 2) * 512;
    uint32_t ovf = 0;
    int g = 0;
    for (int l = 0; l < p.n_layers; ++l) {
      const Layer& L = p.layer[l];
      const int nsl = (L.K + 15) / 16, nchunks = (nsl + 15) / 16;
      for (int pass = 0; pass < L.n_pass; ++pass)
        for (int c = 0; c < nchunks; ++c, ++g) {
          if (L.a_src == A_LOAD && (pass == 0 || nchunks > 1)) {
            // ---- bring K-chunk c of this tile's rows into TMEM.  The producer kernel wrote the operand slice-major
            //      ([tile][K slice][row][16 fp16], hi and lo planes): a warp reads 1 KB contiguous per slice and plane,
            //      and the words ARE the packed TMEM words ----
            const HlIn in = p.in[L.a_buf];
            const int ns_c = min(16, nsl - 16 * c);
            const int s_begin = (cq * ns_c) >> 2, s_end = ((cq + 1) * ns_c) >> 2;
            const __half* src = in.p + (((size_t)blockIdx.x * in.nsl + 16 * c) * BM + (q * 32 + lane)) * 16;
#pragma unroll 2
            for (int s = s_begin; s < s_end; ++s) {
              uint32_t hi[8], lo[8];
              const uint4* ph = reinterpret_cast<const uint4*>(src + (size_t)s * BM * 16);
              const uint4* pl = reinterpret_cast<const uint4*>(src + in.plane + (size_t)s * BM * 16);
              const uint4 h0 = __ldg(ph), h1 = __ldg(ph + 1), l0 = __ldg(pl), l1 = __ldg(pl + 1);
              hi[0] = h0.x; hi[1] = h0.y; hi[2] = h0.z; hi[3] = h0.w; hi[4] = h1.x; hi[5] = h1.y; hi[6] = h1.z; hi[7] = h1.w;
              lo[0] = l0.x; lo[1] = l0.y; lo[2] = l0.z; lo[3] = l0.w; lo[4] = l1.x; lo[5] = l1.y; lo[6] = l1.z; lo[7] = l1.w;
              tmem_st_32x8(t_lane + A_HI_COL + s * 8, hi);
              tmem_st_32x8(t_lane + A_LO_COL + s * 8, lo);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          if (threadIdx.x == 64) CHAIN_TRACE(g, 4);
          mbar_arrive(a_ready);
          mbar_wait(d_full, g & 1);
          tc_fence_after();
          if (threadIdx.x == 64) CHAIN_TRACE(g, 5);
          if (c != nchunks - 1) continue;

          // ---- epilogue of (layer, pass): this thread owns 16-column groups [g_begin, g_end) of its row ----
          const int ng = L.epi == EPI_WHAT ? p.na_off / 16 : L.n_box / 16;
          const int g_begin = (cq * ng) >> 2, g_end = ((cq + 1) * ng) >> 2;
          const float* bias = L.bias + pass * L.n_box;
          if (L.epi == EPI_ELU_A) {
            for (int gi = g_begin; gi < g_end; ++gi) {
              float v[16], b[16];
              tmem_ld_32x16(t_lane + D_COL + gi * 16, v);
              load16(bias + gi * 16, b);
              tmem_ld_wait(v);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = elu_fast(fmaf(v[j], W_UNSCALE, b[j]));
              uint32_t hi[8], lo[8];
              split_pack16(v, hi, lo, ovf);
              tmem_st_32x8(t_lane + A_HI_COL + gi * 8, hi);
              tmem_st_32x8(t_lane + A_LO_COL + gi * 8, lo);
              if (L.save) tile_store(stage, lane, v, L.save, L.N, row_w, gi * 16, L.N, p.M);
            }
          } else if (L.epi == EPI_F32) {
            const int n_base = pass * L.n_box;
            for (int gi = g_begin; gi < g_end; ++gi) {
              const int n0 = n_base + gi * 16;
              if (n0 >= L.N) break;
              float v[16];
              tmem_ld_32x16(t_lane + D_COL + gi * 16, v);
              float b[16];
              load16(bias + gi * 16, b);
              tmem_ld_wait(v);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], W_UNSCALE, b[j]);
              tile_store(stage, lane, v, L.out, L.ldo, row_w, n0, L.N, p.M);
            }
          } else {   // EPI_WHAT
            const int na = p.na;
            for (int gi = g_begin; gi < g_end; ++gi) {
              float vl[16], vs[16], e[16];
              tmem_ld_32x16(t_lane + D_COL + gi * 16, vl);
              tmem_ld_32x16(t_lane + D_COL + p.na_off + gi * 16, vs);
              tile_load(stage, lane, e, p.eps_what, na, row_w, gi * 16, na, p.M);
              float bl[16], bs[16];
              load16(bias + gi * 16, bl);
              load16(bias + p.na_off + gi * 16, bs);
              tmem_ld_wait(vl);
              tmem_ld_wait(vs);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const bool ok = gi * 16 + j < na;
                const float loc = fmaf(vl[j], W_UNSCALE, bl[j]);
                const float sc = softplus_f(fmaf(vs[j], W_UNSCALE, bs[j]) + p.what_offset);
                vl[j] = loc;
                vs[j] = sc;
                e[j] = ok ? __fadd_rn(__fmul_rn(e[j], sc), loc) : 0.f;
              }
              tile_store(stage, lane, vl, p.what_loc, na, row_w, gi * 16, na, p.M);
              tile_store(stage, lane, vs, p.what_scale, na, row_w, gi * 16, na, p.M);
              tile_store(stage, lane, e, p.what, na, row_w, gi * 16, na, p.M);
              uint32_t hi[8], lo[8];
              split_pack16(e, hi, lo, ovf);
              tmem_st_32x8(t_lane + A_HI_COL + gi * 8, hi);
              tmem_st_32x8(t_lane + A_LO_COL + gi * 8, lo);
            }
          }
        }
    }
    if ((ovf & 0x80008000u) && p.range_flag) atomicOr(p.range_flag, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// tensor map over a prepared weight (W^T hl planes, [2 * n_alloc][kpad] fp16) with a box of `n_box` rows x 64 K
inline bool make_weight_tmap(CUtensorMap* tm, const __half* base, int kpad, int n_alloc, int n_box) {
  return make_tmap(tm, base, kpad, 2 * (int64_t)n_alloc, n_box);
}

inline cudaError_t launch_chain(const Params& p, cudaStream_t st) {
  cudaError_t e = ensure_dynamic_smem(chain_kernel, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_k(chain_kernel, dim3((p.M + BM - 1) / BM), dim3(NUM_THREADS), SMEM_BYTES, st, p);
}

}  // namespace chain
}  // namespace air
