// The whole per-(step, canvas) ROW path of the AIR cell in ONE launch (sm_100a):
//   h_t -> where MLP -> where sampling -> steps MLP -> STN glimpse read -> glimpse Encoder -> what head (sample)
//       -> Decoder -> decoded glimpse          (cell.py:129-135,137-138,153-158; modules.py:11-24,41-63,104-109,119-122)
// One 128-row tile per CTA (row = t * B + b).  Activations never leave the SM; the glimpse crop is sampled from the
// L2-resident image straight into the tensor-memory A operand (no crop round trip through HBM, no separate read kernel).
//
// Compared with chain_tc.cuh (one accumulator, epilogue fully exposed between dependent layers) this kernel is a small
// dataflow machine whose three roles run host-built programs:
//   * tensor memory holds TWO 128-column accumulators D[0], D[1] and TWO A-operand halves A[0], A[1] (128 K each: 64
//     columns of fp16 hi pairs + 64 of lo pairs).  A layer's N is cut into sub-tiles of <= 128 columns and its K into
//     parts of <= 128; one UNIT = (layer, n-sub, k-part) = 24 tcgen05.mma (cross sweep lo*hi + hi*lo, then main sweep);
//   * while the tensor core works on n-sub 1 of a layer the 16 epilogue warps drain n-sub 0 (bias, ELU, fp16 hi/lo
//     split) into registers and store it as A[0] of the NEXT layer as soon as the last reader of A[0] has retired; the
//     next layer's first unit starts on A[0] while n-sub 1 is being drained.  The tensor pipe idles for roughly half an
//     epilogue per layer instead of a whole one;
//   * weights stream from the L2-resident prepared arena by TMA into three 64 KB unit buffers, up to three units ahead.
// Synchronisation is six mbarrier families: w_full / w_free (unit buffers), a_ready / a_free (A halves),
// d_full / d_free (accumulators).  build_schedule() below derives units, tasks and flags from the layer list;
// simulate() replays the three programs on the host and proves them deadlock- and hazard-free (CPU test).
#pragma once
#include "chain_tc.cuh"

namespace air {
namespace row {

using namespace air::chain;

constexpr int MAXU = 128;          // units per launch
constexpr int MAXTASK = 80;        // epilogue tasks per launch
constexpr int MAXTM = 16;          // tensor maps (layers)
constexpr int W_SLOTS = 3;
constexpr int W_TILE = 128 * 128;  // [128 rows (N)][64 K] fp16, 128-byte swizzled
constexpr int W_SLOT = 4 * W_TILE; // up to two k-blocks x (hi, lo)
constexpr int EPI_WARPS = 16;       // 4 per TMEM lane quadrant.  18 warps per CTA leave every thread 96 registers and the 227 KB
                                   // shared-memory carve-out leaves an L1 of a few KB, so anything that spills pays L2 round
                                   // trips: the epilogues are written to stay inside that budget, and heavy SIMT work (the
                                   // glimpse gather) lives in its own kernel with full occupancy (measured: fused into this
                                   // kernel it took 4x the time of the whole MLP chain)
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int ROW_THREADS = 64 + EPI_THREADS;
constexpr int BAR_OFFSET = W_SLOTS * W_SLOT;
constexpr int ROW_STAGE_OFFSET = BAR_OFFSET + 1024;     // per-warp 2 KB staging tiles (the where codes alias them)
constexpr int ROW_SMEM_BYTES = ROW_STAGE_OFFSET + EPI_WARPS * 2048 + 1024;

__host__ __device__ constexpr uint32_t d_col(int d) { return (uint32_t)d * 128u; }
__host__ __device__ constexpr uint32_t a_col(int h) { return 256u + (uint32_t)h * 128u; }   // hi; lo at + 64

enum { U_WAIT_A = 1, U_WAIT_D = 2, U_ACC0 = 4, U_COMMIT_D = 8, U_COMMIT_A = 16 };
enum { T_LOAD_HL = 0, T_RESERVED = 1, T_ELU = 2, T_OUT = 3, T_WHAT = 4, T_WHERE = 5 };
enum { OUT_GLOBAL = 0, OUT_M_SMEM = 1 };

struct Unit {
  uint16_t n_row;     // row of the hi tile in the weight tensor map
  uint16_t lo_row;    // rows between the hi and the lo plane (N_alloc)
  uint16_t kb0;       // first 64-wide k-block
  uint8_t tm;         // tensor map index
  uint8_t nkb;        // k-blocks in this unit (1..2)
  uint8_t nsl;        // 16-wide K slices in this unit (1..8)
  uint8_t n16;        // MMA N / 16 (1..8)
  uint8_t a_half, d_idx;
  uint8_t flags;
  uint8_t fill;       // debug / simulate(): id of the A fill this unit reads
  uint8_t use;        // debug / simulate(): id of the accumulator use this unit belongs to
  uint8_t box16;      // rows of the tensor map's box / 16 (128 unless the prepared weight has fewer rows)
};
struct Task {
  uint8_t type;
  uint8_t d_idx;      // epilogue tasks: accumulator
  uint8_t a_half;     // loads, T_ELU, T_WHAT: the A half written
  uint8_t nsl;        // loads: K slices in this part (<= 8)
  uint16_t s0;        // loads: first K slice of the part; epilogues: first column of this n-sub within the layer
  uint16_t n_valid;   // epilogues: columns of the layer covered by this n-sub (<= 128)
  uint8_t buf;        // T_LOAD_HL: which HlIn
  uint8_t out_kind;   // T_OUT: OUT_GLOBAL / OUT_M_SMEM
  uint8_t fill;       // debug: A fill id produced
  uint8_t use;        // debug: accumulator use id consumed
  int32_t ldo;        // T_OUT: row pitch of out
  const float* bias;  // zero-padded bias of this n-sub
  float* out;         // T_OUT: fp32 rows
};

struct Params {
  CUtensorMap tm[MAXTM];
  CUtensorMap tm_out;   // fp32 [M rows][ldo columns] output of the OUT_GLOBAL tasks, box 16 columns x 32 rows, no swizzle
  Unit unit[MAXU];
  Task task[MAXTASK];
  int n_units, n_tasks, n_tm;
  int M;                 // rows = T * B
  int B;
  HlIn in[2];
  // where head (modules.py:41-63; cell.py:129-133)
  const float* eps_where;   // [T*B, 4]
  float* where;          // [T*B, 4]
  float* where_loc;
  float* where_scale;
  float max_crop, scale_bias;
  // what head (modules.py:11-24, cell.py:154-156)
  const float* eps_what;
  float* what;
  float* what_loc;
  float* what_scale;
  int na, na_off;
  float what_offset;
  int out_tma;           // the OUT_GLOBAL tasks store through tm_out (TMA tensor store from the warp's staging tile); legal iff
                         // every other user of the staging tiles runs before the first such task (host checks)
  int prefetch_eps;      // the what head's noise may be staged at kernel start (no T_OUT precedes the T_WHAT task)
  int* range_flag;
  long long* trace;      // debug (AIR_ROW_TRACE): [CTA][MAXU + MAXTASK][4] SM-clock stamps, or null
};
#define ROW_TRACE(idx, slot)                                                                                        \
  do {                                                                                                              \
    if (p.trace) p.trace[((size_t)blockIdx.x * (MAXU + MAXTASK) + (idx)) * 4 + (slot)] = clock64();                 \
  } while (0)

struct WhereCode {
  float sx, tx, sy, ty;
};

// where = loc + scale * eps for one row (modules.py:41-63, cell.py:129-133); m = the 8 outputs of the transform estimator.
// row < 0: a padding row of the last tile.  `write`: this thread stores the row's three outputs.
__device__ __noinline__ WhereCode where_task(const float* m, const float* __restrict__ eps_where, float* __restrict__ where,
                                                float* __restrict__ where_loc, float* __restrict__ where_scale,
                                                float max_crop, float scale_bias, int row, bool write) {
  float wv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float mk = m[k];
    const float loc = (k & 1) ? tanhf(mk) : __fmul_rn(max_crop, sigmoid_f(mk));
    const float sc = softplus_f(m[4 + k] + scale_bias);
    const float e = row >= 0 ? __ldg(eps_where + (size_t)row * 4 + k) : 0.f;
    wv[k] = __fadd_rn(__fmul_rn(e, sc), loc);
    if (write && row >= 0) {
      where_loc[(size_t)row * 4 + k] = loc;
      where_scale[(size_t)row * 4 + k] = sc;
      where[(size_t)row * 4 + k] = wv[k];
    }
  }
  return WhereCode{wv[0], wv[1], wv[2], wv[3]};
}

// tile_store / tile_load of chain_tc.cuh split at the staging tile: the register-array half stays inline, the address
// arithmetic and the global accesses are ONE out-of-line copy each (inlined at every call site they cost the epilogues
// their 96-register budget).
__device__ __forceinline__ void stage_put(float* stage, int lane, const float (&v)[16]) {
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4)
    *reinterpret_cast<float4*>(stage + stage_idx(lane, c4)) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
}
__device__ __forceinline__ void stage_get(const float* stage, int lane, float (&v)[16]) {
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 t = *reinterpret_cast<const float4*>(stage + stage_idx(lane, c4));
    v[4 * c4] = t.x; v[4 * c4 + 1] = t.y; v[4 * c4 + 2] = t.z; v[4 * c4 + 3] = t.w;
  }
}
// staged tile (32 rows x 16 columns) -> out[row_w + r, col0 + c]; columns >= n_cols and rows >= n_rows are not written
__device__ __noinline__ void tile_flush(float* stage, int lane, float* __restrict__ out, int ld, int row_w, int col0,
                                        int n_cols, int n_rows) {
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
  } else if (((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0) && ((n_cols & 1) == 0)) {
    // even row pitch (the what head: na = 50 floats = 200 bytes): 8-byte stores, four rows per instruction
    const int c2 = lane & 7;
#pragma unroll 4
    for (int i = 0; i < 8; ++i) {
      const int r = (lane >> 3) + 4 * i;
      const float2 t = *reinterpret_cast<const float2*>(stage + stage_idx(r, c2 >> 1) + 2 * (c2 & 1));
      if (row_w + r < n_rows && col0 + 2 * c2 < n_cols)
        *reinterpret_cast<float2*>(out + (size_t)(row_w + r) * ld + col0 + 2 * c2) = t;
    }
  } else {
    const int cc = lane & 15;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int r = (lane >> 4) + 2 * i;
      const float t = stage[stage_idx(r, cc >> 2) + (cc & 3)];
      if (row_w + r < n_rows && col0 + cc < n_cols) out[(size_t)(row_w + r) * ld + col0 + cc] = t;
    }
  }
  __syncwarp();
}
// One 16-column group of 32 rows through the TMA engine: the warp parks the values in its staging tile as plain [32][16] rows
// and one lane issues a 2-D tensor store (rows / columns outside the tensor are clipped by the map).  The epilogue warps issue
// no global store themselves -- measured with the staged STG path, the stores were 3 k of an output task's 4.4 k clocks.
__device__ __forceinline__ void tma_store_group(const CUtensorMap* tm, float* stage, int lane, const float (&v)[16], int col0,
                                                int row0) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the tile
  __syncwarp();
  float4* st4 = reinterpret_cast<float4*>(stage + lane * 16);
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) st4[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async proxy
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(col0), "r"(row0),
                 "r"(smem_u32(stage))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// the same as an asynchronous copy (cp.async, 4 bytes per lane and row pair; zero-filled outside): no register, no stall;
// tile_fetch_wait() before the tile is read
__device__ __forceinline__ void tile_fetch_async(float* stage, int lane, const float* __restrict__ in, int ld, int row_w,
                                                 int col0, int n_cols, int n_rows) {
  const int cc = lane & 15;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int r = (lane >> 4) + 2 * i;
    const bool ok = row_w + r < n_rows && col0 + cc < n_cols;
    const float* src = ok ? in + (size_t)(row_w + r) * ld + col0 + cc : in;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(stage + stage_idx(r, cc >> 2) + (cc & 3))),
                 "l"(src), "r"(ok ? 4 : 0)
                 : "memory");
  }
}
__device__ __forceinline__ void tile_fetch_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
}
// in[row_w + r, col0 + c] (0 outside) -> staged tile
__device__ __noinline__ void tile_fetch(float* stage, int lane, const float* __restrict__ in, int ld, int row_w, int col0,
                                        int n_cols, int n_rows) {
  const int cc = lane & 15;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int r = (lane >> 4) + 2 * i;
    float t = 0.f;
    if (row_w + r < n_rows && col0 + cc < n_cols) t = __ldg(in + (size_t)(row_w + r) * ld + col0 + cc);
    stage[stage_idx(r, cc >> 2) + (cc & 3)] = t;
  }
  __syncwarp();
}

// one 16-column group of an ELU layer: accumulator -> bias -> ELU -> fp16 hi/lo words of the next layer's A operand
__device__ __forceinline__ void elu_group(float (&v)[16], const float* __restrict__ bias, uint32_t (&hi)[8], uint32_t (&lo)[8],
                                          uint32_t& ovf) {
  // bias four at a time: a 16-register bias array next to v, the packed words of both groups and the loop state does not
  // fit the 96-register budget
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    v[4 * c4] = elu_fast(fmaf(v[4 * c4], W_UNSCALE, b.x));
    v[4 * c4 + 1] = elu_fast(fmaf(v[4 * c4 + 1], W_UNSCALE, b.y));
    v[4 * c4 + 2] = elu_fast(fmaf(v[4 * c4 + 2], W_UNSCALE, b.z));
    v[4 * c4 + 3] = elu_fast(fmaf(v[4 * c4 + 3], W_UNSCALE, b.w));
  }
  split_pack16(v, hi, lo, ovf);
}
// v = v * 2^-8 + bias, bias four at a time
__device__ __forceinline__ void bias_group(float (&v)[16], const float* __restrict__ bias) {
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    v[4 * c4] = fmaf(v[4 * c4], W_UNSCALE, b.x);
    v[4 * c4 + 1] = fmaf(v[4 * c4 + 1], W_UNSCALE, b.y);
    v[4 * c4 + 2] = fmaf(v[4 * c4 + 2], W_UNSCALE, b.z);
    v[4 * c4 + 3] = fmaf(v[4 * c4 + 3], W_UNSCALE, b.w);
  }
}

__global__ void __launch_bounds__(ROW_THREADS, 1) row_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* w_free = w_full + W_SLOTS;
  uint64_t* a_ready = w_free + W_SLOTS;
  uint64_t* a_free = a_ready + 2;
  uint64_t* d_full = a_free + 2;
  uint64_t* d_free = d_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(d_free + 2);
  float* s_stage = reinterpret_cast<float*>(smem + ROW_STAGE_OFFSET);
  float* s_m = s_stage;   // [128][8] where-MLP outputs; only live between the where head's T_OUT and T_WHERE

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < W_SLOTS; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_free[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_ready[i], EPI_THREADS);
      mbar_init(&a_free[i], 1);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_free[i], EPI_THREADS);
    }
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
    // ===== TMA producer: the weight tiles of every unit, in unit order =====
    if (elect_one()) {
      uint32_t ph = 0;   // bit s: parity of the fills of slot s so far
      for (int l = 0; l < p.n_tm; ++l) prefetch_tmap(&p.tm[l]);
      for (int u = 0; u < p.n_units; ++u) {
        const Unit U = p.unit[u];
        const int s = u % W_SLOTS;
        mbar_wait(&w_free[s], ((ph >> s) & 1u) ^ 1u);
        ph ^= 1u << s;
        mbar_expect_tx(&w_full[s], (uint32_t)U.nkb * 2u * (uint32_t)U.box16 * 16u * 128u);
        uint8_t* slot = smem + s * W_SLOT;
        for (int kb = 0; kb < U.nkb; ++kb) {
          tma_load_2d(slot + (kb * 2) * W_TILE, &p.tm[U.tm], (U.kb0 + kb) * BK, U.n_row, &w_full[s]);
          tma_load_2d(slot + (kb * 2 + 1) * W_TILE, &p.tm[U.tm], (U.kb0 + kb) * BK, U.lo_row + U.n_row, &w_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      uint32_t ph_w = 0, ph_a = 0, ph_d = 0;
      for (int u = 0; u < p.n_units; ++u) {
        const Unit U = p.unit[u];
        const int s = u % W_SLOTS;
        ROW_TRACE(u, 0);
        if (U.flags & U_WAIT_A) {
          mbar_wait(&a_ready[U.a_half], (ph_a >> U.a_half) & 1u);
          ph_a ^= 1u << U.a_half;
        }
        if (U.flags & U_WAIT_D) {
          mbar_wait(&d_free[U.d_idx], ((ph_d >> U.d_idx) & 1u) ^ 1u);
          ph_d ^= 1u << U.d_idx;
        }
        ROW_TRACE(u, 1);
        mbar_wait(&w_full[s], (ph_w >> s) & 1u);
        ph_w ^= 1u << s;
        ROW_TRACE(u, 2);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(BM, (int)U.n16 * 16);
        const uint32_t d = tmem_base + d_col(U.d_idx);
        const uint32_t a_hi = tmem_base + a_col(U.a_half), a_lo = a_hi + 64u;
        const uint8_t* slot = smem + s * W_SLOT;
        const bool acc0 = (U.flags & U_ACC0) != 0;
        // cross terms of this unit's K range, then its main terms (see chain_tc.cuh on the accumulator truncation)
        for (int kb = 0; kb < U.nkb; ++kb) {
          const uint64_t db_hi = make_smem_desc_sw128(slot + (kb * 2) * W_TILE);
          const uint64_t db_lo = make_smem_desc_sw128(slot + (kb * 2 + 1) * W_TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int sl = kb * 4 + k;
            if (sl < U.nsl) {
              const uint64_t adv = (uint64_t)(k * 2);
              umma_f16_ts(d, a_lo + (uint32_t)sl * 8u, db_hi + adv, idesc, (acc0 && sl == 0) ? 0u : 1u);
              umma_f16_ts(d, a_hi + (uint32_t)sl * 8u, db_lo + adv, idesc, 1u);
            }
          }
        }
        for (int kb = 0; kb < U.nkb; ++kb) {
          const uint64_t db_hi = make_smem_desc_sw128(slot + (kb * 2) * W_TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int sl = kb * 4 + k;
            if (sl < U.nsl) umma_f16_ts(d, a_hi + (uint32_t)sl * 8u, db_hi + (uint64_t)(k * 2), idesc, 1u);
          }
        }
        umma_commit(&w_free[s]);
        if (U.flags & U_COMMIT_A) umma_commit(&a_free[U.a_half]);
        if (U.flags & U_COMMIT_D) umma_commit(&d_full[U.d_idx]);
        ROW_TRACE(u, 3);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue / operand warps: thread <-> (row = TMEM lane, quarter cq of a 128-column sub-tile) =====
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int row_w = m0 + q * 32;
    const int rit = q * 32 + lane;
    const int row = m0 + rit;
    const bool row_ok = row < p.M;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stage = s_stage + (warp - 2) * 512;
    uint32_t ovf = 0;
    uint32_t ph_a = 0, ph_d = 0;        // a_free (free-style) / d_full (full-style) parities
    // The what head's noise (16 latents x this warp's 32 rows) is fetched into the warp's staging tile right away: nothing
    // else touches the tile before the what head's epilogue (only T_OUT / T_WHAT stage through it, and T_WHAT comes first),
    // so the L2 round trips hide behind the glimpse Encoder instead of stalling the chain.
    if (p.prefetch_eps && cq * 16 < p.na_off) tile_fetch_async(stage, lane, p.eps_what, p.na, row_w, cq * 16, p.na, p.M);

    int type = p.task[0].type;
    for (int ti = 0; ti < p.n_tasks; ++ti) {
      const Task& K = p.task[ti];
      const int type_now = type;
      type = p.task[min(ti + 1, p.n_tasks - 1)].type;   // touches the next descriptor's constant-cache line early
      if (threadIdx.x == 64) ROW_TRACE(MAXU + ti, 0);
      if (type_now == T_LOAD_HL) {
        const int ah = K.a_half;
        // the operand rows do not depend on the A half being free: fetch them while its last reader retires
        const HlIn in = p.in[K.buf];
        uint4 h0[2], h1[2], l0[2], l1[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int ls = 2 * cq + g;
          if (ls < K.nsl) {
            const __half* src = in.p + (((size_t)blockIdx.x * in.nsl + K.s0 + ls) * BM + rit) * 16;
            const uint4* ph4 = reinterpret_cast<const uint4*>(src);
            const uint4* pl4 = reinterpret_cast<const uint4*>(src + in.plane);
            h0[g] = __ldg(ph4); h1[g] = __ldg(ph4 + 1); l0[g] = __ldg(pl4); l1[g] = __ldg(pl4 + 1);
          }
        }
        mbar_wait(&a_free[ah], ((ph_a >> ah) & 1u) ^ 1u);
        ph_a ^= 1u << ah;
        tc_fence_after();
        if (threadIdx.x == 64) ROW_TRACE(MAXU + ti, 1);
        const uint32_t a_hi = t_lane + a_col(ah), a_lo = a_hi + 64u;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int ls = 2 * cq + g;
          if (ls < K.nsl) {
            const uint32_t hi[8] = {h0[g].x, h0[g].y, h0[g].z, h0[g].w, h1[g].x, h1[g].y, h1[g].z, h1[g].w};
            const uint32_t lo[8] = {l0[g].x, l0[g].y, l0[g].z, l0[g].w, l1[g].x, l1[g].y, l1[g].z, l1[g].w};
            tmem_st_32x8(a_hi + ls * 8, hi);
            tmem_st_32x8(a_lo + ls * 8, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a_ready[ah]);
      } else if (type_now == T_WHERE) {
        // where = loc + scale * eps (cell.py:129-133): one thread per row (cq 0) samples and writes the row's code
        named_bar_sync(1, EPI_THREADS);   // s_m complete
        if (cq == 0)
          where_task(s_m + rit * 8, p.eps_where, p.where, p.where_loc, p.where_scale, p.max_crop, p.scale_bias,
                     row_ok ? row : -1, true);
        named_bar_sync(1, EPI_THREADS);   // s_m is free again (it aliases the staging tiles)
      } else if (type_now == T_WHAT) {
        // D columns [0, na_off) = loc, [na_off, 2 na_off) = raw scale: ParametrisedGaussian + sample (modules.py:11-24)
        const int d = K.d_idx, ah = K.a_half, na = p.na;
        const int c0 = cq * 16;
        const bool mine = c0 < p.na_off;
        float e[16];
        // the noise of this thread's latents (coalesced through the staging tile), fetched ahead of the wait
        if (mine) {
          // this thread's two bias lines (location and raw scale), pulled into L1 behind the accumulator wait
          asm volatile("prefetch.global.L1 [%0];" ::"l"(K.bias + c0));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(K.bias + p.na_off + c0));
          if (!p.prefetch_eps) tile_fetch(stage, lane, p.eps_what, na, row_w, c0, na, p.M);
          else tile_fetch_wait();
          stage_get(stage, lane, e);
          __syncwarp();
        }
        mbar_wait(&d_full[d], (ph_d >> d) & 1u);
        ph_d ^= 1u << d;
        tc_fence_after();
        if (threadIdx.x == 64) ROW_TRACE(MAXU + ti, 1);
        const uint32_t dcol = t_lane + d_col(d);
        uint32_t hi[8], lo[8];
        if (mine) {
          {
            float vs[16];
            tmem_ld_32x16(dcol + p.na_off + c0, vs);
            tmem_ld_wait(vs);
            bias_group(vs, K.bias + p.na_off + c0);
#pragma unroll
            for (int j = 0; j < 16; ++j) vs[j] = softplus_lean(vs[j] + p.what_offset);
            if (threadIdx.x == 64) ROW_TRACE(MAXU + ti, 2);
            stage_put(stage, lane, vs);
            tile_flush(stage, lane, p.what_scale, na, row_w, c0, na, p.M);
#pragma unroll
            for (int j = 0; j < 16; ++j) e[j] = __fmul_rn(e[j], vs[j]);
          }
          {
            float vl[16];
            tmem_ld_32x16(dcol + c0, vl);
            tmem_ld_wait(vl);
            bias_group(vl, K.bias + c0);
            stage_put(stage, lane, vl);
            tile_flush(stage, lane, p.what_loc, na, row_w, c0, na, p.M);
#pragma unroll
            for (int j = 0; j < 16; ++j) e[j] = c0 + j < na ? __fadd_rn(e[j], vl[j]) : 0.f;
          }
          stage_put(stage, lane, e);
          tile_flush(stage, lane, p.what, na, row_w, c0, na, p.M);
          split_pack16(e, hi, lo, ovf);
        }
        tc_fence_before();
        mbar_arrive(&d_free[d]);
        mbar_wait(&a_free[ah], ((ph_a >> ah) & 1u) ^ 1u);
        ph_a ^= 1u << ah;
        tc_fence_after();
        if (mine) {
          tmem_st_32x8(t_lane + a_col(ah) + cq * 8, hi);
          tmem_st_32x8(t_lane + a_col(ah) + 64u + cq * 8, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a_ready[ah]);
      } else {
        // ---- T_ELU / T_OUT: this thread's two 16-column groups of the n-sub ----
        const int d = K.d_idx;
        // the L1 is a few KB with this carve-out: pull this thread's bias lines in while the accumulator completes
        asm volatile("prefetch.global.L1 [%0];" ::"l"(K.bias + 32 * cq));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(K.bias + 32 * cq + 16));
        mbar_wait(&d_full[d], (ph_d >> d) & 1u);
        ph_d ^= 1u << d;
        tc_fence_after();
        if (threadIdx.x == 64) ROW_TRACE(MAXU + ti, 1);
        const uint32_t dcol = t_lane + d_col(d);
        const int cA = 32 * cq, cB = cA + 16;
        const bool okA = cA < K.n_valid, okB = cB < K.n_valid;
        if (type_now == T_ELU) {
          // One 16-column group at a time: drain, bias + ELU, split, store as the next layer's operand.  Holding both
          // groups' packed words until the A half is free costs 32 registers the 96-register budget does not have (every
          // spill is an L2 round trip with this shared-memory carve-out); the half is normally free by the time the first
          // group has been computed (its last reader is one unit behind the unit that completed this accumulator).
          const int ah = K.a_half;
          const uint32_t a_hi = t_lane + a_col(ah), a_lo = a_hi + 64u;
          uint32_t hi[8], lo[8];
          if (okA) {
            float v[16];
            tmem_ld_32x16(dcol + cA, v);
            tmem_ld_wait(v);
            elu_group(v, K.bias + cA, hi, lo, ovf);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) hi[j] = lo[j] = 0u;
          }
          if (threadIdx.x == 64) ROW_TRACE(MAXU + ti, 2);
          mbar_wait(&a_free[ah], ((ph_a >> ah) & 1u) ^ 1u);
          ph_a ^= 1u << ah;
          tc_fence_after();
          tmem_st_32x8(a_hi + (2 * cq) * 8, hi);
          tmem_st_32x8(a_lo + (2 * cq) * 8, lo);
          if (okB) {
            float v[16];
            tmem_ld_32x16(dcol + cB, v);
            tmem_ld_wait(v);
            tc_fence_before();
            mbar_arrive(&d_free[d]);
            elu_group(v, K.bias + cB, hi, lo, ovf);
          } else {
            tc_fence_before();
            mbar_arrive(&d_free[d]);
#pragma unroll
            for (int j = 0; j < 8; ++j) hi[j] = lo[j] = 0u;
          }
          tmem_st_32x8(a_hi + (2 * cq + 1) * 8, hi);
          tmem_st_32x8(a_lo + (2 * cq + 1) * 8, lo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&a_ready[ah]);
        } else {
          float v[16];
          if (okA) {
            tmem_ld_32x16(dcol + cA, v);
            tmem_ld_wait(v);
            bias_group(v, K.bias + cA);
            if (K.out_kind == OUT_M_SMEM) {
              if (cq == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) s_m[rit * 8 + j] = v[j];
              }
            } else if (p.out_tma) {
              tma_store_group(&p.tm_out, stage, lane, v, K.s0 + cA, row_w);
            } else {
              stage_put(stage, lane, v);
              tile_flush(stage, lane, K.out, K.ldo, row_w, K.s0 + cA, K.s0 + K.n_valid, p.M);
            }
          }
          if (okB && K.out_kind != OUT_M_SMEM) {
            tmem_ld_32x16(dcol + cB, v);
            tmem_ld_wait(v);
            tc_fence_before();
            mbar_arrive(&d_free[d]);
            bias_group(v, K.bias + cB);
            if (p.out_tma) {
              tma_store_group(&p.tm_out, stage, lane, v, K.s0 + cB, row_w);
            } else {
              stage_put(stage, lane, v);
              tile_flush(stage, lane, K.out, K.ldo, row_w, K.s0 + cB, K.s0 + K.n_valid, p.M);
            }
          } else {
            tc_fence_before();
            mbar_arrive(&d_free[d]);
          }
        }
      }
      if (threadIdx.x == 64) ROW_TRACE(MAXU + ti, 3);
    }
    if ((ovf & 0x80008000u) && p.range_flag) atomicOr(p.range_flag, 1);
    // the staging tiles must outlive the TMA engine's reads of them
    if (p.out_tma && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---- host side: schedule construction ------------------------------------------------------------------------------
struct LayerDesc {
  int K = 0;            // contraction length
  int N = 0;            // output columns to compute (what head: 2 * na_off)
  int lo_row = 0;       // N_alloc of the prepared weight
  int tm = 0;           // tensor map index
  int box_rows = 128;   // rows of that map's box
  int epi = T_ELU;      // T_ELU / T_OUT / T_WHAT
  int a_src = 0;        // 0: operand left by the previous layer; 1: T_LOAD_HL from in[a_buf]
  int a_buf = 0;
  int out_kind = OUT_GLOBAL;
  const float* bias = nullptr;
  float* out = nullptr;
  int ldo = 0;
  bool where_after = false;   // the where head's output layer: a T_WHERE task follows, after the NEXT layer's operand loads
                              // have been issued (so the next MMAs run while the where code is sampled)
};
struct Schedule {
  std::vector<Unit> units;
  std::vector<Task> tasks;
  std::string error;
};

// Units and tasks of a layer list.  Returns false (with s.error set) when the list does not fit the machine.
inline bool build_schedule(const std::vector<LayerDesc>& layers, Schedule& s) {
  s.units.clear();
  s.tasks.clear();
  s.error.clear();
  int use_counter = 0;                  // accumulator uses so far (d_idx alternates)
  int fill_counter = 0;                 // A fills so far
  int cur_fill[2] = {-1, -1};           // fill currently held by each A half
  int fill_slices[2] = {0, 0};          // K slices that fill holds
  bool fill_waited[256] = {};           // a unit already carries U_WAIT_A for this fill
  int fill_last_reader[256];
  for (int i = 0; i < 256; ++i) fill_last_reader[i] = -1;
  auto new_fill = [&](int half, int nsl) {
    cur_fill[half] = fill_counter++;
    fill_slices[half] = nsl;
    return cur_fill[half];
  };
  bool pending_where = false;
  for (size_t li = 0; li < layers.size(); ++li) {
    const LayerDesc& L = layers[li];
    const int nsl_total = (L.K + 15) / 16;
    const int nparts = (nsl_total + 7) / 8;
    const int nsubs = (L.N + 127) / 128;
    if (fill_counter + nparts + nsubs >= 250) { s.error = "too many A fills"; return false; }
    if (L.epi != T_OUT && nsubs > 2) { s.error = "a hidden layer is wider than 256"; return false; }
    if (L.epi == T_WHAT && nsubs != 1) { s.error = "what head wider than 128 columns"; return false; }
    int part_fill[64];
    if (nparts > 64) { s.error = "K too long"; return false; }
    if (L.a_src == 0) {
      if (nparts > 2) { s.error = "resident operand longer than 256"; return false; }
      for (int pt = 0; pt < nparts; ++pt) {
        const int need = std::min(8, nsl_total - 8 * pt);
        if (cur_fill[pt] < 0 || fill_slices[pt] < need) { s.error = "operand half not produced by the previous layer"; return false; }
        part_fill[pt] = cur_fill[pt];
      }
    } else {
      for (int pt = 0; pt < nparts; ++pt) {
        Task t = {};
        t.type = (uint8_t)T_LOAD_HL;
        t.a_half = (uint8_t)(pt & 1);
        t.nsl = (uint8_t)std::min(8, nsl_total - 8 * pt);
        t.s0 = (uint16_t)(8 * pt);
        t.buf = (uint8_t)L.a_buf;
        part_fill[pt] = new_fill(pt & 1, t.nsl);
        t.fill = (uint8_t)part_fill[pt];
        if (pending_where && pt == 2) {   // later loads wait for this layer's MMAs
          Task tw = {};
          tw.type = T_WHERE;
          s.tasks.push_back(tw);
          pending_where = false;
        }
        s.tasks.push_back(t);
      }
    }
    if (pending_where) {
      Task tw = {};
      tw.type = T_WHERE;
      s.tasks.push_back(tw);
      pending_where = false;
    }
    const bool p_outer = nparts > 2;
    if (p_outer && nsubs > 2) { s.error = "K > 256 with N > 256 in one layer"; return false; }
    const int use0 = use_counter;
    use_counter += nsubs;
    const size_t unit0 = s.units.size();
    auto add_unit = [&](int j, int pt) {
      Unit u = {};
      u.n_row = (uint16_t)(128 * j);
      u.lo_row = (uint16_t)L.lo_row;
      u.kb0 = (uint16_t)(2 * pt);
      u.tm = (uint8_t)L.tm;
      u.box16 = (uint8_t)(L.box_rows / 16);
      u.nsl = (uint8_t)std::min(8, nsl_total - 8 * pt);
      u.nkb = (uint8_t)((u.nsl + 3) / 4);
      u.n16 = (uint8_t)((std::min(128, L.N - 128 * j) + 15) / 16);
      u.a_half = (uint8_t)(pt & 1);
      u.d_idx = (uint8_t)((use0 + j) & 1);
      u.use = (uint8_t)(use0 + j);
      u.fill = (uint8_t)part_fill[pt];
      u.flags = 0;
      if (pt == 0) u.flags |= U_WAIT_D | U_ACC0;
      if (pt == nparts - 1) u.flags |= U_COMMIT_D;
      if (!fill_waited[u.fill]) {
        u.flags |= U_WAIT_A;
        fill_waited[u.fill] = true;
      }
      fill_last_reader[u.fill] = (int)s.units.size();
      s.units.push_back(u);
    };
    if (p_outer) {
      for (int pt = 0; pt < nparts; ++pt)
        for (int j = 0; j < nsubs; ++j) add_unit(j, pt);
    } else {
      for (int j = 0; j < nsubs; ++j)
        for (int pt = 0; pt < nparts; ++pt) add_unit(j, pt);
    }
    for (int pt = 0; pt < nparts; ++pt) s.units[fill_last_reader[part_fill[pt]]].flags |= U_COMMIT_A;
    (void)unit0;
    // epilogue tasks, one per n-sub
    for (int j = 0; j < nsubs; ++j) {
      Task t = {};
      t.type = (uint8_t)L.epi;
      t.d_idx = (uint8_t)((use0 + j) & 1);
      t.use = (uint8_t)(use0 + j);
      t.s0 = (uint16_t)(128 * j);
      t.n_valid = (uint16_t)std::min(128, L.N - 128 * j);
      t.bias = L.bias ? L.bias + 128 * j : nullptr;
      t.out = L.out;
      t.ldo = L.ldo;
      t.out_kind = (uint8_t)L.out_kind;
      if (L.epi == T_ELU || L.epi == T_WHAT) {
        t.a_half = (uint8_t)j;
        t.fill = (uint8_t)new_fill(j, (t.n_valid + 15) / 16);
      }
      s.tasks.push_back(t);
    }
    if (L.where_after) pending_where = true;
  }
  if (pending_where) {
    Task tw = {};
    tw.type = T_WHERE;
    s.tasks.push_back(tw);
  }
  if ((int)s.units.size() > MAXU) { s.error = "too many units"; return false; }
  if ((int)s.tasks.size() > MAXTASK) { s.error = "too many tasks"; return false; }
  return true;
}

// Replays the producer / MMA / epilogue programs against mbarrier semantics on the host.  Checks: no deadlock, every
// unit reads the A fill and accumulator use the builder meant it to read, no A half or accumulator is overwritten while
// a unit that reads it is still in flight, every accumulator is drained exactly once.  Returns "" or a description.
inline std::string simulate(const Schedule& s) {
  const int nu = (int)s.units.size(), nt = (int)s.tasks.size();
  // completion counts of each barrier ("phases completed")
  int w_full[W_SLOTS] = {}, w_free[W_SLOTS] = {}, a_ready[2] = {}, a_free[2] = {}, d_full[2] = {}, d_free[2] = {};
  // role-local wait counters
  int pw_free[W_SLOTS] = {};                       // producer: waits done on w_free[s]
  int mw_full[W_SLOTS] = {}, ma_ready[2] = {}, md_free[2] = {};
  int ea_free[2] = {}, ed_full[2] = {};
  int a_fill[2] = {-1, -1}, d_use[2] = {-1, -1};   // contents
  bool d_drained[2] = {true, true};
  int pu = 0, mu = 0, et = 0;                      // program counters
  int e_stage = 0;                                 // epilogue sub-step within a task
  // MMAs complete in order, "some time" after issue: model completion as immediate at issue (optimistic) AND verify the
  // ordering hazards pessimistically through the barrier protocol itself: a writer may only touch A / D after the
  // corresponding *_free completion that the last reader's commit produces.
  std::vector<int> unit_of_fill_last(256, -1);
  for (int u = 0; u < nu; ++u) unit_of_fill_last[s.units[u].fill] = u;
  int guard = 0;
  while ((pu < nu || mu < nu || et < nt) && guard++ < 100000) {
    bool progress = false;
    // producer: free-style wait #k passes when completions >= k (k = waits done so far)
    if (pu < nu) {
      const int sl = pu % W_SLOTS;
      if (w_free[sl] >= pw_free[sl]) {
        ++pw_free[sl];
        ++w_full[sl];   // TMA lands
        ++pu;
        progress = true;
      }
    }
    if (mu < nu) {
      const Unit& U = s.units[mu];
      const int sl = mu % W_SLOTS;
      bool ok = true;
      if ((U.flags & U_WAIT_A) && a_ready[U.a_half] < ma_ready[U.a_half] + 1) ok = false;
      if ((U.flags & U_WAIT_D) && d_free[U.d_idx] < md_free[U.d_idx]) ok = false;
      if (w_full[sl] < mw_full[sl] + 1) ok = false;
      if (ok) {
        if (U.flags & U_WAIT_A) ++ma_ready[U.a_half];
        if (U.flags & U_WAIT_D) {
          ++md_free[U.d_idx];
          if (!d_drained[U.d_idx]) return "unit " + std::to_string(mu) + " overwrites an undrained accumulator";
          d_use[U.d_idx] = U.use;
          d_drained[U.d_idx] = false;
        }
        ++mw_full[sl];
        if (a_fill[U.a_half] != U.fill)
          return "unit " + std::to_string(mu) + " reads A half " + std::to_string(U.a_half) + " holding fill " +
                 std::to_string(a_fill[U.a_half]) + ", expected " + std::to_string(U.fill);
        if (d_use[U.d_idx] != U.use) return "unit " + std::to_string(mu) + " accumulates into the wrong use";
        if (((U.flags & U_ACC0) != 0) != ((U.flags & U_WAIT_D) != 0)) return "ACC0 / WAIT_D mismatch";
        ++w_free[sl];
        if (U.flags & U_COMMIT_A) ++a_free[U.a_half];
        if (U.flags & U_COMMIT_D) ++d_full[U.d_idx];
        ++mu;
        progress = true;
      }
    }
    if (et < nt) {
      const Task& K = s.tasks[et];
      if (K.type == T_WHERE) {
        ++et;
        progress = true;
      } else if (K.type == T_LOAD_HL) {
        if (a_free[K.a_half] >= ea_free[K.a_half]) {
          ++ea_free[K.a_half];
          // the previous fill of this half must have been fully consumed: its last reader has been issued
          if (a_fill[K.a_half] >= 0 && unit_of_fill_last[a_fill[K.a_half]] >= mu)
            return "task " + std::to_string(et) + " overwrites A half " + std::to_string(K.a_half) + " before its last reader";
          a_fill[K.a_half] = K.fill;
          ++a_ready[K.a_half];
          ++et;
          progress = true;
        }
      } else {
        if (e_stage == 0) {
          if (d_full[K.d_idx] >= ed_full[K.d_idx] + 1) {
            ++ed_full[K.d_idx];
            if (d_use[K.d_idx] != K.use) return "task " + std::to_string(et) + " drains the wrong accumulator use";
            if (d_drained[K.d_idx]) return "task " + std::to_string(et) + " drains an accumulator twice";
            if (K.type != T_ELU) {   // T_ELU releases the accumulator only after its A half has become free (second stage)
              d_drained[K.d_idx] = true;
              ++d_free[K.d_idx];
            }
            if (K.type == T_OUT) {
              ++et;
            } else {
              e_stage = 1;
            }
            progress = true;
          }
        } else {
          if (a_free[K.a_half] >= ea_free[K.a_half]) {
            ++ea_free[K.a_half];
            if (a_fill[K.a_half] >= 0 && unit_of_fill_last[a_fill[K.a_half]] >= mu)
              return "task " + std::to_string(et) + " overwrites A half " + std::to_string(K.a_half) + " before its last reader";
            a_fill[K.a_half] = K.fill;
            if (K.type == T_ELU) {
              d_drained[K.d_idx] = true;
              ++d_free[K.d_idx];
            }
            ++a_ready[K.a_half];
            e_stage = 0;
            ++et;
            progress = true;
          }
        }
      }
    }
    if (!progress)
      return "deadlock at producer " + std::to_string(pu) + " / mma " + std::to_string(mu) + " / task " + std::to_string(et);
  }
  if (pu < nu || mu < nu || et < nt) return "did not terminate";
  for (int i = 0; i < 2; ++i)
    if (!d_drained[i]) return "an accumulator was never drained";
  return "";
}

// tensor map over a prepared weight (W^T hl planes, [2 * n_alloc][kpad] fp16) with a box of (up to) 128 rows x 64 K
inline int row_box_rows(int n_alloc) { return 2 * n_alloc < 128 ? 2 * n_alloc : 128; }
inline bool make_row_weight_tmap(CUtensorMap* tm, const __half* base, int kpad, int n_alloc) {
  return make_tmap(tm, base, kpad, 2 * (int64_t)n_alloc, row_box_rows(n_alloc));
}

// fp32 [rows][ld] output tensor, box 16 columns x 32 rows (one warp's staging tile), no swizzle
inline bool make_row_out_tmap(CUtensorMap* tm, float* base, int ld, long long rows) {
  air::tc::EncodeTiledFn fn = air::tc::get_encode_fn();
  if (!fn || (ld % 4) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {16, 32};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline cudaError_t launch_row(const Params& p, cudaStream_t st) {
  cudaError_t e = ensure_dynamic_smem(row_kernel, ROW_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_k(row_kernel, dim3((p.M + BM - 1) / BM), dim3(ROW_THREADS), ROW_SMEM_BYTES, st, p);
}

}  // namespace row
}  // namespace air
