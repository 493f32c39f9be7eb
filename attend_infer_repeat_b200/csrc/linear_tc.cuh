// Tensor-core dense layer for sm_100a: tcgen05.mma (kind::f16) with TMEM accumulators, TMA-fed, warp-specialised.
//
//   out[M,N] = act((A @ W) + bias (+ addend))          A[M,K], W[K,N] given in fp32
//
// fp32-class accuracy on the fp16 tensor pipe ("fp16x2 split"): every fp32 operand x is carried as two halves
// x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 significand bits), and each K-block issues three MMAs:
//   acc0 += A_hi*W_hi        acc1 += A_lo*W_hi + A_hi*W_lo        (the lo*lo term is below 2^-22 relative)
// fp16 x fp16 products are exact in fp32.  The tensor core truncates (round-toward-zero) its fp32 accumulator once
// per MMA, a bias proportional to the accumulator magnitude and to K/16; keeping the ~2^-11-sized cross terms in
// their own accumulator leaves one truncation per K=16 slice on the main one (measured: same error as an fp32 FMA
// chain for K <= 512, ~1e-5 absolute on O(1) outputs at K = 2500).  Weights are pre-scaled by 2^8 (exact) so that
// their lo halves stay in the fp16 normal range; the epilogue multiplies by 2^-8.
//
// Operand format in HBM ("hl" buffers): fp16 [2][rows_alloc][Kpad] -- the hi plane followed by the lo plane, K
// contiguous, Kpad a multiple of 64 with zero padding.  Producers (this kernel's epilogue, the glimpse-read kernel,
// the LSTM gate kernel, ...) write the next layer's A operand directly in that format, so activations cross HBM at
// 4 bytes per element like fp32 would.  Weights are transposed/split once per parameter update into W^T [2][N_alloc][Kpad].
//
// Kernel shape: one 128 x BN output tile per CTA, BK = 64 (one 128-byte swizzle atom of fp16), STAGES-deep TMA ring,
// sized so that two CTAs share an SM (one's epilogue overlaps the other's main loop).
//   warp 0 : TMA producer (one elected lane): 4 tile loads per stage (A_hi, A_lo, W_hi, W_lo) -> full[stage]
//   warp 1 : TMEM alloc/dealloc + MMA issuer (one elected lane): 12 tcgen05.mma per stage, tcgen05.commit -> empty[stage]
//   warps 2.. : epilogue, one warp per (32 TMEM lanes x 32 columns): tcgen05.ld, bias (staged in smem) / addend /
//               activation, fp32 and/or hi/lo fp16 vector stores
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "cell_kernels.cuh"   // mbarrier helpers
#include "common.cuh"

namespace air {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr float W_SCALE = 256.0f;        // weights are stored as fp16 split of (w * 2^8)
constexpr float W_UNSCALE = 1.0f / 256.0f;

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr int num_threads(int BN) { return 64 + 128 * (BN / 32); }

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c_inner, int c_outer, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns of the accumulator -> 16 registers per thread (thread i <-> TMEM lane base+i).
// Asynchronous: tmem_ld_wait() before the registers are used.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same wait, tied to the destination registers of the load so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), LBO (ignored for swizzled K-major, canonical value 1) in [16,30), SBO = 1024 B (8 rows x 128 B) >> 4 in
// [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B (= 2) in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(const void* smem_tile) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16, A = B = F16 (format 0), D = F32 (c_format 1), both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct GemmParams {
  const float* bias;     // [N] or null
  const float* addend;   // fp32 [M, ldadd] added before the activation (gx of the LSTM) or null
  int ldadd;
  float* out_f32;        // [M, ldc] or null
  int ldc;
  __half* out_hl;        // hl buffer of the consumer: hi plane at out_hl, lo plane at out_hl + hl_plane; or null
  size_t hl_plane;       // elements between the hi and the lo plane (rows_alloc * ld_hl)
  int ld_hl;             // Kpad of the consumer
  int hl_nsl;            // > 0: out_hl is slice-major tiled, [row / 128][hl_nsl slices][128 rows][16] per plane (chain_tc.cuh)
  int M, N;
  int num_k_blocks;      // Kpad / 64
  int a_lo_row;          // row coordinate of the lo plane in the A tensor map (= rows_alloc of A)
  int a_row0;            // first row of this GEMM inside the A buffer (LSTM step t reads rows of step t-1)
  int b_lo_row;          // row coordinate of the lo plane in the W^T tensor map (= N_alloc)
  int act;
  int* range_flag;       // set to 1 if a produced activation overflows fp16
  // ---- backward-pass extensions (all zero = the forward behaviour) ----
  float out_scale;       // factor applied to the accumulator; 0 selects W_UNSCALE (operand B prepared with W_SCALE)
  int kb_per_z;          // > 0: the K loop is split over gridDim.z, kb_per_z blocks of 64 per slice (every slice non-empty)
  int atomic_out;        // out_f32 += result with atomicAdd (split-K weight gradients into the zeroed gradient buffer)
  const float* mask_y;   // [M, ld_mask] forward ELU output: result *= (y > 0 ? 1 : y + 1)   (tf.nn.elu gradient)
  int ld_mask;
  int hl_bf16;           // out_hl receives bf16 hi/lo planes (the next gradient GEMM's A operand) instead of fp16 ones
  int ab_bf16;           // both operands hold bf16 hi/lo planes (fp32 exponent range, 16 significant bits) instead of fp16
  long long* trace;      // AIR_TC_TRACE builds only: [n_ctas][16] SM-clock timestamps of the pipeline phases
};

#ifdef AIR_TC_TRACE
#define TC_TRACE(slot)                                                                                     \
  do {                                                                                                     \
    if (p.trace) p.trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64();         \
  } while (0)
#else
#define TC_TRACE(slot) do {} while (0)
#endif

template <int BN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16;
  static constexpr int TOTAL = BIAS_OFFSET + BN * 4 + 1024 /* alignment slack */;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(num_threads(BN), 2)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, GemmParams p) {
  using L = Smem<BN, STAGES>;
  constexpr int N_EPI = 128 * (BN / 32);
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb0 = p.kb_per_z > 0 ? (int)blockIdx.z * p.kb_per_z : 0;
  const int nkb = p.kb_per_z > 0 ? min(p.kb_per_z, p.num_k_blocks - kb0) : p.num_k_blocks;
  if (threadIdx.x == 0) TC_TRACE(0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) TC_TRACE(1);
  // PDL: everything above overlapped the previous kernel's tail; operands below come from earlier kernels
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (kb == 0) TC_TRACE(2);
        uint8_t* st = smem + s * L::STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        const int k0 = (kb0 + kb) * BK;
        tma_load_2d(st, &tm_a, k0, p.a_row0 + m0, &full_bar[s]);
        tma_load_2d(st + L::A_BYTES, &tm_a, k0, p.a_lo_row + p.a_row0 + m0, &full_bar[s]);
        tma_load_2d(st + 2 * L::A_BYTES, &tm_b, k0, n0, &full_bar[s]);
        tma_load_2d(st + 2 * L::A_BYTES + L::B_BYTES, &tm_b, k0, p.b_lo_row + n0, &full_bar[s]);
        if (kb == nkb - 1) TC_TRACE(3);
      }
    }
    __syncwarp();   // reconverge before the CTA-wide barrier below (bar.sync counts warps, not lanes)
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      // instruction descriptor bits [7,10) / [10,13): A / B element format, 0 = f16, 1 = bf16 (kind::f16 wants both alike:
      // a mixed f16 x bf16 descriptor traps as an illegal instruction on sm_100a); bf16 x bf16 products are exact in fp32
      const uint32_t idesc = make_idesc_f16(BM, BN) | (p.ab_bf16 ? ((1u << 7) | (1u << 10)) : 0u);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        if (kb == 0) TC_TRACE(4);
        if (kb == nkb - 1) TC_TRACE(5);
        tc_fence_after();
        uint8_t* st = smem + s * L::STAGE_BYTES;
        const uint64_t da_hi = make_smem_desc_sw128(st);
        const uint64_t da_lo = make_smem_desc_sw128(st + L::A_BYTES);
        const uint64_t db_hi = make_smem_desc_sw128(st + 2 * L::A_BYTES);
        const uint64_t db_lo = make_smem_desc_sw128(st + 2 * L::A_BYTES + L::B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adv = (uint64_t)(k * 16 * 2 >> 4);   // 16 fp16 = 32 bytes along K inside the swizzle atom
          umma_f16(tmem_base, da_hi + adv, db_hi + adv, idesc, (kb | k) != 0);        // main term  -> columns [0, BN)
          umma_f16(tmem_base + BN, da_lo + adv, db_hi + adv, idesc, (kb | k) != 0);   // cross terms -> columns [BN, 2BN)
          umma_f16(tmem_base + BN, da_hi + adv, db_lo + adv, idesc, 1);
        }
        umma_commit(&empty_bar[s]);          // frees the smem stage when these MMAs have read it
      }
      umma_commit(tmem_full_bar);            // accumulator complete
      TC_TRACE(6);
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w owns TMEM lanes 32*(w%4).. (hardware restriction) and column chunk (w-2)/4 =====
    const int et = threadIdx.x - 64;
    const int lane_grp = warp & 3;
    const int chunk = (warp - 2) >> 2;
    const int row = m0 + lane_grp * 32 + lane;
    const int nb = n0 + chunk * 32;
    const bool row_ok = row < p.M;
    // stage the tile's bias while the main loop runs
    for (int j = et; j < BN; j += N_EPI) s_bias[j] = (p.bias && n0 + j < p.N) ? p.bias[n0 + j] : 0.f;
    named_bar_sync(1, N_EPI);
    // prefetch the addend (gx rows of the LSTM) for this thread's 32 columns
    float4 add4[8];
    const bool add_vec = p.addend && row_ok && ((p.ldadd & 3) == 0) && nb + 32 <= p.N &&
                         ((reinterpret_cast<uintptr_t>(p.addend) & 15) == 0);
    if (add_vec) {
      const float4* src = reinterpret_cast<const float4*>(p.addend + (size_t)row * p.ldadd + nb);
#pragma unroll
      for (int j = 0; j < 8; ++j) add4[j] = __ldg(src + j);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) add4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.addend && row_ok) {
        float* a = reinterpret_cast<float*>(add4);
        for (int j = 0; j < 32; ++j)
          if (nb + j < p.N) a[j] = p.addend[(size_t)row * p.ldadd + nb + j];
      }
    }
    const float* addf = reinterpret_cast<const float*>(add4);

    mbar_wait(tmem_full_bar, 0);
    if (threadIdx.x == 64) TC_TRACE(7);
    tc_fence_after();
    const bool f32_vec = p.out_f32 && !p.atomic_out && ((p.ldc & 3) == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0);
    const float osc = p.out_scale != 0.f ? p.out_scale : W_UNSCALE;
    const bool mask_vec = p.mask_y && ((p.ld_mask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask_y) & 15) == 0);
    float amax = 0.f;   // NaN-propagating running max |x| of what is written as fp16 hi halves
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int c0 = chunk * 32 + hf * 16;   // column inside the tile
      float v[16], vx[16];
      tmem_ld_32x16(tmem_base + t_lane + (uint32_t)c0, v);
      tmem_ld_32x16(tmem_base + t_lane + (uint32_t)(BN + c0), vx);
      tmem_ld_wait();
      if (threadIdx.x == 64) TC_TRACE(10 + 3 * hf);
      if (row_ok) {
        // warp-uniform variants keep the per-element instruction count down (the epilogue is issue-bound)
        if (p.act == ACT_ELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = (v[j] + vx[j]) * osc + s_bias[c0 + j] + addf[hf * 16 + j];
            v[j] = x > 0.f ? x : __expf(x) - 1.0f;   // ex2.approx path: <= 2.4e-7 absolute on (-1, 0]
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (v[j] + vx[j]) * osc + s_bias[c0 + j] + addf[hf * 16 + j];
        }
        if (p.mask_y) {   // backward: gradient through the ELU that produced the saved activation y
          const float* my = p.mask_y + (size_t)row * p.ld_mask + n0 + c0;
          if (mask_vec && n0 + c0 + 16 <= p.N) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 y = __ldg(reinterpret_cast<const float4*>(my + j));
              v[j] *= (y.x > 0.f) ? 1.0f : y.x + 1.0f;
              v[j + 1] *= (y.y > 0.f) ? 1.0f : y.y + 1.0f;
              v[j + 2] *= (y.z > 0.f) ? 1.0f : y.z + 1.0f;
              v[j + 3] *= (y.w > 0.f) ? 1.0f : y.w + 1.0f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n0 + c0 + j < p.N) {
                const float y = my[j];
                v[j] *= (y > 0.f) ? 1.0f : y + 1.0f;
              }
            }
          }
        }
        if (n0 + c0 + 16 > p.N) {   // only the last, partial column tile
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j >= p.N) v[j] = 0.f;
        }
        if (threadIdx.x == 64) TC_TRACE(11 + 3 * hf);
        if (p.out_f32) {
          float* dst = p.out_f32 + (size_t)row * p.ldc + n0 + c0;
          if (p.atomic_out) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < p.N) atomicAdd(dst + j, v[j]);
          } else if (f32_vec && n0 + c0 + 16 <= p.N) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < p.N) dst[j] = v[j];
          }
        }
        if (p.out_hl && n0 + c0 < (p.hl_nsl ? p.hl_nsl * 16 : p.ld_hl)) {
          __align__(16) __half hi[16], lo[16];
          if (p.hl_bf16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __nv_bfloat16 bh = __float2bfloat16_rn(v[j]);
              const __nv_bfloat16 bl = __float2bfloat16_rn(v[j] - __bfloat162float(bh));
              hi[j] = __ushort_as_half(__bfloat16_as_ushort(bh));
              lo[j] = __ushort_as_half(__bfloat16_as_ushort(bl));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              split_f16(v[j], hi[j], lo[j]);
              amax = fmax_nan(amax, fabsf(v[j]));
            }
          }
          __half* dh = p.hl_nsl ? p.out_hl + ((((size_t)row >> 7) * p.hl_nsl + ((n0 + c0) >> 4)) * 128 + (row & 127)) * 16
                                : p.out_hl + (size_t)row * p.ld_hl + n0 + c0;
          __half* dl = dh + p.hl_plane;
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            *reinterpret_cast<uint4*>(dh + j) = *reinterpret_cast<const uint4*>(hi + j);
            *reinterpret_cast<uint4*>(dl + j) = *reinterpret_cast<const uint4*>(lo + j);
          }
        }
      }
    }
    if (!(amax <= 65504.f) && p.range_flag) atomicOr(p.range_flag, 1);
    if (threadIdx.x == 64) TC_TRACE(8);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_TRACE(9);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ---- operand preparation -----------------------------------------------------------------------------------------
// fp32 rows -> hl buffer (hi plane, lo plane), zero-padded K is left untouched (buffers are zeroed at creation).
__global__ void split_rows_kernel(const float* __restrict__ src, int ld_src, __half* __restrict__ dst, size_t plane,
                                  int ld_dst, int M, int K, int* range_flag) {
  griddep_launch();
  griddep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 4 elements
  const int kq = (K + 3) / 4;
  if (idx >= (size_t)M * kq) return;
  const int row = (int)(idx / kq), k = (int)(idx % kq) * 4;
  const float* s = src + (size_t)row * ld_src + k;
  __half* dh = dst + (size_t)row * ld_dst + k;
  bool overflow = false;
  if (k + 3 < K && (ld_src & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4 f = *reinterpret_cast<const float4*>(s);
    __align__(8) __half hi[4], lo[4];
    split_f16(f.x, hi[0], lo[0]);
    split_f16(f.y, hi[1], lo[1]);
    split_f16(f.z, hi[2], lo[2]);
    split_f16(f.w, hi[3], lo[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) overflow |= __hisinf(hi[j]) || __hisnan(hi[j]);
    *reinterpret_cast<uint2*>(dh) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(dh + plane) = *reinterpret_cast<const uint2*>(lo);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k + j < K) {
        __half hi, lo;
        split_f16(s[j], hi, lo);
        overflow |= __hisinf(hi) || __hisnan(hi);
        dh[j] = hi;
        dh[plane + j] = lo;
      }
    }
  }
  if (overflow && range_flag) atomicOr(range_flag, 1);
}

// fp32 [M, C] rows (row pitch ld) -> hl planes of the TRANSPOSE, [2][C_alloc][ld_dst] with M contiguous: the operands of a
// weight-gradient GEMM dW = X^T @ dY, whose contraction runs over the batch rows.  `scale` (a power of two) lifts small
// gradients into the fp16 normal range; columns m in [M, ld_dst) are zero-filled so that the contraction padding is
// exact whatever the buffer held before.  One 32 x 32 tile per CTA through shared memory.
// as_bf16: the planes hold bf16 hi / lo (x = hi + lo to 16 significant bits, fp32 exponent range) -- per-sample gradients
// span too many decades for fp16 (1 / s_x factors of the inverse transformer).
constexpr int ST_M = 64;    // batch rows per CTA of split_transpose_kernel
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ src, int ld, int M, int C, float scale, __half* __restrict__ dst,
                       size_t plane, int ld_dst, int* range_flag, int as_bf16, __half* __restrict__ rm, size_t plane_rm,
                       int ld_rm, float* __restrict__ colsum) {
  // optional by-products of the same read (gradient tensors dY): `rm` = the row-major bf16 planes [2][..][ld_rm] with
  // columns [C, ld_rm) zero-filled (A operand of dX = dY @ W^T; the grid must then cover ld_rm columns), `colsum` += the
  // column sums (bias gradient, fp32 atomics).
  // A 64 (rows) x 32 (columns) tile per CTA: every thread stores PAIRS of neighbouring rows (4 bytes, 128 per warp) into the
  // transposed planes (ld_dst and `plane` are even, m0 is a multiple of 64).
  __shared__ float tile[ST_M][33];
  griddep_launch();
  griddep_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int c0 = blockIdx.x * 32, m0 = blockIdx.y * ST_M;
#pragma unroll
  for (int i = 0; i < ST_M; i += 8) {
    const int m = m0 + ty + i, c = c0 + tx;
    const float x = (m < M && c < C) ? src[(size_t)m * ld + c] * scale : 0.f;
    tile[ty + i][tx] = x;
    if (rm && m < M && c < ld_rm) {
      const __nv_bfloat16 bh = __float2bfloat16_rn(x);
      const __nv_bfloat16 bl = __float2bfloat16_rn(x - __bfloat162float(bh));
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(rm) + (size_t)m * ld_rm + c;
      d[0] = bh;
      d[plane_rm] = bl;
    }
  }
  __syncthreads();
  if (colsum && ty == 0 && c0 + tx < C) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < ST_M; ++i) sum += tile[i][tx];
    atomicAdd(colsum + c0 + tx, sum);
  }
  bool overflow = false;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, m = m0 + 2 * tx;
    if (c < C && m < ld_dst) {
      const float x0 = tile[2 * tx][ty + i], x1 = tile[2 * tx + 1][ty + i];
      __half* d = dst + (size_t)c * ld_dst + m;
      if (as_bf16) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
        overflow |= !(fabsf(x0) <= 3.0e38f) || !(fabsf(x1) <= 3.0e38f);
        *reinterpret_cast<__nv_bfloat162*>(d) = __halves2bfloat162(h0, h1);
        *reinterpret_cast<__nv_bfloat162*>(d + plane) = __halves2bfloat162(l0, l1);
      } else {
        __half hi0, lo0, hi1, lo1;
        split_f16(x0, hi0, lo0);
        split_f16(x1, hi1, lo1);
        overflow |= __hisinf(hi0) || __hisnan(hi0) || __hisinf(hi1) || __hisnan(hi1);
        *reinterpret_cast<__half2*>(d) = __halves2half2(hi0, hi1);
        *reinterpret_cast<__half2*>(d + plane) = __halves2half2(lo0, lo1);
      }
    }
  }
  if (overflow && range_flag) atomicOr(range_flag, 1);
}

// every weight matrix of the model, as stored ([in][out] row-major), -> bf16 hi/lo planes [2][round_up(in,64)][np] with
// columns [N, np) zero-filled, one launch: the B operands of the input-gradient GEMMs.
struct RowsEntry {
  int64_t src_off;   // float offset into params
  int64_t dst_off;   // half offset of the hi plane in the arena
  int64_t plane;     // halves between the hi and the lo plane
  int K, N, np;
  int block_begin;   // first CTA of this matrix
};
__global__ void __launch_bounds__(256)
prep_weights_rows_kernel(const float* __restrict__ params, __half* __restrict__ arena, const RowsEntry* __restrict__ table,
                         int n_entries) {
  griddep_launch();
  griddep_wait();
  int e = 0;
  while (e + 1 < n_entries && (int)blockIdx.x >= table[e + 1].block_begin) ++e;
  const RowsEntry t = table[e];
  const size_t idx = (size_t)(blockIdx.x - t.block_begin) * blockDim.x + threadIdx.x;   // 4 destination columns each
  const int cq = t.np >> 2;
  if (idx >= (size_t)t.K * cq) return;
  const int row = (int)(idx / cq), c = (int)(idx % cq) * 4;
  const float* sp = params + t.src_off + (size_t)row * t.N + c;
  __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float f = (c + j < t.N) ? sp[j] : 0.f;
    hi[j] = __float2bfloat16_rn(f);
    lo[j] = __float2bfloat16_rn(f - __bfloat162float(hi[j]));
  }
  __half* d = arena + t.dst_off + (size_t)row * t.np + c;
  *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(d + t.plane) = *reinterpret_cast<const uint2*>(lo);
}

// fp32 [M, C] rows (row pitch ld) -> bf16 hi/lo planes [2][rows_alloc][ld_dst], row-major, columns [C, ld_dst) zero-filled:
// the A operand (dY, contraction over the layer's outputs) and the B operand (W as stored, [in][out]) of the
// input-gradient GEMM dX = dY @ W^T.
__global__ void split_rows_bf16_kernel(const float* __restrict__ src, int ld, int M, int C, __half* __restrict__ dst,
                                       size_t plane, int ld_dst, int* range_flag) {
  griddep_launch();
  griddep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 4 destination columns
  const int cq = ld_dst >> 2;
  if (idx >= (size_t)M * cq) return;
  const int row = (int)(idx / cq), c = (int)(idx % cq) * 4;
  const float* sp = src + (size_t)row * ld + c;
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  if (c + 3 < C && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4 v = *reinterpret_cast<const float4*>(sp);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < C) f[j] = sp[j];
  }
  __align__(8) __nv_bfloat16 hi[4], lo[4];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = __float2bfloat16_rn(f[j]);
    lo[j] = __float2bfloat16_rn(f[j] - __bfloat162float(hi[j]));
    bad |= !(fabsf(f[j]) <= 3.0e38f);
  }
  __half* d = dst + (size_t)row * ld_dst + c;
  *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(d + plane) = *reinterpret_cast<const uint2*>(lo);
  if (bad && range_flag) atomicOr(range_flag, 1);
}

// uint8 pixels (the reference's dataset format, data.py:35-107) -> float32 / 255 (load_data, data.py:116) written both
// as fp32 rows (read by the glimpse-read and paint kernels) and, for the tensor-core engine, as the hl operand of the
// first encoder layer.  One pass over the image batch.
// gather != null: row r of the batch is row gather[r] of a device-resident uint8 dataset (the minibatch indices of
// tensors_from_data, data.py:121-158), so the host never touches the pixels.
__global__ void u8_to_f32_hl_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, __half* __restrict__ hl,
                                    size_t plane, int ld_hl, int M, int K, const int32_t* __restrict__ gather,
                                    long long n_src = 0) {
  griddep_launch();
  griddep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 4 pixels
  const int kq = (K + 3) / 4;
  if (idx >= (size_t)M * kq) return;
  const int row = (int)(idx / kq), k = (int)(idx % kq) * 4;
  long long srow = gather ? (long long)gather[row] : (long long)row;
  if (gather && n_src > 0) srow = srow < 0 ? 0 : (srow >= n_src ? n_src - 1 : srow);   // a bad index must not read out of bounds
  const uint8_t* s = src + (size_t)srow * K + k;
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = (k + 3 < K) && ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
  if (vec) {
    const uchar4 u = *reinterpret_cast<const uchar4*>(s);
    f[0] = __fdiv_rn((float)u.x, 255.0f);
    f[1] = __fdiv_rn((float)u.y, 255.0f);
    f[2] = __fdiv_rn((float)u.z, 255.0f);
    f[3] = __fdiv_rn((float)u.w, 255.0f);
    *reinterpret_cast<float4*>(dst + (size_t)row * K + k) = make_float4(f[0], f[1], f[2], f[3]);
  } else {
    for (int j = 0; j < 4 && k + j < K; ++j) {
      f[j] = __fdiv_rn((float)s[j], 255.0f);
      dst[(size_t)row * K + k + j] = f[j];
    }
  }
  if (hl) {
    __half* dh = hl + (size_t)row * ld_hl + k;
    __align__(8) __half hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_f16(f[j], hi[j], lo[j]);
    if (k + 3 < K) {
      *reinterpret_cast<uint2*>(dh) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(dh + plane) = *reinterpret_cast<const uint2*>(lo);
    } else {
      for (int j = 0; j < 4 && k + j < K; ++j) {
        dh[j] = hi[j];
        dh[plane + j] = lo[j];
      }
    }
  }
}

// One launch converts every weight matrix of the model: W[K,N] fp32 (row-major, ld = N) -> W^T hl [2][N_alloc][Kpad]
// fp16 split of (w * 2^8).  Each CTA transposes one 32x32 tile through shared memory.
struct PrepEntry {
  int64_t src_off;     // float offset into params
  int64_t dst_off;     // half offset into the prepared-weight arena (hi plane)
  int64_t plane;       // halves between hi and lo plane
  int K, N, Kpad;
  int tile_begin;      // first tile index of this matrix in the global tile list
  int tiles_n;         // tiles along N
  int split_n;         // > 0: source columns n >= split_n land at destination row n - split_n + split_off (the what
  int split_off;       //      head's loc / scale halves, each starting on a 16-row boundary for chain_tc.cuh)
  int64_t bias_src;    // float offset of the bias in params (< 0: none) and of its zero-padded copy in the bias arena
  int64_t bias_dst;
  int perm_nh;         // > 0: LSTM gate columns (gate-major, [4][nh]) regrouped per cluster CTA for lstm_tc.cuh:
                       //      column gate * nh + unit -> row (unit / upc) * nh + gate * upc + unit % upc, upc = nh / 4
};
__device__ __forceinline__ int prep_dst_row(const PrepEntry& t, int n) {
  if (t.perm_nh > 0) {
    const int upc = t.perm_nh >> 2, gate = n / t.perm_nh, unit = n % t.perm_nh;
    return (unit / upc) * t.perm_nh + gate * upc + unit % upc;
  }
  return (t.split_n > 0 && n >= t.split_n) ? n - t.split_n + t.split_off : n;
}
__global__ void __launch_bounds__(256)
prep_weights_kernel(const float* __restrict__ params, __half* __restrict__ arena, const PrepEntry* __restrict__ table,
                    int n_entries, int* range_flag, float* __restrict__ bias_arena) {
  __shared__ float tile[32][33];
  griddep_launch();
  griddep_wait();
  int e = 0;
  while (e + 1 < n_entries && (int)blockIdx.x >= table[e + 1].tile_begin) ++e;
  const PrepEntry t = table[e];
  const int local = blockIdx.x - t.tile_begin;
  const int tk = local / t.tiles_n, tn = local % t.tiles_n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* W = params + t.src_off;
  if (tk == 0 && bias_arena && t.bias_src >= 0 && threadIdx.x < 32) {   // padded bias copy (same row remap as W^T)
    const int n = tn * 32 + threadIdx.x;
    if (n < t.N) {
      bias_arena[t.bias_dst + prep_dst_row(t, n)] = params[t.bias_src + n];
    }
  }
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int k = tk * 32 + ty + i, n = tn * 32 + tx;
    tile[ty + i][tx] = (k < t.K && n < t.N) ? W[(size_t)k * t.N + n] : 0.f;
  }
  __syncthreads();
  bool overflow = false;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int n = tn * 32 + ty + i, k = tk * 32 + tx;
    if (n < t.N && k < t.K) {
      __half hi, lo;
      split_f16(tile[tx][ty + i] * W_SCALE, hi, lo);
      overflow |= __hisinf(hi) || __hisnan(hi);
      const int n_dst = prep_dst_row(t, n);
      __half* d = arena + t.dst_off + (size_t)n_dst * t.Kpad + k;
      d[0] = hi;
      d[t.plane] = lo;
    }
  }
  if (overflow && range_flag) atomicOr(range_flag, 1);
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 tensor map over an hl buffer viewed as [rows_total][kpad], box = [box_rows][64], 128-byte swizzle.
inline bool make_tmap(CUtensorMap* tm, const __half* base, int kpad, int64_t rows_total, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES>
inline cudaError_t launch_gemm_cfg(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmParams& p, int n_alloc,
                                   cudaStream_t st) {
  using L = Smem<BN, STAGES>;
  cudaError_t e = ensure_dynamic_smem(linear_tc_kernel<BN, STAGES>, L::TOTAL);
  if (e != cudaSuccess) return e;
  const int nz = p.kb_per_z > 0 ? (p.num_k_blocks + p.kb_per_z - 1) / p.kb_per_z : 1;
  dim3 grid(n_alloc / BN, (p.M + BM - 1) / BM, nz);
  return launch_k(linear_tc_kernel<BN, STAGES>, grid, dim3(num_threads(BN)), L::TOTAL, st, tm_a, tm_b, p);
}

// Tile width `bn` (32 or 64) is fixed by how the weight was prepared.  Short K loops use a 2-stage ring (<= 100 KB of
// smem: two CTAs per SM, one's epilogue overlapping the other's main loop); long ones (the 2500-wide input encoder
// layer) use 4 stages.
inline cudaError_t launch_gemm(int bn, const CUtensorMap& tm_a, const CUtensorMap& tm_b, const GemmParams& p,
                               int n_alloc, cudaStream_t st) {
  const bool deep = (p.kb_per_z > 0 ? p.kb_per_z : p.num_k_blocks) > 8;
  if (bn == 32) return deep ? launch_gemm_cfg<32, 4>(tm_a, tm_b, p, n_alloc, st) : launch_gemm_cfg<32, 2>(tm_a, tm_b, p, n_alloc, st);
  return deep ? launch_gemm_cfg<64, 4>(tm_a, tm_b, p, n_alloc, st) : launch_gemm_cfg<64, 2>(tm_a, tm_b, p, n_alloc, st);
}

}  // namespace tc
}  // namespace air
