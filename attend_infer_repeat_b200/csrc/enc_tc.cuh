// First layer of the input Encoder (modules.py:66-76, neural.py:63-102): e1 = ELU(img @ W1 + b1), img [B, P] fp32 straight
// from HBM, W1 [P, 256] -- 39.7 % of the cell's MACs as written, 67 % at 100x100, and the only GEMM of the
// path whose A operand is an INPUT (everything else is produced by a kernel that can emit the tensor-core operand format).
//
// The layer-by-layer engine handled it as: split_rows (fp32 -> fp16 hi/lo planes, 2 x 41 MB of HBM traffic at B = 4096)
// followed by a 128 x 64-tile GEMM whose four N-split CTAs each re-read the A planes (L2-bound).  This kernel:
//   * splits K, not N: a 4-CTA cluster owns one 128-canvas row tile, CTA r contracts k-blocks [r * nkb, (r + 1) * nkb) with
//     the full N = 256 (main and cross accumulators = all 512 TMEM columns).  Every image byte is read once.  Four short
//     accumulations instead of one of P / 16 steps also cut the tensor core's accumulator truncation (one round-toward-zero
//     per MMA): the K = 2500 layer was the largest single error source of the split engine;
//   * converts in the kernel: eight converter warps load fp32 image rows (128-byte segments), split them into fp16 hi/lo
//     and store them in the 128-byte-swizzled K-major layout tcgen05.mma reads -- no operand planes in HBM at all;
//   * reduces through distributed shared memory: each CTA parks its 128 x 256 fp32 partial in its own (now idle) stage
//     ring, column-major; CTA r then sums columns [64 r, 64 r + 64) over the four CTAs in a fixed order (deterministic),
//     adds the bias, applies ELU and writes e1 as the hl operand planes of the next layer (and fp32 rows in training mode).
// Roles: warp 0 = TMA producer (W1 tiles), warp 1 = TMEM alloc + MMA issuer, warps 2..9 = converters, then epilogue.
#pragma once
#include "lstm_tc.cuh"

namespace air {
namespace enc {

using namespace air::chain;
using air::lstm::cluster_ctarank;
using air::lstm::cluster_sync_all;

constexpr int KSPLIT = 4;
constexpr int N1 = 256;                     // width of the first hidden layer this kernel is specialised for
constexpr int STAGES = 2;
constexpr int A_TILE = BM * BK * 2;         // 16 KB: 128 rows x 64 fp16, 128-byte swizzled
constexpr int W_TILE_BYTES = N1 * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * W_TILE_BYTES;   // A_hi, A_lo, W_hi, W_lo = 96 KB
constexpr int CONV_WARPS = 8;
constexpr int CONV_THREADS = 32 * CONV_WARPS;
constexpr int ENC_THREADS = 64 + CONV_THREADS;
constexpr int ENC_BAR_OFFSET = STAGES * STAGE_BYTES;
constexpr int ENC_SMEM_BYTES = ENC_BAR_OFFSET + 256 + 1024;
static_assert(N1 * BM * 4 <= STAGES * STAGE_BYTES, "the fp32 partial must fit the stage ring");

struct Params {
  CUtensorMap tm_w;        // prepared W1^T hl planes [2 * N_alloc][Kpad], box 256 rows x 64 K
  const float* img;        // [B, P] fp32, rows 16-byte aligned (P % 4 == 0)
  const float* bias;       // [N1]
  int B, P;
  int nkb_total;           // Kpad / 64
  int nkb_per_cta;         // ceil(nkb_total / KSPLIT)
  int w_lo_row;            // rows between the hi and the lo plane of W1^T (N_alloc)
  __half* out_hl;          // e1 as hl planes [2][rows_alloc][ld_hl] (operand of the next layer), or null
  size_t hl_plane;
  int ld_hl;
  int hl_nsl;              // > 0: out_hl is slice-major tiled, [row / 128][hl_nsl][128 rows][16] per plane (the cluster LSTM's operand)
  float* out_f32;          // e1 fp32 rows [B, N1] (training mode / fp32 consumers), or null
  int* range_flag;
};

__device__ __forceinline__ uint32_t mapa_u32(const void* smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__global__ void __cluster_dims__(KSPLIT, 1, 1) __launch_bounds__(ENC_THREADS, 1) enc1_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + ENC_BAR_OFFSET);   // W tiles landed + A tile converted
  uint64_t* empty_bar = full_bar + STAGES;                                   // the stage's MMAs have retired
  uint64_t* acc_bar = empty_bar + STAGES;                                    // all MMAs of this CTA have retired
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.y * BM;
  const int kb0 = (int)rank * p.nkb_per_cta;
  const int nkb = max(0, min(p.nkb_per_cta, p.nkb_total - kb0));

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + CONV_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
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
    // ===== TMA producer: W1 hi / lo tiles of this CTA's k-blocks =====
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], 2 * W_TILE_BYTES);
        tma_load_2d(st + 2 * A_TILE, &p.tm_w, (kb0 + kb) * BK, 0, &full_bar[s]);
        tma_load_2d(st + 2 * A_TILE + W_TILE_BYTES, &p.tm_w, (kb0 + kb) * BK, p.w_lo_row, &full_bar[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: main terms -> columns [0, 256), cross terms -> [256, 512) (linear_tc.cuh) =====
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(BM, N1);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
        tc_fence_after();
        uint8_t* st = smem + s * STAGE_BYTES;
        const uint64_t da_hi = make_smem_desc_sw128(st), da_lo = make_smem_desc_sw128(st + A_TILE);
        const uint64_t db_hi = make_smem_desc_sw128(st + 2 * A_TILE);
        const uint64_t db_lo = make_smem_desc_sw128(st + 2 * A_TILE + W_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);
          umma_f16(tmem_base, da_hi + adv, db_hi + adv, idesc, (kb | k) != 0);
          umma_f16(tmem_base + N1, da_lo + adv, db_hi + adv, idesc, (kb | k) != 0);
          umma_f16(tmem_base + N1, da_hi + adv, db_lo + adv, idesc, 1);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(acc_bar);
    }
    __syncwarp();
  } else {
    // ===== converter warps: fp32 image rows -> fp16 hi / lo, swizzled K-major tiles =====
    // Thread ct covers, in each of four 32-row passes, row (pass * 32 + ct / 8) and the two 16-byte pieces (ct % 8) and
    // (ct % 8 + 8) of the row's 256-byte k-block segment: eight lanes read 128 contiguous bytes.  The loads of k-block
    // kb + 1 are in flight while k-block kb is converted (the tensor core needs a k-block every ~0.8 us, HBM answers in ~1).
    const int ct = threadIdx.x - 64;          // 0..255
    const int seg = ct & 7, r0 = ct >> 3;
    uint32_t ovf = 0;
    // three register buffers: the loads of k-blocks kb + 1 and kb + 2 are in flight while kb is converted -- 64 KB per CTA,
    // ~8 MB chip-wide, what Little's law asks for at HBM latency (one buffer ahead measured 1.4 TB/s)
    float4 f[3][8];
    auto issue = [&](int kb, float4 (&dst)[8]) {
      const int kbase = (kb0 + kb) * BK;
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int row = m0 + ps * 32 + r0;
        const float* src_row = p.img + (size_t)(row < p.B ? row : 0) * p.P;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = kbase + (seg + 8 * j) * 4;
          dst[ps * 2 + j] = (row < p.B && k < p.P) ? __ldg(reinterpret_cast<const float4*>(src_row + k))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    if (nkb > 0) issue(0, f[0]);
    if (nkb > 1) issue(1, f[1]);
#pragma unroll 1
    for (int kb = 0; kb < nkb; kb += 3) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int kk = kb + u;
        if (kk >= nkb) break;
        if (kk + 2 < nkb) issue(kk + 2, f[(u + 2) % 3]);
        const int s = kk % STAGES;
        mbar_wait(&empty_bar[s], ((kk / STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int r = ps * 32 + r0;
          const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float4 x = f[u][ps * 2 + j];
            const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
            const float2 g0 = __half22float2(h0), g1 = __half22float2(h1);
            const __half2 l0 = __floats2half2_rn(x.x - g0.x, x.y - g0.y), l1 = __floats2half2_rn(x.z - g1.x, x.w - g1.y);
            const uint32_t hw0 = *reinterpret_cast<const uint32_t*>(&h0), hw1 = *reinterpret_cast<const uint32_t*>(&h1);
            ovf |= ((hw0 & 0x7C007C00u) + 0x04000400u) | ((hw1 & 0x7C007C00u) + 0x04000400u);
            const int q4 = seg + 8 * j;                       // which 4-element piece of the 64-wide k-block
            const uint32_t chunk = (uint32_t)(q4 >> 1);       // 16-byte chunk (8 fp16)
            const uint32_t off = row_off + ((chunk ^ (uint32_t)(r & 7)) << 4) + (uint32_t)(q4 & 1) * 8u;
            *reinterpret_cast<uint2*>(st + off) = make_uint2(hw0, hw1);
            *reinterpret_cast<uint2*>(st + A_TILE + off) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
          }
        }
        fence_proxy_async();                    // generic-proxy stores -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);   // one arrival per warp: 256 arrivals on one barrier serialise
      }
    }
    if ((ovf & 0x80008000u) && p.range_flag) atomicOr(p.range_flag, 1);

    // ===== epilogue part 1: this CTA's partial -> its own stage ring as fp32 [256 columns][128 rows] =====
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    {
      const int q = warp & 3, chh = (warp - 2) >> 2;     // TMEM lane quadrant, column half
      const int rit = q * 32 + lane;
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      float* part = reinterpret_cast<float*>(smem);
#pragma unroll 2
      for (int g = 0; g < 8; ++g) {
        const int c0 = chh * 128 + g * 16;
        float v[16], vx[16];
        if (nkb > 0) {
          tmem_ld_32x16(t_lane + (uint32_t)c0, v);
          tmem_ld_32x16(t_lane + (uint32_t)(N1 + c0), vx);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = vx[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) part[(size_t)(c0 + j) * BM + rit] = (v[j] + vx[j]) * W_UNSCALE;
      }
    }
  }
  // every CTA's partial is in place
  tc_fence_before();
  cluster_sync_all();

  if (warp >= 2) {
    // ===== epilogue part 2: CTA `rank` finishes columns [64 rank, 64 rank + 64) =====
    const int ct = threadIdx.x - 64;
    const int r = ct & 127, sb = ct >> 7;     // row of the tile, which 32 of this CTA's 64 columns
    const int row = m0 + r;
    const int cbase = (int)rank * 64 + sb * 32;
    const float* part = reinterpret_cast<const float*>(smem);
    uint32_t remote[KSPLIT];
#pragma unroll
    for (uint32_t j = 0; j < KSPLIT; ++j) remote[j] = mapa_u32(part + (size_t)cbase * BM + r, j);
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      float acc = ld_dsmem_f32(remote[0] + (uint32_t)c * BM * 4u);
#pragma unroll
      for (uint32_t j = 1; j < KSPLIT; ++j) acc += ld_dsmem_f32(remote[j] + (uint32_t)c * BM * 4u);
      const float x = acc + __ldg(p.bias + cbase + c);
      v[c] = elu_fast(x);
    }
    if (row < p.B) {
      if (p.out_f32) {
        float4* dst = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * N1 + cbase);
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
      if (p.out_hl) {
        uint32_t hi[16], lo[16];
        uint32_t ovf = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          const float2 hf2 = __half22float2(h);
          const __half2 l = __floats2half2_rn(v[2 * j] - hf2.x, v[2 * j + 1] - hf2.y);
          hi[j] = *reinterpret_cast<const uint32_t*>(&h);
          lo[j] = *reinterpret_cast<const uint32_t*>(&l);
          ovf |= (hi[j] & 0x7C007C00u) + 0x04000400u;
        }
        if (p.hl_nsl > 0) {   // two 16-column slices of this row, each one 32-byte piece of its [128 rows][16] block
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            __half* d = p.out_hl + ((((size_t)row >> 7) * p.hl_nsl + (size_t)(cbase >> 4) + sl) * 128 + (row & 127)) * 16;
            uint4* dh = reinterpret_cast<uint4*>(d);
            uint4* dl = reinterpret_cast<uint4*>(d + p.hl_plane);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int q = sl * 2 + c;
              dh[c] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
              dl[c] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
          }
        } else {
        uint4* dh = reinterpret_cast<uint4*>(p.out_hl + (size_t)row * p.ld_hl + cbase);
        uint4* dl = reinterpret_cast<uint4*>(p.out_hl + p.hl_plane + (size_t)row * p.ld_hl + cbase);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          dh[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          dl[c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
        }
        if ((ovf & 0x80008000u) && p.range_flag) atomicOr(p.range_flag, 1);
      }
    }
  }
  // no CTA leaves (or frees tensor memory) while a peer may still read its partial
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

inline cudaError_t launch_enc1(const Params& p, cudaStream_t st) {
  cudaError_t e = ensure_dynamic_smem(enc1_kernel, ENC_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_k(enc1_kernel, dim3(KSPLIT, (p.B + BM - 1) / BM), dim3(ENC_THREADS), ENC_SMEM_BYTES, st, p);
}

}  // namespace enc
}  // namespace air
