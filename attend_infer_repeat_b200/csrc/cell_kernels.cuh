// Non-GEMM stages of the AIR cell: LSTM gate math, where/what sampling, presence scan, the
// spatial-transformer glimpse read and the inverse-transformer canvas paint fused with the ELBO terms.
// All HBM-bound byte work: one CTA per canvas, the image / glimpse tile staged once in shared memory
// (cp.async.bulk where the tile is 16-byte sized), coalesced output rows, warp-shuffle reductions.
//
// Both transformer directions are axis-aligned (no shear), so the bilinear footprint of an output pixel is the outer
// product of a per-column and a per-row tap pair.  Each CTA first builds small shared-memory tables of those taps
// (index pair + weight pair per output column / row, coordinates computed exactly as Sonnet does), then every output
// pixel costs four shared-memory loads and the reference kernel's arithmetic, in the reference's association order.
#pragma once
#include "common.cuh"
#include "../../include/air_b200.h"

namespace air {

// ---------------------------------------------------------------------------------------------------
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP), completion on an mbarrier.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool bulk_ok(const void* src, int n_floats) {
  return ((n_floats & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
}

// ---------------------------------------------------------------------------------------------------
// bilinear taps along one axis (snt.resampler semantics: zero outside (-1, n); taps outside [0, n-1] read 0)
// ---------------------------------------------------------------------------------------------------
struct __align__(16) Tap {
  float wf, wc;   // weight of the floor tap (= ceil - coord) and of the ceil tap (= 1 - that); 0 where the tap reads 0
  int i_f, i_c;   // clamped indices of the two taps
};
__device__ __forceinline__ Tap make_tap(float coord, int n, int stride = 1) {
  Tap t;
  const bool inside = coord > -1.0f && coord < (float)n;
  const float f = floorf(coord);
  const int fi = (int)f, ci = fi + 1;
  const float d = (f + 1.0f) - coord;
  t.wf = (inside && fi >= 0 && fi <= n - 1) ? d : 0.f;
  t.wc = (inside && ci >= 0 && ci <= n - 1) ? 1.0f - d : 0.f;
  t.i_f = min(max(fi, 0), n - 1) * stride;   // stride > 1: row taps pre-multiplied by the source row pitch
  t.i_c = min(max(ci, 0), n - 1) * stride;
  return t;
}
// same association as the reference kernel: dx*dy*D(fx,fy) + (1-dx)(1-dy)*D(cx,cy) + dx(1-dy)*D(fx,cy) + (1-dx)dy*D(cx,fy)
__device__ __forceinline__ float bilinear(const float* __restrict__ D, int Ws, const Tap& x, const Tap& y) {
  const float* rf = D + y.i_f * Ws;
  const float* rc = D + y.i_c * Ws;
  float r = __fmul_rn(__fmul_rn(x.wf, y.wf), rf[x.i_f]);
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(x.wc, y.wc), rc[x.i_c]));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(x.wf, y.wc), rc[x.i_f]));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(x.wc, y.wf), rf[x.i_c]));
  return r;
}

// bilinear with the row taps' indices pre-multiplied by the source row pitch (make_tap(..., stride = Ws))
__device__ __forceinline__ float bilinear_pre(const float* __restrict__ D, const Tap& x, const Tap& y) {
  const float* rf = D + y.i_f;
  const float* rc = D + y.i_c;
  float r = __fmul_rn(__fmul_rn(x.wf, y.wf), rf[x.i_f]);
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(x.wc, y.wc), rc[x.i_c]));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(x.wf, y.wc), rc[x.i_f]));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(x.wc, y.wf), rf[x.i_c]));
  return r;
}

// ---------------------------------------------------------------------------------------------------
// In-library noise (SURVEY 8d "perf mode"): the reference draws where / what / presence noise inside the graph
// (cell.py:133,147,156), so a caller that feeds only images needs no noise bytes from the host.  Counter-based
// Philox4x32-10 (Salmon et al., SC'11): element i of stream `stream_id` under `seed` is a pure function of
// (seed, stream_id, i) -- reproducible across launches, batch shards and devices.  Normals by Box-Muller.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// (0, 1]: never 0, so log() below is finite; 24 random bits
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// one thread per 4 outputs; normal != 0: N(0,1) (two Box-Muller pairs), else U[0,1)
__global__ void philox_fill_kernel(float* __restrict__ out, size_t n, unsigned long long seed, uint32_t stream_id,
                                   int normal) {
  griddep_launch();
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream_id, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float v[4];
  if (normal) {
    const float r0 = sqrtf(-2.0f * logf(u01_open(r.x))), r1 = sqrtf(-2.0f * logf(u01_open(r.z)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u01_open(r.y), &s0, &c0);
    sincospif(2.0f * u01_open(r.w), &s1, &c1);
    v[0] = r0 * c0; v[1] = r0 * s0; v[2] = r1 * c1; v[3] = r1 * s1;
  } else {
    v[0] = (float)(r.x >> 8) * (1.0f / 16777216.0f);   // [0, 1)
    v[1] = (float)(r.y >> 8) * (1.0f / 16777216.0f);
    v[2] = (float)(r.z >> 8) * (1.0f / 16777216.0f);
    v[3] = (float)(r.w >> 8) * (1.0f / 16777216.0f);
  }
  griddep_wait();   // the buffer may still be read by the previous pass
  for (int j = 0; j < 4 && q * 4 + j < n; ++j) out[q * 4 + j] = v[j];
}

// ---------------------------------------------------------------------------------------------------
// LSTM state handling (snt.LSTM [upstream], mnist_model.py:35; cell.py:101-114,126-127)
// ---------------------------------------------------------------------------------------------------
// trainable initial state (h0, c0) [nh] tiled to the batch (cell.py:103)
__global__ void lstm_init_state_kernel(const float* __restrict__ h0, const float* __restrict__ c0,
                                       float* __restrict__ h, float* __restrict__ c, int B, int nh, HlOut h_hl) {
  griddep_launch();
  griddep_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * nh) return;
  const int u = (int)(i % nh);
  h[i] = h0[u];
  c[i] = c0[u];
  if (h_hl.p) hl_store(h_hl, i / nh, u, h0[u]);
}
// explicit incoming hidden state (air_cell_step) -> hl operand of the recurrent GEMM
__global__ void split_state_kernel(const float* __restrict__ h, int B, int nh, HlOut h_hl) {
  griddep_launch();
  griddep_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * nh) return;
  hl_store(h_hl, i / nh, (int)(i % nh), h[i]);
}

// gates[B,4nh] (order i, j, f, o) already hold [x,h] @ W + b.  c <- sig(f + fb) * c + sig(i) * tanh(j);
// h <- tanh(c) * sig(o).
// c_out may alias c (in-place state) or be the next slice of the saved cell-state history (training mode).
__global__ void lstm_pointwise_kernel(const float* __restrict__ gates, const float* c, float* c_out,
                                      float* __restrict__ h_out, int B, int nh, float forget_bias, HlOut h_hl,
                                      HlOut h_hl2, size_t row0_hl2) {
  griddep_launch();
  griddep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * nh) return;
  const size_t b = idx / nh;
  const int u = (int)(idx % nh);
  const float* g = gates + b * 4 * (size_t)nh;
  const float gi = g[u], gj = g[nh + u], gf = g[2 * nh + u], go = g[3 * nh + u];
  const float c_new = __fadd_rn(__fmul_rn(sigmoid_f(gf + forget_bias), c[idx]), __fmul_rn(sigmoid_f(gi), tanhf(gj)));
  c_out[idx] = c_new;
  const float h_new = __fmul_rn(tanhf(c_new), sigmoid_f(go));
  h_out[idx] = h_new;
  if (h_hl.p) hl_store(h_hl, b, u, h_new);
  if (h_hl2.p) hl_store(h_hl2, row0_hl2 + b, u, h_new);   // second copy in the fused chains' layout, row = t * B + b
}

// ---------------------------------------------------------------------------------------------------
// presence: StepsPredictor sigmoid + explore-eps mix + Bernoulli draw + cumulative product over steps
// (modules.py:119-122, cell.py:137-151)
// ---------------------------------------------------------------------------------------------------
// (implemented by presence_scan() below: one thread of each canvas's where_read CTA)

// q = n / d, r = n % d for 0 <= n < 2^20, 1 <= d without the ~25-instruction integer division: the float quotient of
// (n + 0.5) / d is off by at most one, which the remainder check repairs
__device__ __forceinline__ void small_divmod(int n, int d, int& q, int& r) {
  q = (int)(((float)n + 0.5f) * __frcp_rn((float)d));
  r = n - q * d;
  if (r < 0) { --q; r += d; }
  else if (r >= d) { ++q; r -= d; }
}

// ---------------------------------------------------------------------------------------------------
// where head + glimpse read.  One CTA per canvas b: the image is staged once in shared memory and all
// T glimpses of that canvas are cropped from it (the T where-codes depend only on h_t, never on the crop).
//   m[T,B,8] = transform-estimator MLP output; loc = (sig, tanh, sig, tanh)(m[0:4]) * (max_crop,1,max_crop,1);
//   scale = softplus(m[4:8] + scale_bias); where = eps * scale + loc       (modules.py:41-63, cell.py:129-133)
//   crop[t,b] = resampler(img[b], AffineGridWarper(where))                  (modules.py:104-109, cell.py:135)
// dynamic smem: H*W floats (image) + T*(w+h) taps.
// ---------------------------------------------------------------------------------------------------
// 128 threads, 16 CTAs per SM: the kernel is a chain of dependent latencies (where code -> taps -> image copy -> crop), so
// more, smaller CTAs in flight hide more of it than 8 of 256 did (measured at B = 4096: 35 -> see profiles/r02)
constexpr int WHERE_READ_THREADS = 128;
__host__ __device__ inline size_t where_read_smem(int T, int H, int W, int h, int w) {
  // image + tap tables + (tensor-core engine) the fp16 hi / lo staging of the T glimpses
  return sizeof(float) * ((size_t)H * W + 4) / 16 * 16 + 16 + sizeof(Tap) * (size_t)T * (w + h) +
         2 * sizeof(__half) * (size_t)T * h * w;
}

// presence scan of one canvas, run by one thread of the canvas's where_read CTA
struct PresenceArgs {
  const float* logit;        // [T,B]; null: no scan
  const float* u_pres;       // [T,B]
  const float* presence_in;  // [B] or null (= 1)
  float* presence_prob;      // [T,B]
  float* presence;           // [T,B]
  float step_bias, explore_eps;
  int discrete;
};
// one step of the scan: presence probability p_t and the carried presence (cell.py:137-151)
__device__ __forceinline__ void presence_step(const PresenceArgs& pa, size_t i, float& pres, float& p) {
  p = sigmoid_f(pa.logit[i] + pa.step_bias);
  if (pa.explore_eps >= 0.f) p = __fadd_rn(pa.explore_eps / 2.0f, __fmul_rn(1.0f - pa.explore_eps, p));
  if (pa.discrete) {
    const float z = (pa.u_pres[i] < p) ? 1.0f : 0.0f;
    pres *= z;
  } else {
    pres = p;
  }
}
__device__ __forceinline__ void presence_scan(const PresenceArgs& pa, int b, int T, int B) {
  float pres = pa.presence_in ? pa.presence_in[b] : 1.0f;
  // every logit / uniform draw of the canvas is loaded before the first store: the stores below may alias them as far as the
  // compiler knows, and the scan would otherwise be T dependent memory round trips instead of one
  float lg[AIR_MAX_STEPS], uu[AIR_MAX_STEPS];
#pragma unroll
  for (int t = 0; t < AIR_MAX_STEPS; ++t) {
    lg[t] = t < T ? __ldg(pa.logit + (size_t)t * B + b) : 0.f;
    uu[t] = (t < T && pa.discrete) ? __ldg(pa.u_pres + (size_t)t * B + b) : 0.f;
  }
#pragma unroll
  for (int t = 0; t < AIR_MAX_STEPS; ++t) {
    if (t < T) {
      const size_t i = (size_t)t * B + b;
      float p = sigmoid_f(lg[t] + pa.step_bias);
      if (pa.explore_eps >= 0.f) p = __fadd_rn(pa.explore_eps / 2.0f, __fmul_rn(1.0f - pa.explore_eps, p));
      if (pa.discrete) pres *= (uu[t] < p) ? 1.0f : 0.0f;
      else pres = p;
      pa.presence_prob[i] = p;
      pa.presence[i] = pres;
    }
  }
}

// FT .. FNT > 0 fix the step count, the shapes and the thread count at compile time (the quoted configuration: every divisor and
// trip count folds); 0 = taken from the arguments
template <int FT, int FH, int FW, int Fh, int Fw, int FNT>
__global__ void __launch_bounds__(WHERE_READ_THREADS, 2048 / WHERE_READ_THREADS)
where_read_kernel(const float* __restrict__ m, const float* __restrict__ eps_where, const float* __restrict__ img,
                  float* __restrict__ where, float* __restrict__ where_loc, float* __restrict__ where_scale,
                  float* __restrict__ crop, HlOut crop_hl, int T_, int B, int H_, int W_, int h_, int w_, float max_crop,
                  float scale_bias, double step_w, double step_h, PresenceArgs pa, long long* trace) {
  const int T = FT ? FT : T_, H = FH ? FH : H_, W = FW ? FW : W_, h = Fh ? Fh : h_, w = Fw ? Fw : w_;
  const int NTC = FNT ? FNT : (int)blockDim.x;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ float s_where[AIR_MAX_STEPS][4];
  const int b = blockIdx.x;
  long long* tr = trace ? trace + (size_t)blockIdx.x * 8 : nullptr;   // debug (AIR_READ_TRACE): phase stamps, ns
#define READ_STAMP(i) do { if (tr && threadIdx.x == 0) tr[i] = (long long)globaltimer_ns(); } while (0)
  const int P = H * W, G = h * w;
  float* s_img = reinterpret_cast<float*>(smem_raw);
  Tap* s_tx = reinterpret_cast<Tap*>(smem_raw + (sizeof(float) * ((size_t)P + 4) / 16 * 16 + 16));   // [T][w]
  Tap* s_ty = s_tx + (size_t)T * w;                                                                 // [T][h]
  const float* src = img + (size_t)b * P;
  const bool bulk = bulk_ok(src, P);
  if (threadIdx.x == 0 && bulk) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  griddep_launch();
  griddep_wait();   // m / eps / where come from earlier kernels of the chain -- and so may the image (the uint8 entry
                    // points convert it on this stream; every kernel in between releases its dependents early, so the
                    // prologue is NOT transitively ordered after that producer): the copy starts after the wait
  READ_STAMP(0);
  if (threadIdx.x == 0 && bulk) {
    mbar_expect_tx(&bar, (uint32_t)P * 4u);
    bulk_g2s(s_img, src, (uint32_t)P * 4u, &bar);
  }
  // the T where codes of this canvas (4 components each) while the image is in flight
  if (threadIdx.x < 4 * T) {
    const int t = threadIdx.x >> 2, k = threadIdx.x & 3;
    const size_t row = (size_t)t * B + b;
    if (m) {
      const float mk = m[row * 8 + k];
      const float loc = (k & 1) ? tanhf(mk) : __fmul_rn(max_crop, sigmoid_f(mk));
      const float sc = softplus_f(m[row * 8 + 4 + k] + scale_bias);
      const float wv = __fadd_rn(__fmul_rn(eps_where[row * 4 + k], sc), loc);
      where_loc[row * 4 + k] = loc;
      where_scale[row * 4 + k] = sc;
      where[row * 4 + k] = wv;
      s_where[t][k] = wv;
    } else {   // m == null: the where code has been sampled by the row kernel's heads launch (row_tc.cuh)
      s_where[t][k] = where[row * 4 + k];
    }
  }
  if (pa.logit && threadIdx.x == 64) presence_scan(pa, b, T, B);   // StepsPredictor + Bernoulli draw (cell.py:137-151)
  if (!bulk)
    for (int i = threadIdx.x; i < P; i += NTC) s_img[i] = src[i];
  __syncthreads();
  READ_STAMP(1);
  for (int i = threadIdx.x; i < T * (w + h); i += NTC) {
    const int t = i / (w + h), j = i - t * (w + h);
    if (j < w) s_tx[t * w + j] = make_tap(fwd_coord_s(s_where[t][0], s_where[t][1], j, step_w, W), W, 1);
    else       s_ty[t * h + (j - w)] = make_tap(fwd_coord_s(s_where[t][2], s_where[t][3], j - w, step_h, H), H, W);
  }
  __syncthreads();
  READ_STAMP(2);
  if (bulk) mbar_wait(&bar, 0);
  READ_STAMP(3);
  const int NT = NTC;
  if (!crop && crop_hl.p && crop_hl.nsl > 0 && (G & 7) == 0) {
    // tensor-core engine, fused-chain operand layout ([row tile of 128][K slice of 16][128 rows][16 fp16]): 16-byte
    // stores of 8 consecutive glimpse pixels into the hi plane and into the lo plane.  The gather runs with lane <->
    // glimpse pixel (neighbouring lanes read neighbouring image pixels: few bank conflicts, taps of a row broadcast) and
    // stages the split halves in shared memory; a second sweep with thread <-> 8-pixel chunk does the global stores.
    // (Round 1 gathered with thread <-> chunk: half of its shared-memory wavefronts were bank-conflict replays and the
    // per-thread hi[8] / lo[8] arrays went through local memory.)
    __half* s_hi = reinterpret_cast<__half*>(s_ty + (size_t)T * h);   // [T][G]
    __half* s_lo = s_hi + (size_t)T * G;
    const int TG = T * G;
    for (int i = threadIdx.x; i < TG; i += NT) {
      int t, g, r, c;
      small_divmod(i, G, t, g);
      small_divmod(g, w, r, c);
      const Tap tx = s_tx[t * w + c], ty = s_ty[t * h + r];
      float v = 0.f;
      if (((tx.wf != 0.f) | (tx.wc != 0.f)) & ((ty.wf != 0.f) | (ty.wc != 0.f))) v = bilinear_pre(s_img, tx, ty);
      __half hi, lo;
      split_f16(v, hi, lo);
      s_hi[i] = hi;
      s_lo[i] = lo;
    }
    __syncthreads();
    const int chunks = G >> 3;
    for (int i = threadIdx.x; i < T * chunks; i += NT) {
      int t, ch;
      small_divmod(i, chunks, t, ch);
      const int g0 = ch << 3;
      const size_t row = (size_t)t * B + b;
      __half* dst = crop_hl.p + (((row >> 7) * (size_t)crop_hl.nsl + (size_t)(g0 >> 4)) * 128 + (row & 127)) * 16 + (g0 & 15);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(s_hi + t * G + g0);
      *reinterpret_cast<uint4*>(dst + crop_hl.plane) = *reinterpret_cast<const uint4*>(s_lo + t * G + g0);
    }
    READ_STAMP(4);
    return;
  }
  const int dr = NT / w, dc = NT - dr * w;
  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)t * B + b;
    float* cf = crop ? crop + row * (size_t)G : nullptr;
    const Tap* txs = s_tx + t * w;
    const Tap* tys = s_ty + t * h;
    int r = (int)threadIdx.x / w, c = (int)threadIdx.x - r * w;
    for (int g = threadIdx.x; g < G; g += NT) {
      const Tap tx = txs[c], ty = tys[r];
      float v = 0.f;
      if (((tx.wf != 0.f) | (tx.wc != 0.f)) & ((ty.wf != 0.f) | (ty.wc != 0.f))) v = bilinear_pre(s_img, tx, ty);
      if (cf) cf[g] = v;
      if (crop_hl.p) hl_store(crop_hl, row, g, v);
      c += dc;
      r += dr;
      if (c >= w) {
        c -= w;
        ++r;
      }
    }
  }
}

// stand-alone forward STN with explicit where codes (air_stn_read); dynamic smem: H*W floats + (w+h) taps
__global__ void __launch_bounds__(256)
stn_read_kernel(const float* __restrict__ img, const float* __restrict__ where, float* __restrict__ crop, int H,
                int W, int h, int w) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int b = blockIdx.x;
  const int P = H * W, G = h * w;
  float* s_img = reinterpret_cast<float*>(smem_raw);
  Tap* s_tx = reinterpret_cast<Tap*>(smem_raw + (sizeof(float) * ((size_t)P + 4) / 16 * 16 + 16));
  Tap* s_ty = s_tx + w;
  for (int i = threadIdx.x; i < P; i += blockDim.x) s_img[i] = img[(size_t)b * P + i];
  const float sx = where[b * 4 + 0], tx = where[b * 4 + 1], sy = where[b * 4 + 2], ty = where[b * 4 + 3];
  for (int j = threadIdx.x; j < w + h; j += blockDim.x) {
    if (j < w) s_tx[j] = make_tap(fwd_coord(sx, tx, j, w, W), W);
    else       s_ty[j - w] = make_tap(fwd_coord(sy, ty, j - w, h, H), H);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const int r = g / w, c = g - r * w;
    crop[(size_t)b * G + g] = bilinear(s_img, W, s_tx[c], s_ty[r]);
  }
}

// ---------------------------------------------------------------------------------------------------
// what head: ParametrisedGaussian (modules.py:11-24, cell.py:154-156)
//   r[rows, 2na] -> loc = r[:, :na]; scale = softplus(r[:, na:] + offset); what = eps * scale + loc
// ---------------------------------------------------------------------------------------------------
__global__ void what_kernel(const float* __restrict__ r, const float* __restrict__ eps, float* __restrict__ what,
                            float* __restrict__ what_loc, float* __restrict__ what_scale, size_t rows, int na,
                            float offset, HlOut what_hl) {
  griddep_launch();
  griddep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * (size_t)na) return;
  const size_t row = idx / na;
  const int j = (int)(idx % na);
  const float loc = r[row * 2 * na + j];
  const float sc = softplus_f(r[row * 2 * na + na + j] + offset);
  what_loc[idx] = loc;
  what_scale[idx] = sc;
  const float wv = __fadd_rn(__fmul_rn(eps[idx], sc), loc);
  what[idx] = wv;
  if (what_hl.p) hl_store(what_hl, row, j, wv);
}

// ---------------------------------------------------------------------------------------------------
// Normal || Normal KL [upstream tf.contrib.distributions _kl_normal_normal], model.py:177-209
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float normal_kl(float mu_a, float s_a, float mu_b, float s_b) {
  const float sa2 = __fmul_rn(s_a, s_a);
  const float sb2 = __fmul_rn(s_b, s_b);
  const float ratio = __fdiv_rn(sa2, sb2);
  const float d = mu_a - mu_b;
  const float t1 = __fdiv_rn(__fmul_rn(d, d), __fmul_rn(2.0f, sb2));
  const float t2 = __fmul_rn(0.5f, (ratio - 1.0f) - logf(ratio));
  return __fadd_rn(t1, t2);
}

struct ElboArgs {
  // inputs
  const float* img;            // [B,P] obs
  const float* glimpse;        // [T,B,G] raw decoder output
  const float* where;          // [T,B,4]
  const float* where_loc;      // [T,B,4]
  const float* where_scale;    // [T,B,4]
  const float* what_loc;       // [T,B,na]
  const float* what_scale;     // [T,B,na]
  const float* presence;       // [T,B]
  const float* presence_prob;  // [T,B]
  const float* canvas_in;      // [B,P] or null: raw canvas carried in (air_cell_step)
  // outputs
  float* canvas;               // [T,B,P] or null
  float* glimpse_viz;          // [T,B,G] or null
  float* num_steps_posterior;  // [B,T+1]
  float* num_step_per_sample;  // [B]
  float* prior_step_weight;    // [T,B]
  float* rec_loss_per_sample;
  float* kl_num_steps_per_sample;
  float* kl_what_per_sample;
  float* kl_where_per_sample;
  float* loss_per_sample;
  float* num_steps_log_prob;
  int T, B, H, W, h, w, na;
  float output_std, output_multiplier;
  float lp_const;              // 0.5 log(2 pi) + log(output_std), the constant of Normal.log_prob (host, float64)
  float inv_sigma;             // 1 / output_std (host)
  double step_W, step_H;       // np.linspace(-1, 1, n) step 2/(n-1) of the canvas columns / rows (host, float64)
  int do_elbo;
  float* prior_part;           // [B] scratch: prior_weight * prior_per_sample (prior CTAs -> elbo_scalars_kernel)
  int n_prior_ctas;            // leading CTAs of the paint grid that compute the prior terms (launch_paint_elbo)
  long long t_stride;          // B * H * W: elements between the canvases of consecutive steps (host)
  int separable;               // paint_use_separable(T, H, W, h, w) (host)
  int col_ni;                  // column pass: glimpse rows per thread, ceil(h / col_rg) (host)
  int col_cp, col_rg;          // column pass: columns resident in one sweep min(W, threads), glimpse rows per sweep (host)
  int row_tprb, row_rpp;       // row pass: threads per canvas row resident in one pass, rows per pass (host)
  long long* trace;            // debug (AIR_PAINT_TRACE): 8 stamps per CTA (globaltimer ns at the phase boundaries, SM id)
  PresenceArgs scan;           // scan.logit != null: presence / presence_prob are not inputs -- this grid runs the presence scan
                               // itself (the fused row kernel has no per-canvas CTA to do it): the paint CTA of a canvas
                               // writes the two outputs, the prior warp of the canvas recomputes the same values locally
  air_prior prior;
  double steps_prior[AIR_MAX_STEPS + 1];   // geometric_prior(success_prob, T) (prior.py:26-32): the same table for
                                           // every canvas, computed once on the host (air_api.cu:steps_prior_table)
  const double* steps_prior_dev;           // non-null: the table is read from device memory instead (air_prior_table_device:
                                           // an annealed prior under a replayed CUDA graph, whose kernel arguments are frozen)
};

// prior.py:26-32 geometric_prior in float64 [upstream Geometric(probs=1-s).prob(k) = exp(k*log1p(-probs) + log(probs))]
__host__ __device__ inline double geom_prior_f64(double s, int k) {
  s = fmin(fmax(s, 1e-7), 1.0 - 1e-15);
  const double probs = 1.0 - s;
  return exp((double)k * log1p(-probs) + log(probs));
}
// same in float32 (python-float success probability -> tf.float32 graph constants)
__host__ __device__ inline float geom_prior_f32(float s, int k) {
  s = fminf(fmaxf(s, 1e-7f), (float)(1.0 - 1e-15));
  const float probs = 1.0f - s;
  return expf((float)k * log1pf(-probs) + logf(probs));
}

// prior.py:62-68 bernoulli_to_modified_geometric for one row (float64 island); p has T entries, q gets T+1.
__device__ __forceinline__ void modified_geometric_row(const float* p, int stride, int T, float* q) {
  double pi[AIR_MAX_STEPS + 1];
  double cum = 1.0, sum = 0.0;
  for (int k = 0; k <= T; ++k) {
    double v;
    if (k < T) {
      const double pk = (double)p[(size_t)k * stride];
      v = (1.0 - pk) * cum;     // inv[k] * prod_{j<k} p_j   (k = 0: cum = 1)
      cum *= pk;
    } else {
      v = cum;
    }
    pi[k] = v;
    sum += v;
  }
  for (int k = 0; k <= T; ++k) q[k] = (float)(pi[k] / sum);
}

// prior.py:71-90 tabular_kl entry: float32(p * log(p / q)) in float64 where p > zero, else exactly 0
__device__ __forceinline__ float tabular_kl_entry(float p, double q, double zero_prob_value) {
  const double pd = (double)p;
  if (!(pd > zero_prob_value)) return 0.f;
  return (float)(pd * log(pd / q));
}

// ---------------------------------------------------------------------------------------------------
// Per-canvas prior terms (model.py:126-216, prior.py:62-90,148-151).  A group of L = 8 lanes per canvas (four canvases
// per warp), no shared memory:
//   q(n) (float64 island, lane k owns n = k), KL(q(n) || prior), log q(n_b), the per-step weights,
//   KL(what) (lanes over the na latents, one group reduction per step), KL(where) (lane t owns step t).
// (Round 1 used a whole warp per canvas; most of the work is scalar per canvas, so 31 lanes replayed it: 1560 warp
// instructions per canvas, 15 % of the paint grid's issue slots.  One THREAD per canvas needs the fewest instructions
// but its dependent chain -- 150 KL terms with two IEEE divisions and a log each -- runs for 100 us.)
// Runs inside the paint grid (its leading CTAs) and leaves prior_weight * prior_per_sample in prior_part[b];
// elbo_scalars_kernel adds the reconstruction term on top (Loss.add, ops.py:12-29).  `finalize` != 0 (air_prior_terms,
// stand-alone kernel): there is no canvas, the reconstruction term is exactly 0 and loss_per_sample is complete.
// ---------------------------------------------------------------------------------------------------
// lanes of a warp that share one canvas: T + 1 <= L (lane k owns n = k)
template <int T>
struct PriorGroup { static constexpr int L = (T + 1 <= 8) ? 8 : 16; };

template <int T>
__device__ __forceinline__ void prior_terms_group(const ElboArgs& a, int b_in, int finalize) {
  constexpr int L = PriorGroup<T>::L;
  const int B = a.B;
  const int lane = threadIdx.x & (L - 1);
  const bool valid = b_in < B;             // every lane of the warp takes part in the shuffles; only valid groups store
  const int b = valid ? b_in : B - 1;
  const air_prior& pr = a.prior;
  float pp[T], pz[T];   // presence_prob / presence of this canvas
  if (a.scan.logit) {
    float carry = a.scan.presence_in ? a.scan.presence_in[b] : 1.0f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      presence_step(a.scan, (size_t)t * B + b, carry, pp[t]);
      pz[t] = carry;
    }
  } else {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      pp[t] = a.presence_prob[(size_t)t * B + b];
      pz[t] = a.presence[(size_t)t * B + b];
    }
  }
  // step-count posterior q(n): lane k owns n = k (float64 island)
  const int k = lane;
  double pi = 0.0;
  if (k <= T) {
    double cum = 1.0;
#pragma unroll
    for (int j = 0; j < T; ++j)
      if (j < k) cum *= (double)pp[j];
    double pk = 0.0;
#pragma unroll
    for (int j = 0; j < T; ++j)
      if (j == k) pk = (double)pp[j];
    pi = (k < T) ? (1.0 - pk) * cum : cum;
  }
  double sum = pi;
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float q = 0.f, kl = 0.f;
  if (k <= T) {
    q = (float)(pi / sum);
    kl = tabular_kl_entry(q, a.steps_prior_dev ? a.steps_prior_dev[k] : a.steps_prior[k], 0.0);
    if (valid) a.num_steps_posterior[(size_t)b * (T + 1) + k] = q;
  }
  float kl_n = 0.f;   // fp32 sum over n in index order (model.py:149)
  float qs[T + 1];
#pragma unroll
  for (int j = 0; j <= T; ++j) {
    const float v = __shfl_sync(0xffffffffu, kl, j, L);
    kl_n = (j == 0) ? v : __fadd_rn(kl_n, v);
    qs[j] = __shfl_sync(0xffffffffu, q, j, L);
  }
  // KL(what) per step (model.py:174-186): lanes over the latents
  float klw[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float v = 0.f;
    const size_t base = ((size_t)t * B + b) * a.na;
    for (int i = lane; i < a.na; i += L)
      v += normal_kl(a.what_loc[base + i], a.what_scale[base + i], pr.what_loc, pr.what_scale);
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    klw[t] = v;
  }
  // KL(where) per step (model.py:188-214): lane t owns step t; (sx, sy) vs scale prior, (tx, ty) vs shift prior
  float klwh_l = 0.f;
  if (lane < T) {
    const float* wlp = a.where_loc + ((size_t)lane * B + b) * 4;
    const float* wsp = a.where_scale + ((size_t)lane * B + b) * 4;
    const float4 wl = make_float4(wlp[0], wlp[1], wlp[2], wlp[3]);
    const float4 ws = make_float4(wsp[0], wsp[1], wsp[2], wsp[3]);
    const float k_sx = normal_kl(wl.x, ws.x, pr.where_scale_loc, pr.where_scale_scale);
    const float k_sy = normal_kl(wl.z, ws.z, pr.where_scale_loc, pr.where_scale_scale);
    const float k_tx = normal_kl(wl.y, ws.y, pr.where_shift_has_loc ? pr.where_shift_loc : wl.y, pr.where_shift_scale);
    const float k_ty = normal_kl(wl.w, ws.w, pr.where_shift_has_loc ? pr.where_shift_loc : wl.w, pr.where_shift_scale);
    klwh_l = __fadd_rn(__fadd_rn(k_sx, k_tx), __fadd_rn(k_sy, k_ty));
  }
  float klwh[T], pres[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    klwh[t] = __shfl_sync(0xffffffffu, klwh_l, t, L);
    pres[t] = pz[t];
  }
  if (lane != 0 || !valid) return;
  a.kl_num_steps_per_sample[b] = kl_n;
  float n = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) n += pres[t];
  a.num_step_per_sample[b] = n;
  int idx = (int)n;
  idx = idx < 0 ? 0 : (idx > T ? T : idx);
  float q_sel = qs[0];
#pragma unroll
  for (int j = 1; j <= T; ++j) q_sel = (idx == j) ? qs[j] : q_sel;
  a.num_steps_log_prob[b] = logf(fmaxf(q_sel, 1e-32f));
  float sw[T];
  if (pr.analytic) {   // reverse cumsum of q(n)[1:]   (model.py:157-161)
    float cs = 0.f;
#pragma unroll
    for (int t = T - 1; t >= 0; --t) {
      cs = (t == T - 1) ? qs[t + 1] : __fadd_rn(cs, qs[t + 1]);
      sw[t] = cs;
    }
  } else {
#pragma unroll
    for (int t = 0; t < T; ++t) sw[t] = pres[t];
  }
  float kl_what = 0.f, kl_where = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    a.prior_step_weight[(size_t)t * B + b] = sw[t];
    const float v = __fmul_rn(klw[t], sw[t]);
    kl_what = (t == 0) ? v : __fadd_rn(kl_what, v);
    const float u = __fmul_rn(klwh[t], sw[t]);
    kl_where = (t == 0) ? u : __fadd_rn(kl_where, u);
  }
  a.kl_what_per_sample[b] = kl_what;
  a.kl_where_per_sample[b] = kl_where;
  // Loss.add bookkeeping (ops.py:12-29; model.py:154,186,214,332): the prior part of the per-sample loss
  const float prior_ps = __fadd_rn(__fadd_rn(__fmul_rn(kl_n, pr.steps_weight), kl_what), kl_where);
  const float part = __fmul_rn(prior_ps, pr.use_prior ? 1.0f : 0.0f);
  if (finalize) {
    a.rec_loss_per_sample[b] = 0.f;
    a.loss_per_sample[b] = __fadd_rn(0.f, part);
  } else {
    a.prior_part[b] = part;
  }
}

template <int T>
__global__ void __launch_bounds__(128) prior_terms_kernel(ElboArgs a, int finalize) {
  griddep_launch();
  griddep_wait();
  prior_terms_group<T>(a, (int)((blockIdx.x * blockDim.x + threadIdx.x) / PriorGroup<T>::L), finalize);
}

inline cudaError_t launch_prior_terms(const ElboArgs& a, int finalize, cudaStream_t st) {
  const int L = (a.T + 1 <= 8) ? 8 : 16;       // PriorGroup<T>::L
  const dim3 grid((a.B * L + 127) / 128), block(128);
  switch (a.T) {
    case 1: return launch_k(prior_terms_kernel<1>, grid, block, 0, st, a, finalize);
    case 2: return launch_k(prior_terms_kernel<2>, grid, block, 0, st, a, finalize);
    case 3: return launch_k(prior_terms_kernel<3>, grid, block, 0, st, a, finalize);
    case 4: return launch_k(prior_terms_kernel<4>, grid, block, 0, st, a, finalize);
    case 5: return launch_k(prior_terms_kernel<5>, grid, block, 0, st, a, finalize);
    case 6: return launch_k(prior_terms_kernel<6>, grid, block, 0, st, a, finalize);
    case 7: return launch_k(prior_terms_kernel<7>, grid, block, 0, st, a, finalize);
    case 8: return launch_k(prior_terms_kernel<8>, grid, block, 0, st, a, finalize);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------------------------------
// paint + reconstruction term.  One CTA per canvas b.
//   canvas_t = canvas_{t-1} + presence_t * resampler(glimpse_t, inverse warp(where_t))   (cell.py:159-164)
//   rec[b]   = sum_px 0.5 ((x - mu)/sigma)^2 + log sigma + 0.5 log 2 pi, mu = multiplier * canvas_T (model.py:319-321)
//   (loss_per_sample[b] = rec[b] + prior part is completed by elbo_scalars_kernel)
// The canvas is never read back from HBM: it accumulates in registers across the T steps and is written once per step.
// The bilinear resampler is separable -- canvas[r][c] = sum_i wy[r][i] (sum_j wx[c][j] glimpse[i][j]) with two non-zero
// taps per axis -- and the kernel was issue-bound (ncu r02a: 310 warp instructions per 32 pixel pairs, 68 % issue
// utilisation, 30 % of DRAM), so it runs in two passes over shared memory:
//   columns: s_col[t][i][c] = presence_t * (wx_f glimpse[i][j_f] + wx_c glimpse[i][j_c])     T*h*W values, 2 taps each
//   rows:    canvas_t[r][c] += wy_f s_col[t][i_f][c] + wy_c s_col[t][i_c][c]                  T*H*W values, 2 taps each
// The row pass touches no tap table of the columns and no mask: a canvas pixel costs two 8-byte shared loads and four
// FMAs per step.  (Association differs from the four-product form of the stand-alone resampler by a few ulp.)
// Thread mapping of the row pass: a thread owns a fixed column (a column PAIR when W is even: 8-byte stores) and walks
// down the rows; in pass k the CTA's threads cover one contiguous run of W * rows_per_pass pixels, so every store
// instruction is fully coalesced.
// dynamic smem: T*G floats (glimpses) + T*(W+H) taps + T*h*W floats (column pass).
// ---------------------------------------------------------------------------------------------------
// direct form (large canvases): glimpses + tap tables
__host__ __device__ inline size_t paint_smem_direct(int T, int H, int W, int h, int w) {
  return (sizeof(float) * (size_t)T * h * w + 15) / 16 * 16 + sizeof(Tap) * (size_t)T * (W + H);
}
// separable form: + the column pass's T*h*W floats
__host__ __device__ inline size_t paint_smem_separable(int T, int H, int W, int h, int w) {
  return paint_smem_direct(T, H, W, h, w) + sizeof(float) * (size_t)T * h * W;
}
// The separable form needs fewer instructions but T*h*W more floats of shared memory per CTA; it is used while at least six
// CTAs still fit on an SM (50x50 / 20x20 / T = 3: 21.6 KB).  At 100x100 / 28x28 / T = 5 it would need 88 KB -- two CTAs per
// SM, measured 191 us against 155 us for the direct form (four bilinear taps per pixel straight from the glimpse).
__host__ __device__ inline bool paint_use_separable(int T, int H, int W, int h, int w) {
  return paint_smem_separable(T, H, W, h, w) <= 36 * 1024;
}
__host__ __device__ inline size_t paint_smem(int T, int H, int W, int h, int w) {
  return paint_use_separable(T, H, W, h, w) ? paint_smem_separable(T, H, W, h, w) : paint_smem_direct(T, H, W, h, w);
}

#ifndef PAINT_MIN_CTAS
#define PAINT_MIN_CTAS 8
#endif
// column pairs in the row pass need an even row pitch and 8-byte aligned rows
__host__ __device__ inline bool paint_pairs(const ElboArgs& a) {
  return ((a.W & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.canvas) & 7) == 0) &&
         ((reinterpret_cast<uintptr_t>(a.canvas_in) & 7) == 0) && ((reinterpret_cast<uintptr_t>(a.img) & 7) == 0);
}

// Shape policy of the paint kernel.  PaintRt reads the shape and the host-computed loop shapes from the arguments; PaintFx
// makes them compile-time constants for one (canvas, glimpse, threads) combination -- every index product, trip count and
// divisor then folds, the short loops unroll -- and is selected at launch when the shapes match (the quoted configuration).
struct PaintRt {
  static constexpr bool fixed = false;
  __device__ static int H(const ElboArgs& a) { return a.H; }
  __device__ static int W(const ElboArgs& a) { return a.W; }
  __device__ static int h(const ElboArgs& a) { return a.h; }
  __device__ static int w(const ElboArgs& a) { return a.w; }
  __device__ static int NT(const ElboArgs&) { return (int)blockDim.x; }
  __device__ static int col_cp(const ElboArgs& a) { return a.col_cp; }
  __device__ static int col_rg(const ElboArgs& a) { return a.col_rg; }
  __device__ static int col_ni(const ElboArgs& a) { return a.col_ni; }
  __device__ static int row_tprb(const ElboArgs& a) { return a.row_tprb; }
  __device__ static int row_rpp(const ElboArgs& a) { return a.row_rpp; }
};
template <int FH, int FW, int Fh, int Fw, int FNT>
struct PaintFx {   // column pairs assumed (FW even, aligned pointers: checked at launch)
  static constexpr bool fixed = true;
  static constexpr int CP = FW < FNT ? FW : FNT, RG = FNT / CP, NI = (Fh + RG - 1) / RG;
  static constexpr int TPRB = (FW / 2) < FNT ? (FW / 2) : FNT, RPP = FNT / TPRB;
  __device__ static constexpr int H(const ElboArgs&) { return FH; }
  __device__ static constexpr int W(const ElboArgs&) { return FW; }
  __device__ static constexpr int h(const ElboArgs&) { return Fh; }
  __device__ static constexpr int w(const ElboArgs&) { return Fw; }
  __device__ static constexpr int NT(const ElboArgs&) { return FNT; }
  __device__ static constexpr int col_cp(const ElboArgs&) { return CP; }
  __device__ static constexpr int col_rg(const ElboArgs&) { return RG; }
  __device__ static constexpr int col_ni(const ElboArgs&) { return NI; }
  __device__ static constexpr int row_tprb(const ElboArgs&) { return TPRB; }
  __device__ static constexpr int row_rpp(const ElboArgs&) { return RPP; }
};

// column pass: s_col[t][i][c] for every glimpse row i and canvas column c (zero outside the footprint / absent steps)
template <int T, class D>
__device__ __forceinline__ void paint_columns(const ElboArgs& a, const float* __restrict__ s_gl,
                                              const Tap* __restrict__ s_tx, const float* __restrict__ s_pres,
                                              float* __restrict__ s_col) {
  const int W = D::W(a), h = D::h(a), w = D::w(a);
  const int CP = D::col_cp(a);                  // columns resident in one sweep
  const int RG = D::col_rg(a);                  // glimpse rows per sweep
  int cslot, islot;
  small_divmod((int)threadIdx.x, CP, islot, cslot);
  if (islot >= RG) return;
#pragma unroll 1
  for (int c = cslot; c < W; c += CP) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float pres = s_pres[t];
      if (pres == 0.f) continue;                // the row pass skips absent steps as well
      const Tap tx = s_tx[t * W + c];
      const float pm = __fmul_rn(pres, a.output_multiplier);   // the row pass accumulates multiplier * canvas
      const float wf = __fmul_rn(pm, tx.wf), wc = __fmul_rn(pm, tx.wc);
      const float* g = s_gl + t * h * w;
      float* dst = s_col + (size_t)t * h * W + c;
      const float* gf = g + islot * w + tx.i_f;
      const float* gc = g + islot * w + tx.i_c;
      float* d = dst + islot * W;
      const int n_i = D::col_ni(a), gstep = RG * w, dstep = RG * W;
#pragma unroll 4
      for (int k = 0; k < n_i; ++k) {
        if (islot + k * RG < h) d[k * dstep] = fmaf(wc, gc[k * gstep], __fmul_rn(wf, gf[k * gstep]));
      }
    }
  }
}

// row pass.  s_ty is laid out [H][T] (the T taps of a canvas row are adjacent: one base pointer, immediate offsets) and its
// indices are shared-window BYTE ADDRESSES of s_col rows, the step's slab included; i_f < 0 marks a (row, step) that paints
// nothing (outside the footprint, or the step is absent).  Every address in the loop is a pointer induction variable.
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// FAST: the hot configuration (canvases written, reconstruction term wanted, no initial canvas) with those three facts as
// compile-time constants -- no predicates, no dead pointer arithmetic in the loop
template <int T, int CPT, bool FAST, class D>
__device__ __forceinline__ float paint_rows(const ElboArgs& a, int b, const Tap* __restrict__ s_ty) {
  const int H = D::H(a), W = D::W(a);
  const int TPRB = D::row_tprb(a);              // threads per row (W / CPT) resident in one pass
  const int RPP = D::row_rpp(a);                // rows per pass
  int cslot, rslot;
  small_divmod((int)threadIdx.x, TPRB, rslot, cslot);
  const float mult = a.output_multiplier;
  const bool do_elbo = FAST ? true : a.do_elbo != 0;
  const bool has_cin = FAST ? false : a.canvas_in != nullptr, has_dst = FAST ? true : a.canvas != nullptr;
  const long long tstride = a.t_stride;
  const size_t base = (size_t)b * H * W;
  const int dp = RPP * W;
  float rec = 0.f;
  if (rslot >= RPP) return rec;
#pragma unroll 1
  for (int c = cslot * CPT; c < W; c += TPRB * CPT) {
    const uint32_t c4 = (uint32_t)c * 4u;
    const Tap* typ = s_ty + rslot * T;
    const size_t off = base + rslot * W + c;
    const float* obs = a.img + off;
    const float* cin = a.canvas_in + off;        // only dereferenced when has_cin / has_dst
    float* dst = a.canvas + off;
    // the observation of the NEXT row is fetched one iteration ahead (its latency would otherwise sit in front of the
    // residual at the end of every iteration)
    float xn[CPT];
    if (CPT == 2) {
      const float2 xv = do_elbo ? *reinterpret_cast<const float2*>(obs) : make_float2(0.f, 0.f);
      xn[0] = xv.x; xn[CPT - 1] = xv.y;
    } else {
      xn[0] = do_elbo ? obs[0] : 0.f;
    }
#pragma unroll 1
    for (int r = rslot; r < H; r += RPP) {
      float acc[CPT], xo[CPT];
      xo[0] = xn[0]; xo[CPT - 1] = xn[CPT - 1];
      obs += dp;
      if (do_elbo && r + RPP < H) {
        if (CPT == 2) {
          const float2 xv = *reinterpret_cast<const float2*>(obs);
          xn[0] = xv.x; xn[CPT - 1] = xv.y;
        } else {
          xn[0] = obs[0];
        }
      }
      if (CPT == 2) {
        const float2 ci = has_cin ? *reinterpret_cast<const float2*>(cin) : make_float2(0.f, 0.f);
        acc[0] = __fmul_rn(ci.x, mult); acc[CPT - 1] = __fmul_rn(ci.y, mult);
      } else {
        acc[0] = has_cin ? __fmul_rn(cin[0], mult) : 0.f;
      }
      float* d = dst;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const Tap ty = typ[t];
        if (ty.i_f >= 0) {
          if (CPT == 2) {
            const float2 u = lds_f2((uint32_t)ty.i_f + c4), v = lds_f2((uint32_t)ty.i_c + c4);
            acc[0] = __fadd_rn(acc[0], fmaf(ty.wc, v.x, __fmul_rn(ty.wf, u.x)));
            acc[CPT - 1] = __fadd_rn(acc[CPT - 1], fmaf(ty.wc, v.y, __fmul_rn(ty.wf, u.y)));
          } else {
            acc[0] = __fadd_rn(acc[0], fmaf(ty.wc, lds_f1((uint32_t)ty.i_c + c4), __fmul_rn(ty.wf, lds_f1((uint32_t)ty.i_f + c4))));
          }
        }
        if (has_dst) {
          if (CPT == 2) *reinterpret_cast<float2*>(d) = make_float2(acc[0], acc[CPT - 1]);
          else          d[0] = acc[0];
          d += tstride;
        }
      }
      if (do_elbo) {   // sum of squared residuals; the constants of Normal.log_prob are applied once per canvas
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float dd = xo[j] - acc[j];
          rec = fmaf(dd, dd, rec);
        }
      }
      if (!FAST) cin += dp;
      dst += dp;
      typ += RPP * T;
    }
  }
  return rec;
}

// direct form of the row pass (round 1): four taps per pixel from the staged glimpse; s_tx [T][W], s_ty [T][H] with indices
// pre-multiplied by the glimpse pitch w
template <int T, int CPT, class D>
__device__ __forceinline__ float paint_rows_direct(const ElboArgs& a, int b, const float* __restrict__ s_gl,
                                            const Tap* __restrict__ s_tx, const Tap* __restrict__ s_ty,
                                            const float* __restrict__ s_pres) {
  const int B = a.B, H = D::H(a), W = D::W(a);
  const int P = H * W, G = D::h(a) * D::w(a);
  const int NT = D::NT(a);
  const int TPR = W / CPT;                      // threads per row
  const int TPRB = TPR < NT ? TPR : NT;         // ... resident in one pass
  const int RPP = NT / TPRB;                    // rows per pass
  const int cslot = (int)threadIdx.x % TPRB, rslot = (int)threadIdx.x / TPRB;
  const float mult = a.output_multiplier;
  const bool do_elbo = a.do_elbo != 0;
  float pres[T];
#pragma unroll
  for (int t = 0; t < T; ++t) pres[t] = s_pres[t];
  const float* cin = a.canvas_in ? a.canvas_in + (size_t)b * P : nullptr;
  const float* obs = a.img + (size_t)b * P;
  float* cbase = a.canvas ? a.canvas + (size_t)b * P : nullptr;
  const size_t tstride = (size_t)B * P;
  float rec = 0.f;
  if (rslot >= RPP) return rec;
  for (int c = cslot * CPT; c < W; c += TPRB * CPT) {
    // loop-invariant: which (step, column) pairs lie inside the glimpse footprint (and are present at all)
    uint32_t colmask = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (pres[t] != 0.f) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const Tap tx = s_tx[t * W + c + j];
          if ((tx.wf != 0.f) | (tx.wc != 0.f)) colmask |= 1u << (t * CPT + j);
        }
      }
    }
    for (int r = rslot; r < H; r += RPP) {
      const int p = r * W + c;
      float acc[CPT], xo[CPT];
      if (CPT == 2) {
        const float2 ci = cin ? *reinterpret_cast<const float2*>(cin + p) : make_float2(0.f, 0.f);
        const float2 xv = do_elbo ? *reinterpret_cast<const float2*>(obs + p) : make_float2(0.f, 0.f);
        acc[0] = ci.x; acc[CPT - 1] = ci.y;
        xo[0] = xv.x; xo[CPT - 1] = xv.y;
      } else {
        acc[0] = cin ? cin[p] : 0.f;
        xo[0] = do_elbo ? obs[p] : 0.f;
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const uint32_t cm = (colmask >> (t * CPT)) & ((1u << CPT) - 1u);
        if (cm) {
          const Tap ty = s_ty[t * H + r];
          if ((ty.wf != 0.f) | (ty.wc != 0.f)) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              if (cm & (1u << j)) {
                const Tap tx = s_tx[t * W + c + j];
                const float v = bilinear_pre(s_gl + t * G, tx, ty);
                acc[j] = __fadd_rn(acc[j], __fmul_rn(pres[t], v));
              }
            }
          }
        }
        if (cbase) {
          float* dst = cbase + (size_t)t * tstride + p;
          if (CPT == 2) *reinterpret_cast<float2*>(dst) = make_float2(__fmul_rn(acc[0], mult), __fmul_rn(acc[CPT - 1], mult));
          else          dst[0] = __fmul_rn(acc[0], mult);
        }
      }
      if (do_elbo) {   // sum of squared residuals; the constants of Normal.log_prob are applied once per canvas
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float d = xo[j] - __fmul_rn(acc[j], mult);
          rec = fmaf(d, d, rec);
        }
      }
    }
  }
  return rec;
}

template <int T, class D>
__global__ void __launch_bounds__(256, PAINT_MIN_CTAS) paint_elbo_kernel(ElboArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ float s_pres[AIR_MAX_STEPS], s_red[32];
  __shared__ float4 s_inv[AIR_MAX_STEPS];
  const int B = a.B, H = D::H(a), W = D::W(a), h = D::h(a), w = D::w(a);
  const int NT = D::NT(a);
  const int G = h * w;
  float* s_gl = reinterpret_cast<float*>(smem_raw);                                                      // [T][G]
  Tap* s_tx = reinterpret_cast<Tap*>(smem_raw + (sizeof(float) * (size_t)T * G + 15) / 16 * 16);         // [T][W]
  Tap* s_ty = s_tx + (size_t)T * W;                                                                      // [H][T]
  float* s_col = reinterpret_cast<float*>(s_ty + (size_t)T * H);                                         // [T][h][W]

  griddep_launch();
  griddep_wait();
  long long* tr = a.trace ? a.trace + (size_t)blockIdx.x * 8 : nullptr;
#define PAINT_STAMP(i) do { if (tr && threadIdx.x == 0) tr[i] = (long long)globaltimer_ns(); } while (0)
  PAINT_STAMP(0);
  // The first n_prior_ctas CTAs of the grid compute the prior terms (8 lanes per canvas) while the others paint: the
  // latency-bound float64 / log chains hide behind the paint CTAs instead of costing a launch.
  if ((int)blockIdx.x < a.n_prior_ctas) {
    prior_terms_group<T>(a, (int)((blockIdx.x * NT + threadIdx.x) / PriorGroup<T>::L), 0);
    PAINT_STAMP(6);
    return;
  }
  const int b = blockIdx.x - a.n_prior_ctas;
  // stage the T decoded glimpses of this canvas (TMA bulk copies, one mbarrier)
  const bool bulk = bulk_ok(a.glimpse, G);
  if (threadIdx.x == 0 && bulk) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    mbar_expect_tx(&bar, (uint32_t)(T * G) * 4u);
    for (int t = 0; t < T; ++t)
      bulk_g2s(s_gl + (size_t)t * G, a.glimpse + ((size_t)t * B + b) * G, (uint32_t)G * 4u, &bar);
  }
  if (!bulk)
    for (int i = threadIdx.x; i < T * G; i += NT) {
      const int t = i / G, g = i - t * G;
      s_gl[i] = a.glimpse[((size_t)t * B + b) * G + g];
    }
  // the T inverse transforms of this canvas (two divisions each), one thread per step
  if (threadIdx.x < T) {
    const float* wh = a.where + ((size_t)threadIdx.x * B + b) * 4;
    float4 iv;   // (a', d', -tx', -ty')
    inv_params(wh[0], wh[1], wh[2], wh[3], iv.x, iv.y, iv.z, iv.w);
    s_inv[threadIdx.x] = iv;
    if (!a.scan.logit) s_pres[threadIdx.x] = a.presence[(size_t)threadIdx.x * B + b];
  }
  if (a.scan.logit && threadIdx.x == 32) {   // presence scan of this canvas (cell.py:137-151)
    float carry = a.scan.presence_in ? a.scan.presence_in[b] : 1.0f;
    for (int t = 0; t < T; ++t) {
      const size_t i = (size_t)t * B + b;
      float p;
      presence_step(a.scan, i, carry, p);
      a.scan.presence_prob[i] = p;
      a.scan.presence[i] = carry;
      s_pres[t] = carry;
    }
  }
  __syncthreads();
  PAINT_STAMP(1);
  // inverse-warp tap tables while the copy is in flight: glimpse-space taps of every canvas column / row
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float4 iv = s_inv[t];
    for (int j = threadIdx.x; j < W + H; j += NT) {
      if (j < W) s_tx[t * W + j] = make_tap(inv_coord_s(iv.x, iv.z, j, a.step_W, w), w, 1);
      else if (!a.separable) s_ty[t * H + (j - W)] = make_tap(inv_coord_s(iv.y, iv.w, j - W, a.step_H, h), h, w);
      else {
        Tap tp = make_tap(inv_coord_s(iv.y, iv.w, j - W, a.step_H, h), h, W);
        const bool live = (s_pres[t] != 0.f) & ((tp.wf != 0.f) | (tp.wc != 0.f));
        const int col0 = (int)smem_u32(s_col) + t * h * W * 4;   // shared-window address of step t's slab
        tp.i_f = live ? col0 + tp.i_f * 4 : -1;
        tp.i_c = col0 + tp.i_c * 4;
        s_ty[(j - W) * T + t] = tp;
      }
    }
  }
  __syncthreads();
  PAINT_STAMP(2);
  if (bulk) mbar_wait(&bar, 0);
  PAINT_STAMP(3);
  if (a.separable) paint_columns<T, D>(a, s_gl, s_tx, s_pres, s_col);

  // optional visualisation output: presence * sigmoid(glimpse)   (model.py:90)
  if (a.glimpse_viz) {
    if (((G & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.glimpse_viz) & 15) == 0) && T * (G >> 2) < 2048) {
      const int G4 = G >> 2;                       // 16-byte pieces, all T glimpses in one flat loop
      for (int i = threadIdx.x; i < T * G4; i += NT) {
        int t, g;
        small_divmod(i, G4, t, g);
        const float pres_t = s_pres[t];
        float4 v = reinterpret_cast<const float4*>(s_gl)[i];
        v.x = __fmul_rn(pres_t, sigmoid_lean(v.x)); v.y = __fmul_rn(pres_t, sigmoid_lean(v.y));
        v.z = __fmul_rn(pres_t, sigmoid_lean(v.z)); v.w = __fmul_rn(pres_t, sigmoid_lean(v.w));
        reinterpret_cast<float4*>(a.glimpse_viz + ((size_t)t * B + b) * G)[g] = v;
      }
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float* dst = a.glimpse_viz + ((size_t)t * B + b) * G;
        const float pres_t = s_pres[t];
        for (int g = threadIdx.x; g < G; g += NT) dst[g] = __fmul_rn(pres_t, sigmoid_lean(s_gl[t * G + g]));
      }
    }
  }

  const bool pair = paint_pairs(a);
  __syncthreads();
  PAINT_STAMP(4);
  float rec;
  const bool fast = a.do_elbo && a.canvas && !a.canvas_in;
  if (a.separable && pair && fast) rec = paint_rows<T, 2, true, D>(a, b, s_ty);
  else if (a.separable) rec = pair ? paint_rows<T, 2, false, D>(a, b, s_ty) : paint_rows<T, 1, false, D>(a, b, s_ty);
  else rec = pair ? paint_rows_direct<T, 2, D>(a, b, s_gl, s_tx, s_ty, s_pres) : paint_rows_direct<T, 1, D>(a, b, s_gl, s_tx, s_ty, s_pres);
  PAINT_STAMP(5);
  if (!a.do_elbo) return;
  rec = block_sum(rec, s_red);
  PAINT_STAMP(6);
  if (tr && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[7] = smid;
  }
  // sum_px 0.5 ((x - mu) / sigma)^2 + (log sigma + 0.5 log 2 pi); elbo_scalars_kernel adds the prior part
  if (threadIdx.x == 0)
    a.rec_loss_per_sample[b] = fmaf(__fmul_rn(0.5f * a.inv_sigma, a.inv_sigma), rec, __fmul_rn((float)(H * W), a.lp_const));
}

template <int T>
inline cudaError_t launch_paint_elbo_t(const ElboArgs& a, size_t smem, cudaStream_t st) {
  static const int nt = getenv("AIR_PAINT_THREADS") ? atoi(getenv("AIR_PAINT_THREADS")) : 256;
  static const bool no_fixed = getenv("AIR_PAINT_NO_FIXED") != nullptr;
  // the configuration the metric is quoted on (50x50 canvas, 20x20 glimpse, three steps, 256 threads, separable, column
  // pairs): shapes and loop shapes as compile-time constants
  using Fx = PaintFx<50, 50, 20, 20, 256>;
  if constexpr (T == 5) {   // BASELINE configs[3]: 100x100 canvas, 28x28 glimpse, five steps (direct form)
    using Fx4 = PaintFx<100, 100, 28, 28, 256>;
    if (!no_fixed && nt == 256 && a.H == 100 && a.W == 100 && a.h == 28 && a.w == 28 && !a.separable && paint_pairs(a)) {
      cudaError_t e = ensure_dynamic_smem(paint_elbo_kernel<T, Fx4>, smem);
      if (e != cudaSuccess) return e;
      return launch_k(paint_elbo_kernel<T, Fx4>, dim3(a.B + a.n_prior_ctas), dim3(nt), smem, st, a);
    }
  }
  if constexpr (T == 3) {
    if (!no_fixed && nt == 256 && a.H == 50 && a.W == 50 && a.h == 20 && a.w == 20 && a.separable && paint_pairs(a)) {
      cudaError_t e = ensure_dynamic_smem(paint_elbo_kernel<T, Fx>, smem);
      if (e != cudaSuccess) return e;
      return launch_k(paint_elbo_kernel<T, Fx>, dim3(a.B + a.n_prior_ctas), dim3(nt), smem, st, a);
    }
  }
  cudaError_t e = ensure_dynamic_smem(paint_elbo_kernel<T, PaintRt>, smem);
  if (e != cudaSuccess) return e;
  return launch_k(paint_elbo_kernel<T, PaintRt>, dim3(a.B + a.n_prior_ctas), dim3(nt), smem, st, a);
}
// host-side constants of ElboArgs (float64 maths on the host, exactly what the device code computed per tap before)
inline void fill_elbo_consts(ElboArgs& a) {
  a.inv_sigma = 1.0f / a.output_std;
  a.step_W = a.W > 1 ? 2.0 / (double)(a.W - 1) : 0.0;
  a.step_H = a.H > 1 ? 2.0 / (double)(a.H - 1) : 0.0;
  a.t_stride = (long long)a.B * a.H * a.W;
}
// prior terms (when a prior is given) followed by paint + reconstruction term
inline cudaError_t launch_paint_elbo(ElboArgs& a, cudaStream_t st) {
  fill_elbo_consts(a);
  static const int nt = getenv("AIR_PAINT_THREADS") ? atoi(getenv("AIR_PAINT_THREADS")) : 256;
  // loop shapes of the two passes (integer divisions the kernel would otherwise do per thread)
  a.separable = paint_use_separable(a.T, a.H, a.W, a.h, a.w) ? 1 : 0;
  const bool pair = paint_pairs(a);
  const int tpr = pair ? a.W / 2 : a.W;
  a.col_cp = a.W < nt ? a.W : nt;
  a.col_rg = nt / a.col_cp;
  a.col_ni = (a.h + a.col_rg - 1) / a.col_rg;
  a.row_tprb = tpr < nt ? tpr : nt;
  a.row_rpp = nt / a.row_tprb;
  const int cpc = nt / ((a.T + 1 <= 8) ? 8 : 16);   // canvases per prior CTA (PriorGroup<T>::L lanes each)
  a.n_prior_ctas = a.do_elbo ? (a.B + cpc - 1) / cpc : 0;
  const size_t smem = paint_smem(a.T, a.H, a.W, a.h, a.w);
  switch (a.T) {
    case 1: return launch_paint_elbo_t<1>(a, smem, st);
    case 2: return launch_paint_elbo_t<2>(a, smem, st);
    case 3: return launch_paint_elbo_t<3>(a, smem, st);
    case 4: return launch_paint_elbo_t<4>(a, smem, st);
    case 5: return launch_paint_elbo_t<5>(a, smem, st);
    case 6: return launch_paint_elbo_t<6>(a, smem, st);
    case 7: return launch_paint_elbo_t<7>(a, smem, st);
    case 8: return launch_paint_elbo_t<8>(a, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

// stand-alone inverse STN (air_stn_paint): out[b] = resampler(glimpse[b], inverse warp(where[b]))
// dynamic smem: h*w floats + (W+H) taps
__global__ void __launch_bounds__(256)
stn_paint_kernel(const float* __restrict__ glimpse, const float* __restrict__ where, float* __restrict__ out, int H,
                 int W, int h, int w) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int b = blockIdx.x;
  const int P = H * W, G = h * w;
  float* s_gl = reinterpret_cast<float*>(smem_raw);
  Tap* s_tx = reinterpret_cast<Tap*>(smem_raw + (sizeof(float) * (size_t)G + 15) / 16 * 16);
  Tap* s_ty = s_tx + W;
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_gl[i] = glimpse[(size_t)b * G + i];
  float a_inv, d_inv, ntx, nty;
  inv_params(where[b * 4 + 0], where[b * 4 + 1], where[b * 4 + 2], where[b * 4 + 3], a_inv, d_inv, ntx, nty);
  for (int j = threadIdx.x; j < W + H; j += blockDim.x) {
    if (j < W) s_tx[j] = make_tap(inv_coord(a_inv, ntx, j, W, w), w);
    else       s_ty[j - W] = make_tap(inv_coord(d_inv, nty, j - W, H, h), h);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    out[(size_t)b * P + p] = bilinear(s_gl, w, s_tx[c], s_ty[r]);
  }
}

// model.py:90: the visualisation tensor presence * sigmoid(glimpse), [rows = T * B][G].  The paint kernel writes it when asked
// (ElboArgs.glimpse_viz); this is the same arithmetic as a stand-alone pass for callers that fetch it only now and then --
// the reference's graph computes it only when `air.glimpse` is fetched, which its training loop never does.
__global__ void __launch_bounds__(256)
glimpse_viz_kernel(const float* __restrict__ glimpse, const float* __restrict__ presence, float* __restrict__ out,
                   long long rows, int G) {
  const long long n = rows * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fmul_rn(presence[i / G], sigmoid_lean(glimpse[i]));
}

// ---------------------------------------------------------------------------------------------------
// batch means of the per-sample terms (model.py:103,151,184,212,247-248,322; ops.py:12-29).  One CTA,
// fixed summation order (deterministic).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
elbo_scalars_kernel(const float* __restrict__ rec, const float* __restrict__ kl_n, const float* __restrict__ kl_what,
                    const float* __restrict__ kl_where, const float* __restrict__ nsteps,
                    const float* __restrict__ logq, const float* __restrict__ baseline, float* __restrict__ scalars,
                    int B, air_prior pr, const float* __restrict__ prior_part, float* __restrict__ loss_per_sample) {
  __shared__ float s_part[32][11];
  griddep_launch();
  griddep_wait();
  float s[11] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float r = rec[b], lq = logq[b], k_n = kl_n[b], k_wt = kl_what[b], k_wh = kl_where[b];
    if (prior_part) loss_per_sample[b] = __fadd_rn(r, prior_part[b]);   // Loss.add (ops.py:12-29), model.py:324,332
    float iw = r;   // REINFORCE importance weight (model.py:337-339)
    if (!pr.analytic) iw = __fadd_rn(r, __fadd_rn(__fadd_rn(__fmul_rn(k_n, pr.steps_weight), k_wt), k_wh));
    s[0] += r;
    s[1] += k_n;
    s[2] += k_wt;
    s[3] += k_wh;
    s[4] += nsteps[b];
    s[5] += iw * lq;
    s[6] += lq;
    const float bl = baseline ? baseline[b] : 0.f;
    s[7] += bl;
    s[8] += iw;
    s[9] += iw * iw;
    s[10] += bl * bl;
  }
  // one pass: warp sums -> shared -> the first warp adds the per-warp partials (fixed order: deterministic)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < 11; ++i) {
    s[i] = warp_sum(s[i]);
    if (lane == 0) s_part[wid][i] = s[i];
  }
  __syncthreads();
  if (wid != 0) return;
#pragma unroll
  for (int i = 0; i < 11; ++i) s[i] = warp_sum(lane < nw ? s_part[lane][i] : 0.f);
  if (threadIdx.x == 0) {
    const float inv = 1.0f / (float)B;
    const float m_rec = s[0] * inv, m_kln = s[1] * inv, m_klw = s[2] * inv, m_klwh = s[3] * inv;
    const float prior_loss = m_kln * pr.steps_weight + m_klw + m_klwh;
    const float loss = m_rec + prior_loss * (pr.use_prior ? 1.0f : 0.0f);
    const float m_iwlq = s[5] * inv, m_lq = s[6] * inv, m_base = s[7] * inv;
    // mean over the [B,B] broadcast of (iw_j - baseline_i) * logq_j  ==  mean(iw*logq) - mean(baseline)*mean(logq)
    // with NVIL normalisation (model.py:232-239): iw' = (iw - baseline - shift) * scale
    const float nv_scale = pr.nvil_scale != 0.f ? pr.nvil_scale : 1.0f, nv_shift = pr.nvil_scale != 0.f ? pr.nvil_shift : 0.f;
    const float reinforce = pr.use_reinforce ? nv_scale * (m_iwlq - (m_base + nv_shift) * m_lq) : 0.f;
    scalars[AIR_S_REC_LOSS] = m_rec;
    scalars[AIR_S_KL_NUM_STEPS] = m_kln;
    scalars[AIR_S_KL_WHAT] = m_klw;
    scalars[AIR_S_KL_WHERE] = m_klwh;
    scalars[AIR_S_PRIOR_LOSS] = prior_loss;
    scalars[AIR_S_LOSS] = loss;
    scalars[AIR_S_REINFORCE] = reinforce;
    scalars[AIR_S_OPT_LOSS] = loss + reinforce;
    scalars[AIR_S_NUM_STEP] = s[4] * inv;
    scalars[AIR_S_MEAN_REC_LOGQ] = m_iwlq;
    scalars[AIR_S_MEAN_LOGQ] = m_lq;
    scalars[AIR_S_MEAN_BASELINE] = m_base;
    scalars[AIR_S_MEAN_IW] = s[8] * inv;
    scalars[AIR_S_MEAN_IW2] = s[9] * inv;
    scalars[AIR_S_MEAN_BASELINE2] = s[10] * inv;
    for (int i = AIR_S_MEAN_BASELINE2 + 1; i < AIR_N_SCALARS; ++i) scalars[i] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------
// Importance-weighted bound (BASELINE.json configs[4]; an EXTENSION -- the reference has no IWAE).
// K particles per canvas are K rows of the ordinary forward pass (row = canvas * K + particle, same image, own noise).
//   log w = log p(x | z) + log p(z, n) - log q(z, n | x)
//   log p(x | z)      = -rec_loss                                           (Normal(canvas, sigma), model.py:319-321)
//   log q(z, n | x)   = log q(n) + sum_{t < n} [log N(what_t; loc, scale) + log N(where_t; loc, scale)]
//   log p(z, n)       = log prior(n) + sum_{t < n} [log N(what_t; what prior) + log N(where_t; scale / shift priors)]
// with n = the sampled number of steps (steps whose presence is 1 -- later latents never reach the canvas), q(n) the
// NumStepsDistribution (prior.py:119-151), prior(n) the geometric prior table (prior.py:26-32), Normal priors as in
// multi_mnist.py:49-51.  One warp per row (lanes over the what latents), float64 for the step-count terms.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float normal_logpdf(float x, float mu, float s) {
  const float z = (x - mu) / s;
  return -0.5f * z * z - logf(s) - 0.9189385332046727f;   // 0.5 log(2 pi)
}
struct IwaeArgs {
  const float *what, *what_loc, *what_scale;      // [T, R, na]
  const float *where, *where_loc, *where_scale;   // [T, R, 4]
  const float* presence;                          // [T, R]
  const float* rec_loss;                          // [R]
  const float* log_q_n;                           // [R]  num_steps_log_prob
  float* log_w;                                   // [R]
  int T, R, na;
  air_prior prior;
  double steps_prior[AIR_MAX_STEPS + 1];
};
__global__ void __launch_bounds__(128) iwae_logw_kernel(IwaeArgs a) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= a.R) return;
  const air_prior& pr = a.prior;
  float lq = 0.f, lp = 0.f;
  int n = 0;
  for (int t = 0; t < a.T; ++t) {
    const size_t row = (size_t)t * a.R + r;
    if (a.presence[row] == 0.f) break;   // presence is a cumulative product: once 0 it stays 0 (cell.py:148)
    ++n;
    const size_t base = row * a.na;
    for (int i = lane; i < a.na; i += 32) {
      const float x = a.what[base + i];
      lq += normal_logpdf(x, a.what_loc[base + i], a.what_scale[base + i]);
      lp += normal_logpdf(x, pr.what_loc, pr.what_scale);
    }
    if (lane < 4) {
      const float x = a.where[row * 4 + lane];
      const bool shift = lane & 1;
      lq += normal_logpdf(x, a.where_loc[row * 4 + lane], a.where_scale[row * 4 + lane]);
      lp += normal_logpdf(x, shift ? pr.where_shift_loc : pr.where_scale_loc,
                          shift ? pr.where_shift_scale : pr.where_scale_scale);
    }
  }
  lq = warp_sum(lq);
  lp = warp_sum(lp);
  if (lane == 0) {
    const double log_pn = log(a.steps_prior[n]);
    a.log_w[r] = (float)((double)(-a.rec_loss[r]) + ((double)lp + log_pn) - ((double)lq + (double)a.log_q_n[r]));
  }
}
// bound[b] = logsumexp_k log_w[b*K + k] - log K; *mean = batch mean (one CTA, fixed order)
__global__ void __launch_bounds__(1024)
iwae_reduce_kernel(const float* __restrict__ log_w, float* __restrict__ bound, float* __restrict__ mean, int n, int K) {
  __shared__ float s_red[32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < n; b += blockDim.x) {
    const float* lw = log_w + (size_t)b * K;
    float m = lw[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, lw[k]);
    double sum = 0.0;
    for (int k = 0; k < K; ++k) sum += exp((double)lw[k] - (double)m);
    const float v = (float)((double)m + log(sum) - log((double)K));
    bound[b] = v;
    acc += v;
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0 && mean) *mean = acc / (float)n;
}

// ---------------------------------------------------------------------------------------------------
// prior.py building blocks as stand-alone kernels (unit parity with test/prior_test.py)
// ---------------------------------------------------------------------------------------------------
__global__ void modified_geometric_kernel(const float* __restrict__ probs, float* __restrict__ pmf, int64_t n, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float q[AIR_MAX_STEPS + 1];
  modified_geometric_row(probs + i * T, 1, T, q);
  for (int k = 0; k <= T; ++k) pmf[i * (T + 1) + k] = q[k];
}

__global__ void geometric_prior_kernel(double s, int n_steps, int is_f64, void* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_steps) return;
  if (is_f64) reinterpret_cast<double*>(out)[k] = geom_prior_f64(s, k);
  else        reinterpret_cast<float*>(out)[k] = geom_prior_f32((float)s, k);
}

__global__ void tabular_kl_kernel(const float* __restrict__ p, const double* __restrict__ q, float* __restrict__ kl,
                                  int64_t n, int m, double zero_prob_value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * m) return;
  kl[i] = tabular_kl_entry(p[i], q[i % m], zero_prob_value);
}

__global__ void num_steps_log_prob_kernel(const float* __restrict__ pmf, const float* __restrict__ samples,
                                          float* __restrict__ out, int64_t n, int m, int take_log) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int idx = (int)samples[i];
  idx = idx < 0 ? 0 : (idx >= m ? m - 1 : idx);
  const float pr = pmf[i * m + idx];
  out[i] = take_log ? logf(fmaxf(pr, 1e-32f)) : pr;
}

}  // namespace air
