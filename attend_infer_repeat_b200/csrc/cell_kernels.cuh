// Non-GEMM stages of the AIR cell: LSTM gate math, where/what sampling, presence scan, the
// spatial-transformer glimpse read and the inverse-transformer canvas paint fused with the ELBO terms.
// All HBM-bound byte work: one CTA per canvas, the image / glimpse tile staged once in shared memory
// (cp.async.bulk where the tile is 16-byte sized), coalesced output rows, warp-shuffle reductions.
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

// Stage `n` floats from global into shared: TMA bulk copy when the tile is 16-byte aligned/sized, plain
// coalesced loads otherwise (tiny test shapes such as the 3x3 images of test/cell_test.py).  Ends with a barrier.
__device__ __forceinline__ void stage_tile(float* dst, const float* __restrict__ src, int n, uint64_t* bar,
                                           uint32_t& parity) {
  const bool bulk_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (bulk_ok) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, (uint32_t)n * 4u);
      bulk_g2s(dst, src, (uint32_t)n * 4u, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// LSTM state handling (snt.LSTM [upstream], mnist_model.py:35; cell.py:101-114,126-127)
// ---------------------------------------------------------------------------------------------------
// trainable initial state (h0, c0) [nh] tiled to the batch (cell.py:103)
__global__ void lstm_init_state_kernel(const float* __restrict__ h0, const float* __restrict__ c0,
                                       float* __restrict__ h, float* __restrict__ c, int B, int nh) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * nh) return;
  const int u = (int)(i % nh);
  h[i] = h0[u];
  c[i] = c0[u];
}

// gates[B,4nh] (order i, j, f, o) already hold [x,h] @ W + b.  c <- sig(f + fb) * c + sig(i) * tanh(j);
// h <- tanh(c) * sig(o).
__global__ void lstm_pointwise_kernel(const float* __restrict__ gates, float* __restrict__ c,
                                      float* __restrict__ h_out, int B, int nh, float forget_bias) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * nh) return;
  const size_t b = idx / nh;
  const int u = (int)(idx % nh);
  const float* g = gates + b * 4 * (size_t)nh;
  const float gi = g[u], gj = g[nh + u], gf = g[2 * nh + u], go = g[3 * nh + u];
  const float c_new = __fadd_rn(__fmul_rn(sigmoid_f(gf + forget_bias), c[idx]), __fmul_rn(sigmoid_f(gi), tanhf(gj)));
  c[idx] = c_new;
  h_out[idx] = __fmul_rn(tanhf(c_new), sigmoid_f(go));
}

// ---------------------------------------------------------------------------------------------------
// presence: StepsPredictor sigmoid + explore-eps mix + Bernoulli draw + cumulative product over steps
// (modules.py:119-122, cell.py:137-151)
// ---------------------------------------------------------------------------------------------------
__global__ void presence_kernel(const float* __restrict__ logit /*[T,B]*/, const float* __restrict__ u_pres /*[T,B]*/,
                                const float* __restrict__ presence_in /*[B] or null (=1)*/,
                                float* __restrict__ presence_prob /*[T,B]*/, float* __restrict__ presence /*[T,B]*/,
                                int T, int B, float step_bias, float explore_eps, int discrete) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float pres = presence_in ? presence_in[b] : 1.0f;
  for (int t = 0; t < T; ++t) {
    const size_t i = (size_t)t * B + b;
    float p = sigmoid_f(logit[i] + step_bias);
    if (explore_eps >= 0.f) p = __fadd_rn(explore_eps / 2.0f, __fmul_rn(1.0f - explore_eps, p));
    presence_prob[i] = p;
    if (discrete) {
      const float z = (u_pres[i] < p) ? 1.0f : 0.0f;
      pres *= z;
    } else {
      pres = p;
    }
    presence[i] = pres;
  }
}

// ---------------------------------------------------------------------------------------------------
// where head + glimpse read.  One CTA per canvas b: the image is staged once in shared memory and all
// T glimpses of that canvas are cropped from it (the T where-codes depend only on h_t, never on the crop).
//   m[T,B,8] = transform-estimator MLP output; loc = (sig, tanh, sig, tanh)(m[0:4]) * (max_crop,1,max_crop,1);
//   scale = softplus(m[4:8] + scale_bias); where = eps * scale + loc       (modules.py:41-63, cell.py:129-133)
//   crop[t,b] = resampler(img[b], AffineGridWarper(where))                  (modules.py:104-109, cell.py:135)
// dynamic smem: H*W floats.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
where_read_kernel(const float* __restrict__ m, const float* __restrict__ eps_where, const float* __restrict__ img,
                  float* __restrict__ where, float* __restrict__ where_loc, float* __restrict__ where_scale,
                  float* __restrict__ crop, int T, int B, int H, int W, int h, int w, float max_crop,
                  float scale_bias) {
  extern __shared__ __align__(16) float s_img[];
  __shared__ uint64_t bar;
  __shared__ float s_where[4];
  const int b = blockIdx.x;
  const int P = H * W, G = h * w;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t parity = 0;
  stage_tile(s_img, img + (size_t)b * P, P, &bar, parity);

  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)t * B + b;
    if (threadIdx.x < 4) {
      const int k = threadIdx.x;
      const float mk = m[row * 8 + k];
      const float loc = (k & 1) ? tanhf(mk) : __fmul_rn(max_crop, sigmoid_f(mk));
      const float sc = softplus_f(m[row * 8 + 4 + k] + scale_bias);
      const float wv = __fadd_rn(__fmul_rn(eps_where[row * 4 + k], sc), loc);
      where_loc[row * 4 + k] = loc;
      where_scale[row * 4 + k] = sc;
      where[row * 4 + k] = wv;
      s_where[k] = wv;
    }
    __syncthreads();
    const float sx = s_where[0], tx = s_where[1], sy = s_where[2], ty = s_where[3];
    float* out = crop + row * (size_t)G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const int r = g / w, c = g - r * w;
      const float x = fwd_coord(sx, tx, c, w, W);
      const float y = fwd_coord(sy, ty, r, h, H);
      out[g] = resample_plane(s_img, H, W, x, y);
    }
    __syncthreads();
  }
}

// stand-alone forward STN with explicit where codes (air_stn_read)
__global__ void __launch_bounds__(256)
stn_read_kernel(const float* __restrict__ img, const float* __restrict__ where, float* __restrict__ crop, int H,
                int W, int h, int w) {
  extern __shared__ __align__(16) float s_img[];
  __shared__ uint64_t bar;
  const int b = blockIdx.x;
  const int P = H * W, G = h * w;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t parity = 0;
  stage_tile(s_img, img + (size_t)b * P, P, &bar, parity);
  const float sx = where[b * 4 + 0], tx = where[b * 4 + 1], sy = where[b * 4 + 2], ty = where[b * 4 + 3];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const int r = g / w, c = g - r * w;
    crop[(size_t)b * G + g] = resample_plane(s_img, H, W, fwd_coord(sx, tx, c, w, W), fwd_coord(sy, ty, r, h, H));
  }
}

// ---------------------------------------------------------------------------------------------------
// what head: ParametrisedGaussian (modules.py:11-24, cell.py:154-156)
//   r[rows, 2na] -> loc = r[:, :na]; scale = softplus(r[:, na:] + offset); what = eps * scale + loc
// ---------------------------------------------------------------------------------------------------
__global__ void what_kernel(const float* __restrict__ r, const float* __restrict__ eps, float* __restrict__ what,
                            float* __restrict__ what_loc, float* __restrict__ what_scale, size_t rows, int na,
                            float offset) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * (size_t)na) return;
  const size_t row = idx / na;
  const int j = (int)(idx % na);
  const float loc = r[row * 2 * na + j];
  const float sc = softplus_f(r[row * 2 * na + na + j] + offset);
  what_loc[idx] = loc;
  what_scale[idx] = sc;
  what[idx] = __fadd_rn(__fmul_rn(eps[idx], sc), loc);
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
  int do_elbo;
  air_prior prior;
};

// prior.py:26-32 geometric_prior in float64 [upstream Geometric(probs=1-s).prob(k) = exp(k*log1p(-probs) + log(probs))]
__device__ __forceinline__ double geom_prior_f64(double s, int k) {
  s = fmin(fmax(s, 1e-7), 1.0 - 1e-15);
  const double probs = 1.0 - s;
  return exp((double)k * log1p(-probs) + log(probs));
}
// same in float32 (python-float success probability -> tf.float32 graph constants)
__device__ __forceinline__ float geom_prior_f32(float s, int k) {
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
// paint + ELBO.  One CTA per canvas b.
//   canvas_t = canvas_{t-1} + presence_t * resampler(glimpse_t, inverse warp(where_t))   (cell.py:159-164)
//   rec[b]   = sum_px 0.5 ((x - mu)/sigma)^2 + log sigma + 0.5 log 2 pi, mu = multiplier * canvas_T (model.py:319-321)
//   KL terms per sample (model.py:126-216), q(n) (prior.py:62-68), log q(n_b) (prior.py:148-151)
// The canvas is never read back from HBM: it accumulates in a register per pixel across the T steps and is
// written once per step (coalesced rows).  dynamic smem: T*G (glimpses) + T*(W+H) (inverse-grid tables) floats.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) paint_elbo_kernel(ElboArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t bar;
  __shared__ float s_pres[AIR_MAX_STEPS], s_w[AIR_MAX_STEPS], s_red[32];
  __shared__ float s_q[AIR_MAX_STEPS + 1];
  const int T = a.T, B = a.B, H = a.H, W = a.W, h = a.h, w = a.w;
  const int P = H * W, G = h * w;
  const int b = blockIdx.x;
  float* s_gl = smem;                 // [T][G]
  float* s_x = s_gl + (size_t)T * G;  // [T][W] glimpse-space x of every canvas column
  float* s_y = s_x + (size_t)T * W;   // [T][H]

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // stage the T decoded glimpses of this canvas
  {
    const bool bulk_ok = ((G & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.glimpse) & 15) == 0);
    if (bulk_ok) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)(T * G) * 4u);
        for (int t = 0; t < T; ++t)
          bulk_g2s(s_gl + (size_t)t * G, a.glimpse + ((size_t)t * B + b) * G, (uint32_t)G * 4u, &bar);
      }
    } else {
      for (int i = threadIdx.x; i < T * G; i += blockDim.x) {
        const int t = i / G, g = i - t * G;
        s_gl[i] = a.glimpse[((size_t)t * B + b) * G + g];
      }
    }
    // inverse-grid tables while the copy is in flight
    for (int i = threadIdx.x; i < T * (W + H); i += blockDim.x) {
      const int t = i / (W + H), j = i - t * (W + H);
      const float* wh = a.where + ((size_t)t * B + b) * 4;
      float a_inv, d_inv, ntx, nty;
      inv_params(wh[0], wh[1], wh[2], wh[3], a_inv, d_inv, ntx, nty);
      if (j < W) s_x[t * W + j] = inv_coord(a_inv, ntx, j, W, w);
      else       s_y[t * H + (j - W)] = inv_coord(d_inv, nty, j - W, H, h);
    }
    if (threadIdx.x < T) s_pres[threadIdx.x] = a.presence[(size_t)threadIdx.x * B + b];
    if (bulk_ok) mbar_wait(&bar, 0);
    __syncthreads();
  }

  // optional visualisation output: presence * sigmoid(glimpse)   (model.py:90)
  if (a.glimpse_viz) {
    for (int i = threadIdx.x; i < T * G; i += blockDim.x) {
      const int t = i / G, g = i - t * G;
      a.glimpse_viz[((size_t)t * B + b) * G + g] = __fmul_rn(s_pres[t], sigmoid_f(s_gl[i]));
    }
  }

  const float mult = a.output_multiplier, sigma = a.output_std;
  const float lp_const = (float)(0.5 * 1.8378770664093453 /*log(2 pi)*/ + log((double)sigma));
  float rec = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    float acc = a.canvas_in ? a.canvas_in[(size_t)b * P + p] : 0.f;
    for (int t = 0; t < T; ++t) {
      const float v = resample_plane(s_gl + (size_t)t * G, h, w, s_x[t * W + c], s_y[t * H + r]);
      acc = __fadd_rn(acc, __fmul_rn(s_pres[t], v));
      if (a.canvas) a.canvas[((size_t)t * B + b) * P + p] = __fmul_rn(acc, mult);
    }
    if (a.do_elbo) {
      const float mu = __fmul_rn(acc, mult);
      const float z = __fdiv_rn(a.img[(size_t)b * P + p] - mu, sigma);
      rec += __fadd_rn(__fmul_rn(__fmul_rn(0.5f, z), z), lp_const);
    }
  }
  if (!a.do_elbo) return;
  rec = block_sum(rec, s_red);

  // step-count posterior, its KL and the per-step weights: tiny float64 island, one thread
  const air_prior& pr = a.prior;
  if (threadIdx.x == 0) {
    modified_geometric_row(a.presence_prob + b, B, T, s_q);
    float kl_n = 0.f;
    for (int k = 0; k <= T; ++k) {
      const double prior_k = pr.steps_prob_is_f64 ? geom_prior_f64(pr.steps_success_prob, k)
                                                  : (double)geom_prior_f32((float)pr.steps_success_prob, k);
      kl_n += tabular_kl_entry(s_q[k], prior_k, 0.0);
      a.num_steps_posterior[(size_t)b * (T + 1) + k] = s_q[k];
    }
    a.kl_num_steps_per_sample[b] = kl_n;
    float n = 0.f;
    for (int t = 0; t < T; ++t) n += s_pres[t];
    a.num_step_per_sample[b] = n;
    int idx = (int)n;
    idx = idx < 0 ? 0 : (idx > T ? T : idx);
    a.num_steps_log_prob[b] = logf(fmaxf(s_q[idx], 1e-32f));
    if (pr.analytic) {   // reverse cumsum of q(n)[1:]   (model.py:157-161)
      float cs = 0.f;
      for (int t = T - 1; t >= 0; --t) {
        cs = (t == T - 1) ? s_q[t + 1] : __fadd_rn(cs, s_q[t + 1]);
        s_w[t] = cs;
      }
    } else {
      for (int t = 0; t < T; ++t) s_w[t] = s_pres[t];
    }
    for (int t = 0; t < T; ++t) a.prior_step_weight[(size_t)t * B + b] = s_w[t];
  }
  __syncthreads();

  // KL(what)   (model.py:174-186)
  float kl_what = 0.f;
  for (int t = 0; t < T; ++t) {
    float v = 0.f;
    const size_t base = ((size_t)t * B + b) * a.na;
    for (int i = threadIdx.x; i < a.na; i += blockDim.x)
      v += normal_kl(a.what_loc[base + i], a.what_scale[base + i], pr.what_loc, pr.what_scale);
    v = block_sum(v, s_red);
    kl_what = (t == 0) ? __fmul_rn(v, s_w[t]) : __fadd_rn(kl_what, __fmul_rn(v, s_w[t]));
  }

  if (threadIdx.x == 0) {
    // KL(where)   (model.py:188-214): (sx, sy) vs scale prior, (tx, ty) vs shift prior
    float kl_where = 0.f;
    for (int t = 0; t < T; ++t) {
      const float* wl = a.where_loc + ((size_t)t * B + b) * 4;
      const float* ws = a.where_scale + ((size_t)t * B + b) * 4;
      const float k_sx = normal_kl(wl[0], ws[0], pr.where_scale_loc, pr.where_scale_scale);
      const float k_sy = normal_kl(wl[2], ws[2], pr.where_scale_loc, pr.where_scale_scale);
      const float k_tx = normal_kl(wl[1], ws[1], pr.where_shift_has_loc ? pr.where_shift_loc : wl[1], pr.where_shift_scale);
      const float k_ty = normal_kl(wl[3], ws[3], pr.where_shift_has_loc ? pr.where_shift_loc : wl[3], pr.where_shift_scale);
      const float s = __fadd_rn(__fadd_rn(k_sx, k_tx), __fadd_rn(k_sy, k_ty));
      const float ws_t = __fmul_rn(s, s_w[t]);
      kl_where = (t == 0) ? ws_t : __fadd_rn(kl_where, ws_t);
    }
    const float kl_n = a.kl_num_steps_per_sample[b];
    a.rec_loss_per_sample[b] = rec;
    a.kl_what_per_sample[b] = kl_what;
    a.kl_where_per_sample[b] = kl_where;
    // Loss.add bookkeeping (ops.py:12-29; model.py:154,186,214,324,332)
    const float prior_ps = __fadd_rn(__fadd_rn(__fmul_rn(kl_n, pr.steps_weight), kl_what), kl_where);
    a.loss_per_sample[b] = __fadd_rn(rec, __fmul_rn(prior_ps, pr.use_prior ? 1.0f : 0.0f));
  }
}

// stand-alone inverse STN (air_stn_paint): out[b] = resampler(glimpse[b], inverse warp(where[b]))
__global__ void __launch_bounds__(256)
stn_paint_kernel(const float* __restrict__ glimpse, const float* __restrict__ where, float* __restrict__ out, int H,
                 int W, int h, int w) {
  extern __shared__ __align__(16) float s_gl[];
  const int b = blockIdx.x;
  const int P = H * W, G = h * w;
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_gl[i] = glimpse[(size_t)b * G + i];
  __syncthreads();
  float a_inv, d_inv, ntx, nty;
  inv_params(where[b * 4 + 0], where[b * 4 + 1], where[b * 4 + 2], where[b * 4 + 3], a_inv, d_inv, ntx, nty);
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int r = p / W, c = p - r * W;
    out[(size_t)b * P + p] = resample_plane(s_gl, h, w, inv_coord(a_inv, ntx, c, W, w), inv_coord(d_inv, nty, r, H, h));
  }
}

// ---------------------------------------------------------------------------------------------------
// batch means of the per-sample terms (model.py:103,151,184,212,247-248,322; ops.py:12-29).  One CTA,
// fixed summation order (deterministic).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
elbo_scalars_kernel(const float* __restrict__ rec, const float* __restrict__ kl_n, const float* __restrict__ kl_what,
                    const float* __restrict__ kl_where, const float* __restrict__ nsteps,
                    const float* __restrict__ logq, const float* __restrict__ baseline, float* __restrict__ scalars,
                    int B, air_prior pr) {
  __shared__ float s_red[32];
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float r = rec[b], lq = logq[b];
    float iw = r;   // REINFORCE importance weight (model.py:337-339)
    if (!pr.analytic) iw = __fadd_rn(r, __fadd_rn(__fadd_rn(__fmul_rn(kl_n[b], pr.steps_weight), kl_what[b]), kl_where[b]));
    s[0] += r;
    s[1] += kl_n[b];
    s[2] += kl_what[b];
    s[3] += kl_where[b];
    s[4] += nsteps[b];
    s[5] += iw * lq;
    s[6] += lq;
    s[7] += baseline ? baseline[b] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = block_sum(s[i], s_red);
  if (threadIdx.x == 0) {
    const float inv = 1.0f / (float)B;
    const float m_rec = s[0] * inv, m_kln = s[1] * inv, m_klw = s[2] * inv, m_klwh = s[3] * inv;
    const float prior_loss = m_kln * pr.steps_weight + m_klw + m_klwh;
    const float loss = m_rec + prior_loss * (pr.use_prior ? 1.0f : 0.0f);
    const float m_iwlq = s[5] * inv, m_lq = s[6] * inv, m_base = s[7] * inv;
    // mean over the [B,B] broadcast of (iw_j - baseline_i) * logq_j  ==  mean(iw*logq) - mean(baseline)*mean(logq)
    const float reinforce = pr.use_reinforce ? (m_iwlq - m_base * m_lq) : 0.f;
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
    for (int i = AIR_S_MEAN_BASELINE + 1; i < AIR_N_SCALARS; ++i) scalars[i] = 0.f;
  }
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
