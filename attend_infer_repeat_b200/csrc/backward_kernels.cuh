// Non-GEMM stages of the backward pass: the hand-derived gradients of the stages in cell_kernels.cuh,
// i.e. what tf.gradients(opt_loss, model_vars) (model.py:355-356) computes for
//   * the reconstruction term through the canvas and the inverse spatial transformer (cell.py:159-164, model.py:319-321;
//     snt.resampler's registered gradient [upstream]: d/d data = scatter of the bilinear weights, d/d warp = data
//     differences, floor() treated as a constant),
//   * the glimpse read (cell.py:135) with respect to the where code,
//   * the what / where reparameterisation and their Normal||Normal KL terms (modules.py:11-63, model.py:174-214),
//   * the step-count posterior: KL(q(n)||prior), the q(n)-weighted KL terms and the REINFORCE term
//     (prior.py:62-90,148-151; model.py:143-161,218-251), float64 island like the forward,
//   * the LSTM gates (snt.LSTM [upstream]),
// plus the centered RMSProp update (tf.train.RMSPropOptimizer(centered=True, momentum=.9), model.py:265,355-360).
// Per-canvas kernels: one CTA (or warp) per canvas, the tile staged in shared memory, warp-shuffle reductions.
#pragma once
#include <cuda_bf16.h>
#include "cell_kernels.cuh"

namespace air {

// bilinear tap along one axis for the backward pass: the raw fractional weight plus validity bits (the forward Tap
// folds validity into zeroed weights, which loses the distinction the derivative needs)
struct __align__(16) BTap {
  float d;       // ceil - coord: weight of the floor tap
  float aux;     // paint: the coordinate itself; read: the pre-scaled grid feature u * S of this output column / row
  int i_f;       // clamped floor index * stride
  int i_c_fl;    // clamped ceil index * stride | flags << 24   (bit 0: floor tap valid, 1: ceil tap valid, 2: inside)
};
__device__ __forceinline__ BTap make_btap(float coord, int n, int stride, float aux) {
  BTap t;
  const bool inside = coord > -1.0f && coord < (float)n;
  const float f = floorf(coord);
  const int fi = (int)f, ci = fi + 1;
  t.d = (f + 1.0f) - coord;
  t.aux = aux;
  const int fl = ((fi >= 0 && fi <= n - 1) ? 1 : 0) | ((ci >= 0 && ci <= n - 1) ? 2 : 0) | (inside ? 4 : 0);
  t.i_f = min(max(fi, 0), n - 1) * stride;
  t.i_c_fl = (min(max(ci, 0), n - 1) * stride) | (fl << 24);
  return t;
}
// value-gradient pieces of one bilinear sample: the four (validity-masked) data values
struct Quad {
  float ff, cc, fc, cf;
};
__device__ __forceinline__ Quad load_quad(const float* __restrict__ D, const BTap& x, const BTap& y) {
  const int xc = x.i_c_fl & 0xffffff, yc = y.i_c_fl & 0xffffff;
  const int xf_ok = (x.i_c_fl >> 24) & 1, xc_ok = (x.i_c_fl >> 25) & 1;
  const int yf_ok = (y.i_c_fl >> 24) & 1, yc_ok = (y.i_c_fl >> 25) & 1;
  Quad q;
  q.ff = (xf_ok & yf_ok) ? D[y.i_f + x.i_f] : 0.f;
  q.cc = (xc_ok & yc_ok) ? D[yc + xc] : 0.f;
  q.fc = (xf_ok & yc_ok) ? D[yc + x.i_f] : 0.f;   // (fx, cy)
  q.cf = (xc_ok & yf_ok) ? D[y.i_f + xc] : 0.f;   // (cx, fy)
  return q;
}
__device__ __forceinline__ bool btap_inside(const BTap& t) { return (t.i_c_fl >> 26) & 1; }

struct BwdArgs {
  // forward tensors
  const float* img;            // [B,P]
  const float* canvas_final;   // [B,P] = output_multiplier * canvas_T (the last slice of the `canvas` output)
  const float* glimpse;        // [T,B,G]
  const float* where;          // [T,B,4]
  const float* where_loc;
  const float* where_scale;
  const float* eps_where;
  const float* what_loc;       // [T,B,na]
  const float* what_scale;
  const float* eps_what;
  const float* presence;       // [T,B]
  const float* presence_prob;  // [T,B]
  const float* posterior;      // [B,T+1]
  const float* num_step;       // [B]
  const float* step_weight;    // [T,B]
  const float* rec_ps;         // [B]
  const float* kl_n_ps;
  const float* kl_what_ps;
  const float* kl_where_ps;
  // gradient tensors
  float* dglimpse;             // [T,B,G]
  float* dwhere_paint;         // [T,B,4]
  const float* dcrop;          // [T,B,G]
  float* dwhere_read;          // [T,B,4]
  const float* dwhat;          // [T,B,na]
  float* dr;                   // [T,B,2na]
  float* dm;                   // [T,B,8]
  float* dlogit;               // [T,B]
  float* dpresence;            // [T,B] non-discrete steps only (cell.py:150-151: presence = presence_prob): d rec / d presence_t
  int discrete;                // cfg.discrete_steps
  int T, B, H, W, h, w, na;
  float output_std, output_multiplier, max_crop, explore_eps;
  float inv_batch;             // 1 / (global batch): every per-sample term enters the loss through a batch mean
  float baseline_mean;         // mean over the global batch of the REINFORCE baseline (0 without one)
  const float* baseline_mean_dev;   // non-null: read the mean from device memory instead (scalars[AIR_S_MEAN_BASELINE])
  double step_W, step_H, step_w, step_h;   // np.linspace steps (host, float64) as in the forward
  air_prior prior;
  double steps_prior[AIR_MAX_STEPS + 1];
  const double* steps_prior_dev;    // non-null: the table is read from device memory (air_prior_table_device)
};

// ---------------------------------------------------------------------------------------------------
// d rec / d glimpse_t and d rec / d where_t through the inverse transformer.  One CTA per canvas.
//   rec = sum_p 0.5 ((x_p - mu_p)/sigma)^2 + const, mu = mult * (sum_t presence_t * inv_t)
//   d rec / d inv_t[p] = dC_p * presence_t,  dC_p = -mult * (x_p - mu_p) / sigma^2 / batch
// snt.resampler's gradient with respect to its data is the scatter of the bilinear weights.  The inverse warp being
// axis-aligned, the weights factor, w(p; j, i) = wy(r, j) * wx(c, i), and being monotone, the canvas columns that touch
// glimpse column i (rows that touch glimpse row j) form ONE contiguous range.  So the scatter is evaluated as a two-pass
// GATHER with no atomics (shared-memory fp32 atomics are CAS loops on sm_100):
//   U[r][i]            = sum_{c in cols(i)} wx(c, i) * dC[r][c]          (rows of the footprint rectangle x glimpse columns)
//   d glimpse_t[j][i]  = presence_t * sum_{r in rows(j)} wy(r, j) * U[r][i]
// d rec / d where_t comes from the data differences along each axis (one pass over the footprint rectangle).
// dynamic smem: T*G floats (glimpses) + T*(W+H) taps + P floats (dC) + H*w floats (U) + 2*T*(w+h) ints (ranges)
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t paint_bwd_smem(int T, int H, int W, int h, int w) {
  return (sizeof(float) * (size_t)T * h * w + 15) / 16 * 16 + sizeof(BTap) * (size_t)T * (W + H) +
         sizeof(float) * ((size_t)H * W + (size_t)H * w) + sizeof(int) * 2 * (size_t)T * (w + h);
}

// weight with which a canvas column / row (tap) contributes to glimpse column / row `idx` (idx pre-multiplied by the tap's
// index stride)
__device__ __forceinline__ float btap_weight(const BTap& t, int idx) {
  const int ic = t.i_c_fl & 0xffffff;
  const int f_ok = (t.i_c_fl >> 24) & 1, c_ok = (t.i_c_fl >> 25) & 1;
  float wgt = 0.f;
  if (f_ok && t.i_f == idx) wgt += t.d;
  if (c_ok && ic == idx) wgt += 1.0f - t.d;
  return wgt;
}

// FH .. Fw > 0: canvas / glimpse shape as compile-time constants (the quoted configuration), 0: from the arguments
template <int T, int FH, int FW, int Fh, int Fw>
__global__ void __launch_bounds__(256) paint_bwd_kernel(BwdArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float4 s_inv[AIR_MAX_STEPS];
  __shared__ float s_pres[AIR_MAX_STEPS];
  __shared__ float s_red[8][4 * T];
  __shared__ float s_redp[8][T];
  __shared__ int s_rect[T][4];   // canvas rectangle inside glimpse t's footprint: c_lo, c_hi, r_lo, r_hi
  const int B = a.B, H = FH ? FH : a.H, W = FW ? FW : a.W, h = Fh ? Fh : a.h, w = Fw ? Fw : a.w;
  const int P = H * W, G = h * w;
  const int b = blockIdx.x;
  float* s_gl = reinterpret_cast<float*>(smem_raw);                                                      // [T][G]
  BTap* s_tx = reinterpret_cast<BTap*>(smem_raw + (sizeof(float) * (size_t)T * G + 15) / 16 * 16);       // [T][W]
  BTap* s_ty = s_tx + (size_t)T * W;                                                                     // [T][H]
  float* s_dC = reinterpret_cast<float*>(s_ty + (size_t)T * H);                                          // [P]
  float* s_U = s_dC + P;                                                                                 // [H][w]
  int* s_clo = reinterpret_cast<int*>(s_U + (size_t)H * w);   // [T][w] first / last canvas column touching glimpse column i
  int* s_chi = s_clo + T * w;
  int* s_rlo = s_chi + T * w;                                 // [T][h] first / last canvas row touching glimpse row j
  int* s_rhi = s_rlo + T * h;
  griddep_launch();
  griddep_wait();
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float* src = a.glimpse + ((size_t)t * B + b) * G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) s_gl[t * G + g] = src[g];
  }
  {
    const float cf = -a.inv_batch * a.output_multiplier / (a.output_std * a.output_std);
    const float* obs = a.img + (size_t)b * P;
    const float* mu = a.canvas_final + (size_t)b * P;
    for (int p = threadIdx.x; p < P; p += blockDim.x) s_dC[p] = cf * (obs[p] - mu[p]);
  }
  for (int i = threadIdx.x; i < T * w; i += blockDim.x) {
    s_clo[i] = W;
    s_chi[i] = -1;
  }
  for (int i = threadIdx.x; i < T * h; i += blockDim.x) {
    s_rlo[i] = H;
    s_rhi[i] = -1;
  }
  if (threadIdx.x < T) {
    const float* wh = a.where + ((size_t)threadIdx.x * B + b) * 4;
    float4 iv;
    inv_params(wh[0], wh[1], wh[2], wh[3], iv.x, iv.y, iv.z, iv.w);
    s_inv[threadIdx.x] = iv;
    s_pres[threadIdx.x] = a.presence[(size_t)threadIdx.x * B + b];
    s_rect[threadIdx.x][0] = W;
    s_rect[threadIdx.x][1] = -1;
    s_rect[threadIdx.x][2] = H;
    s_rect[threadIdx.x][3] = -1;
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float4 iv = s_inv[t];
    for (int j = threadIdx.x; j < W + H; j += blockDim.x) {
      if (j < W) {
        const float xg = inv_coord_s(iv.x, iv.z, j, a.step_W, w);
        const BTap tp = make_btap(xg, w, 1, xg);
        s_tx[t * W + j] = tp;
        if (btap_inside(tp)) {
          atomicMin(&s_rect[t][0], j);
          atomicMax(&s_rect[t][1], j);
          if ((tp.i_c_fl >> 24) & 1) {
            atomicMin(&s_clo[t * w + tp.i_f], j);
            atomicMax(&s_chi[t * w + tp.i_f], j);
          }
          if ((tp.i_c_fl >> 25) & 1) {
            atomicMin(&s_clo[t * w + (tp.i_c_fl & 0xffffff)], j);
            atomicMax(&s_chi[t * w + (tp.i_c_fl & 0xffffff)], j);
          }
        }
      } else {
        const int r = j - W;
        const float yg = inv_coord_s(iv.y, iv.w, r, a.step_H, h);
        const BTap tp = make_btap(yg, h, w, yg);   // indices pre-multiplied by the glimpse row pitch
        s_ty[t * H + r] = tp;
        if (btap_inside(tp)) {
          atomicMin(&s_rect[t][2], r);
          atomicMax(&s_rect[t][3], r);
          if ((tp.i_c_fl >> 24) & 1) {
            atomicMin(&s_rlo[t * h + tp.i_f / w], r);
            atomicMax(&s_rhi[t * h + tp.i_f / w], r);
          }
          if ((tp.i_c_fl >> 25) & 1) {
            atomicMin(&s_rlo[t * h + (tp.i_c_fl & 0xffffff) / w], r);
            atomicMax(&s_rhi[t * h + (tp.i_c_fl & 0xffffff) / w], r);
          }
        }
      }
    }
  }
  __syncthreads();

  const float S_w = ((float)w - 1.0f) * 0.5f, S_h = ((float)h - 1.0f) * 0.5f;
  float acc[T][4];   // per step: sum gx * (xg - S_w), sum gx, sum gy * (yg - S_h), sum gy
  float dp_acc[T];   // per step: sum_g glimpse_t[g] * (sum_p w(p, g) dC[p]) = d rec / d presence_t (non-discrete steps)
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = dp_acc[t] = 0.f;
  const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float pres = s_pres[t];
    const int c_lo = s_rect[t][0], c_hi = s_rect[t][1], r_lo = s_rect[t][2], r_hi = s_rect[t][3];
    float* dgl = a.dglimpse + ((size_t)t * B + b) * G;
    // (a presence of exactly 0 still has a gradient when presence is the continuous probability)
    if ((a.discrete && pres == 0.f) || c_hi < c_lo || r_hi < r_lo) {   // CTA-uniform: this glimpse never reached the canvas
      for (int g = threadIdx.x; g < G; g += blockDim.x) dgl[g] = 0.f;
      continue;
    }
    const float* D = s_gl + t * G;
    // (1) d rec / d where_t: the footprint rectangle, a warp per row, a lane per column
    for (int r = r_lo + warp_; r <= r_hi; r += n_warps) {
      const BTap by = s_ty[t * H + r];
      const float dy = by.d;
      for (int c = c_lo + lane_; c <= c_hi; c += 32) {
        const BTap bx = s_tx[t * W + c];
        const float gv = pres * s_dC[r * W + c];
        const float dx = bx.d;
        const Quad q = load_quad(D, bx, by);
        const float gx = gv * (((1.0f - dy) * q.cc + dy * q.cf) - (dy * q.ff + (1.0f - dy) * q.fc));
        const float gy = gv * ((dx * q.fc + (1.0f - dx) * q.cc) - (dx * q.ff + (1.0f - dx) * q.cf));
        acc[t][0] = fmaf(gx, bx.aux - S_w, acc[t][0]);
        acc[t][1] += gx;
        acc[t][2] = fmaf(gy, by.aux - S_h, acc[t][2]);
        acc[t][3] += gy;
      }
    }
    // (2) U[r][i] = sum_c wx(c, i) dC[r][c] over the rectangle's rows
    const int nr = r_hi - r_lo + 1;
    for (int idx = threadIdx.x; idx < nr * w; idx += blockDim.x) {
      const int rr = idx / w, i = idx - rr * w;
      const int lo = s_clo[t * w + i], hi = s_chi[t * w + i];
      const float* dCr = s_dC + (r_lo + rr) * W;
      float sum = 0.f;
      for (int c = lo; c <= hi; ++c) sum = fmaf(btap_weight(s_tx[t * W + c], i), dCr[c], sum);
      s_U[rr * w + i] = sum;
    }
    __syncthreads();
    // (3) d glimpse_t[j][i] = presence_t * sum_r wy(r, j) U[r][i]
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const int j = g / w, i = g - j * w;
      const int lo = s_rlo[t * h + j], hi = s_rhi[t * h + j];
      float sum = 0.f;
      for (int r = lo; r <= hi; ++r) sum = fmaf(btap_weight(s_ty[t * H + r], j * w), s_U[(r - r_lo) * w + i], sum);
      dgl[g] = pres * sum;
      dp_acc[t] = fmaf(sum, D[g], dp_acc[t]);
    }
    __syncthreads();   // s_U is reused by the next step
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = warp_sum(acc[t][k]);
      if (lane == 0) s_red[wid][t * 4 + k] = v;
    }
  if (a.dpresence) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float v = warp_sum(dp_acc[t]);
      if (lane == 0) s_redp[wid][t] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < T) {
    const int t = threadIdx.x;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi)
      for (int k = 0; k < 4; ++k) s[k] += s_red[wi][t * 4 + k];
    const float* wh = a.where + ((size_t)t * B + b) * 4;
    float* o = a.dwhere_paint + ((size_t)t * B + b) * 4;
    // x_g - S_w = (U S_w - tx S_w) / sx  ->  d x_g / d sx = -(x_g - S_w) / sx,  d x_g / d tx = -S_w / sx
    // A sampled scale can be exactly 0 (loc + scale * eps cancels in fp32: ~2e-8 per draw, i.e. once every few hundred
    // steps at B = 4096) or too small for 1 / s to be finite.  The glimpse then covers no canvas pixel, the sums are
    // exactly 0 and the gradient through the painted canvas is 0 -- not the 0 / 0 the formula gives (which turned every
    // parameter upstream of `where` into NaN; the reference's graph divides by the same determinant).
    o[0] = s[0] == 0.f ? 0.f : -s[0] / wh[0];
    o[1] = s[1] == 0.f ? 0.f : -S_w * s[1] / wh[0];
    o[2] = s[2] == 0.f ? 0.f : -s[2] / wh[2];
    o[3] = s[3] == 0.f ? 0.f : -S_h * s[3] / wh[2];
    if (a.dpresence) {
      float dp = 0.f;
      for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) dp += s_redp[wi][t];
      a.dpresence[(size_t)t * B + b] = dp;
    }
  }
}

template <int T>
inline cudaError_t launch_paint_bwd_t(const BwdArgs& a, cudaStream_t st) {
  const size_t smem = paint_bwd_smem(a.T, a.H, a.W, a.h, a.w);
  static const bool no_fixed = getenv("AIR_BWD_NO_FIXED") != nullptr;
  if constexpr (T == 3) {
    if (!no_fixed && a.H == 50 && a.W == 50 && a.h == 20 && a.w == 20) {
      cudaError_t e = ensure_dynamic_smem(paint_bwd_kernel<T, 50, 50, 20, 20>, smem);
      if (e != cudaSuccess) return e;
      return launch_k(paint_bwd_kernel<T, 50, 50, 20, 20>, dim3(a.B), dim3(256), smem, st, a);
    }
  }
  cudaError_t e = ensure_dynamic_smem(paint_bwd_kernel<T, 0, 0, 0, 0>, smem);
  if (e != cudaSuccess) return e;
  return launch_k(paint_bwd_kernel<T, 0, 0, 0, 0>, dim3(a.B), dim3(256), smem, st, a);
}
inline cudaError_t launch_paint_bwd(const BwdArgs& a, cudaStream_t st) {
  switch (a.T) {
    case 1: return launch_paint_bwd_t<1>(a, st);
    case 2: return launch_paint_bwd_t<2>(a, st);
    case 3: return launch_paint_bwd_t<3>(a, st);
    case 4: return launch_paint_bwd_t<4>(a, st);
    case 5: return launch_paint_bwd_t<5>(a, st);
    case 6: return launch_paint_bwd_t<6>(a, st);
    case 7: return launch_paint_bwd_t<7>(a, st);
    case 8: return launch_paint_bwd_t<8>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------------------------------
// d L / d where_t through the glimpse read: crop_t[g] = resampler(img, x = sx u S_W + tx S_W + S_W, y likewise).
// One CTA per canvas; dynamic smem: H*W floats (image) + T*(w+h) taps.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t read_bwd_smem(int T, int H, int W, int h, int w) {
  return (sizeof(float) * (size_t)H * W + 15) / 16 * 16 + sizeof(BTap) * (size_t)T * (w + h);
}
template <int FT, int FH, int FW, int Fh, int Fw>
__global__ void __launch_bounds__(256) read_bwd_kernel(BwdArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float s_red[8][4];
  const int T = FT ? FT : a.T, B = a.B, H = FH ? FH : a.H, W = FW ? FW : a.W, h = Fh ? Fh : a.h, w = Fw ? Fw : a.w;
  const int P = H * W, G = h * w;
  const int b = blockIdx.x;
  float* s_img = reinterpret_cast<float*>(smem_raw);
  BTap* s_tx = reinterpret_cast<BTap*>(smem_raw + (sizeof(float) * (size_t)P + 15) / 16 * 16);   // [T][w]
  BTap* s_ty = s_tx + (size_t)T * w;                                                             // [T][h]
  griddep_launch();
  griddep_wait();
  for (int i = threadIdx.x; i < P; i += blockDim.x) s_img[i] = a.img[(size_t)b * P + i];
  const float S_W = ((float)W - 1.0f) * 0.5f, S_H = ((float)H - 1.0f) * 0.5f;
  for (int i = threadIdx.x; i < T * (w + h); i += blockDim.x) {
    const int t = i / (w + h), j = i - t * (w + h);
    const float* wh = a.where + ((size_t)t * B + b) * 4;
    if (j < w) {
      const float uS = __fmul_rn((float)(-1.0 + (double)j * a.step_w), S_W);
      s_tx[t * w + j] = make_btap(fwd_coord_s(wh[0], wh[1], j, a.step_w, W), W, 1, uS);
    } else {
      const float vS = __fmul_rn((float)(-1.0 + (double)(j - w) * a.step_h), S_H);
      s_ty[t * h + (j - w)] = make_btap(fwd_coord_s(wh[2], wh[3], j - w, a.step_h, H), H, W, vS);
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = 0; t < T; ++t) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* dc = a.dcrop + ((size_t)t * B + b) * G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const int r = g / w, c = g - r * w;
      const BTap bx = s_tx[t * w + c], by = s_ty[t * h + r];
      if (!(btap_inside(bx) && btap_inside(by))) continue;
      const float gv = dc[g];
      const float dx = bx.d, dy = by.d;
      const Quad q = load_quad(s_img, bx, by);
      const float gx = gv * (((1.0f - dy) * q.cc + dy * q.cf) - (dy * q.ff + (1.0f - dy) * q.fc));
      const float gy = gv * ((dx * q.fc + (1.0f - dx) * q.cc) - (dx * q.ff + (1.0f - dx) * q.cf));
      acc[0] = fmaf(gx, bx.aux, acc[0]);
      acc[1] += gx;
      acc[2] = fmaf(gy, by.aux, acc[2]);
      acc[3] += gy;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) s_red[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      float s = 0.f;
      for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) s += s_red[wi][threadIdx.x];
      const int k = threadIdx.x;
      // d x / d sx = u S_W, d x / d tx = S_W, d y / d sy = v S_H, d y / d ty = S_H
      a.dwhere_read[((size_t)t * B + b) * 4 + k] = (k == 1) ? S_W * s : (k == 3) ? S_H * s : s;
    }
    __syncthreads();
  }
}
inline cudaError_t launch_read_bwd(const BwdArgs& a, cudaStream_t st) {
  const size_t smem = read_bwd_smem(a.T, a.H, a.W, a.h, a.w);
  static const bool no_fixed = getenv("AIR_BWD_NO_FIXED") != nullptr;
  if (!no_fixed && a.T == 3 && a.H == 50 && a.W == 50 && a.h == 20 && a.w == 20) {
    cudaError_t e = ensure_dynamic_smem(read_bwd_kernel<3, 50, 50, 20, 20>, smem);
    if (e != cudaSuccess) return e;
    return launch_k(read_bwd_kernel<3, 50, 50, 20, 20>, dim3(a.B), dim3(256), smem, st, a);
  }
  cudaError_t e = ensure_dynamic_smem(read_bwd_kernel<0, 0, 0, 0, 0>, smem);
  if (e != cudaSuccess) return e;
  return launch_k(read_bwd_kernel<0, 0, 0, 0, 0>, dim3(a.B), dim3(256), smem, st, a);
}

// d Normal||Normal KL / d (mu_a, s_a)   [upstream _kl_normal_normal]
__device__ __forceinline__ void normal_kl_grad(float mu_a, float s_a, float mu_b, float s_b, float& d_mu, float& d_s) {
  const float sb2 = s_b * s_b;
  d_mu = (mu_a - mu_b) / sb2;
  d_s = s_a / sb2 - 1.0f / s_a;
}

// ---------------------------------------------------------------------------------------------------
// what head backward (modules.py:11-24, cell.py:154-156, model.py:174-186): what = loc + softplus(raw + offset) * eps
//   d loc = d what + c * w_t * dKL/d loc;  d scale = d what * eps + c * w_t * dKL/d scale;  d raw = d scale * sigmoid(.)
// with sigmoid(x) = 1 - exp(-softplus(x)), c = prior_weight / batch, w_t = the q(n) step weight.  -> dr [T*B, 2 na]
// ---------------------------------------------------------------------------------------------------
__global__ void what_bwd_kernel(BwdArgs a) {
  griddep_launch();
  griddep_wait();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t rows = (size_t)a.T * a.B;
  if (idx >= rows * (size_t)a.na) return;
  const size_t row = idx / a.na;
  const int j = (int)(idx % a.na);
  const float c = a.inv_batch * (a.prior.use_prior ? 1.0f : 0.0f) * a.step_weight[row];
  const float loc = a.what_loc[idx], sc = a.what_scale[idx], g = a.dwhat[idx];
  float k_mu, k_s;
  normal_kl_grad(loc, sc, a.prior.what_loc, a.prior.what_scale, k_mu, k_s);
  const float d_loc = fmaf(c, k_mu, g);
  const float d_sc = fmaf(c, k_s, g * a.eps_what[idx]);
  a.dr[row * 2 * a.na + j] = d_loc;
  a.dr[row * 2 * a.na + a.na + j] = d_sc * (1.0f - expf(-sc));
}

// ---------------------------------------------------------------------------------------------------
// where head + step-count backward.  One warp per canvas.
//   where: d where_t = d(paint) + d(read); where = loc + scale * eps; loc = (max_crop sig, tanh, max_crop sig, tanh)(m[0:4]);
//          scale = softplus(m[4:8] + bias); KL(where) terms weighted by w_t                       -> dm [T*B, 8]
//   steps: L depends on p_t = presence_prob through q(n) (prior.py:62-68): KL(q||prior), the weights w_t = sum_{k>t} q_k
//          (analytic mode, model.py:157-161) and REINFORCE's log q(n_b) (clip straight-through, ops.py:67-76),
//          float64 like the forward island; p = eps/2 + (1 - eps) sigmoid(logit + bias)            -> dlogit [T*B]
// ---------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(128) latent_bwd_kernel(BwdArgs a) {
  griddep_launch();
  griddep_wait();
  const int B = a.B;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const air_prior& pr = a.prior;
  const float pw = pr.use_prior ? 1.0f : 0.0f;
  const float coef = a.inv_batch * pw;
  // KL(what) per step, recomputed (model.py:174-186)
  float klw[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float v = 0.f;
    const size_t base = ((size_t)t * B + b) * a.na;
    for (int i = lane; i < a.na; i += 32)
      v += normal_kl(a.what_loc[base + i], a.what_scale[base + i], pr.what_loc, pr.what_scale);
    klw[t] = warp_sum(v);
  }
  // where head: lane t owns step t
  float klwh_l = 0.f;
  if (lane < T) {
    const size_t row = (size_t)lane * B + b;
    const float sw = a.step_weight[row];
    float dmv[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float loc = a.where_loc[row * 4 + k], sc = a.where_scale[row * 4 + k];
      const float g = a.dwhere_paint[row * 4 + k] + a.dwhere_read[row * 4 + k];
      const bool shift = k & 1;
      const float mu_b = shift ? (pr.where_shift_has_loc ? pr.where_shift_loc : loc) : pr.where_scale_loc;
      const float s_b = shift ? pr.where_shift_scale : pr.where_scale_scale;
      float k_mu, k_s;
      normal_kl_grad(loc, sc, mu_b, s_b, k_mu, k_s);
      klwh_l += normal_kl(loc, sc, mu_b, s_b);
      const float d_loc = fmaf(coef * sw, k_mu, g);
      const float d_sc = fmaf(coef * sw, k_s, g * a.eps_where[row * 4 + k]);
      dmv[k] = d_loc * (shift ? (1.0f - loc * loc) : loc * (1.0f - loc / a.max_crop));
      dmv[4 + k] = d_sc * (1.0f - expf(-sc));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) a.dm[row * 8 + k] = dmv[k];
  }
  float klwh[T];
#pragma unroll
  for (int t = 0; t < T; ++t) klwh[t] = __shfl_sync(0xffffffffu, klwh_l, t);
  if (lane != 0) return;

  // ---- step-count posterior backward (float64) ----
  double p[T], q[T + 1], dq[T + 1];
#pragma unroll
  for (int t = 0; t < T; ++t) p[t] = (double)a.presence_prob[(size_t)t * B + b];
#pragma unroll
  for (int k = 0; k <= T; ++k) q[k] = (double)a.posterior[(size_t)b * (T + 1) + k];
  // REINFORCE coefficient on log q(n_b): stop_gradient(importance weight - baseline) / batch   (model.py:242-248)
  double cq = 0.0;
  int n_idx = (int)a.num_step[b];
  n_idx = n_idx < 0 ? 0 : (n_idx > T ? T : n_idx);
  if (pr.use_reinforce) {
    float iw = a.rec_ps[b];
    if (!pr.analytic)
      iw = __fadd_rn(iw, __fadd_rn(__fadd_rn(__fmul_rn(a.kl_n_ps[b], pr.steps_weight), a.kl_what_ps[b]), a.kl_where_ps[b]));
    const double nv_scale = pr.nvil_scale != 0.f ? (double)pr.nvil_scale : 1.0;
    const double nv_shift = pr.nvil_scale != 0.f ? (double)pr.nvil_shift : 0.0;
    const double bmean = a.baseline_mean_dev ? (double)__ldg(a.baseline_mean_dev) : (double)a.baseline_mean;
    cq = (double)a.inv_batch * ((double)iw - bmean - nv_shift) * nv_scale;
  }
  double dsw_run = 0.0;   // sum_{t < k} dL/dw_t
#pragma unroll
  for (int k = 0; k <= T; ++k) {
    double g = 0.0;
    if (q[k] > 0.0) g = (double)coef * (double)pr.steps_weight * (log(q[k] / (a.steps_prior_dev ? a.steps_prior_dev[k] : a.steps_prior[k])) + 1.0);
    if (k >= 1 && pr.analytic) {
      dsw_run += (double)coef * ((double)klw[k - 1] + (double)klwh[k - 1]);
      g += dsw_run;
    }
    if (k == n_idx) g += cq / fmax(q[k], 1e-32);
    dq[k] = g;
  }
  // q = pi / sum(pi)
  double pi[T + 1], cum[T + 1];
  double run = 1.0, S = 0.0;
#pragma unroll
  for (int k = 0; k <= T; ++k) {
    cum[k] = run;                                   // prod_{j<k} p_j
    pi[k] = (k < T) ? (1.0 - p[k]) * run : run;
    if (k < T) run *= p[k];
    S += pi[k];
  }
  double dot = 0.0;
#pragma unroll
  for (int k = 0; k <= T; ++k) dot += dq[k] * (pi[k] / S);
  double dpi[T + 1];
#pragma unroll
  for (int k = 0; k <= T; ++k) dpi[k] = (dq[k] - dot) / S;
  const float ee = a.explore_eps;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    // d pi_j / d p_j = -cum_j;  d pi_k / d p_j (k > j) = (1 - p_k [k < T]) * prod_{i < k, i != j} p_i
    double g = -dpi[j] * cum[j];
#pragma unroll
    for (int k = j + 1; k <= T; ++k) {
      double prod = 1.0;
#pragma unroll
      for (int i = 0; i < k; ++i)
        if (i != j) prod *= p[i];
      g += dpi[k] * ((k < T) ? (1.0 - p[k]) * prod : prod);
    }
    if (!a.discrete) {
      // cell.py:150-151: presence_t IS presence_prob_t -- the painted canvas (d rec / d presence_t from paint_bwd) and,
      // with sampled-style step weights (analytic = False, model.py:163: the weights are the presence), both KL terms
      // depend on p_t directly
      if (a.dpresence) g += (double)a.dpresence[(size_t)j * B + b];
      if (!pr.analytic) g += (double)coef * ((double)klw[j] + (double)klwh[j]);
    }
    double s = p[j], scale = 1.0;   // p = eps/2 + (1 - eps) * sigmoid(.)   (cell.py:140-141)
    if (ee >= 0.f) {
      scale = 1.0 - (double)ee;
      s = (p[j] - 0.5 * (double)ee) / scale;
    }
    a.dlogit[(size_t)j * B + b] = (float)(g * scale * s * (1.0 - s));
  }
}
inline cudaError_t launch_latent_bwd(const BwdArgs& a, cudaStream_t st) {
  const dim3 grid((a.B + 3) / 4), block(128);
  switch (a.T) {
    case 1: return launch_k(latent_bwd_kernel<1>, grid, block, 0, st, a);
    case 2: return launch_k(latent_bwd_kernel<2>, grid, block, 0, st, a);
    case 3: return launch_k(latent_bwd_kernel<3>, grid, block, 0, st, a);
    case 4: return launch_k(latent_bwd_kernel<4>, grid, block, 0, st, a);
    case 5: return launch_k(latent_bwd_kernel<5>, grid, block, 0, st, a);
    case 6: return launch_k(latent_bwd_kernel<6>, grid, block, 0, st, a);
    case 7: return launch_k(latent_bwd_kernel<7>, grid, block, 0, st, a);
    case 8: return launch_k(latent_bwd_kernel<8>, grid, block, 0, st, a);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------------------------------
// LSTM gate backward for one step (snt.LSTM [upstream]; gates order i, j, f, o; forget bias added inside the sigmoid).
//   gates [B,4nh] pre-activation of this step, c_prev / c_new [B,nh], dh_heads [B,nh] (from the where / steps heads),
//   dh_rec [B,nh] (from step t+1 through W_h, null at the last step), dc [B,nh] in: d L / d c_t from step t+1, out: for t-1
//   -> dgates [B,4nh]
// ---------------------------------------------------------------------------------------------------
// U = units per thread (2 when nh is even: 8-byte loads / stores of the fp32 rows, 4-byte stores into the bf16 planes)
template <int U>
__global__ void __launch_bounds__(256)
lstm_bwd_pointwise_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_new,
                          const float* __restrict__ dh_heads, const float* __restrict__ dh_rec, float* __restrict__ dc,
                          float* __restrict__ dgates, int B, int nh, float forget_bias, int first, float* __restrict__ dgx,
                          __half* __restrict__ hl_dst, size_t hl_plane, int hl_ld, int* range_flag) {
  griddep_launch();
  griddep_wait();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int upr = nh / U;                     // threads per canvas row
  if (tid >= (size_t)B * upr) return;
  const size_t b = tid / upr;
  const int u = (int)(tid % upr) * U;
  const size_t idx = b * nh + u;
  const float* g = gates + b * 4 * (size_t)nh + u;
  float gi[U], gj[U], gf[U], go[U], cp[U], cn[U], dhh[U], dhr[U], dcc[U];
  auto ld = [&](const float* p, float* out) {
    if (U == 2) {
      const float2 v = *reinterpret_cast<const float2*>(p);
      out[0] = v.x;
      out[U - 1] = v.y;
    } else {
      out[0] = p[0];
    }
  };
  auto st = [&](float* p, const float* v) {
    if (U == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[U - 1]);
    else p[0] = v[0];
  };
  ld(g, gi); ld(g + nh, gj); ld(g + 2 * nh, gf); ld(g + 3 * nh, go);
  ld(c_prev + idx, cp); ld(c_new + idx, cn); ld(dh_heads + idx, dhh);
#pragma unroll
  for (int q = 0; q < U; ++q) dhr[q] = dcc[q] = 0.f;
  if (dh_rec) ld(dh_rec + idx, dhr);
  if (!first) ld(dc + idx, dcc);
  float d_i[U], d_j[U], d_f[U], d_o[U], dc_out[U];
#pragma unroll
  for (int q = 0; q < U; ++q) {
    const float si = sigmoid_f(gi[q]), tj = tanhf(gj[q]), sf = sigmoid_f(gf[q] + forget_bias), so = sigmoid_f(go[q]);
    const float tc = tanhf(cn[q]);
    const float dh = dhh[q] + dhr[q];
    const float dcv = dcc[q] + dh * so * (1.0f - tc * tc);
    d_i[q] = dcv * tj * si * (1.0f - si);
    d_j[q] = dcv * si * (1.0f - tj * tj);
    d_f[q] = dcv * cp[q] * sf * (1.0f - sf);
    d_o[q] = dh * tc * so * (1.0f - so);
    dc_out[q] = dcv * sf;
  }
  float* dg = dgates + b * 4 * (size_t)nh + u;
  st(dg, d_i); st(dg + nh, d_j); st(dg + 2 * nh, d_f); st(dg + 3 * nh, d_o);
  st(dc + idx, dc_out);
  if (hl_dst) {   // the row-major bf16 hi/lo planes of dgates_t: the A operand of d h_{t-1} = dgates_t @ W_h^T, which follows
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(hl_dst) + b * (size_t)hl_ld + u;
    const float* v[4] = {d_i, d_j, d_f, d_o};
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat16 hi[U], lo[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        hi[q] = __float2bfloat16_rn(v[k][q]);
        lo[q] = __float2bfloat16_rn(v[k][q] - __bfloat162float(hi[q]));
        bad |= !(fabsf(v[k][q]) <= 3.0e38f);
      }
      __nv_bfloat16* dk = d + (size_t)k * nh;
      if (U == 2) {
        *reinterpret_cast<__nv_bfloat162*>(dk) = __halves2bfloat162(hi[0], hi[U - 1]);
        *reinterpret_cast<__nv_bfloat162*>(dk + hl_plane) = __halves2bfloat162(lo[0], lo[U - 1]);
      } else {
        dk[0] = hi[0];
        dk[hl_plane] = lo[0];
      }
    }
    if (bad && range_flag) atomicOr(range_flag, 1);
  }
  if (dgx) {   // the LSTM's input half sees the same encoder output at every step: d gx = sum_t dgates_t, accumulated here
    float* sx = dgx + b * 4 * (size_t)nh + u;
    const float* v[4] = {d_i, d_j, d_f, d_o};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float acc[U];
#pragma unroll
      for (int q = 0; q < U; ++q) acc[q] = 0.f;
      if (!first) ld(sx + (size_t)k * nh, acc);
#pragma unroll
      for (int q = 0; q < U; ++q) acc[q] += v[k][q];
      st(sx + (size_t)k * nh, acc);
    }
  }
}

// grad[i] += l2 * w[i]   (l2_weight * tf.nn.l2_loss(w) on the 2-D variables, model.py:345-350)
__global__ void l2_grad_kernel(const float* __restrict__ w, float* __restrict__ grad, float l2, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad[i] = fmaf(l2, w[i], grad[i]);
}

// baseline_loss = .5 * mean((stop_gradient(iw) - baseline)^2) with iw [B] and baseline [B,1] broadcasting to [B,B]
// (model.py:253-259): d / d baseline_i = -(mean_j iw_j - baseline_i) / B
// BaselineMLP input rows (modules.py:131-141): x[b] = concat[img[b] (P), what[:, b] (T * na), where[:, b] (T * 4),
// presence[:, b] (T), h[b] (nh), c[b] (nh)] -- the time-major cell outputs transposed to batch-major exactly as
// tf.transpose(t, (1, 0, 2)) + reshape does.  One pass: fp32 rows and (tensor-core engine) the fp16 hi/lo operand planes.
__global__ void __launch_bounds__(256)
baseline_gather_kernel(const float* __restrict__ img, const float* __restrict__ what, const float* __restrict__ where,
                       const float* __restrict__ presence, const float* __restrict__ hfin, const float* __restrict__ cfin,
                       float* __restrict__ x, __half* __restrict__ hl, size_t plane, int ld_hl, int B, int T, int P, int na,
                       int nh, int n_in) {
  griddep_launch();
  griddep_wait();
  // one CTA per canvas, two neighbouring columns per thread (32-bit index arithmetic, 4-byte stores into the planes)
  const int b = blockIdx.x;
  auto fetch = [&](int k) -> float {
    if (k < P) return img[(size_t)b * P + k];
    if ((k -= P) < T * na) return what[((size_t)(k / na) * B + b) * na + k % na];
    if ((k -= T * na) < T * 4) return where[((size_t)(k >> 2) * B + b) * 4 + (k & 3)];
    if ((k -= T * 4) < T) return presence[(size_t)k * B + b];
    if ((k -= T) < nh) return hfin[(size_t)b * nh + k];
    return cfin[(size_t)b * nh + (k - nh)];
  };
  float* xr = x + (size_t)b * n_in;
  __half* hr = hl ? hl + (size_t)b * ld_hl : nullptr;
  for (int k = 2 * threadIdx.x; k < n_in; k += 2 * blockDim.x) {
    const bool two = k + 1 < n_in;
    const float v0 = fetch(k), v1 = two ? fetch(k + 1) : 0.f;
    xr[k] = v0;
    if (two) xr[k + 1] = v1;
    if (hr) {
      __half h0, l0, h1, l1;
      split_f16(v0, h0, l0);
      split_f16(v1, h1, l1);
      if (two) {
        *reinterpret_cast<__half2*>(hr + k) = __halves2half2(h0, h1);
        *reinterpret_cast<__half2*>(hr + plane + k) = __halves2half2(l0, l1);
      } else {
        hr[k] = h0;
        hr[plane + k] = l0;
      }
    }
  }
}

__global__ void baseline_grad_kernel(const float* __restrict__ baseline, float target_mean, float inv_batch,
                                     float* __restrict__ d_baseline, int B, const float* __restrict__ target_mean_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float tm = target_mean_dev ? __ldg(target_mean_dev) : target_mean;
  if (i < B) d_baseline[i] = -inv_batch * (tm - baseline[i]);
}

// ---------------------------------------------------------------------------------------------------
// tf.train.RMSPropOptimizer(lr, decay, momentum, epsilon, centered=True) [upstream ApplyCenteredRMSProp]:
//   mg <- mg + (1 - rho)(g - mg);  ms <- ms + (1 - rho)(g^2 - ms);  mom <- mu mom + lr g / sqrt(ms - mg^2 + eps);
//   theta <- theta - mom.     (ms starts at 1, mg and mom at 0; epsilon inside the square root)
// One pass over the flat parameter buffer; grad_scale folds the 1 / world of a summed all-reduce.
// ---------------------------------------------------------------------------------------------------
__global__ void rmsprop_centered_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ mg,
                                        float* __restrict__ ms, float* __restrict__ mom, size_t n, float lr, float rho,
                                        float mu, float eps, float grad_scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = grad[i] * grad_scale;
  const float mgi = mg[i] + (1.0f - rho) * (g - mg[i]);
  const float msi = ms[i] + (1.0f - rho) * (g * g - ms[i]);
  // ms - mg^2 is a variance estimate, but in fp32 it rounds below zero once the gradient of an element stops changing
  // (ms -> g^2, mg -> g; reproduced on the host after ~140 steps of a gradient with 1e-4 relative jitter): the literal
  // formula then takes the square root of a negative number and the parameter is NaN for good.  Clamping the variance at 0 changes nothing wherever TF's result is finite
  // and the estimate non-negative.
  const float mo = mu * mom[i] + lr * g / sqrtf(fmaxf(msi - mgi * mgi, 0.0f) + eps);
  mg[i] = mgi;
  ms[i] = msi;
  mom[i] = mo;
  theta[i] -= mo;
}

}  // namespace air
