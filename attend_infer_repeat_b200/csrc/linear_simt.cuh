// fp32 CUDA-core dense layer: out[M,N] = act(A[M,K] @ W[K,N] + bias[N] (+ addend[M,N])).
// This is the AIR_PREC_FP32 engine: the arithmetic is fp32 FMA with an fp32 accumulator per output, so the
// result differs from an fp32 reference only by summation order.  It is also the general-shape engine
// (any M, N, K, any row alignment) that the tensor-core engine falls back to for odd layer widths.
//
// Replaces: snt.Linear + transfer inside neural.Affine / neural.MLP (neural.py:42-102), reached from
// Encoder / Decoder / StochasticTransformParam / StepsPredictor / ParametrisedGaussian (modules.py:11-122).
#pragma once
#include "common.cuh"

namespace air {

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
linear_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wt, int ldw,
                   const float* __restrict__ bias, const float* __restrict__ addend, int ldadd,
                   float* __restrict__ C, int ldc, int M, int N, int K, int act) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int APAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Ws[2][BK][BN];

  griddep_launch();
  griddep_wait();
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool w_vec = ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(Wt) & 15) == 0);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  constexpr int A_VECS = BM * BK / 4;            // float4 per A tile
  constexpr int A_ITERS = (A_VECS + NT - 1) / NT;
  constexpr int W_VECS = BK * BN / 4;
  constexpr int W_ITERS = (W_VECS + NT - 1) / NT;
  float4 a_reg[A_ITERS], w_reg[W_ITERS];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int it = 0; it < A_ITERS; ++it) {
      const int v = tid + it * NT;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < A_VECS) {
        const int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
        const int gm = m0 + row, gk = k0 + kq;
        if (gm < M) {
          const float* p = A + (size_t)gm * lda + gk;
          if (a_vec && gk + 3 < K) {
            r = *reinterpret_cast<const float4*>(p);
          } else {
            if (gk + 0 < K) r.x = p[0];
            if (gk + 1 < K) r.y = p[1];
            if (gk + 2 < K) r.z = p[2];
            if (gk + 3 < K) r.w = p[3];
          }
        }
      }
      a_reg[it] = r;
    }
#pragma unroll
    for (int it = 0; it < W_ITERS; ++it) {
      const int v = tid + it * NT;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < W_VECS) {
        const int kr = v / (BN / 4), nq = (v % (BN / 4)) * 4;
        const int gk = k0 + kr, gn = n0 + nq;
        if (gk < K) {
          const float* p = Wt + (size_t)gk * ldw + gn;
          if (w_vec && gn + 3 < N) {
            r = *reinterpret_cast<const float4*>(p);
          } else {
            if (gn + 0 < N) r.x = p[0];
            if (gn + 1 < N) r.y = p[1];
            if (gn + 2 < N) r.z = p[2];
            if (gn + 3 < N) r.w = p[3];
          }
        }
      }
      w_reg[it] = r;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int it = 0; it < A_ITERS; ++it) {
      const int v = tid + it * NT;
      if (v < A_VECS) {
        const int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
        As[buf][kq + 0][row] = a_reg[it].x;
        As[buf][kq + 1][row] = a_reg[it].y;
        As[buf][kq + 2][row] = a_reg[it].z;
        As[buf][kq + 3][row] = a_reg[it].w;
      }
    }
#pragma unroll
    for (int it = 0; it < W_ITERS; ++it) {
      const int v = tid + it * NT;
      if (v < W_VECS) {
        const int kr = v / (BN / 4), nq = (v % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Ws[buf][kr][nq]) = w_reg[it];
      }
    }
  };

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tiles((kb + 1) * BK);   // global loads in flight during the FMA block
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * TN + j]);
        w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias (+ addend) + transfer
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (addend) v += addend[(size_t)gm * ldadd + gn];
      C[(size_t)gm * ldc + gn] = apply_act(v, act);
    }
  }
}

inline cudaError_t launch_linear_simt(const float* A, int lda, const float* Wt, int ldw, const float* bias,
                                      const float* addend, int ldadd, float* C, int ldc, int M, int N, int K,
                                      int act, cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (N > 32 && M > 32 && (long long)((N + 63) / 64) * ((M + 127) / 128) < 148) {
    // too few 128-row tiles to fill the machine (the BaselineMLP's [4096, 256] x [256, 128] layer: 64 CTAs): 32-row tiles.
    // The order of the k accumulation per output does not depend on the tile, so the results are the same bit for bit.
    constexpr int BM = 32, BN = 64, BK = 16, TM = 4, TN = 4;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    return launch_k(linear_simt_kernel<BM, BN, BK, TM, TN>, grid, dim3((BM / TM) * (BN / TN)), 0, st, A, lda, Wt, ldw,
                    bias, addend, ldadd, C, ldc, M, N, K, act);
  } else if (N > 32) {
    constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    return launch_k(linear_simt_kernel<BM, BN, BK, TM, TN>, grid, dim3((BM / TM) * (BN / TN)), 0, st, A, lda, Wt, ldw,
                    bias, addend, ldadd, C, ldc, M, N, K, act);
  } else {
    // skinny heads (N = 8 where-head, N = 1 presence logit): tall tiles, little wasted width
    constexpr int BM = 128, BN = 16, BK = 16, TM = 4, TN = 4;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    return launch_k(linear_simt_kernel<BM, BN, BK, TM, TN>, grid, dim3((BM / TM) * (BN / TN)), 0, st, A, lda, Wt, ldw,
                    bias, addend, ldadd, C, ldc, M, N, K, act);
  }
}

}  // namespace air
