// fp32 CUDA-core GEMMs of the backward pass: the two products every dense layer
//   y = act(x @ W + b)            (neural.py:42-60)
// needs under tf.gradients:
//   dX[M,K] = dY[M,N] @ W[K,N]^T   (* elu'(x) when x is itself the output of an ELU layer)      -> "NT"
//   dW[K,N] += X[M,K]^T @ dY[M,N]  (contraction over the M rows of the batch; split over gridDim.z)  -> "TN"
// plus the column sums db[N] = sum_m dY[m,N].  Same register-tiled FMA scheme as linear_simt.cuh (128 x 64 tile,
// BK = 16, 8 x 4 outputs per thread, double-buffered shared-memory tiles); the operand that is stored contraction-major
// is scattered into the k-major tile on the way in.  fp32 FMA with an fp32 accumulator: results differ from an fp32
// reference only by summation order (dW additionally by the order of the split-K atomics).
#pragma once
#include "common.cuh"

namespace air {

// C[M,N] (=|+=) op(A) @ op(B)
//   TA = false: A is [M,K] row-major (lda);   TA = true: A is stored [K,M] row-major (lda)
//   TB = false: B is [K,N] row-major (ldb);   TB = true: B is stored [N,K] row-major (ldb)
//   elu_y != null: C = acc * (Y > 0 ? 1 : Y + 1), Y[M,N] (ldy) = the forward ELU output that C is the gradient of
//   accumulate: C += acc (plain store/add when gridDim.z == 1, atomicAdd when the contraction is split over gridDim.z)
template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, float* __restrict__ C,
                 int ldc, int M, int N, int K, int k_chunk, int accumulate, const float* __restrict__ elu_y, int ldy) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  griddep_launch();
  griddep_wait();
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  if (k_begin >= k_end) return;

  const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(Bm) & 15) == 0);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  constexpr int A_VECS = BM * BK / 4, A_ITERS = (A_VECS + NT - 1) / NT;
  constexpr int B_VECS = BN * BK / 4, B_ITERS = (B_VECS + NT - 1) / NT;
  float4 a_reg[A_ITERS], b_reg[B_ITERS];

  // a [rows, cols] tile of a row-major matrix, 4 consecutive columns per thread, zero outside (r_lim, c_lim)
  auto ld4 = [](const float* base, int ld, int r, int c, int r_lim, int c_lim, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < r_lim) {
      const float* p = base + (size_t)r * ld + c;
      if (vec && c + 3 < c_lim) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        if (c + 0 < c_lim) v.x = p[0];
        if (c + 1 < c_lim) v.y = p[1];
        if (c + 2 < c_lim) v.z = p[2];
        if (c + 3 < c_lim) v.w = p[3];
      }
    }
    return v;
  };
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int it = 0; it < A_ITERS; ++it) {
      const int v = tid + it * NT;
      if (v < A_VECS) {
        if (TA) {   // stored [K,M]: vector along M
          const int kr = v / (BM / 4), mq = (v % (BM / 4)) * 4;
          a_reg[it] = ld4(A, lda, k0 + kr, m0 + mq, k_end, M, a_vec);
        } else {    // stored [M,K]: vector along K
          const int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
          a_reg[it] = ld4(A, lda, m0 + row, k0 + kq, M, k_end, a_vec);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < B_ITERS; ++it) {
      const int v = tid + it * NT;
      if (v < B_VECS) {
        if (TB) {   // stored [N,K]: vector along K
          const int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
          b_reg[it] = ld4(Bm, ldb, n0 + row, k0 + kq, N, k_end, b_vec);
        } else {    // stored [K,N]: vector along N
          const int kr = v / (BN / 4), nq = (v % (BN / 4)) * 4;
          b_reg[it] = ld4(Bm, ldb, k0 + kr, n0 + nq, k_end, N, b_vec);
        }
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int it = 0; it < A_ITERS; ++it) {
      const int v = tid + it * NT;
      if (v < A_VECS) {
        if (TA) {
          const int kr = v / (BM / 4), mq = (v % (BM / 4)) * 4;
          *reinterpret_cast<float4*>(&As[buf][kr][mq]) = a_reg[it];
        } else {
          const int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
          As[buf][kq + 0][row] = a_reg[it].x;
          As[buf][kq + 1][row] = a_reg[it].y;
          As[buf][kq + 2][row] = a_reg[it].z;
          As[buf][kq + 3][row] = a_reg[it].w;
        }
      }
    }
#pragma unroll
    for (int it = 0; it < B_ITERS; ++it) {
      const int v = tid + it * NT;
      if (v < B_VECS) {
        if (TB) {
          const int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
          Bs[buf][kq + 0][row] = b_reg[it].x;
          Bs[buf][kq + 1][row] = b_reg[it].y;
          Bs[buf][kq + 2][row] = b_reg[it].z;
          Bs[buf][kq + 3][row] = b_reg[it].w;
        } else {
          const int kr = v / (BN / 4), nq = (v % (BN / 4)) * 4;
          *reinterpret_cast<float4*>(&Bs[buf][kr][nq]) = b_reg[it];
        }
      }
    }
  };

  const int nk = (k_end - k_begin + BK - 1) / BK;
  load_tiles(k_begin);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tiles(k_begin + (kb + 1) * BK);   // global loads in flight during the FMA block
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
        const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
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

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (elu_y) {
        const float y = elu_y[(size_t)gm * ldy + gn];
        v *= (y > 0.f) ? 1.0f : y + 1.0f;   // tf.nn.elu gradient from the saved activation [upstream EluGrad]
      }
      float* dst = C + (size_t)gm * ldc + gn;
      if (split) atomicAdd(dst, v);
      else if (accumulate) *dst += v;
      else *dst = v;
    }
  }
}

// C (=|+=) op(A) @ op(B).  split_k > 1 splits the contraction over gridDim.z and accumulates with atomicAdd: C must
// then already hold the value to add to (the zeroed gradient buffer) and elu_y must be null.
inline cudaError_t launch_gemm_simt(bool ta, bool tb, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                                    int M, int N, int K, bool accumulate, const float* elu_y, int ldy, int split_k,
                                    cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  if (split_k < 1) split_k = 1;
  const int k_chunk = ((K + split_k - 1) / split_k + 15) / 16 * 16;
  const int nz = (K + k_chunk - 1) / k_chunk;
  const int acc = (accumulate || nz > 1) ? 1 : 0;
#define AIR_GEMM_LAUNCH(BM, BN, TM, TN, TA_, TB_)                                                                   \
  return launch_k(gemm_simt_kernel<BM, BN, 16, TM, TN, TA_, TB_>, dim3((N + BN - 1) / BN, (M + BM - 1) / BM, nz),  \
                  dim3((BM / TM) * (BN / TN)), 0, st, A, lda, B, ldb, C, ldc, M, N, K, k_chunk, acc, elu_y, ldy)
  if (N > 32 && M > 32 && !ta && tb && (long long)((N + 63) / 64) * ((M + 127) / 128) * nz < 148) {
    AIR_GEMM_LAUNCH(32, 64, 4, 4, false, true);   // input gradients of small layers: 32-row tiles fill the machine
  }
  if (N > 32) {
    if (!ta && tb) { AIR_GEMM_LAUNCH(128, 64, 8, 4, false, true); }
    if (ta && !tb) { AIR_GEMM_LAUNCH(128, 64, 8, 4, true, false); }
    if (!ta && !tb) { AIR_GEMM_LAUNCH(128, 64, 8, 4, false, false); }
  } else {
    if (!ta && tb) { AIR_GEMM_LAUNCH(128, 16, 4, 4, false, true); }
    if (ta && !tb) { AIR_GEMM_LAUNCH(128, 16, 4, 4, true, false); }
    if (!ta && !tb) { AIR_GEMM_LAUNCH(128, 16, 4, 4, false, false); }
  }
#undef AIR_GEMM_LAUNCH
  return cudaErrorInvalidValue;
}

// out[n] += sum_m X[m, n]   (bias gradients; rows split over gridDim.y, atomicAdd into the zeroed gradient buffer)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int ldx, float* __restrict__ out, int M, int N, int rows_per_cta) {
  __shared__ float s[8][33];
  griddep_launch();
  griddep_wait();
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float v = 0.f;
  if (col < N)
    for (int r = r0 + rl; r < r1; r += 8) v += X[(size_t)r * ldx + col];
  s[rl][threadIdx.x & 31] = v;
  __syncthreads();
  if (rl == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
    atomicAdd(out + col, t);
  }
}
inline cudaError_t launch_colsum(const float* X, int ldx, float* out, int M, int N, cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  // enough CTAs to pull the operand at HBM speed: ~1200 CTAs for a [12288, 256] gradient
  int rows_per_cta = 512;
  while (rows_per_cta > 64 && (long long)((N + 31) / 32) * ((M + rows_per_cta - 1) / rows_per_cta) < 1184) rows_per_cta >>= 1;
  return launch_k(colsum_kernel, dim3((N + 31) / 32, (M + rows_per_cta - 1) / rows_per_cta), dim3(256), 0, st, X, ldx, out,
                  M, N, rows_per_cta);
}

}  // namespace air
