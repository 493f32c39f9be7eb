// Stand-alone micro-benchmark + phase tracer of the tcgen05 dense-layer kernel (linear_tc.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DAIR_TC_TRACE -Iinclude tools/tc_trace.cu -o /tmp/tc_trace
//   /tmp/tc_trace M N K [iters]
// Prints the CUDA-event time per launch (back-to-back launches, warm) and, for a few CTAs, the SM-clock timestamps of
// the pipeline phases (kernel start, setup done, first TMA, last TMA, first/last stage consumed, last MMA commit,
// accumulator ready, epilogue done, teardown).  Operands are random hl planes; results are not checked here (the
// parity tests do that) -- this tool only answers "where does a CTA spend its time".
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../attend_infer_repeat_b200/csrc/linear_tc.cuh"

using namespace air::tc;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                         \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

__global__ void fill_kernel(__half* p, size_t n, unsigned seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)i * 2654435761u + seed;
  x ^= x >> 13;
  x *= 0x5bd1e995u;
  x ^= x >> 15;
  p[i] = __float2half(((x & 0xffff) / 65536.0f - 0.5f) * 0.25f);
}

template <int BN, int STAGES>
int run(int M, int N, int K, int iters, bool want_f32, bool want_hl) {
  const int Kpad = round_up(K, BK), N_alloc = round_up(N, BN), M_alloc = round_up(M, BM);
  const int ld_out = round_up(N, 64);
  __half *a, *w, *ohl;
  float *of32, *bias;
  long long* trace;
  const size_t a_n = 2 * (size_t)M_alloc * Kpad, w_n = 2 * (size_t)N_alloc * Kpad, o_n = 2 * (size_t)M_alloc * ld_out;
  CK(cudaMalloc(&a, a_n * 2));
  CK(cudaMalloc(&w, w_n * 2));
  CK(cudaMalloc(&ohl, o_n * 2));
  CK(cudaMalloc(&of32, (size_t)M * N * 4));
  CK(cudaMalloc(&bias, N_alloc * 4));
  CK(cudaMemset(bias, 0, N_alloc * 4));
  dim3 grid(N_alloc / BN, (M + BM - 1) / BM);
  const int n_ctas = grid.x * grid.y;
  CK(cudaMalloc(&trace, (size_t)n_ctas * 16 * 8));
  CK(cudaMemset(trace, 0, (size_t)n_ctas * 16 * 8));
  fill_kernel<<<(unsigned)((a_n + 255) / 256), 256>>>(a, a_n, 1);
  fill_kernel<<<(unsigned)((w_n + 255) / 256), 256>>>(w, w_n, 2);
  CUtensorMap tm_a, tm_b;
  if (!make_tmap(&tm_a, a, Kpad, 2 * (int64_t)M_alloc, BM) || !make_tmap(&tm_b, w, Kpad, 2 * (int64_t)N_alloc, BN)) {
    printf("tensor map creation failed\n");
    return 1;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias;
  p.out_f32 = want_f32 ? of32 : nullptr;
  p.ldc = N;
  p.out_hl = want_hl ? ohl : nullptr;
  p.hl_plane = (size_t)M_alloc * ld_out;
  p.ld_hl = ld_out;
  p.M = M;
  p.N = N;
  p.num_k_blocks = Kpad / BK;
  p.a_lo_row = M_alloc;
  p.b_lo_row = N_alloc;
  p.act = 1;
  p.trace = trace;
  for (int i = 0; i < 3; ++i) CK((launch_gemm_cfg<BN, STAGES>(tm_a, tm_b, p, N_alloc, 0)));
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) CK((launch_gemm_cfg<BN, STAGES>(tm_a, tm_b, p, N_alloc, 0)));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = 1e3 * ms / iters;
  const double flops = 2.0 * M * (double)N * K;
  printf("BN=%d STAGES=%d M=%d N=%d K=%d grid=(%d,%d) f32=%d hl=%d : %.2f us/launch, %.1f TFLOP/s (algorithmic), "
         "%.1f TFLOP/s (3 MMAs)\n",
         BN, STAGES, M, N, K, grid.x, grid.y, (int)want_f32, (int)want_hl, us, flops / us * 1e-6, 3 * flops / us * 1e-6);
  std::vector<long long> t((size_t)n_ctas * 16);
  CK(cudaMemcpy(t.data(), trace, t.size() * 8, cudaMemcpyDeviceToHost));
  const char* names[16] = {"start", "setup", "tma0", "tmaN", "full0", "fullN", "commit", "accrdy", "epi", "teardn",
                           "ld0", "math0", "-", "ld1", "math1", "-"};
  for (int c : {0, n_ctas / 2, n_ctas - 1}) {
    printf("  cta %4d:", c);
    for (int i : {1, 2, 4, 5, 6, 7, 10, 11, 13, 14, 8}) printf(" %s+%lld", names[i], t[(size_t)c * 16 + i] - t[(size_t)c * 16]);
    printf("\n");
  }
  cudaFree(a); cudaFree(w); cudaFree(ohl); cudaFree(of32); cudaFree(bias); cudaFree(trace);
  return 0;
}

int main(int argc, char** argv) {
  int M = argc > 1 ? atoi(argv[1]) : 12288, N = argc > 2 ? atoi(argv[2]) : 256, K = argc > 3 ? atoi(argv[3]) : 256;
  int iters = argc > 4 ? atoi(argv[4]) : 20;
  int rc = 0;
  rc |= run<64, 2>(M, N, K, iters, false, true);
  rc |= run<64, 2>(M, N, K, iters, true, false);
  return rc;
}
