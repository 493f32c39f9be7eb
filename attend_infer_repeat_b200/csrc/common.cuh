// Shared device helpers for the AIR hot path (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

namespace air {

constexpr int ACT_NONE = 0;
constexpr int ACT_ELU = 1;

// tf.nn.elu [upstream Eigen functor]: x if x > 0 else exp(x) - 1  (neural.py:20)
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expf(x) - 1.0f; }

// tf.nn.softplus [upstream functor]: threshold = log(eps) + 2; x > -thr -> x; x < thr -> exp(x);
// else log(exp(x) + 1)   (NormalWithSoftplusScale, cell.py:130, modules.py:17)
__device__ __forceinline__ float softplus_f(float x) {
  const float thr = -13.942385f;  // logf(FLT_EPSILON) + 2
  if (x > -thr) return x;
  float ex = expf(x);
  if (x < thr) return ex;
  return logf(ex + 1.0f);
}

// the same three branches with MUFU.EX2 / MUFU.LG2 (ex2.approx, lg2.approx): absolute error <= 4e-7 -- the 16 accurate
// expf + logf pairs per thread were two thirds of the row kernel's what-head epilogue (14 k clocks in which the tensor
// core waits for the sampled code)
__device__ __forceinline__ float softplus_lean(float x) {
  const float thr = -13.942385f;
  if (x > -thr) return x;
  float ex, lg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(x * 1.4426950408889634f));
  if (x < thr) return ex;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(ex + 1.0f));
  return lg * 0.6931471805599453f;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// visualisation-only sigmoid (glimpse_viz, model.py:90): ex2.approx + fast division, ~2^-21 relative
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// four instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP; flush-to-zero so no range fix-ups): 1 / (1 + 2^(-x log2 e)), ~3e-7
// relative; exact limits 0 and 1 at -inf / +inf.  For visualisation outputs.
__device__ __forceinline__ float sigmoid_lean(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act) { return act == ACT_ELU ? elu_f(v) : v; }

// fp16x2 split operand format of the tensor-core engine (linear_tc.cuh): x = hi + lo, hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
// Optional second output of a producer kernel: the consumer GEMM's A operand in "hl" format
// (fp16 [2][rows_alloc][ld]: hi plane, then lo plane; see linear_tc.cuh).  p == nullptr -> not written.
// nsl > 0 selects the slice-major tiled layout read by the fused chains (chain_tc.cuh):
// [row tile of 128][K slice of 16][128 rows][16 fp16] per plane, nsl = K slices per row.
struct HlOut {
  __half* p;
  size_t plane;
  int ld;
  int nsl;
};
__device__ __forceinline__ size_t hl_index(const HlOut& o, size_t row, int col) {
  if (o.nsl > 0) return (((row >> 7) * (size_t)o.nsl + (size_t)(col >> 4)) * 128 + (row & 127)) * 16 + (col & 15);
  return row * (size_t)o.ld + col;
}
__device__ __forceinline__ void hl_store(const HlOut& o, size_t row, int col, float x) {
  __half hi, lo;
  split_f16(x, hi, lo);
  __half* d = o.p + hl_index(o, row, col);
  d[0] = hi;
  d[o.plane] = lo;
}

// NaN-propagating max (fmaxf drops NaNs)
__device__ __forceinline__ float fmax_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// Programmatic dependent launch (PDL).  Every kernel of the forward chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may start while the previous kernel drains, run their
// prologue (barrier init, TMEM allocation, tensor-map prefetch, index setup) and then block in griddep_wait() until the
// previous grid has completed and its writes are visible.  griddep_launch() lets the NEXT kernel do the same with us.
// Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = getenv("AIR_NO_PDL") == nullptr;
  return on;
}
// Per-thread switch consulted by launch_k, set for the duration of a call by the entry points that act on a handle
// (air_set_launch_overlap).  Programmatic dependent launch lets the next kernel's CTAs take their SMs before the previous
// kernel has finished: a gain for ONE stream (prologues overlap tails), a loss when several handles share the device --
// an early tensor-kernel CTA then holds a whole SM while it only waits, and a neighbouring batch's kernel cannot use it.
inline bool& pdl_thread_flag() {
  static thread_local bool on = true;
  return on;
}
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(pdl_thread_flag()) { pdl_thread_flag() = on; }
  ~PdlScope() { pdl_thread_flag() = prev; }
};

#ifdef __CUDACC__
// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: the opt-in is remembered per (device, kernel),
// not per process, so a second handle on another GPU of the same process gets it too.
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kernel, size_t bytes) {
  // keyed by (device, kernel ADDRESS): instantiations of one kernel template share a function-pointer type
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> configured;
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(kernel));
  std::lock_guard<std::mutex> lock(mu);
  auto it = configured.find(key);
  if (it != configured.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) configured[key] = bytes;
  return e;
}

// <<<>>> replacement that sets the PDL attribute (or not)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && pdl_thread_flag()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; every thread gets the result.  `red` must hold >= 32 floats.  Ends with a barrier.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// Sonnet builds the warp grid with np.linspace(-1, 1, n, dtype=float32): float64 maths, cast at the end.
__device__ __forceinline__ float linspace_pm1(int i, int n) {
  if (n <= 1) return -1.0f;
  double step = 2.0 / (double)(n - 1);
  return (float)(-1.0 + (double)i * step);
}

// snt.resampler [upstream C++/CUDA kernel] on one single-channel plane held in (shared or global) memory:
// bilinear sample at pixel coords (x, y); zero outside (-1, Ws) x (-1, Hs); taps outside the plane read 0.
__device__ __forceinline__ float resample_plane(const float* __restrict__ D, int Hs, int Ws, float x, float y) {
  if (!(x > -1.0f && y > -1.0f && x < (float)Ws && y < (float)Hs)) return 0.f;
  const float fxf = floorf(x), fyf = floorf(y);
  const int fx = (int)fxf, fy = (int)fyf, cx = fx + 1, cy = fy + 1;
  const float dx = (fxf + 1.0f) - x, dy = (fyf + 1.0f) - y;
  const bool fx_ok = fx >= 0 && fx <= Ws - 1, cx_ok = cx >= 0 && cx <= Ws - 1;
  const bool fy_ok = fy >= 0 && fy <= Hs - 1, cy_ok = cy >= 0 && cy <= Hs - 1;
  const float v_ff = (fx_ok && fy_ok) ? D[fy * Ws + fx] : 0.f;
  const float v_cc = (cx_ok && cy_ok) ? D[cy * Ws + cx] : 0.f;
  const float v_fc = (fx_ok && cy_ok) ? D[cy * Ws + fx] : 0.f;   // (fx, cy)
  const float v_cf = (cx_ok && fy_ok) ? D[fy * Ws + cx] : 0.f;   // (cx, fy)
  const float one = 1.0f;
  // same association as the reference kernel: dx*dy*ff + (1-dx)(1-dy)*cc + dx(1-dy)*fc + (1-dx)dy*cf
  float r = __fmul_rn(__fmul_rn(dx, dy), v_ff);
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(one - dx, one - dy), v_cc));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(dx, one - dy), v_fc));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(one - dx, dy), v_cf));
  return r;
}

// AffineGridWarper(source = img HxW, output = hxw, no_shear_2d): pixel coordinate along one axis,
//   coord = s * (u * S) + t * S + S,  S = (n_src - 1) / 2,  u = linspace(-1, 1, n_out)[i]
__device__ __forceinline__ float fwd_coord(float s, float t, int i, int n_out, int n_src) {
  const float S = ((float)n_src - 1.0f) * 0.5f;
  const float uS = __fmul_rn(linspace_pm1(i, n_out), S);
  return __fadd_rn(__fadd_rn(__fmul_rn(s, uS), __fmul_rn(t, S)), S);
}

// fwd_coord with np.linspace's float64 step 2/(n_out - 1) computed once on the host (same value, no per-tap division)
__device__ __forceinline__ float fwd_coord_s(float s, float t, int i, double step, int n_src) {
  const float S = ((float)n_src - 1.0f) * 0.5f;
  const float uS = __fmul_rn((float)(-1.0 + (double)i * step), S);
  return __fadd_rn(__fadd_rn(__fmul_rn(s, uS), __fmul_rn(t, S)), S);
}

// AffineGridWarper.inverse() for the no-shear case (a = sx, d = sy, b = c = 0):
//   det = sx * sy; a' = sy / det; d' = sx / det; tx' = a' * tx; ty' = d' * ty
//   x_g = a' * (U * S_w) + (-tx') * S_w + S_w   (U = linspace(-1, 1, W) over canvas columns), y likewise.
__device__ __forceinline__ void inv_params(float sx, float tx, float sy, float ty, float& a_inv, float& d_inv,
                                           float& ntx, float& nty) {
  const float det = __fmul_rn(sx, sy);
  a_inv = __fdiv_rn(sy, det);
  d_inv = __fdiv_rn(sx, det);
  ntx = -__fmul_rn(a_inv, tx);
  nty = -__fmul_rn(d_inv, ty);
}
__device__ __forceinline__ float inv_coord(float a_inv, float nt, int i, int n_canvas, int n_glimpse) {
  const float S = ((float)n_glimpse - 1.0f) * 0.5f;
  const float US = __fmul_rn(linspace_pm1(i, n_canvas), S);
  return __fadd_rn(__fadd_rn(__fmul_rn(a_inv, US), __fmul_rn(nt, S)), S);
}

// inv_coord with np.linspace's float64 step 2/(n_canvas - 1) computed once on the host (same value, no per-tap division)
__device__ __forceinline__ float inv_coord_s(float a_inv, float nt, int i, double step, int n_glimpse) {
  const float S = ((float)n_glimpse - 1.0f) * 0.5f;
  const float U = (float)(-1.0 + (double)i * step);
  const float US = __fmul_rn(U, S);
  return __fadd_rn(__fadd_rn(__fmul_rn(a_inv, US), __fmul_rn(nt, S)), S);
}

}  // namespace air
