// C ABI of the B200-native AIR hot path (see include/air_b200.h for the contract and the reference
// interfaces each entry point replaces).
//
// Data flow of one air_forward (B canvases, T steps), re-associated for the GPU -- not the reference's
// per-step graph order (cell.py:116-171):
//   * the input encoder sees only the raw image (cell.py:121-125), so e = Encoder(img) is computed ONCE;
//   * the LSTM's input half is step-invariant too: gx = e @ W[:n_enc] + b once, then per step only
//     gates = gx + h_{t-1} @ W[n_enc:]  (the only truly sequential chain: T small GEMMs + gate math);
//   * nothing downstream of h_t feeds back into the recurrence (what/where/canvas are outputs only), so all
//     heads, the glimpse read, the glimpse VAE and the paint run ONCE over the T*B stacked rows [T,B,.];
//   * the canvas accumulates in registers over t inside the paint kernel, which also produces every
//     per-sample ELBO term; a last tiny kernel forms the batch means.
// Every row's arithmetic is unchanged; only the batching differs.
//
// Two dense-layer engines share that data flow (air_config.precision):
//   AIR_PREC_FP32     linear_simt.cuh  fp32 FMA, activations cross HBM as fp32 rows
//   AIR_PREC_TC_SPLIT linear_tc.cuh    tcgen05 fp16x2-split MMAs, activations cross HBM as "hl" fp16 planes
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <utility>
#include <vector>

#include "../../include/air_b200.h"
#include "backward_kernels.cuh"
#include "cell_kernels.cuh"
#include "chain_tc.cuh"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "linear_simt.cuh"
#include "linear_tc.cuh"
#include "enc_tc.cuh"
#include "lstm_tc.cuh"
#include "row_tc.cuh"

namespace {

thread_local std::string g_last_error;

int32_t fail(air_status st, const std::string& msg) {
  g_last_error = msg;
  return (int32_t)st;
}

#define AIR_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(AIR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
  } while (0)

using air::tc::round_up;

struct Layer {
  int64_t w_off = 0, b_off = -1;   // float offsets into the flat parameter buffer (b_off < 0: no bias)
  int K = 0, N = 0;
  int tc = -1;                     // index into air_handle::tcw (prepared tensor-core weight) or -1
};
struct Mlp {
  std::vector<Layer> layers;   // hidden layers (ELU) followed by the optional linear output layer
  int n_hidden = 0;
};
struct ParamEntry {
  std::string name;
  int64_t offset;
  int rows, cols;
};

// One activation matrix in the representation(s) the active engine needs.
struct Buf {
  float* f32 = nullptr;   // fp32 rows (SIMT engine operands; final outputs of either engine)
  int ld = 0;
  __half* hl = nullptr;   // "hl" fp16 planes (tensor-core engine operands), see linear_tc.cuh
  int rows_alloc = 0;     // rows per plane
  int kpad = 0;           // row pitch in halves = round_up(width, 64)
  size_t plane() const { return (size_t)rows_alloc * kpad; }
  air::HlOut hl_out() const { return air::HlOut{hl, plane(), kpad, 0}; }
  // slice-major tiled copy for the fused chains (chain_tc.cuh): [rows_alloc / 128][nsl][128][16] fp16 per plane
  __half* hlt = nullptr;
  int nsl = 0;
  size_t plane_t() const { return (size_t)rows_alloc * nsl * 16; }
  air::HlOut hlt_out() const { return air::HlOut{hlt, plane_t(), 0, nsl}; }
  bool hl_tiled = false;   // as the OUTPUT of dense(): write the hl result into hlt (slice-major tiles) instead of hl
};

// A weight matrix prepared for the tensor-core engine: W^T, fp16 split of (w * 2^8), [2][N_alloc][Kpad].
struct TcWeight {
  int64_t src_off = 0;
  int K = 0, N = 0, Kpad = 0, N_alloc = 0, BN = 0;
  int64_t arena_off = 0;   // halves
  CUtensorMap tm;
  // chain_tc.cuh view of the same prepared weight: n_pass tiles of n_box rows
  int split_n = 0, split_off = 0;   // what head only: scale half starts at row split_off
  int n_box = 0, n_pass = 0;
  CUtensorMap tm_chain;
  CUtensorMap tm_row;                    // row_tc.cuh view: boxes of 128 rows x 64 K
  int64_t bias_src = -1, bias_off = 0;   // bias in params; zero-padded copy in the bias arena (floats)
  int perm_nh = 0;                       // lstm_tc.cuh column regrouping
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct air_handle {
  air_config cfg;
  int P = 0, G = 0, n_enc = 0;
  bool use_tc = false;
  std::vector<ParamEntry> entries;
  int64_t n_params = 0;
  Mlp enc, where_mlp, steps_mlp, glenc, dec;
  Layer what_lin, lstm_x, lstm_h;
  Layer what_chain;                // what_lin with the loc / scale halves on 16-row boundaries (chain_tc.cuh)
  bool chain_ok = false;           // the fused-chain kernels cover this configuration
  bool row_ok = false;             // the fused row kernel (row_tc.cuh) covers this configuration
  bool enc1_ok = false;            // the split-K first encoder layer (enc_tc.cuh) covers this configuration
  int na_off = 0;
  bool lstm_ok = false;            // the cluster LSTM kernel (lstm_tc.cuh) covers this configuration
  int lstm_x_perm = -1, lstm_h_perm = -1;   // tcw indices of the regrouped W[:n_enc] / W[n_enc:]
  float* gx_scr = nullptr;
  __half* hx = nullptr;
  int64_t lstm_w = 0, lstm_b = 0, lstm_h0 = 0, lstm_c0 = 0;
  int max_width = 0;
  // workspace (one cudaMalloc)
  char* ws = nullptr;
  size_t ws_bytes = 0;
  // activations
  Buf ping, pong;                                  // hidden activations of the MLP chains [T*B, max_width]
  Buf x, e, h_init, hs, crop, what_in;             // GEMM A operands produced by non-GEMM kernels (+ e)
  float *gx = nullptr, *gates = nullptr, *cbuf = nullptr, *m = nullptr, *logit = nullptr, *r = nullptr;
  float* prior_part = nullptr;   // [B] prior part of the per-sample loss (paint grid -> elbo_scalars_kernel)
  // staging for air_forward_host / air_cell_step
  uint8_t* st_img_u8 = nullptr;
  float *st_img = nullptr, *st_eps_where = nullptr, *st_eps_what = nullptr, *st_u = nullptr, *st_pres_in = nullptr;
  // tensor-core engine state
  std::vector<TcWeight> tcw;
  __half* arena = nullptr;
  float* bias_arena = nullptr;
  air::tc::PrepEntry* prep_table = nullptr;
  int prep_tiles = 0;
  int* range_flag = nullptr;
  bool launch_overlap = true;      // programmatic dependent launch between the kernels of a pass (air_set_launch_overlap)
  float* hw0 = nullptr;            // h0 @ W_h of the cluster LSTM (lstm_h0w_kernel), rebuilt with the weight arena
  std::map<std::pair<const void*, int>, CUtensorMap> tmap_cache;
  // training (air_train_enable / air_backward; either engine): saved activations + gradient scratch, one cudaMalloc
  // inference: the prepared fp16-split weight arena is reused while the caller vouches that `params` is unchanged
  bool cache_weights = false;
  const float* weights_ready = nullptr;
  const float* hw0_ready = nullptr;   // parameters for which hw0 (lstm_h0w_kernel) is current
  double* prior_dev = nullptr;        // air_prior_table_device: geometric_prior table in device memory ([AIR_MAX_STEPS + 1])
  bool prior_dev_on = false;
  bool train = false;
  bool fwd_saved = false;          // the last forward on this handle ran in training mode with a prior (backward is valid)
  char* tws = nullptr;
  size_t tws_bytes = 0;
  std::vector<float*> sv_enc, sv_where, sv_steps, sv_glenc, sv_dec;   // hidden activations of each MLP (layer outputs)
  float *sv_q = nullptr;           // glimpse-encoder output [T*B, n_gl]
  float *gates_all = nullptr;      // [T,B,4nh] pre-activation gates of every step
  float *c_all = nullptr;          // [T+1,B,nh] cell states (slice 0 = initial)
  float *hprev = nullptr;          // [T,B,nh] h_{t-1} of every step (slice 0 = initial)
  float *g_a = nullptr, *g_b = nullptr;            // [T*B, max_width] gradient ping-pong
  float *g_glimpse = nullptr, *g_crop = nullptr;   // [T*B, G]
  float *g_what = nullptr, *g_r = nullptr;         // [T*B, na], [T*B, 2na]
  float *g_wh_paint = nullptr, *g_wh_read = nullptr, *g_m = nullptr, *g_logit = nullptr;   // [T*B,4] x2, [T*B,8], [T*B]
  float* g_pres = nullptr;         // [T*B] d rec / d presence (non-discrete steps)
  float *g_h = nullptr, *g_gates = nullptr;        // [T*B, nh], [T*B, 4nh]
  float *g_gx = nullptr, *g_hrec = nullptr, *g_c = nullptr, *g_e = nullptr;   // [B,4nh], [B,nh], [B,nh], [B,n_enc]
  // tensor-core weight gradients (dW = X^T @ dY on the tcgen05 split engine): transposed hl operands, M contiguous
  bool tc_bwd = false;
  static constexpr int MAX_SIDE = 4;
  __half *hl_xt[MAX_SIDE] = {}, *hl_yt[MAX_SIDE] = {};   // one pair of transposed-operand buffers per side stream
  size_t hl_xt_halves = 0, hl_yt_halves = 0;   // per plane
  // tensor-core input gradients (dX = dY @ W^T): dY row-major planes, W as stored ([in][out]) planes per layer
  __half *hl_dy2[2] = {nullptr, nullptr}, *wnt_arena = nullptr;   // dY planes, double buffered: a dX GEMM's epilogue
                                                                  // writes the planes the next dX GEMM reads
  int dy_ready_buf = 0;
  size_t hl_dy_halves = 0;
  std::map<int64_t, std::pair<size_t, int>> wnt_index;   // Layer::w_off -> (half offset of the hi plane, Npad)
  air::tc::RowsEntry* wnt_table = nullptr;               // device copy of the per-matrix table (prep_weights_rows_kernel)
  int wnt_entries = 0, wnt_blocks = 0;
  // the weight-gradient work of a backward pass runs on side streams, beside the dX / pointwise critical path: the layers'
  // weight gradients are independent of each other, so they are dealt round-robin to n_side streams (AIR_SIDE_STREAMS,
  // default 2; each launch alone is too small to fill the machine)
  cudaStream_t side[MAX_SIDE] = {};
  int n_side = 0, side_next = 0;
  int n_side_bufs = 1;             // operand buffer pairs carved into the training workspace
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  std::map<const float*, cudaEvent_t> dy_consumed;       // gradient buffer -> "its last reader on a side stream has re-laid it out"
  const float* dy_ready = nullptr;                       // dY whose row-major planes currently sit in hl_dy2[dy_ready_buf]
  int dy_ready_m = 0, dy_ready_n = 0;                    // ... with these dimensions
  int* t_range_flag = nullptr;
  std::map<std::tuple<const void*, int, long long, int>, CUtensorMap> tmap_cache2;
  // BaselineMLP on the engine (air_baseline_attach / _forward / _backward; modules.py:125-143, model.py:253-259)
  struct Baseline {
    bool attached = false;
    int n_in = 0;
    Mlp mlp;                         // offsets into the caller's flat baseline parameter / gradient buffers
    int64_t n_params = 0;
    char* ws = nullptr;
    Buf x;                           // gathered input rows: fp32 [B, n_in] (+ hl planes for the tensor-core first layer)
    std::vector<float*> act;         // hidden activations [B, N_i] fp32
    float* g[2] = {nullptr, nullptr};   // gradient ping-pong [B, widest layer]
    __half* w_arena = nullptr;       // prepared first-layer weight (hl planes)
    air::tc::PrepEntry* prep_table = nullptr;
    int prep_tiles = 0;
  } bl;
  // instrumentation: kernel-launch counter and optional per-stage CUDA-event timing (air_profile_*)
  uint64_t launches = 0;
  bool profile = false;
  cudaEvent_t ev[AIR_N_STAGES + 1] = {};
  long long* trace = nullptr;      // AIR_CHAIN_TRACE=<file prefix>: per-group clock stamps of the chain kernels (debug)
  int trace_seq = 0;
  // double-buffered host feed (air_feed_host_u8 / air_forward_fed_u8_rng / air_feed_wait), created on first use
  cudaStream_t feed_stream = nullptr;
  uint8_t* feed_buf[2] = {nullptr, nullptr};       // one cudaMalloc, two uint8 batches
  cudaEvent_t feed_fed[2] = {}, feed_consumed[2] = {}, feed_done[2] = {};
  bool feed_pending[2] = {false, false};           // a copy into the slot has been enqueued and not yet run through a pass
};

namespace {

int64_t add_entry(air_handle* h, const std::string& name, int rows, int cols) {
  const int64_t off = h->n_params;
  h->entries.push_back({name, off, rows, cols});
  h->n_params += (int64_t)rows * cols;
  return off;
}

int build_mlp(air_handle* h, Mlp& mlp, const char* prefix, int n_in, const int32_t* hidden, int n_hidden, int n_out) {
  int d = n_in;
  for (int i = 0; i < n_hidden; ++i) {
    Layer l;
    l.K = d;
    l.N = hidden[i];
    l.w_off = add_entry(h, std::string(prefix) + "." + std::to_string(i) + ".w", d, hidden[i]);
    l.b_off = add_entry(h, std::string(prefix) + "." + std::to_string(i) + ".b", 1, hidden[i]);
    mlp.layers.push_back(l);
    d = hidden[i];
    if (d > h->max_width) h->max_width = d;
  }
  mlp.n_hidden = n_hidden;
  if (n_out > 0) {
    Layer l;
    l.K = d;
    l.N = n_out;
    l.w_off = add_entry(h, std::string(prefix) + ".out.w", d, n_out);
    l.b_off = add_entry(h, std::string(prefix) + ".out.b", 1, n_out);
    mlp.layers.push_back(l);
    d = n_out;
  }
  return d;
}

bool valid_hidden(const int32_t* v, int n) {
  if (n < 1 || n > AIR_MAX_HIDDEN) return false;
  for (int i = 0; i < n; ++i)
    if (v[i] < 1) return false;
  return true;
}

inline void mark(air_handle* h, int stage, cudaStream_t st) {
  if (h->profile) cudaEventRecord(h->ev[stage], st);
}

void register_tc_weight(air_handle* h, Layer& l, bool feeds_gemm) {
  TcWeight w;
  w.src_off = l.w_off;
  w.K = l.K;
  w.N = l.N;
  w.Kpad = round_up(l.K, air::tc::BK);
  w.BN = (l.N <= 32 && !feeds_gemm) ? 32 : 64;   // hl outputs (operands of a following GEMM) need the 64-wide tile
  w.N_alloc = round_up(l.N, w.BN);
  w.bias_src = l.b_off;
  l.tc = (int)h->tcw.size();
  h->tcw.push_back(w);
}

int32_t get_tmap_a(air_handle* h, const Buf& b, const CUtensorMap** out) {
  const auto key = std::make_pair((const void*)b.hl, b.kpad);
  auto it = h->tmap_cache.find(key);
  if (it == h->tmap_cache.end()) {
    CUtensorMap tm;
    if (!air::tc::make_tmap(&tm, b.hl, b.kpad, 2 * (int64_t)b.rows_alloc, air::tc::BM))
      return fail(AIR_ERR_CUDA, "cuTensorMapEncodeTiled failed for an activation buffer");
    it = h->tmap_cache.emplace(key, tm).first;
  }
  *out = &it->second;
  return AIR_OK;
}

// One dense layer: out = act(in[row0 : row0 + M] @ W + bias (+ addend)).  `want_f32` / `want_hl` select which
// representation(s) of the result are materialised; every GEMM of the path goes through here.
int32_t dense(air_handle* h, const float* params, const Buf& in, int row0, const Layer& l, bool use_bias,
              const float* addend, int ldadd, const Buf& out, bool want_f32, bool want_hl, int M, int act,
              cudaStream_t st) {
  ++h->launches;
  const float* bias = (use_bias && l.b_off >= 0) ? params + l.b_off : nullptr;
  if (!h->use_tc) {
    AIR_CUDA(air::launch_linear_simt(in.f32 + (size_t)row0 * in.ld, in.ld, params + l.w_off, l.N, bias, addend, ldadd,
                                     out.f32, out.ld, M, l.N, l.K, act, st));
    return AIR_OK;
  }
  const TcWeight& w = h->tcw[l.tc];
  if (in.kpad != w.Kpad) return fail(AIR_ERR_ARG, "internal: operand pitch does not match the prepared weight");
  const CUtensorMap* tm_a = nullptr;
  const int32_t rc = get_tmap_a(h, in, &tm_a);
  if (rc != AIR_OK) return rc;
  air::tc::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias;
  p.addend = addend;
  p.ldadd = ldadd;
  p.out_f32 = want_f32 ? out.f32 : nullptr;
  p.ldc = out.ld;
  p.out_hl = want_hl ? out.hl : nullptr;
  p.hl_plane = out.plane();
  p.ld_hl = out.kpad;
  if (want_hl && out.hl_tiled) {   // slice-major tiles [row tile][slice][128][16]: the consumer reads whole 1 KB runs
    p.out_hl = out.hlt;
    p.hl_plane = out.plane_t();
    p.hl_nsl = out.nsl;
  }
  p.M = M;
  p.N = l.N;
  p.num_k_blocks = w.Kpad / air::tc::BK;
  p.a_lo_row = in.rows_alloc;
  p.a_row0 = row0;
  p.b_lo_row = w.N_alloc;
  p.act = act;
  p.range_flag = h->range_flag;
  if (want_hl && w.BN != 64) return fail(AIR_ERR_ARG, "internal: hl output needs a 64-wide tile");
  AIR_CUDA(air::tc::launch_gemm(w.BN, *tm_a, w.tm, p, w.N_alloc, st));
  return AIR_OK;
}

// neural.MLP (neural.py:63-102): ELU hidden layers, linear output layer.  Hidden activations ping-pong between the two
// workspace buffers (fp32 rows or hl planes depending on the engine); the last layer writes `out`.
int32_t run_mlp_from(air_handle* h, const float* params, const Mlp& mlp, const Buf& in, int M, const Buf& out,
                     bool out_f32, bool out_hl, cudaStream_t st, const std::vector<float*>* saves, bool first_to_ping);
int32_t run_mlp(air_handle* h, const float* params, const Mlp& mlp, const Buf& in, int M, const Buf& out,
                bool out_f32, bool out_hl, cudaStream_t st, const std::vector<float*>* saves = nullptr) {
  return run_mlp_from(h, params, mlp, in, M, out, out_f32, out_hl, st, saves, true);
}
// first_to_ping = false: `in` already sits in the ping buffer (the split-K first encoder layer wrote it there)
int32_t run_mlp_from(air_handle* h, const float* params, const Mlp& mlp, const Buf& in, int M, const Buf& out,
                     bool out_f32, bool out_hl, cudaStream_t st, const std::vector<float*>* saves, bool first_to_ping) {
  Buf cur = in;
  bool to_ping = first_to_ping;
  const int nl = (int)mlp.layers.size();
  for (int i = 0; i < nl; ++i) {
    const Layer& l = mlp.layers[i];
    const bool last = (i == nl - 1);
    Buf dst;
    if (last) {
      dst = out;
    } else if (saves) {   // training mode: every hidden activation is kept (fp32 rows) for the backward pass
      if (h->use_tc) {
        dst = to_ping ? h->ping : h->pong;   // + the hl planes the next tensor-core layer reads
        to_ping = !to_ping;
        dst.kpad = round_up(l.N, air::tc::BK);
      }
      dst.f32 = (*saves)[i];
      dst.ld = l.N;
    } else {
      dst = to_ping ? h->ping : h->pong;
      to_ping = !to_ping;
      dst.ld = l.N;
      dst.kpad = round_up(l.N, air::tc::BK);
    }
    const int act = (i < mlp.n_hidden) ? air::ACT_ELU : air::ACT_NONE;
    const int32_t rc = dense(h, params, cur, 0, l, true, nullptr, 0, dst, last ? out_f32 : (!h->use_tc || saves != nullptr),
                             last ? out_hl : h->use_tc, M, act, st);
    if (rc != AIR_OK) return rc;
    cur = dst;
  }
  return AIR_OK;
}

int32_t check_outs(const air_outputs* o, bool need_elbo) {
  if (!o) return fail(AIR_ERR_ARG, "air_outputs is NULL");
  if (!o->glimpse || !o->what || !o->what_loc || !o->what_scale || !o->where || !o->where_loc || !o->where_scale ||
      !o->presence_prob || !o->presence)
    return fail(AIR_ERR_ARG, "air_outputs: glimpse/what*/where*/presence* buffers are mandatory");
  if (need_elbo &&
      (!o->num_steps_posterior || !o->num_step_per_sample || !o->prior_step_weight || !o->rec_loss_per_sample ||
       !o->kl_num_steps_per_sample || !o->kl_what_per_sample || !o->kl_where_per_sample || !o->loss_per_sample ||
       !o->num_steps_log_prob || !o->scalars))
    return fail(AIR_ERR_ARG, "air_outputs: ELBO buffers are mandatory when a prior is given");
  return AIR_OK;
}

int32_t check_outs_elbo_only(const air_outputs* o) {
  if (!o || !o->num_step_per_sample || !o->rec_loss_per_sample || !o->kl_num_steps_per_sample ||
      !o->kl_what_per_sample || !o->kl_where_per_sample || !o->num_steps_log_prob || !o->scalars)
    return fail(AIR_ERR_ARG, "air_outputs: per-sample ELBO buffers and `scalars` are mandatory");
  return AIR_OK;
}

// One layer of a fused chain (chain_tc.cuh) from a prepared weight.
void chain_add(const air_handle* h, air::chain::Params& p, const float* params, const Layer& l, int epi, int a_src,
               int a_buf, float* out, int ldo, float* save = nullptr) {
  const TcWeight& w = h->tcw[l.tc];
  air::chain::Layer& L = p.layer[p.n_layers];
  p.tm[p.n_layers] = w.tm_chain;
  L.K = l.K;
  L.N = w.split_n > 0 ? w.N_alloc : l.N;
  L.n_box = w.n_box;
  L.n_pass = w.n_pass;
  L.lo_row = w.N_alloc;
  L.epi = epi;
  L.a_src = a_src;
  L.a_buf = a_buf;
  L.bias = h->bias_arena + w.bias_off;
  L.out = out;
  L.ldo = ldo;
  L.save = epi == air::chain::EPI_ELU_A ? save : nullptr;
  ++p.n_layers;
}
// a whole neural.MLP whose first layer reads in[a_buf] and whose last layer writes fp32 rows to `out`
// `saves` (training mode): fp32 destinations of the hidden activations, one per ELU layer in order (null: not kept)
void chain_add_mlp(const air_handle* h, air::chain::Params& p, const float* params, const Mlp& mlp, int a_buf,
                   bool first_loads, float* out, int ldo, const std::vector<float*>* saves = nullptr) {
  const int nl = (int)mlp.layers.size();
  for (int i = 0; i < nl; ++i) {
    const bool last = i == nl - 1;
    float* save = (saves && i < (int)saves->size()) ? (*saves)[i] : nullptr;
    chain_add(h, p, params, mlp.layers[i], last && out ? air::chain::EPI_F32 : air::chain::EPI_ELU_A,
              (i == 0 && first_loads) ? air::chain::A_LOAD : air::chain::A_KEEP, a_buf, last ? out : nullptr, ldo, save);
  }
}
// debug: AIR_CHAIN_TRACE=<prefix> dumps the per-group SM-clock stamps of every chain launch to <prefix>.<seq>.bin
int32_t launch_chain_traced(air_handle* h, air::chain::Params& cp, cudaStream_t st) {
  static const char* prefix = getenv("AIR_CHAIN_TRACE");
  if (!prefix) {
    AIR_CUDA(air::chain::launch_chain(cp, st));
    return AIR_OK;
  }
  const int n_cta = (cp.M + air::tc::BM - 1) / air::tc::BM;
  const size_t n = (size_t)n_cta * air::chain::TRACE_GROUPS * 8;
  if (!h->trace) AIR_CUDA(cudaMalloc(&h->trace, sizeof(long long) * 4096 * air::chain::TRACE_GROUPS * 8));
  AIR_CUDA(cudaMemsetAsync(h->trace, 0, sizeof(long long) * n, st));
  cp.trace = h->trace;
  AIR_CUDA(air::chain::launch_chain(cp, st));
  std::vector<long long> host(n);
  AIR_CUDA(cudaMemcpyAsync(host.data(), h->trace, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
  AIR_CUDA(cudaStreamSynchronize(st));
  const std::string path = std::string(prefix) + "." + std::to_string(h->trace_seq++) + ".bin";
  if (FILE* f = fopen(path.c_str(), "wb")) {
    fwrite(host.data(), sizeof(long long), n, f);
    fclose(f);
  }
  return AIR_OK;
}

// the split-K first encoder layer reads fp32 images itself: no operand planes of the image are needed
bool enc1_active(const air_handle* h) {
  static const bool off = getenv("AIR_NO_ENC1") != nullptr;
  return h->use_tc && h->enc1_ok && !off;
}

air::chain::HlIn chain_in(const Buf& b) { return air::chain::HlIn{b.hlt, b.plane_t(), b.nsl}; }

// ---- fused row kernel (row_tc.cuh) ---------------------------------------------------------------------------------
// The layer list of the row path from the configuration alone (dimensions only; no device state): where MLP, steps MLP,
// glimpse Encoder, what head, Decoder.
std::vector<air::row::LayerDesc> row_layer_dims(const air_config& c, int which) {
  using namespace air::row;
  std::vector<LayerDesc> v;
  const int na_off = round_up(c.na, 16);
  auto mlp = [&](int n_in, const int32_t* hidden, int n_hidden, int n_out, int a_src) {
    int d = n_in;
    for (int i = 0; i < n_hidden; ++i) {
      LayerDesc L;
      L.K = d;
      L.N = hidden[i];
      L.epi = T_ELU;
      L.a_src = (i == 0) ? a_src : 0;
      v.push_back(L);
      d = hidden[i];
    }
    if (n_out > 0) {
      LayerDesc L;
      L.K = d;
      L.N = n_out;
      L.epi = T_OUT;
      L.a_src = (n_hidden == 0) ? a_src : 0;
      v.push_back(L);
    }
  };
  if (which == 0) {   // heads: h_t -> where MLP -> where code ; h_t -> steps MLP -> logit
    mlp(c.nh, c.where_hidden, c.n_where_hidden, 8, 1);
    v.back().out_kind = OUT_M_SMEM;
    v.back().where_after = true;
    mlp(c.nh, c.steps_hidden, c.n_steps_hidden, 1, 1);
  } else {            // glimpse VAE: crop -> glimpse Encoder -> what head -> Decoder
    mlp(c.h * c.w, c.glenc_hidden, c.n_glenc_hidden, 0, 1);
    LayerDesc L;
    L.K = c.glenc_hidden[c.n_glenc_hidden - 1];
    L.N = 2 * na_off;
    L.epi = T_WHAT;
    v.push_back(L);
    mlp(c.na, c.dec_hidden, c.n_dec_hidden, c.h * c.w, 0);
  }
  return v;
}

// "" when the row kernel covers the configuration and its schedule replays cleanly on the host
std::string row_schedule_check(const air_config& c) {
  if (2 * round_up(c.na, 16) > 128) return "what head wider than 128 columns";
  if (c.nh > 256) return "hidden state wider than the 256-K operand";
  for (const int32_t* hv : {c.where_hidden, c.steps_hidden, c.glenc_hidden, c.dec_hidden})
    for (int i = 0; i < AIR_MAX_HIDDEN; ++i)
      if (hv[i] > 256) return "a hidden layer is wider than 256";
  for (int which = 0; which < 2; ++which) {
    std::vector<air::row::LayerDesc> layers = row_layer_dims(c, which);
    if ((int)layers.size() > air::row::MAXTM) return "too many layers";
    air::row::Schedule sch;
    if (!air::row::build_schedule(layers, sch)) return sch.error;
    if (getenv("AIR_ROW_DUMP") && getenv("AIR_ROW_DUMP")[0]) {   // debug: the three programs, one line per unit / task
      for (size_t i = 0; i < sch.units.size(); ++i) {
        const air::row::Unit& u = sch.units[i];
        fprintf(stderr, "unit %3zu  layer %2d  n_row %3d  kb0 %2d  nkb %d  nsl %d  N %3d  A%d D%d  fill %2d use %2d %s%s%s%s%s\n", i,
                u.tm, u.n_row, u.kb0, u.nkb, u.nsl, u.n16 * 16, u.a_half, u.d_idx, u.fill, u.use,
                (u.flags & air::row::U_WAIT_A) ? " WAIT_A" : "", (u.flags & air::row::U_WAIT_D) ? " WAIT_D" : "",
                (u.flags & air::row::U_ACC0) ? " ACC0" : "", (u.flags & air::row::U_COMMIT_D) ? " COMMIT_D" : "",
                (u.flags & air::row::U_COMMIT_A) ? " COMMIT_A" : "");
      }
      static const char* tn[] = {"LOAD_HL", "LOAD_CROP", "ELU", "OUT", "WHAT", "WHERE"};
      for (size_t i = 0; i < sch.tasks.size(); ++i) {
        const air::row::Task& t = sch.tasks[i];
        fprintf(stderr, "task %3zu  %-9s  D%d use %2d  A%d fill %2d  s0 %3d nsl %d n_valid %3d\n", i, tn[t.type], t.d_idx, t.use,
                t.a_half, t.fill, t.s0, t.nsl, t.n_valid);
      }
    }
    const std::string r = air::row::simulate(sch);
    if (!r.empty()) return r;
  }
  return "";
}

// which = 0: heads (where MLP + where sampling, steps MLP); which = 1: glimpse VAE (crop -> what -> decoded glimpse)
int32_t launch_row_path(air_handle* h, int which, const float* eps_where, const float* eps_what, const air_outputs* o,
                        int T_run, cudaStream_t st) {
  using namespace air::row;
  const air_config& c = h->cfg;
  std::vector<LayerDesc> layers = row_layer_dims(c, which);
  // bind weights / biases / outputs, in the order row_layer_dims() lists the layers
  std::vector<const Layer*> src;
  if (which == 0) {
    for (const Mlp* m : {&h->where_mlp, &h->steps_mlp})
      for (const Layer& l : m->layers) src.push_back(&l);
  } else {
    for (const Layer& l : h->glenc.layers) src.push_back(&l);
    src.push_back(&h->what_chain);
    for (const Layer& l : h->dec.layers) src.push_back(&l);
  }
  if (src.size() != layers.size()) return fail(AIR_ERR_ARG, "internal: row layer list mismatch");
  static thread_local air::row::Params p;   // 17 KB: kept off the stack
  memset(&p, 0, sizeof(p));
  for (size_t i = 0; i < layers.size(); ++i) {
    const TcWeight& w = h->tcw[src[i]->tc];
    layers[i].tm = (int)i;
    layers[i].lo_row = w.N_alloc;
    layers[i].box_rows = air::row::row_box_rows(w.N_alloc);
    layers[i].bias = h->bias_arena + w.bias_off;
    p.tm[i] = w.tm_row;
  }
  if (which == 0) {
    layers.back().out = h->logit;
    layers.back().ldo = 1;
  } else {
    layers.back().out = o->glimpse;
    layers.back().ldo = h->G;
  }
  Schedule sch;
  if (!build_schedule(layers, sch)) return fail(AIR_ERR_ARG, "internal: row schedule: " + sch.error);
  p.n_units = (int)sch.units.size();
  p.n_tasks = (int)sch.tasks.size();
  p.n_tm = (int)layers.size();
  memcpy(p.unit, sch.units.data(), sizeof(Unit) * sch.units.size());
  memcpy(p.task, sch.tasks.data(), sizeof(Task) * sch.tasks.size());
  p.M = T_run * c.B;
  p.B = c.B;
  p.in[0] = chain_in(which == 0 ? h->hs : h->crop);
  p.eps_where = eps_where;
  p.where = o->where;
  p.where_loc = o->where_loc;
  p.where_scale = o->where_scale;
  p.max_crop = c.max_crop_size;
  p.scale_bias = c.scale_bias;
  p.eps_what = eps_what;
  p.what = o->what;
  p.what_loc = o->what_loc;
  p.what_scale = o->what_scale;
  p.na = c.na;
  p.na_off = h->na_off;
  p.what_offset = c.what_scale_offset;
  p.range_flag = h->range_flag;
  {   // the noise prefetch parks in the staging tile until the what head: legal iff no T_OUT task runs before it
    bool seen_what = false, out_before = false;
    for (const Task& t : sch.tasks) {
      if (t.type == T_WHAT) seen_what = true;
      if (t.type == T_OUT && !seen_what) out_before = true;
    }
    p.prefetch_eps = (seen_what && !out_before) ? 1 : 0;
  }
  // glimpse VAE launch: the decoder's output tasks are the last users of the staging tiles, so they may store through the
  // TMA engine (the map is cached per output pointer)
  static const bool no_out_tma = getenv("AIR_ROW_NO_OUT_TMA") != nullptr;
  if (which == 1 && !no_out_tma) {
    bool out_last = true, seen_out = false;
    for (const Task& t : sch.tasks) {
      if (t.type == T_OUT && t.out_kind == OUT_GLOBAL) seen_out = true;
      else if (seen_out && t.type != T_LOAD_HL && t.type != T_ELU) out_last = false;   // T_ELU does not touch the tiles
    }
    if (out_last && seen_out) {
      const auto key = std::make_tuple((const void*)o->glimpse, (int)h->G, (long long)p.M, 777);
      auto it = h->tmap_cache2.find(key);
      if (it == h->tmap_cache2.end()) {
        CUtensorMap tm;
        if (air::row::make_row_out_tmap(&tm, o->glimpse, h->G, p.M)) it = h->tmap_cache2.emplace(key, tm).first;
      }
      if (it != h->tmap_cache2.end()) {
        p.tm_out = it->second;
        p.out_tma = 1;
      }
    }
  }
  // debug: AIR_ROW_TRACE=<prefix> dumps the per-unit / per-task SM-clock stamps of every launch to <prefix>.<seq>.bin
  static const char* trace_prefix = getenv("AIR_ROW_TRACE");
  if (trace_prefix) {
    const size_t n = (size_t)((p.M + air::tc::BM - 1) / air::tc::BM) * (MAXU + MAXTASK) * 4;
    if (!h->trace) AIR_CUDA(cudaMalloc(&h->trace, sizeof(long long) * 1024 * (MAXU + MAXTASK) * 4));
    if (n > (size_t)1024 * (MAXU + MAXTASK) * 4) return fail(AIR_ERR_ARG, "AIR_ROW_TRACE: too many tiles");
    AIR_CUDA(cudaMemsetAsync(h->trace, 0, sizeof(long long) * n, st));
    p.trace = h->trace;
    AIR_CUDA(launch_row(p, st));
    std::vector<long long> host(n);
    AIR_CUDA(cudaMemcpyAsync(host.data(), h->trace, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    AIR_CUDA(cudaStreamSynchronize(st));
    const std::string path = std::string(trace_prefix) + "." + std::to_string(h->trace_seq++) + ".bin";
    if (FILE* f = fopen(path.c_str(), "wb")) {
      const int hdr[4] = {p.n_units, p.n_tasks, MAXU, MAXTASK};
      fwrite(hdr, sizeof(int), 4, f);
      fwrite(p.unit, sizeof(Unit), p.n_units, f);
      for (int i = 0; i < p.n_tasks; ++i) fwrite(&p.task[i].type, 1, 1, f);
      fwrite(host.data(), sizeof(long long), n, f);
      fclose(f);
    }
    ++h->launches;
    return AIR_OK;
  }
  AIR_CUDA(launch_row(p, st));
  ++h->launches;
  return AIR_OK;
}

// The shared body of air_forward / air_cell_step: T_run steps starting from explicit or initial state.
int32_t forward_impl(air_handle* h, const float* params, const float* img, const float* eps_where,
                     const float* eps_what, const float* u_pres, const float* baseline, const air_prior* prior,
                     const air_outputs* o, int T_run, const float* h_in, const float* c_in, const float* presence_in,
                     const float* canvas_in, float* canvas_step_out, float mult, cudaStream_t st,
                     bool x_hl_ready = false) {
  const air_config& c = h->cfg;
  const int B = c.B, nh = c.nh, P = h->P, G = h->G, na = c.na;
  const int TB = T_run * B;
  const int thr = 256;
  const bool tc = h->use_tc;
  const air::HlOut no_hl{nullptr, 0, 0, 0};
  const air::PdlScope pdl_scope(h->launch_overlap);
  int32_t rc;
  // training mode (air_train_enable, fp32 engine): every activation the backward pass needs is kept
  const bool train = h->train && prior != nullptr && T_run == c.T && !h_in && !canvas_in;
  // training mode: the fused chains / the cluster LSTM additionally write the activations they otherwise keep in TMEM /
  // registers as fp32 rows (Layer::save, gates_save / c_save); AIR_TRAIN_LAYERWISE=1 selects the layer-by-layer path
  static const bool layerwise = getenv("AIR_TRAIN_LAYERWISE") != nullptr;
  const bool chain = tc && h->chain_ok && !(train && layerwise);
  h->fwd_saved = false;

  // 0. tensor-core engine: (re)build the fp16-split W^T arena from the current parameters and split the images
  mark(h, AIR_ST_ENCODER, st);
  Buf x = h->x;
  x.f32 = const_cast<float*>(img);
  x.ld = P;
  if (tc && !(h->cache_weights && h->weights_ready == params)) {
    AIR_CUDA(air::launch_k(air::tc::prep_weights_kernel, dim3(h->prep_tiles), dim3(256), 0, st, params, h->arena,
                           h->prep_table, (int)h->tcw.size(), h->range_flag, h->bias_arena));
    ++h->launches;
    h->weights_ready = params;
  }
  // h0 @ W_h of the cluster LSTM (only passes that broadcast the trainable initial state read it: not air_cell_step with
  // explicit state)
  if (tc && h->lstm_ok && h->hw0 && !h_in && !(h->cache_weights && h->hw0_ready == params)) {
    AIR_CUDA(air::launch_k(air::lstm::lstm_h0w_kernel, dim3((4 * nh + 31) / 32), dim3(256), 0, st,
                           params + h->lstm_h.w_off, params + h->lstm_h0, h->hw0, nh));
    ++h->launches;
    h->hw0_ready = params;
  }
  const bool enc1 = enc1_active(h);
  if (tc && !enc1) {
    if (!x_hl_ready) {
      const size_t n4 = (size_t)B * ((P + 3) / 4);
      AIR_CUDA(air::launch_k(air::tc::split_rows_kernel, dim3((unsigned)((n4 + thr - 1) / thr)), dim3(thr), 0, st, img,
                             P, x.hl, x.plane(), x.kpad, B, P, h->range_flag));
      ++h->launches;
    }
  }

  // 1. e = Encoder(img)   (modules.py:72-76; step-invariant, cell.py:125)
  const bool lstm_fused = tc && h->lstm_ok && !(train && layerwise);
  // the cluster LSTM reads e as its A operand, thread <-> row: from row-major planes every lane touches its own sector
  // (measured: 19 k of the kernel's 90 k clocks); the last encoder layer therefore writes slice-major tiles for it
  const bool e_tiled = lstm_fused && h->e.hlt != nullptr && h->enc.layers.size() > 1;
  // inference with a two-layer input Encoder: the cluster LSTM can compute the second layer itself (lstm_tc.cuh, fuse_e2); the
  // first layer's kernel then writes its activation straight into the LSTM's tiled operand buffer.  Opt-in
  // (AIR_LSTM_FUSE_E2=1): measured at B = 4096 it shortens ONE pass (0.2679 -> 0.2652 ms: encoder 40.9 -> 32.2 us, LSTM
  // 54.0 -> 60.6 us) but slows three batches in flight (0.223 -> 0.243 ms per batch) -- the small separate GEMM co-runs with
  // the neighbouring batches' kernels, a longer 4-CTA-cluster kernel that needs whole SMs does not.
  static const bool no_fuse_e2 = getenv("AIR_LSTM_FUSE_E2") == nullptr;
  const bool fuse_e2 = e_tiled && !train && !no_fuse_e2 && enc1_active(h) && h->enc.layers.size() == 2 && h->enc.n_hidden == 2 &&
                       h->enc.layers[0].N == h->n_enc && h->n_enc == air::lstm::NH &&
                       h->tcw[h->enc.layers[1].tc].n_box == air::lstm::NH;
  Buf e_out = h->e;
  e_out.hl_tiled = e_tiled;
  if (enc1) {
    // first layer: split-K cluster kernel straight from the fp32 image (enc_tc.cuh); the remaining layers as before
    const Layer& l0 = h->enc.layers[0];
    const TcWeight& w0 = h->tcw[l0.tc];
    const bool only = h->enc.layers.size() == 1 || fuse_e2;
    Buf dst = only ? h->e : h->ping;
    dst.kpad = round_up(l0.N, air::tc::BK);
    air::enc::Params ep;
    memset(&ep, 0, sizeof(ep));
    ep.tm_w = w0.tm_chain;
    ep.img = img;
    ep.bias = params + l0.b_off;
    ep.B = B;
    ep.P = P;
    ep.nkb_total = w0.Kpad / air::tc::BK;
    ep.nkb_per_cta = (ep.nkb_total + air::enc::KSPLIT - 1) / air::enc::KSPLIT;
    ep.w_lo_row = w0.N_alloc;
    const bool want_hl = true;
    const bool want_f32 = train;
    ep.out_hl = want_hl ? dst.hl : nullptr;
    ep.hl_plane = dst.plane();
    ep.ld_hl = dst.kpad;
    if (fuse_e2) {
      ep.out_hl = h->e.hlt;
      ep.hl_plane = h->e.plane_t();
      ep.hl_nsl = h->e.nsl;
    }
    ep.out_f32 = want_f32 ? (only ? h->e.f32 : h->sv_enc[0]) : nullptr;
    ep.range_flag = h->range_flag;
    AIR_CUDA(air::enc::launch_enc1(ep, st));
    ++h->launches;
    if (!only) {
      Mlp rest;
      rest.layers.assign(h->enc.layers.begin() + 1, h->enc.layers.end());
      rest.n_hidden = h->enc.n_hidden - 1;
      std::vector<float*> rest_saves;
      if (train) rest_saves.assign(h->sv_enc.begin() + 1, h->sv_enc.end());
      if ((rc = run_mlp_from(h, params, rest, dst, B, e_out, !tc || train, tc, st,
                             train ? &rest_saves : nullptr, /*first_to_ping=*/false)) != AIR_OK)
        return rc;
    }
  } else if ((rc = run_mlp(h, params, h->enc, x, B, e_out, !tc || train, tc, st,
                           train ? &h->sv_enc : nullptr)) != AIR_OK)
    return rc;
  mark(h, AIR_ST_LSTM, st);

  // 2. gx = e @ W[:n_enc] + b   (input half of snt.LSTM's [x,h] @ W + b)
  Buf gx;
  gx.f32 = h->gx;
  gx.ld = 4 * nh;
  if (!lstm_fused &&
      (rc = dense(h, params, h->e, 0, h->lstm_x, true, nullptr, 0, gx, true, false, B, air::ACT_NONE, st)) != AIR_OK)
    return rc;

  // 3. recurrence: gates = gx + h_{t-1} @ W[n_enc:]; (c, h_t) pointwise
  if (h_in) {
    AIR_CUDA(cudaMemcpyAsync(h->h_init.f32, h_in, sizeof(float) * B * nh, cudaMemcpyDeviceToDevice, st));
    AIR_CUDA(cudaMemcpyAsync(h->cbuf, c_in, sizeof(float) * B * nh, cudaMemcpyDeviceToDevice, st));
    if (tc && !lstm_fused) {
      AIR_CUDA(air::launch_k(air::split_state_kernel, dim3((B * nh + thr - 1) / thr), dim3(thr), 0, st, h_in, B, nh,
                             h->h_init.hl_out()));
      ++h->launches;
    }
  } else if (!lstm_fused || train) {   // (the cluster LSTM kernel broadcasts the trainable initial state itself)
    AIR_CUDA(air::launch_k(air::lstm_init_state_kernel, dim3((B * nh + thr - 1) / thr), dim3(thr), 0, st,
                           params + h->lstm_h0, params + h->lstm_c0, h->h_init.f32, train ? h->c_all : h->cbuf, B, nh,
                           (tc && !lstm_fused) ? h->h_init.hl_out() : no_hl));
    ++h->launches;
  }
  Buf gates;
  gates.f32 = h->gates;
  gates.ld = 4 * nh;
  if (lstm_fused) {
    // gx + all T recurrent steps in one cluster launch (lstm_tc.cuh)
    air::lstm::Params lp;
    memset(&lp, 0, sizeof(lp));
    lp.tm_x = h->tcw[h->lstm_x_perm].tm_chain;
    lp.tm_h = h->tcw[h->lstm_h_perm].tm_chain;
    lp.bias = h->bias_arena + h->tcw[h->lstm_x_perm].bias_off;
    lp.e = h->e.f32;
    lp.e_hl = e_tiled ? h->e.hlt : h->e.hl;   // the encoder's last layer wrote hl planes (tensor-core engine)
    lp.e_plane = e_tiled ? h->e.plane_t() : h->e.plane();
    lp.e_ld = h->e.kpad;
    lp.e_nsl = e_tiled ? h->e.nsl : 0;
    if (fuse_e2) {
      const TcWeight& w2 = h->tcw[h->enc.layers[1].tc];
      lp.fuse_e2 = 1;
      lp.tm_e2 = w2.tm_chain;
      lp.e2_lo_row = w2.N_alloc;
      lp.n_e1 = h->enc.layers[0].N;
      lp.bias_e2 = h->bias_arena + w2.bias_off;
    }
    lp.hs_last_only = train ? 0 : 1;
    lp.n_enc = h->n_enc;
    // explicit per-canvas state (air_cell_step): rows.  The training forward broadcasts (h0, c0) like inference and folds
    // step 1's recurrent product into the constant hw0; the tiled copies the backward pass reads (hprev slice 0, c_all
    // slice 0) are written by lstm_init_state_kernel, which the LSTM kernel no longer depends on.
    static const bool train_rows = getenv("AIR_LSTM_TRAIN_ROWS") != nullptr;   // (the previous behaviour, for comparison)
    const bool rows = h_in || (train && train_rows);
    lp.h_init = rows ? h->h_init.f32 : params + h->lstm_h0;   // cell.py:103: (h0, c0) [1,nh] tiled to the batch
    lp.h_init_ld = rows ? nh : 0;
    static const bool no_fold = getenv("AIR_LSTM_NO_FOLD") != nullptr;
    lp.hw0 = (rows || no_fold) ? nullptr : h->hw0;   // broadcast initial state: step 1 needs no recurrent GEMM
    lp.c_in = rows ? (train ? h->c_all : h->cbuf) : params + h->lstm_c0;
    lp.c_in_ld = rows ? nh : 0;
    lp.c = h->cbuf;
    lp.gates_save = train ? h->gates_all : nullptr;
    lp.c_save = train ? h->c_all + (size_t)B * nh : nullptr;
    lp.hs = h->hs.f32;
    lp.hs_hlt = h->hs.hlt_out();
    lp.gx_scr = h->gx_scr;
    lp.hx = h->hx;
    lp.hx_plane = (size_t)round_up(B, air::tc::BM) * nh;
    lp.B = B;
    lp.T = T_run;
    lp.forget_bias = c.forget_bias;
    lp.range_flag = h->range_flag;
    static const char* lstm_trace = getenv("AIR_LSTM_TRACE");
    if (lstm_trace) {   // debug: per-phase SM-clock stamps of every CTA -> <AIR_LSTM_TRACE>.<seq>.bin
      const size_t n = (size_t)((B + air::tc::BM - 1) / air::tc::BM) * air::lstm::CLUSTER * 64;
      if (!h->trace) AIR_CUDA(cudaMalloc(&h->trace, sizeof(long long) * 1024 * (air::row::MAXU + air::row::MAXTASK) * 4));
      AIR_CUDA(cudaMemsetAsync(h->trace, 0, sizeof(long long) * n, st));
      lp.trace = h->trace;
      AIR_CUDA(air::lstm::launch_lstm(lp, st));
      std::vector<long long> host(n);
      AIR_CUDA(cudaMemcpyAsync(host.data(), h->trace, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
      AIR_CUDA(cudaStreamSynchronize(st));
      const std::string path = std::string(lstm_trace) + "." + std::to_string(h->trace_seq++) + ".bin";
      if (FILE* f = fopen(path.c_str(), "wb")) {
        fwrite(host.data(), sizeof(long long), n, f);
        fclose(f);
      }
    } else {
      AIR_CUDA(air::lstm::launch_lstm(lp, st));
    }
    ++h->launches;
  }
  for (int t = 0; t < (lstm_fused ? 0 : T_run); ++t) {
    const Buf& h_prev = (t == 0) ? h->h_init : h->hs;
    const int row0 = (t == 0) ? 0 : (t - 1) * B;
    if (train) gates.f32 = h->gates_all + (size_t)t * B * 4 * nh;
    if ((rc = dense(h, params, h_prev, row0, h->lstm_h, false, h->gx, 4 * nh, gates, true, false, B, air::ACT_NONE,
                    st)) != AIR_OK)
      return rc;
    const float* c_src = train ? h->c_all + (size_t)t * B * nh : h->cbuf;
    float* c_dst = train ? h->c_all + (size_t)(t + 1) * B * nh : h->cbuf;
    air::HlOut hs_hl = no_hl;
    if (tc) {
      hs_hl = h->hs.hl_out();
      hs_hl.p += (size_t)t * B * h->hs.kpad;
    }
    AIR_CUDA(air::launch_k(air::lstm_pointwise_kernel, dim3((B * nh + thr - 1) / thr), dim3(thr), 0, st,
                           (const float*)gates.f32, c_src, c_dst, h->hs.f32 + (size_t)t * B * nh, B, nh, c.forget_bias, hs_hl,
                           chain ? h->hs.hlt_out() : no_hl, (size_t)t * B));
    ++h->launches;
  }
  if (o->final_h)
    AIR_CUDA(cudaMemcpyAsync(o->final_h, h->hs.f32 + (size_t)(T_run - 1) * B * nh, sizeof(float) * B * nh,
                             cudaMemcpyDeviceToDevice, st));
  if (o->final_c)
    AIR_CUDA(cudaMemcpyAsync(o->final_c, train ? h->c_all + (size_t)T_run * B * nh : h->cbuf, sizeof(float) * B * nh,
                             cudaMemcpyDeviceToDevice, st));
  if (train) {   // h_{t-1} of every step, stacked: the A operand of dW_h = h_prev^T @ dgates
    AIR_CUDA(cudaMemcpyAsync(h->hprev, h->h_init.f32, sizeof(float) * B * nh, cudaMemcpyDeviceToDevice, st));
    if (T_run > 1)
      AIR_CUDA(cudaMemcpyAsync(h->hprev + (size_t)B * nh, h->hs.f32, sizeof(float) * (size_t)(T_run - 1) * B * nh,
                               cudaMemcpyDeviceToDevice, st));
  }

  // 4.-7. row kernel (row_tc.cuh), two launches around the glimpse read: heads + where sampling, then the glimpse VAE
  static const bool no_row = getenv("AIR_NO_ROW") != nullptr;
  const bool rowk = chain && h->row_ok && !train && !no_row;
  // The heads keep chain_kernel by default: the where code is the one quantity the canvas amplifies (1 / s_x in the
  // inverse transformer), and chain_kernel's sweep order -- every cross term of a layer before any main term -- leaves
  // about half the accumulator-truncation bias of the row kernel's per-unit sweeps (measured on the ill-conditioned
  // canvases of test_forward_tc_trained_like_weights); the heads are 4 us of the pass.  AIR_ROW_HEADS=1 selects the row kernel.
  static const bool row_heads_env = getenv("AIR_ROW_HEADS") != nullptr;
  const bool row_heads = rowk && row_heads_env;
  // 4. heads over all T*B hidden states at once
  mark(h, AIR_ST_WHERE_MLP, st);
  Buf m;
  m.f32 = h->m;
  m.ld = 8;
  Buf logit;
  logit.f32 = h->logit;
  logit.ld = 1;
  if (row_heads) {
    if ((rc = launch_row_path(h, 0, eps_where, eps_what, o, T_run, st)) != AIR_OK) return rc;
    mark(h, AIR_ST_STEPS, st);
  } else if (chain) {
    // both heads of every (t, canvas) row in ONE launch: hs -> where MLP -> m ; hs -> steps MLP -> logit
    air::chain::Params cp;
    memset(&cp, 0, sizeof(cp));
    cp.M = TB;
    cp.in[0] = chain_in(h->hs);
    cp.range_flag = h->range_flag;
    chain_add_mlp(h, cp, params, h->where_mlp, 0, true, h->m, 8, train ? &h->sv_where : nullptr);       // modules.py:58-63
    chain_add_mlp(h, cp, params, h->steps_mlp, 0, true, h->logit, 1, train ? &h->sv_steps : nullptr);   // modules.py:119-122
    if ((rc = launch_chain_traced(h, cp, st)) != AIR_OK) return rc;
    ++h->launches;
    mark(h, AIR_ST_STEPS, st);
  } else {
    if ((rc = run_mlp(h, params, h->where_mlp, h->hs, TB, m, true, false, st, train ? &h->sv_where : nullptr)) != AIR_OK)
      return rc;   // modules.py:58-63
    mark(h, AIR_ST_STEPS, st);
    if ((rc = run_mlp(h, params, h->steps_mlp, h->hs, TB, logit, true, false, st, train ? &h->sv_steps : nullptr)) !=
        AIR_OK)
      return rc;   // modules.py:119-122
  }
  // the presence scan (cell.py:137-151) rides in the glimpse-read kernel: one thread of each canvas's CTA
  air::PresenceArgs pa;
  pa.logit = h->logit;
  pa.u_pres = u_pres;
  pa.presence_in = presence_in;
  pa.presence_prob = o->presence_prob;
  pa.presence = o->presence;
  pa.step_bias = c.step_bias;
  pa.explore_eps = c.explore_eps;
  pa.discrete = c.discrete_steps;
  mark(h, AIR_ST_READ, st);

  // 5. where sampling + glimpse read   (cell.py:129-135); the heads launch of the row kernel has sampled `where` already
  static const int read_threads = getenv("AIR_READ_THREADS") ? atoi(getenv("AIR_READ_THREADS")) : air::WHERE_READ_THREADS;
  static const char* read_trace = getenv("AIR_READ_TRACE");   // debug: 8 stamps per CTA to <prefix>.<seq>.bin
  long long* rtr = nullptr;
  if (read_trace) {
    AIR_CUDA(cudaMalloc(&rtr, sizeof(long long) * 8 * (size_t)B));
    AIR_CUDA(cudaMemsetAsync(rtr, 0, sizeof(long long) * 8 * (size_t)B, st));
  }
  {
    const float* m_arg = row_heads ? (const float*)nullptr : (const float*)h->m;
    float* crop_arg = (tc && !train) ? nullptr : h->crop.f32;
    const air::HlOut hl_arg = tc ? (chain ? h->crop.hlt_out() : h->crop.hl_out()) : no_hl;
    const double sw = c.w > 1 ? 2.0 / (double)(c.w - 1) : 0.0, sh = c.h > 1 ? 2.0 / (double)(c.h - 1) : 0.0;
    static const bool no_fixed = getenv("AIR_READ_NO_FIXED") != nullptr;
    // the quoted configuration (three steps, 50x50 image, 20x20 glimpse, 128 threads) has its own instantiation
    if (!no_fixed && T_run == 3 && c.H == 50 && c.W == 50 && c.h == 20 && c.w == 20 && read_threads == 128)
      AIR_CUDA(air::launch_k(air::where_read_kernel<3, 50, 50, 20, 20, 128>, dim3(B), dim3(read_threads),
                             air::where_read_smem(T_run, c.H, c.W, c.h, c.w), st, m_arg, eps_where, img, o->where,
                             o->where_loc, o->where_scale, crop_arg, hl_arg, T_run, B, c.H, c.W, c.h, c.w,
                             c.max_crop_size, c.scale_bias, sw, sh, pa, rtr));
    else if (!no_fixed && T_run == 5 && c.H == 100 && c.W == 100 && c.h == 28 && c.w == 28 && read_threads == 128)   // configs[3]
      AIR_CUDA(air::launch_k(air::where_read_kernel<5, 100, 100, 28, 28, 128>, dim3(B), dim3(read_threads),
                             air::where_read_smem(T_run, c.H, c.W, c.h, c.w), st, m_arg, eps_where, img, o->where,
                             o->where_loc, o->where_scale, crop_arg, hl_arg, T_run, B, c.H, c.W, c.h, c.w,
                             c.max_crop_size, c.scale_bias, sw, sh, pa, rtr));
    else
      AIR_CUDA(air::launch_k(air::where_read_kernel<0, 0, 0, 0, 0, 0>, dim3(B), dim3(read_threads),
                             air::where_read_smem(T_run, c.H, c.W, c.h, c.w), st, m_arg, eps_where, img, o->where,
                             o->where_loc, o->where_scale, crop_arg, hl_arg, T_run, B, c.H, c.W, c.h, c.w,
                             c.max_crop_size, c.scale_bias, sw, sh, pa, rtr));
  }
  if (read_trace) {
    std::vector<long long> host((size_t)B * 8);
    AIR_CUDA(cudaMemcpyAsync(host.data(), rtr, sizeof(long long) * host.size(), cudaMemcpyDeviceToHost, st));
    AIR_CUDA(cudaStreamSynchronize(st));
    cudaFree(rtr);
    const std::string path = std::string(read_trace) + "." + std::to_string(h->trace_seq++) + ".bin";
    if (FILE* f = fopen(path.c_str(), "wb")) {
      const int hdr[4] = {B, 0, 8, 0};
      fwrite(hdr, sizeof(int), 4, f);
      fwrite(host.data(), sizeof(long long), host.size(), f);
      fclose(f);
    }
  }
  ++h->launches;
  mark(h, AIR_ST_GLIMPSE_ENC, st);

  // 6. glimpse encoder -> what   (cell.py:153-156)   7. decoder   (cell.py:158)
  if (rowk) {
    if ((rc = launch_row_path(h, 1, eps_where, eps_what, o, T_run, st)) != AIR_OK) return rc;
    mark(h, AIR_ST_DECODER, st);
  } else if (chain) {
    // crop -> glimpse Encoder -> what head (sample) -> Decoder -> glimpse, one launch, activations resident in TMEM
    air::chain::Params cp;
    memset(&cp, 0, sizeof(cp));
    cp.M = TB;
    cp.in[0] = chain_in(h->crop);
    cp.range_flag = h->range_flag;
    cp.eps_what = eps_what;
    cp.what = o->what;
    cp.what_loc = o->what_loc;
    cp.what_scale = o->what_scale;
    cp.na = na;
    cp.na_off = h->na_off;
    cp.what_offset = c.what_scale_offset;
    std::vector<float*> glenc_saves;   // every glimpse-encoder layer is an ELU layer; the last one feeds the what head
    if (train) {
      glenc_saves = h->sv_glenc;
      glenc_saves.push_back(h->sv_q);
    }
    chain_add_mlp(h, cp, params, h->glenc, 0, true, nullptr, 0, train ? &glenc_saves : nullptr);
    chain_add(h, cp, params, h->what_chain, air::chain::EPI_WHAT, air::chain::A_KEEP, 0, nullptr, 0);
    {
      Mlp dec = h->dec;
      dec.layers[0].K = na;
      chain_add_mlp(h, cp, params, dec, 0, false, o->glimpse, G, train ? &h->sv_dec : nullptr);
    }
    if ((rc = launch_chain_traced(h, cp, st)) != AIR_OK) return rc;
    ++h->launches;
    mark(h, AIR_ST_DECODER, st);
  } else {
    {
      const Layer& last = h->glenc.layers.back();
      Buf q = (h->glenc.layers.size() & 1) ? h->ping : h->pong;   // where the chain's last layer may land
      if (train) q.f32 = h->sv_q;
      q.ld = last.N;
      q.kpad = round_up(last.N, air::tc::BK);
      if ((rc = run_mlp(h, params, h->glenc, h->crop, TB, q, !tc || train, tc, st, train ? &h->sv_glenc : nullptr)) != AIR_OK)
        return rc;
      Buf r;
      r.f32 = h->r;
      r.ld = 2 * na;
      if ((rc = dense(h, params, q, 0, h->what_lin, true, nullptr, 0, r, true, false, TB, air::ACT_NONE, st)) != AIR_OK)
        return rc;
      const size_t n = (size_t)TB * na;
      AIR_CUDA(air::launch_k(air::what_kernel, dim3((unsigned)((n + thr - 1) / thr)), dim3(thr), 0, st, h->r, eps_what,
                             o->what, o->what_loc, o->what_scale, (size_t)TB, na, c.what_scale_offset,
                             tc ? h->what_in.hl_out() : no_hl));
      ++h->launches;
    }
    mark(h, AIR_ST_DECODER, st);
    {
      Buf what = h->what_in;
      what.f32 = o->what;
      what.ld = na;
      Buf glimpse;
      glimpse.f32 = o->glimpse;
      glimpse.ld = G;
      if ((rc = run_mlp(h, params, h->dec, what, TB, glimpse, true, false, st, train ? &h->sv_dec : nullptr)) != AIR_OK)
        return rc;
    }
  }
  mark(h, AIR_ST_PAINT_ELBO, st);

  // 8. paint + ELBO   (cell.py:159-165, model.py:89-104,126-251,319-343)
  air::ElboArgs a;
  memset(&a, 0, sizeof(a));
  a.img = img;
  a.glimpse = o->glimpse;
  a.where = o->where;
  a.where_loc = o->where_loc;
  a.where_scale = o->where_scale;
  a.what_loc = o->what_loc;
  a.what_scale = o->what_scale;
  a.presence = o->presence;
  a.presence_prob = o->presence_prob;
  a.canvas_in = canvas_in;
  a.canvas = canvas_step_out ? canvas_step_out : o->canvas;
  a.glimpse_viz = o->glimpse_viz;
  a.num_steps_posterior = o->num_steps_posterior;
  a.num_step_per_sample = o->num_step_per_sample;
  a.prior_step_weight = o->prior_step_weight;
  a.rec_loss_per_sample = o->rec_loss_per_sample;
  a.kl_num_steps_per_sample = o->kl_num_steps_per_sample;
  a.kl_what_per_sample = o->kl_what_per_sample;
  a.kl_where_per_sample = o->kl_where_per_sample;
  a.loss_per_sample = o->loss_per_sample;
  a.num_steps_log_prob = o->num_steps_log_prob;
  a.T = T_run; a.B = B; a.H = c.H; a.W = c.W; a.h = c.h; a.w = c.w; a.na = na;
  a.output_std = c.output_std;
  a.output_multiplier = mult;
  a.do_elbo = prior ? 1 : 0;
  if (prior) {
    a.prior = *prior;
    // geometric_prior(success_prob, T) (prior.py:26-32): one table for the whole batch, float64 or float32 island
    for (int k = 0; k <= T_run; ++k)
      a.steps_prior[k] = prior->steps_prob_is_f64 ? air::geom_prior_f64(prior->steps_success_prob, k)
                                                  : (double)air::geom_prior_f32((float)prior->steps_success_prob, k);
    a.steps_prior_dev = h->prior_dev_on ? h->prior_dev : nullptr;
  }
  a.lp_const = (float)(0.5 * 1.8378770664093453 /* log(2 pi) */ + std::log((double)c.output_std));
  a.prior_part = h->prior_part;
  a.trace = nullptr;
  // debug: AIR_PAINT_TRACE=<prefix> dumps 8 stamps per CTA of the paint grid (globaltimer ns, SM id) to <prefix>.<seq>.bin
  static const char* paint_trace = getenv("AIR_PAINT_TRACE");
  if (paint_trace) {
    const size_t n = (size_t)(B + (B + 7) / 8 + 8) * 8;
    long long* dbuf = nullptr;
    AIR_CUDA(cudaMalloc(&dbuf, sizeof(long long) * n));
    AIR_CUDA(cudaMemsetAsync(dbuf, 0, sizeof(long long) * n, st));
    a.trace = dbuf;
    AIR_CUDA(air::launch_paint_elbo(a, st));
    std::vector<long long> host(n);
    AIR_CUDA(cudaMemcpyAsync(host.data(), dbuf, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    AIR_CUDA(cudaStreamSynchronize(st));
    cudaFree(dbuf);
    const std::string path = std::string(paint_trace) + "." + std::to_string(h->trace_seq++) + ".bin";
    if (FILE* f = fopen(path.c_str(), "wb")) {
      const int hdr[4] = {B, a.n_prior_ctas, 8, 0};
      fwrite(hdr, sizeof(int), 4, f);
      fwrite(host.data(), sizeof(long long), n, f);
      fclose(f);
    }
  } else {
    AIR_CUDA(air::launch_paint_elbo(a, st));   // paint CTAs + (when a prior is given) the prior-term CTAs, one grid
  }
  ++h->launches;

  if (prior) {
    AIR_CUDA(air::launch_k(air::elbo_scalars_kernel, dim3(1), dim3(1024), 0, st, o->rec_loss_per_sample,
                           o->kl_num_steps_per_sample, o->kl_what_per_sample, o->kl_where_per_sample,
                           o->num_step_per_sample, o->num_steps_log_prob, baseline, o->scalars, B, *prior,
                           (const float*)h->prior_part, o->loss_per_sample));
    ++h->launches;
  }
  mark(h, AIR_N_STAGES, st);
  h->fwd_saved = train && o->canvas != nullptr;
  return AIR_OK;
}

const char* const kStageNames[AIR_N_STAGES] = {"input_encoder", "lstm",        "where_mlp", "steps_presence",
                                               "where_read",    "glimpse_enc", "decoder",   "paint_elbo"};

// carve a buffer out of the workspace (two passes: sizing with base == nullptr, then assignment)
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(char* b) : base(b) {}
  template <class T>
  T* take(size_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(n * sizeof(T), 1024);
    return p;
  }
};

void carve_workspace(air_handle* h, Carver& cv) {
  const air_config& c = h->cfg;
  const size_t TB = (size_t)c.T * c.B, B = c.B;
  const bool tc = h->use_tc;
  const int B_alloc = round_up(c.B, air::tc::BM), TB_alloc = round_up((int)TB, air::tc::BM);
  h->prior_dev = cv.take<double>(AIR_MAX_STEPS + 1);
  auto f32 = [&](Buf& b, size_t rows, int width) {
    b.f32 = cv.take<float>(rows * width);
    b.ld = width;
  };
  auto hl = [&](Buf& b, int rows_alloc, int width) {
    b.rows_alloc = rows_alloc;
    b.kpad = round_up(width, air::tc::BK);
    b.hl = cv.take<__half>(2 * (size_t)rows_alloc * b.kpad);
  };
  // fp32 side
  if (!tc) {
    f32(h->ping, TB, h->max_width);
    f32(h->pong, TB, h->max_width);
    f32(h->e, B, h->n_enc);
    f32(h->crop, TB, h->G);
  }
  f32(h->h_init, B, c.nh);
  f32(h->hs, TB, c.nh);
  if (tc) f32(h->e, B, h->n_enc);
  h->gx = cv.take<float>(B * 4 * c.nh);
  h->gates = cv.take<float>(B * 4 * c.nh);
  h->cbuf = cv.take<float>(B * c.nh);
  h->m = cv.take<float>(TB * 8);
  h->logit = cv.take<float>(TB);
  h->r = cv.take<float>(TB * 2 * c.na);
  h->st_img = cv.take<float>(B * h->P);
  h->st_img_u8 = cv.take<uint8_t>(B * h->P);
  h->st_eps_where = cv.take<float>(TB * 4);
  h->st_eps_what = cv.take<float>(TB * c.na);
  h->st_u = cv.take<float>(TB);
  h->st_pres_in = cv.take<float>(B);
  h->prior_part = cv.take<float>(B);
  // tensor-core side
  if (tc) {
    hl(h->ping, TB_alloc, h->max_width);
    hl(h->pong, TB_alloc, h->max_width);
    hl(h->x, B_alloc, h->P);
    hl(h->e, B_alloc, h->n_enc);
    hl(h->h_init, B_alloc, c.nh);
    hl(h->hs, TB_alloc, c.nh);
    hl(h->crop, TB_alloc, h->G);
    hl(h->what_in, TB_alloc, c.na);
    auto hlt = [&](Buf& b, int width) {
      b.nsl = (width + 15) / 16;
      b.hlt = cv.take<__half>(2 * (size_t)b.rows_alloc * b.nsl * 16);
    };
    hlt(h->hs, c.nh);
    hlt(h->crop, h->G);
    if (h->lstm_ok) hlt(h->e, h->n_enc);
    if (h->lstm_ok) {
      h->hw0 = cv.take<float>((size_t)4 * c.nh);
      h->gx_scr = cv.take<float>((size_t)B_alloc * 4 * c.nh);
      h->hx = cv.take<__half>(2 * 2 * (size_t)B_alloc * c.nh);
    }
    size_t halves = 0, bias_floats = 0;
    for (TcWeight& w : h->tcw) {
      w.arena_off = (int64_t)halves;
      halves += align_up(2 * (size_t)w.N_alloc * w.Kpad, 512);
      w.bias_off = (int64_t)bias_floats;
      bias_floats += (size_t)w.N_alloc;
    }
    h->arena = cv.take<__half>(halves);
    h->bias_arena = cv.take<float>(bias_floats);
    h->prep_table = cv.take<air::tc::PrepEntry>(h->tcw.size());
    h->range_flag = cv.take<int>(1);
  }
}

// ---- training workspace (air_train_enable): saved activations + gradient scratch --------------------------------
void carve_train(air_handle* h, Carver& cv) {
  const air_config& c = h->cfg;
  const size_t TB = (size_t)c.T * c.B, B = c.B;
  auto hidden = [&](std::vector<float*>& v, const Mlp& mlp, size_t rows) {
    v.clear();
    for (size_t i = 0; i + 1 < mlp.layers.size(); ++i) v.push_back(cv.take<float>(rows * mlp.layers[i].N));
  };
  hidden(h->sv_enc, h->enc, B);
  hidden(h->sv_where, h->where_mlp, TB);
  hidden(h->sv_steps, h->steps_mlp, TB);
  hidden(h->sv_glenc, h->glenc, TB);
  hidden(h->sv_dec, h->dec, TB);
  h->sv_q = cv.take<float>(TB * h->what_lin.K);
  if (h->use_tc) {   // the tensor-core engine keeps the crops as hl planes only; the backward pass wants fp32 rows too
    h->crop.f32 = cv.take<float>(TB * h->G);
    h->crop.ld = h->G;
  }
  h->gates_all = cv.take<float>(TB * 4 * c.nh);
  h->c_all = cv.take<float>((TB + B) * c.nh);
  h->hprev = cv.take<float>(TB * c.nh);
  int wmax = h->max_width;
  for (int v : {h->G, 2 * c.na, c.nh, 8}) wmax = v > wmax ? v : wmax;
  h->g_a = cv.take<float>(TB * wmax);
  h->g_b = cv.take<float>(TB * wmax);
  h->g_glimpse = cv.take<float>(TB * h->G);
  h->g_crop = cv.take<float>(TB * h->G);
  h->g_what = cv.take<float>(TB * c.na);
  h->g_r = cv.take<float>(TB * 2 * c.na);
  h->g_wh_paint = cv.take<float>(TB * 4);
  h->g_wh_read = cv.take<float>(TB * 4);
  h->g_m = cv.take<float>(TB * 8);
  h->g_logit = cv.take<float>(TB);
  h->g_pres = cv.take<float>(TB);
  h->g_h = cv.take<float>(TB * c.nh);
  h->g_gates = cv.take<float>(TB * 4 * c.nh);
  h->g_gx = cv.take<float>(B * 4 * c.nh);
  h->g_hrec = cv.take<float>(B * c.nh);
  h->g_c = cv.take<float>(B * c.nh);
  h->g_e = cv.take<float>(B * h->n_enc);
  if (h->tc_bwd) {
    // largest transposed operands over all dense layers: X^T [round_up(K,128)][round_up(rows,64)], dY^T [round_up(N,64)][..]
    size_t xt = 0, yt = 0;
    auto visit = [&](const Layer& l, size_t rows) {
      const size_t mp = (size_t)round_up((int)rows, 64);
      xt = std::max(xt, (size_t)round_up(l.K, 128) * mp);
      yt = std::max(yt, (size_t)round_up(l.N, 64) * mp);
    };
    for (const Layer& l : h->enc.layers) visit(l, B);
    for (const Mlp* m : {&h->where_mlp, &h->steps_mlp, &h->glenc, &h->dec})
      for (const Layer& l : m->layers) visit(l, TB);
    visit(h->what_lin, TB);
    visit(h->lstm_h, TB);
    visit(h->lstm_x, B);
    if (h->bl.attached)
      for (const Layer& l : h->bl.mlp.layers) visit(l, B);
    h->hl_xt_halves = xt;
    h->hl_yt_halves = yt;
    for (int i = 0; i < h->n_side_bufs; ++i) {
      h->hl_xt[i] = cv.take<__half>(2 * xt);
      h->hl_yt[i] = cv.take<__half>(2 * yt);
    }
    h->t_range_flag = cv.take<int>(1);
    // input gradients: the widest dY (rows padded to the 128-row tile) and one [round_up(K,64)][round_up(N,64)] pair
    // of planes per weight matrix
    size_t dy = 0, wnt = 0;
    h->wnt_index.clear();
    auto visit2 = [&](const Layer& l, size_t rows) {
      dy = std::max(dy, (size_t)round_up((int)rows, 128) * round_up(l.N, 64));
      h->wnt_index[l.w_off] = std::make_pair(wnt, round_up(l.N, 64));
      wnt += 2 * (size_t)round_up(l.K, 64) * round_up(l.N, 64);
    };
    for (const Layer& l : h->enc.layers) visit2(l, B);
    for (const Mlp* m : {&h->where_mlp, &h->steps_mlp, &h->glenc, &h->dec})
      for (const Layer& l : m->layers) visit2(l, TB);
    visit2(h->what_lin, TB);
    visit2(h->lstm_h, TB);
    visit2(h->lstm_x, B);
    if (h->bl.attached)   // (only the dY plane size matters for the baseline: its input gradients run on the SIMT GEMMs)
      for (const Layer& l : h->bl.mlp.layers) dy = std::max(dy, (size_t)round_up((int)B, 128) * round_up(l.N, 64));
    h->hl_dy_halves = dy;
    h->hl_dy2[0] = cv.take<__half>(2 * dy);
    h->hl_dy2[1] = cv.take<__half>(2 * dy);
    h->wnt_arena = cv.take<__half>(wnt);
    h->wnt_table = cv.take<air::tc::RowsEntry>(h->wnt_index.size());
  }
}

// split of the batch-row contraction of a weight-gradient GEMM so that the grid fills the machine
int pick_split(int out_rows, int out_cols, int contraction) {
  const int tiles = ((out_rows + 127) / 128) * ((out_cols + (out_cols > 32 ? 63 : 15)) / (out_cols > 32 ? 64 : 16));
  int s = (2 * 148 + tiles - 1) / tiles;
  const int max_s = contraction / 64 > 1 ? contraction / 64 : 1;
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

// dW += X^T @ dY, db += colsum(dY) for one dense layer (X [M, l.K] with row pitch ldx, dY [M, l.N] with row pitch ldy)
int32_t get_tmap2(air_handle* h, const __half* base, int kpad, long long rows_total, int box_rows, const CUtensorMap** out) {
  const auto key = std::make_tuple((const void*)base, kpad, rows_total, box_rows);
  auto it = h->tmap_cache2.find(key);
  if (it == h->tmap_cache2.end()) {
    CUtensorMap tm;
    if (!air::tc::make_tmap(&tm, base, kpad, rows_total, box_rows))
      return fail(AIR_ERR_CUDA, "cuTensorMapEncodeTiled failed for a backward operand");
    it = h->tmap_cache2.emplace(key, tm).first;
  }
  *out = &it->second;
  return AIR_OK;
}

// dW += X^T @ dY on the tensor-core split engine: both operands are re-laid out contraction-major (the batch rows) by
// split_transpose_kernel as bf16 hi/lo planes (16 significant bits each; per-sample gradients span the fp32 exponent
// range -- the 1 / s_x factors of the inverse transformer -- which fp16 planes cannot hold), three tcgen05.mma per K slice
// as in the forward, the contraction split over gridDim.z with fp32 atomics into the zeroed gradient buffer.
cudaEvent_t next_event(air_handle* h) {
  cudaEvent_t e = h->ev_pool[h->ev_next];
  h->ev_next = (h->ev_next + 1) % h->ev_pool.size();
  return e;
}
// before the main stream overwrites a gradient buffer: wait until the side streams have finished reading it
int32_t wait_consumed(air_handle* h, const float* buf, cudaStream_t st) {
  auto it = h->dy_consumed.find(buf);
  if (it != h->dy_consumed.end()) {
    AIR_CUDA(cudaStreamWaitEvent(st, it->second, 0));
    h->dy_consumed.erase(it);
  }
  return AIR_OK;
}

int32_t layer_weight_grad_tc(air_handle* h, float* grad, const Layer& l, const float* X, int ldx, const float* dY, int ldy,
                             int M, bool dx_follows, cudaStream_t main_st) {
  namespace tc = air::tc;
  // fork: everything below runs on a side stream once X and dY are complete on the main stream; the main stream goes on
  // with the input gradient of this layer (its own row-major copy of dY) and the layers below
  cudaStream_t st = main_st;
  int lane = 0;                    // which side stream / operand buffer pair this layer uses
  if (h->n_side) {
    lane = h->side_next;
    h->side_next = (h->side_next + 1) % h->n_side;
    cudaEvent_t ready = next_event(h);
    AIR_CUDA(cudaEventRecord(ready, main_st));
    AIR_CUDA(cudaStreamWaitEvent(h->side[lane], ready, 0));
    st = h->side[lane];
    dx_follows = false;
  }
  __half* const hl_xt = h->hl_xt[lane];
  __half* const hl_yt = h->hl_yt[lane];
  const int mp = round_up(M, 64), KA = round_up(l.K, 128), NA = round_up(l.N, 64);
  const int np = NA, MA = round_up(M, 128);
  if ((size_t)KA * mp > h->hl_xt_halves || (size_t)NA * mp > h->hl_yt_halves || (size_t)MA * np > h->hl_dy_halves)
    return fail(AIR_ERR_ARG, "internal: transposed operand does not fit the training workspace");
  AIR_CUDA(air::launch_k(tc::split_transpose_kernel, dim3((l.K + 31) / 32, (mp + tc::ST_M - 1) / tc::ST_M), dim3(256), 0, st, X, ldx, M, l.K,
                         1.0f, hl_xt, (size_t)KA * mp, mp, h->t_range_flag, 1, (__half*)nullptr, (size_t)0, 0,
                         (float*)nullptr));
  // one read of dY: transposed planes (this GEMM), the bias gradient, and the row-major planes of the dX GEMM that follows
  AIR_CUDA(air::launch_k(tc::split_transpose_kernel, dim3((dx_follows ? np : l.N + 31) / 32, (mp + tc::ST_M - 1) / tc::ST_M), dim3(256), 0,
                         st, dY, ldy, M, l.N, 1.0f, hl_yt, (size_t)NA * mp, mp, h->t_range_flag, 1,
                         dx_follows ? h->hl_dy2[0] : (__half*)nullptr, (size_t)MA * np, np,
                         l.b_off >= 0 ? grad + l.b_off : (float*)nullptr));
  if (dx_follows) {
    h->dy_ready = dY;
    h->dy_ready_m = M;
    h->dy_ready_n = l.N;
    h->dy_ready_buf = 0;
  }
  if (h->n_side) {   // dY (and X, which nobody overwrites within a pass) have been read
    auto prev = h->dy_consumed.find(dY);   // an earlier reader of the same buffer on another side stream: chain the events
    if (prev != h->dy_consumed.end()) AIR_CUDA(cudaStreamWaitEvent(st, prev->second, 0));
    cudaEvent_t done = next_event(h);
    AIR_CUDA(cudaEventRecord(done, st));
    h->dy_consumed[dY] = done;
  }
  const CUtensorMap *tm_a = nullptr, *tm_b = nullptr;
  int32_t rc = get_tmap2(h, hl_xt, mp, 2LL * KA, tc::BM, &tm_a);
  if (rc != AIR_OK) return rc;
  if ((rc = get_tmap2(h, hl_yt, mp, 2LL * NA, 64, &tm_b)) != AIR_OK) return rc;
  tc::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out_f32 = grad + l.w_off;
  p.ldc = l.N;
  p.M = l.K;
  p.N = l.N;
  p.num_k_blocks = mp / tc::BK;
  p.a_lo_row = KA;
  p.b_lo_row = NA;
  p.act = air::ACT_NONE;
  p.range_flag = h->t_range_flag;
  p.out_scale = 1.0f;
  p.atomic_out = 1;
  p.ab_bf16 = 1;
  const int tiles = (KA / tc::BM) * (NA / 64);
  // enough slices to fill the machine, but at least 8 K blocks per slice: a slice ends in 128 x 64 fp32 atomics
  int split = (2 * 148 + tiles - 1) / tiles;
  split = std::max(1, std::min(split, p.num_k_blocks / 8));
  p.kb_per_z = (p.num_k_blocks + split - 1) / split;
  AIR_CUDA(tc::launch_gemm(64, *tm_a, *tm_b, p, NA, st));
  h->launches += 3;
  return AIR_OK;
}

int32_t layer_param_grads(air_handle* h, float* grad, const Layer& l, const float* X, int ldx, const float* dY, int ldy,
                          int M, cudaStream_t st, bool dx_follows = false) {
  if (h->tc_bwd && M >= 64) return layer_weight_grad_tc(h, grad, l, X, ldx, dY, ldy, M, dx_follows, st);
  AIR_CUDA(air::launch_gemm_simt(true, false, X, ldx, dY, ldy, grad + l.w_off, l.N, l.K, l.N, M, true, nullptr, 0,
                                 pick_split(l.K, l.N, M), st));
  ++h->launches;
  if (l.b_off >= 0) {
    AIR_CUDA(air::launch_colsum(dY, ldy, grad + l.b_off, M, l.N, st));
    ++h->launches;
  }
  return AIR_OK;
}
// dX = dY @ W^T on the tensor-core split engine: dY [M, N] as row-major bf16 hi/lo planes (A operand, contraction over the
// layer's N outputs), W [K, N] exactly as stored (its rows are the output columns of dX) from the per-step weight arena
// prepared by prep_backward_weights; elu' mask, accumulation (atomicAdd) in the epilogue.
int32_t layer_input_grad_tc(air_handle* h, const Layer& l, const float* dY, int ldy, float* dX, int ldx, int M,
                            bool accumulate, const float* elu_x, int ld_elu, bool planes_for_next, cudaStream_t st) {
  namespace tc = air::tc;
  const int np = round_up(l.N, 64), MA = round_up(M, 128), KA = round_up(l.K, 64);
  if ((size_t)MA * np > h->hl_dy_halves) return fail(AIR_ERR_ARG, "internal: dY does not fit the training workspace");
  const auto it = h->wnt_index.find(l.w_off);
  if (it == h->wnt_index.end()) return fail(AIR_ERR_ARG, "internal: weight not in the backward arena");
  int buf = h->dy_ready_buf;
  if (h->dy_ready != dY || h->dy_ready_m != M || h->dy_ready_n != l.N) {   // not left behind by the previous GEMM's epilogue
    buf = 0;
    const size_t n4 = (size_t)M * (np / 4);
    AIR_CUDA(air::launch_k(tc::split_rows_bf16_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, dY, ldy, M, l.N,
                           h->hl_dy2[0], (size_t)MA * np, np, h->t_range_flag));
    ++h->launches;
  }
  h->dy_ready = nullptr;
  const CUtensorMap *tm_a = nullptr, *tm_b = nullptr;
  int32_t rc = get_tmap2(h, h->hl_dy2[buf], np, 2LL * MA, tc::BM, &tm_a);
  if (rc != AIR_OK) return rc;
  if ((rc = get_tmap2(h, h->wnt_arena + it->second.first, np, 2LL * KA, 64, &tm_b)) != AIR_OK) return rc;
  tc::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out_f32 = dX;
  p.ldc = ldx;
  p.M = M;
  p.N = l.K;
  p.num_k_blocks = np / tc::BK;
  p.a_lo_row = MA;
  p.b_lo_row = KA;
  p.act = air::ACT_NONE;
  p.range_flag = h->t_range_flag;
  p.out_scale = 1.0f;
  p.atomic_out = accumulate ? 1 : 0;
  p.ab_bf16 = 1;
  p.mask_y = elu_x;
  p.ld_mask = ld_elu;
  if (planes_for_next && !accumulate && (size_t)MA * KA <= h->hl_dy_halves) {
    // dX is the dY of the layer below: its row-major bf16 planes come straight out of this epilogue
    p.out_hl = h->hl_dy2[buf ^ 1];
    p.hl_plane = (size_t)MA * KA;
    p.ld_hl = KA;
    p.hl_bf16 = 1;
    h->dy_ready = dX;
    h->dy_ready_m = M;
    h->dy_ready_n = l.K;
    h->dy_ready_buf = buf ^ 1;
  }
  AIR_CUDA(tc::launch_gemm(64, *tm_a, *tm_b, p, KA, st));
  ++h->launches;
  return AIR_OK;
}

// all weight matrices, as stored, -> bf16 hi/lo planes with zero-padded columns (once per backward pass)
int32_t prep_backward_weights(air_handle* h, const float* params, cudaStream_t st) {
  AIR_CUDA(air::launch_k(air::tc::prep_weights_rows_kernel, dim3(h->wnt_blocks), dim3(256), 0, st, params, h->wnt_arena,
                         (const air::tc::RowsEntry*)h->wnt_table, h->wnt_entries));
  ++h->launches;
  return AIR_OK;
}
// host table of prep_weights_rows_kernel (uploaded once by air_train_enable)
int32_t upload_backward_weight_table(air_handle* h) {
  std::vector<air::tc::RowsEntry> table;
  int blocks = 0;
  auto one = [&](const Layer& l) {
    const auto it = h->wnt_index.find(l.w_off);
    air::tc::RowsEntry e;
    e.src_off = l.w_off;
    e.dst_off = (int64_t)it->second.first;
    e.np = it->second.second;
    e.plane = (int64_t)round_up(l.K, 64) * e.np;
    e.K = l.K;
    e.N = l.N;
    e.block_begin = blocks;
    blocks += (int)(((size_t)l.K * (e.np / 4) + 255) / 256);
    table.push_back(e);
  };
  for (const Mlp* m : {&h->enc, &h->where_mlp, &h->steps_mlp, &h->glenc, &h->dec})
    for (size_t i = (m == &h->enc ? 1 : 0); i < m->layers.size(); ++i) one(m->layers[i]);   // enc layer 0 needs no dX
  for (const Layer* l : {&h->what_lin, &h->lstm_h, &h->lstm_x}) one(*l);
  h->wnt_entries = (int)table.size();
  h->wnt_blocks = blocks;
  AIR_CUDA(cudaMemcpy(h->wnt_table, table.data(), sizeof(air::tc::RowsEntry) * table.size(), cudaMemcpyHostToDevice));
  return AIR_OK;
}

// dX = dY @ W^T (* elu'(X) when elu_x is the saved forward value of X)
int32_t layer_input_grad(air_handle* h, const float* params, const Layer& l, const float* dY, int ldy, float* dX, int ldx,
                         int M, bool accumulate, const float* elu_x, int ld_elu, cudaStream_t st,
                         bool planes_for_next = false) {
  const int32_t rcw = wait_consumed(h, dX, st);
  if (rcw != AIR_OK) return rcw;
  if (h->tc_bwd && M >= 64)
    return layer_input_grad_tc(h, l, dY, ldy, dX, ldx, M, accumulate, elu_x, ld_elu, planes_for_next, st);
  AIR_CUDA(air::launch_gemm_simt(false, true, dY, ldy, params + l.w_off, l.N, dX, ldx, M, l.K, l.N, accumulate, elu_x,
                                 ld_elu, 1, st));
  ++h->launches;
  return AIR_OK;
}

// Backward of a whole neural.MLP (neural.py:63-102).  x0 [M, layers[0].K] is the MLP input, `saves` its hidden
// activations (outputs of layers 0 .. L-2), dY the gradient with respect to the LAST layer's pre-activation (callers mask
// with elu' first when the last layer is an ELU layer).  dY is consumed; intermediate gradients ping-pong between
// h->g_a / h->g_b.  dx0 != null: gradient with respect to x0 (no activation mask) is written (or accumulated) there.
int32_t mlp_backward(air_handle* h, const float* params, float* grad, const Mlp& mlp, const float* x0, int ld0,
                     const std::vector<float*>& saves, const float* dY, int ldy, int M, float* dx0, int ld_dx0,
                     bool dx0_accumulate, cudaStream_t st) {
  const int nl = (int)mlp.layers.size();
  const float* cur = dY;
  int ld_cur = ldy;
  for (int i = nl - 1; i >= 0; --i) {
    const Layer& l = mlp.layers[i];
    const float* X = i == 0 ? x0 : saves[i - 1];
    const int ldx = i == 0 ? ld0 : mlp.layers[i - 1].N;
    int32_t rc = layer_param_grads(h, grad, l, X, ldx, cur, ld_cur, M, st, i > 0 || dx0 != nullptr);
    if (rc != AIR_OK) return rc;
    if (i > 0) {
      float* dst = (cur == h->g_a) ? h->g_b : h->g_a;
      rc = layer_input_grad(h, params, l, cur, ld_cur, dst, l.K, M, false, X, ldx, st,   // X is an ELU output
                            /*planes_for_next=*/i - 1 > 0 || dx0 != nullptr);
      if (rc != AIR_OK) return rc;
      cur = dst;
      ld_cur = l.K;
    } else if (dx0) {
      rc = layer_input_grad(h, params, l, cur, ld_cur, dx0, ld_dx0, M, dx0_accumulate, nullptr, 0, st);
      if (rc != AIR_OK) return rc;
    }
  }
  return AIR_OK;
}

int32_t join_side_streams(air_handle* h, cudaStream_t st) {
  for (int i = 0; i < h->n_side; ++i) {
    cudaEvent_t done = next_event(h);
    AIR_CUDA(cudaEventRecord(done, h->side[i]));
    AIR_CUDA(cudaStreamWaitEvent(st, done, 0));
  }
  h->dy_consumed.clear();
  h->side_next = 0;
  return AIR_OK;
}

int32_t backward_impl(air_handle* h, const float* params, const float* img, const float* eps_where,
                      const float* eps_what, const air_prior* prior, const air_outputs* o, float baseline_mean,
                      float inv_batch, float l2_weight, float* grad, cudaStream_t st) {
  const air_config& c = h->cfg;
  const int B = c.B, T = c.T, nh = c.nh, P = h->P, G = h->G, na = c.na;
  const int TB = T * B;
  const int thr = 256;
  int32_t rc;
  AIR_CUDA(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)h->n_params, st));
  h->dy_consumed.clear();
  if (h->tc_bwd && TB >= 64 && (rc = prep_backward_weights(h, params, st)) != AIR_OK) return rc;

  air::BwdArgs a;
  memset(&a, 0, sizeof(a));
  a.img = img;
  a.canvas_final = o->canvas + (size_t)(T - 1) * B * P;
  a.glimpse = o->glimpse;
  a.where = o->where;
  a.where_loc = o->where_loc;
  a.where_scale = o->where_scale;
  a.eps_where = eps_where;
  a.what_loc = o->what_loc;
  a.what_scale = o->what_scale;
  a.eps_what = eps_what;
  a.presence = o->presence;
  a.presence_prob = o->presence_prob;
  a.posterior = o->num_steps_posterior;
  a.num_step = o->num_step_per_sample;
  a.step_weight = o->prior_step_weight;
  a.rec_ps = o->rec_loss_per_sample;
  a.kl_n_ps = o->kl_num_steps_per_sample;
  a.kl_what_ps = o->kl_what_per_sample;
  a.kl_where_ps = o->kl_where_per_sample;
  a.dglimpse = h->g_glimpse;
  a.dwhere_paint = h->g_wh_paint;
  a.dcrop = h->g_crop;
  a.dwhere_read = h->g_wh_read;
  a.dwhat = h->g_what;
  a.dr = h->g_r;
  a.dm = h->g_m;
  a.dlogit = h->g_logit;
  a.discrete = c.discrete_steps;
  a.dpresence = c.discrete_steps ? nullptr : h->g_pres;   // non-discrete steps: the painted canvas depends on presence = p
  a.T = T; a.B = B; a.H = c.H; a.W = c.W; a.h = c.h; a.w = c.w; a.na = na;
  a.output_std = c.output_std;
  a.output_multiplier = c.output_multiplier;
  a.max_crop = c.max_crop_size;
  a.explore_eps = c.explore_eps;
  a.inv_batch = inv_batch > 0.f ? inv_batch : 1.0f / (float)B;
  a.baseline_mean = baseline_mean;
  // NaN: the mean is whatever scalars[AIR_S_MEAN_BASELINE] holds on the device when the kernel runs (the last
  // air_forward / air_elbo_scalars with a baseline, or the all-reduced block under sharding) -- no host round trip
  a.baseline_mean_dev = (baseline_mean != baseline_mean && o && o->scalars) ? o->scalars + AIR_S_MEAN_BASELINE : nullptr;
  if (baseline_mean != baseline_mean && !a.baseline_mean_dev)
    return fail(AIR_ERR_ARG, "air_backward: baseline_mean = NaN needs outs->scalars");
  a.step_W = c.W > 1 ? 2.0 / (double)(c.W - 1) : 0.0;
  a.step_H = c.H > 1 ? 2.0 / (double)(c.H - 1) : 0.0;
  a.step_w = c.w > 1 ? 2.0 / (double)(c.w - 1) : 0.0;
  a.step_h = c.h > 1 ? 2.0 / (double)(c.h - 1) : 0.0;
  a.prior = *prior;
  for (int k = 0; k <= T; ++k)
    a.steps_prior[k] = prior->steps_prob_is_f64 ? air::geom_prior_f64(prior->steps_success_prob, k)
                                                : (double)air::geom_prior_f32((float)prior->steps_success_prob, k);
  a.steps_prior_dev = h->prior_dev_on ? h->prior_dev : nullptr;

  // 1. reconstruction term -> d glimpse, d where (inverse transformer)            cell.py:159-164, model.py:319-321
  AIR_CUDA(air::launch_paint_bwd(a, st));
  ++h->launches;
  // 2. decoder                                                                    cell.py:158
  if ((rc = mlp_backward(h, params, grad, h->dec, o->what, na, h->sv_dec, h->g_glimpse, G, TB, h->g_what, na, false,
                         st)) != AIR_OK)
    return rc;
  // 3. what sample + KL(what)                                                     cell.py:154-156, model.py:174-186
  {
    const size_t n = (size_t)TB * na;
    AIR_CUDA(air::launch_k(air::what_bwd_kernel, dim3((unsigned)((n + thr - 1) / thr)), dim3(thr), 0, st, a));
    ++h->launches;
  }
  // 4. what head (linear) and glimpse encoder -> d crop                           cell.py:153, modules.py:20
  if ((rc = layer_param_grads(h, grad, h->what_lin, h->sv_q, h->what_lin.K, h->g_r, 2 * na, TB, st, true)) != AIR_OK)
    return rc;
  if ((rc = layer_input_grad(h, params, h->what_lin, h->g_r, 2 * na, h->g_a, h->what_lin.K, TB, false, h->sv_q,
                             h->what_lin.K, st, /*planes_for_next=*/true)) != AIR_OK)
    return rc;
  if ((rc = mlp_backward(h, params, grad, h->glenc, h->crop.f32, G, h->sv_glenc, h->g_a, h->what_lin.K, TB, h->g_crop, G,
                         false, st)) != AIR_OK)
    return rc;
  // 5. glimpse read -> d where                                                    cell.py:135
  AIR_CUDA(air::launch_read_bwd(a, st));
  ++h->launches;
  // 6. where sample + KL(where) -> dm ; step-count posterior (KL(n), q(n) weights, REINFORCE) -> dlogit
  AIR_CUDA(air::launch_latent_bwd(a, st));
  ++h->launches;
  // 7. the two heads on h_t -> dh                                                 modules.py:58-63,119-122
  if ((rc = mlp_backward(h, params, grad, h->where_mlp, h->hs.f32, nh, h->sv_where, h->g_m, 8, TB, h->g_h, nh, false,
                         st)) != AIR_OK)
    return rc;
  if ((rc = mlp_backward(h, params, grad, h->steps_mlp, h->hs.f32, nh, h->sv_steps, h->g_logit, 1, TB, h->g_h, nh, true,
                         st)) != AIR_OK)
    return rc;
  // 8. LSTM through time                                                          cell.py:126-127
  // (tensor-core gradient GEMMs: the gate kernel also writes dgates_t as the bf16 hi/lo planes the recurrent product reads)
  const int g_np = 4 * nh, g_ma = round_up(B, 128);
  const bool gate_planes = h->tc_bwd && B >= 64 && g_np % 64 == 0 && (size_t)g_ma * g_np <= h->hl_dy_halves;
  for (int t = T - 1; t >= 0; --t) {
    const size_t off = (size_t)t * B;
    const bool pair = nh % 2 == 0;
    AIR_CUDA(air::launch_k(pair ? air::lstm_bwd_pointwise_kernel<2> : air::lstm_bwd_pointwise_kernel<1>,
                           dim3((B * (pair ? nh / 2 : nh) + thr - 1) / thr), dim3(thr), 0, st,
                           (const float*)(h->gates_all + off * 4 * nh), (const float*)(h->c_all + off * nh),
                           (const float*)(h->c_all + (off + B) * nh), (const float*)(h->g_h + off * nh),
                           (const float*)(t == T - 1 ? nullptr : h->g_hrec), h->g_c, h->g_gates + off * 4 * nh, B, nh,
                           c.forget_bias, t == T - 1 ? 1 : 0, h->g_gx, gate_planes ? h->hl_dy2[0] : (__half*)nullptr,
                           (size_t)g_ma * g_np, g_np, h->t_range_flag));
    ++h->launches;
    if (gate_planes) {
      h->dy_ready = h->g_gates + off * 4 * nh;
      h->dy_ready_m = B;
      h->dy_ready_n = g_np;
      h->dy_ready_buf = 0;
    }
    // d h_{t-1} = dgates_t @ W_h^T  (t = 0: gradient of the trainable initial state)
    if ((rc = layer_input_grad(h, params, h->lstm_h, h->g_gates + off * 4 * nh, 4 * nh, h->g_hrec, nh, B, false, nullptr,
                               0, st)) != AIR_OK)
      return rc;
  }
  AIR_CUDA(air::launch_colsum(h->g_hrec, nh, grad + h->lstm_h0, B, nh, st));
  AIR_CUDA(air::launch_colsum(h->g_c, nh, grad + h->lstm_c0, B, nh, st));
  h->launches += 2;
  {
    Layer wh = h->lstm_h;
    wh.b_off = h->lstm_b;   // db = colsum over all T*B gate rows
    if ((rc = layer_param_grads(h, grad, wh, h->hprev, nh, h->g_gates, 4 * nh, TB, st)) != AIR_OK) return rc;
    Layer wx = h->lstm_x;   // (d gx = sum_t dgates_t was accumulated by the gate kernels)
    wx.b_off = -1;
    if ((rc = layer_param_grads(h, grad, wx, h->e.f32, h->n_enc, h->g_gx, 4 * nh, B, st, true)) != AIR_OK) return rc;
    // 9. input encoder (its last layer is an ELU layer: mask with the saved output e)   cell.py:125
    if ((rc = layer_input_grad(h, params, wx, h->g_gx, 4 * nh, h->g_e, h->n_enc, B, false, h->e.f32, h->n_enc, st,
                               /*planes_for_next=*/h->enc.layers.size() > 1)) != AIR_OK)
      return rc;
    if ((rc = mlp_backward(h, params, grad, h->enc, img, P, h->sv_enc, h->g_e, h->n_enc, B, nullptr, 0, false, st)) !=
        AIR_OK)
      return rc;
  }
  // join: the weight gradients of the side streams are part of this call's result
  if ((rc = join_side_streams(h, st)) != AIR_OK) return rc;
  // l2_weight * sum of tf.nn.l2_loss over the 2-D variables (model.py:345-350): weights and the [1,nh] initial state
  if (l2_weight > 0.f) {
    for (const ParamEntry& e : h->entries) {
      const bool two_d = e.name.size() > 2 && (e.name.compare(e.name.size() - 2, 2, ".w") == 0 || e.name == "lstm.h0" ||
                                               e.name == "lstm.c0");
      if (!two_d) continue;
      const size_t n = (size_t)e.rows * e.cols;
      air::l2_grad_kernel<<<(unsigned)((n + thr - 1) / thr), thr, 0, st>>>(params + e.offset, grad + e.offset, l2_weight, n);
      ++h->launches;
    }
    AIR_CUDA(cudaGetLastError());
  }
  return AIR_OK;
}

}  // namespace

extern "C" {

int32_t air_abi_version(void) { return AIR_ABI_VERSION; }

const char* air_last_error(void) { return g_last_error.c_str(); }

int32_t air_create(const air_config* cfg, air_handle** out) {
  if (!cfg || !out) return fail(AIR_ERR_ARG, "air_create: NULL argument");
  *out = nullptr;
  const air_config& c = *cfg;
  if (c.B < 1 || c.H < 1 || c.W < 1 || c.h < 1 || c.w < 1 || c.na < 1 || c.nh < 1)
    return fail(AIR_ERR_ARG, "air_create: sizes must be positive");
  if (c.T < 1 || c.T > AIR_MAX_STEPS) return fail(AIR_ERR_ARG, "air_create: T must be in [1, AIR_MAX_STEPS]");
  if (!valid_hidden(c.enc_hidden, c.n_enc_hidden) || !valid_hidden(c.glenc_hidden, c.n_glenc_hidden) ||
      !valid_hidden(c.dec_hidden, c.n_dec_hidden) || !valid_hidden(c.where_hidden, c.n_where_hidden) ||
      !valid_hidden(c.steps_hidden, c.n_steps_hidden))
    return fail(AIR_ERR_ARG, "air_create: every MLP needs 1..AIR_MAX_HIDDEN positive hidden widths");
  if (c.precision != AIR_PREC_FP32 && c.precision != AIR_PREC_TC_SPLIT)
    return fail(AIR_ERR_ARG, "air_create: unknown precision");
  if (!(c.output_std > 0.f)) return fail(AIR_ERR_ARG, "air_create: output_std must be > 0");

  int dev = 0;
  AIR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  AIR_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(AIR_ERR_ARCH, std::string("air_create: built for sm_100a only, device is sm_") +
                                  std::to_string(prop.major) + std::to_string(prop.minor));

  air_handle* h = new air_handle();
  h->cfg = c;
  h->P = c.H * c.W;
  h->G = c.h * c.w;
  h->use_tc = c.precision == AIR_PREC_TC_SPLIT;
  // canonical flat parameter order == the Sonnet variables of cell.py:61-69 in creation order
  h->n_enc = build_mlp(h, h->enc, "input_encoder", h->P, c.enc_hidden, c.n_enc_hidden, 0);
  h->lstm_w = add_entry(h, "lstm.w", h->n_enc + c.nh, 4 * c.nh);
  h->lstm_b = add_entry(h, "lstm.b", 1, 4 * c.nh);
  h->lstm_h0 = add_entry(h, "lstm.h0", 1, c.nh);
  h->lstm_c0 = add_entry(h, "lstm.c0", 1, c.nh);
  h->lstm_x.w_off = h->lstm_w;                                  // rows [0, n_enc) of lstm.w: the input half
  h->lstm_x.b_off = h->lstm_b;
  h->lstm_x.K = h->n_enc;
  h->lstm_x.N = 4 * c.nh;
  h->lstm_h.w_off = h->lstm_w + (int64_t)h->n_enc * 4 * c.nh;   // rows [n_enc, n_enc + nh): the recurrent half
  h->lstm_h.K = c.nh;
  h->lstm_h.N = 4 * c.nh;
  build_mlp(h, h->where_mlp, "transform_estimator", c.nh, c.where_hidden, c.n_where_hidden, 8);
  build_mlp(h, h->steps_mlp, "steps_predictor", c.nh, c.steps_hidden, c.n_steps_hidden, 1);
  const int n_gl = build_mlp(h, h->glenc, "glimpse_encoder", h->G, c.glenc_hidden, c.n_glenc_hidden, 0);
  h->what_lin.K = n_gl;
  h->what_lin.N = 2 * c.na;
  h->what_lin.w_off = add_entry(h, "what.w", n_gl, 2 * c.na);
  h->what_lin.b_off = add_entry(h, "what.b", 1, 2 * c.na);
  build_mlp(h, h->dec, "glimpse_decoder", c.na, c.dec_hidden, c.n_dec_hidden, h->G);

  if (h->use_tc) {
    if (!air::tc::get_encode_fn()) {
      delete h;
      return fail(AIR_ERR_CUDA, "air_create: cuTensorMapEncodeTiled is not available from the driver");
    }
    for (Mlp* mlp : {&h->enc, &h->where_mlp, &h->steps_mlp, &h->glenc, &h->dec}) {
      const bool chain_feeds_gemm = (mlp == &h->enc || mlp == &h->glenc);   // their last layer is another GEMM's input
      for (size_t i = 0; i < mlp->layers.size(); ++i)
        register_tc_weight(h, mlp->layers[i], chain_feeds_gemm || i + 1 < mlp->layers.size());
    }
    register_tc_weight(h, h->lstm_x, false);
    register_tc_weight(h, h->lstm_h, false);
    register_tc_weight(h, h->what_lin, false);
    // fused chains (chain_tc.cuh): every hidden activation must fit the 256-K TMEM operand
    h->na_off = round_up(c.na, 16);
    bool ok = getenv("AIR_NO_CHAIN") == nullptr && 2 * h->na_off <= 256;
    for (const Mlp* mlp : {&h->where_mlp, &h->steps_mlp, &h->glenc, &h->dec})
      for (int i = 0; i < mlp->n_hidden; ++i) ok = ok && mlp->layers[i].N <= 256;
    h->chain_ok = ok;
    if (ok) {
      h->what_chain = h->what_lin;
      TcWeight w;
      w.src_off = h->what_lin.w_off;
      w.K = h->what_lin.K;
      w.N = h->what_lin.N;
      w.Kpad = round_up(w.K, air::tc::BK);
      w.BN = 64;
      w.N_alloc = 2 * h->na_off;
      w.split_n = c.na;
      w.split_off = h->na_off;
      w.bias_src = h->what_lin.b_off;
      h->what_chain.tc = (int)h->tcw.size();
      h->tcw.push_back(w);
    }
    // cluster LSTM kernel: snt.LSTM(256) with an encoder output that fits the TMEM operand
    h->lstm_ok = h->chain_ok && getenv("AIR_NO_LSTM_CLUSTER") == nullptr && c.nh == air::lstm::NH && h->n_enc <= 256;
    if (h->lstm_ok) {
      for (int which = 0; which < 2; ++which) {
        const Layer& l = which == 0 ? h->lstm_x : h->lstm_h;
        TcWeight w;
        w.src_off = l.w_off;
        w.K = l.K;
        w.N = l.N;
        w.Kpad = round_up(w.K, air::tc::BK);
        w.BN = 64;
        w.N_alloc = l.N;
        w.perm_nh = c.nh;
        w.bias_src = which == 0 ? h->lstm_b : -1;
        (which == 0 ? h->lstm_x_perm : h->lstm_h_perm) = (int)h->tcw.size();
        h->tcw.push_back(w);
      }
    }
  }

  // shared-memory budgets of the per-canvas kernels
  const size_t smem_read = air::where_read_smem(c.T, c.H, c.W, c.h, c.w);
  const size_t smem_paint = air::paint_smem(c.T, c.H, c.W, c.h, c.w);
  if (smem_read > 200 * 1024 || smem_paint > 200 * 1024) {
    delete h;
    return fail(AIR_ERR_ARG, "air_create: image / glimpse tile does not fit in shared memory");
  }
  if (smem_read > 48 * 1024) {
    cudaFuncSetAttribute(air::where_read_kernel<0, 0, 0, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_read);
    cudaFuncSetAttribute(air::where_read_kernel<3, 50, 50, 20, 20, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_read);
    cudaFuncSetAttribute(air::where_read_kernel<5, 100, 100, 28, 28, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_read);
    cudaFuncSetAttribute(air::stn_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_read);
  }

  // workspace: size it, allocate once, carve it
  Carver sizing(nullptr);
  carve_workspace(h, sizing);
  cudaError_t e = cudaMalloc(&h->ws, sizing.off);
  if (e != cudaSuccess) {
    delete h;
    return fail(AIR_ERR_NOMEM, std::string("air_create: cudaMalloc workspace: ") + cudaGetErrorString(e));
  }
  h->ws_bytes = sizing.off;
  Carver real(h->ws);
  carve_workspace(h, real);
  if (h->use_tc) {
    // zero once: the K padding of every hl operand and the N padding of the weight arena must stay zero
    e = cudaMemset(h->ws, 0, h->ws_bytes);
    std::vector<air::tc::PrepEntry> table;
    int tiles = 0;
    bool ok = e == cudaSuccess, row_tm_ok = true;
    for (TcWeight& w : h->tcw) {
      air::tc::PrepEntry pe;
      pe.src_off = w.src_off;
      pe.dst_off = w.arena_off;
      pe.plane = (int64_t)w.N_alloc * w.Kpad;
      pe.K = w.K;
      pe.N = w.N;
      pe.Kpad = w.Kpad;
      pe.tile_begin = tiles;
      pe.tiles_n = (w.N + 31) / 32;
      pe.split_n = w.split_n;
      pe.split_off = w.split_off;
      pe.bias_src = w.bias_src;
      pe.bias_dst = w.bias_off;
      pe.perm_nh = w.perm_nh;
      tiles += pe.tiles_n * ((w.K + 31) / 32);
      table.push_back(pe);
      ok = ok && air::tc::make_tmap(&w.tm, h->arena + w.arena_off, w.Kpad, 2 * (int64_t)w.N_alloc, w.BN);
      if (h->chain_ok) {
        const int n_eff = w.split_n > 0 ? w.N_alloc : w.N;
        w.n_pass = (n_eff + 255) / 256;
        w.n_box = w.perm_nh > 0 ? w.perm_nh : round_up((n_eff + w.n_pass - 1) / w.n_pass, 16);
        if (w.n_pass * w.n_box > w.N_alloc) h->chain_ok = false;
        else
          ok = ok && air::chain::make_weight_tmap(&w.tm_chain, h->arena + w.arena_off, w.Kpad, w.N_alloc, w.n_box);
      }
      row_tm_ok = row_tm_ok && air::row::make_row_weight_tmap(&w.tm_row, h->arena + w.arena_off, w.Kpad, w.N_alloc);
    }
    h->row_ok = h->chain_ok && row_tm_ok && row_schedule_check(c).empty();
    h->enc1_ok = h->chain_ok && h->enc.layers[0].N == air::enc::N1 && (h->P % 4) == 0;
    h->prep_tiles = tiles;
    if (ok)
      ok = cudaMemcpy(h->prep_table, table.data(), sizeof(air::tc::PrepEntry) * table.size(),
                      cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
      cudaFree(h->ws);
      delete h;
      return fail(AIR_ERR_CUDA, "air_create: tensor-core engine set-up failed (memset / tensor maps / table upload)");
    }
  }
  *out = h;
  return AIR_OK;
}

int32_t air_destroy(air_handle* h) {
  if (!h) return AIR_OK;
  for (cudaEvent_t e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->ws) cudaFree(h->ws);
  if (h->tws) cudaFree(h->tws);
  if (h->bl.ws) cudaFree(h->bl.ws);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (int i = 0; i < h->n_side; ++i) cudaStreamDestroy(h->side[i]);
  if (h->trace) cudaFree(h->trace);
  if (h->feed_stream) {
    cudaStreamSynchronize(h->feed_stream);
    cudaStreamDestroy(h->feed_stream);
    for (int s = 0; s < 2; ++s) {
      cudaEventDestroy(h->feed_fed[s]);
      cudaEventDestroy(h->feed_consumed[s]);
      cudaEventDestroy(h->feed_done[s]);
    }
    cudaFree(h->feed_buf[0]);
  }
  delete h;
  return AIR_OK;
}

int32_t air_row_schedule_check(const air_config* cfg, char* msg, int32_t msg_len) {
  if (!cfg) return fail(AIR_ERR_ARG, "air_row_schedule_check: NULL config");
  const std::string r = row_schedule_check(*cfg);
  if (msg && msg_len > 0) {
    strncpy(msg, r.c_str(), (size_t)msg_len - 1);
    msg[msg_len - 1] = 0;
  }
  return r.empty() ? AIR_OK : fail(AIR_ERR_ARG, "row schedule: " + r);
}

int64_t air_param_count(const air_handle* h) { return h ? h->n_params : 0; }
int32_t air_param_entries(const air_handle* h) { return h ? (int32_t)h->entries.size() : 0; }
int32_t air_param_entry(const air_handle* h, int32_t i, const char** name, int64_t* offset, int32_t* rows,
                        int32_t* cols) {
  if (!h || i < 0 || i >= (int32_t)h->entries.size()) return fail(AIR_ERR_ARG, "air_param_entry: bad index");
  const ParamEntry& e = h->entries[i];
  if (name) *name = e.name.c_str();
  if (offset) *offset = e.offset;
  if (rows) *rows = e.rows;
  if (cols) *cols = e.cols;
  return AIR_OK;
}
int64_t air_workspace_bytes(const air_handle* h) { return h ? (int64_t)h->ws_bytes : 0; }

int64_t air_launch_count(const air_handle* h) { return h ? (int64_t)h->launches : 0; }

int32_t air_profile_enable(air_handle* h, int32_t on) {
  if (!h) return fail(AIR_ERR_ARG, "air_profile_enable: NULL handle");
  if (on && !h->ev[0])
    for (cudaEvent_t& e : h->ev) AIR_CUDA(cudaEventCreate(&e));
  h->profile = on != 0;
  return AIR_OK;
}

int32_t air_profile_read(air_handle* h, float* ms_per_stage, int32_t n) {
  if (!h || !ms_per_stage || n < AIR_N_STAGES) return fail(AIR_ERR_ARG, "air_profile_read: bad argument");
  if (!h->ev[0]) return fail(AIR_ERR_ARG, "air_profile_read: profiling was never enabled");
  AIR_CUDA(cudaEventSynchronize(h->ev[AIR_N_STAGES]));
  for (int i = 0; i < AIR_N_STAGES; ++i) AIR_CUDA(cudaEventElapsedTime(&ms_per_stage[i], h->ev[i], h->ev[i + 1]));
  return AIR_OK;
}

const char* air_stage_name(int32_t i) { return (i >= 0 && i < AIR_N_STAGES) ? kStageNames[i] : ""; }

int32_t air_check_range(air_handle* h, void* stream) {
  if (!h) return fail(AIR_ERR_ARG, "air_check_range: NULL handle");
  if (!h->use_tc && !h->t_range_flag) return AIR_OK;
  int flag = 0, tflag = 0;
  if (h->use_tc)
    AIR_CUDA(cudaMemcpyAsync(&flag, h->range_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (h->t_range_flag)
    AIR_CUDA(cudaMemcpyAsync(&tflag, h->t_range_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  AIR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (tflag) {
    AIR_CUDA(cudaMemsetAsync(h->t_range_flag, 0, sizeof(int), (cudaStream_t)stream));
    return fail(AIR_ERR_RANGE, "a non-finite activation or gradient reached the tensor-core gradient GEMMs (bf16 hi/lo "
                               "planes cover the whole fp32 range); AIR_NO_TC_BWD=1 selects the fp32 SIMT GEMMs");
  }
  if (flag) {
    AIR_CUDA(cudaMemsetAsync(h->range_flag, 0, sizeof(int), (cudaStream_t)stream));
    return fail(AIR_ERR_RANGE, "a weight*2^8 or an activation exceeded the fp16 range (65504) in the tensor-core "
                               "split engine; use AIR_PREC_FP32 for this model");
  }
  return AIR_OK;
}

int32_t air_train_enable(air_handle* h, int32_t on) {
  if (!h) return fail(AIR_ERR_ARG, "air_train_enable: NULL handle");
  if (on && !h->tws) {
    // (checked before anything is allocated: a failed enable must leave the handle exactly as it was, so that a retry
    // runs the whole initialisation again instead of finding a half-built workspace)
    const size_t smem = air::paint_bwd_smem(h->cfg.T, h->cfg.H, h->cfg.W, h->cfg.h, h->cfg.w);
    const size_t smem_r = air::read_bwd_smem(h->cfg.T, h->cfg.H, h->cfg.W, h->cfg.h, h->cfg.w);
    if (smem > 200 * 1024 || smem_r > 200 * 1024)
      return fail(AIR_ERR_ARG, "air_train_enable: image / glimpse tile does not fit in shared memory");
    const int32_t rc_init = [&]() -> int32_t {
    h->tc_bwd = getenv("AIR_NO_TC_BWD") == nullptr && air::tc::get_encode_fn() != nullptr;
    // side streams for the weight-gradient work: AIR_SIDE_STREAMS=n (1..MAX_SIDE, default 2), AIR_NO_SIDE_STREAM=1 for none
    int want_side = 2;
    if (const char* v = getenv("AIR_SIDE_STREAMS")) want_side = std::max(0, std::min(atoi(v), (int)air_handle::MAX_SIDE));
    if (getenv("AIR_NO_SIDE_STREAM") != nullptr || !h->tc_bwd) want_side = 0;
    h->n_side_bufs = std::max(1, want_side);
    Carver sizing(nullptr);
    carve_train(h, sizing);
    cudaError_t e = cudaMalloc(&h->tws, sizing.off);
    if (e != cudaSuccess) return fail(AIR_ERR_NOMEM, std::string("air_train_enable: cudaMalloc: ") + cudaGetErrorString(e));
    h->tws_bytes = sizing.off;
    Carver real(h->tws);
    carve_train(h, real);
    if (h->t_range_flag) AIR_CUDA(cudaMemset(h->t_range_flag, 0, sizeof(int)));
    if (h->tc_bwd) {
      const int32_t rc = upload_backward_weight_table(h);
      if (rc != AIR_OK) return rc;
      for (int i = 0; i < want_side; ++i) {
        AIR_CUDA(cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
        h->n_side = i + 1;
      }
      if (h->n_side) {
        h->ev_pool.assign(128, nullptr);
        for (cudaEvent_t& e : h->ev_pool) AIR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      }
    }
    return AIR_OK;
    }();
    if (rc_init != AIR_OK) {   // undo: streams, events, workspace
      for (int i = 0; i < h->n_side; ++i) cudaStreamDestroy(h->side[i]);
      h->n_side = 0;
      for (cudaEvent_t e : h->ev_pool)
        if (e) cudaEventDestroy(e);
      h->ev_pool.clear();
      if (h->tws) cudaFree(h->tws);
      h->tws = nullptr;
      h->tws_bytes = 0;
      h->train = false;
      return rc_init;
    }
  }
  h->train = on != 0;
  h->fwd_saved = false;
  return AIR_OK;
}

int32_t air_backward(air_handle* h, const float* params, const float* img, const float* eps_where,
                     const float* eps_what, const air_prior* prior, const air_outputs* outs, float baseline_mean,
                     float inv_batch, float l2_weight, float* grad_params, void* stream) {
  if (!h || !params || !img || !eps_where || !eps_what || !prior || !grad_params)
    return fail(AIR_ERR_ARG, "air_backward: NULL argument");
  if (!h->train || !h->fwd_saved)
    return fail(AIR_ERR_ARG, "air_backward: needs air_train_enable(h, 1) and a preceding air_forward on this handle with a "
                             "prior and a materialised canvas");
  const int32_t rc = check_outs(outs, true);
  if (rc != AIR_OK) return rc;
  if (!outs->canvas) return fail(AIR_ERR_ARG, "air_backward: the canvas output of the forward pass is required");
  return backward_impl(h, params, img, eps_where, eps_what, prior, outs, baseline_mean, inv_batch, l2_weight,
                       grad_params, (cudaStream_t)stream);
}

int32_t air_rmsprop_step(float* params, const float* grad, float* mg, float* ms, float* mom, int64_t n,
                         float learning_rate, float decay, float momentum, float epsilon, float grad_scale,
                         void* stream) {
  if (!params || !grad || !mg || !ms || !mom || n < 0) return fail(AIR_ERR_ARG, "air_rmsprop_step: bad argument");
  if (n == 0) return AIR_OK;
  air::rmsprop_centered_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      params, grad, mg, ms, mom, (size_t)n, learning_rate, decay, momentum, epsilon, grad_scale);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int64_t air_train_workspace_bytes(const air_handle* h) { return h ? (int64_t)h->tws_bytes : 0; }

int32_t air_cache_weights(air_handle* h, int32_t on) {
  if (!h) return fail(AIR_ERR_ARG, "air_cache_weights: NULL handle");
  h->cache_weights = on != 0;
  h->weights_ready = nullptr;
  h->hw0_ready = nullptr;
  return AIR_OK;
}
int32_t air_set_launch_overlap(air_handle* h, int32_t on) {
  if (!h) return fail(AIR_ERR_ARG, "air_set_launch_overlap: NULL handle");
  h->launch_overlap = on != 0;
  return AIR_OK;
}
namespace {
struct PriorTable { double v[AIR_MAX_STEPS + 1]; };
__global__ void prior_table_kernel(PriorTable t, double* dst) {
  if (threadIdx.x <= AIR_MAX_STEPS) dst[threadIdx.x] = t.v[threadIdx.x];
}
}   // namespace
int32_t air_prior_table_device(air_handle* h, const air_prior* prior, void* stream) {
  if (!h) return fail(AIR_ERR_ARG, "air_prior_table_device: NULL handle");
  if (!prior) {
    h->prior_dev_on = false;
    return AIR_OK;
  }
  PriorTable t;
  for (int k = 0; k <= AIR_MAX_STEPS; ++k)
    t.v[k] = k <= h->cfg.T ? (prior->steps_prob_is_f64 ? air::geom_prior_f64(prior->steps_success_prob, k)
                                                      : (double)air::geom_prior_f32((float)prior->steps_success_prob, k))
                           : 1.0;
  prior_table_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t, h->prior_dev);
  AIR_CUDA(cudaGetLastError());
  h->prior_dev_on = true;
  return AIR_OK;
}
int32_t air_params_updated(air_handle* h) {
  if (!h) return fail(AIR_ERR_ARG, "air_params_updated: NULL handle");
  h->weights_ready = nullptr;
  h->hw0_ready = nullptr;
  return AIR_OK;
}

int32_t air_linear_backward(const float* X, const float* W, const float* dY, const float* elu_x, float* dW, float* db,
                            float* dX, int32_t M, int32_t N, int32_t K, void* stream) {
  if (!X || !W || !dY || M < 1 || N < 1 || K < 1) return fail(AIR_ERR_ARG, "air_linear_backward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dW)
    AIR_CUDA(air::launch_gemm_simt(true, false, X, K, dY, N, dW, N, K, N, M, true, nullptr, 0, pick_split(K, N, M), st));
  if (db) AIR_CUDA(air::launch_colsum(dY, N, db, M, N, st));
  if (dX) AIR_CUDA(air::launch_gemm_simt(false, true, dY, N, W, N, dX, K, M, K, N, false, elu_x, K, 1, st));
  return AIR_OK;
}

int32_t air_baseline_grad(const float* target, const float* baseline, float target_mean, float inv_batch, float* d_baseline,
                          int32_t B, void* stream) {
  if (!target || !baseline || !d_baseline || B < 1) return fail(AIR_ERR_ARG, "air_baseline_grad: bad argument");
  (void)target;
  air::baseline_grad_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(baseline, target_mean, inv_batch,
                                                                                d_baseline, B, nullptr);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_baseline_grad_dev(const float* baseline, const float* target_mean_dev, float inv_batch, float* d_baseline,
                              int32_t B, void* stream) {
  if (!baseline || !target_mean_dev || !d_baseline || B < 1) return fail(AIR_ERR_ARG, "air_baseline_grad_dev: bad argument");
  air::baseline_grad_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(baseline, 0.f, inv_batch, d_baseline, B,
                                                                                target_mean_dev);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

// ---- BaselineMLP on the engine ------------------------------------------------------------------------------------
int32_t air_baseline_attach(air_handle* h, int32_t n_hidden, const int32_t* hidden) {
  if (!h || !hidden || n_hidden < 1 || n_hidden > AIR_MAX_HIDDEN) return fail(AIR_ERR_ARG, "air_baseline_attach: bad argument");
  if (h->bl.attached) return fail(AIR_ERR_ARG, "air_baseline_attach: already attached");
  if (h->tws) return fail(AIR_ERR_ARG, "air_baseline_attach: attach before air_train_enable (the training workspace is sized "
                                       "for the widest layer)");
  const air_config& c = h->cfg;
  air_handle::Baseline& bl = h->bl;
  bl.n_in = h->P + c.T * (c.na + 4 + 1) + 2 * c.nh;   // concat[img, what, where, presence, h, c] (modules.py:131-141)
  int d = bl.n_in;
  int64_t off = 0;
  int wmax = 1;
  for (int i = 0; i <= n_hidden; ++i) {
    Layer l;
    l.K = d;
    l.N = i < n_hidden ? hidden[i] : 1;
    if (l.N < 1) return fail(AIR_ERR_ARG, "air_baseline_attach: hidden widths must be positive");
    l.w_off = off;
    off += (int64_t)l.K * l.N;
    l.b_off = off;
    off += l.N;
    bl.mlp.layers.push_back(l);
    d = l.N;
    wmax = std::max(wmax, l.N);
  }
  bl.mlp.n_hidden = n_hidden;
  bl.n_params = off;
  const int B = c.B, B_alloc = round_up(B, air::tc::BM);
  Layer& l0 = bl.mlp.layers[0];
  const int Kpad = round_up(l0.K, air::tc::BK), N_alloc = round_up(l0.N, 64);
  for (int pass = 0; pass < 2; ++pass) {
    Carver cv(pass == 0 ? nullptr : bl.ws);
    bl.x.f32 = cv.take<float>((size_t)B * bl.n_in);
    bl.x.ld = bl.n_in;
    bl.act.clear();
    for (int i = 0; i < n_hidden; ++i) bl.act.push_back(cv.take<float>((size_t)B * bl.mlp.layers[i].N));
    bl.g[0] = cv.take<float>((size_t)B * wmax);
    bl.g[1] = cv.take<float>((size_t)B * wmax);
    if (h->use_tc) {
      bl.x.rows_alloc = B_alloc;
      bl.x.kpad = Kpad;
      bl.x.hl = cv.take<__half>(2 * (size_t)B_alloc * Kpad);
      bl.w_arena = cv.take<__half>(2 * (size_t)N_alloc * Kpad);
      bl.prep_table = cv.take<air::tc::PrepEntry>(1);
    }
    if (pass == 0) {
      AIR_CUDA(cudaMalloc(&bl.ws, cv.off));
      AIR_CUDA(cudaMemset(bl.ws, 0, cv.off));   // K / N padding of the operand planes stays zero
    }
  }
  if (h->use_tc) {
    TcWeight w;
    w.src_off = l0.w_off;
    w.K = l0.K;
    w.N = l0.N;
    w.Kpad = Kpad;
    w.BN = 64;
    w.N_alloc = N_alloc;
    w.bias_src = -1;
    if (!air::tc::make_tmap(&w.tm, bl.w_arena, Kpad, 2 * (int64_t)N_alloc, 64))
      return fail(AIR_ERR_CUDA, "air_baseline_attach: cuTensorMapEncodeTiled failed");
    air::tc::PrepEntry pe;
    memset(&pe, 0, sizeof(pe));
    pe.src_off = l0.w_off;
    pe.dst_off = 0;
    pe.plane = (int64_t)N_alloc * Kpad;
    pe.K = l0.K;
    pe.N = l0.N;
    pe.Kpad = Kpad;
    pe.tile_begin = 0;
    pe.tiles_n = (l0.N + 31) / 32;
    pe.bias_src = -1;
    bl.prep_tiles = pe.tiles_n * ((l0.K + 31) / 32);
    AIR_CUDA(cudaMemcpy(bl.prep_table, &pe, sizeof(pe), cudaMemcpyHostToDevice));
    l0.tc = (int)h->tcw.size();
    h->tcw.push_back(w);
  }
  bl.attached = true;
  return AIR_OK;
}

int64_t air_baseline_param_count(const air_handle* h) { return (h && h->bl.attached) ? h->bl.n_params : 0; }
int32_t air_baseline_input_width(const air_handle* h) { return (h && h->bl.attached) ? h->bl.n_in : 0; }

int32_t air_baseline_forward(air_handle* h, const float* bparams, const float* img, const air_outputs* o, float* baseline,
                             void* stream) {
  if (!h || !bparams || !img || !o || !baseline) return fail(AIR_ERR_ARG, "air_baseline_forward: NULL argument");
  if (!h->bl.attached) return fail(AIR_ERR_ARG, "air_baseline_forward: air_baseline_attach first");
  if (!o->what || !o->where || !o->presence || !o->final_h || !o->final_c)
    return fail(AIR_ERR_ARG, "air_baseline_forward: outs needs what / where / presence / final_h / final_c");
  cudaStream_t st = (cudaStream_t)stream;
  const air_config& c = h->cfg;
  air_handle::Baseline& bl = h->bl;
  const int B = c.B;
  const size_t n = (size_t)B * bl.n_in;
  AIR_CUDA(air::launch_k(air::baseline_gather_kernel, dim3((unsigned)B), dim3(256), 0, st, img,
                         (const float*)o->what, (const float*)o->where, (const float*)o->presence, (const float*)o->final_h,
                         (const float*)o->final_c, bl.x.f32, bl.x.hl, bl.x.plane(), bl.x.kpad, B, c.T, h->P, c.na, c.nh, bl.n_in));
  ++h->launches;
  const int nl = (int)bl.mlp.layers.size();
  for (int i = 0; i < nl; ++i) {
    const Layer& l = bl.mlp.layers[i];
    const bool last = i == nl - 1;
    float* dst = last ? baseline : bl.act[i];
    const int act = last ? air::ACT_NONE : air::ACT_ELU;
    if (i == 0 && h->use_tc) {
      AIR_CUDA(air::launch_k(air::tc::prep_weights_kernel, dim3(bl.prep_tiles), dim3(256), 0, st, bparams, bl.w_arena,
                             bl.prep_table, 1, h->range_flag, (float*)nullptr));
      ++h->launches;
      Buf out;
      out.f32 = dst;
      out.ld = l.N;
      const int32_t rc = dense(h, bparams, bl.x, 0, l, true, nullptr, 0, out, true, false, B, act, st);
      if (rc != AIR_OK) return rc;
    } else {
      const float* A = i == 0 ? bl.x.f32 : bl.act[i - 1];
      AIR_CUDA(air::launch_linear_simt(A, l.K, bparams + l.w_off, l.N, bparams + l.b_off, nullptr, 0, dst, l.N, B, l.N, l.K,
                                       act, st));
      ++h->launches;
    }
  }
  return AIR_OK;
}

namespace {
int32_t baseline_backward_impl(air_handle* h, const float* bparams, const float* d_baseline, float* bgrad, void* stream,
                               bool join) {
  if (!h || !bparams || !d_baseline || !bgrad) return fail(AIR_ERR_ARG, "air_baseline_backward: NULL argument");
  if (!h->bl.attached) return fail(AIR_ERR_ARG, "air_baseline_backward: air_baseline_attach first");
  cudaStream_t st = (cudaStream_t)stream;
  air_handle::Baseline& bl = h->bl;
  const int B = h->cfg.B;
  AIR_CUDA(cudaMemsetAsync(bgrad, 0, sizeof(float) * (size_t)bl.n_params, st));
  const int nl = (int)bl.mlp.layers.size();
  const float* cur = d_baseline;
  int ld_cur = 1;
  int32_t rc;
  for (int i = nl - 1; i >= 0; --i) {
    const Layer& l = bl.mlp.layers[i];
    const float* X = i == 0 ? bl.x.f32 : bl.act[i - 1];
    // dW += X^T dY, db += colsum(dY): the tensor-core split-K GEMM on a side stream when the handle trains on it
    if ((rc = layer_param_grads(h, bgrad, l, X, l.K, cur, ld_cur, B, st)) != AIR_OK) return rc;
    if (i > 0) {
      float* dst = (cur == bl.g[0]) ? bl.g[1] : bl.g[0];
      if ((rc = wait_consumed(h, dst, st)) != AIR_OK) return rc;
      AIR_CUDA(air::launch_gemm_simt(false, true, cur, ld_cur, bparams + l.w_off, l.N, dst, l.K, B, l.K, l.N, false, X, l.K, 1,
                                     st));   // (X is the ELU output of the layer below: the mask is elu'(X))
      ++h->launches;
      cur = dst;
      ld_cur = l.K;
    }
  }
  return join ? join_side_streams(h, st) : AIR_OK;
}
}   // namespace
int32_t air_baseline_backward(air_handle* h, const float* bparams, const float* d_baseline, float* bgrad, void* stream) {
  return baseline_backward_impl(h, bparams, d_baseline, bgrad, stream, true);
}
// The weight-gradient GEMMs stay on the handle's side streams without being joined: they overlap the air_backward that
// follows on the same stream, whose final join covers them (the big one, [B, 3177]^T [B, 256], is 50 MB of operand traffic
// that nothing on the main gradient path waits for).
int32_t air_baseline_backward_async(air_handle* h, const float* bparams, const float* d_baseline, float* bgrad,
                                    void* stream) {
  return baseline_backward_impl(h, bparams, d_baseline, bgrad, stream, false);
}

int32_t air_forward(air_handle* h, const float* params, const float* img, const float* eps_where,
                    const float* eps_what, const float* u_pres, const float* baseline, const air_prior* prior,
                    const air_outputs* outs, void* stream) {
  if (!h || !params || !img || !eps_where || !eps_what) return fail(AIR_ERR_ARG, "air_forward: NULL argument");
  if (h->cfg.discrete_steps && !u_pres) return fail(AIR_ERR_ARG, "air_forward: u_pres is required for discrete steps");
  const int32_t rc = check_outs(outs, prior != nullptr);
  if (rc != AIR_OK) return rc;
  return forward_impl(h, params, img, eps_where, eps_what, u_pres, baseline, prior, outs, h->cfg.T, nullptr, nullptr,
                      nullptr, nullptr, nullptr, h->cfg.output_multiplier, (cudaStream_t)stream);
}

int32_t air_forward_host(air_handle* h, const float* params, const float* img_host, const float* eps_where_host,
                         const float* eps_what_host, const float* u_pres_host, const air_prior* prior,
                         const air_outputs* outs, float* scalars_host, float* loss_per_sample_host, void* stream) {
  const air::PdlScope pdl_scope(h ? h->launch_overlap : true);
  if (!h || !params || !img_host || !eps_where_host || !eps_what_host || !u_pres_host)
    return fail(AIR_ERR_ARG, "air_forward_host: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const air_config& c = h->cfg;
  const size_t TB = (size_t)c.T * c.B;
  AIR_CUDA(cudaMemcpyAsync(h->st_img, img_host, sizeof(float) * c.B * h->P, cudaMemcpyHostToDevice, st));
  AIR_CUDA(cudaMemcpyAsync(h->st_eps_where, eps_where_host, sizeof(float) * TB * 4, cudaMemcpyHostToDevice, st));
  AIR_CUDA(cudaMemcpyAsync(h->st_eps_what, eps_what_host, sizeof(float) * TB * c.na, cudaMemcpyHostToDevice, st));
  AIR_CUDA(cudaMemcpyAsync(h->st_u, u_pres_host, sizeof(float) * TB, cudaMemcpyHostToDevice, st));
  const int32_t rc = air_forward(h, params, h->st_img, h->st_eps_where, h->st_eps_what, h->st_u, nullptr, prior, outs,
                                 stream);
  if (rc != AIR_OK) return rc;
  if (prior && scalars_host)
    AIR_CUDA(cudaMemcpyAsync(scalars_host, outs->scalars, sizeof(float) * AIR_N_SCALARS, cudaMemcpyDeviceToHost, st));
  if (prior && loss_per_sample_host)
    AIR_CUDA(cudaMemcpyAsync(loss_per_sample_host, outs->loss_per_sample, sizeof(float) * c.B, cudaMemcpyDeviceToHost,
                             st));
  AIR_CUDA(cudaStreamSynchronize(st));
  return AIR_OK;
}

int32_t air_forward_host_u8(air_handle* h, const float* params, const uint8_t* img_u8_host,
                            const float* eps_where_host, const float* eps_what_host, const float* u_pres_host,
                            const air_prior* prior, const air_outputs* outs, float* scalars_host,
                            float* loss_per_sample_host, void* stream) {
  const air::PdlScope pdl_scope(h ? h->launch_overlap : true);
  if (!h || !params || !img_u8_host || !eps_where_host || !eps_what_host || !u_pres_host)
    return fail(AIR_ERR_ARG, "air_forward_host_u8: NULL argument");
  int32_t rc = check_outs(outs, prior != nullptr);
  if (rc != AIR_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const air_config& c = h->cfg;
  const size_t TB = (size_t)c.T * c.B;
  AIR_CUDA(cudaMemcpyAsync(h->st_img_u8, img_u8_host, (size_t)c.B * h->P, cudaMemcpyHostToDevice, st));
  AIR_CUDA(cudaMemcpyAsync(h->st_eps_where, eps_where_host, sizeof(float) * TB * 4, cudaMemcpyHostToDevice, st));
  AIR_CUDA(cudaMemcpyAsync(h->st_eps_what, eps_what_host, sizeof(float) * TB * c.na, cudaMemcpyHostToDevice, st));
  AIR_CUDA(cudaMemcpyAsync(h->st_u, u_pres_host, sizeof(float) * TB, cudaMemcpyHostToDevice, st));
  // uint8 -> float32 / 255 (data.py:116) on the device, fused with the first layer's operand split
  const size_t n4 = (size_t)c.B * ((h->P + 3) / 4);
  AIR_CUDA(air::launch_k(air::tc::u8_to_f32_hl_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st,
                         (const uint8_t*)h->st_img_u8, h->st_img, (h->use_tc && !enc1_active(h)) ? h->x.hl : (__half*)nullptr, h->x.plane(),
                         h->x.kpad, c.B, h->P, (const int32_t*)nullptr, 0LL));
  ++h->launches;
  rc = forward_impl(h, params, h->st_img, h->st_eps_where, h->st_eps_what, h->st_u, nullptr, prior, outs, c.T, nullptr,
                    nullptr, nullptr, nullptr, nullptr, c.output_multiplier, st, /*x_hl_ready=*/h->use_tc);
  if (rc != AIR_OK) return rc;
  if (prior && scalars_host)
    AIR_CUDA(cudaMemcpyAsync(scalars_host, outs->scalars, sizeof(float) * AIR_N_SCALARS, cudaMemcpyDeviceToHost, st));
  if (prior && loss_per_sample_host)
    AIR_CUDA(cudaMemcpyAsync(loss_per_sample_host, outs->loss_per_sample, sizeof(float) * c.B, cudaMemcpyDeviceToHost,
                             st));
  AIR_CUDA(cudaStreamSynchronize(st));
  return AIR_OK;
}

namespace {
int32_t draw_noise_impl(air_handle* h, uint64_t seed, float* eps_where, float* eps_what, float* u_pres, cudaStream_t st) {
  const air_config& c = h->cfg;
  const size_t TB = (size_t)c.T * c.B;
  struct { float* p; size_t n; uint32_t id; int normal; } jobs[3] = {
      {eps_where, TB * 4, 0u, 1}, {eps_what, TB * (size_t)c.na, 1u, 1}, {u_pres, TB, 2u, 0}};
  for (auto& j : jobs) {
    if (!j.p) continue;
    const size_t threads = (j.n + 3) / 4;
    AIR_CUDA(air::launch_k(air::philox_fill_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, j.p, j.n,
                           (unsigned long long)seed, j.id, j.normal));
    ++h->launches;
  }
  return AIR_OK;
}
}  // namespace

int32_t air_draw_noise(air_handle* h, uint64_t seed, float* eps_where, float* eps_what, float* u_pres, void* stream) {
  if (!h) return fail(AIR_ERR_ARG, "air_draw_noise: NULL handle");
  return draw_noise_impl(h, seed, eps_where, eps_what, u_pres, (cudaStream_t)stream);
}

int32_t air_forward_host_u8_rng(air_handle* h, const float* params, const uint8_t* img_u8_host, uint64_t seed,
                                const air_prior* prior, const air_outputs* outs, float* scalars_host,
                                float* loss_per_sample_host, void* stream) {
  const air::PdlScope pdl_scope(h ? h->launch_overlap : true);
  if (!h || !params || !img_u8_host) return fail(AIR_ERR_ARG, "air_forward_host_u8_rng: NULL argument");
  int32_t rc = check_outs(outs, prior != nullptr);
  if (rc != AIR_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const air_config& c = h->cfg;
  AIR_CUDA(cudaMemcpyAsync(h->st_img_u8, img_u8_host, (size_t)c.B * h->P, cudaMemcpyHostToDevice, st));
  if ((rc = draw_noise_impl(h, seed, h->st_eps_where, h->st_eps_what, h->st_u, st)) != AIR_OK) return rc;
  const size_t n4 = (size_t)c.B * ((h->P + 3) / 4);
  AIR_CUDA(air::launch_k(air::tc::u8_to_f32_hl_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st,
                         (const uint8_t*)h->st_img_u8, h->st_img, (h->use_tc && !enc1_active(h)) ? h->x.hl : (__half*)nullptr, h->x.plane(),
                         h->x.kpad, c.B, h->P, (const int32_t*)nullptr, 0LL));
  ++h->launches;
  rc = forward_impl(h, params, h->st_img, h->st_eps_where, h->st_eps_what, h->st_u, nullptr, prior, outs, c.T, nullptr,
                    nullptr, nullptr, nullptr, nullptr, c.output_multiplier, st, /*x_hl_ready=*/h->use_tc);
  if (rc != AIR_OK) return rc;
  if (prior && scalars_host)
    AIR_CUDA(cudaMemcpyAsync(scalars_host, outs->scalars, sizeof(float) * AIR_N_SCALARS, cudaMemcpyDeviceToHost, st));
  if (prior && loss_per_sample_host)
    AIR_CUDA(cudaMemcpyAsync(loss_per_sample_host, outs->loss_per_sample, sizeof(float) * c.B, cudaMemcpyDeviceToHost,
                             st));
  AIR_CUDA(cudaStreamSynchronize(st));
  return AIR_OK;
}

// ---- double-buffered host feed -------------------------------------------------------------------------------------
namespace {
int32_t feed_init(air_handle* h) {
  if (h->feed_stream) return AIR_OK;
  const size_t bytes = align_up((size_t)h->cfg.B * h->P, 256);
  uint8_t* buf = nullptr;
  AIR_CUDA(cudaMalloc(&buf, 2 * bytes));
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
    cudaFree(buf);
    return fail(AIR_ERR_CUDA, "air_feed_host_u8: cudaStreamCreate failed");
  }
  for (int i = 0; i < 2; ++i) {
    AIR_CUDA(cudaEventCreateWithFlags(&h->feed_fed[i], cudaEventDisableTiming));
    AIR_CUDA(cudaEventCreateWithFlags(&h->feed_consumed[i], cudaEventDisableTiming));
    AIR_CUDA(cudaEventCreateWithFlags(&h->feed_done[i], cudaEventDisableTiming));
  }
  h->feed_buf[0] = buf;
  h->feed_buf[1] = buf + bytes;
  h->feed_stream = s;
  return AIR_OK;
}
}  // namespace

int32_t air_feed_host_u8(air_handle* h, int32_t slot, const uint8_t* img_u8_host) {
  if (!h || !img_u8_host || slot < 0 || slot > 1) return fail(AIR_ERR_ARG, "air_feed_host_u8: bad argument");
  int32_t rc = feed_init(h);
  if (rc != AIR_OK) return rc;
  // the pass that last read this slot must have converted it (an event that was never recorded is complete)
  AIR_CUDA(cudaStreamWaitEvent(h->feed_stream, h->feed_consumed[slot], 0));
  AIR_CUDA(cudaMemcpyAsync(h->feed_buf[slot], img_u8_host, (size_t)h->cfg.B * h->P, cudaMemcpyHostToDevice,
                           h->feed_stream));
  AIR_CUDA(cudaEventRecord(h->feed_fed[slot], h->feed_stream));
  h->feed_pending[slot] = true;
  return AIR_OK;
}

int32_t air_forward_fed_u8_rng(air_handle* h, const float* params, int32_t slot, uint64_t seed, const air_prior* prior,
                               const air_outputs* outs, float* scalars_host, float* loss_per_sample_host, void* stream) {
  const air::PdlScope pdl_scope(h ? h->launch_overlap : true);
  if (!h || !params || slot < 0 || slot > 1) return fail(AIR_ERR_ARG, "air_forward_fed_u8_rng: bad argument");
  if (!h->feed_stream || !h->feed_pending[slot])
    return fail(AIR_ERR_ARG, "air_forward_fed_u8_rng: nothing was fed into this slot (call air_feed_host_u8 first)");
  int32_t rc = check_outs(outs, prior != nullptr);
  if (rc != AIR_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const air_config& c = h->cfg;
  if ((rc = draw_noise_impl(h, seed, h->st_eps_where, h->st_eps_what, h->st_u, st)) != AIR_OK) return rc;
  AIR_CUDA(cudaStreamWaitEvent(st, h->feed_fed[slot], 0));
  const size_t n4 = (size_t)c.B * ((h->P + 3) / 4);
  AIR_CUDA(air::launch_k(air::tc::u8_to_f32_hl_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st,
                         (const uint8_t*)h->feed_buf[slot], h->st_img, (h->use_tc && !enc1_active(h)) ? h->x.hl : (__half*)nullptr,
                         h->x.plane(), h->x.kpad, c.B, h->P, (const int32_t*)nullptr, 0LL));
  ++h->launches;
  AIR_CUDA(cudaEventRecord(h->feed_consumed[slot], st));
  h->feed_pending[slot] = false;
  rc = forward_impl(h, params, h->st_img, h->st_eps_where, h->st_eps_what, h->st_u, nullptr, prior, outs, c.T, nullptr,
                    nullptr, nullptr, nullptr, nullptr, c.output_multiplier, st, /*x_hl_ready=*/h->use_tc);
  if (rc != AIR_OK) return rc;
  if (prior && scalars_host)
    AIR_CUDA(cudaMemcpyAsync(scalars_host, outs->scalars, sizeof(float) * AIR_N_SCALARS, cudaMemcpyDeviceToHost, st));
  if (prior && loss_per_sample_host)
    AIR_CUDA(cudaMemcpyAsync(loss_per_sample_host, outs->loss_per_sample, sizeof(float) * c.B, cudaMemcpyDeviceToHost,
                             st));
  AIR_CUDA(cudaEventRecord(h->feed_done[slot], st));
  return AIR_OK;
}

int32_t air_feed_wait(air_handle* h, int32_t slot) {
  if (!h || slot < 0 || slot > 1) return fail(AIR_ERR_ARG, "air_feed_wait: bad argument");
  if (!h->feed_stream) return fail(AIR_ERR_ARG, "air_feed_wait: the feed was never used on this handle");
  AIR_CUDA(cudaEventSynchronize(h->feed_done[slot]));
  return AIR_OK;
}

int32_t air_forward_dataset_u8(air_handle* h, const float* params, const uint8_t* dataset_u8, int64_t n_dataset,
                               const int32_t* idx, const float* eps_where, const float* eps_what, const float* u_pres,
                               const float* baseline, const air_prior* prior, const air_outputs* outs, float* img_out,
                               void* stream) {
  const air::PdlScope pdl_scope(h ? h->launch_overlap : true);
  if (!h || !params || !dataset_u8 || !idx || !eps_where || !eps_what || n_dataset < 1)
    return fail(AIR_ERR_ARG, "air_forward_dataset_u8: bad argument");
  if (h->cfg.discrete_steps && !u_pres) return fail(AIR_ERR_ARG, "air_forward_dataset_u8: u_pres is required");
  int32_t rc = check_outs(outs, prior != nullptr);
  if (rc != AIR_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const air_config& c = h->cfg;
  float* img = img_out ? img_out : h->st_img;
  // gather + uint8 -> float32 / 255 (data.py:116,131-132) + the first layer's operand split, one pass, all on the device
  const size_t n4 = (size_t)c.B * ((h->P + 3) / 4);
  AIR_CUDA(air::launch_k(air::tc::u8_to_f32_hl_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, dataset_u8,
                         img, (h->use_tc && !enc1_active(h)) ? h->x.hl : (__half*)nullptr, h->x.plane(), h->x.kpad, c.B, h->P, idx,
                         (long long)n_dataset));   // indices are clamped to [0, n_dataset)
  ++h->launches;
  return forward_impl(h, params, img, eps_where, eps_what, u_pres, baseline, prior, outs, c.T, nullptr, nullptr, nullptr,
                      nullptr, nullptr, c.output_multiplier, st, /*x_hl_ready=*/h->use_tc);
}

int32_t air_gather_u8(const uint8_t* dataset_u8, const int32_t* idx, float* img_out, int32_t B, int32_t P, void* stream) {
  if (!dataset_u8 || !idx || !img_out || B < 1 || P < 1) return fail(AIR_ERR_ARG, "air_gather_u8: bad argument");
  const size_t n4 = (size_t)B * ((P + 3) / 4);
  air::tc::u8_to_f32_hl_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      dataset_u8, img_out, nullptr, 0, 0, B, P, idx);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_elbo_scalars(air_handle* h, const float* baseline, const air_prior* prior, const air_outputs* outs,
                         void* stream) {
  if (!h || !prior) return fail(AIR_ERR_ARG, "air_elbo_scalars: NULL argument");
  const int32_t rc = check_outs(outs, true);
  if (rc != AIR_OK) return rc;
  air::elbo_scalars_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
      outs->rec_loss_per_sample, outs->kl_num_steps_per_sample, outs->kl_what_per_sample, outs->kl_where_per_sample,
      outs->num_step_per_sample, outs->num_steps_log_prob, baseline, outs->scalars, h->cfg.B, *prior, nullptr, nullptr);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_elbo_scalars_raw(int32_t B, const float* baseline, const air_prior* prior, const air_outputs* outs,
                             void* stream) {
  if (B < 1 || !prior) return fail(AIR_ERR_ARG, "air_elbo_scalars_raw: bad argument");
  const int32_t rc = check_outs_elbo_only(outs);
  if (rc != AIR_OK) return rc;
  AIR_CUDA(air::launch_k(air::elbo_scalars_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream,
                         outs->rec_loss_per_sample, outs->kl_num_steps_per_sample, outs->kl_what_per_sample,
                         outs->kl_where_per_sample, outs->num_step_per_sample, outs->num_steps_log_prob, baseline,
                         outs->scalars, B, *prior, (const float*)nullptr, (float*)nullptr));
  return AIR_OK;
}

int32_t air_prior_terms(int32_t B, int32_t T, int32_t na, const float* what_loc, const float* what_scale,
                        const float* where_loc, const float* where_scale, const float* presence_prob,
                        const float* presence, const air_prior* prior, const air_outputs* outs, void* stream) {
  if (B < 1 || T < 1 || T > AIR_MAX_STEPS || na < 1 || !what_loc || !what_scale || !where_loc || !where_scale ||
      !presence_prob || !presence || !prior || !outs)
    return fail(AIR_ERR_ARG, "air_prior_terms: bad argument");
  if (!outs->num_steps_posterior || !outs->num_step_per_sample || !outs->prior_step_weight ||
      !outs->rec_loss_per_sample || !outs->kl_num_steps_per_sample || !outs->kl_what_per_sample ||
      !outs->kl_where_per_sample || !outs->loss_per_sample || !outs->num_steps_log_prob)
    return fail(AIR_ERR_ARG, "air_prior_terms: ELBO output buffers are mandatory");
  cudaStream_t st = (cudaStream_t)stream;
  air::ElboArgs a;
  memset(&a, 0, sizeof(a));
  a.where_loc = where_loc;
  a.where_scale = where_scale;
  a.what_loc = what_loc;
  a.what_scale = what_scale;
  a.presence = presence;
  a.presence_prob = presence_prob;
  a.num_steps_posterior = outs->num_steps_posterior;
  a.num_step_per_sample = outs->num_step_per_sample;
  a.prior_step_weight = outs->prior_step_weight;
  a.rec_loss_per_sample = outs->rec_loss_per_sample;
  a.kl_num_steps_per_sample = outs->kl_num_steps_per_sample;
  a.kl_what_per_sample = outs->kl_what_per_sample;
  a.kl_where_per_sample = outs->kl_where_per_sample;
  a.loss_per_sample = outs->loss_per_sample;
  a.num_steps_log_prob = outs->num_steps_log_prob;
  a.T = T; a.B = B; a.na = na;
  a.do_elbo = 1;
  a.prior = *prior;
  for (int k = 0; k <= T; ++k)
    a.steps_prior[k] = prior->steps_prob_is_f64 ? air::geom_prior_f64(prior->steps_success_prob, k)
                                                : (double)air::geom_prior_f32((float)prior->steps_success_prob, k);
  // no canvas: the reconstruction term is exactly 0 and prior_terms_kernel completes loss_per_sample itself
  AIR_CUDA(air::launch_prior_terms(a, /*finalize=*/1, (cudaStream_t)stream));
  return AIR_OK;
}

int32_t air_iwae_bound(int32_t n_canvases, int32_t K, int32_t T, int32_t na, const float* what, const float* what_loc,
                       const float* what_scale, const float* where, const float* where_loc, const float* where_scale,
                       const float* presence, const float* rec_loss_per_row, const float* num_steps_log_prob_per_row,
                       const air_prior* prior, float* log_w, float* bound_per_canvas, float* bound_mean, void* stream) {
  if (n_canvases < 1 || K < 1 || T < 1 || T > AIR_MAX_STEPS || na < 1 || !what || !what_loc || !what_scale || !where ||
      !where_loc || !where_scale || !presence || !rec_loss_per_row || !num_steps_log_prob_per_row || !prior || !log_w ||
      !bound_per_canvas)
    return fail(AIR_ERR_ARG, "air_iwae_bound: bad argument");
  if (!prior->where_shift_has_loc)
    return fail(AIR_ERR_ARG, "air_iwae_bound: the where-shift prior needs a location (a density, not a KL, is evaluated)");
  air::IwaeArgs a;
  memset(&a, 0, sizeof(a));
  a.what = what; a.what_loc = what_loc; a.what_scale = what_scale;
  a.where = where; a.where_loc = where_loc; a.where_scale = where_scale;
  a.presence = presence;
  a.rec_loss = rec_loss_per_row;
  a.log_q_n = num_steps_log_prob_per_row;
  a.log_w = log_w;
  a.T = T;
  a.R = n_canvases * K;
  a.na = na;
  a.prior = *prior;
  for (int k = 0; k <= T; ++k)
    a.steps_prior[k] = prior->steps_prob_is_f64 ? air::geom_prior_f64(prior->steps_success_prob, k)
                                                : (double)air::geom_prior_f32((float)prior->steps_success_prob, k);
  cudaStream_t st = (cudaStream_t)stream;
  air::iwae_logw_kernel<<<(a.R + 3) / 4, 128, 0, st>>>(a);
  AIR_CUDA(cudaGetLastError());
  air::iwae_reduce_kernel<<<1, 1024, 0, st>>>(log_w, bound_per_canvas, bound_mean, n_canvases, K);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_cell_step(air_handle* h, const float* params, const float* img, float* canvas, float* hstate,
                      float* cstate, float* presence, const float* eps_where, const float* eps_what,
                      const float* u_pres, float* out_glimpse, float* out_what, float* out_what_loc,
                      float* out_what_scale, float* out_where, float* out_where_loc, float* out_where_scale,
                      float* out_presence_prob, void* stream) {
  if (!h || !params || !img || !canvas || !hstate || !cstate || !presence || !eps_where || !eps_what || !out_glimpse ||
      !out_what || !out_what_loc || !out_what_scale || !out_where || !out_where_loc || !out_where_scale ||
      !out_presence_prob)
    return fail(AIR_ERR_ARG, "air_cell_step: NULL argument");
  if (h->cfg.discrete_steps && !u_pres) return fail(AIR_ERR_ARG, "air_cell_step: u_pres is required");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = h->cfg.B;
  // presence is both carried-in state and output: snapshot the incoming value
  AIR_CUDA(cudaMemcpyAsync(h->st_pres_in, presence, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  air_outputs o;
  memset(&o, 0, sizeof(o));
  o.glimpse = out_glimpse;
  o.what = out_what;
  o.what_loc = out_what_loc;
  o.what_scale = out_what_scale;
  o.where = out_where;
  o.where_loc = out_where_loc;
  o.where_scale = out_where_scale;
  o.presence_prob = out_presence_prob;
  o.presence = presence;
  o.final_h = hstate;
  o.final_c = cstate;
  return forward_impl(h, params, img, eps_where, eps_what, u_pres, nullptr, nullptr, &o, 1, hstate, cstate,
                      h->st_pres_in, canvas, canvas, 1.0f, st);
}

int32_t air_linear(const float* A, const float* Wt, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                   int32_t act, int32_t precision, void* stream) {
  if (!A || !Wt || !out || M < 1 || N < 1 || K < 1) return fail(AIR_ERR_ARG, "air_linear: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == AIR_PREC_FP32) {
    AIR_CUDA(air::launch_linear_simt(A, K, Wt, N, bias, nullptr, 0, out, N, M, N, K, act, st));
    return AIR_OK;
  }
  if (precision != AIR_PREC_TC_SPLIT) return fail(AIR_ERR_ARG, "air_linear: unknown precision");
  // stand-alone tensor-core layer: build the hl operands in a temporary arena, run one GEMM, free
  namespace tc = air::tc;
  if (!tc::get_encode_fn()) return fail(AIR_ERR_CUDA, "air_linear: cuTensorMapEncodeTiled is not available");
  const int Kpad = round_up(K, tc::BK), BN = N <= 32 ? 32 : 64, N_alloc = round_up(N, BN), M_alloc = round_up(M, tc::BM);
  const size_t a_halves = 2 * (size_t)M_alloc * Kpad, w_halves = 2 * (size_t)N_alloc * Kpad;
  const size_t bytes = align_up(a_halves * 2, 1024) + align_up(w_halves * 2, 1024) + 1024;
  char* tmp = nullptr;
  AIR_CUDA(cudaMallocAsync(&tmp, bytes, st));
  AIR_CUDA(cudaMemsetAsync(tmp, 0, bytes, st));
  __half* a_hl = reinterpret_cast<__half*>(tmp);
  __half* w_hl = reinterpret_cast<__half*>(tmp + align_up(a_halves * 2, 1024));
  tc::PrepEntry* table = reinterpret_cast<tc::PrepEntry*>(tmp + align_up(a_halves * 2, 1024) + align_up(w_halves * 2, 1024));
  int* flag = reinterpret_cast<int*>(table + 1);
  tc::PrepEntry pe;
  pe.src_off = 0;
  pe.dst_off = 0;
  pe.plane = (int64_t)N_alloc * Kpad;
  pe.K = K;
  pe.N = N;
  pe.Kpad = Kpad;
  pe.tile_begin = 0;
  pe.tiles_n = (N + 31) / 32;
  pe.split_n = 0;
  pe.split_off = 0;
  pe.bias_src = -1;
  pe.bias_dst = 0;
  pe.perm_nh = 0;
  AIR_CUDA(cudaMemcpyAsync(table, &pe, sizeof(pe), cudaMemcpyHostToDevice, st));
  tc::prep_weights_kernel<<<pe.tiles_n * ((K + 31) / 32), 256, 0, st>>>(Wt, w_hl, table, 1, flag, nullptr);
  AIR_CUDA(cudaGetLastError());
  const size_t n4 = (size_t)M * ((K + 3) / 4);
  tc::split_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(A, K, a_hl, (size_t)M_alloc * Kpad, Kpad, M, K, flag);
  AIR_CUDA(cudaGetLastError());
  CUtensorMap tm_a, tm_b;
  if (!tc::make_tmap(&tm_a, a_hl, Kpad, 2 * (int64_t)M_alloc, tc::BM) ||
      !tc::make_tmap(&tm_b, w_hl, Kpad, 2 * (int64_t)N_alloc, BN))
    return fail(AIR_ERR_CUDA, "air_linear: cuTensorMapEncodeTiled failed");
  tc::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias;
  p.out_f32 = out;
  p.ldc = N;
  p.M = M;
  p.N = N;
  p.num_k_blocks = Kpad / tc::BK;
  p.a_lo_row = M_alloc;
  p.b_lo_row = N_alloc;
  p.act = act;
  p.range_flag = flag;
  AIR_CUDA(tc::launch_gemm(BN, tm_a, tm_b, p, N_alloc, st));
  int host_flag = 0;
  AIR_CUDA(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  AIR_CUDA(cudaStreamSynchronize(st));
  AIR_CUDA(cudaFreeAsync(tmp, st));
  if (host_flag) return fail(AIR_ERR_RANGE, "air_linear: operand outside the fp16 range of the split engine");
  return AIR_OK;
}

int32_t air_lstm_step(const float* x, float* hstate, float* cstate, const float* W, const float* b, int32_t B,
                      int32_t nx, int32_t nh, float forget_bias, void* stream) {
  if (!x || !hstate || !cstate || !W || !b || B < 1 || nx < 1 || nh < 1)
    return fail(AIR_ERR_ARG, "air_lstm_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* gates = nullptr;
  AIR_CUDA(cudaMallocAsync(&gates, sizeof(float) * (size_t)B * 4 * nh * 2, st));
  float* gx = gates + (size_t)B * 4 * nh;
  AIR_CUDA(air::launch_linear_simt(x, nx, W, 4 * nh, b, nullptr, 0, gx, 4 * nh, B, 4 * nh, nx, air::ACT_NONE, st));
  AIR_CUDA(air::launch_linear_simt(hstate, nh, W + (size_t)nx * 4 * nh, 4 * nh, nullptr, gx, 4 * nh, gates, 4 * nh, B,
                                   4 * nh, nh, air::ACT_NONE, st));
  air::lstm_pointwise_kernel<<<(B * nh + 255) / 256, 256, 0, st>>>(gates, cstate, cstate, hstate, B, nh, forget_bias,
                                                                  air::HlOut{nullptr, 0, 0, 0},
                                                                  air::HlOut{nullptr, 0, 0, 0}, 0);
  AIR_CUDA(cudaGetLastError());
  AIR_CUDA(cudaFreeAsync(gates, st));
  return AIR_OK;
}

int32_t air_stn_read(const float* img, const float* where, float* crop, int32_t B, int32_t H, int32_t W, int32_t h,
                     int32_t w, void* stream) {
  if (!img || !where || !crop || B < 1 || H < 1 || W < 1 || h < 1 || w < 1)
    return fail(AIR_ERR_ARG, "air_stn_read: bad argument");
  const size_t smem = air::where_read_smem(1, H, W, h, w);
  if (smem > 200 * 1024) return fail(AIR_ERR_ARG, "air_stn_read: image does not fit in shared memory");
  if (smem > 48 * 1024)
    AIR_CUDA(cudaFuncSetAttribute(air::stn_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  air::stn_read_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(img, where, crop, H, W, h, w);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_stn_paint(const float* glimpse, const float* where, float* out, int32_t B, int32_t H, int32_t W, int32_t h,
                      int32_t w, void* stream) {
  if (!glimpse || !where || !out || B < 1 || H < 1 || W < 1 || h < 1 || w < 1)
    return fail(AIR_ERR_ARG, "air_stn_paint: bad argument");
  const size_t smem = air::paint_smem_direct(1, H, W, h, w);
  if (smem > 200 * 1024) return fail(AIR_ERR_ARG, "air_stn_paint: glimpse does not fit in shared memory");
  if (smem > 48 * 1024)
    AIR_CUDA(cudaFuncSetAttribute(air::stn_paint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  air::stn_paint_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(glimpse, where, out, H, W, h, w);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_glimpse_viz(const float* glimpse, const float* presence, float* out, int64_t rows, int32_t G, void* stream) {
  if (!glimpse || !presence || !out || rows < 0 || G < 1) return fail(AIR_ERR_ARG, "air_glimpse_viz: bad argument");
  if (rows == 0) return AIR_OK;
  const long long n = (long long)rows * G;
  const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, 148 * 16);
  air::glimpse_viz_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(glimpse, presence, out, (long long)rows, G);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_bernoulli_to_modified_geometric(const float* probs, float* pmf, int64_t n, int32_t T, void* stream) {
  if (!probs || !pmf || n < 0 || T < 1 || T > AIR_MAX_STEPS)
    return fail(AIR_ERR_ARG, "air_bernoulli_to_modified_geometric: bad argument (T must be in [1, AIR_MAX_STEPS])");
  if (n == 0) return AIR_OK;
  air::modified_geometric_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(probs, pmf, n, T);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_geometric_prior(double success_prob, int32_t n_steps, int32_t is_f64, void* out, void* stream) {
  if (!out || n_steps < 0) return fail(AIR_ERR_ARG, "air_geometric_prior: bad argument");
  air::geometric_prior_kernel<<<(n_steps + 1 + 63) / 64, 64, 0, (cudaStream_t)stream>>>(success_prob, n_steps, is_f64,
                                                                                       out);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_tabular_kl(const float* p, const double* q, float* kl, int64_t n, int32_t m, double zero_prob_value,
                       void* stream) {
  if (!p || !q || !kl || n < 0 || m < 1) return fail(AIR_ERR_ARG, "air_tabular_kl: bad argument");
  if (n == 0) return AIR_OK;
  air::tabular_kl_kernel<<<(unsigned)((n * m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, q, kl, n, m,
                                                                                           zero_prob_value);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_num_steps_log_prob(const float* pmf, const float* samples, float* out, int64_t n, int32_t m,
                               void* stream) {
  if (!pmf || !samples || !out || n < 0 || m < 1) return fail(AIR_ERR_ARG, "air_num_steps_log_prob: bad argument");
  if (n == 0) return AIR_OK;
  air::num_steps_log_prob_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pmf, samples, out, n, m,
                                                                                                1);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

int32_t air_sample_from_tensor(const float* pmf, const float* samples, float* out, int64_t n, int32_t m,
                               void* stream) {
  if (!pmf || !samples || !out || n < 0 || m < 1) return fail(AIR_ERR_ARG, "air_sample_from_tensor: bad argument");
  if (n == 0) return AIR_OK;
  air::num_steps_log_prob_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pmf, samples, out, n, m,
                                                                                                0);
  AIR_CUDA(cudaGetLastError());
  return AIR_OK;
}

double air_anneal_weight(double init_val, double final_val, int32_t anneal_type, double global_step,
                         double anneal_steps, double hold_for, double steps_div) {
  // tf.cast(python_float, tf.float64) first makes a float32 constant (ops.convert_to_tensor), then widens it: the
  // schedule constants of multi_mnist.py:40-47 enter the float64 island rounded to float32 (1 - 1e-15 -> 1.0).
  init_val = (double)(float)init_val;
  final_val = (double)(float)final_val;
  anneal_steps = (double)(float)anneal_steps;
  hold_for = (double)(float)hold_for;
  steps_div = (double)(float)steps_div;
  double step = global_step - hold_for;
  if (step < 0.0) step = 0.0;
  double val = init_val;
  if (anneal_type == 0) {
    // tf.train.exponential_decay(val, step, steps_div, rate) = val * rate ** (step / steps_div)  [upstream]
    const double rate = pow(final_val / init_val, steps_div / anneal_steps);
    val = init_val * pow(rate, step / steps_div);
  } else if (anneal_type == 1) {
    val = final_val + (init_val - final_val) * (1.0 - step / anneal_steps);
  } else {
    return NAN;   // NotImplementedError (model.py:121)
  }
  return val > final_val ? val : final_val;
}

}  // extern "C"
