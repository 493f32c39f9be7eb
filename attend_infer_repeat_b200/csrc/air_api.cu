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
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/air_b200.h"
#include "cell_kernels.cuh"
#include "common.cuh"
#include "linear_simt.cuh"

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

struct Layer {
  int64_t w_off = 0, b_off = 0;
  int K = 0, N = 0;
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

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct air_handle {
  air_config cfg;
  int P = 0, G = 0, n_enc = 0;
  std::vector<ParamEntry> entries;
  int64_t n_params = 0;
  Mlp enc, where_mlp, steps_mlp, glenc, dec;
  Layer what_lin;
  int64_t lstm_w = 0, lstm_b = 0, lstm_h0 = 0, lstm_c0 = 0;
  int max_width = 0;
  // workspace (one cudaMalloc)
  char* ws = nullptr;
  size_t ws_bytes = 0;
  float *buf_a = nullptr, *buf_b = nullptr;   // ping-pong activations [T*B, max_width]
  float *e = nullptr, *gx = nullptr, *gates = nullptr, *h_init = nullptr, *cbuf = nullptr, *hs = nullptr;
  float *m = nullptr, *logit = nullptr, *crop = nullptr, *r = nullptr;
  // staging for air_forward_host
  float *st_img = nullptr, *st_eps_where = nullptr, *st_eps_what = nullptr, *st_u = nullptr;
  float *st_pres_in = nullptr;
  // instrumentation: kernel-launch counter and optional per-stage CUDA-event timing (air_profile_*)
  uint64_t launches = 0;
  bool profile = false;
  cudaEvent_t ev[AIR_N_STAGES + 1] = {};
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

// one dense layer; every GEMM of the path goes through here (engine selection + launch accounting)
cudaError_t gemm(air_handle* h, const float* A, int lda, const float* Wt, int ldw, const float* bias,
                 const float* addend, int ldadd, float* C, int ldc, int M, int N, int K, int act, cudaStream_t st) {
  ++h->launches;
  return air::launch_linear_simt(A, lda, Wt, ldw, bias, addend, ldadd, C, ldc, M, N, K, act, st);
}

cudaError_t linear(air_handle* h, const float* A, int lda, const float* params, const Layer& l,
                   const float* addend, int ldadd, float* C, int ldc, int M, int act, cudaStream_t st) {
  return gemm(h, A, lda, params + l.w_off, l.N, params + l.b_off, addend, ldadd, C, ldc, M, l.N, l.K, act, st);
}

inline void mark(air_handle* h, int stage, cudaStream_t st) {
  if (h->profile) cudaEventRecord(h->ev[stage], st);
}

// neural.MLP (neural.py:63-102): ELU hidden layers, linear output layer.  Intermediate activations ping-pong
// between the two workspace buffers; the last layer writes `out`.
cudaError_t run_mlp(air_handle* h, const float* params, const Mlp& mlp, const float* in, int ld_in, int M,
                    float* out, int ld_out, cudaStream_t st) {
  const float* cur = in;
  int ld = ld_in;
  const int nl = (int)mlp.layers.size();
  for (int i = 0; i < nl; ++i) {
    const Layer& l = mlp.layers[i];
    const bool last = (i == nl - 1);
    float* dst = last ? out : ((cur == h->buf_a) ? h->buf_b : h->buf_a);
    const int ldd = last ? ld_out : l.N;
    const int act = (i < mlp.n_hidden) ? air::ACT_ELU : air::ACT_NONE;
    cudaError_t e = linear(h, cur, ld, params, l, nullptr, 0, dst, ldd, M, act, st);
    if (e != cudaSuccess) return e;
    cur = dst;
    ld = ldd;
  }
  return cudaSuccess;
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

// The shared body of air_forward / air_cell_step: T_run steps starting from explicit or initial state.
int32_t forward_impl(air_handle* h, const float* params, const float* img, const float* eps_where,
                     const float* eps_what, const float* u_pres, const float* baseline, const air_prior* prior,
                     const air_outputs* o, int T_run, const float* h_in, const float* c_in, const float* presence_in,
                     const float* canvas_in, float* canvas_step_out, float mult, cudaStream_t st) {
  const air_config& c = h->cfg;
  const int B = c.B, nh = c.nh, P = h->P, G = h->G, na = c.na;
  const int TB = T_run * B;
  const int thr = 256;

  // 1. e = Encoder(img)   (modules.py:72-76; step-invariant, cell.py:125)
  mark(h, AIR_ST_ENCODER, st);
  AIR_CUDA(run_mlp(h, params, h->enc, img, P, B, h->e, h->n_enc, st));
  mark(h, AIR_ST_LSTM, st);

  // 2. gx = e @ W[:n_enc] + b   (input half of snt.LSTM's [x,h] @ W + b)
  AIR_CUDA(gemm(h, h->e, h->n_enc, params + h->lstm_w, 4 * nh, params + h->lstm_b, nullptr, 0, h->gx, 4 * nh, B, 4 * nh,
                h->n_enc, air::ACT_NONE, st));

  // 3. recurrence: gates = gx + h_{t-1} @ W[n_enc:]; (c, h_t) pointwise
  if (h_in) {
    AIR_CUDA(cudaMemcpyAsync(h->h_init, h_in, sizeof(float) * B * nh, cudaMemcpyDeviceToDevice, st));
    AIR_CUDA(cudaMemcpyAsync(h->cbuf, c_in, sizeof(float) * B * nh, cudaMemcpyDeviceToDevice, st));
  } else {
    air::lstm_init_state_kernel<<<(B * nh + thr - 1) / thr, thr, 0, st>>>(params + h->lstm_h0, params + h->lstm_c0,
                                                                         h->h_init, h->cbuf, B, nh);
    AIR_CUDA(cudaGetLastError());
    ++h->launches;
  }
  for (int t = 0; t < T_run; ++t) {
    const float* h_prev = (t == 0) ? h->h_init : h->hs + (size_t)(t - 1) * B * nh;
    AIR_CUDA(gemm(h, h_prev, nh, params + h->lstm_w + (int64_t)h->n_enc * 4 * nh, 4 * nh, nullptr, h->gx, 4 * nh,
                  h->gates, 4 * nh, B, 4 * nh, nh, air::ACT_NONE, st));
    air::lstm_pointwise_kernel<<<(B * nh + thr - 1) / thr, thr, 0, st>>>(h->gates, h->cbuf,
                                                                        h->hs + (size_t)t * B * nh, B, nh,
                                                                        c.forget_bias);
    AIR_CUDA(cudaGetLastError());
    ++h->launches;
  }
  if (o->final_h)
    AIR_CUDA(cudaMemcpyAsync(o->final_h, h->hs + (size_t)(T_run - 1) * B * nh, sizeof(float) * B * nh,
                             cudaMemcpyDeviceToDevice, st));
  if (o->final_c)
    AIR_CUDA(cudaMemcpyAsync(o->final_c, h->cbuf, sizeof(float) * B * nh, cudaMemcpyDeviceToDevice, st));

  // 4. heads over all T*B hidden states at once
  mark(h, AIR_ST_WHERE_MLP, st);
  AIR_CUDA(run_mlp(h, params, h->where_mlp, h->hs, nh, TB, h->m, 8, st));          // modules.py:58-63
  mark(h, AIR_ST_STEPS, st);
  AIR_CUDA(run_mlp(h, params, h->steps_mlp, h->hs, nh, TB, h->logit, 1, st));      // modules.py:119-122
  air::presence_kernel<<<(B + 127) / 128, 128, 0, st>>>(h->logit, u_pres, presence_in, o->presence_prob, o->presence,
                                                        T_run, B, c.step_bias, c.explore_eps, c.discrete_steps);
  AIR_CUDA(cudaGetLastError());
  ++h->launches;
  mark(h, AIR_ST_READ, st);

  // 5. where sampling + glimpse read   (cell.py:129-135)
  air::where_read_kernel<<<B, 256, sizeof(float) * P, st>>>(h->m, eps_where, img, o->where, o->where_loc,
                                                            o->where_scale, h->crop, T_run, B, c.H, c.W, c.h, c.w,
                                                            c.max_crop_size, c.scale_bias);
  AIR_CUDA(cudaGetLastError());
  ++h->launches;
  mark(h, AIR_ST_GLIMPSE_ENC, st);

  // 6. glimpse encoder -> what   (cell.py:153-156)
  AIR_CUDA(run_mlp(h, params, h->glenc, h->crop, G, TB, (h->glenc.layers.size() & 1) ? h->buf_a : h->buf_b,
                   h->glenc.layers.back().N, st));
  {
    const float* q = (h->glenc.layers.size() & 1) ? h->buf_a : h->buf_b;
    AIR_CUDA(linear(h, q, h->glenc.layers.back().N, params, h->what_lin, nullptr, 0, h->r, 2 * na, TB, air::ACT_NONE,
                    st));
  }
  {
    const size_t n = (size_t)TB * na;
    air::what_kernel<<<(unsigned)((n + thr - 1) / thr), thr, 0, st>>>(h->r, eps_what, o->what, o->what_loc,
                                                                      o->what_scale, (size_t)TB, na,
                                                                      c.what_scale_offset);
    AIR_CUDA(cudaGetLastError());
    ++h->launches;
  }
  mark(h, AIR_ST_DECODER, st);

  // 7. decoder   (cell.py:158)
  AIR_CUDA(run_mlp(h, params, h->dec, o->what, na, TB, o->glimpse, G, st));
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
  if (prior) a.prior = *prior;
  const size_t smem = sizeof(float) * ((size_t)T_run * G + (size_t)T_run * (c.W + c.H));
  air::paint_elbo_kernel<<<B, 256, smem, st>>>(a);
  AIR_CUDA(cudaGetLastError());
  ++h->launches;

  if (prior) {
    air::elbo_scalars_kernel<<<1, 1024, 0, st>>>(o->rec_loss_per_sample, o->kl_num_steps_per_sample,
                                                 o->kl_what_per_sample, o->kl_where_per_sample, o->num_step_per_sample,
                                                 o->num_steps_log_prob, baseline, o->scalars, B, *prior);
    AIR_CUDA(cudaGetLastError());
    ++h->launches;
  }
  mark(h, AIR_N_STAGES, st);
  return AIR_OK;
}

const char* const kStageNames[AIR_N_STAGES] = {"input_encoder", "lstm",        "where_mlp", "steps_presence",
                                               "where_read",    "glimpse_enc", "decoder",   "paint_elbo"};

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
  // canonical flat parameter order == the Sonnet variables of cell.py:61-69 in creation order
  h->n_enc = build_mlp(h, h->enc, "input_encoder", h->P, c.enc_hidden, c.n_enc_hidden, 0);
  h->lstm_w = add_entry(h, "lstm.w", h->n_enc + c.nh, 4 * c.nh);
  h->lstm_b = add_entry(h, "lstm.b", 1, 4 * c.nh);
  h->lstm_h0 = add_entry(h, "lstm.h0", 1, c.nh);
  h->lstm_c0 = add_entry(h, "lstm.c0", 1, c.nh);
  build_mlp(h, h->where_mlp, "transform_estimator", c.nh, c.where_hidden, c.n_where_hidden, 8);
  build_mlp(h, h->steps_mlp, "steps_predictor", c.nh, c.steps_hidden, c.n_steps_hidden, 1);
  const int n_gl = build_mlp(h, h->glenc, "glimpse_encoder", h->G, c.glenc_hidden, c.n_glenc_hidden, 0);
  h->what_lin.K = n_gl;
  h->what_lin.N = 2 * c.na;
  h->what_lin.w_off = add_entry(h, "what.w", n_gl, 2 * c.na);
  h->what_lin.b_off = add_entry(h, "what.b", 1, 2 * c.na);
  build_mlp(h, h->dec, "glimpse_decoder", c.na, c.dec_hidden, c.n_dec_hidden, h->G);

  // shared-memory budgets of the per-canvas kernels
  const size_t smem_read = sizeof(float) * h->P;
  const size_t smem_paint = sizeof(float) * ((size_t)c.T * h->G + (size_t)c.T * (c.W + c.H));
  if (smem_read > 200 * 1024 || smem_paint > 200 * 1024) {
    delete h;
    return fail(AIR_ERR_ARG, "air_create: image / glimpse tile does not fit in shared memory");
  }
  if (smem_read > 48 * 1024) {
    cudaFuncSetAttribute(air::where_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_read);
    cudaFuncSetAttribute(air::stn_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_read);
  }
  if (smem_paint > 48 * 1024)
    cudaFuncSetAttribute(air::paint_elbo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_paint);

  // workspace carve-up
  const size_t TB = (size_t)c.T * c.B, B = c.B;
  struct Slot { float** p; size_t n; };
  const Slot slots[] = {
      {&h->buf_a, TB * h->max_width}, {&h->buf_b, TB * h->max_width}, {&h->e, B * h->n_enc},
      {&h->gx, B * 4 * c.nh},         {&h->gates, B * 4 * c.nh},      {&h->h_init, B * c.nh},
      {&h->cbuf, B * c.nh},           {&h->hs, TB * c.nh},            {&h->m, TB * 8},
      {&h->logit, TB},                {&h->crop, TB * h->G},          {&h->r, TB * 2 * c.na},
      {&h->st_img, B * h->P},         {&h->st_eps_where, TB * 4},     {&h->st_eps_what, TB * c.na},
      {&h->st_u, TB},                 {&h->st_pres_in, B},
  };
  size_t total = 0;
  for (const Slot& s : slots) total += align_up(s.n * sizeof(float), 256);
  cudaError_t e = cudaMalloc(&h->ws, total);
  if (e != cudaSuccess) {
    delete h;
    return fail(AIR_ERR_NOMEM, std::string("air_create: cudaMalloc workspace: ") + cudaGetErrorString(e));
  }
  h->ws_bytes = total;
  size_t off = 0;
  for (const Slot& s : slots) {
    *s.p = reinterpret_cast<float*>(h->ws + off);
    off += align_up(s.n * sizeof(float), 256);
  }
  *out = h;
  return AIR_OK;
}

int32_t air_destroy(air_handle* h) {
  if (!h) return AIR_OK;
  for (cudaEvent_t e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->ws) cudaFree(h->ws);
  delete h;
  return AIR_OK;
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

int32_t air_elbo_scalars(air_handle* h, const float* baseline, const air_prior* prior, const air_outputs* outs,
                         void* stream) {
  if (!h || !prior) return fail(AIR_ERR_ARG, "air_elbo_scalars: NULL argument");
  const int32_t rc = check_outs(outs, true);
  if (rc != AIR_OK) return rc;
  air::elbo_scalars_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
      outs->rec_loss_per_sample, outs->kl_num_steps_per_sample, outs->kl_what_per_sample, outs->kl_where_per_sample,
      outs->num_step_per_sample, outs->num_steps_log_prob, baseline, outs->scalars, h->cfg.B, *prior);
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
  if (!A || !Wt || !out || M < 0 || N < 1 || K < 1) return fail(AIR_ERR_ARG, "air_linear: bad argument");
  if (precision != AIR_PREC_FP32) return fail(AIR_ERR_ARG, "air_linear: only AIR_PREC_FP32 stand-alone");
  AIR_CUDA(air::launch_linear_simt(A, K, Wt, N, bias, nullptr, 0, out, N, M, N, K, act, (cudaStream_t)stream));
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
  air::lstm_pointwise_kernel<<<(B * nh + 255) / 256, 256, 0, st>>>(gates, cstate, hstate, B, nh, forget_bias);
  AIR_CUDA(cudaGetLastError());
  AIR_CUDA(cudaFreeAsync(gates, st));
  return AIR_OK;
}

int32_t air_stn_read(const float* img, const float* where, float* crop, int32_t B, int32_t H, int32_t W, int32_t h,
                     int32_t w, void* stream) {
  if (!img || !where || !crop || B < 1 || H < 1 || W < 1 || h < 1 || w < 1)
    return fail(AIR_ERR_ARG, "air_stn_read: bad argument");
  const size_t smem = sizeof(float) * (size_t)H * W;
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
  const size_t smem = sizeof(float) * (size_t)h * w;
  if (smem > 200 * 1024) return fail(AIR_ERR_ARG, "air_stn_paint: glimpse does not fit in shared memory");
  if (smem > 48 * 1024)
    AIR_CUDA(cudaFuncSetAttribute(air::stn_paint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  air::stn_paint_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(glimpse, where, out, H, W, h, w);
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
