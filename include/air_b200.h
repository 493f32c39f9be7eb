/*
 * air_b200.h -- C ABI of the B200-native AIR (Attend-Infer-Repeat) hot path.
 *
 * The reference (akosiorek/attend_infer_repeat) has no FFI of its own: its boundary is a Python
 * class surface (AIRCell / AIRModel / NumStepsDistribution) on top of Sonnet's custom-op library
 * (snt.resampler) and the TensorFlow kernels.  Every entry point below names the reference
 * interface (file:line under /root/reference) that it replaces.  The Python mirror of the reference
 * classes (attend_infer_repeat_b200/*.py) binds these symbols with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller unless the
 *     parameter name ends in `_host`;
 *   - all tensors are dense row-major float32 unless stated, time-major [T,B,...] like
 *     tf.nn.dynamic_rnn(time_major=True) (model.py:83-84);
 *   - every call enqueues on the caller's stream (`stream` is a cudaStream_t passed as void*) and
 *     returns without synchronising unless the name ends in `_host`;
 *   - return value: 0 = ok, negative = air_status; air_last_error() gives the message of the last
 *     failure on the calling thread;
 *   - the library never allocates persistent device memory except the workspace owned by an
 *     air_handle (air_create .. air_destroy).
 */
#ifndef AIR_B200_H_
#define AIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIR_ABI_VERSION 1
#define AIR_MAX_HIDDEN 4   /* hidden layers per MLP (reference script uses 2) */
#define AIR_MAX_STEPS 8    /* max_steps (reference script uses 3)            */

typedef enum air_status {
  AIR_OK = 0,
  AIR_ERR_ARG = -1,       /* bad shape / null pointer / unsupported configuration */
  AIR_ERR_CUDA = -2,      /* a CUDA runtime call or kernel launch failed          */
  AIR_ERR_ARCH = -3,      /* device is not sm_100 (no other target is built)      */
  AIR_ERR_NOMEM = -4,
  AIR_ERR_RANGE = -5      /* tensor-core split path saw a value outside fp16 range */
} air_status;

typedef enum air_precision {
  AIR_PREC_FP32 = 0,      /* fp32 FMA on CUDA cores: bit-faithful to an fp32 reference up to summation order */
  AIR_PREC_TC_SPLIT = 1   /* tcgen05 fp16x2-split (3 MMAs / product), fp32-class accuracy on tensor cores    */
} air_precision;

/* Hyper-parameters of the path: AIRCell.__init__ (cell.py:15-69), AIRModel.__init__ (model.py:18-64),
 * AIRonMNIST.__init__ (mnist_model.py:13-44), values in scripts/multi_mnist.py:24-94. */
typedef struct air_config {
  int32_t B;                 /* batch size (model.py:60)                                */
  int32_t H, W;              /* image size (cell.py:38)                                 */
  int32_t h, w;              /* glimpse / crop size (cell.py:40)                        */
  int32_t T;                 /* max_steps (model.py:83)                                 */
  int32_t na;                /* n_appearance (cell.py:41)                               */
  int32_t nh;                /* LSTM units, snt.LSTM(256) (mnist_model.py:35)           */
  int32_t n_enc_hidden;   int32_t enc_hidden[AIR_MAX_HIDDEN];    /* Encoder (input),  modules.py:66-76  */
  int32_t n_glenc_hidden; int32_t glenc_hidden[AIR_MAX_HIDDEN];  /* Encoder (glimpse)                   */
  int32_t n_dec_hidden;   int32_t dec_hidden[AIR_MAX_HIDDEN];    /* Decoder, modules.py:79-91           */
  int32_t n_where_hidden; int32_t where_hidden[AIR_MAX_HIDDEN];  /* StochasticTransformParam, :53-63    */
  int32_t n_steps_hidden; int32_t steps_hidden[AIR_MAX_HIDDEN];  /* StepsPredictor, :112-122            */
  float output_std;          /* model.py:97 (mnist_model.py:42 -> .3)                   */
  float output_multiplier;   /* model.py:58,93                                          */
  float explore_eps;         /* cell.py:140-141; < 0 means None                         */
  float scale_bias;          /* transform_var_bias, modules.py:54,63                    */
  float step_bias;           /* modules.py:121                                          */
  float what_scale_offset;   /* cell.py:66 (0.5)                                        */
  float forget_bias;         /* snt.LSTM default 1.0                                    */
  float max_crop_size;       /* modules.py:29                                           */
  int32_t discrete_steps;    /* cell.py:143-151                                         */
  int32_t precision;         /* air_precision                                           */
} air_config;

/* Priors and loss switches: AIRModel.train_step arguments (model.py:261-265) and the AttrDicts of
 * scripts/multi_mnist.py:38-51.  The annealed success probability (model.py:106-124,133-142) is a
 * scalar schedule: compute it with air_anneal_weight() and pass it in. */
typedef struct air_prior {
  float what_loc, what_scale;                 /* what_prior                               */
  float where_scale_loc, where_scale_scale;   /* where_scale_prior                        */
  float where_shift_loc, where_shift_scale;   /* where_shift_prior                        */
  int32_t where_shift_has_loc;                /* 'loc' in where_shift_prior (model.py:202) */
  double steps_success_prob;                  /* model.py:133-142                         */
  int32_t steps_prob_is_f64;                  /* 1: annealed (float64 island); 0: python float -> float32 (prior.py:28) */
  float steps_weight;                         /* num_steps_prior.weight (model.py:154)    */
  int32_t analytic;                           /* num_steps_prior.analytic (model.py:157)  */
  int32_t use_prior;                          /* prior_weight = float(use_prior), model.py:331 */
  int32_t use_reinforce;                      /* model.py:335                             */
  /* NVIL normalisation of the REINFORCE importance weight (decay_rate, model.py:232-239):
   * iw <- (iw - nvil_shift) * nvil_scale, shift = imp_weight_moving_mean, scale = 1 / max(sqrt(moving_var), 1).
   * nvil_scale == 0 selects "off" (shift 0, scale 1). */
  float nvil_shift, nvil_scale;
} air_prior;

/* Caller-owned output buffers of one unrolled forward pass (+ ELBO terms).  A NULL pointer means
 * "do not materialise" for the optional ones (marked opt).
 * Names follow AIRCell.output_names (cell.py:97-99) and the AIRModel attributes (model.py:86-104,
 * 143-372). */
typedef struct air_outputs {
  /* AIRCell outputs stacked time-major by dynamic_rnn (model.py:84-87) */
  float* canvas;            /* opt [T,B,H*W]  canvas * output_multiplier (model.py:92-93)          */
  float* glimpse;           /*     [T,B,h*w]  raw decoder output (cell output "glimpse")            */
  float* glimpse_viz;       /* opt [T,B,h*w]  presence * sigmoid(glimpse) (model.py:90)             */
  float* what;              /*     [T,B,na]                                                         */
  float* what_loc;          /*     [T,B,na]                                                         */
  float* what_scale;        /*     [T,B,na]                                                         */
  float* where;             /*     [T,B,4]   (sx,tx,sy,ty), modules.py:42                           */
  float* where_loc;         /*     [T,B,4]                                                          */
  float* where_scale;       /*     [T,B,4]                                                          */
  float* presence_prob;     /*     [T,B,1]                                                          */
  float* presence;          /*     [T,B,1]                                                          */
  float* final_h;           /*     [B,nh]    final_state (model.py:89)                              */
  float* final_c;           /*     [B,nh]                                                           */
  /* NumStepsDistribution + ELBO terms (prior.py:119-151, model.py:126-251,319-343) */
  float* num_steps_posterior;     /* [B,T+1] num_steps_distrib.prob()                               */
  float* num_step_per_sample;     /* [B]     model.py:102                                           */
  float* prior_step_weight;       /* [T,B]   model.py:157-165                                       */
  float* rec_loss_per_sample;     /* [B]     model.py:320-321                                       */
  float* kl_num_steps_per_sample; /* [B]     model.py:149                                           */
  float* kl_what_per_sample;      /* [B]     model.py:181-182                                       */
  float* kl_where_per_sample;     /* [B]     model.py:209-210                                       */
  float* loss_per_sample;         /* [B]     loss.per_sample (ops.py:12-29)                         */
  float* num_steps_log_prob;      /* [B]     num_steps_distrib.log_prob(num_step_per_sample)        */
  float* scalars;                 /* [AIR_N_SCALARS] batch SUMS /B, see air_scalar                  */
} air_outputs;

typedef enum air_scalar {
  AIR_S_REC_LOSS = 0,        /* model.py:322                                   */
  AIR_S_KL_NUM_STEPS = 1,    /* model.py:151                                   */
  AIR_S_KL_WHAT = 2,         /* model.py:184                                   */
  AIR_S_KL_WHERE = 3,        /* model.py:212                                   */
  AIR_S_PRIOR_LOSS = 4,      /* prior_loss.value                               */
  AIR_S_LOSS = 5,            /* loss.value ; ELBO = -loss.value                */
  AIR_S_REINFORCE = 6,       /* model.py:247-248 (0 if !use_reinforce)         */
  AIR_S_OPT_LOSS = 7,        /* model.py:335-343 (without L2)                  */
  AIR_S_NUM_STEP = 8,        /* model.py:103                                   */
  AIR_S_MEAN_REC_LOGQ = 9,   /* mean_j rec_j * log q(n_j)   (REINFORCE pieces) */
  AIR_S_MEAN_LOGQ = 10,      /* mean_j log q(n_j)                              */
  AIR_S_MEAN_BASELINE = 11,  /* mean_i baseline_i                              */
  AIR_S_MEAN_IW = 12,        /* mean_j iw_j            (moments of the importance weight, model.py:233-234) */
  AIR_S_MEAN_IW2 = 13,       /* mean_j iw_j^2                                  */
  AIR_S_MEAN_BASELINE2 = 14, /* mean_i baseline_i^2                            */
  AIR_N_SCALARS = 16
} air_scalar;

typedef struct air_handle air_handle;

/* ---- library ---------------------------------------------------------------------------------- */
int32_t air_abi_version(void);
const char* air_last_error(void);

/* ---- handle: replaces graph construction in AIRCell.__init__ / AIRModel._build (cell.py:15-69,
 *      model.py:66-104).  Allocates the activation workspace for cfg->B samples on the current device. */
int32_t air_create(const air_config* cfg, air_handle** out);
int32_t air_destroy(air_handle* h);
/* number of float32 parameters and the canonical flat layout (name, offset, rows, cols) -- the
 * variables Sonnet would create for the modules of cell.py:61-69. */
int64_t air_param_count(const air_handle* h);
int32_t air_param_entries(const air_handle* h);
int32_t air_param_entry(const air_handle* h, int32_t i, const char** name, int64_t* offset,
                        int32_t* rows, int32_t* cols);
int64_t air_workspace_bytes(const air_handle* h);
/* Host-only self-check of the fused row kernel's schedule (csrc/row_tc.cuh) for a configuration: builds the unit / task
 * programs of the producer, MMA and epilogue roles for the row path of cell.py:129-158 and replays them against the
 * mbarrier protocol.  AIR_OK when the configuration is covered and the replay is deadlock- and hazard-free; otherwise
 * AIR_ERR_ARG with the reason in `msg` (the engine then uses the per-stage kernels).  Needs no GPU. */
int32_t air_row_schedule_check(const air_config* cfg, char* msg, int32_t msg_len);

/* ---- instrumentation (bench.py): kernels launched so far through this handle, and per-stage device
 *      time of the LAST air_forward measured with CUDA events on the caller's stream. */
typedef enum air_stage {
  AIR_ST_ENCODER = 0,     /* input Encoder GEMMs (cell.py:125)                       */
  AIR_ST_LSTM = 1,        /* gx GEMM + T x (recurrent GEMM + gate math) (cell.py:126) */
  AIR_ST_WHERE_MLP = 2,   /* transform-estimator GEMMs (cell.py:129)                 */
  AIR_ST_STEPS = 3,       /* steps-predictor GEMMs + presence scan (cell.py:137-151) */
  AIR_ST_READ = 4,        /* where sampling + STN glimpse read (cell.py:130-135)     */
  AIR_ST_GLIMPSE_ENC = 5, /* glimpse Encoder + what head (cell.py:153-156)           */
  AIR_ST_DECODER = 6,     /* Decoder GEMMs (cell.py:158)                             */
  AIR_ST_PAINT_ELBO = 7,  /* inverse STN paint + ELBO terms + batch means            */
  AIR_N_STAGES = 8
} air_stage;
int64_t air_launch_count(const air_handle* h);
int32_t air_profile_enable(air_handle* h, int32_t on);
int32_t air_profile_read(air_handle* h, float* ms_per_stage, int32_t n);
const char* air_stage_name(int32_t i);

/* AIR_PREC_TC_SPLIT carries operands as fp16 hi/lo pairs: a weight * 2^8 or an activation beyond +-65504
 * cannot be represented.  The kernels raise a device flag instead of producing silent infinities; this call
 * reads it (synchronises the stream), clears it and returns AIR_ERR_RANGE if it was set. */
int32_t air_check_range(air_handle* h, void* stream);

/* ---- the hot path ----------------------------------------------------------------------------- */
/* T unrolled AIRCell steps + post-processing + ELBO terms in one enqueue:
 * tf.nn.dynamic_rnn over AIRCell._build (model.py:81-104, cell.py:116-171) followed by the loss
 * assembly of AIRModel.train_step (model.py:319-343), _prior_loss (126-216), _reinforce (218-251).
 *   params     [air_param_count]   flat float32 parameters
 *   img        [B,H,W]             obs
 *   eps_where  [T,B,4]             N(0,1) draws of where_distrib.sample()    (cell.py:133)
 *   eps_what   [T,B,na]            N(0,1) draws of what_distrib.sample()     (cell.py:156)
 *   u_pres     [T,B,1]             U[0,1) draws of presence_distrib.sample() (cell.py:147)
 *   baseline   [B] or NULL         BaselineMLP output (model.py:224-230)
 *   prior      NULL -> only the cell outputs are produced (no ELBO terms) */
int32_t air_forward(air_handle* h, const float* params, const float* img, const float* eps_where,
                    const float* eps_what, const float* u_pres, const float* baseline,
                    const air_prior* prior, const air_outputs* outs, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies img and the noise host->device, runs
 * air_forward, copies `scalars` and the [B] ELBO vectors back and synchronises the stream.  `outs`
 * still points at DEVICE buffers; the *_host arguments receive the copies (may be NULL).
 * This is the end-to-end call a sess.run([loss...]) of the reference maps to (multi_mnist.py:136). */
int32_t air_forward_host(air_handle* h, const float* params, const float* img_host,
                         const float* eps_where_host, const float* eps_what_host,
                         const float* u_pres_host, const air_prior* prior, const air_outputs* outs,
                         float* scalars_host, float* loss_per_sample_host, void* stream);

/* Same with the images in the reference's DATASET format: uint8 [B,H,W] (data.py:35-107; load_data divides by 255 on
 * the host, data.py:116).  The division runs on the device, fused with the first layer's operand preparation, so the
 * host->device copy is 4x smaller. */
int32_t air_forward_host_u8(air_handle* h, const float* params, const uint8_t* img_u8_host,
                            const float* eps_where_host, const float* eps_what_host,
                            const float* u_pres_host, const air_prior* prior, const air_outputs* outs,
                            float* scalars_host, float* loss_per_sample_host, void* stream);

/* In-library noise: fills eps_where [T,B,4] ~ N(0,1), eps_what [T,B,na] ~ N(0,1), u_pres [T,B,1] ~ U[0,1)
 * (DEVICE buffers, any may be NULL) from a counter-based Philox4x32-10 generator: a pure function of (seed, tensor,
 * element index), like the in-graph draws of cell.py:133,147,156 under a fixed graph seed. */
int32_t air_draw_noise(air_handle* h, uint64_t seed, float* eps_where, float* eps_what, float* u_pres, void* stream);
/* air_forward_host_u8 with the noise drawn on the device (air_draw_noise(seed)): the only host->device traffic is the
 * uint8 image batch -- what sess.run(train_step, feed_dict={imgs}) moves in the reference. */
int32_t air_forward_host_u8_rng(air_handle* h, const float* params, const uint8_t* img_u8_host, uint64_t seed,
                                const air_prior* prior, const air_outputs* outs, float* scalars_host,
                                float* loss_per_sample_host, void* stream);

/* Double-buffered host feed: the role of the reference's input queue (data.py:121-158, tf.train.slice_input_producer /
 * tf.train.batch prefetching batch i+1 while sess.run works on batch i, multi_mnist.py:73,136).
 *   air_feed_host_u8(h, slot, img)        enqueue the host->device copy of one uint8 batch [B,H,W] (pinned host memory)
 *                                         into staging slot 0 or 1 on the handle's own copy stream; returns at once.  It
 *                                         waits (on the device) until the last pass that read the slot has consumed it.
 *   air_forward_fed_u8_rng(h, ..., slot)  air_forward_host_u8_rng on the batch fed into `slot`: `stream` waits for the
 *                                         copy, runs the pass, enqueues the device->host copies of scalars /
 *                                         loss_per_sample into the caller's (pinned) host buffers and returns WITHOUT
 *                                         synchronising.
 *   air_feed_wait(h, slot)                blocks the host until the results of the last pass on `slot` are in host memory.
 * Every step still moves its own batch in and its own loss out; only the waiting is overlapped. */
int32_t air_feed_host_u8(air_handle* h, int32_t slot, const uint8_t* img_u8_host);
int32_t air_forward_fed_u8_rng(air_handle* h, const float* params, int32_t slot, uint64_t seed, const air_prior* prior,
                               const air_outputs* outs, float* scalars_host, float* loss_per_sample_host, void* stream);
int32_t air_feed_wait(air_handle* h, int32_t slot);

/* Same with the whole DATASET resident on the device: dataset_u8 [n_dataset,H,W] uint8 as pickled by
 * data.py:35-107, idx [B] int32 = the minibatch indices tensors_from_data draws (data.py:121-158).  Gather, /255 and the
 * first layer's operand preparation run in one device pass; no host buffer is touched.  img_out [B,H,W] (optional)
 * receives the float32 minibatch (what a following air_backward needs); every pointer is a DEVICE pointer. */
int32_t air_forward_dataset_u8(air_handle* h, const float* params, const uint8_t* dataset_u8, int64_t n_dataset,
                               const int32_t* idx, const float* eps_where, const float* eps_what, const float* u_pres,
                               const float* baseline, const air_prior* prior, const air_outputs* outs, float* img_out,
                               void* stream);
/* stand-alone minibatch gather: img_out[b] = float32(dataset_u8[idx[b]]) / 255  (data.py:116,131-132) */
int32_t air_gather_u8(const uint8_t* dataset_u8, const int32_t* idx, float* img_out, int32_t B, int32_t P, void* stream);

/* Inference loops (AIR_PREC_TC_SPLIT): by default every forward call re-derives the fp16-split weight arena from `params`
 * (7 MB, one kernel), so that an optimiser step needs no extra call.  air_cache_weights(h, 1) keeps the arena across calls
 * made with the SAME params pointer; the caller must then call air_params_updated(h) after changing the buffer's contents
 * (a TF graph has the same contract between a variable assignment and the ops that read it). */
int32_t air_cache_weights(air_handle* h, int32_t on);

/* Programmatic dependent launch between the kernels of a forward pass (default on): the next kernel's CTAs are scheduled while
 * the previous kernel drains, which shortens ONE pass.  Turn it off for handles that share the device with other handles'
 * passes (several batches in flight, EnginePool): an early tensor-kernel CTA holds a whole SM while it only waits for its
 * predecessor, and a neighbouring batch's kernel cannot use that SM (measured at B = 4096, four batches in flight: 0.2155 ->
 * 0.2036 ms per batch without it; one batch alone: 0.2670 -> 0.2716). */
int32_t air_set_launch_overlap(air_handle* h, int32_t on);
int32_t air_params_updated(air_handle* h);

/* Replayed CUDA graphs (the training step of AIRModel.train_step, model.py:261-376, captured once and replayed): kernel
 * arguments are frozen at capture, but the annealed prior on the number of steps (model.py:133-142, a new success
 * probability every iteration) is not.  With a non-NULL `prior` this call writes geometric_prior(success_prob, T)
 * (prior.py:26-32; float64 or float32 island as `steps_prob_is_f64` says) into the handle's device table from a one-warp
 * kernel on `stream`, and every later air_forward / air_backward on the handle reads the table from there instead of from
 * its own `prior` argument (whose other fields are used as before).  prior == NULL switches back. */
int32_t air_prior_table_device(air_handle* h, const air_prior* prior, void* stream);

/* Importance-weighted bound (BASELINE.json configs[4]; an EXTENSION: the reference has no IWAE).  The K particles of a
 * canvas are K consecutive rows (row = canvas * K + particle: same image, own noise) of an ordinary air_forward over
 * R = n_canvases * K rows with a prior; this call turns that pass's outputs into
 *   log_w[R]              = log p(x|z) + log p(z,n) - log q(z,n|x)   (Normal priors, geometric step prior, NumStepsDistribution)
 *   bound_per_canvas[n]   = logsumexp_k log_w - log K,    *bound_mean = their batch mean (may be NULL).
 * All particles of a canvas live on one device, so batch sharding needs no extra collective. */
int32_t air_iwae_bound(int32_t n_canvases, int32_t K, int32_t T, int32_t na, const float* what, const float* what_loc,
                       const float* what_scale, const float* where, const float* where_loc, const float* where_scale,
                       const float* presence, const float* rec_loss_per_row, const float* num_steps_log_prob_per_row,
                       const air_prior* prior, float* log_w, float* bound_per_canvas, float* bound_mean, void* stream);

/* ---- training step: opt.compute_gradients(opt_loss, model_vars) + opt.apply_gradients of
 *      AIRModel.train_step (model.py:261-265,335-360) ------------------------------------------------------- */
/* Switch a handle (either engine, discrete_steps = 1) to training mode: allocates the second workspace (saved
 * activations + gradient scratch, air_train_workspace_bytes) once; from then on air_forward with a prior keeps every
 * activation the backward pass needs.  air_backward runs its weight-gradient work on side streams owned by the handle
 * (AIR_SIDE_STREAMS = 0..4, default 2) and joins them on the caller's stream before it returns. */
int32_t air_train_enable(air_handle* h, int32_t on);
int64_t air_train_workspace_bytes(const air_handle* h);
/* Gradient of opt_loss = loss.value + reinforce_loss (+ l2) with respect to the flat parameter buffer, for the batch the
 * LAST air_forward on this handle saw (same params / img / noise / prior / outs must be passed again; `outs` must have a
 * materialised canvas).  tf.gradients semantics: reparameterised what / where samples, presence samples and the REINFORCE
 * importance weight are constants (stop_gradient, model.py:246), log q(n) clips straight-through (ops.py:67-76).
 *   baseline_mean  mean over the GLOBAL batch of the baseline (0 without one): the [B]-[B,1] broadcast of model.py:231
 *                  reduces the REINFORCE coefficient of sample j to (iw_j - mean_i baseline_i) / B  
 *   inv_batch      1 / (global batch size); <= 0 selects 1 / B.  With batch sharding every rank passes 1 / (N * B) and the
 *                  gradient buffers are SUMMED by one all-reduce
 *   l2_weight      model.py:345-350 (2-D variables only)
 *   grad_params    [air_param_count] overwritten */
/*   (baseline_mean = NaN: the mean is read on the device from outs->scalars[AIR_S_MEAN_BASELINE] when the kernel runs --
 *    the value air_forward / air_elbo_scalars left there, or the all-reduced block under sharding; no host round trip) */
int32_t air_backward(air_handle* h, const float* params, const float* img, const float* eps_where,
                     const float* eps_what, const air_prior* prior, const air_outputs* outs, float baseline_mean,
                     float inv_batch, float l2_weight, float* grad_params, void* stream);
/* tf.train.RMSPropOptimizer(learning_rate, decay=.9, momentum=.9, epsilon=1e-10, centered=True) on a flat buffer
 * [upstream ApplyCenteredRMSProp]: mg <- mg + (1-decay)(g - mg); ms <- ms + (1-decay)(g^2 - ms);
 * mom <- momentum * mom + lr * g / sqrt(ms - mg^2 + epsilon); params <- params - mom.  Slots: ms starts at 1, mg and mom
 * at 0 (caller-owned).  g = grad * grad_scale. */
int32_t air_rmsprop_step(float* params, const float* grad, float* mg, float* ms, float* mom, int64_t n,
                         float learning_rate, float decay, float momentum, float epsilon, float grad_scale,
                         void* stream);

/* Backward of ONE dense layer y = act(x @ W + b) (neural.py:42-60) for callers that own their own MLP parameters -- the
 * BaselineMLP of modules.py:125-143, trained by its own optimiser (model.py:253-259,362-367):
 *   dW[K,N] += X^T @ dY, db[N] += colsum(dY)  (accumulated: zero them first), dX[M,K] = dY @ W^T (overwritten),
 *   multiplied by elu'(x) when elu_x (the forward value of X, itself an ELU layer's output) is given.
 * dY is the gradient with respect to the layer's PRE-activation.  Any of dW / db / dX may be NULL. */
int32_t air_linear_backward(const float* X, const float* W, const float* dY, const float* elu_x, float* dW, float* db,
                            float* dX, int32_t M, int32_t N, int32_t K, void* stream);
/* d baseline_loss / d baseline for baseline_loss = .5 mean((stop_gradient(target) - baseline)^2) with target [B] and
 * baseline [B,1] broadcasting to [B,B] (model.py:253-259): -(target_mean - baseline_i) * inv_batch. */
int32_t air_baseline_grad(const float* target, const float* baseline, float target_mean, float inv_batch,
                          float* d_baseline, int32_t B, void* stream);
/* The same with the target mean read from device memory when the kernel runs (e.g. outs->scalars + AIR_S_MEAN_IW after
 * air_elbo_scalars, or its all-reduced value under sharding): the training step needs no host round trip for it. */
int32_t air_baseline_grad_dev(const float* baseline, const float* target_mean_dev, float inv_batch, float* d_baseline,
                              int32_t B, void* stream);

/* BaselineMLP of modules.py:125-143 ON THE ENGINE: concat[img, what, where, presence, h, c] (batch-major) -> MLP(hidden, 1).
 *   air_baseline_attach    once per handle, BEFORE air_train_enable (the training workspace is sized for the widest
 *                          layer); the flat parameter / gradient layout is (w_0 [n_in, h_0], b_0 [h_0], ..., w_out
 *                          [h_last, 1], b_out [1]) -- air_baseline_param_count floats, air_baseline_input_width = n_in
 *   air_baseline_forward   gathers the input rows from `img` and the cell outputs in `outs` (what, where, presence,
 *                          final_h, final_c of the last air_forward) in one pass, runs the first (n_in x h_0) layer on
 *                          the handle's engine (tcgen05 split GEMM on an AIR_PREC_TC_SPLIT handle) and the small layers
 *                          on the fp32 GEMMs; baseline [B]
 *   air_baseline_backward  d baseline_loss / d parameters for the last air_baseline_forward given d loss / d baseline [B]
 *                          (air_baseline_grad / air_baseline_grad_dev); bgrad is overwritten.  The n_in x h_0 weight
 *                          gradient runs on the tensor-core split-K GEMM of the training workspace when that exists. */
int32_t air_baseline_attach(air_handle* h, int32_t n_hidden, const int32_t* hidden);
int64_t air_baseline_param_count(const air_handle* h);
int32_t air_baseline_input_width(const air_handle* h);
int32_t air_baseline_forward(air_handle* h, const float* bparams, const float* img, const air_outputs* outs, float* baseline,
                             void* stream);
int32_t air_baseline_backward(air_handle* h, const float* bparams, const float* d_baseline, float* bgrad, void* stream);
/* Same, but the weight-gradient GEMMs are left running on the handle's side streams: `bgrad` is complete only after the NEXT
 * air_backward on the same handle and stream has returned (its final join covers them).  The training step of
 * model.py:362-367 uses it so that the baseline's gradient overlaps the cell's backward pass. */
int32_t air_baseline_backward_async(air_handle* h, const float* bparams, const float* d_baseline, float* bgrad,
                                    void* stream);

/* Re-form the batch means in outs->scalars from the per-sample vectors an earlier air_forward left in
 * `outs`, now with a baseline[B] (BaselineMLP is evaluated on the cell outputs, so it can only be
 * known after the forward pass): AIRModel._reinforce, model.py:218-251. */
int32_t air_elbo_scalars(air_handle* h, const float* baseline, const air_prior* prior,
                         const air_outputs* outs, void* stream);

/* air_elbo_scalars without a handle (batch size given explicitly); used with air_prior_terms. */
int32_t air_elbo_scalars_raw(int32_t B, const float* baseline, const air_prior* prior, const air_outputs* outs,
                             void* stream);

/* The KL / step-count part of the loss on explicit posterior tensors (no images): AIRModel._prior_loss
 * (model.py:126-216) + NumStepsDistribution (prior.py:119-151).  Inputs are [T,B,.] device tensors; fills the
 * num_steps_posterior, prior_step_weight, kl_*_per_sample, num_step_per_sample and num_steps_log_prob members of
 * `outs` (rec_loss_per_sample is set to 0, loss_per_sample to the weighted prior terms).  Follow with
 * air_elbo_scalars() for the batch means / REINFORCE term.  Runs the same fused kernel as air_forward. */
int32_t air_prior_terms(int32_t B, int32_t T, int32_t na, const float* what_loc, const float* what_scale,
                        const float* where_loc, const float* where_scale, const float* presence_prob,
                        const float* presence, const air_prior* prior, const air_outputs* outs, void* stream);

/* One AIRCell step with explicit state, the RNNCore contract of cell.py:116-171:
 * state = [img, canvas, what, where, (h, c), presence]; `canvas`, `h`, `c`, `presence` are updated
 * in place; outputs (10 tensors of cell.py:167-168) are written as [B,.] (canvas NOT multiplied). */
int32_t air_cell_step(air_handle* h, const float* params, const float* img, float* canvas,
                      float* hstate, float* cstate, float* presence, const float* eps_where,
                      const float* eps_what, const float* u_pres, float* out_glimpse, float* out_what,
                      float* out_what_loc, float* out_what_scale, float* out_where,
                      float* out_where_loc, float* out_where_scale, float* out_presence_prob,
                      void* stream);

/* ---- building blocks (stand-alone, for unit parity) -------------------------------------------- */
/* snt.Linear + transfer (neural.py:42-60): out[M,N] = act(A[M,K] @ Wt[K,N] + bias[N]); act 0 none, 1 ELU */
int32_t air_linear(const float* A, const float* Wt, const float* bias, float* out, int32_t M, int32_t N,
                   int32_t K, int32_t act, int32_t precision, void* stream);
/* snt.LSTM step (mnist_model.py:35): gates = [x,h] @ W[nx+nh,4nh] + b, order i,j,f,o; h,c updated in place */
int32_t air_lstm_step(const float* x, float* hstate, float* cstate, const float* W, const float* b,
                      int32_t B, int32_t nx, int32_t nh, float forget_bias, void* stream);
/* SpatialTransformer (modules.py:94-109; cell.py:58,135): crop[B,h,w] from img[B,H,W], where[B,4] */
int32_t air_stn_read(const float* img, const float* where, float* crop, int32_t B, int32_t H, int32_t W,
                     int32_t h, int32_t w, void* stream);
/* inverse SpatialTransformer (modules.py:100-102; cell.py:59,159): out[B,H,W] gathered from glimpse[B,h,w] */
int32_t air_stn_paint(const float* glimpse, const float* where, float* out, int32_t B, int32_t H,
                      int32_t W, int32_t h, int32_t w, void* stream);

/* model.py:90  `self.glimpse = presence * sigmoid(glimpse)`: the visualisation tensor, [rows = T * B][G] from the decoded
 * glimpses and the presence column.  air_forward writes it into outs->glimpse_viz when that pointer is given; this entry
 * computes it on request (the reference's graph evaluates it only when the attribute is fetched). */
int32_t air_glimpse_viz(const float* glimpse, const float* presence, float* out, int64_t rows, int32_t G, void* stream);
/* prior.py:62-68: probs[n,T] -> pmf[n,T+1] (float64 island inside) */
int32_t air_bernoulli_to_modified_geometric(const float* probs, float* pmf, int64_t n, int32_t T,
                                            void* stream);
/* prior.py:26-32: prior[T+1]; is_f64 selects float64 (out is double*) or float32 (out is float*) maths */
int32_t air_geometric_prior(double success_prob, int32_t n_steps, int32_t is_f64, void* out, void* stream);
/* prior.py:71-90: kl[n,m] = float32(p * log(p / q)) where p > zero_prob_value else 0; q is [m] float64 */
int32_t air_tabular_kl(const float* p, const double* q, float* kl, int64_t n, int32_t m,
                       double zero_prob_value, void* stream);
/* prior.py:103-116 sample_from_tensor: out[i] = pmf[i, int32(samples[i])] (flat gather; index clamped to [0, m-1]) */
int32_t air_sample_from_tensor(const float* pmf, const float* samples, float* out, int64_t n, int32_t m,
                               void* stream);
/* prior.py:141-151: out[n] = log(max(pmf[i, int(samples[i])], 1e-32)) */
int32_t air_num_steps_log_prob(const float* pmf, const float* samples, float* out, int64_t n, int32_t m,
                               void* stream);
/* model.py:106-124 (host scalar, float64): anneal_type 0 = 'exp', 1 = 'linear' */
double air_anneal_weight(double init_val, double final_val, int32_t anneal_type, double global_step,
                         double anneal_steps, double hold_for, double steps_div);

#ifdef __cplusplus
}
#endif
#endif /* AIR_B200_H_ */
