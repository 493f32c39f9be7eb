#!/usr/bin/env python
"""bench.py -- AIR cell-steps/sec (batch x N_steps) on synthetic 50x50 multi-MNIST-shaped canvases.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B] [--precision fp32|tc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: T=3 unrolled AIRCell steps + post-processing + every ELBO term
(air_forward through the C ABI) on B=4096 canvases per GPU (BASELINE.json configs[1]); all ten AIRCell outputs are
materialised as [T,B,.] float32 like the reference's dynamic_rnn does.  One rank per GPU, batches sharded (weak
scaling), the 16 loss scalars all-reduced over NCCL every step when N > 1.  Rank 0 prints ONE JSON line.

`--impl reference` times the reference algorithm on the host cores instead: the reference is TF1/Sonnet/Python-2 and
cannot run in this image, so this is its torch-CPU restatement (oracle/air_oracle.py, encoder recomputed every step
as written in cell.py:125), with every host thread.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "AIR cell-steps/sec (batch x N_steps), fused AIRCell forward + ELBO, 50x50 multi-MNIST"
UNIT = "cell-steps/s"
SCRIPT_SHAPE = dict(H=50, W=50, h=20, w=20, T=3, na=50, nh=256)
SHAPE = SCRIPT_SHAPE          # (kept for the tests that import it)

# BASELINE.json configs[0..4].  `batch` is per GPU; `mode` selects what a "step" is:
#   forward  one pass of the hot path (T cell steps + every ELBO term) over the batch           -> `value`
#   train    one FULL training step through the public API (AIRonMNIST.train_step -> train_op): forward, BaselineMLP,
#            backward, gradient all-reduce, both centered-RMSProp updates                        -> `value`
#   iwae     forward over batch x K particle rows + the importance-weighted bound               -> `value` (particle-steps)
CONFIGS = {
    "c1": dict(index=0, shape=SCRIPT_SHAPE, batch=64, mode="forward",
               name="multi-MNIST 50x50, max_steps=3, batch=64 (scripts/train_multi_mnist.sh) -- plumbing/parity"),
    "c2": dict(index=1, shape=SCRIPT_SHAPE, batch=4096, mode="forward",
               name="multi-MNIST 50x50, max_steps=3, batch=4096 per GPU, fused AIRCell forward+ELBO"),
    "c3": dict(index=2, shape=SCRIPT_SHAPE, batch=4096, mode="train",
               name="multi-MNIST 50x50, max_steps=3, batch=4096 per GPU (32768 on 8), full training step through the "
                    "public API (AIRonMNIST.train_step), NCCL gradient all-reduce"),
    "c4": dict(index=3, shape=dict(H=100, W=100, h=28, w=28, T=5, na=50, nh=256), batch=2048, mode="forward",
               name="multi-MNIST 100x100 canvas, 28x28 glimpse, max_steps=5, batch=2048 per GPU -- STN-bandwidth-bound regime"),
    "c5": dict(index=4, shape=SCRIPT_SHAPE, batch=256, mode="iwae", K=5,
               name="IWAE K=5 importance-weighted ELBO, max_steps=3, 256 canvases x 5 particles per GPU (batch 1024 on 4)"),
}


def executed_macs_per_sample(cfg, T):
    """Dense MACs this implementation EXECUTES per canvas (encoder hoisted, LSTM input half hoisted) -- DESIGN.md."""
    def chain(n_in, widths):
        m, d = 0, n_in
        for n in widths:
            m += d * n
            d = n
        return m, d
    enc, n_enc = chain(cfg.H * cfg.W, cfg.enc_hidden)
    gx = n_enc * 4 * cfg.nh
    rec = cfg.nh * 4 * cfg.nh
    where, _ = chain(cfg.nh, list(cfg.where_hidden) + [8])
    steps, _ = chain(cfg.nh, list(cfg.steps_hidden) + [1])
    glenc, n_gl = chain(cfg.h * cfg.w, cfg.glenc_hidden)
    what = n_gl * 2 * cfg.na
    dec, _ = chain(cfg.na, list(cfg.dec_hidden) + [cfg.h * cfg.w])
    per_stage = dict(input_encoder=enc, lstm=gx + T * rec, where_mlp=T * where, steps_presence=T * steps,
                     glimpse_enc=T * (glenc + what), decoder=T * dec)
    return per_stage, sum(per_stage.values())


def algorithmic_bytes_per_sample(cfg, T):
    """image in + noise in + all ten API outputs out, float32."""
    P, G, na = cfg.H * cfg.W, cfg.h * cfg.w, cfg.na
    noise = T * (4 + na + 1) * 4
    outs = T * (P + G + 3 * na + 3 * 4 + 2) * 4
    return P * 4 + noise + outs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 8] or \
               [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------------------------
def oracle_setup(args):
    """The oracle's view of the selected configuration, on a bounded sample of the per-GPU batch."""
    from oracle import air_oracle as O
    conf = CONFIGS[args.config]
    ocfg = O.AirConfig(**conf["shape"])
    cap = {"c4": 256}.get(args.config, 1024)       # ~10-30 s of host work for the whole timed loop
    B = min(args.batch, cap)
    pc = O.PriorConfig()
    params = O.init_params(ocfg, 0)
    img, _ = O.synthetic_multi_mnist(B, ocfg.H, ocfg.W, seed=1)
    K = conf.get("K", 1)
    if K > 1:
        img = img.repeat_interleave(K, 0)
    noise = O.make_noise(ocfg, B * K, 1)
    return O, conf, ocfg, pc, params, img, noise, B, K


def oracle_step(O, conf, ocfg, pc, params, img, noise, K):
    """One step of the selected configuration on the host: forward (+ IWAE bound, + autograd backward for `train`)."""
    if conf["mode"] == "train":
        p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        res = O.forward(ocfg, pc, p, img, *noise, global_step=20000)
        res["opt_loss"].backward()
        return res
    with torch.no_grad():
        res = O.forward(ocfg, pc, params, img, *noise, global_step=20000)
        if K > 1:
            O.iwae_bound(ocfg, pc, res, K, global_step=20000)
    return res


def run_reference(args, rank):
    """The reference algorithm on the host cores (torch-CPU restatement; TF1 cannot run here)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, conf, ocfg, pc, params, img, noise, B, K = oracle_setup(args)
    for _ in range(args.warmup):
        oracle_step(O, conf, ocfg, pc, params, img, noise, K)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(O, conf, ocfg, pc, params, img, noise, K)
    dt = time.perf_counter() - t0
    value = B * K * ocfg.T * args.steps / dt
    sample = (f"{B} of {args.batch} canvases per step" + (f" x {K} particles" if K > 1 else "") + f", {args.steps} steps, "
              f"fp32 torch-CPU, encoder not hoisted" + (", forward + autograd backward (no optimiser)" if conf["mode"] == "train" else ""))
    cfgd = workload_config(args, args.gpus)
    cfgd["reference_sample"] = sample
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, n):
    conf = CONFIGS[args.config]
    sh = conf["shape"]
    K = conf.get("K", 1)
    out_mb = args.batch * K * sh["T"] * (sh["H"] * sh["W"] + sh["h"] * sh["w"] + 3 * sh["na"] + 14) * 4 / 1e6
    return {"workload": f"{conf['name']} (BASELINE.json configs[{conf['index']}])",
            "bench_config": args.config, "mode": conf["mode"],
            "global_batch": args.batch * n, "canvas": f"{sh['H']}x{sh['W']}", "glimpse": f"{sh['h']}x{sh['w']}",
            "max_steps": sh["T"], "particles": K,
            "outputs": "all 10 AIRCell outputs materialised [T,B,.] fp32 + per-sample ELBO terms (the model-level visualisation "
                       "tensor presence * sigmoid(glimpse), model.py:90, is computed on request as in the reference's graph)",
            "parallelism": f"dp{n}", "precision": args.precision,
            "arithmetic": ("fp32-equivalent: every GEMM operand carried as an fp16 hi/lo pair, three tcgen05 MMAs per product, fp32 "
                           "accumulation in tensor memory; the per-canvas stages in fp32" if args.precision == "tc" else
                           "fp32 FMA on CUDA cores"),
            "streams": (args.streams if conf["mode"] != "train" else 1),
            "batches_in_flight": (f"{args.streams} independent batches of {args.batch} canvases, one per CUDA stream / handle "
                                  "(EnginePool); every step is one full pass over one batch"
                                  if conf["mode"] != "train" and args.streams > 1 else "one batch at a time"),
            "weights": ("updated every step (the tensor-core operand arena is rebuilt from the fp32 parameters each pass)"
                        if conf["mode"] == "train" else
                        "constant over the timed loop; tensor-core operand arena prepared once (air_cache_weights)"),
            "l2": f"{args.input_sets} rotating input sets ({args.input_sets * args.batch * K * sh['H'] * sh['W'] * 4 / 1e6:.0f} MB of "
                  f"images) + {out_mb:.0f} MB of outputs written per step (L2 is 126 MB)",
            "reference_sample": "the --impl reference arm times min(batch, 1024) canvases per step (256 for c4) and "
                                "normalises per canvas"}


def cpu_baseline(args, device=None, precision=None):
    """Oracle port timed on the host cores of the GPU box, bounded to ~10-30 s.  With `device`, the same sample also goes
    through the CUDA engine once and the line gets the second half of BASELINE.json's metric, "ELBO delta vs ref"."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, conf, ocfg, pc, params, img, noise, B, K = oracle_setup(args)
    with torch.no_grad():
        ref = O.forward(ocfg, pc, params, img, *noise, global_step=20000)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
        oracle_step(O, conf, ocfg, pc, params, img, noise, K)
        n += 1
    dt = time.perf_counter() - t0
    base = {"value": B * K * ocfg.T * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/air_oracle.py (torch-CPU fp32 restatement of the TF1 path; TF1 cannot run here), "
                      f"{B} canvases" + (f" x {K} particles" if K > 1 else "") + f" x {n} steps in {dt:.1f} s"
                      + (", forward + autograd backward" if conf["mode"] == "train" else "")}
    delta = None
    if device is not None:
        delta = _elbo_delta(O, ocfg, pc, params, img, noise, ref, img.shape[0], device, precision)
    return base, delta


TAU = 0.02   # conditioning threshold of tests/test_gpu_full_batch.py: painted steps with |s_x| or |s_y| below it are ill-conditioned


def _elbo_delta(O, ocfg, pc, params, img, noise, ref, B, device, precision):
    """CUDA engine vs the oracle on the cpu_baseline sample (the checker's own leg: the oracle is not on the timed path).
    The canvas is compared under the explicit conditioning rule of tests/test_gpu_full_batch.py: 1e-4 (abs + rel) on every
    canvas whose painted steps all have min(|s_x|, |s_y|) >= TAU; the canvases below TAU are counted and reported apart
    (their error is the where code's ~1e-6 rounding amplified by 1 / |s|, for the fp32 oracle as much as for the kernel)."""
    import attend_infer_repeat_b200 as air
    eng = None
    try:
        ccfg = air.CellConfig(H=ocfg.H, W=ocfg.W, h=ocfg.h, w=ocfg.w, na=ocfg.na, nh=ocfg.nh, precision=precision)
        T = ocfg.T
        eng = air.Engine(ccfg, B, T, device=device)
        pr = air.make_prior(dict(loc=pc.what_loc, scale=pc.what_scale),
                            dict(loc=pc.where_scale_loc, scale=pc.where_scale_scale),
                            dict(loc=pc.where_shift_loc, scale=pc.where_shift_scale),
                            float(O.steps_prior_success_prob(pc, 20000)), True)
        out = eng.forward(O.flatten_params(ocfg, params).to(device), img.to(device).contiguous(),
                          *(t.to(device).contiguous() for t in noise), pr)
        torch.cuda.synchronize()
        elbo, elbo_ref = -float(out["scalars"][air._lib.SCALAR_INDEX["loss"]]), float(ref["elbo"])
        lps, lps_ref = out["loss_per_sample"].cpu().double(), ref["loss_per_sample"].double()
        canvas = out["canvas"].cpu().reshape(T, B, -1).double()
        canvas_ref = ref["canvas"].reshape(T, B, -1).double()
        pres_equal = bool(torch.equal(out["presence"].reshape(-1).cpu(), ref["outs"]["presence"].reshape(-1)))
        scale = max(1.0, float(lps_ref.abs().mean()))
        # conditioning of every canvas: smallest |s| over its painted steps
        where = ref["outs"]["where"].reshape(T, B, 4)
        pres = ref["outs"]["presence"].reshape(T, B)
        s_min = torch.minimum(where[..., 0].abs(), where[..., 2].abs())
        s_min = torch.where(pres > 0, s_min, torch.full_like(s_min, 1e9)).min(0).values
        good = s_min >= TAU
        c_err = (canvas - canvas_ref).abs()
        outside = c_err > 1e-4 + 1e-4 * canvas_ref.abs()
        # the fp32 reference is itself only defined up to its rounding noise: the same sample through the oracle in float64
        with torch.no_grad():
            r64 = O.forward(ocfg, pc, {k: v.double() for k, v in params.items()}, img.double(),
                            *(t.double() for t in noise), global_step=20000)
        c64, l64 = r64["canvas"].reshape(T, B, -1), r64["loss_per_sample"]
        floor = {"oracle_fp32_vs_fp64_canvas_max_abs": float((canvas_ref - c64).abs().max()),
                 "cuda_vs_fp64_canvas_max_abs": float((canvas - c64).abs().max()),
                 "oracle_fp32_vs_fp64_loss_per_sample_max_abs": float((lps_ref - l64).abs().max()),
                 "cuda_vs_fp64_loss_per_sample_max_abs": float((lps - l64).abs().max())}
        return {"elbo_cuda": elbo, "elbo_oracle": elbo_ref, "rel": abs(elbo - elbo_ref) / abs(elbo_ref),
                "loss_per_sample_max_abs": float((lps - lps_ref).abs().max()), "loss_per_sample_mean_magnitude": scale,
                "loss_per_sample_max_rel_to_mean_magnitude": float((lps - lps_ref).abs().max()) / scale,
                "conditioning_rule": f"canvases whose painted steps all have min(|s_x|, |s_y|) >= {TAU}",
                "canvases": B, "canvases_below_tau": int((~good).sum()),
                "canvas_max_abs_well_conditioned": float(c_err[:, good].max()) if bool(good.any()) else 0.0,
                "canvas_elements_outside_1e-4_abs_plus_1e-4_rel": int(outside[:, good].sum()),
                "canvas_max_abs_below_tau": float(c_err[:, ~good].max()) if bool((~good).any()) else 0.0,
                "canvas_elements_outside_below_tau": int(outside[:, ~good].sum()),
                "canvas_max_magnitude": float(canvas_ref.abs().max()),
                "presence_bit_exact": pres_equal, "tolerance": 1e-4, "float64_floor": floor,
                "sample": f"the {B} rows of cpu_baseline, same weights / images / noise on both sides"}
    except Exception as e:                      # never lose the bench line over the accuracy report
        return {"error": f"{type(e).__name__}: {e}"}
    finally:
        if eng is not None:
            eng.close()


# ------------------------------------------------------------------------------------------------------------
def _max_over_ranks(ms, dist, dev):
    if dist is None:
        return ms
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_native(args, rank, local_rank, world):
    import attend_infer_repeat_b200 as air
    from attend_infer_repeat_b200.cell import _init_flat
    from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # several ranks on one host: each binds to its GPU's local CPUs BEFORE any pinned buffer exists (NUMA-local feed)
    bound_cpus = None
    if world > 1 and os.environ.get("AIR_NO_CPU_BIND") is None:
        from attend_infer_repeat_b200.sharding import bind_host_to_device
        bound_cpus = bind_host_to_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    conf = CONFIGS[args.config]
    mode, sh, K = conf["mode"], conf["shape"], conf.get("K", 1)
    prec = air.AIR_PREC_TC_SPLIT if args.precision == "tc" else air.AIR_PREC_FP32
    cfg = air.CellConfig(H=sh["H"], W=sh["W"], h=sh["h"], w=sh["w"], na=sh["na"], nh=sh["nh"], precision=prec)
    T, B = sh["T"], args.batch
    R = B * K                                  # rows of one pass (canvases x particles)
    prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.),
                           air.functional.anneal_weight(1 - 1e-15, 1e-7, "exp", 20000, 1e5, 1e3, 1e4), True)

    # synthetic multi-MNIST-shaped inputs: several distinct resident sets so successive steps never re-read L2-hot data
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    base_u8 = torch.from_numpy(synthetic_multi_mnist_u8(256, sh["H"], sh["W"], seed=rank)[0])      # dataset format: uint8
    sets, host_u8 = [], []
    for s_ in range(args.input_sets):
        idx = torch.randint(0, base_u8.shape[0], (B,), generator=torch.Generator().manual_seed(s_))
        u8 = base_u8[idx].contiguous()
        img = (u8.to(torch.float32) / 255.0).to(dev).contiguous()                          # load_data, data.py:116
        if K > 1:
            img = img.repeat_interleave(K, 0).contiguous()                                 # K particles of a canvas = K rows
        sets.append((img, torch.randn(T, R, 4, device=dev, generator=g), torch.randn(T, R, cfg.na, device=dev, generator=g),
                     torch.rand(T, R, 1, device=dev, generator=g)))
        host_u8.append(u8.pin_memory())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_steps, join=None):
        """exactly n_steps timed calls of fn(i): device time by CUDA events on the launching stream, max over ranks.
        ``join`` makes the launching stream wait for the side streams the steps ran on (EnginePool) before the closing event;
        every such stream starts by waiting for the launching stream, i.e. for the opening event."""
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(n_steps):
            fn(i)
        if join is not None:
            join()
        ev1.record()
        barrier()
        return _max_over_ranks(ev0.elapsed_time(ev1), dist, dev)

    hbm_peak, tf_peak, peak_kind = measured_peaks()
    macs, total_macs = executed_macs_per_sample(cfg, T)
    alg_bytes = algorithmic_bytes_per_sample(cfg, T)
    # the two rooflines of one pass over R rows (bytes / FLOPs per canvas x canvases per launch set)
    t_hbm_ms = R * alg_bytes / (hbm_peak * 1e9) * 1e3
    t_tensor_ms = 2.0 * R * total_macs / (tf_peak * 1e12) * 1e3

    extra = {}
    sampler = ClockSampler(local_rank)

    # =====================================================================================================
    if mode in ("forward", "iwae"):
        # Independent batches are enqueued round-robin on args.streams engines / CUDA streams (EnginePool): every step is
        # still one full pass over one batch of R rows, but kernels of neighbouring batches may co-run.  --streams 1 is the
        # one-batch-at-a-time latency figure, also measured below and reported as "single_stream".
        pool = air.EnginePool(cfg, R, T, n_streams=args.streams, device=dev, materialise_viz=False)
        pool.cache_weights(True)   # forward-only loop with constant parameters: the fp16-split weight arena is built once
        eng = pool.engines[0]
        params, _ = _init_flat(air.param_spec(cfg), dev, seed=0)
        pending = [None]

        # The only cross-rank exchange of forward + ELBO is 16 floats per step.  The blocks of len(pool) consecutive steps are
        # copied into one staging tensor and all-reduced together from a stream of their own, so neither the collective's ~30 us
        # latency nor its join ever sits on a stream that runs passes (two staging tensors alternate; a staging tensor is
        # rewritten only after the event that follows its previous all-reduce).
        S_ = len(pool)
        comm = torch.cuda.Stream(device=dev) if dist is not None else None
        stage = [torch.zeros(S_, air._lib.AIR_N_SCALARS, device=dev) for _ in range(2)]
        done = [None, None]          # event after the last all-reduce of each staging tensor
        evs = []

        def step(i):
            img, ew, ea, u = sets[i % len(sets)]
            j, g = i % S_, (i // S_) % 2
            with pool.next() as e:
                out = e.forward(params, img, ew, ea, u, prior)
                if K > 1:
                    e.iwae_bound(K, prior)
                if dist is not None:
                    if done[g] is not None:
                        torch.cuda.current_stream().wait_event(done[g])
                    stage[g][j].copy_(out["scalars"])
                    ev = torch.cuda.Event()
                    ev.record()
                    evs.append(ev)
            if dist is not None and j == S_ - 1:
                for ev in evs:
                    comm.wait_event(ev)
                evs.clear()
                with torch.cuda.stream(comm):
                    if pending[0] is not None:
                        pending[0].wait()
                    pending[0] = dist.all_reduce(stage[g], async_op=True)
                    pending[0].wait()                      # the COMM stream waits; the host and the pass streams do not
                    done[g] = torch.cuda.Event()
                    done[g].record()
            return out

        def join_all():
            pool.join()
            if comm is not None:
                torch.cuda.current_stream().wait_stream(comm)

        def step_single(i):
            img, ew, ea, u = sets[i % len(sets)]
            eng.forward(params, img, ew, ea, u, prior)
            if K > 1:
                eng.iwae_bound(K, prior)

        for i in range(max(args.warmup, 3) * len(pool)):
            step(i)
        join_all()
        barrier()
        launches0 = sum(e.launch_count for e in pool.engines)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        t_wall0 = time.time()
        ms = timed(step, args.steps, join=join_all)
        t_wall1 = time.time()
        launches = sum(e.launch_count for e in pool.engines) - launches0
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        value = world * R * T * args.steps / (ms * 1e-3)
        elbo = -float(eng.scalar("loss"))
        eng.set_launch_overlap(True)     # from here on engine 0 runs alone: dependent-launch overlap back on (the pool turns it off)
        if len(pool) > 1:
            n1 = max(50, args.steps // 4)
            for i in range(3):
                step_single(i)
            ms1 = timed(step_single, n1)
            eng.set_launch_overlap(False)
            extra["single_stream"] = {"ms_per_step": ms1 / n1, "value": world * R * T * n1 / (ms1 * 1e-3), "steps": n1,
                                      "note": "one batch at a time on one stream, dependent-launch overlap on (latency of a pass); the headline value "
                                              f"keeps {len(pool)} independent batches in flight on {len(pool)} streams"}

        # ---- end to end through the C ABI with HOST buffers (uint8 images in, loss out, every step) ----------
        scal2 = [torch.empty(air._lib.AIR_N_SCALARS).pin_memory() for _ in range(2)]
        lps2 = [torch.empty(R).pin_memory() for _ in range(2)]
        if K == 1:
            def run_fed(n):
                acc_, n_out = 0.0, 0
                for scal_h, lps_h in pool.stream_host_u8(params, (host_u8[i % len(host_u8)] for i in range(n)), prior, 1000):
                    acc_ += float(scal_h[0])                          # the host reads every step's loss
                    n_out += 1
                assert n_out == n
                return acc_
            run_fed(4 * len(pool))
            barrier()
            t0 = time.perf_counter()
            run_fed(args.steps)
            torch.cuda.synchronize()
            fed_ms = _max_over_ranks((time.perf_counter() - t0) * 1e3, dist, dev)

            def sync_step(i):
                eng.forward_host_u8_rng(params, host_u8[i % len(host_u8)], 1000 + i, prior, scal2[0], lps2[0])
            for i in range(3):
                sync_step(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                sync_step(i)
            torch.cuda.synchronize()
            sync_ms = _max_over_ranks((time.perf_counter() - t0) * 1e3, dist, dev)
            d2h = (scal2[0].numel() + lps2[0].numel()) * 4
            e2e = {"value": world * R * T * args.steps / (fed_ms * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": host_u8[0].numel(), "d2h_bytes_per_step": d2h, "ms_per_step": fed_ms / args.steps,
                   "api": "EnginePool.stream_host_u8 = air_feed_host_u8 + air_forward_fed_u8_rng + air_feed_wait per batch, "
                          f"round-robin over {len(pool)} handles / streams (double-buffered feed per handle: pinned host uint8 "
                          "images in on a copy stream while earlier batches are processed, in-library Philox noise, loss "
                          "scalars + per-sample loss out and read on the host every step, in order; host wall clock)",
                   "synchronous": {"value": world * R * T * args.steps / (sync_ms * 1e-3), "ms_per_step": sync_ms / args.steps,
                                   "h2d_bytes_per_step": host_u8[0].numel(),
                                   "api": "air_forward_host_u8_rng (copy in, pass, copy out, host synchronisation, one call "
                                          "per step)"}}
        else:
            # IWAE: the K particle rows of a canvas are replicated on the device from ONE uploaded uint8 canvas
            stage_u8 = torch.empty(B, sh["H"], sh["W"], dtype=torch.uint8, device=dev)
            mean_h = torch.empty(1).pin_memory()

            def e2e_step(i):
                stage_u8.copy_(host_u8[i % len(host_u8)], non_blocking=True)
                img = (stage_u8.to(torch.float32) / 255.0).repeat_interleave(K, 0)
                _, ew, ea, u = sets[i % len(sets)]
                eng.forward(params, img, ew, ea, u, prior)
                m, _, _ = eng.iwae_bound(K, prior)
                mean_h.copy_(m.reshape(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return float(mean_h[0])
            for i in range(3):
                e2e_step(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                e2e_step(i)
            e_ms = _max_over_ranks((time.perf_counter() - t0) * 1e3, dist, dev)
            e2e = {"value": world * R * T * args.steps / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": host_u8[0].numel(),
                   "d2h_bytes_per_step": 4, "ms_per_step": e_ms / args.steps,
                   "api": "pinned uint8 canvases -> device, K-fold particle rows, Engine.forward + Engine.iwae_bound, the mean "
                          "bound read on the host every step"}

        # ---- per-stage device time of the hot path (CUDA events on the launching stream, separate pass) ----------
        eng.set_launch_overlap(True)     # the stage times are those of one pass alone
        eng.profile(True)
        acc, n_prof = {}, 5
        for i in range(n_prof):
            img, ew, ea, u = sets[i % len(sets)]
            eng.forward(params, img, ew, ea, u, prior)
            for k, v in eng.stage_times_ms().items():
                acc[k] = acc.get(k, 0.0) + v / n_prof
        eng.profile(False)
        step_ms = ms / args.steps
        tighter = "hbm" if t_hbm_ms >= t_tensor_ms else "tensor"
        t_roof = max(t_hbm_ms, t_tensor_ms)
        P, G = cfg.P, cfg.G
        paint_bytes = R * (T * P * 4 + P * 4 + T * G * 4)     # canvases out, image + decoded glimpses in
        dom = max(acc, key=acc.get)
        gl_ms = acc["glimpse_enc"] + acc["decoder"]
        gl_flops = 2.0 * R * (macs["glimpse_enc"] + macs["decoder"])
        per_kernel = {
            "paint_elbo": {"bound": "hbm", "algorithmic_bytes": paint_bytes, "ms": acc["paint_elbo"],
                           "achieved_gbs": paint_bytes / (acc["paint_elbo"] * 1e-3) / 1e9,
                           "frac": paint_bytes / (acc["paint_elbo"] * 1e-3) / 1e9 / hbm_peak},
            "glimpse_vae_row_kernel": {"bound": "tensor", "useful_flops": gl_flops, "ms": gl_ms,
                                       "achieved_tflops": gl_flops / (gl_ms * 1e-3) / 1e12},
            "lstm_cluster": {"bound": "tensor", "useful_flops": 2.0 * R * macs["lstm"], "ms": acc["lstm"],
                             "achieved_tflops": 2.0 * R * macs["lstm"] / (acc["lstm"] * 1e-3) / 1e12},
            "input_encoder": {"bound": "tensor", "useful_flops": 2.0 * R * macs["input_encoder"], "ms": acc["input_encoder"],
                              "achieved_tflops": 2.0 * R * macs["input_encoder"] / (acc["input_encoder"] * 1e-3) / 1e12},
        }
        for k_ in ("glimpse_vae_row_kernel", "lstm_cluster", "input_encoder"):
            per_kernel[k_]["frac"] = per_kernel[k_]["achieved_tflops"] / tf_peak
        roofline = {
            # WHOLE-STEP fraction against the tighter of the two rooflines of one pass (the number north_star asks for)
            "bound": tighter, "scope": "whole fused step (all launches of one pass)",
            "achieved": (R * alg_bytes / (step_ms * 1e-3) / 1e9) if tighter == "hbm" else
                        (2.0 * R * total_macs / (step_ms * 1e-3) / 1e12),
            "peak": hbm_peak if tighter == "hbm" else tf_peak, "unit": "GB/s" if tighter == "hbm" else "TFLOP/s",
            "frac": t_roof / step_ms,
            "peak_kind": f"{'STREAM copy' if tighter == 'hbm' else 'bf16 dense sustained'}, {peak_kind}",
            "algorithmic_bytes_per_step": R * alg_bytes, "useful_flops_per_step": 2.0 * R * total_macs,
            "t_hbm_ms": t_hbm_ms, "t_tensor_ms": t_tensor_ms, "ms_per_step": step_ms,
            # dram__bytes_read + write summed over the launches of one pass, ncu --set full (profiles/r02p_full.md), c2 only
            "traffic": TRAFFIC_C2 if (args.config == "c2" and prec == air.AIR_PREC_TC_SPLIT) else None,
            "traffic_unit": "bytes per step (all launches)",
            "dominant_stage": dom, "stage_ms": {k: round(v, 4) for k, v in acc.items()},
            "stage_share": {k: round(v / sum(acc.values()), 4) for k, v in acc.items()},
            "per_kernel_note": "tensor entries count useful (fp32-equivalent) FLOPs; the fp16 hi/lo split issues 3 MMAs per "
                               "useful one, so 1/3 of the peak is the ceiling of those fractions",
            "per_kernel": per_kernel}

        # ---- kernel-level training step (engine calls only: forward with kept activations, backward, all-reduce,
        #      centered RMSProp; no BaselineMLP -- the full step through the public API is --config c3) ----------
        if mode == "forward" and not args.no_train:
            extra["train_step"] = kernel_train_loop(air, cfg, prec, R, T, dev, sets, params, prior, world, dist, barrier,
                                                    max(20, args.steps // 5))
        pool.close()

    # =====================================================================================================
    else:   # mode == "train": the reference's train step through the public API
        nums = torch.zeros(3, B, 1, device=dev)
        model = air.AIRonMNIST(sets[0][0], nums, max_steps=T, explore_eps=1e-3, inpt_encoder_hidden=[256, 256],
                               glimpse_encoder_hidden=[256, 256], glimpse_decoder_hidden=[256, 256],
                               transform_estimator_hidden=[256, 256], steps_pred_hidden=[128, 64], baseline_hidden=[256, 128],
                               transform_var_bias=.5, step_bias=.75, output_multiplier=.5, precision=prec, seed=0)
        pr = dict(loc=0., scale=1.)
        nsp = dict(anneal='exp', init=1. - 1e-15, final=1e-7, steps_div=1e4, steps=1e5, hold_init=1e3, analytic=True)
        train_op, _ = model.train_step(1e-5, 0., pr, pr, pr, nsp, cuda_graph=not args.no_train_graph)
        model.global_step = 20000
        torch.cuda.manual_seed(1234 + rank)

        def step(i):
            img, ew, ea, u = sets[i % len(sets)]
            return train_op(img, None, (ew, ea, u))

        for i in range(max(args.warmup, 3)):
            step(i)
        barrier()
        launches0, replays0 = model.engine.launch_count, model.graph_replays
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        t_wall0 = time.time()
        ms = timed(step, args.steps)
        t_wall1 = time.time()
        # (a replayed step launches its kernels as graph nodes: the library's launch counter does not see them)
        launches = (model.engine.launch_count - launches0) + (model.graph_replays - replays0) * model.graph_launches_per_step
        extra["cuda_graph"] = {"replays": model.graph_replays - replays0, "kernel_nodes_per_step": model.graph_launches_per_step}
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        value = world * B * T * args.steps / (ms * 1e-3)
        elbo = -float(model.engine.scalar("loss"))
        step_ms = ms / args.steps

        # e2e: pinned uint8 batch in, /255 on the device, the step, the loss read on the host -- every step.  The batch of
        # step i + 1 travels on a copy stream while step i computes (two staging buffers); the float32 image is written
        # straight into the model's input tensor (model.obs, the reference's `obs` placeholder), so train_op() takes no copy.
        stage_u8 = [torch.empty(B, sh["H"], sh["W"], dtype=torch.uint8, device=dev) for _ in range(2)]
        staged = [torch.cuda.Event() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        loss_h = torch.empty(1).pin_memory()

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                stage_u8[i % 2].copy_(host_u8[i % len(host_u8)], non_blocking=True)
                staged[i % 2].record(copy_stream)

        def e2e_step(i):
            prefetch(i + 1)       # (its staging buffer was last read by step i - 1, which has completed: see the synchronize)
            torch.cuda.current_stream().wait_event(staged[i % 2])
            torch.div(stage_u8[i % 2], 255.0, out=model.obs)
            train_op()
            loss_h.copy_(model.engine.scalar("loss").reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(loss_h[0])
        prefetch(0)
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(3, 3 + args.steps):
            e2e_step(i)
        e_ms = _max_over_ranks((time.perf_counter() - t0) * 1e3, dist, dev)
        e2e = {"value": world * B * T * args.steps / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": host_u8[0].numel(),
               "d2h_bytes_per_step": 4, "ms_per_step": e_ms / args.steps,
               "api": "AIRonMNIST.train_step -> train_op(): pinned uint8 batch -> device on a copy stream (double-buffered, one "
                      "batch ahead), /255 into model.obs, forward + BaselineMLP + backward + gradient all-reduce + two "
                      "centered-RMSProp updates, noise drawn on the device, the loss read on the host and a stream "
                      "synchronisation every step (host wall clock)"}
        # roofline of a training step: ~3x the forward's useful FLOPs (forward, input gradients, weight gradients)
        roofline = {"bound": "tensor", "scope": "whole training step (all launches)",
                    "achieved": 3.0 * 2.0 * B * total_macs / (step_ms * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": 3.0 * t_tensor_ms / step_ms, "peak_kind": f"bf16 dense sustained, {peak_kind}",
                    "useful_flops_per_step": 3.0 * 2.0 * B * total_macs, "ms_per_step": step_ms, "traffic": None,
                    "note": "useful FLOPs = 3 x forward (forward, input gradients, weight gradients), BaselineMLP excluded; "
                            "every useful MMA is issued three times (fp16 / bf16 hi-lo split)"}
        extra["train_step_kernel_loop"] = kernel_train_loop(air, cfg, prec, B, T, dev, sets, model.params.clone(), prior, world,
                                                            dist, barrier, max(20, args.steps // 5))
        extra["allreduce_bytes_per_step"] = ((model.engine.n_params + model.baseline_module.params.numel()) * 4
                                             if world > 1 else 0)
        extra["baseline_mlp"] = "3177->256->128->1, forward + backward + its own RMSProp at 10x lr (model.py:253-259,362-367)"

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": None, "elbo_delta_vs_oracle": None, "elbo": elbo}
        line.update(extra)
        line["host_binding"] = (f"rank bound to the {len(bound_cpus)} NVML-local CPUs of its GPU (sched_setaffinity)"
                                if bound_cpus else "none")
        if world == 1:
            line["cpu_baseline"], line["elbo_delta_vs_oracle"] = cpu_baseline(args, dev, prec)
        emit(line)
    if mode == "train":
        model.release_graphs()      # (the captured step holds NCCL kernels: before the process group goes away)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum over the launches of ONE c2 pass (profiles/r02p_full.md, ncu --set full):
# enc1 43.6 + encoder layer 2 4.5 + lstm 6.5 + heads 13.4 + where_read 42.4 + glimpse row kernel 24.3 + paint/ELBO 156.0
# + scalars 0.1 MB (the canvases' last 33 MB are still dirty in L2 when the pass ends)
TRAFFIC_C2 = 290.8e6


def kernel_train_loop(air, cfg, prec, B, T, dev, sets, params, prior, world, dist, barrier, n_train):
    """forward (activations kept) + air_backward + gradient all-reduce + centered RMSProp through Engine calls only."""
    teng = air.Engine(air.CellConfig(H=cfg.H, W=cfg.W, h=cfg.h, w=cfg.w, na=cfg.na, nh=cfg.nh, precision=prec), B, T,
                      device=dev)
    teng.train_enable(True)
    tparams = params.clone()
    n = tparams.numel()
    grad = torch.empty(n, device=dev)
    mg, ms_, mom = torch.zeros(n, device=dev), torch.ones(n, device=dev), torch.zeros(n, device=dev)

    def train_step(i):
        img, ew, ea, u = sets[i % len(sets)]
        teng.forward(tparams, img, ew, ea, u, prior)
        teng.backward(tparams, img, ew, ea, prior, grad, inv_batch=1.0 / (world * B))
        if dist is not None:
            dist.all_reduce(grad)
        teng.rmsprop_step(tparams, grad, mg, ms_, mom, 1e-5)

    for i in range(3):
        train_step(i)
    barrier()
    l0 = teng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(n_train):
        train_step(i)
    ev1.record()
    barrier()
    tms = _max_over_ranks(ev0.elapsed_time(ev1), dist, dev)
    res = {"value": world * B * T * n_train / (tms * 1e-3), "unit": UNIT, "ms_per_step": tms / n_train,
           "steps": n_train, "global_batch": world * B,
           "engine": ("tcgen05 split engine: fp16 hi/lo forward layers, bf16 hi/lo gradient GEMMs"
                      if prec == air.AIR_PREC_TC_SPLIT else "AIR_PREC_FP32 forward (SIMT), tcgen05 bf16 hi/lo gradient GEMMs"),
           "gpu_launches_per_step": (teng.launch_count - l0) / n_train + 1,
           "allreduce_bytes_per_step": n * 4 if world > 1 else 0,
           "what": "forward+ELBO (activations kept) + backward + gradient all-reduce + centered RMSProp; no BaselineMLP",
           "train_workspace_mb": round(teng.train_workspace_bytes / 1e6, 1),
           "final_loss": float(teng.scalar("loss"))}
    teng.close()
    return res


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # libraries print to fd 1 behind our back (NCCL's version banner at communicator creation): keep the original stdout
    # for the JSON line and point fd 1 at stderr for everything else
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configs[0..4] = c1..c5")
    ap.add_argument("--batch", type=int, default=None, help="canvases per GPU (default: the configuration's)")
    ap.add_argument("--precision", default="tc", choices=["fp32", "tc"])
    ap.add_argument("--input-sets", type=int, default=4)
    ap.add_argument("--streams", type=int, default=4,
                    help="forward / IWAE configs: independent batches in flight (EnginePool); 1 = one pass at a time")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement")
    ap.add_argument("--no-train-graph", action="store_true",
                    help="c3: enqueue every training step eagerly instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = CONFIGS[args.config]["batch"]
    if args.steps is None:
        args.steps = {"forward": 1000, "iwae": 1000, "train": 300}[CONFIGS[args.config]["mode"]]
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: spawn one rank per GPU ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
