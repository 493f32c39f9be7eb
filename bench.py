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
SHAPE = dict(H=50, W=50, h=20, w=20, T=3, na=50, nh=256)


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
def run_reference(args, rank):
    """The reference algorithm on the host cores (torch-CPU restatement; TF1 cannot run here)."""
    if rank != 0:
        return
    from oracle import air_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ocfg = O.AirConfig(**SHAPE)
    B = min(args.batch, 1024)          # bounded sample: one step = the first `B` canvases of the 4096-batch workload
    pc = O.PriorConfig()
    params = O.init_params(ocfg, 0)
    img, _ = O.synthetic_multi_mnist(B, ocfg.H, ocfg.W, seed=1)
    noise = O.make_noise(ocfg, B, 1)
    with torch.no_grad():
        for _ in range(args.warmup):
            O.forward(ocfg, pc, params, img, *noise, global_step=20000)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.forward(ocfg, pc, params, img, *noise, global_step=20000)
        dt = time.perf_counter() - t0
    value = B * ocfg.T * args.steps / dt
    sample = f"{B} of {args.batch} canvases per step, {args.steps} steps, fp32 torch-CPU, encoder not hoisted"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, n):
    return {"workload": f"multi-MNIST 50x50, max_steps=3, batch={args.batch} per GPU, fused AIRCell forward+ELBO "
                        f"(BASELINE.json configs[1])",
            "global_batch": args.batch * n, "canvas": "50x50", "glimpse": "20x20", "max_steps": 3,
            "outputs": "all 10 AIRCell outputs materialised [T,B,.] fp32 + per-sample ELBO terms",
            "parallelism": f"dp{n}", "precision": args.precision,
            "weights": "constant over the timed loop; tensor-core operand arena prepared once (air_cache_weights)",
            "l2": f"{args.input_sets} rotating input sets ({args.input_sets * args.batch * 10000 / 1e6:.0f} MB of "
                  f"images) + {args.batch * 3 * 11.6e3 / 1e6:.0f} MB of outputs written per step > 126 MB L2"}


def cpu_baseline(args, device=None, precision=None):
    """Oracle port timed on the host cores of the GPU box, bounded to ~10-30 s.  With `device`, the same sample also goes
    through the CUDA engine once and the line gets the second half of BASELINE.json's metric, "ELBO delta vs ref"."""
    from oracle import air_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ocfg = O.AirConfig(**SHAPE)
    B = min(args.batch, 1024)
    pc = O.PriorConfig()
    params = O.init_params(ocfg, 0)
    img, _ = O.synthetic_multi_mnist(B, ocfg.H, ocfg.W, seed=1)
    noise = O.make_noise(ocfg, B, 1)
    with torch.no_grad():
        ref = O.forward(ocfg, pc, params, img, *noise, global_step=20000)
        n, t0 = 0, time.perf_counter()
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
            O.forward(ocfg, pc, params, img, *noise, global_step=20000)
            n += 1
        dt = time.perf_counter() - t0
    base = {"value": B * ocfg.T * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/air_oracle.py (torch-CPU fp32 restatement of the TF1 path; TF1 cannot run here), "
                      f"{B} canvases x {n} steps in {dt:.1f} s"}
    delta = None
    if device is not None:
        delta = _elbo_delta(O, ocfg, pc, params, img, noise, ref, B, device, precision)
    return base, delta


def _elbo_delta(O, ocfg, pc, params, img, noise, ref, B, device, precision):
    """CUDA engine vs the oracle on the cpu_baseline sample (the checker's own leg: the oracle is not on the timed path)."""
    import attend_infer_repeat_b200 as air
    eng = None
    try:
        eng = air.Engine(air.CellConfig(precision=precision), B, ocfg.T, device=device)
        pr = air.make_prior(dict(loc=pc.what_loc, scale=pc.what_scale),
                            dict(loc=pc.where_scale_loc, scale=pc.where_scale_scale),
                            dict(loc=pc.where_shift_loc, scale=pc.where_shift_scale),
                            float(O.steps_prior_success_prob(pc, 20000)), True)
        out = eng.forward(O.flatten_params(ocfg, params).to(device), img.to(device).contiguous(),
                          *(t.to(device).contiguous() for t in noise), pr)
        torch.cuda.synchronize()
        elbo, elbo_ref = -float(out["scalars"][air._lib.SCALAR_INDEX["loss"]]), float(ref["elbo"])
        lps, lps_ref = out["loss_per_sample"].cpu().double(), ref["loss_per_sample"].double()
        canvas, canvas_ref = out["canvas"].cpu().reshape(-1), ref["canvas"].reshape(-1)
        pres_equal = bool(torch.equal(out["presence"].reshape(-1).cpu(), ref["outs"]["presence"].reshape(-1)))
        # the tests' criteria (tests/test_gpu_parity.py::_check_forward): per-sample sums range over several hundred and
        # change sign across the batch, so "relative 1e-4" is taken against the batch's mean magnitude of the term;
        # element-wise tensors are held to 1e-4 absolute + 1e-4 relative
        scale = max(1.0, float(lps_ref.abs().mean()))
        c_err = (canvas.double() - canvas_ref.double()).abs()
        # the fp32 reference is itself only defined up to its rounding noise (near-singular sampled scales amplify it through
        # 1 / s_x in the inverse transformer): the same sample through the oracle in float64 gives the floor
        with torch.no_grad():
            r64 = O.forward(ocfg, pc, {k: v.double() for k, v in params.items()}, img.double(),
                            *(t.double() for t in noise), global_step=20000)
        c64, l64 = r64["canvas"].reshape(-1), r64["loss_per_sample"]
        floor = {"oracle_fp32_vs_fp64_canvas_max_abs": float((canvas_ref.double() - c64).abs().max()),
                 "cuda_vs_fp64_canvas_max_abs": float((canvas.double() - c64).abs().max()),
                 "oracle_fp32_vs_fp64_loss_per_sample_max_abs": float((lps_ref - l64).abs().max()),
                 "cuda_vs_fp64_loss_per_sample_max_abs": float((lps - l64).abs().max())}
        return {"elbo_cuda": elbo, "elbo_oracle": elbo_ref, "rel": abs(elbo - elbo_ref) / abs(elbo_ref),
                "loss_per_sample_max_abs": float((lps - lps_ref).abs().max()), "loss_per_sample_mean_magnitude": scale,
                "loss_per_sample_max_rel_to_mean_magnitude": float((lps - lps_ref).abs().max()) / scale,
                "canvas_max_abs": float(c_err.max()), "canvas_max_magnitude": float(canvas_ref.abs().max()),
                "canvas_elements_outside_1e-4_abs_plus_1e-4_rel": int((c_err > 1e-4 + 1e-4 * canvas_ref.abs().double()).sum()),
                "presence_bit_exact": pres_equal, "tolerance": 1e-4, "float64_floor": floor,
                "sample": f"the {B} canvases of cpu_baseline, same weights / images / noise on both sides"}
    except Exception as e:                      # never lose the bench line over the accuracy report
        return {"error": f"{type(e).__name__}: {e}"}
    finally:
        if eng is not None:
            eng.close()


# ------------------------------------------------------------------------------------------------------------
def run_native(args, rank, local_rank, world):
    import attend_infer_repeat_b200 as air
    from attend_infer_repeat_b200.data import synthetic_multi_mnist_u8

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    prec = air.AIR_PREC_TC_SPLIT if args.precision == "tc" else air.AIR_PREC_FP32
    cfg = air.CellConfig(precision=prec)
    T, B = SHAPE["T"], args.batch
    eng = air.Engine(cfg, B, T, device=dev)
    eng.cache_weights(True)        # forward-only loop with constant parameters: the fp16-split weight arena is built once
    spec = air.param_spec(cfg)
    from attend_infer_repeat_b200.cell import _init_flat
    params, _ = _init_flat(spec, dev, seed=0)
    prior = air.make_prior(dict(loc=0., scale=1.), dict(loc=0., scale=1.), dict(loc=0., scale=1.),
                           air.functional.anneal_weight(1 - 1e-15, 1e-7, "exp", 20000, 1e5, 1e3, 1e4), True)

    # synthetic multi-MNIST-shaped inputs: several distinct resident sets so successive steps never re-read L2-hot data
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    base_u8 = torch.from_numpy(synthetic_multi_mnist_u8(256, 50, 50, seed=rank)[0])      # dataset format: uint8
    sets, host_u8 = [], []
    for s in range(args.input_sets):
        idx = torch.randint(0, base_u8.shape[0], (B,), generator=torch.Generator().manual_seed(s))
        u8 = base_u8[idx].contiguous()
        img = (u8.to(torch.float32) / 255.0).to(dev).contiguous()                          # load_data, data.py:116
        sets.append((img, torch.randn(T, B, 4, device=dev, generator=g), torch.randn(T, B, cfg.na, device=dev, generator=g),
                     torch.rand(T, B, 1, device=dev, generator=g)))
        host_u8.append(u8.pin_memory())
    # pinned host copies for the end-to-end arm
    host = [tuple(t.cpu().pin_memory() for t in s) for s in sets[:2]]
    scal_h = torch.empty(air._lib.AIR_N_SCALARS).pin_memory()
    lps_h = torch.empty(B).pin_memory()

    def step(i):
        img, ew, ea, u = sets[i % len(sets)]
        out = eng.forward(params, img, ew, ea, u, prior)
        if dist is not None:
            dist.all_reduce(out["scalars"])          # the only cross-rank exchange of forward+ELBO: 16 floats
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * T * args.steps / (ms * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers (H2D of images + noise, D2H of the loss) -------------
    # Headline: images in the reference's dataset format (uint8 [B,50,50], data.py:35-107), /255 on the device.
    # Also reported: the same call fed float32 images (what load_data hands to the TF graph, data.py:116).
    def e2e_step_u8(i):
        _, ew, ea, u = host[i % len(host)]
        eng.forward_host_u8(params, host_u8[i % len(host)], ew, ea, u, prior, scal_h, lps_h)

    def e2e_step_u8_rng(i):
        eng.forward_host_u8_rng(params, host_u8[i % len(host)], 1000 + i, prior, scal_h, lps_h)

    def e2e_step_f32(i):
        img, ew, ea, u = host[i % len(host)]
        eng.forward_host(params, img, ew, ea, u, prior, scal_h, lps_h)

    def time_e2e(fn):
        for i in range(3):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            fn(i)
        torch.cuda.synchronize()
        ms_ = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            t = torch.tensor([ms_], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_

    # double-buffered feed: the H2D copy of batch i+1 (copy stream) overlaps the pass over batch i; the loss of batch i-1 is
    # read on the host while batch i runs.  Every step still moves its own 10 MB batch in and its own loss out.
    scal2 = [torch.empty(air._lib.AIR_N_SCALARS).pin_memory() for _ in range(2)]
    lps2 = [torch.empty(B).pin_memory() for _ in range(2)]

    def time_e2e_fed():
        def run(n):
            acc_ = 0.0
            eng.feed_host_u8(0, host_u8[0])
            for i in range(n):
                if i + 1 < n:
                    eng.feed_host_u8((i + 1) % 2, host_u8[(i + 1) % len(host_u8)])
                eng.forward_fed_u8_rng(params, i % 2, 1000 + i, prior, scal2[i % 2], lps2[i % 2])
                if i >= 1:
                    eng.feed_wait((i - 1) % 2)
                    acc_ += float(scal2[(i - 1) % 2][0])          # the host reads every step's loss
            eng.feed_wait((n - 1) % 2)
            return acc_ + float(scal2[(n - 1) % 2][0])
        run(4)
        barrier()
        t0 = time.perf_counter()
        run(args.steps)
        torch.cuda.synchronize()
        ms_ = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            t = torch.tensor([ms_], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_

    e2e_fed_ms = time_e2e_fed()
    e2e_rng_ms = time_e2e(e2e_step_u8_rng)
    e2e_ms = time_e2e(e2e_step_u8)
    e2e_f32_ms = time_e2e(e2e_step_f32)
    noise_bytes = sum(t.numel() * 4 for t in host[0][1:])
    d2h = (scal_h.numel() + lps_h.numel()) * 4
    # headline: what sess.run(train_step, feed_dict={imgs}) moves in the reference -- the uint8 image batch in, the loss out;
    # where / what / presence noise is drawn inside the library like the reference's in-graph draws (cell.py:133,147,156)
    e2e = {"value": world * B * T * args.steps / (e2e_fed_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": host_u8[0].numel(), "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_fed_ms / args.steps,
           "api": "air_feed_host_u8 + air_forward_fed_u8_rng + air_feed_wait (double-buffered feed: pinned host uint8 "
                  "images in on a copy stream while the previous batch is processed, in-library Philox noise, loss scalars "
                  "+ per-sample loss out and read on the host every step; host wall clock)",
           "synchronous": {"value": world * B * T * args.steps / (e2e_rng_ms * 1e-3),
                           "ms_per_step": e2e_rng_ms / args.steps, "h2d_bytes_per_step": host_u8[0].numel(),
                           "api": "air_forward_host_u8_rng (copy in, pass, copy out, host synchronisation, one call per "
                                  "step)"},
           "host_noise": {"value": world * B * T * args.steps / (e2e_ms * 1e-3), "ms_per_step": e2e_ms / args.steps,
                          "h2d_bytes_per_step": host_u8[0].numel() + noise_bytes,
                          "api": "air_forward_host_u8 (images + pre-drawn float32 noise from the host)"},
           "f32_images": {"value": world * B * T * args.steps / (e2e_f32_ms * 1e-3), "ms_per_step": e2e_f32_ms / args.steps,
                          "h2d_bytes_per_step": host[0][0].numel() * 4 + noise_bytes, "api": "air_forward_host"}}

    # ---- per-stage device time of the hot path (CUDA events on the launching stream, separate pass) ----------
    eng.profile(True)
    acc = {}
    n_prof = 5
    for i in range(n_prof):
        step(i)
        for k, v in eng.stage_times_ms().items():
            acc[k] = acc.get(k, 0.0) + v / n_prof
    eng.profile(False)
    macs, total_macs = executed_macs_per_sample(cfg, T)
    hbm_peak, tf_peak, peak_kind = measured_peaks()
    if prec == air.AIR_PREC_FP32:
        # stages that contain ONLY dense-layer launches
        gemm_stages, mac_keys = ["input_encoder", "where_mlp", "decoder"], ["input_encoder", "where_mlp", "decoder"]
        engine_name = "linear_simt_kernel (fp32 FMA)"
    else:
        # the two fused-chain launches (chain_tc.cuh): heads (where + steps MLPs) and glimpse VAE (encoder, what, decoder);
        # with chains on, the `where_mlp` and `glimpse_enc` stage intervals hold exactly one chain_kernel launch each
        gemm_stages = ["where_mlp", "glimpse_enc"]
        mac_keys = ["where_mlp", "steps_presence", "glimpse_enc", "decoder"]
        engine_name = "chain_kernel (tcgen05 TS-form fp16x2 split, activations resident in TMEM)"
    gemm_ms = sum(acc[s] for s in gemm_stages)
    gemm_flops = 2.0 * B * sum(macs[s] for s in mac_keys)
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    stage_share = {k: round(v / sum(acc.values()), 4) for k, v in acc.items()}
    roofline = {"bound": "tensor", "kernel": engine_name, "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": achieved / tf_peak, "peak_kind": f"bf16 dense sustained, {peak_kind}",
                # dram__bytes_read.sum + dram__bytes_write.sum of the two chain_kernel launches of one pass at B=4096, from
                # the committed `ncu --set full` capture (profiles/r01c_full_tc.md: 13.36 + 23.80 MB read, 0.07 MB written;
                # the 37 MB of outputs they produce are still in the 126 MB L2 when the kernels end)
                "traffic": 37.23e6 if (prec == air.AIR_PREC_TC_SPLIT and B == 4096) else None,
                "traffic_unit": "bytes per launch set",
                "flops_per_launch_set": gemm_flops, "ms_per_launch_set": gemm_ms,
                "stages_timed": gemm_stages, "stage_ms": {k: round(v, 4) for k, v in acc.items()},
                "stage_share": stage_share,
                "whole_step": {"executed_tflops": 2.0 * B * total_macs / (ms / args.steps * 1e-3) / 1e12,
                               "algorithmic_gbs": B * algorithmic_bytes_per_sample(cfg, T) / (ms / args.steps * 1e-3) / 1e9,
                               "hbm_peak_gbs": hbm_peak}}
    paint_bytes = B * (T * cfg.P * 4 + cfg.P * 4 + 2 * T * cfg.G * 4)
    roofline["paint_elbo_gbs"] = paint_bytes / (acc["paint_elbo"] * 1e-3) / 1e9

    # ---- full training step (BASELINE.json configs[2] per GPU): forward + ELBO with saved activations, backward,
    # one all-reduce of the flat gradient buffer (N > 1), centered RMSProp -- same engine as the forward arm
    train = None
    if not args.no_train:
        teng = air.Engine(air.CellConfig(precision=prec), B, T, device=dev)
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
        n_train = max(5, args.steps // 5)
        ev0.record()
        for i in range(n_train):
            train_step(i)
        ev1.record()
        barrier()
        tms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([tms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tms = float(t.item())
        train = {"value": world * B * T * n_train / (tms * 1e-3), "unit": UNIT, "ms_per_step": tms / n_train,
                 "steps": n_train, "global_batch": world * B, "engine": ("tcgen05 split engine: fp16 hi/lo forward layers, bf16 hi/lo gradient GEMMs" if prec == air.AIR_PREC_TC_SPLIT
                            else "AIR_PREC_FP32 forward (SIMT), tcgen05 bf16 hi/lo gradient GEMMs"),
                 "gpu_launches_per_step": (teng.launch_count - l0) / n_train + 1,
                 "allreduce_bytes_per_step": n * 4 if world > 1 else 0,
                 "what": "forward+ELBO (activations kept) + backward + gradient all-reduce + centered RMSProp",
                 "train_workspace_mb": round(teng.train_workspace_bytes / 1e6, 1),
                 "final_loss": float(teng.scalar("loss"))}
        teng.close()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "train_step": train, "cpu_baseline": None, "elbo_delta_vs_oracle": None,
                "elbo": -float(eng.scalar("loss")) / world}
        if world == 1:
            line["cpu_baseline"], line["elbo_delta_vs_oracle"] = cpu_baseline(args, dev, prec)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


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
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="canvases per GPU")
    ap.add_argument("--precision", default="tc", choices=["fp32", "tc"])
    ap.add_argument("--input-sets", type=int, default=4)
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement")
    args = ap.parse_args()
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
