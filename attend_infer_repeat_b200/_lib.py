"""ctypes binding of libair_b200.so (include/air_b200.h).

The library is built in-tree by :func:`build` (nvcc, sm_100a only) and loaded with ctypes: plain pointers and
sizes cross the boundary, torch only supplies device memory (``tensor.data_ptr()``) and the current stream.
There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(_HERE, "libair_b200.so")

AIR_MAX_HIDDEN = 4
AIR_MAX_STEPS = 8
AIR_N_SCALARS = 16
AIR_N_STAGES = 8
AIR_PREC_FP32 = 0
AIR_PREC_TC_SPLIT = 1

SCALAR_INDEX = dict(rec_loss=0, kl_num_steps=1, kl_what=2, kl_where=3, prior_loss=4, loss=5, reinforce_loss=6,
                    opt_loss=7, num_step=8, mean_iw_logq=9, mean_logq=10, mean_baseline=11, mean_iw=12, mean_iw2=13,
                    mean_baseline2=14)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class AirError(RuntimeError):
    pass


class air_config(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("T", C.c_int32), ("na", C.c_int32), ("nh", C.c_int32),
        ("n_enc_hidden", C.c_int32), ("enc_hidden", C.c_int32 * AIR_MAX_HIDDEN),
        ("n_glenc_hidden", C.c_int32), ("glenc_hidden", C.c_int32 * AIR_MAX_HIDDEN),
        ("n_dec_hidden", C.c_int32), ("dec_hidden", C.c_int32 * AIR_MAX_HIDDEN),
        ("n_where_hidden", C.c_int32), ("where_hidden", C.c_int32 * AIR_MAX_HIDDEN),
        ("n_steps_hidden", C.c_int32), ("steps_hidden", C.c_int32 * AIR_MAX_HIDDEN),
        ("output_std", C.c_float), ("output_multiplier", C.c_float), ("explore_eps", C.c_float),
        ("scale_bias", C.c_float), ("step_bias", C.c_float), ("what_scale_offset", C.c_float),
        ("forget_bias", C.c_float), ("max_crop_size", C.c_float),
        ("discrete_steps", C.c_int32), ("precision", C.c_int32),
    ]


class air_prior(C.Structure):
    _fields_ = [
        ("what_loc", C.c_float), ("what_scale", C.c_float),
        ("where_scale_loc", C.c_float), ("where_scale_scale", C.c_float),
        ("where_shift_loc", C.c_float), ("where_shift_scale", C.c_float),
        ("where_shift_has_loc", C.c_int32),
        ("steps_success_prob", C.c_double),
        ("steps_prob_is_f64", C.c_int32),
        ("steps_weight", C.c_float),
        ("analytic", C.c_int32), ("use_prior", C.c_int32), ("use_reinforce", C.c_int32),
        ("nvil_shift", C.c_float), ("nvil_scale", C.c_float),
    ]


OUTPUT_FIELDS = ["canvas", "glimpse", "glimpse_viz", "what", "what_loc", "what_scale", "where", "where_loc",
                 "where_scale", "presence_prob", "presence", "final_h", "final_c", "num_steps_posterior",
                 "num_step_per_sample", "prior_step_weight", "rec_loss_per_sample", "kl_num_steps_per_sample",
                 "kl_what_per_sample", "kl_where_per_sample", "loss_per_sample", "num_steps_log_prob", "scalars"]


class air_outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUTPUT_FIELDS]


# every symbol include/air_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "air_abi_version": (C.c_int32, []),
    "air_last_error": (C.c_char_p, []),
    "air_create": (C.c_int32, [C.POINTER(air_config), C.POINTER(_P)]),
    "air_destroy": (C.c_int32, [_P]),
    "air_param_count": (C.c_int64, [_P]),
    "air_param_entries": (C.c_int32, [_P]),
    "air_param_entry": (C.c_int32, [_P, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "air_workspace_bytes": (C.c_int64, [_P]),
    "air_row_schedule_check": (C.c_int32, [C.POINTER(air_config), C.c_char_p, C.c_int32]),
    "air_launch_count": (C.c_int64, [_P]),
    "air_profile_enable": (C.c_int32, [_P, C.c_int32]),
    "air_profile_read": (C.c_int32, [_P, _P, C.c_int32]),
    "air_stage_name": (C.c_char_p, [C.c_int32]),
    "air_check_range": (C.c_int32, [_P, _P]),
    "air_forward": (C.c_int32, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(air_prior), C.POINTER(air_outputs), _P]),
    "air_forward_host": (C.c_int32, [_P, _P, _P, _P, _P, _P, C.POINTER(air_prior), C.POINTER(air_outputs), _P, _P,
                                     _P]),
    "air_forward_host_u8": (C.c_int32, [_P, _P, _P, _P, _P, _P, C.POINTER(air_prior), C.POINTER(air_outputs), _P, _P,
                                        _P]),
    "air_cache_weights": (C.c_int32, [_P, C.c_int32]),
    "air_set_launch_overlap": (C.c_int32, [_P, C.c_int32]),
    "air_params_updated": (C.c_int32, [_P]),
    "air_prior_table_device": (C.c_int32, [_P, C.POINTER(air_prior), _P]),
    "air_train_enable": (C.c_int32, [_P, C.c_int32]),
    "air_train_workspace_bytes": (C.c_int64, [_P]),
    "air_backward": (C.c_int32, [_P, _P, _P, _P, _P, C.POINTER(air_prior), C.POINTER(air_outputs), C.c_float, C.c_float,
                                 C.c_float, _P, _P]),
    "air_rmsprop_step": (C.c_int32, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_float, _P]),
    "air_draw_noise": (C.c_int32, [_P, C.c_uint64, _P, _P, _P, _P]),
    "air_forward_host_u8_rng": (C.c_int32, [_P, _P, _P, C.c_uint64, C.POINTER(air_prior), C.POINTER(air_outputs), _P, _P,
                                            _P]),
    "air_feed_host_u8": (C.c_int32, [_P, C.c_int32, _P]),
    "air_forward_fed_u8_rng": (C.c_int32, [_P, _P, C.c_int32, C.c_uint64, C.POINTER(air_prior), C.POINTER(air_outputs), _P,
                                           _P, _P]),
    "air_feed_wait": (C.c_int32, [_P, C.c_int32]),
    "air_forward_dataset_u8": (C.c_int32, [_P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(air_prior),
                                           C.POINTER(air_outputs), _P, _P]),
    "air_gather_u8": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "air_linear_backward": (C.c_int32, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "air_baseline_grad": (C.c_int32, [_P, _P, C.c_float, C.c_float, _P, C.c_int32, _P]),
    "air_baseline_grad_dev": (C.c_int32, [_P, _P, C.c_float, _P, C.c_int32, _P]),
    "air_baseline_attach": (C.c_int32, [_P, C.c_int32, C.POINTER(C.c_int32)]),
    "air_baseline_param_count": (C.c_int64, [_P]),
    "air_baseline_input_width": (C.c_int32, [_P]),
    "air_baseline_forward": (C.c_int32, [_P, _P, _P, C.POINTER(air_outputs), _P, _P]),
    "air_baseline_backward": (C.c_int32, [_P, _P, _P, _P, _P]),
    "air_baseline_backward_async": (C.c_int32, [_P, _P, _P, _P, _P]),
    "air_elbo_scalars": (C.c_int32, [_P, _P, C.POINTER(air_prior), C.POINTER(air_outputs), _P]),
    "air_elbo_scalars_raw": (C.c_int32, [C.c_int32, _P, C.POINTER(air_prior), C.POINTER(air_outputs), _P]),
    "air_prior_terms": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, C.POINTER(air_prior),
                                    C.POINTER(air_outputs), _P]),
    "air_iwae_bound": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [_P] * 9 + [C.POINTER(air_prior), _P, _P,
                                                                                              _P, _P]),
    "air_cell_step": (C.c_int32, [_P] * 19),
    "air_linear": (C.c_int32, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "air_lstm_step": (C.c_int32, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P]),
    "air_stn_read": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "air_stn_paint": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "air_glimpse_viz": (C.c_int32, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    "air_bernoulli_to_modified_geometric": (C.c_int32, [_P, _P, C.c_int64, C.c_int32, _P]),
    "air_geometric_prior": (C.c_int32, [C.c_double, C.c_int32, C.c_int32, _P, _P]),
    "air_tabular_kl": (C.c_int32, [_P, _P, _P, C.c_int64, C.c_int32, C.c_double, _P]),
    "air_sample_from_tensor": (C.c_int32, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    "air_num_steps_log_prob": (C.c_int32, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    "air_anneal_weight": (C.c_double, [C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_double,
                                       C.c_double]),
}


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _newest_mtime(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "air_b200.h")]
    return _newest_mtime(deps) > os.path.getmtime(LIB_PATH)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libair_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise AirError("nvcc not found: cannot build libair_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-o", LIB_PATH] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise AirError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load (building first if the sources are newer) and type the shared library.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        build()
    if not os.path.exists(LIB_PATH):
        raise AirError(f"{LIB_PATH} is missing and could not be built; there is no fallback path")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(handle, name)       # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if handle.air_abi_version() != 1:
        raise AirError("libair_b200.so ABI version mismatch")
    _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().air_last_error()
        raise AirError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device (or host) pointer of a contiguous float tensor, or NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    return C.c_void_p(t.data_ptr())


def current_stream_ptr():
    """torch's current stream on the current device as a cudaStream_t (torch._C entry point: ~0.3 us instead of the
    ~15 us of torch.cuda.current_stream(), which matters for the launch-bound small-batch training loop)."""
    import torch
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:      # private entry point moved: fall back to the public (slower) API
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
