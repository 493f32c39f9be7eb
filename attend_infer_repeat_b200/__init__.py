"""attend_infer_repeat_b200 -- B200-native (sm_100a) implementation of the AIR per-step inference/generation hot
path behind the reference's Python call surface (AIRCell / AIRModel / AIRonMNIST / NumStepsDistribution).

Everything numerical runs in libair_b200.so (hand-written CUDA behind a C ABI, include/air_b200.h); torch supplies
device memory, streams and torch.distributed.  There is no CPU or library fallback: importing works anywhere (so the
host logic can be tested), but any compute call without the CUDA library and a B200 raises.
"""
from . import evaluation, functional, ops, prior  # noqa: F401
from ._lib import AIR_PREC_FP32, AIR_PREC_TC_SPLIT, AirError, build  # noqa: F401
from .cell import AIRCell  # noqa: F401
from .data import ResidentDataset, load_data, save_data, tensors_from_data  # noqa: F401
from .engine import CellConfig, Engine, EnginePool, c_config, make_prior, param_count, param_spec, row_schedule_check  # noqa: F401
from .mnist_model import AIRonMNIST  # noqa: F401
from .model import AIRModel  # noqa: F401
from .modules import (LSTM, BaselineMLP, Decoder, Encoder, ParametrisedGaussian, SpatialTransformer,  # noqa: F401
                      StepsPredictor, StochasticTransformParam, TransformParam)
from .neural import MLP  # noqa: F401
from .ops import Loss, clip_preserve  # noqa: F401
from .prior import NumStepsDistribution, bernoulli_to_modified_geometric, geometric_prior, tabular_kl  # noqa: F401
