"""AIRonMNIST -- mirror of the reference's mnist_model.py (mnist_model.py:10-44): hyper-parameter binding."""
from __future__ import annotations

from functools import partial

from .model import AIRModel
from .modules import LSTM, BaselineMLP, Decoder, Encoder, StepsPredictor, StochasticTransformParam


class AIRonMNIST(AIRModel):
    """Implements AIR for the MNIST dataset"""

    def __init__(self, obs, nums, glimpse_size=(20, 20),
                 inpt_encoder_hidden=[256] * 2,
                 glimpse_encoder_hidden=[256] * 2,
                 glimpse_decoder_hidden=[252] * 2,
                 transform_estimator_hidden=[256] * 2,
                 steps_pred_hidden=[50] * 1,
                 baseline_hidden=[256, 128] * 1,
                 transform_var_bias=-2.,
                 step_bias=0.,
                 *args, **kwargs):
        self.transform_var_bias = transform_var_bias
        self.step_bias = step_bias
        self.baseline = BaselineMLP(baseline_hidden)

        super(AIRonMNIST, self).__init__(
            *args,
            obs=obs,
            nums=nums,
            glimpse_size=glimpse_size,
            n_appearance=50,
            transition=LSTM(256),
            input_encoder=partial(Encoder, inpt_encoder_hidden),
            glimpse_encoder=partial(Encoder, glimpse_encoder_hidden),
            glimpse_decoder=partial(Decoder, glimpse_decoder_hidden),
            transform_estimator=partial(StochasticTransformParam, transform_estimator_hidden,
                                        scale_bias=self.transform_var_bias),
            steps_predictor=partial(StepsPredictor, steps_pred_hidden, self.step_bias),
            output_std=.3,
            **kwargs
        )
