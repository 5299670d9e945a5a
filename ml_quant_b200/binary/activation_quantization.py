"""Activation quantizer modules: one scale (set) per sample, optional moving average.

Mirror of quant/binary/activation_quantization.py (MovingAverageMode :19-28, the policy in
ActivationQuantizer.forward :68-102, LS1 :117, LS2 :148, LST :179, GF :210).  ``off``: scales are
solved on every call, train or eval.  ``eval_only``: tracked while training, used in eval.
``train_and_eval``: tracked and used while training, used in eval.
"""
from enum import Enum
from typing import List, Tuple

import torch
import torch.nn as nn

from . import quantization
from ..utils.moving_average import MovingAverage


class MovingAverageMode(Enum):
    off = 'off'
    eval_only = 'eval_only'
    train_and_eval = 'train_and_eval'


class ActivationQuantizer(nn.Module):
    """Base class; subclasses supply the per-batch solve and the fixed-scale quantization."""

    scheme = ''

    def __init__(self, num_scaling_factors: int, moving_average_mode: str = 'off',
                 moving_average_momentum: float = 0.99) -> None:
        super().__init__()
        self.num_scaling_factors = num_scaling_factors
        self.moving_avg_module = MovingAverage(torch.tensor([moving_average_momentum] * num_scaling_factors))
        self.moving_average_mode = MovingAverageMode(moving_average_mode)

    def stored_scales(self, batch: int) -> List[torch.Tensor]:
        avg = self.moving_avg_module.moving_average
        return [avg[i].expand(batch) for i in range(avg.size(0))]

    def uses_stored_scales(self) -> bool:
        if self.training:
            return self.moving_average_mode == MovingAverageMode.train_and_eval
        return self.moving_average_mode != MovingAverageMode.off

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        if not self.training:
            if self.moving_average_mode != MovingAverageMode.off:
                return self._moving_average_quantization(x, self.stored_scales(x.shape[0]))
            return self._batch_quantization(x)[1]
        batch_vs, x_q = self._batch_quantization(x)
        if self.moving_average_mode != MovingAverageMode.off:
            tracked = self.moving_avg_module(batch_vs.mean(1))
            if self.moving_average_mode == MovingAverageMode.train_and_eval:
                vs = [tracked[i].expand(x.shape[0]) for i in range(self.num_scaling_factors)]
                x_q = self._moving_average_quantization(x, vs)
        return x_q

    def _batch_quantization(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def _moving_average_quantization(self, x: torch.Tensor, vs: List[torch.Tensor]) -> torch.Tensor:
        raise NotImplementedError


class ActivationQuantizerLS1(ActivationQuantizer):
    scheme = 'ls-1'

    def __init__(self, moving_average_mode: str = 'off', moving_average_momentum: float = 0.99) -> None:
        super().__init__(1, moving_average_mode, moving_average_momentum)

    def _batch_quantization(self, x):
        v1, x_q = quantization.quantizer_ls_1(x)
        return v1.view(1, -1), x_q

    def _moving_average_quantization(self, x, vs):
        return quantization.quantizer_ls_1(x, vs[0])[1]


class ActivationQuantizerLS2(ActivationQuantizer):
    scheme = 'ls-2'

    def __init__(self, moving_average_mode: str = 'off', moving_average_momentum: float = 0.99) -> None:
        super().__init__(2, moving_average_mode, moving_average_momentum)

    def _batch_quantization(self, x):
        v1, v2, x_q = quantization.quantizer_ls_2(x)
        return torch.stack([v1, v2]), x_q

    def _moving_average_quantization(self, x, vs):
        return quantization.quantizer_ls_2(x, vs[0], vs[1])[2]


class ActivationQuantizerLST(ActivationQuantizer):
    scheme = 'ls-T'

    def __init__(self, moving_average_mode: str = 'off', moving_average_momentum: float = 0.99) -> None:
        super().__init__(1, moving_average_mode, moving_average_momentum)

    def _batch_quantization(self, x):
        v1, x_q = quantization.quantizer_ls_ternary(x)
        return v1.view(1, -1), x_q

    def _moving_average_quantization(self, x, vs):
        return quantization.quantizer_ls_ternary(x, vs[0])[1]


class ActivationQuantizerGF(ActivationQuantizer):
    def __init__(self, k: int, moving_average_mode: str = 'off', moving_average_momentum: float = 0.99) -> None:
        super().__init__(k, moving_average_mode, moving_average_momentum)
        self.k = k
        self.scheme = f'gf-{k}'

    def _batch_quantization(self, x):
        vs, x_q = quantization.quantizer_gf(x, self.k)
        return torch.stack(vs), x_q

    def _moving_average_quantization(self, x, vs):
        return quantization.quantizer_gf(x, self.k, vs)[1]
