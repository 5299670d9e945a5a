"""Exponential moving average of activation scales.

Mirror of quant/utils/moving_average.py:12-39 (same buffers: num_batches_tracked, momentum,
moving_average) so checkpoints load unchanged.  The first tracked batch is copied, later ones are
blended as momentum * old + (1 - momentum) * new.  Unlike the reference there is no ``.item()``
host synchronisation: the first-call test uses a host-side counter mirrored into the buffer.
"""
import torch
import torch.nn as nn


class MovingAverage(nn.Module):
    def __init__(self, momentum: torch.Tensor) -> None:
        super().__init__()
        self.register_buffer('num_batches_tracked', torch.tensor(0))
        self.register_buffer('momentum', momentum)
        self.register_buffer('moving_average', torch.zeros(len(momentum)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        if self.training:
            with torch.no_grad():
                seen = (self.num_batches_tracked > 0).to(x.dtype)
                blend = self.momentum * self.moving_average + (torch.ones_like(self.momentum) - self.momentum) * x
                self.moving_average.copy_(seen * blend + (1 - seen) * x)
                self.num_batches_tracked += 1
        return self.moving_average
