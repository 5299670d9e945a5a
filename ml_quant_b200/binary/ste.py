"""Sign with sign(0) = +1 and its straight-through estimator.

Mirror of quant/binary/ste.py (binary_sign :16-18, STESign :21-66, binarize :70); forward and
backward run as CUDA kernels behind the C ABI (lsq_fakequant with a unit scale, lsq_ste_backward).
"""
from typing import Any

import torch
from torch.autograd import Function

from .. import ops


def binary_sign(x: torch.Tensor) -> torch.Tensor:
    """-1 where x < 0, +1 where x >= 0, as float32."""
    ops.require_cuda(x)
    flat = x.detach().reshape(1, -1)
    one = torch.ones(1, dtype=torch.float32, device=x.device)
    return ops.fakequant(flat, [one]).view(x.shape)


class STESign(Function):
    """sign(x) forward; gradient passed where -1 <= x <= 1 (Bengio et al. 2013)."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        ctx.save_for_backward(x)
        return binary_sign(x)

    @staticmethod
    def backward(ctx: Any, grad_output: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        x, = ctx.saved_tensors
        return ops.ste_backward(x, grad_output).view(x.shape)


binarize = STESign.apply
