"""Quantizer functions: (scales..., fake-quant tensor).

Mirror of quant/binary/quantization.py (clamps :17-24, QuantizerFP :27-32, quantizer_ls_1 :35-56,
quantizer_ls_2 :59-92, quantizer_ls_ternary :95-115, quantizer_gf :118-148).  The scale solves are
CUDA kernels; when no gradient is required the dense tensor is produced by one fused kernel
(lsq_fakequant), otherwise it is assembled from STESign terms exactly as the reference does so that
autograd sees the same graph (scales carry no gradient: the reference solves on x.clone().detach()).
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops
from .ste import binarize


def clamp_identity(x: torch.Tensor) -> torch.Tensor:
    return x


def clamp_symmetric(x: torch.Tensor, alpha: float) -> torch.Tensor:
    return x.clamp(-alpha, alpha)


class QuantizerFP(nn.Module):
    """Full precision: the identity."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        return x


def _rows(x: torch.Tensor) -> torch.Tensor:
    return x.detach().reshape(x.shape[0], -1)


def _needs_grad(x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and x.requires_grad


def _col(v: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return v.view(x.shape[0], *([1] * (x.dim() - 1)))


def quantizer_ls_1(x: torch.Tensor, v1: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """1-bit least squares (XNOR-Net): v1 = mean|x| per row, x_q = v1 * sign(x)."""
    ops.require_cuda(x, 'x')
    if v1 is None:
        v1 = ops.row_absmean(_rows(x))
    if _needs_grad(x):
        return v1, _col(v1, x) * binarize(x)
    return v1, ops.fakequant(_rows(x), [v1]).view(x.shape)


def quantizer_ls_2(x: torch.Tensor, v1: Optional[torch.Tensor] = None, v2: Optional[torch.Tensor] = None,
                   skip: int = 3) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """2-bit least squares: x_q = v1 b1 + v2 sign(x - v1 b1)."""
    ops.require_cuda(x, 'x')
    rows = _rows(x)
    v1 = ops.solve_v1(rows, False, skip) if v1 is None else v1.reshape(-1)
    v2 = ops.row_absmean(rows, [v1]) if v2 is None else v2.reshape(-1)
    if _needs_grad(x):
        s1 = _col(v1, x)
        b1 = binarize(x)
        return v1, v2, s1 * b1 + _col(v2, x) * binarize(x - s1 * b1)
    return v1, v2, ops.fakequant(rows, [v1, v2]).view(x.shape)


def quantizer_ls_ternary(x: torch.Tensor, v1: Optional[torch.Tensor] = None,
                         skip: int = 3) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ternary least squares: x_q = v1 (b1 + sign(x - v1 b1)) in {-2 v1, 0, 2 v1}."""
    ops.require_cuda(x, 'x')
    rows = _rows(x)
    v1 = ops.solve_v1(rows, True, skip) if v1 is None else v1.reshape(-1)
    if _needs_grad(x):
        s1 = _col(v1, x)
        b1 = binarize(x)
        return v1, s1 * (b1 + binarize(x - s1 * b1))
    return v1, ops.fakequant(rows, [v1], ternary=True).view(x.shape)


def quantizer_gf(x: torch.Tensor, k: int, vs: Optional[Sequence[torch.Tensor]] = None
                 ) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """Greedy foldable k-bit: v_i = mean|residual_{i-1}|, residual_i = residual_{i-1} - v_i sign(.)."""
    ops.require_cuda(x, 'x')
    if vs is not None and len(vs) != k:
        raise ValueError('If vs is passed in, all vs from v_1 to v_k must be passed in (could be None).')
    rows = _rows(x)
    scales: List[torch.Tensor] = []
    for i in range(k):
        scales.append(vs[i].reshape(-1) if vs is not None else ops.row_absmean(rows, scales))
    if _needs_grad(x):
        out = 0
        for v in scales:
            out = out + _col(v, x) * binarize(x - out)
        return scales, out  # type: ignore[return-value]
    return scales, ops.fakequant(rows, scales).view(x.shape)
