"""Weight quantizer modules: one scale (set) per output channel, cached in buffers.

Mirror of quant/binary/weight_quantization.py: training solves and stores the scales (:29-31, :53-56,
:77-79, :102-105), evaluation re-uses the stored buffers (:32-33, :57-58, :80-81, :106-108).  Buffer
names and shapes (v1, v2 / v1..vk, [size]) are the reference's, so state_dicts are interchangeable.
"""
import torch
import torch.nn as nn

from . import quantization


class _WeightQuantizer(nn.Module):
    names = ('v1',)

    def __init__(self, size: int) -> None:
        super().__init__()
        for n in self.names:
            self.register_buffer(n, torch.zeros(size))

    def scales(self):
        return [getattr(self, n) for n in self.names]

    def _store(self, values) -> None:
        for n, v in zip(self.names, values):
            getattr(self, n).copy_(v)


class WeightQuantizerLS1(_WeightQuantizer):
    def forward(self, w: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        if self.training:
            v1, w_q = quantization.quantizer_ls_1(w)
            self._store([v1])
            return w_q
        return quantization.quantizer_ls_1(w, self.v1)[1]


class WeightQuantizerLS2(_WeightQuantizer):
    names = ('v1', 'v2')

    def forward(self, w: torch.Tensor, skip: int = 3) -> torch.Tensor:  # type: ignore[override]
        if self.training:
            v1, v2, w_q = quantization.quantizer_ls_2(w, skip=skip)
            self._store([v1, v2])
            return w_q
        return quantization.quantizer_ls_2(w, self.v1, self.v2, skip=skip)[2]


class WeightQuantizerLST(_WeightQuantizer):
    def forward(self, w: torch.Tensor, skip: int = 3) -> torch.Tensor:  # type: ignore[override]
        if self.training:
            v1, w_q = quantization.quantizer_ls_ternary(w, skip=skip)
            self._store([v1])
            return w_q
        return quantization.quantizer_ls_ternary(w, self.v1, skip=skip)[1]


class WeightQuantizerGF(nn.Module):
    def __init__(self, size: int, k: int) -> None:
        super().__init__()
        self.k = k
        for i in range(1, k + 1):
            self.register_buffer(f'v{i}', torch.zeros(size))

    def scales(self):
        return [getattr(self, f'v{i + 1}') for i in range(self.k)]

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        if self.training:
            vs, x_q = quantization.quantizer_gf(x, k=self.k)
            for buf, v in zip(self.scales(), vs):
                buf.copy_(v)
            return x_q
        return quantization.quantizer_gf(x, k=self.k, vs=self.scales())[1]
