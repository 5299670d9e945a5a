"""QuantConv2d: convolution on scaled-binary weights and activations.

Mirror of quant/binary/binary_conv.py (constructor and factories :55-159, forward :161-173): same
constructor signature, attributes (``x_approximate``, ``w_approximate``, ``clamping_fn``,
``quantized_parameters``), scheme validation and state_dict keys.

Forward has two routes, both on the GPU:
  * packed  -- inference with ls-1 weights and a 1..4-plane activation code (ls-1, ls-2, ls-T, gf-k):
               clamp + scale solve + bit-plane encode (csrc/lsq_solve.cu, lsq_quant.cu), then a binary
               convolution with integer accumulation and the scales, bias applied in the epilogue
               (csrc/lsq_bconv_tc.cu on tcgen05 tensor cores, csrc/lsq_bconv.cu otherwise).  The dense
               fake-quant tensors of the reference are never materialised.  sign(W) is packed once and
               cached until the weight changes.
  * generic -- everything else (gradients required, other weight schemes, grouped / dilated convs):
               the reference's composition x_approximate(clamp(x)), w_approximate(W), F.conv2d with the
               quantizers running as CUDA kernels.
"""
from collections import defaultdict
from functools import partial
import re
import warnings
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import activation_quantization, quantization, weight_quantization

_PACKED_MAX_PLANES = 4


def bn_affine(bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm as y = x * a + b per channel (a = gamma / sqrt(var + eps), b = beta - mean * a),
    cached on the module until one of its tensors changes."""
    srcs = [t for t in (bn.running_mean, bn.running_var, bn.weight, bn.bias) if t is not None]
    key = tuple((t.data_ptr(), t._version) for t in srcs)
    hit = getattr(bn, '_lsq_affine', None)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            a = torch.rsqrt(bn.running_var + bn.eps)
            if bn.weight is not None:
                a = a * bn.weight
            b = -bn.running_mean * a
            if bn.bias is not None:
                b = b + bn.bias
        hit = (key, a.float().contiguous(), b.float().contiguous())
        bn._lsq_affine = hit
    return hit[1], hit[2]


class QuantConv2d(nn.Conv2d):
    """Conv2d(w_quant(w), x_quant(clamp(x)))."""

    def __init__(self, x_quant: str, w_quant: str, in_channels: int, out_channels: int,
                 kernel_size: Union[int, Tuple[int, int]], clamp: Optional[Dict] = None,
                 moving_average_mode: str = 'off', moving_average_momentum: float = 0.99, **kwargs: Any):
        super().__init__(in_channels, out_channels, kernel_size, **kwargs)
        clamp = {'kind': 'identity'} if clamp is None else clamp
        self.x_approximate = self._get_x_quantizer(x_quant, moving_average_mode, moving_average_momentum)
        self.w_approximate = self._get_w_quantizer(w_quant, out_channels)
        self.clamping_fn = self._get_clamper(**clamp)
        self.quantized_parameters: Dict[str, List[torch.Tensor]] = defaultdict(list)
        if self.bias is not None:
            self.quantized_parameters['fp'].append(self.bias)
        self.quantized_parameters[w_quant].append(self.weight)
        # host-side description used by the packed route
        self.x_quant, self.w_quant = x_quant, w_quant
        self.clamp_alpha: Optional[float] = float(clamp.get('alpha', 2)) if clamp['kind'] == 'symmetric' else None
        self.packed_impl = 0            # 0 auto, 1 CUDA-core kernel, 2 tensor-core kernel (tests / profiling)
        self.allow_packed = True
        self._wpack_cache: Dict[int, Tuple[Tuple[int, int], torch.Tensor]] = {}
        self._wplanes_cache: Dict[int, Tuple[tuple, list]] = {}
        # bit-plane scratch, one per (device, stream): two streams / threads running this module concurrently
        # (nn.DataParallel replicas share this dict; a graph on one stream next to eager calls on another)
        # must not overwrite each other's planes between the encoder and the convolution
        self._planes_cache: Dict[Tuple[int, int], torch.Tensor] = {}

    # ---- factories (same static methods as the reference) -----------------------------------
    @staticmethod
    def _validate_scheme(scheme: str) -> None:
        if scheme not in {'fp', 'ls-1', 'ls-T', 'ls-2'} and not re.fullmatch(r'gf-\d+', scheme):
            raise ValueError(f'Scheme {scheme} is invalid. Please see docs for valid schemes.')

    @staticmethod
    def _get_x_quantizer(scheme: str, moving_average_mode: str = 'off',
                         moving_average_momentum: float = 0.99) -> nn.Module:
        QuantConv2d._validate_scheme(scheme)
        aq = activation_quantization
        if scheme == 'fp':
            return quantization.QuantizerFP()
        if scheme.startswith('gf'):
            return aq.ActivationQuantizerGF(int(scheme.split('-')[1]), moving_average_mode, moving_average_momentum)
        cls = {'ls-1': aq.ActivationQuantizerLS1, 'ls-2': aq.ActivationQuantizerLS2,
               'ls-T': aq.ActivationQuantizerLST}[scheme]
        return cls(moving_average_mode, moving_average_momentum)

    @staticmethod
    def _get_w_quantizer(scheme: str, size: int) -> nn.Module:
        QuantConv2d._validate_scheme(scheme)
        wq = weight_quantization
        if scheme == 'fp':
            return quantization.QuantizerFP()
        if scheme.startswith('gf'):
            return wq.WeightQuantizerGF(size, int(scheme.split('-')[1]))
        return {'ls-1': wq.WeightQuantizerLS1, 'ls-2': wq.WeightQuantizerLS2,
                'ls-T': wq.WeightQuantizerLST}[scheme](size)

    @staticmethod
    def _get_clamper(kind: str, alpha: float = 2) -> Callable[[torch.Tensor], torch.Tensor]:
        table: Dict[str, Callable[[torch.Tensor], torch.Tensor]] = {
            'identity': quantization.clamp_identity,
            'symmetric': partial(quantization.clamp_symmetric, alpha=alpha),
        }
        if kind not in table:
            raise ValueError(f'{kind} is not a valid clamping function.')
        return table[kind]

    # ---- packed route ---------------------------------------------------------------------------
    def _num_planes(self) -> int:
        s = self.x_quant
        return {'ls-1': 1, 'ls-2': 2, 'ls-T': 2}.get(s) or (int(s.split('-')[1]) if s.startswith('gf') else 0)

    def _packed_geometry(self, x: torch.Tensor):
        """Geometry when the packed route applies to this call, else None."""
        if not (self.allow_packed and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            return None
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad
                                        or (self.bias is not None and self.bias.requires_grad)):
            # autograd needs the reference's graph (STESign terms): generic route.  In eval mode this is
            # almost always a forgotten torch.no_grad() -- say so once, the generic route is several times slower.
            if not self.training and not getattr(QuantConv2d, '_warned_grad_eval', False):
                QuantConv2d._warned_grad_eval = True
                warnings.warn('QuantConv2d: eval-mode forward with autograd enabled takes the generic (fake-quant + '
                              'F.conv2d) route; wrap inference in torch.no_grad() for the packed tensor-core path.',
                              RuntimeWarning, stacklevel=3)
            return None
        if self.w_quant == 'fp' or not 1 <= self._num_planes() <= _PACKED_MAX_PLANES:
            return None
        if self.w_quant != 'ls-1' and self.w_approximate.training:
            return None          # multi-plane weight scales are being solved: the reference's composition does that
        if self.groups != 1 or tuple(self.dilation) != (1, 1) or self.padding_mode != 'zeros':
            return None
        if isinstance(self.padding, str) or self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1]:
            return None
        xa = self.x_approximate
        if xa.training and xa.moving_average_mode != activation_quantization.MovingAverageMode.off:
            return None          # the moving average is being tracked: keep the reference's control flow
        n, c, h, w = x.shape
        return ops.act_geometry(n, c, h, w, self.kernel_size[0], self.kernel_size[1], self.stride[0], self.padding[0])

    def packed_weights(self) -> torch.Tensor:
        """sign(W) operand images, repacked only when the weight tensor changes."""
        w = self.weight
        dev = w.device.index or 0
        key = (w.data_ptr(), w._version)
        hit = self._wpack_cache.get(dev)
        if hit is None or hit[0] != key:
            hit = (key, ops.pack_weights(w))
            self._wpack_cache[dev] = hit
        return hit[1]

    def packed_weight_planes(self) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Multi-plane weight schemes (ls-2, ls-T, gf-k; quant/binary/weight_quantization.py:37-109 in eval mode):
        W_q = sum_i v_i b_i with b_i = sign(W - sum_{l<i} v_l b_l) per output channel (quantization.py:89-92,
        :113-115, :139-146).  Returns [(packed sign image of b_i, v_i)], cached until the weight or a scale buffer
        changes; every plane runs through the same binary convolution kernels as an ls-1 weight."""
        w, wa = self.weight, self.w_approximate
        scales = list(wa.scales())
        if self.w_quant == 'ls-T':
            scales = [scales[0], scales[0]]
        dev = w.device.index or 0
        key = (w.data_ptr(), w._version) + tuple((v.data_ptr(), v._version) for v in scales)
        hit = self._wplanes_cache.get(dev)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                acc = torch.zeros_like(w)
                planes = []
                for v in scales:
                    b = torch.where(w - acc >= 0, 1.0, -1.0).to(w.dtype)       # sign with sign(0) = +1 (ste.py:16-18)
                    planes.append((ops.pack_weights(b), v.detach()))
                    acc = acc + v.view(-1, 1, 1, 1) * b
            hit = (key, planes)
            self._wplanes_cache[dev] = hit
        return hit[1]

    def _weight_scale(self) -> torch.Tensor:
        wa = self.w_approximate
        if wa.training:
            # train-mode call under no_grad (e.g. calibration): solve and cache like the reference
            wa.v1.copy_(ops.row_absmean(self.weight.detach().reshape(self.out_channels, -1)))
        return wa.v1

    def quantize_input(self, x: torch.Tensor, g, prologue=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Packed activation code of clamp(x * ch_scale + ch_shift): (bit planes, scale table [planes, batch]).
        ``prologue`` = (ch_scale, ch_shift, H*W) folds a preceding eval-mode BatchNorm into the kernels."""
        xa, alpha, npl = self.x_approximate, self.clamp_alpha, self._num_planes()
        pro = prologue
        n = x.shape[0]
        dev = (x.device.index or 0, torch.cuda.current_stream(x.device).cuda_stream)
        buf = self._planes_cache.get(dev)
        rows = x.reshape(n, -1)
        tern = self.x_quant == 'ls-T'
        if xa.uses_stored_scales():
            scales = [s.contiguous() for s in xa.stored_scales(n)]
            known = scales[:1] if tern else scales
            planes, _ = ops.encode_act(x, g, known[:npl], npl, alpha, False, buf, pro)
            table = scales + scales if tern else scales
        elif self.x_quant == 'ls-1':
            planes, v1 = ops.encode_act(x, g, [], 1, alpha, True, buf, pro)
            table = [v1]
        elif self.x_quant in ('ls-2', 'ls-T'):
            # one fused launch (csrc/lsq_qact.cu): clamp + BatchNorm prologue + v1 solve + v2 + both sign planes with a
            # single read of x from HBM, written straight into the [2, n] table the convolution reads
            planes, tab = ops.quantize_act(x, g, tern, alpha, 3, buf, pro)
            self._planes_cache[dev] = planes
            return planes, tab
        else:
            scales: List[torch.Tensor] = []
            for _ in range(npl - 1):
                scales.append(ops.row_absmean(rows, scales, alpha, pro))
            planes, last = ops.encode_act(x, g, scales, npl, alpha, True, buf, pro)
            table = scales + [last]
        self._planes_cache[dev] = planes
        return planes, torch.stack(table)

    def _forward_packed(self, x: torch.Tensor, g, prologue=None, act: int = 0, prelu=None, residual=None,
                        residual_after_act: bool = True) -> torch.Tensor:
        x = x.contiguous()
        planes, table = self.quantize_input(x, g, prologue)
        if self.w_quant == 'ls-1':
            return ops.bconv2d(planes, g, self._num_planes(), table, self.packed_weights(), self._weight_scale(),
                               self.bias, self.out_channels, self.packed_impl, None, residual, act, prelu,
                               residual_after_act)
        # multi-plane weights: y = sum_i v_i[c] * sum_j s_j[n] * conv(b_j, wb_i) + bias -- one binary convolution per
        # weight plane, each adding to the previous through the residual input of the epilogue (exact integer
        # accumulators per plane pair); the block's own activation / residual, if any, on the way out of the last one
        wplanes = self.packed_weight_planes()
        y = None
        for i, (wpack, v) in enumerate(wplanes):
            lastp = i == len(wplanes) - 1
            res_in = y
            if lastp and residual is not None and not residual_after_act:
                res_in = residual if y is None else y + residual          # act(sum + residual)
            y = ops.bconv2d(planes, g, self._num_planes(), table, wpack, v, self.bias if i == 0 else None,
                            self.out_channels, self.packed_impl, None, res_in, act if lastp else 0,
                            prelu if lastp else None, False)
        if residual is not None and residual_after_act:                   # act(sum) + residual
            y = y + residual
        return y

    def forward_fused(self, x: torch.Tensor, bn: Optional[nn.BatchNorm2d] = None, nonlin: Optional[nn.Module] = None,
                      residual: Optional[torch.Tensor] = None, residual_after_act: bool = True) -> torch.Tensor:
        """nonlin(conv(bn(x))) (+ residual, after or before the non-linearity) with the eval-mode BatchNorm
        folded into the quantizer kernels and bias / non-linearity / residual into the convolution epilogue.
        Falls back to the unfused composition whenever the packed route does not apply."""
        g = self._packed_geometry(x)
        kind = {nn.ReLU: 1, nn.PReLU: 2, nn.Identity: 0, type(None): 0}.get(type(nonlin))
        ok = g is not None and kind is not None and (bn is None or (not bn.training and bn.track_running_stats))
        if ok and torch.is_grad_enabled():
            # the fused kernels build no autograd graph: anything around the convolution that wants a gradient
            # (BatchNorm affine, PReLU slope, the residual branch) sends the call to the unfused composition
            extra = [residual] + [p for m in (bn, nonlin) if m is not None for p in m.parameters()]
            ok = not any(t is not None and t.requires_grad for t in extra)
        if not ok:
            y = self.forward(x if bn is None else bn(x))
            if residual is not None and not residual_after_act:
                y = y + residual
            y = y if nonlin is None else nonlin(y)
            return y + residual if (residual is not None and residual_after_act) else y
        pro = None
        if bn is not None:
            a, b = bn_affine(bn)
            pro = (a, b, x.shape[2] * x.shape[3])
        prelu = nonlin.weight if kind == 2 else None
        return self._forward_packed(x, g, pro, kind, prelu, residual, residual_after_act)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        g = self._packed_geometry(x)
        if g is not None:
            return self._forward_packed(x, g)
        x_q = self.x_approximate(self.clamping_fn(x))
        w_q = self.w_approximate(self.weight)
        return F.conv2d(x_q, w_q, self.bias, self.stride, self.padding, self.dilation, self.groups)


class QuantLinear(nn.Module):
    """Linear layer on scaled-binary weights and activations: ``y = w_quant(W) @ x_quant(clamp(x)) + b``.

    The reference has no quantized linear layer (its classifiers are ``nn.Linear``, quant/models/resnet.py:339,
    lenet.py:75-76); BASELINE.json's north star names one, so it is provided as the 1x1 / no-spatial case of
    ``QuantConv2d``: a row of ``x`` is one sample [in_features, 1, 1], hence per-row activation
    scales and per-output-feature weight scales, the same schemes, clamp and moving-average options, and the same
    kernels (packed route: solve + encode + tcgen05 binary GEMM when in/out features are multiples of 64).
    Inputs of shape [..., in_features] are flattened over the leading dimensions.
    """

    def __init__(self, x_quant: str, w_quant: str, in_features: int, out_features: int, clamp: Optional[Dict] = None,
                 moving_average_mode: str = 'off', moving_average_momentum: float = 0.99, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.conv = QuantConv2d(x_quant, w_quant, in_features, out_features, 1, clamp, moving_average_mode,
                                moving_average_momentum, bias=bias)

    @property
    def weight(self) -> torch.Tensor:
        return self.conv.weight.view(self.out_features, self.in_features)

    @property
    def bias(self) -> Optional[torch.Tensor]:
        return self.conv.bias

    @property
    def x_approximate(self) -> nn.Module:
        return self.conv.x_approximate

    @property
    def w_approximate(self) -> nn.Module:
        return self.conv.w_approximate

    @property
    def quantized_parameters(self) -> Dict[str, List[torch.Tensor]]:
        return self.conv.quantized_parameters

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        if x.shape[-1] != self.in_features:
            raise ValueError(f'QuantLinear: expected last dimension {self.in_features}, got {tuple(x.shape)}')
        lead = x.shape[:-1]
        y = self.conv(x.reshape(-1, self.in_features, 1, 1))
        return y.reshape(*lead, self.out_features)

    def extra_repr(self) -> str:
        return f'in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}'
