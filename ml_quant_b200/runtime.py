"""Inference runtime around the hot path: model construction / calibration for benchmarks and tests,
whole-forward CUDA-graph capture, a double-buffered host->device pipeline (the "public API" call a
user makes with host batches) and batch sharding over the GPUs of one node (one process per GPU,
weights replicated, one NCCL all_gather of logits -- SURVEY.md section 8e).
"""
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import configs
from .binary.binary_conv import QuantConv2d
from .nets import QLeNet5, QResNet


def strict_fp32(disable_cudnn: bool = True) -> None:
    """IEEE fp32 for the full-precision layers around the hot path (stem, shortcuts, classifier) in
    parity tests.  torch defaults cuDNN convolutions to TF32, and on B200 cuDNN's fp32 convolutions stay
    ~1e-3 off an fp32 reference even with TF32 disabled (measured, scripts/dev/gpu_diag_tf32.py); only the
    native ATen convolution (cudnn.enabled = False) reproduces the CPU result to ~6e-7."""
    if disable_cudnn:
        torch.backends.cudnn.enabled = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for mod, attr in ((getattr(torch.backends.cudnn, 'conv', None), 'fp32_precision'),
                      (getattr(torch.backends.cudnn, 'rnn', None), 'fp32_precision'),
                      (torch.backends.cudnn, 'fp32_precision'),
                      (getattr(torch.backends.cuda, 'matmul', None), 'fp32_precision')):
        if mod is not None and hasattr(mod, attr):
            try:
                setattr(mod, attr, 'ieee')
            except Exception:  # noqa: BLE001
                pass


def build_model(config: str, device: Optional[torch.device] = None, seed: int = 0) -> nn.Module:
    """Random-init network of the named shipped config (there are no checkpoints offline)."""
    torch.manual_seed(seed)
    arch = configs.arch(config)
    if 'lenet' in config:
        model: nn.Module = QLeNet5(loss_fn=F.nll_loss, **arch)
    else:
        model = QResNet(loss_fn=F.cross_entropy, **arch)
    return model.to(device) if device is not None else model


@torch.no_grad()
def calibrate(model: nn.Module, input_shape, batches: int = 2, batch: int = 32, seed: int = 100) -> nn.Module:
    """Populate what a checkpoint would carry: weight scales (w_approximate.v*) and BatchNorm running
    statistics, by ``batches`` train-mode forwards on seeded N(0,1) inputs.
    A freshly constructed model in eval() has v1 = 0 and outputs only biases."""
    dev = next(model.parameters()).device
    bns = [m for m in model.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)]
    saved = [m.momentum for m in bns]
    for m in bns:                     # cumulative average: running stats = mean of the batch statistics
        m.reset_running_stats()
        m.momentum = None
    model.train()
    for i in range(batches):
        g = torch.Generator(device='cpu').manual_seed(seed + i)
        model(torch.randn(batch, *input_shape, generator=g).to(dev))
    for m, mom in zip(bns, saved):
        m.momentum = mom
    return model.eval()


def _shortcut_fused(blk, x: torch.Tensor) -> torch.Tensor:
    """The full-precision downsampling shortcut (1x1 strided convolution + eval BatchNorm,
    quant/models/resnet.py:24-39) with the BatchNorm folded into the weights: one tcgen05 3xTF32 kernel
    (lsq_pwconv_fwd) when the channel counts allow it, else one batched fp32 GEMM on the strided slice.
    cuDNN ran it as layout transposes of the whole input + TF32 convolution + bias add + BatchNorm kernel
    (1.2 ms of a 10.5 ms step for the three shortcuts)."""
    sc = blk.shortcut
    if len(sc) == 0:
        return x
    conv, bn = (sc[0], sc[1]) if len(sc) == 2 else (None, None)
    ok = (isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and not bn.training and bn.track_running_stats
          and tuple(conv.kernel_size) == (1, 1) and tuple(conv.padding) == (0, 0) and conv.groups == 1
          and tuple(conv.dilation) == (1, 1) and conv.stride[0] == conv.stride[1] and x.is_cuda)
    if not ok:
        return sc(x)
    from .binary.binary_conv import bn_affine
    a, b = bn_affine(bn)
    key = (conv.weight.data_ptr(), conv.weight._version, a.data_ptr(),
           None if conv.bias is None else conv.bias._version)
    hit = getattr(blk, '_lsq_shortcut', None)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            w = (conv.weight.reshape(conv.out_channels, conv.in_channels) * a.view(-1, 1)).contiguous()
            bias = (b if conv.bias is None else b + a * conv.bias).reshape(1, -1, 1).contiguous()
        from . import ops
        image = ops.pwconv_pack(w) if ops.pwconv_supported(conv.in_channels, conv.out_channels) else None
        hit = (key, w, bias, image)
        blk._lsq_shortcut = hit
    _, w, bias, image = hit
    st = conv.stride[0]
    if image is not None and x.dtype == torch.float32:
        from . import ops
        return ops.pwconv_fwd(x, image, bias.reshape(-1), conv.out_channels, st)
    xs = x[:, :, ::st, ::st] if st > 1 else x
    n, c, h, wd = xs.shape
    xs = xs.reshape(n, c, h * wd)                      # copies the strided slice once
    return torch.baddbmm(bias, w.unsqueeze(0).expand(n, -1, -1), xs).view(n, conv.out_channels, h, wd)


def _xnor_block_fused(blk, x: torch.Tensor) -> torch.Tensor:
    """XnorBasicBlock.forward (quant/models/resnet.py:180-190) with bn1/bn2 folded into the quantizer
    kernels and nonlin / residual adds into the convolution epilogues."""
    sc = _shortcut_fused(blk, x)
    if blk.double_shortcut:
        first = blk.conv1.forward_fused(x, blk.bn1, blk.nonlin1, sc, True)
        return blk.conv2.forward_fused(first, blk.bn2, blk.nonlin2, first, True)
    first = blk.conv1.forward_fused(x, blk.bn1, blk.nonlin1)
    return blk.conv2.forward_fused(first, blk.bn2, blk.nonlin2, sc, False)


def pixel_lut(mean, std) -> torch.Tensor:
    """float[c][256]: the value torchvision's ``ToTensor`` (x / 255) followed by ``Normalize(mean, std)`` ((x - mean) / std)
    gives every uint8 pixel level of every channel, computed with the same fp32 operations on the host."""
    m = torch.as_tensor(mean, dtype=torch.float32).reshape(-1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).reshape(-1, 1)
    levels = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)
    return levels.reshape(1, 256).repeat(m.shape[0], 1).sub_(m).div_(s).contiguous()


def _lut_on(model: nn.Module, device: torch.device) -> torch.Tensor:
    cache = getattr(model, '_lsq_pixel_lut', None)
    if cache is None:
        raise RuntimeError('uint8 input: call runtime.set_pixel_input(model, mean, std) first')
    key = (device.type, device.index)
    if key not in cache:
        cache[key] = cache['host'].to(device)
    return cache[key]


def set_pixel_input(model: nn.Module, mean, std) -> nn.Module:
    """Let ``model`` take uint8 pixel batches [n, c, h, w] besides fp32 tensors: the host-side ``ToTensor`` + ``Normalize`` of the
    reference's loaders (a per-channel map of 256 levels) is applied on the device -- inside the stem kernel for the ImageNet
    stem, by ``lsq_u8_expand`` otherwise -- so a batch is uploaded as 1 byte per pixel instead of 4.  The fp32 input path is
    unchanged and the results are bit-identical to transforming on the host (the same fp32 value per level)."""
    import types
    from . import ops
    object.__setattr__(model, '_lsq_pixel_lut', {'host': pixel_lut(mean, std)})
    if getattr(model, '_lsq_pixel_wrapped', False):
        return model
    inner = model.forward

    def fwd(self, x):
        if x.dtype == torch.uint8:
            stem = getattr(self, '_lsq_stem', None)
            fused = stem is not None and not (self.training or torch.is_grad_enabled()) and stem.takes_u8(x)
            if not fused:
                x = ops.u8_expand(x, _lut_on(self, x.device))
        return inner(x)
    model.forward = types.MethodType(fwd, model)
    object.__setattr__(model, '_lsq_pixel_wrapped', True)
    return model


class _FusedStem(nn.Module):
    """conv1 -> bn1 -> relu -> maxpool of QResNet with the eval BatchNorm folded into the convolution
    weights and the max-pool taken before the ReLU (they commute), so the largest tensor of the network is
    written once and read once."""

    def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d, pool: nn.Module):
        super().__init__()
        self.conv, self.bn, self.pool = conv, bn, pool
        self._key = None
        self._image, self._image_key = None, None
        self.use_kernel = True
        self.owner = None                 # weak reference to the network (its pixel table, set_pixel_input)

    def _folded(self):
        from .binary.binary_conv import bn_affine
        a, b = bn_affine(self.bn)
        key = (self.conv.weight.data_ptr(), self.conv.weight._version, a.data_ptr())
        if self._key != key:
            with torch.no_grad():
                self._w = (self.conv.weight * a.view(-1, 1, 1, 1)).contiguous()
                self._b = b if self.conv.bias is None else (b + a * self.conv.bias)
            self._key = key
        return self._w, self._b

    def _kernel_ok(self, x: torch.Tensor) -> bool:
        return x.is_cuda and x.dtype == torch.float32 and self._layers_ok()

    def _layers_ok(self) -> bool:
        c, p = self.conv, self.pool
        return (c.in_channels == 3 and c.out_channels == 64
                and tuple(c.kernel_size) == (7, 7) and tuple(c.stride) == (2, 2) and tuple(c.padding) == (3, 3)
                and tuple(c.dilation) == (1, 1) and c.groups == 1 and isinstance(p, nn.MaxPool2d)
                and p.kernel_size == 3 and p.stride == 2 and p.padding == 1 and p.dilation == 1 and not p.ceil_mode)

    def takes_u8(self, x: torch.Tensor) -> bool:
        """uint8 pixels go straight into the stem kernel (no fp32 image in HBM) when its one-kernel route applies."""
        from . import ops
        return (self.use_kernel and x.is_cuda and x.dtype == torch.uint8 and x.dim() == 4 and x.shape[1] == 3
                and self._layers_ok()
                and bool(ops._C.lib().lsq_stem_is_fused(int(x.shape[0]), int(x.shape[2]), int(x.shape[3]))))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        if x.dtype == torch.uint8:
            from . import ops
            lut = _lut_on(self.owner() if self.owner is not None else self, x.device)
            if not (self.bn.training or torch.is_grad_enabled()) and self.takes_u8(x):
                w, b = self._folded()
                if self._image is None or self._image_key != self._key:
                    self._image = ops.stem_pack(w)
                    self._image_key = self._key
                return ops.stem_fwd_u8(x, lut, self._image, b.contiguous())
            x = ops.u8_expand(x, lut)
        if self.bn.training or torch.is_grad_enabled():
            return self.pool(F.relu(self.bn(self.conv(x))))
        w, b = self._folded()
        if self.use_kernel and self._kernel_ok(x):
            from . import ops
            if ops.stem_supported(x.shape[0], x.shape[2], x.shape[3]):
                if self._image is None or self._image_key != self._key:
                    self._image = ops.stem_pack(w)
                    self._image_key = self._key
                return ops.stem_fwd(x, self._image, b.contiguous())
        # bias and ReLU commute with the max-pool: apply them on the 4x smaller pooled tensor
        y = F.conv2d(x, w, None, self.conv.stride, self.conv.padding, self.conv.dilation, self.conv.groups)
        return F.relu_(self.pool(y).add_(b.view(1, -1, 1, 1)))


def _classifier_fused(head: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """``linear_classifier`` of QResNet (AdaptiveAvgPool2d((1, 1)), Flatten, Linear ...; quant/models/resnet.py) with the global
    average pool as one warp-per-plane kernel (lsq_plane_mean) when the head has that shape; anything else runs as it is."""
    mods = list(head.children()) if isinstance(head, nn.Sequential) else []
    if (len(mods) >= 2 and isinstance(mods[0], nn.AdaptiveAvgPool2d) and mods[0].output_size in (1, (1, 1))
            and isinstance(mods[1], nn.Flatten) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        from . import ops
        x = ops.plane_mean(x)
        for m in mods[2:]:
            x = m(x)
        return x
    return head(x)


def optimize_for_inference(model: nn.Module) -> nn.Module:
    """Rewrite the callers of the hot path for eval-mode inference: every XnorBasicBlock
    runs its two QuantConv2d through ``forward_fused`` and the stem uses a BatchNorm-folded convolution.
    Parameters, buffers and state_dict keys are untouched; training mode falls back to the original graph."""
    import types
    from .nets import QResNet, XnorBasicBlock
    for m in model.modules():
        if isinstance(m, XnorBasicBlock) and not hasattr(m, '_lsq_orig_forward'):
            m._lsq_orig_forward = m.forward
            # the fused block (folded shortcut / BatchNorm kernels) builds no autograd graph: any call that may
            # need gradients -- training, or eval with autograd on (frozen-BN fine-tuning, saliency, KD
            # teachers) -- runs the original forward
            m.forward = types.MethodType(
                lambda self, x: (_xnor_block_fused(self, x) if not (self.training or torch.is_grad_enabled())
                                 else self._lsq_orig_forward(x)), m)
    if isinstance(model, QResNet) and not isinstance(model.blocks[0], _FusedStem):
        import weakref
        stem = _FusedStem(model.conv1, model.bn1, model.maxpool)
        stem.owner = weakref.ref(model)
        object.__setattr__(model, '_lsq_stem', stem)      # not registered: state_dict stays the reference's
        orig_forward = model.forward

        def fwd(self, x):
            if self.training or torch.is_grad_enabled():
                if x.dtype == torch.uint8:
                    from . import ops
                    x = ops.u8_expand(x, _lut_on(self, x.device))
                return orig_forward(x)
            x = self._lsq_stem(x)
            for blk in list(self.blocks)[1:]:
                x = blk(x)
            return _classifier_fused(self.linear_classifier, x)
        model.forward = types.MethodType(fwd, model)
    return model


class GraphedForward:
    """model(x) for a fixed input shape as one CUDA graph: no per-layer launch gaps, no host work."""

    def __init__(self, model: nn.Module, example: torch.Tensor, warmup: int = 2, clone: bool = True):
        self.model = model
        # clone=False: the graph reads ``example`` itself (a caller-owned device buffer that is refilled
        # between replays, e.g. the upload buffers of HostPipeline)
        self.static_in = example.clone() if clone else example
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.model(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self.model(self.static_in)
        # the captured kernels address the modules' cached buffers (bit planes, packed weights, scratch): keep
        # them alive for the life of the graph even if a later eager call with another shape replaces a cache entry
        from . import ops
        self._keepalive = [(dict(m._planes_cache), dict(m._wpack_cache)) for m in quant_layers(model)]
        self._keepalive.append((dict(ops._stem_ws), dict(getattr(ops._tls, 'ws', None) or {})))

    def __call__(self, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None and x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


class HostPipeline:
    """Classify host batches: pinned host memory -> GPU (copy stream) -> forward -> logits to host.

    The upload of batch i+1 overlaps the forward of batch i (two device input buffers)."""

    def __init__(self, model: nn.Module, batch_shape, device: torch.device, use_graph: bool = True,
                 dtype: torch.dtype = torch.float32):
        # dtype=torch.uint8: pixel batches for a model prepared with set_pixel_input (1 byte per pixel over PCIe)
        self.device = device
        self.bufs = [torch.empty(batch_shape, device=device, dtype=dtype) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.freed = [torch.cuda.Event() for _ in range(2)]
        self.model = model
        for b in self.bufs:
            b.zero_()                     # defined contents for the capture warm-up runs
        self.fwd = [GraphedForward(model, b, clone=False) if use_graph else None for b in self.bufs]
        n_out = self._run(0).shape
        self.host_out = torch.empty(n_out, pin_memory=True)
        for e in self.freed:
            e.record()

    def _run(self, i: int) -> torch.Tensor:
        if self.fwd[i] is not None:
            return self.fwd[i]()
        with torch.no_grad():
            return self.model(self.bufs[i])

    def upload(self, i: int, host_batch: torch.Tensor) -> None:
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.freed[i])
            self.bufs[i].copy_(host_batch, non_blocking=True)
            self.ready[i].record(self.copy_stream)

    def infer(self, i: int) -> torch.Tensor:
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[i])
        out = self._run(i)
        self.freed[i].record(cur)
        self.host_out.copy_(out, non_blocking=True)
        return self.host_out

    def run(self, host_batches) -> torch.Tensor:
        """host_batches: sequence of pinned tensors; returns the logits of the last one (on host)."""
        n = len(host_batches)
        self.upload(0, host_batches[0])
        for k in range(n):
            if k + 1 < n:
                self.upload((k + 1) & 1, host_batches[k + 1])
            self.infer(k & 1)
        torch.cuda.current_stream().synchronize()
        return self.host_out


@torch.no_grad()
def refresh_weight_scales(model: nn.Module, skip: int = 3) -> nn.Module:
    """Solve and store the weight-quantizer scales of every ``QuantConv2d`` / ``QuantLinear`` in ``model`` -- what one
    train-mode forward does layer by layer (quant/binary/weight_quantization.py:27-34, :51-59, :75-82) -- with ONE
    multi-tensor launch per scheme (``lsq_row_absmean_multi`` for ls-1, ``lsq_solve_v1_multi`` for ls-2 / ls-T; the
    values are bit-identical to the per-layer solves).  gf-k weights keep their per-layer greedy solve."""
    from . import ops
    from .binary import quantization
    groups: Dict[str, list] = {'ls-1': [], 'ls-2': [], 'ls-T': []}
    for m in quant_layers(model):
        if m.w_quant in groups:
            ops.require_cuda(m.weight, 'weight')
            groups[m.w_quant].append(m)
        elif m.w_quant.startswith('gf'):
            vs, _ = quantization.quantizer_gf(m.weight.detach(), k=m.w_approximate.k)
            for buf, v in zip(m.w_approximate.scales(), vs):
                buf.copy_(v)
    rows = {k: [m.weight.detach().reshape(m.out_channels, -1) for m in ms] for k, ms in groups.items()}
    if groups['ls-1']:
        for m, v1 in zip(groups['ls-1'], ops.row_absmean_multi(rows['ls-1'])):
            m.w_approximate.v1.copy_(v1)
    for scheme, tern in (('ls-2', False), ('ls-T', True)):
        if groups[scheme]:
            for m, w, v1 in zip(groups[scheme], rows[scheme], ops.solve_v1_multi(rows[scheme], tern, skip)):
                m.w_approximate.v1.copy_(v1)
                if not tern:
                    m.w_approximate.v2.copy_(ops.row_absmean(w, [v1]))
    return model


PACKED_FORMAT = 'lsq_b200.packed.v1'


def export_packed(model: nn.Module) -> Dict[str, torch.Tensor]:
    """Deployable checkpoint: the model's ``state_dict`` with the fp32 weight of every ls-1
    ``QuantConv2d`` (in eval mode the layer only uses sign(W) and the stored ``w_approximate.v1``,
    quant/binary/weight_quantization.py:32-33) replaced by its sign image -- ``<layer>.weight_bits`` int32
    [cout, kh*kw, ceil(cin/32)] plus ``<layer>.weight_shape`` -- 1 bit instead of 32 per quantized weight.
    Everything else (scales, biases, BatchNorm, fp layers, moving averages) is copied unchanged, on the CPU."""
    from . import ops
    from .binary.binary_conv import QuantConv2d
    out: Dict[str, torch.Tensor] = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    for name, m in model.named_modules():
        if isinstance(m, QuantConv2d) and m.w_quant == 'ls-1':
            key = f'{name}.weight' if name else 'weight'
            ops.require_cuda(m.weight, f'{key} (export packs on the GPU)')
            out[f'{key}_bits'] = ops.weight_bits(m.weight).cpu()
            out[f'{key}_shape'] = torch.tensor(tuple(m.weight.shape), dtype=torch.int64)
            del out[key]
    out['_format'] = torch.tensor(list(PACKED_FORMAT.encode()), dtype=torch.uint8)
    return out


def load_packed(model: nn.Module, packed: Dict[str, torch.Tensor], strict: bool = True):
    """Load an ``export_packed`` checkpoint into ``model`` (already on its CUDA device): the sign images are
    expanded to W = +-v1[co], whose sign(W) and per-channel mean|W| are exactly the exported ones, so eval-mode
    outputs equal those of the original fp32 checkpoint bit for bit and a train-mode re-solve returns the same
    v1.  Returns ``load_state_dict``'s result."""
    from . import ops
    fmt = packed.get('_format')
    if fmt is None or bytes(fmt.tolist()).decode() != PACKED_FORMAT:
        raise ValueError(f'not a {PACKED_FORMAT} checkpoint')
    dev = next(model.parameters()).device
    state = {k: v for k, v in packed.items() if k != '_format' and not k.endswith(('.weight_bits', '.weight_shape'))
             and k not in ('weight_bits', 'weight_shape')}
    for k, bits in packed.items():
        if k.endswith('weight_bits'):
            base = k[:-len('_bits')]
            prefix = base[:-len('weight')]
            v1 = packed[f'{prefix}w_approximate.v1'].to(dev, torch.float32)
            state[base] = ops.unpack_weights(bits.to(dev), v1, packed[f'{base}_shape'].tolist())
    return model.load_state_dict(state, strict=strict)


def bind_host_to_gpu(device_index: int) -> Optional[str]:
    """Pin this process (one process per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned
    upload buffers are allocated, so that they are first-touched on that node: with eight ranks uploading 308 MB per
    step each, buffers that all sit on one socket make seven of the eight copies cross the inter-socket link
    (round 1: 22 GB/s per GPU at 8 ranks against 55 GB/s for one).  Returns a description, or None when the
    topology cannot be read (no sysfs, single node): nothing is changed then."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f'gpu {device_index} ({bdf}) -> numa node {node}, {len(cpus)} cpus'
    except Exception:  # noqa: BLE001
        return None


def gather_logits(local: torch.Tensor, world: int) -> torch.Tensor:
    """The path's only collective: all ranks' logits (one NCCL all_gather over NVLink)."""
    if world == 1:
        return local
    import torch.distributed as dist
    local = local.contiguous()
    if dist.get_backend() == 'nccl':
        out = torch.empty(world * local.shape[0], *local.shape[1:], dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local)
        return out
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local)
    return torch.cat(parts)


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous batch split (samples are independent units: per-sample scales, eval BatchNorm)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def quant_layers(model: nn.Module):
    return [m for m in model.modules() if isinstance(m, QuantConv2d)]
