"""Host-side launchers: torch tensors in, C-ABI calls on the current CUDA stream, torch tensors out.

Torch is plumbing here (device memory, streams); every value is computed by liblsq_b200.so.
All entry points require contiguous fp32 CUDA tensors and raise otherwise -- there is no CPU path.
"""
import ctypes as C
import threading
from typing import List, Optional, Sequence, Tuple

import torch

from . import _C

_tls = threading.local()

# ---- instrumentation (bench.py): launches of OUR kernels and optional per-launch CUDA-event timing ----
LAUNCHES = {}          # kernel name -> number of launches since reset_counters()
PROFILE = None         # None, or a list receiving (name, start_event, end_event, alg_bytes, alg_ops)


def reset_counters() -> None:
    LAUNCHES.clear()


class _launch:
    """Context manager around one C-ABI kernel launch: counts it, optionally brackets it with events."""

    def __init__(self, name: str, alg_bytes: float = 0.0, alg_ops: float = 0.0):
        self.name, self.b, self.o = name, alg_bytes, alg_ops

    def __enter__(self):
        LAUNCHES[self.name] = LAUNCHES.get(self.name, 0) + 1
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, self.e0, self.e1, self.b, self.o))
        return False


def require_cuda(x: torch.Tensor, what: str = 'tensor', dtype: torch.dtype = torch.float32) -> None:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise _C.LsqError(
            f'ml_quant_b200: {what} must be a CUDA tensor -- the quantizer runs only as sm_100a CUDA kernels '
            '(no CPU fallback; the CPU restatement used for testing lives in oracle/).')
    if x.dtype != dtype:
        raise _C.LsqError(f'ml_quant_b200: {what} must be {str(dtype).replace("torch.", "")}, got {x.dtype}')


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch, cached per (thread, device, stream); kernels leave it zeroed."""
    cache = getattr(_tls, 'ws', None)
    if cache is None:
        cache = _tls.ws = {}
    key = (device.index, _stream())
    buf = cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        cache[key] = buf
    return buf


def scale_table(scales: Sequence[torch.Tensor], rows: int, device) -> Optional[torch.Tensor]:
    if len(scales) == 0:
        return None
    if isinstance(scales, torch.Tensor) and scales.dim() == 2:     # already a [nscales, rows] table: no copy kernel
        if scales.shape[1] != rows or scales.dtype != torch.float32 or not scales.is_contiguous():
            raise ValueError(f'scale table must be contiguous float32 [k, {rows}]')
        return scales.detach()
    tab = torch.stack([s.detach().reshape(-1).to(device=device, dtype=torch.float32) for s in scales]).contiguous()
    if tab.shape[1] != rows:
        raise ValueError(f'scale vectors must have {rows} entries, got {tab.shape[1]}')
    return tab


def _alpha(alpha: Optional[float]) -> float:
    return float(alpha) if alpha is not None and alpha > 0 else 0.0


def _prologue(pro, keep):
    """pro = (ch_scale, ch_shift, inner) or None -> ctypes pointer (tensors kept alive in ``keep``)."""
    if pro is None:
        return None
    a, b, inner = pro
    a, b = a.contiguous(), b.contiguous()
    keep.extend([a, b])
    return C.byref(_C.Prologue(a.data_ptr(), b.data_ptr(), a.numel(), int(inner)))


def row_absmean(x2d: torch.Tensor, scales: Sequence[torch.Tensor] = (), alpha: Optional[float] = None,
                prologue=None) -> torch.Tensor:
    """mean |residual| per row after folding ``scales`` (include/lsq_b200.h: lsq_row_absmean)."""
    require_cuda(x2d)
    x2d = x2d.contiguous()
    rows, length = x2d.shape
    out = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    tab = scale_table(scales, rows, x2d.device)
    L = _C.lib()
    need = L.lsq_reduce_workspace_bytes(rows, length)
    ws = workspace(x2d.device, need)
    with torch.cuda.device(x2d.device), _launch('row_absmean', 4.0 * rows * length):
        keep = []
        _C.check(L.lsq_row_absmean_ex(x2d.data_ptr(), rows, length, _alpha(alpha), _ptr(tab), len(scales),
                                      out.data_ptr(), ws.data_ptr(), ws.numel(), _prologue(prologue, keep), _stream()),
                 'lsq_row_absmean')
    return out


def solve_v1(x2d: torch.Tensor, ternary: bool, skip: int = 1, alpha: Optional[float] = None,
             diag: bool = False, prologue=None, out: Optional[torch.Tensor] = None):
    """Optimal v1 per row (lsq_solve_v1); returns [rows] (and the int32 [rows,4] diagnostics)."""
    require_cuda(x2d)
    x2d = x2d.contiguous()
    rows, length = x2d.shape
    if out is None:
        out = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    elif out.shape != (rows,) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x2d.device:
        raise ValueError(f'out must be a contiguous float32 [{rows}] tensor on {x2d.device}')
    dg = torch.zeros(rows, 16, dtype=torch.int32, device=x2d.device) if diag else None
    with torch.cuda.device(x2d.device), _launch('solve_v1', 4.0 * rows * length):
        keep = []
        _C.check(_C.lib().lsq_solve_v1_ex(x2d.data_ptr(), rows, length, int(skip), int(bool(ternary)), _alpha(alpha),
                                          out.data_ptr(), _ptr(dg), _prologue(prologue, keep), _stream()), 'lsq_solve_v1')
    return (out, dg) if diag else out


def _row_tensor_table(xs: Sequence[torch.Tensor]):
    """Host descriptor table (include/lsq_b200.h: lsq_row_tensor) for 2-D tensors on one device + their outputs."""
    if len(xs) == 0:
        raise ValueError('need at least one tensor')
    dev = xs[0].device
    keep, outs = [], []
    tab = (_C.RowTensor * len(xs))()
    for i, x in enumerate(xs):
        require_cuda(x)
        if x.device != dev or x.dim() != 2:
            raise ValueError('multi-tensor calls need 2-D tensors on one device')
        x = x.detach().contiguous()
        out = torch.empty(x.shape[0], dtype=torch.float32, device=dev)
        keep.append(x)
        outs.append(out)
        tab[i] = _C.RowTensor(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1])
    return dev, tab, keep, outs


def solve_v1_multi(xs: Sequence[torch.Tensor], ternary: bool, skip: int = 1,
                   alpha: Optional[float] = None) -> List[torch.Tensor]:
    """Optimal v1 per row of several [rows, len] tensors in one launch (lsq_solve_v1_multi); same values as
    ``solve_v1`` on each tensor."""
    dev, tab, keep, outs = _row_tensor_table(xs)
    total = float(sum(x.numel() for x in keep))
    with torch.cuda.device(dev), _launch('solve_v1_multi', 4.0 * total):
        _C.check(_C.lib().lsq_solve_v1_multi(tab, len(keep), int(skip), int(bool(ternary)), _alpha(alpha), _stream()),
                 'lsq_solve_v1_multi')
    return outs


def row_absmean_multi(xs: Sequence[torch.Tensor], alpha: Optional[float] = None) -> List[torch.Tensor]:
    """mean |x| per row of several [rows, len] tensors in one launch (lsq_row_absmean_multi)."""
    dev, tab, keep, outs = _row_tensor_table(xs)
    total = float(sum(x.numel() for x in keep))
    with torch.cuda.device(dev), _launch('row_absmean_multi', 4.0 * total):
        _C.check(_C.lib().lsq_row_absmean_multi(tab, len(keep), _alpha(alpha), _stream()), 'lsq_row_absmean_multi')
    return outs


def fakequant(x2d: torch.Tensor, scales: Sequence[torch.Tensor], ternary: bool = False,
              alpha: Optional[float] = None) -> torch.Tensor:
    """Dense sum_j s_j b_j in the reference's operation order (lsq_fakequant)."""
    require_cuda(x2d)
    x2d = x2d.contiguous()
    rows, length = x2d.shape
    nplanes = 2 if ternary else len(scales)
    tab = scale_table(scales, rows, x2d.device)
    out = torch.empty_like(x2d)
    with torch.cuda.device(x2d.device), _launch('fakequant', 8.0 * rows * length):
        _C.check(_C.lib().lsq_fakequant(x2d.data_ptr(), rows, length, _alpha(alpha), tab.data_ptr(), nplanes,
                                        int(bool(ternary)), out.data_ptr(), _stream()), 'lsq_fakequant')
    return out


def ste_backward(x: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    require_cuda(x)
    x = x.contiguous()
    grad_out = grad_out.contiguous().to(torch.float32)
    gin = torch.empty_like(x)
    with torch.cuda.device(x.device), _launch('ste_backward', 12.0 * x.numel()):
        _C.check(_C.lib().lsq_ste_backward(x.data_ptr(), grad_out.data_ptr(), gin.data_ptr(), x.numel(), _stream()),
                 'lsq_ste_backward')
    return gin


def act_geometry(n: int, c: int, h: int, w: int, kh: int, kw: int, stride: int, pad: int) -> Optional[_C.ActGeom]:
    """Packed-plane geometry, or None when the packed path does not cover the shape."""
    g = _C.ActGeom()
    st = _C.lib().lsq_act_geometry(n, c, h, w, kh, kw, stride, pad, C.byref(g))
    if st == -4:
        return None
    _C.check(st, 'lsq_act_geometry')
    return g


def encode_act(x: torch.Tensor, g: _C.ActGeom, scales: Sequence[torch.Tensor], nplanes: int,
               alpha: Optional[float] = None, want_next_scale: bool = False,
               planes: Optional[torch.Tensor] = None, prologue=None,
               next_scale_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """x [n,c,h,w] -> bit planes (int32 buffer) and optionally the next per-sample scale (lsq_encode_act)."""
    require_cuda(x)
    x = x.contiguous()
    L = _C.lib()
    nbytes = L.lsq_act_planes_bytes(C.byref(g), nplanes)
    if planes is None or planes.numel() * 4 < nbytes:
        planes = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=x.device)
    tab = scale_table(scales, g.n, x.device)
    nxt = None
    if want_next_scale:
        nxt = next_scale_out if next_scale_out is not None else torch.empty(g.n, dtype=torch.float32, device=x.device)
        if nxt.shape != (g.n,) or nxt.dtype != torch.float32 or not nxt.is_contiguous() or nxt.device != x.device:
            raise ValueError(f'next_scale_out must be a contiguous float32 [{g.n}] tensor on {x.device}')
    need = L.lsq_reduce_workspace_bytes(g.n, g.c * g.h * g.w)
    ws = workspace(x.device, need)
    with torch.cuda.device(x.device), _launch('encode_act', x.numel() * (4.0 + nplanes / 8.0)):
        keep = []
        _C.check(L.lsq_encode_act_ex(x.data_ptr(), C.byref(g), _alpha(alpha), _ptr(tab), len(scales), nplanes,
                                     planes.data_ptr(), _ptr(nxt), ws.data_ptr(), ws.numel(),
                                     _prologue(prologue, keep), _stream()), 'lsq_encode_act')
    return planes, nxt


def quantize_act(x: torch.Tensor, g: _C.ActGeom, ternary: bool, alpha: Optional[float] = None, skip: int = 3,
                 planes: Optional[torch.Tensor] = None, prologue=None, table: Optional[torch.Tensor] = None,
                 diag: bool = False):
    """Fused ls-2 / ls-T activation quantizer (lsq_quantize_act): x [n,c,h,w] -> (bit planes, scale table [2, n])
    with one read of x from HBM; ``diag=True`` also returns the int32 [n, 16] per-row diagnostics."""
    require_cuda(x)
    x = x.contiguous()
    L = _C.lib()
    nbytes = L.lsq_act_planes_bytes(C.byref(g), 2)
    if planes is None or planes.numel() * 4 < nbytes:
        planes = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=x.device)
    if table is None:
        table = torch.empty(2, g.n, dtype=torch.float32, device=x.device)
    elif tuple(table.shape) != (2, g.n) or table.dtype != torch.float32 or not table.is_contiguous() or table.device != x.device:
        raise ValueError(f'table must be a contiguous float32 [2, {g.n}] tensor on {x.device}')
    ws = workspace(x.device, L.lsq_quantize_act_workspace_bytes(C.byref(g)))
    dg = torch.zeros(g.n, 16, dtype=torch.int32, device=x.device) if diag else None
    with torch.cuda.device(x.device), _launch('quant_act', x.numel() * (4.0 + 2.0 / 8.0)):
        keep = []
        _C.check(L.lsq_quantize_act(x.data_ptr(), C.byref(g), _alpha(alpha), int(bool(ternary)), int(skip),
                                    planes.data_ptr(), table.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _prologue(prologue, keep), _ptr(dg), _stream()), 'lsq_quantize_act')
    # lsq_quantize_act launches two kernels either way: the fused kernel + its (normally empty) fallback pass, or the generic
    # solver + encoder for shapes outside the fused kernel's domain
    LAUNCHES['quant_act_second'] = LAUNCHES.get('quant_act_second', 0) + 1
    return (planes, table, dg) if diag else (planes, table)


def pack_weights(w: torch.Tensor) -> torch.Tensor:
    """sign(W) images for the convolution kernels (lsq_pack_weights)."""
    require_cuda(w, 'weight')
    w = w.detach().contiguous()
    cout, cin, kh, kw = w.shape
    L = _C.lib()
    nbytes = L.lsq_wpack_bytes(cout, cin, kh, kw)
    buf = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=w.device)
    off = (-buf.data_ptr()) % 1024
    view = buf[off:off + nbytes]
    with torch.cuda.device(w.device), _launch('pack_weights', 5.0 * w.numel()):
        _C.check(L.lsq_pack_weights(w.data_ptr(), cout, cin, kh, kw, view.data_ptr(), _stream()), 'lsq_pack_weights')
    return view


def weight_bits(w: torch.Tensor) -> torch.Tensor:
    """The deployable sign image of a weight tensor: int32 [cout, kh*kw, ceil(cin/32)], bit c&31 of word c>>5 set
    when W >= 0 (the `bits` part of lsq_pack_weights, include/lsq_b200.h)."""
    cout, cin, kh, kw = w.shape
    nbytes = _C.lib().lsq_wbits_bytes(cout, cin, kh, kw)
    return pack_weights(w)[:nbytes].clone().view(torch.int32).view(cout, kh * kw, (cin + 31) // 32)


def unpack_weights(bits: torch.Tensor, scale: Optional[torch.Tensor], shape) -> torch.Tensor:
    """Dense fp32 weights +-scale[co] from a sign image (lsq_unpack_weights)."""
    cout, cin, kh, kw = (int(v) for v in shape)
    if not bits.is_cuda or bits.dtype != torch.int32:
        raise _C.LsqError('ml_quant_b200: unpack_weights needs an int32 CUDA tensor')
    bits = bits.contiguous()
    if bits.numel() * 4 != _C.lib().lsq_wbits_bytes(cout, cin, kh, kw):
        raise ValueError(f'sign image of {bits.numel() * 4} bytes does not match weight shape {(cout, cin, kh, kw)}')
    if scale is not None:
        require_cuda(scale, 'scale')
        scale = scale.detach().contiguous()
        if scale.numel() != cout:
            raise ValueError(f'scale must have {cout} entries')
    w = torch.empty(cout, cin, kh, kw, dtype=torch.float32, device=bits.device)
    with torch.cuda.device(bits.device), _launch('unpack_weights', 4.0 * w.numel()):
        _C.check(_C.lib().lsq_unpack_weights(bits.data_ptr(), _ptr(scale), cout, cin, kh, kw, w.data_ptr(), _stream()),
                 'lsq_unpack_weights')
    return w


def bconv2d(planes: torch.Tensor, g: _C.ActGeom, nplanes: int, act_scales: torch.Tensor, wpack: torch.Tensor,
            w_scale: torch.Tensor, bias: Optional[torch.Tensor], cout: int, impl: int = 0,
            out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act: int = 0,
            prelu: Optional[torch.Tensor] = None, residual_after_act: bool = True) -> torch.Tensor:
    """Binary convolution forward (lsq_bconv2d_fwd_ex): act_scales is the [nplanes, n] table; the optional
    epilogue applies act (0 none, 1 ReLU, 2 PReLU) and adds ``residual`` after or before it."""
    if out is None:
        out = torch.empty(g.n, cout, g.ho, g.wo, dtype=torch.float32, device=planes.device)
    act_scales = act_scales.contiguous()
    w_scale = w_scale.detach().contiguous()
    b = None if bias is None else bias.detach().contiguous()
    n_out = g.n * cout * g.ho * g.wo
    macs = float(n_out) * g.c * g.kh * g.kw * nplanes
    name = 'bconv_tc' if (impl == 2 or (impl == 0 and tc_supported(g, nplanes, cout))) else 'bconv_simple'
    nres = 4.0 * n_out if residual is not None else 0.0
    with torch.cuda.device(planes.device), _launch(name, 4.0 * n_out + nres + nplanes * g.n * g.c * g.h * g.w / 8.0, 2.0 * macs):
        if residual is not None:
            residual = residual.contiguous()
            if tuple(residual.shape) != tuple(out.shape):
                raise ValueError(f'residual shape {tuple(residual.shape)} != output shape {tuple(out.shape)}')
        pr = None if prelu is None else prelu.detach().contiguous()
        epi = _C.Epilogue(_ptr(residual), _ptr(pr), 0 if pr is None else pr.numel(), int(act), int(bool(residual_after_act)))
        _C.check(_C.lib().lsq_bconv2d_fwd_ex(planes.data_ptr(), C.byref(g), nplanes, act_scales.data_ptr(),
                                             wpack.data_ptr(), w_scale.data_ptr(), _ptr(b), cout, out.data_ptr(),
                                             int(impl), C.byref(epi), _stream()), 'lsq_bconv2d_fwd')
    return out


def tc_supported(g: _C.ActGeom, nplanes: int, cout: int) -> bool:
    return bool(_C.lib().lsq_bconv2d_tc_supported(C.byref(g), nplanes, cout))


def stem_supported(n: int, h: int, w: int) -> bool:
    return bool(_C.lib().lsq_stem_supported(int(n), int(h), int(w)))


def stem_pack(w: torch.Tensor) -> torch.Tensor:
    """Tensor-core operand image of folded stem weights [64, 3, 7, 7] (lsq_stem_pack_weights)."""
    require_cuda(w, 'weight')
    if tuple(w.shape) != (64, 3, 7, 7):
        raise ValueError('stem_pack handles 3 -> 64 channel 7x7 stems only')
    w = w.detach().contiguous()
    L = _C.lib()
    image = torch.empty(L.lsq_stem_image_bytes() // 4, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device), _launch('stem_pack', 4.0 * w.numel()):
        _C.check(L.lsq_stem_pack_weights(w.data_ptr(), image.data_ptr(), _stream()), 'lsq_stem_pack_weights')
    return image


_stem_ws = {}


def stem_fwd(x: torch.Tensor, image: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """maxpool3x3s2p1(relu(conv7x7s2p3(x, w) + bias)) (lsq_stem_fwd); image comes from stem_pack.
    Images up to 250 pixels wide run as one kernel (|x| clamped to the fp16 range 65504)."""
    require_cuda(x, 'x')
    x = x.contiguous()
    n, c, h, w = x.shape
    if c != 3:
        raise ValueError('stem_fwd handles 3 -> 64 channel 7x7 stems only')
    L = _C.lib()
    hc, wc = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    hp, wp = (hc - 1) // 2 + 1, (wc - 1) // 2 + 1
    key = (x.device.index, _stream())
    ws = _stem_ws.get(key)
    need = L.lsq_stem_workspace_bytes(n, h, w) // 4
    if ws is None or ws.numel() < need:
        ws = _stem_ws[key] = torch.empty(need, dtype=torch.float32, device=x.device)
    out = torch.empty(n, 64, hp, wp, dtype=torch.float32, device=x.device)
    macs = float(n) * 64 * hc * wc * 147
    with torch.cuda.device(x.device), _launch('stem', 4.0 * (x.numel() + out.numel()), 2.0 * macs):
        _C.check(L.lsq_stem_fwd(x.data_ptr(), n, h, w, image.data_ptr(), bias.contiguous().data_ptr(), ws.data_ptr(),
                                out.data_ptr(), _stream()), 'lsq_stem_fwd')
    if not L.lsq_stem_is_fused(n, h, w):
        LAUNCHES['stem_pool'] = LAUNCHES.get('stem_pool', 0) + 1      # the two-kernel route (conv, pool) of wide images
    return out


def plane_mean(x: torch.Tensor) -> torch.Tensor:
    """Global average pool [n, c, h, w] -> [n, c] (lsq_plane_mean: one warp per plane, fixed summation tree)."""
    require_cuda(x, 'x')
    if x.dim() != 4:
        raise ValueError('plane_mean takes a [n, c, h, w] tensor')
    x = x.contiguous()
    n, c, h, w = x.shape
    out = torch.empty(n, c, dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return out
    with torch.cuda.device(x.device), _launch('plane_mean', 4.0 * x.numel()):
        _C.check(_C.lib().lsq_plane_mean(x.data_ptr(), n * c, h * w, out.data_ptr(), _stream()), 'lsq_plane_mean')
    return out


def u8_expand(x: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """fp32 tensor lut[c][x] of a uint8 tensor [n, c, ...] (lsq_u8_expand): the caller's ToTensor + Normalize, evaluated
    once per pixel level in ``lut`` float[c][256], applied on the device."""
    if not isinstance(x, torch.Tensor) or x.dtype != torch.uint8 or x.dim() < 2:
        raise ValueError('u8_expand takes a uint8 tensor [n, c, ...]')
    require_cuda(x, 'x', torch.uint8)
    x = x.contiguous()
    n, c = x.shape[0], x.shape[1]
    lut = lut.to(device=x.device, dtype=torch.float32).contiguous()
    if tuple(lut.shape) != (c, 256):
        raise ValueError(f'lut must be [{c}, 256]')
    inner = x.numel() // max(n * c, 1)
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return out
    with torch.cuda.device(x.device), _launch('u8_expand', 5.0 * x.numel()):
        _C.check(_C.lib().lsq_u8_expand(x.data_ptr(), n * c, c, inner, lut.data_ptr(), out.data_ptr(), _stream()), 'lsq_u8_expand')
    return out


def stem_fwd_u8(x: torch.Tensor, lut: torch.Tensor, image: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """stem_fwd on lut[c][x] for uint8 pixels x [n, 3, h, w] without materialising the fp32 image (lsq_stem_fwd_u8; images up
    to 250 pixels wide -- check with lsq_stem_is_fused -- else u8_expand + stem_fwd)."""
    if not isinstance(x, torch.Tensor) or x.dtype != torch.uint8 or x.dim() != 4 or x.shape[1] != 3:
        raise ValueError('stem_fwd_u8 takes uint8 pixels [n, 3, h, w]')
    require_cuda(x, 'x', torch.uint8)
    x = x.contiguous()
    n, _, h, w = x.shape
    L = _C.lib()
    lut = lut.to(device=x.device, dtype=torch.float32).contiguous()
    if tuple(lut.shape) != (3, 256):
        raise ValueError('lut must be [3, 256]')
    hc, wc = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    hp, wp = (hc - 1) // 2 + 1, (wc - 1) // 2 + 1
    out = torch.empty(n, 64, hp, wp, dtype=torch.float32, device=x.device)
    macs = float(n) * 64 * hc * wc * 147
    with torch.cuda.device(x.device), _launch('stem', 1.0 * x.numel() + 4.0 * out.numel(), 2.0 * macs):
        _C.check(L.lsq_stem_fwd_u8(x.data_ptr(), n, h, w, lut.data_ptr(), image.data_ptr(), bias.contiguous().data_ptr(),
                                   out.data_ptr(), _stream()), 'lsq_stem_fwd_u8')
    return out


def pwconv_supported(cin: int, cout: int) -> bool:
    return bool(_C.lib().lsq_pwconv_supported(int(cin), int(cout)))


def pwconv_pack(w2d: torch.Tensor) -> torch.Tensor:
    """Tensor-core operand image of pointwise-convolution weights [cout, cin] (lsq_pwconv_pack_weights)."""
    require_cuda(w2d, 'weight')
    w2d = w2d.detach().contiguous()
    cout, cin = w2d.shape
    L = _C.lib()
    image = torch.empty(L.lsq_pwconv_image_bytes(cout, cin) // 4, dtype=torch.float32, device=w2d.device)
    with torch.cuda.device(w2d.device), _launch('pwconv_pack', 12.0 * w2d.numel()):
        _C.check(L.lsq_pwconv_pack_weights(w2d.data_ptr(), cout, cin, image.data_ptr(), _stream()), 'lsq_pwconv_pack_weights')
    return image


def pwconv_fwd(x: torch.Tensor, image: torch.Tensor, bias: torch.Tensor, cout: int, stride: int) -> torch.Tensor:
    """y = conv1x1(x, w, stride) + bias in fp32 (3xTF32 on the tensor cores; lsq_pwconv_fwd)."""
    require_cuda(x, 'x')
    x = x.contiguous()
    n, cin, h, w = x.shape
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    out = torch.empty(n, cout, ho, wo, dtype=torch.float32, device=x.device)
    macs = float(n) * cout * ho * wo * cin
    with torch.cuda.device(x.device), _launch('pwconv', 4.0 * (n * cin * ho * wo + out.numel()), 2.0 * macs):
        _C.check(_C.lib().lsq_pwconv_fwd(x.data_ptr(), n, cin, h, w, int(stride), image.data_ptr(), bias.contiguous().data_ptr(),
                                         int(cout), out.data_ptr(), _stream()), 'lsq_pwconv_fwd')
    return out
