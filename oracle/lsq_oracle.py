"""CPU oracle for the binary-quantized inference path of apple/ml-quant.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ml_quant_b200/`` imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may.  The product path is the CUDA library and fails loudly without it.

What it is: a functional restatement, in torch *CPU* ops, of the reference algorithm for the
hot path.  The reference itself is pure PyTorch, so issuing the same
ATen operators in the same order on the same dtype reproduces its CPU results bit for bit;
that is the strongest oracle available and is why torch-CPU (not numpy) is used here.
Each function cites the reference lines it follows (paths relative to /root/reference).

Parity pin: ``tests/golden/*.pt`` hold input/output vectors produced by importing the
*unmodified* reference in the build container (``oracle/gen_golden.py``).
``tests/test_oracle.py`` checks this file against them bit-exactly, against the
reference's own known-answer tests, and (when /root/reference is present) against the live
reference on fresh random inputs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- sign / STE
def sign_pm1(x: Tensor) -> Tensor:
    """+1 where x >= 0, -1 where x < 0 (sign(0) = +1).  quant/binary/ste.py:16-18."""
    return x.sign() + (x == 0).type(torch.float)


def ste_grad(x: Tensor, grad_out: Tensor) -> Tensor:
    """Straight-through gradient: pass where -1 <= x <= 1.  quant/binary/ste.py:50-66."""
    g = grad_out.clone()
    g[x.gt(1)] = 0
    g[x.lt(-1)] = 0
    return g


def clamp_symmetric(x: Tensor, alpha: float) -> Tensor:
    """quant/binary/quantization.py:22-24."""
    return x.clamp(-alpha, alpha)


# --------------------------------------------------------------------------- LS solve
def candidate_mask(abs_rows: Tensor, ternary: bool) -> Tuple[Tensor, Tensor]:
    """Sorted values and the interior-candidate mask.  quant/binary/optimal.py:41-83.

    abs_rows: [R, n] non-negative.  Returns (sorted [R, n], mask [R, n-2]); mask[:, j] refers to
    sorted position j+1.  A position i is a candidate when the half-sum of the prefix mean and the
    suffix mean (2-bit only) or half the suffix mean (both) falls in [a_i, a_{i+1}].
    """
    srt, _ = torch.sort(abs_rows, dim=1)
    run = srt.cumsum(dim=1)
    n = abs_rows.shape[1]
    k_lo = torch.arange(1, n + 1, device=abs_rows.device)
    k_hi = torch.flip(k_lo, [0]) - 1
    k_hi[-1] = 1  # optimal.py:61 -- last column is never used, avoid 0-division
    hi_half = ((run[:, -1:] - run) / k_hi)[:, 1:-1]
    inner, nxt = srt[:, 1:-1], srt[:, 2:]
    if ternary:
        hi_half = 0.5 * hi_half
        mask = (inner <= hi_half) * (hi_half <= nxt)
    else:
        mid = 0.5 * ((run / k_lo)[:, 1:-1] + hi_half)
        hi_half = 0.5 * hi_half
        mask = (inner <= hi_half) * (hi_half <= nxt)
        mask = mask + (inner <= mid) * (mid <= nxt)
    return srt, mask


def candidate_table(abs_rows: Tensor, ternary: bool) -> Tuple[Tensor, Tensor]:
    """Per-row candidate values, ascending, zero padded to the batch-wide maximum count.

    Follows optimal.py:135-148 (masked_select -> split -> pad_sequence) including the ternary
    edge case optimal.py:86-118 (append mean/2 when min > mean/2).  Returns (table [R, C], counts).
    """
    srt, mask = candidate_mask(abs_rows, ternary)
    inner = srt[:, 1:-1]
    counts = mask.sum(dim=1)
    extra = torch.zeros_like(counts, dtype=torch.bool)
    half_mean = None
    if ternary:
        means = abs_rows.mean(dim=1)
        mins, _ = abs_rows.min(dim=1)
        extra = mins > 0.5 * means
        # optimal.py:116: rows_mean[i].item() / 2 is a python double, rounded back to fp32
        half_mean = (means.double() / 2).float()
    total = counts + extra.long()
    width = int(total.max().item()) if total.numel() else 0
    table = torch.zeros(abs_rows.shape[0], width, dtype=abs_rows.dtype)
    # stable: ascending order of position == ascending order of value
    pos = torch.cumsum(mask.long(), dim=1) - 1
    rr, cc = torch.nonzero(mask, as_tuple=True)
    table[rr, pos[rr, cc]] = inner[rr, cc]
    if ternary and bool(extra.any()):
        er = torch.nonzero(extra, as_tuple=True)[0]
        table[er, counts[er]] = half_mean[er]
    return table, total


def candidate_cost(abs_rows: Tensor, table: Tensor, ternary: bool) -> Tensor:
    """L2 residual of every candidate.  quant/binary/optimal.py:16-38 (fp32, same op order)."""
    a = abs_rows.view(abs_rows.shape[0], 1, -1)
    c = table.view(table.shape[0], table.shape[1], 1)
    r = a - c * sign_pm1(a)
    v2 = c if ternary else r.abs().mean(dim=-1, keepdim=True)
    return torch.norm(r - v2 * sign_pm1(r), dim=-1)


def solve_v1(rows: Tensor, ternary: bool, skip: int = 1, chunk: int = 0) -> Tensor:
    """Optimal v1 per row, shape [R, 1].  quant/binary/optimal.py:121-155.

    ``chunk`` > 0 evaluates the [rows, candidates, n] cost tensor ``chunk`` rows at a time (the
    reference materialises it whole); results are identical because every row is independent
    once the batch-wide padding width is fixed.
    """
    with torch.no_grad():
        a = rows[..., ::skip].abs()
        table, _ = candidate_table(a, ternary)
        if table.shape[1] == 0:
            # pad_sequence of empty splits: the reference would raise; define as v1 = 0 (ls-1)
            return torch.zeros(rows.shape[0], 1, dtype=rows.dtype)
        if chunk <= 0:
            cost = candidate_cost(a, table, ternary)
        else:
            cost = torch.cat([candidate_cost(a[i:i + chunk], table[i:i + chunk], ternary)
                              for i in range(0, a.shape[0], chunk)])
        pick = torch.argmin(cost, dim=-1, keepdim=True)
        return torch.gather(table, 1, pick)


# --------------------------------------------------------------------------- quantizers
def quant_ls1(x: Tensor, v1: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """quant/binary/quantization.py:35-56.  x is 4-D; scale per index of dim 0."""
    if v1 is None:
        v1 = x.detach().abs().mean(dim=-1).mean(dim=-1).mean(dim=-1)
    return v1, v1.view(-1, 1, 1, 1) * sign_pm1(x)


def quant_ls2(x: Tensor, v1: Optional[Tensor] = None, v2: Optional[Tensor] = None,
              skip: int = 3, chunk: int = 0) -> Tuple[Tensor, Tensor, Tensor]:
    """quant/binary/quantization.py:59-92."""
    flat = x.detach().reshape(x.shape[0], -1)
    v1 = solve_v1(flat, False, skip, chunk) if v1 is None else v1.view(-1, 1)
    if v2 is None:
        v2 = (flat - v1 * sign_pm1(flat)).abs().mean(dim=-1, keepdim=True)
    else:
        v2 = v2.view(-1, 1)
    s1 = v1.view(-1, 1, 1, 1)
    b1 = sign_pm1(x)
    return v1.view(-1), v2.view(-1), s1 * b1 + v2.view(-1, 1, 1, 1) * sign_pm1(x - s1 * b1)


def quant_lsT(x: Tensor, v1: Optional[Tensor] = None, skip: int = 3,
              chunk: int = 0) -> Tuple[Tensor, Tensor]:
    """quant/binary/quantization.py:95-115."""
    flat = x.detach().reshape(x.shape[0], -1)
    if v1 is None:
        v1 = solve_v1(flat, True, skip, chunk)
    s1 = v1.view(-1, 1, 1, 1)
    b1 = sign_pm1(x)
    return v1.view(-1), s1 * (b1 + sign_pm1(x - s1 * b1))


def quant_gf(x: Tensor, k: int, vs: Optional[Sequence[Tensor]] = None) -> Tuple[List[Tensor], Tensor]:
    """Greedy foldable k-bit.  quant/binary/quantization.py:118-148."""
    if vs is not None and len(vs) != k:
        raise ValueError('all of v_1..v_k must be given')
    res = x.detach().reshape(x.shape[0], -1).clone()
    out = 0
    kept = []
    for i in range(k):
        v = vs[i] if vs is not None else res.abs().mean(dim=-1)
        kept.append(v)
        res = res - v.view(-1, 1) * sign_pm1(res)
        out = out + v.view(-1, 1, 1, 1) * sign_pm1(x - out)
    return kept, out


def bit_planes(x: Tensor, scheme: str, scales: Sequence[Tensor]) -> List[Tensor]:
    """The +-1 planes b_j of the fake-quant value  x_q = sum_j s_j * b_j  as boolean (b_j == +1).

    ls-1: [x>=0].  ls-2: [x>=0, (x - v1*b1)>=0] (quantization.py:89-92).  ls-T: same planes, both
    scaled by v1 (:113-115).  gf-k: b_j = sign(x - sum_{i<j} v_i b_i) (:139-146).
    """
    planes, acc = [], 0
    nb = {'ls-1': 1, 'ls-2': 2, 'ls-T': 2}.get(scheme) or int(scheme.split('-')[1])
    sc = list(scales) + ([scales[0]] if scheme == 'ls-T' else [])
    for j in range(nb):
        b = sign_pm1(x - acc)
        planes.append(b > 0)
        acc = acc + sc[j].view(-1, 1, 1, 1) * b
    return planes


def quantize_activation_stored(x: Tensor, scheme: str, scales: Sequence[Tensor]) -> Tensor:
    """Eval-mode activation quantizer with GIVEN per-sample scales: the moving-average path
    (quant/binary/activation_quantization.py:90-98, _moving_average_quantization :141-145, :172-176,
    :203-207, :234-239) and the public ``v1=`` / ``vs=`` arguments of the quantizer functions."""
    if scheme == 'fp':
        return x
    if scheme == 'ls-1':
        return quant_ls1(x, scales[0])[1]
    if scheme == 'ls-2':
        return quant_ls2(x, scales[0], scales[1])[2]
    if scheme == 'ls-T':
        return quant_lsT(x, scales[0])[1]
    return quant_gf(x, int(scheme.split('-')[1]), list(scales))[1]


def quantize_activation(x: Tensor, scheme: str, chunk: int = 16) -> Tuple[List[Tensor], Tensor]:
    """Eval-mode, moving_average_mode='off' activation quantizer: scales are re-solved per sample.
    quant/binary/activation_quantization.py:99-100 and the _batch_quantization methods."""
    if scheme == 'fp':
        return [], x
    if scheme == 'ls-1':
        v1, q = quant_ls1(x)
        return [v1], q
    if scheme == 'ls-2':
        v1, v2, q = quant_ls2(x, chunk=chunk)
        return [v1, v2], q
    if scheme == 'ls-T':
        v1, q = quant_lsT(x, chunk=chunk)
        return [v1], q
    vs, q = quant_gf(x, int(scheme.split('-')[1]))
    return list(vs), q


def quantize_weight(w: Tensor, scheme: str, scales: Optional[Sequence[Tensor]] = None,
                    skip: int = 3) -> Tuple[List[Tensor], Tensor]:
    """Weight quantizer; ``scales=None`` is the train-mode solve, else the eval-mode reuse of the
    cached buffers.  quant/binary/weight_quantization.py:27-34,51-59,75-82,100-109."""
    if scheme == 'fp':
        return [], w
    if scheme == 'ls-1':
        v1, q = quant_ls1(w, None if scales is None else scales[0])
        return [v1], q
    if scheme == 'ls-2':
        v1, v2, q = quant_ls2(w, *(scales if scales is not None else (None, None)), skip=skip)
        return [v1, v2], q
    if scheme == 'ls-T':
        v1, q = quant_lsT(w, None if scales is None else scales[0], skip=skip)
        return [v1], q
    vs, q = quant_gf(w, int(scheme.split('-')[1]), scales)
    return list(vs), q


def ema_step(avg: Tensor, momentum: Tensor, x: Tensor, seen: int) -> Tensor:
    """quant/utils/moving_average.py:26-37: first call copies, later momentum*old+(1-momentum)*new."""
    if seen > 0:
        return momentum * avg + (torch.ones_like(momentum) - momentum) * x
    return x.clone()


# --------------------------------------------------------------------------- QuantConv2d forward
def quant_conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor], x_scheme: str, w_scheme: str,
                 w_scales: Optional[Sequence[Tensor]], alpha: Optional[float] = None,
                 stride=1, padding=0, dilation=1, groups=1, chunk: int = 16,
                 x_scales: Optional[Sequence[Tensor]] = None, record: Optional[list] = None) -> Tensor:
    """Eval-mode QuantConv2d.forward.  quant/binary/binary_conv.py:161-173.  ``x_scales``: stored activation
    scales (moving-average modes) instead of the per-sample solve; ``record``: receives the scales used."""
    xin = x if alpha is None else clamp_symmetric(x, alpha)
    if x_scales is not None:
        xq, used = quantize_activation_stored(xin, x_scheme, x_scales), list(x_scales)
    else:
        used, xq = quantize_activation(xin, x_scheme, chunk)
    if record is not None:
        record.append([u.clone() for u in used])
    _, wq = quantize_weight(weight, w_scheme, w_scales)
    return F.conv2d(xq, wq, bias, stride, padding, dilation, groups)


def plane_conv_identity(x: Tensor, weight: Tensor, bias: Optional[Tensor], x_scheme: str,
                        x_scales: Sequence[Tensor], w_v1: Tensor, stride=1, padding=0) -> Tuple[Tensor, List[Tensor]]:
    """Integer formulation the CUDA kernel uses, for ls-1 weights:
        y[n,c] = vw[c] * sum_j s_j[n] * I_j[n,c] + bias[c],  I_j = conv(b_j, sign(W)) exactly integer.
    Returns (y, [I_j]).  Zero padding contributes 0 (binary_conv.py:165-173 pads the quantized tensor).
    """
    sw = sign_pm1(weight)
    planes = bit_planes(x, x_scheme, x_scales)
    sc = list(x_scales) + ([x_scales[0]] if x_scheme == 'ls-T' else [])
    ints, acc = [], 0
    for b, s in zip(planes, sc):
        pm = b.to(torch.float64) * 2 - 1
        i_j = F.conv2d(pm, sw.double(), None, stride, padding)
        ints.append(i_j.round().to(torch.int32))
        acc = acc + s.view(-1, 1, 1, 1).double() * i_j
    y = w_v1.view(1, -1, 1, 1).double() * acc
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1).double()
    return y.float(), ints


# --------------------------------------------------------------------------- callers (whole nets)
def _bn_eval(x: Tensor, sd: Dict[str, Tensor], p: str, eps: float = 1e-5) -> Tensor:
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd.get(p + '.weight'), sd.get(p + '.bias'), False, 0.0, eps)


def _nonlin(x: Tensor, kind: str, sd: Dict[str, Tensor], p: str) -> Tensor:
    if kind == 'relu':
        return F.relu(x)
    if kind == 'prelu':
        return F.prelu(x, sd[p + '.weight'])
    return x


def _qconv(x: Tensor, sd: Dict[str, Tensor], p: str, cfg: dict, stride: int, padding: int,
           record: Optional[list] = None) -> Tensor:
    w_s = cfg['w_quant']
    nsc = {'fp': 0, 'ls-1': 1, 'ls-2': 2, 'ls-T': 1}.get(w_s)
    if nsc is None:
        nsc = int(w_s.split('-')[1])
    scales = [sd[f'{p}.w_approximate.v{i + 1}'] for i in range(nsc)]
    clamp = cfg.get('clamp') or {'kind': 'identity'}
    alpha = clamp.get('alpha', 2) if clamp['kind'] == 'symmetric' else None
    return quant_conv2d(x, sd[p + '.weight'], sd.get(p + '.bias'), cfg['x_quant'], w_s, scales,
                        alpha, stride, padding, record=record)


def resnet_forward(sd: Dict[str, Tensor], arch: dict, x: Tensor, record: Optional[list] = None) -> Tensor:
    """Eval forward of QResNet from a state_dict and the YAML ``arch_config``.
    quant/models/resnet.py:91-97 (regular), :180-190 (xnor), :283-340,393-397 (stem/classifier).
    ``record`` (a list) receives the activation scales of every QuantConv2d in call order."""
    l0 = arch['layer0']
    x = F.conv2d(x, sd['conv1.weight'], sd.get('conv1.bias'), l0['stride'], l0['padding'])
    x = F.relu(_bn_eval(x, sd, 'bn1'))
    mp = l0['maxpool']
    if mp['type'] == 'maxpool2d':
        x = F.max_pool2d(x, mp['kernel_size'], mp['stride'], mp['padding'])
    nl = arch['nonlins']
    bi = 1
    planes = l0['n_in_channels']
    layers = [arch['layer1'], arch['layer2'], arch['layer3'], arch.get('layer4')]
    for li, (cfg, nb) in enumerate(zip(layers, arch['num_blocks'])):
        if cfg is None:
            continue
        out_planes = l0['n_in_channels'] * (2 ** li)
        for j in range(nb):
            stride = (1 if li == 0 else 2) if j == 0 else 1
            p = f'blocks.{bi}'
            has_sc = (p + '.shortcut.0.weight') in sd

            def shortcut(t):
                if not has_sc:
                    return t
                t = F.conv2d(t, sd[p + '.shortcut.0.weight'], sd.get(p + '.shortcut.0.bias'), stride)
                return _bn_eval(t, sd, p + '.shortcut.1')

            if arch['block'] == 'xnor':
                o1 = _nonlin(_qconv(_bn_eval(x, sd, p + '.bn1'), sd, p + '.conv1', cfg, stride, 1, record),
                             nl[0], sd, p + '.nonlin1')
                if cfg.get('double_shortcut', False):
                    o1 = o1 + shortcut(x)
                o2 = _qconv(_bn_eval(o1, sd, p + '.bn2'), sd, p + '.conv2', cfg, 1, 1, record)
                if cfg.get('double_shortcut', False):
                    x = _nonlin(o2, nl[1], sd, p + '.nonlin2') + o1
                else:
                    x = _nonlin(o2 + shortcut(x), nl[1], sd, p + '.nonlin2')
            else:
                o = _nonlin(_bn_eval(_qconv(x, sd, p + '.conv1', cfg, stride, 1, record), sd, p + '.bn1'),
                            nl[0], sd, p + '.nonlin1')
                o = _bn_eval(_qconv(o, sd, p + '.conv2', cfg, 1, 1, record), sd, p + '.bn2')
                x = _nonlin(o + shortcut(x), nl[1], sd, p + '.nonlin2')
            planes = out_planes
            bi += 1
    x = F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)
    return F.linear(x, sd['linear_classifier.2.weight'], sd['linear_classifier.2.bias'])


def lenet_forward(sd: Dict[str, Tensor], arch: dict, x: Tensor) -> Tensor:
    """Eval forward of QLeNet5.  quant/models/lenet.py:78-94."""
    c2 = sd['conv2.weight'].shape[0]
    x = F.conv2d(x, sd['conv1.weight'], sd['conv1.bias'])
    x = _bn_eval(F.relu(x), sd, 'bn_conv1', 1e-4)
    x = F.max_pool2d(x, 2, 2)
    cfg = {'x_quant': arch.get('x_quant', 'fp'), 'w_quant': arch.get('w_quant', 'fp'),
           'clamp': arch.get('clamp')}
    x = F.relu(_qconv(_bn_eval(x, sd, 'bn_conv2', 1e-4), sd, 'conv2', cfg, 1, 0))
    x = F.max_pool2d(x, 2, 2).reshape(-1, c2 * 16)
    x = F.relu(F.linear(x, sd['fc1.weight'], sd['fc1.bias']))
    return F.log_softmax(F.linear(x, sd['fc2.weight'], sd['fc2.bias']), dim=1)


# --------------------------------------------------------------------------- staged-parity helpers
def exact_cost(rows: Tensor, v1: Tensor, ternary: bool, skip: int = 1) -> Tensor:
    """fp64 cost (SURVEY.md section 3.4 closed meaning) of a given v1 on the solve's view of the row:
    ||r - v2*sign(r)||_2 with r = |x| - v1, v2 = mean|r| (2-bit) or v1 (ternary)."""
    a = rows[..., ::skip].abs().double()
    c = v1.view(-1, 1).double()
    r = a - c
    v2 = c if ternary else r.abs().mean(dim=1, keepdim=True)
    s = torch.where(r >= 0, 1.0, -1.0).double()
    return torch.linalg.vector_norm(r - v2 * s, dim=1)
