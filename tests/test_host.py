"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol of include/lsq_b200.h,
geometry arithmetic, host logic of the reference-API mirror (construction, validation, state_dict
layout, moving average), loud failure on CPU tensors, and the multi-process sharding helpers (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from ml_quant_b200 import _C
    hdr = open(os.path.join(ROOT, 'include', 'lsq_b200.h')).read()
    declared = set(re.findall(r'LSQ_API\s+[\w\s\*]+?\b(lsq_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_C.EXPORTS), declared ^ set(_C.EXPORTS)
    L = _C.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.lsq_abi_version() == 1


def test_argument_errors_are_reported_without_a_gpu():
    from ml_quant_b200 import _C
    L = _C.lib()
    assert L.lsq_row_absmean(None, 1, 1, 0.0, None, 0, None, None, 0, None) == -1
    assert b'null pointer' in L.lsq_last_error()
    g = _C.ActGeom()
    assert L.lsq_act_geometry(1, 64, 8, 8, 3, 3, 3, 1, ctypes.byref(g)) == -4     # stride 3: not packed


def test_multi_tensor_and_packed_weight_argument_checks_without_a_gpu():
    """Entry points added for the weight sweep / packed checkpoint validate their arguments before any CUDA call."""
    from ml_quant_b200 import _C, runtime
    L = _C.lib()
    assert L.lsq_solve_v1_multi(None, 1, 3, 0, 0.0, None) == -1
    tab = (_C.RowTensor * 2)()
    tab[0] = _C.RowTensor(0x1000, 0x2000, 4, 64)
    tab[1] = _C.RowTensor(0x1000, None, 4, 64)                     # second tensor has no output vector
    assert L.lsq_solve_v1_multi(tab, 2, 3, 0, 0.0, None) == -1
    assert b'tensor 1' in L.lsq_last_error()
    assert L.lsq_row_absmean_multi(tab, 2, 0.0, None) == -1
    assert L.lsq_solve_v1_multi(tab, 1, 0, 0, 0.0, None) == -1     # skip < 1
    assert L.lsq_wbits_bytes(64, 64, 3, 3) == 64 * 9 * 2 * 4
    assert L.lsq_wbits_bytes(50, 20, 5, 5) == 50 * 25 * 1 * 4
    assert L.lsq_wbits_bytes(0, 20, 5, 5) == 0
    assert L.lsq_unpack_weights(None, None, 64, 64, 3, 3, None, None) == -1
    model = runtime.build_model('mnist_lenet5_ls1w_fpa')
    with pytest.raises(ValueError):
        runtime.load_packed(model, {'conv2.weight_bits': torch.zeros(1, dtype=torch.int32)})      # no format tag
    with pytest.raises(Exception):
        runtime.export_packed(model)                               # packing runs on the GPU: CPU weights fail loudly


@pytest.mark.parametrize('shape', [(512, 64, 56, 56, 3, 1, 1), (4, 64, 56, 56, 3, 2, 1), (2, 128, 7, 9, 3, 2, 1),
                                   (4, 20, 12, 12, 5, 1, 0), (3, 16, 9, 9, 3, 2, 1), (2, 32, 5, 5, 1, 1, 0)])
def test_geometry_positions_are_unique_and_padding_is_shared(shape):
    """Every (sample, phase row, phase col) maps to its own position; every tap of every valid output
    lands either on a valid position of the right phase or on a padding position."""
    from ml_quant_b200 import ops
    n, c, h, w, k, s, p = shape
    n = min(n, 3)
    g = ops.act_geometry(n, c, h, w, k, k, s, p)
    assert g.ho == (h + 2 * p - k) // s + 1 and g.wo == (w + 2 * p - k) // s + 1

    def vpos(smp, a, b):
        return g.lead + (smp * g.rows_per_sample + g.ph + a) * g.pitch + b

    owner = {}
    for smp in range(n):
        for yi in range(h):
            for xi in range(w):
                ph = ((yi % s) * 2 + (xi % s)) if s == 2 else 0
                key = (ph, vpos(smp, yi // s, xi // s))
                assert key not in owner
                owner[key] = (smp, yi, xi)
    assert max(v for _, v in owner) < g.vtot
    for smp in range(n):
        for yo in range(g.ho):
            for xo in range(g.wo):
                q = vpos(smp, yo, xo)
                for dy in range(k):
                    for dx in range(k):
                        ey, ex = dy - p, dx - p
                        qy, qx = ey // s, ex // s
                        ph = ((ey - qy * s) * 2 + (ex - qx * s)) if s == 2 else 0
                        yi, xi = yo * s + ey, xo * s + ex
                        hit = owner.get((ph, q + qy * g.pitch + qx))
                        if 0 <= yi < h and 0 <= xi < w:
                            assert hit == (smp, yi, xi)
                        else:
                            assert hit is None


def test_quantconv_construction_and_validation():
    # reference tests/binary/test_binary_conv.py:70-107
    import itertools
    from quant.binary.binary_conv import QuantConv2d
    clamp = {'alpha': 2, 'kind': 'symmetric'}
    c = QuantConv2d('ls-2', 'ls-1', 3, 1, (4, 4), clamp=clamp, stride=4, bias=False)
    assert len(c.quantized_parameters['fp']) == 0 and len(c.quantized_parameters['ls-1']) == 1
    assert set(c.quantized_parameters.keys()) - {'fp', 'ls-1'} == set() and len(list(c.parameters())) == 1
    c = QuantConv2d('ls-2', 'ls-2', 3, 1, (4, 4), clamp=clamp, stride=4)
    assert len(c.quantized_parameters['fp']) == 1 and len(c.quantized_parameters['ls-2']) == 1
    schemes = ['fp', 'ls-1', 'ls-2', 'ls-T', 'gf-2', 'gf-3']
    for xs, ws in itertools.product(schemes, schemes):
        QuantConv2d(xs, ws, 3, 1, (4, 4))
    for bad in [('ls', 'ls-1'), ('l2', 'ls-1'), ('ls-1', 'ls-3'), ('ls-1', 'l2')]:
        with pytest.raises(ValueError):
            QuantConv2d(bad[0], bad[1], 3, 1, (4, 4))
    with pytest.raises(ValueError):
        QuantConv2d('ls-1', 'ls-2', 3, 1, (4, 4), clamp={'kind': 'sym'})


def test_fp_quantconv_equals_conv2d_on_cpu():
    # reference tests/binary/test_binary_conv.py:18-38 -- the fp/fp route has no quantizer in it
    from quant.binary.binary_conv import QuantConv2d
    torch.manual_seed(1234)
    x = torch.randn(2, 3, 20, 20, requires_grad=True)
    x2 = x.clone().detach().requires_grad_(True)
    ref = nn.Conv2d(3, 8, 5)
    mine = QuantConv2d('fp', 'fp', 3, 8, 5)
    mine.weight, mine.bias = nn.Parameter(ref.weight), nn.Parameter(ref.bias)
    a, b = ref(x), mine(x2)
    a.sum().backward()
    b.sum().backward()
    assert torch.equal(a, b) and torch.equal(x.grad, x2.grad)


def test_cpu_tensors_fail_loudly():
    from ml_quant_b200._C import LsqError
    from quant.binary import quantization
    from quant.binary.binary_conv import QuantConv2d
    with pytest.raises(LsqError, match='no CPU fallback'):
        quantization.quantizer_ls_2(torch.randn(2, 3, 4, 4))
    with pytest.raises(LsqError):
        QuantConv2d('ls-1', 'ls-1', 3, 4, 3).eval()(torch.randn(1, 3, 8, 8))


def test_state_dict_layout_matches_reference(golden_nets):
    from ml_quant_b200.nets import QLeNet5, QResNet
    for name, rec in golden_nets.items():
        cls, loss = (QLeNet5, F.nll_loss) if name.startswith('mnist') else (QResNet, F.cross_entropy)
        m = cls(loss_fn=loss, **rec['arch'])
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        ref = {k: tuple(v.shape) for k, v in rec['state'].items()}
        assert list(mine.items()) == list(ref.items()), name
        m.load_state_dict(rec['state'], strict=True)


def test_moving_average_closed_form():
    # reference tests/utils/test_moving_average.py, tests/binary/test_activation_quantization.py (2.0 / 4.0 / 2.2)
    from ml_quant_b200.utils.moving_average import MovingAverage
    ma = MovingAverage(torch.tensor([0.9]))
    ma.train()
    assert ma(torch.tensor([2.0])).item() == 2.0
    assert ma(torch.tensor([4.0])).item() == pytest.approx(0.9 * 2 + 0.1 * 4)
    assert ma.num_batches_tracked.item() == 2
    ma.eval()
    assert ma(torch.tensor([100.0])).item() == pytest.approx(2.2)
    assert set(ma.state_dict()) == {'num_batches_tracked', 'momentum', 'moving_average'}


def test_shard_bounds_cover_the_batch():
    from ml_quant_b200.runtime import shard_bounds
    for total, world in [(4096, 8), (10, 4), (7, 8)]:
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_gather_logits_two_ranks_gloo(tmp_path):
    """N>1 host logic on CPU: contiguous batch shards, one all_gather restores the full batch order."""
    code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from ml_quant_b200.runtime import gather_logits, shard_bounds
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
full = torch.arange(8 * 5, dtype=torch.float32).view(8, 5)
lo, hi = shard_bounds(8, w, r)
out = gather_logits(full[lo:hi] * 1.0, w)
assert torch.equal(out, full), (r, out)
dist.destroy_process_group()
print("rank", r, "ok")
'''
    f = tmp_path / 'w.py'
    f.write_text(code)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29577', str(f), ROOT],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count('ok') == 2


def test_quantlinear_construction_and_fp_equivalence():
    """QuantLinear: same scheme validation as QuantConv2d; 'fp'/'fp' equals F.linear on the CPU."""
    import torch.nn.functional as F
    from quant.binary.binary_conv import QuantLinear
    with pytest.raises(ValueError):
        QuantLinear('ls-3', 'ls-1', 8, 4)
    m = QuantLinear('ls-2', 'ls-1', 128, 64, clamp={'kind': 'symmetric', 'alpha': 2.0})
    assert tuple(m.weight.shape) == (64, 128) and m.w_approximate.v1.shape == (64,)
    assert 'ls-1' in m.quantized_parameters
    fp = QuantLinear('fp', 'fp', 16, 8)
    x = torch.randn(3, 5, 16)
    assert torch.allclose(fp(x), F.linear(x, fp.weight, fp.bias), atol=1e-6)
    with pytest.raises(ValueError):
        fp(torch.randn(3, 15))


def test_pixel_lut_is_the_host_transform():
    """runtime.pixel_lut evaluates torchvision's ToTensor + Normalize once per pixel level with the host's own fp32 ops."""
    import torch
    from ml_quant_b200 import runtime
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    lut = runtime.pixel_lut(mean, std)
    assert lut.shape == (3, 256) and lut.dtype == torch.float32
    img = torch.arange(256, dtype=torch.uint8).view(1, 1, 16, 16).repeat(2, 3, 1, 1)
    want = img.to(torch.float32).div(255).sub_(torch.tensor(mean).view(1, 3, 1, 1)).div_(torch.tensor(std).view(1, 3, 1, 1))
    got = torch.stack([lut[c][img[:, c].long()] for c in range(3)], dim=1)
    assert torch.equal(got, want)
    assert runtime.pixel_lut((0.1307,), (0.3081,)).shape == (1, 256)


def test_uint8_input_needs_a_pixel_table():
    """A model that was not prepared with set_pixel_input must reject uint8 batches loudly (no silent cast)."""
    import pytest
    import torch
    from ml_quant_b200 import runtime

    class M(torch.nn.Module):
        def forward(self, x):
            return x

    with pytest.raises(RuntimeError):
        runtime._lut_on(M(), torch.device('cpu'))
    m = runtime.set_pixel_input(M(), (0.5,), (0.5,))
    assert tuple(runtime._lut_on(m, torch.device('cpu')).shape) == (1, 256)
