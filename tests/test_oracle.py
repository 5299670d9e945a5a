"""Pin the CPU oracle: bit-exact against vectors produced by the unmodified reference
(tests/golden, made by oracle/gen_golden.py), against the reference's known-answer tests
(tests/binary/test_ste.py, test_quantization.py in /root/reference), and -- when the reference
is mounted -- against the live reference on fresh inputs."""
import os
import sys

import pytest
import torch

from oracle import lsq_oracle as O


def test_sign_kat():
    # reference tests/binary/test_ste.py:13-18
    x = torch.tensor([42, -42, 42, 42, 0, -1, 1, -4.2, 4.2])
    assert O.sign_pm1(x).tolist() == [1, -1, 1, 1, 1, -1, 1, -1, 1]


def test_ste_grad_kat():
    # reference tests/binary/test_ste.py:21-36
    x = torch.tensor([42, -42, 0, -1, 1, -0.2, 0.2])
    assert O.ste_grad(x, torch.ones(7)).tolist() == [0, 0, 1, 1, 1, 1, 1]


def test_ternary_all_equal_kat():
    # reference tests/binary/test_quantization.py:95-111
    x = torch.ones(32, 3, 16, 16) * 2
    assert torch.all(O.quant_lsT(x)[1] == 2.0)
    torch.manual_seed(1234)
    x = torch.rand(32, 3, 16, 16)
    x[1] = 2
    x[9] = -3
    xq = O.quant_lsT(x)[1]
    assert torch.all(xq[1] == 2) and torch.all(xq[9] == -3)


def test_functions_bit_exact(golden_functions):
    for rec in golden_functions:
        x = rec['x']
        v1, xq = O.quant_ls1(x)
        assert torch.equal(v1, rec['ls1']['v1'])
        if 'xq' in rec['ls1']:
            assert torch.equal(xq, rec['ls1']['xq'])
        for skip in (1, 3):
            g = rec[f'ls2_s{skip}']
            v1, v2, xq = O.quant_ls2(x, skip=skip, chunk=2)
            assert torch.equal(v1, g['v1']) and torch.equal(v2, g['v2'])
            if 'xq' in g:
                assert torch.equal(xq, g['xq'])
            g = rec[f'lsT_s{skip}']
            v1, xq = O.quant_lsT(x, skip=skip)
            assert torch.equal(v1, g['v1'])
            if 'xq' in g:
                assert torch.equal(xq, g['xq'])
            for tern in (False, True):
                key = f'cand_s{skip}_t{int(tern)}'
                if key not in rec:
                    continue
                a = x.view(x.shape[0], -1)[..., ::skip].abs()
                srt, mask = O.candidate_mask(a, tern)
                assert torch.equal(mask.sum(1), rec[key]['counts'])
                assert torch.equal(torch.masked_select(srt[:, 1:-1], mask), rec[key]['values'])
        for k in (1, 2, 3):
            vs, xq = O.quant_gf(x, k)
            assert torch.equal(torch.stack(vs), rec[f'gf{k}']['vs'])
            if 'xq' in rec[f'gf{k}']:
                assert torch.equal(xq, rec[f'gf{k}']['xq'])


def _w_scales(state, scheme):
    n = {'fp': 0, 'ls-1': 1, 'ls-2': 2, 'ls-T': 1}.get(scheme)
    if n is None:
        n = int(scheme.split('-')[1])
    return [state[f'w_approximate.v{i + 1}'] for i in range(n)]


def test_layers_bit_exact(golden_layers):
    for rec in golden_layers:
        s = rec['spec']
        st = rec['state']
        y = O.quant_conv2d(rec['x'], st['weight'], st.get('bias'), s['x_quant'], s['w_quant'],
                           _w_scales(st, s['w_quant']), s['alpha'], s['stride'], s['padding'])
        assert torch.equal(y, rec['y']), s
        xin = rec['x'] if s['alpha'] is None else O.clamp_symmetric(rec['x'], s['alpha'])
        sc, _ = O.quantize_activation(xin, s['x_quant'])
        for a, b in zip(sc, rec['x_scales']):
            assert torch.equal(a, b)


def test_plane_identity(golden_layers):
    """The integer plane formulation equals the fake-quant conv within 1e-5 of max|y|."""
    for rec in golden_layers:
        s = rec['spec']
        if s['w_quant'] != 'ls-1' or s['x_quant'] == 'fp':
            continue
        st = rec['state']
        xin = rec['x'] if s['alpha'] is None else O.clamp_symmetric(rec['x'], s['alpha'])
        y, ints = O.plane_conv_identity(xin, st['weight'], st.get('bias'), s['x_quant'], rec['x_scales'],
                                        st['w_approximate.v1'], s['stride'], s['padding'])
        err = (y - rec['y']).abs().max() / rec['y'].abs().max()
        assert err < 1e-5, (s, err)
        k = s['cin'] * s['k'] * s['k']
        assert all(int(i.abs().max()) <= k for i in ints)


def test_nets_bit_exact(golden_nets):
    for name, rec in golden_nets.items():
        fwd = O.lenet_forward if name.startswith('mnist') else O.resnet_forward
        y = fwd(rec['state'], rec['arch'], rec['x'])
        assert torch.equal(y, rec['y']), name


def test_moving_average_layers_bit_exact():
    """tests/golden/ma_layers.pt: eval forward of QuantConv2d with STORED activation scales (moving-average modes,
    activation_quantization.py:90-98) -- the oracle's quant_conv2d(x_scales=...) reproduces it bit for bit, and the
    tracked averages follow ema_step over the per-batch mean scales of the two training batches."""
    from tests.conftest import load_golden
    for rec in load_golden('ma_layers.pt'):
        sp, st, x = rec['spec'], rec['state'], rec['x']
        avg = st['x_approximate.moving_avg_module.moving_average']
        scales = [avg[i].expand(x.shape[0]) for i in range(avg.numel())]
        y = O.quant_conv2d(x, st['weight'], st['bias'], sp['x_quant'], 'ls-1', [st['w_approximate.v1']], sp['alpha'], 1, 1,
                           x_scales=scales)
        assert torch.equal(y, rec['y']), sp
        mom = st['x_approximate.moving_avg_module.momentum']
        track = torch.zeros_like(avg)
        for i, xt in enumerate(rec['x_train']):
            vs, _ = O.quantize_activation(O.clamp_symmetric(xt, sp['alpha']), sp['x_quant'], chunk=4)
            track = O.ema_step(track, mom, torch.stack(vs).mean(1), i)
        assert torch.equal(track, avg), (sp, track, avg)


def test_cpu_cumsum_accumulates_in_double():
    """ATen's CPU cumsum of a float tensor accumulates in double and rounds every prefix to fp32 (acc_type<float,
    /*is_cuda=*/false> = double): the reference's `values.cumsum(dim=1)` (optimal.py:56) on the CPU is therefore
    bit-identical to an fp64 prefix sum cast to fp32 -- which is what the mirror's compute_mask and the CUDA
    solver's candidate test ((float) of an fp64 prefix) compute."""
    torch.manual_seed(0)
    for n in (100, 5000, 66902):
        s = torch.sort(torch.randn(4, n).abs(), dim=1).values
        assert torch.equal(s.cumsum(1), s.double().cumsum(1).float())


def test_ema():
    m = torch.tensor([0.9])
    a = O.ema_step(torch.zeros(1), m, torch.tensor([2.0]), 0)
    a = O.ema_step(a, m, torch.tensor([4.0]), 1)
    assert a.item() == pytest.approx(2.2)   # reference tests/binary/test_activation_quantization.py


@pytest.mark.skipif(not os.path.isdir('/root/reference/quant'), reason='reference not mounted')
def test_live_reference_random():
    """Fresh random inputs through the live reference (subprocess: its package is also named quant)."""
    import subprocess
    import tempfile
    code = r'''
import sys, torch
sys.path.insert(0, "/root/reference")
import quant.binary.quantization as q
torch.manual_seed(77)
x = torch.randn(6, 8, 10, 10).clamp(-2.5, 2.5)
out = {"x": x, "ls2": q.quantizer_ls_2(x), "lsT": q.quantizer_ls_ternary(x), "gf3": q.quantizer_gf(x, 3),
       "ls1": q.quantizer_ls_1(x)}
torch.save(out, sys.argv[1])
'''
    with tempfile.NamedTemporaryFile(suffix='.pt') as f:
        subprocess.run([sys.executable, '-c', code, f.name], check=True, cwd='/tmp')
        ref = torch.load(f.name, weights_only=False)
    x = ref['x']
    for a, b in zip(O.quant_ls2(x), ref['ls2']):
        assert torch.equal(a, b)
    for a, b in zip(O.quant_lsT(x), ref['lsT']):
        assert torch.equal(a, b)
    assert torch.equal(O.quant_gf(x, 3)[1], ref['gf3'][1])
    assert torch.equal(O.quant_ls1(x)[1], ref['ls1'][1])


@pytest.mark.skipif(not os.path.isdir('/root/reference/quant'), reason='reference not mounted')
def test_live_reference_stored_scales_and_moving_average_conv():
    """The stored-scale (moving-average) eval path of the live reference -- QuantConv2d built with
    moving_average_mode='eval_only', one train-mode step to track the scales, then eval
    (activation_quantization.py:68-102, binary_conv.py:161-173) -- against the oracle's
    quantize_activation_stored / quant_conv2d(x_scales=...), bit for bit, for every activation scheme."""
    import subprocess
    import tempfile
    code = r'''
import sys, torch
sys.path.insert(0, "/root/reference")
from quant.binary.binary_conv import QuantConv2d
out = {}
for scheme in ("ls-1", "ls-2", "ls-T", "gf-2"):
    torch.manual_seed(5)
    m = QuantConv2d(scheme, "ls-1", 8, 16, 3, clamp={"kind": "symmetric", "alpha": 2.0},
                    moving_average_mode="eval_only", moving_average_momentum=0.9, padding=1)
    m.train()
    with torch.no_grad():
        m(torch.randn(4, 8, 9, 9))
        m(torch.randn(4, 8, 9, 9))
    m.eval()
    x = torch.randn(3, 8, 9, 9)
    with torch.no_grad():
        y = m(x)
    out[scheme] = {"state": m.state_dict(), "x": x, "y": y}
torch.save(out, sys.argv[1])
'''
    with tempfile.NamedTemporaryFile(suffix='.pt') as f:
        subprocess.run([sys.executable, '-c', code, f.name], check=True, cwd='/tmp')
        ref = torch.load(f.name, weights_only=False)
    for scheme, rec in ref.items():
        st, x = rec['state'], rec['x']
        avg = st['x_approximate.moving_avg_module.moving_average']
        scales = [avg[i].expand(x.shape[0]) for i in range(avg.numel())]
        y = O.quant_conv2d(x, st['weight'], st['bias'], scheme, 'ls-1', [st['w_approximate.v1']], 2.0, 1, 1,
                           x_scales=scales)
        assert torch.equal(y, rec['y']), scheme


def test_sort_free_solver_model_matches_oracle():
    """tests/solver_model.py (the serial numpy specification of the CUDA solver's window / flag / collect /
    evaluate logic) picks a candidate of the oracle's own candidate set with no worse exact cost -- on rows long
    enough to take the windowed path, for both quantizers.  CPU only: the algorithm is validated before any GPU time."""
    import numpy as np
    from tests import solver_model as M
    torch.manual_seed(21)
    x = torch.randn(3, 24000).clamp(-3, 3)
    x[2] = x[2].abs() + 0.5                     # a row whose minimum exceeds half its mean (ternary edge case)
    for tern in (False, True):
        v, infos = M.solve(x.numpy(), tern, skip=1)
        v_ref = O.solve_v1(x, tern, 1, chunk=1).view(-1)
        a = x.abs()
        srt, mask = O.candidate_mask(a, tern)
        for r in range(3):
            cands = torch.masked_select(srt[r, 1:-1], mask[r])
            edge = tern and bool(a[r].min() > 0.5 * a[r].mean())
            assert bool((cands == float(v[r])).any()) or edge or cands.numel() == 0, (tern, r, float(v[r]))
        c_my = O.exact_cost(x, torch.from_numpy(np.asarray(v, dtype=np.float32)), tern, 1)
        c_or = O.exact_cost(x, v_ref, tern, 1)
        assert bool((c_my <= c_or * (1 + 1e-5) + 1e-12).all())
