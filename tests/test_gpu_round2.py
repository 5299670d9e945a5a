"""GPU parity tests added in round 2 (VERDICT r1 "what's missing" 1-4, 6 and the advisor's findings), all through
the C ABI: stored-scale (moving-average) packed route, compute_mask / cost_function, injected-scale end-to-end
logits of the headline network, the CIFAR configuration at full width, re-entrancy, large feature maps, many rows.
"""
import threading

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import lsq_oracle as O
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def runtime_strict():
    from ml_quant_b200 import runtime
    runtime.strict_fp32()


# ---------------------------------------------------------------------------------------------------------------
# QuantConv2d with moving_average_mode != 'off' (reference activation_quantization.py:68-102)
# ---------------------------------------------------------------------------------------------------------------
def test_moving_average_layers_against_golden():
    """tests/golden/ma_layers.pt (made by oracle/gen_golden.py from the unmodified reference): QuantConv2d with
    'eval_only' / 'train_and_eval' activation quantizers for ls-1 / ls-2 / ls-T / gf-2.
      (a) reference state loaded, eval forward through forward() AND forward_fused(): the packed route with the
          STORED scales (encode only, no solve) reproduces the reference output to 1e-5 of max|y|;
      (b) the two train-mode steps re-run here track the same moving average (1e-5: the ls-2 / ls-T batch solve
          picks a v1 under the staged solver contract, the tracked value is a mean over 4 samples and 2 steps)."""
    runtime_strict()
    from quant.binary.binary_conv import QuantConv2d
    from ml_quant_b200 import ops
    recs = load_golden('ma_layers.pt')
    assert len(recs) == 8
    for rec in recs:
        sp = rec['spec']
        m = QuantConv2d(sp['x_quant'], 'ls-1', 64, 64, 3, {'kind': 'symmetric', 'alpha': sp['alpha']}, sp['mode'],
                        sp['momentum'], padding=1)
        m.load_state_dict(rec['state'])
        m = m.to(DEV).eval()
        x = rec['x'].to(DEV)
        ops.reset_counters()
        with torch.no_grad():
            y = m(x)
            yf = m.forward_fused(x)
        # the stored-scale branch ran: encoder + packed convolution, and NO solver launch
        assert ops.LAUNCHES.get('encode_act', 0) + ops.LAUNCHES.get('quant_act', 0) >= 2, ops.LAUNCHES
        assert ops.LAUNCHES.get('solve_v1', 0) == 0 and ops.LAUNCHES.get('row_absmean', 0) == 0, ops.LAUNCHES
        assert ops.LAUNCHES.get('bconv_tc', 0) == 2, ops.LAUNCHES
        want = rec['y']
        for out in (y, yf):
            err = float((out.cpu() - want).abs().max() / want.abs().max())
            assert err < 1e-5, (sp, err)
        # oracle with the stored scales agrees with the golden output bit for bit (pins the oracle function)
        st = rec['state']
        avg = st['x_approximate.moving_avg_module.moving_average']
        scales = [avg[i].expand(x.shape[0]) for i in range(avg.numel())]
        y_or = O.quant_conv2d(rec['x'], st['weight'], st['bias'], sp['x_quant'], 'ls-1', [st['w_approximate.v1']],
                              sp['alpha'], 1, 1, x_scales=scales)
        assert torch.equal(y_or, want)
        # (b) tracking: fresh module with the reference's weights, the reference's two training batches
        m2 = QuantConv2d(sp['x_quant'], 'ls-1', 64, 64, 3, {'kind': 'symmetric', 'alpha': sp['alpha']}, sp['mode'],
                         sp['momentum'], padding=1)
        with torch.no_grad():
            m2.weight.copy_(st['weight'])
            m2.bias.copy_(st['bias'])
        m2 = m2.to(DEV).train()
        with torch.no_grad():
            for xt in rec['x_train']:
                yt = m2(xt.to(DEV))
        got_avg = m2.x_approximate.moving_avg_module.moving_average.cpu()
        assert torch.allclose(got_avg, avg, rtol=2e-5, atol=0), (sp, got_avg, avg)
        assert int(m2.x_approximate.moving_avg_module.num_batches_tracked) == 2
        assert torch.allclose(m2.w_approximate.v1.cpu(), st['w_approximate.v1'], rtol=1e-6, atol=0)
        if sp['mode'] == 'eval_only' and sp['x_quant'] in ('ls-1', 'gf-2'):
            # train-mode output uses the batch scales; ls-1 / gf have no ill-posed pick: compare directly
            err = float((yt.cpu() - rec['y_train_last']).abs().max() / rec['y_train_last'].abs().max())
            assert err < 1e-5, (sp, err)


def test_activation_quantizer_lst_moving_average_kat():
    """Reference KAT tests/binary/test_activation_quantization.py (ternary, 1.0 -> 1.1 with momentum 0.9)."""
    from quant.binary import quantization
    from quant.binary.activation_quantization import ActivationQuantizerLST
    torch.manual_seed(1234)
    x = torch.ones(32, 16, 3, 3, device=DEV) * 2
    x2 = torch.rand(32, 16, 3, 3, device=DEV)
    x3 = torch.ones(32, 16, 3, 3, device=DEV) * 4
    for mode in ('eval_only', 'train_and_eval'):
        q = ActivationQuantizerLST(mode, 0.9).to(DEV)
        q.train()
        out = q(x)                       # all-equal rows: v1 = mean / 2 = 1.0 (optimal.py:86-118), x_q = 2 v1
        assert torch.all(out == 2.0)
        assert torch.allclose(q.moving_avg_module.moving_average.cpu(), torch.tensor([1.0]))
        out = q(x3)                      # tracked v1 -> 1 * 0.9 + 2 * 0.1 = 1.1
        assert torch.allclose(q.moving_avg_module.moving_average.cpu(), torch.tensor([1.1]))
        if mode == 'eval_only':
            assert torch.all(out == 4.0)
        else:
            assert torch.equal(out, quantization.quantizer_ls_ternary(x3, torch.tensor([1.1] * 32, device=DEV))[1])
        q.eval()
        want = quantization.quantizer_ls_ternary(x2, torch.tensor([1.1] * 32, device=DEV))[1]
        assert torch.equal(q(x2), want)


def test_moving_average_reference_closed_form_on_cuda():
    """The CUDA half of the reference's tests/utils/test_moving_average.py::test_moving_average_{train_and_eval,
    eval_only} (:41-122; their loops start with an explicit CPU device, which this implementation rejects by design)."""
    from quant.binary.activation_quantization import ActivationQuantizerLS1
    from quant.binary.quantization import quantizer_ls_1

    def closed_form(i, alpha):
        return (1 - alpha) * sum(alpha ** (i - j) * j for j in range(1, i + 1))
    alpha, device = 0.9, torch.device(DEV)
    for mode in ('train_and_eval', 'eval_only'):
        q = ActivationQuantizerLS1(mode, alpha).to(device)
        q.train()
        for i in range(10):
            x = i * torch.ones(8, 1, 20, 20, requires_grad=True, device=device)
            x_q = q(x)
            x_q.sum().backward()
            ma = q.moving_avg_module.moving_average
            want = torch.tensor(closed_form(i, alpha), device=device).expand_as(ma)
            assert torch.allclose(want, ma)
            if mode == 'train_and_eval':
                _, expected = quantizer_ls_1(x, torch.tensor([closed_form(i, alpha)], device=device).expand(8))
                assert torch.allclose(expected, x_q)
            else:
                assert torch.allclose(x, x_q)
        q.eval()
        for i in range(5):
            x = i * torch.ones(8, 1, 20, 20, requires_grad=True, device=device)
            q(x).sum().backward()
            ma = q.moving_avg_module.moving_average
            assert torch.allclose(torch.tensor(closed_form(9, alpha), device=device).expand_as(ma), ma)


# ---------------------------------------------------------------------------------------------------------------
# compute_mask / cost_function (reference optimal.py:16-83) against the golden candidate sets
# ---------------------------------------------------------------------------------------------------------------
def test_compute_mask_and_cost_function_against_golden(golden_functions):
    """compute_mask on the GPU returns the reference's candidate set -- per-row counts and the selected values bit
    for bit (tests/golden/functions.pt: cand_*).  The prefix sums are accumulated in fp64 and rounded to fp32 per
    position, which IS what the reference computes on the CPU: ATen's CPU cumsum accumulates a float tensor in
    double (acc_type<float, false>) and rounds each prefix (checked bit-exactly in tests/test_oracle.py).
    cost_function agrees with the reference formula evaluated by the oracle to 1e-6."""
    from quant.binary import optimal
    for rec in golden_functions:
        rows = rec['x'].reshape(rec['x'].shape[0], -1)
        for skip in (1, 3):
            a = rows[..., ::skip].abs()
            if a.shape[1] < 3:
                continue
            for tern in (False, True):
                gold = rec[f'cand_s{skip}_t{int(tern)}']
                mask, vals = optimal.compute_mask(a.to(DEV), tern)
                assert mask.dtype == torch.bool and tuple(mask.shape) == (a.shape[0], a.shape[1] - 2)
                assert torch.equal(mask.sum(1).cpu(), gold['counts']), (skip, tern)
                assert torch.equal(vals.cpu(), gold['values']), (skip, tern)
                # cost of every candidate: reference formula on the CPU (oracle) vs the device expression
                if vals.numel() == 0:
                    continue
                ncand = int(gold['counts'].max())
                table = torch.zeros(a.shape[0], max(ncand, 1))
                off = 0
                for r, c in enumerate(gold['counts'].tolist()):
                    table[r, :c] = gold['values'][off:off + c]
                    off += c
                want = O.candidate_cost(a, table, tern)
                got = optimal.cost_function(a.to(DEV), table.to(DEV), tern).cpu()
                assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (skip, tern, float((got - want).abs().max()))


# ---------------------------------------------------------------------------------------------------------------
# end-to-end logits with the oracle's v1 injected per layer
# ---------------------------------------------------------------------------------------------------------------
def _inject_v1(model, per_layer_v1):
    """Make every ls-2 QuantConv2d of ``model`` quantize with a GIVEN v1 (the public ``v1=`` argument of
    quantizer_ls_2, quantization.py:61-62) instead of solving; v2 and the planes are computed by the kernels."""
    from ml_quant_b200 import ops, runtime
    used = []
    for i, m in enumerate(runtime.quant_layers(model)):
        def q_given(x, g, prologue=None, _m=m, _i=i):
            n = x.shape[0]
            tab = torch.empty(2, n, dtype=torch.float32, device=x.device)
            tab[0].copy_(per_layer_v1[_i].to(x.device))
            planes, _ = ops.encode_act(x, g, tab[:1], 2, _m.clamp_alpha, True, None, prologue, next_scale_out=tab[1])
            used.append(tab.detach().cpu())
            return planes, tab
        m.quantize_input = q_given
    return used


@pytest.mark.parametrize('cfg,hw', [('imagenet_resnet18_ls1w_ls2a', 224), ('cifar100_resnet18_ls1w_ls2a', 32)])
def test_logits_with_injected_oracle_scales(cfg, hw):
    """SURVEY H1(d) experiment: with the oracle's v1 of every layer injected, the ill-posed arg-min is
    out of the loop.  The whole forward -- stem, 16 x (BatchNorm, clamp, encode, v2, binary convolution, ReLU,
    shortcuts), classifier -- is then compared with the oracle's logits (plain and fused-for-inference graph).  See the
    comment at the assertions for what was measured and why the bar is the coarse one."""
    runtime_strict()
    from ml_quant_b200 import configs, runtime
    model = runtime.build_model(cfg, torch.device(DEV))
    runtime.calibrate(model, (3, hw, hw), batches=1, batch=8)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(2, 3, hw, hw, generator=g)
    rec = []
    y_ref = O.resnet_forward(sd, configs.arch(cfg), x, record=rec)
    assert len(rec) == 16
    used = _inject_v1(model, [r[0] for r in rec])
    with torch.no_grad():
        y = model(x.to(DEV)).cpu()
    assert len(used) == 16
    v2_err = 0.0
    for i, (tab, r) in enumerate(zip(used, rec)):
        assert torch.equal(tab[0], r[0])
        v2_err = max(v2_err, float(((tab[1] - r[1]).abs() / r[1].abs()).max()))
    err = float((y - y_ref).abs().max() / y_ref.abs().max())
    used.clear()
    fused = runtime.optimize_for_inference(model)
    with torch.no_grad():
        yf = fused(x.to(DEV)).cpu()
    errf = float((yf - y_ref).abs().max() / y_ref.abs().max())
    print(f'[injected {cfg}] logits err plain {err:.3e} fused {errf:.3e}  max v2 err over layers {v2_err:.3e}')
    # Measured (B200, round 2): logits 4.5e-2 / 8.1e-2 (ImageNet / CIFAR), v2 up to 2e-3 / 4e-3 -- the same size as WITHOUT
    # injection.  Fixing v1 does not make the network well conditioned: every layer thresholds ~10^5..10^6 activations
    # per sample at 0 and at +-v1, a few of them lie within fp32 rounding of a threshold, each flipped bit moves 9 x Cout
    # outputs by 2 v w, and 16 random-init layers amplify that (the oracle's own logits move 2-3 % under a 1e-7 input
    # perturbation, scripts/dev/oracle_sensitivity.py).  SURVEY H1(d)'s 1e-5 end-to-end bar is therefore not attainable
    # by any implementation with a different summation order; the strict statements are the teacher-forced per-layer
    # tests (1e-5 / 1e-6 / bit-exact).  Here: the coarse end-to-end bar, equal for both graphs.
    assert v2_err < 2e-2, v2_err
    assert err < 0.15 and errf < 0.15, (err, errf)


def test_cifar_full_width_layer_by_layer():
    """BASELINE config 2 (CIFAR-100 ResNet-18, ls-1 weights / ls-2 activations, clamp 2, full width) teacher-forced
    like the headline network: every layer, fed with the GPU's own input, meets the solver contract, reproduces the
    oracle's v2 for that v1 to 1e-6 and the oracle's convolution output for those scales to 1e-5 of max|y|."""
    runtime_strict()
    from ml_quant_b200 import runtime
    from tests.test_gpu_quantizers import _solver_contract
    model = runtime.build_model('cifar100_resnet18_ls1w_ls2a', torch.device(DEV))
    runtime.calibrate(model, (3, 32, 32), batches=1, batch=16)
    layers = runtime.quant_layers(model)
    assert len(layers) == 16
    rec = {}
    for i, m in enumerate(layers):
        orig = m.quantize_input

        def wrapped(x, g, prologue=None, _i=i, _orig=orig):
            planes, table = _orig(x, g, prologue)
            rec[_i] = [x.detach().cpu(), table.detach().cpu(), None]
            return planes, table
        m.quantize_input = wrapped
        m.register_forward_hook(lambda mod, inp, out, _i=i: rec[_i].__setitem__(2, out.detach().cpu()))
    g = torch.Generator().manual_seed(79)
    x = torch.randn(4, 3, 32, 32, generator=g)
    with torch.no_grad():
        model(x.to(DEV))
    for i, m in enumerate(layers):
        xi, table, y = rec[i]
        xin = xi.clamp(-m.clamp_alpha, m.clamp_alpha)
        rows = xin.reshape(4, -1)
        v1, v2 = table[0], table[1]
        _solver_contract(rows, v1, O.solve_v1(rows, False, 3, chunk=1).view(-1), False, 3)
        b1 = torch.where(xin >= 0, 1.0, -1.0)
        v2_ref = (xin - v1.view(-1, 1, 1, 1) * b1).abs().mean(dim=(1, 2, 3))
        assert torch.allclose(v2, v2_ref, rtol=1e-6, atol=0), (i, v2, v2_ref)
        y_ref, _ = O.plane_conv_identity(xin, m.weight.detach().cpu(), m.bias.detach().cpu(), 'ls-2', [v1, v2],
                                         m.w_approximate.v1.detach().cpu(), m.stride[0], m.padding[0])
        err = float((y - y_ref).abs().max() / y_ref.abs().max())
        assert err < 1e-5, (i, err)


# ---------------------------------------------------------------------------------------------------------------
# re-entrancy: replicas / threads / streams sharing one module
# ---------------------------------------------------------------------------------------------------------------
def test_two_threads_two_streams_share_a_module():
    """nn.DataParallel replicas are shallow copies that share the module's caches (packed weights, plane scratch)
    and run concurrently, one host thread each.  Here: two threads, each on its own stream, push different batches
    through the SAME QuantConv2d 20 times; every result must equal the serial result bit for bit."""
    from quant.binary.binary_conv import QuantConv2d
    torch.manual_seed(3)
    m = QuantConv2d('ls-2', 'ls-1', 64, 64, 3, {'kind': 'symmetric', 'alpha': 3.0}, padding=1).to(DEV)
    with torch.no_grad():
        m.train(); m(torch.randn(2, 64, 28, 28, device=DEV)); m.eval()
    xs = [torch.randn(8, 64, 28, 28, device=DEV) * (1 + 0.1 * i) for i in range(2)]
    with torch.no_grad():
        want = [m(x).clone() for x in xs]
    torch.cuda.synchronize()
    bad = []

    def worker(i):
        st = torch.cuda.Stream(device=DEV)
        with torch.cuda.stream(st), torch.no_grad():
            for _ in range(20):
                y = m(xs[i])
                if not torch.equal(y, want[i]):
                    bad.append(i)
        st.synchronize()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not bad, bad


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='requires >= 2 GPUs')
def test_data_parallel_two_gpus():
    """Counterpart of the reference's only multi-GPU test (tests/utils/test_moving_average.py:125-166): a
    QuantConv2d with an eval_only moving average under nn.DataParallel on 2 GPUs -- statistics tracked on replica
    0, eval output equal to the single-GPU module's."""
    from quant.binary.binary_conv import QuantConv2d
    torch.manual_seed(4)
    m = QuantConv2d('ls-1', 'ls-1', 64, 64, 3, {'kind': 'symmetric', 'alpha': 2.0}, 'eval_only', 0.9, padding=1).to(DEV)
    dp = nn.DataParallel(m, device_ids=[0, 1])
    dp.train()
    with torch.no_grad():
        for i in range(3):
            dp(torch.randn(8, 64, 14, 14, device=DEV))
    assert int(m.x_approximate.moving_avg_module.num_batches_tracked) == 3
    dp.eval()
    x = torch.randn(8, 64, 14, 14, device=DEV)
    with torch.no_grad():
        y_dp, y_single = dp(x), m(x)
    assert torch.equal(y_dp, y_single)


# ---------------------------------------------------------------------------------------------------------------
# advisor findings
# ---------------------------------------------------------------------------------------------------------------
def test_fused_prologue_on_large_feature_map():
    """The fused BatchNorm prologue maps an element to its channel by a reciprocal multiplication that is exact only
    while row_length * inner < 2^40; 64 channels of 500 x 500 (inner = 250 000, not a power of two) is beyond that
    and takes the exact division.  Scales with the fused prologue must equal the scales of the materialised affine
    map (same fmaf) bit for bit -- row means at that size, the solvers (which share prologue_channel) at a
    non-power-of-two inner size they solve quickly."""
    from ml_quant_b200 import ops
    torch.manual_seed(6)
    for c, inner, solve in ((64, 500 * 500, False), (5, 70001, True)):
        x = torch.randn(2, c * inner, device=DEV)
        a = (torch.rand(c, device=DEV) + 0.5)
        b = torch.randn(c, device=DEV)
        xb = torch.addcmul(b.repeat_interleave(inner).unsqueeze(0), x, a.repeat_interleave(inner).unsqueeze(0))
        xb2 = (x.double() * a.double().repeat_interleave(inner) + b.double().repeat_interleave(inner)).float()   # one rounding
        assert torch.equal(xb, xb2)          # addcmul on the GPU is a fused multiply-add, like the kernels' fmaf
        pro = (a, b, inner)
        assert torch.equal(ops.row_absmean(x, [], 2.0, pro), ops.row_absmean(xb, [], 2.0))
        if solve:
            assert torch.equal(ops.solve_v1(x, False, 3, 2.0, prologue=pro), ops.solve_v1(xb, False, 3, 2.0))
            assert torch.equal(ops.solve_v1(x, True, 3, 2.0, prologue=pro), ops.solve_v1(xb, True, 3, 2.0))


def test_quantlinear_many_rows_and_solver_contract():
    """QuantLinear flattens every leading dimension into rows: more than 65 535 rows (the gridDim.y limit of the row
    kernels) must work on both routes; the ls-2 / ls-T activation scales obey the solver contract and, with OUR
    scales, the output equals the oracle's composition to 1e-5."""
    from quant.binary.binary_conv import QuantLinear
    from tests.test_gpu_quantizers import _solver_contract
    torch.manual_seed(13)
    for xs, fin, fout, alpha in [('ls-2', 512, 256, 3.0), ('ls-T', 192, 64, 3.0), ('ls-2', 100, 10, None)]:
        clamp = None if alpha is None else {'kind': 'symmetric', 'alpha': alpha}
        m = QuantLinear(xs, 'ls-1', fin, fout, clamp=clamp).to(DEV)
        x = torch.randn(33, fin) * 1.5
        with torch.no_grad():
            m.train(); m(x.to(DEV)); m.eval()
            y = m(x.to(DEV)).cpu()
        xin = x if alpha is None else x.clamp(-alpha, alpha)
        tern = xs == 'ls-T'
        from ml_quant_b200 import ops
        v1 = ops.solve_v1(xin.to(DEV), tern, 3).cpu()
        _solver_contract(xin, v1, O.solve_v1(xin, tern, 3, chunk=8).view(-1), tern, 3)
        x4 = xin.view(33, fin, 1, 1)
        xq = (O.quant_lsT(x4, v1)[1] if tern else O.quant_ls2(x4, v1)[2]).view(33, fin)
        w = m.weight.detach().cpu()
        wq = w.abs().mean(1, keepdim=True) * torch.where(w >= 0, 1.0, -1.0)
        want = F.linear(xq, wq, m.bias.detach().cpu())
        err = float((y - want).abs().max() / want.abs().max())
        assert err < 1e-5, (xs, fin, fout, err)
    big = QuantLinear('ls-1', 'ls-1', 64, 64, clamp={'kind': 'symmetric', 'alpha': 2.0}).to(DEV)
    x = torch.randn(70000, 64, device=DEV)
    with torch.no_grad():
        big.train(); big(x[:16]); big.eval()
        y = big(x)
        y_head = big(x[:100])
        y_tail = big(x[-100:])
    assert y.shape == (70000, 64)
    assert torch.equal(y[:100], y_head) and torch.equal(y[-100:], y_tail)


def test_eval_with_grad_enabled_keeps_gradients_and_warns():
    """optimize_for_inference'd blocks in eval mode WITH autograd on (frozen-BN fine-tuning, saliency) must build the
    reference's graph: the input and the shortcut convolution receive gradients; a RuntimeWarning names the slow route."""
    import warnings
    from ml_quant_b200 import runtime
    from ml_quant_b200.binary.binary_conv import QuantConv2d
    model = runtime.build_model('cifar100_resnet18_ls1w_ls2a', torch.device(DEV))
    runtime.calibrate(model, (3, 32, 32), batches=1, batch=8)
    runtime.optimize_for_inference(model)
    x = torch.randn(2, 3, 32, 32, device=DEV, requires_grad=True)
    QuantConv2d._warned_grad_eval = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        model(x).sum().backward()
    assert any('torch.no_grad' in str(i.message) for i in w)
    assert x.grad is not None and float(x.grad.abs().sum()) > 0
    sc = [m for m in model.modules() if hasattr(m, 'shortcut') and len(m.shortcut) == 2]
    assert sc and all(b.shortcut[0].weight.grad is not None and float(b.shortcut[0].weight.grad.abs().sum()) > 0 for b in sc)


# ---------------------------------------------------------------------------------------------------------------
# fused activation quantizer (csrc/lsq_qact.cu): one HBM read per QuantConv2d input
# ---------------------------------------------------------------------------------------------------------------
def _fused_vs_generic(x, g, tern, alpha, pro):
    """lsq_quantize_act against the generic kernels it replaces: v1 under the solver contract (and normally the same
    pick), planes bit-exact given v1, v2 to 1e-6."""
    from ml_quant_b200 import ops
    from tests.test_gpu_quantizers import _solver_contract
    n = x.shape[0]
    nwords = ops._C.lib().lsq_act_planes_bytes(ops.C.byref(g), 2) // 4
    # zero-filled plane buffers: the padding positions of the raster are written by neither kernel
    planes, tab, dg = ops.quantize_act(x, g, tern, alpha, 3, torch.zeros(nwords, dtype=torch.int32, device=x.device), pro, diag=True)
    torch.cuda.synchronize()
    # reference composition on the CPU for the contract: clamp(bn(x)) rows
    xin = x.detach().cpu()
    if pro is not None:
        a, b, inner = pro
        xin = torch.addcmul(b.cpu().view(1, -1, 1, 1), xin, a.cpu().view(1, -1, 1, 1))       # one rounding, like fmaf
        xin = (x.detach().cpu().double() * a.cpu().double().view(1, -1, 1, 1) + b.cpu().double().view(1, -1, 1, 1)).float()
    xin = xin.clamp(-alpha, alpha)
    rows = xin.reshape(n, -1)
    v1 = tab[0].cpu()
    _solver_contract(rows, v1, O.solve_v1(rows, tern, 3, chunk=1).view(-1), tern, 3)
    # generic encoder with the fused kernel's v1: identical planes, v2 to rounding
    planes2, v2 = ops.encode_act(x, g, tab[:1].clone(), 2, alpha, not tern, torch.zeros(nwords, dtype=torch.int32, device=x.device), pro)
    assert torch.equal(planes[:nwords], planes2[:nwords])
    if tern:
        assert torch.equal(tab[1], tab[0])
    else:
        assert torch.allclose(tab[1], v2, rtol=1e-6, atol=0), (tab[1], v2)
        b1 = torch.where(xin >= 0, 1.0, -1.0)
        v2_ref = (xin - v1.view(-1, 1, 1, 1) * b1).abs().mean(dim=(1, 2, 3))
        assert torch.allclose(tab[1].cpu(), v2_ref, rtol=1e-6, atol=0)
    return tab, dg.cpu()


@pytest.mark.parametrize('tern', [False, True])
def test_fused_activation_quantizer_layer_shapes(tern):
    """Every QuantConv2d input shape of the ImageNet and CIFAR networks (cluster sizes 4, 2 and 1; 16-byte and scalar
    encoder paths; stride-1 and stride-2 rasters; with and without the fused BatchNorm prologue): all rows solved by
    the fused kernel itself (status 0)."""
    from ml_quant_b200 import ops
    torch.manual_seed(31)
    shapes = [(3, 64, 56, 56, 1, 3.0), (3, 64, 56, 56, 2, 3.0), (3, 128, 28, 28, 1, 3.0), (4, 256, 14, 14, 1, 2.0),
              (5, 512, 7, 7, 1, 3.0), (4, 64, 32, 32, 1, 2.0), (4, 128, 16, 16, 2, 2.0), (6, 512, 4, 4, 1, 2.0),
              (2, 96, 13, 11, 1, 2.5)]
    for i, (n, c, h, w, st, alpha) in enumerate(shapes):
        x = torch.randn(n, c, h, w, device=DEV) * (0.6 + 0.3 * i)
        g = ops.act_geometry(n, c, h, w, 3, 3, st, 1)
        for use_pro in (False, True):
            pro = None
            if use_pro:
                pro = (torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.3, h * w)
            tab, dg = _fused_vs_generic(x, g, tern, alpha, pro)
            assert int(dg[:, 0].abs().sum()) == 0, (n, c, h, w, st, use_pro, dg)
            assert int(dg[:, 3].min()) >= 1          # candidates found on every row
    # the same pick as the generic solver on a bare tensor (both take the first minimum of the same exact costs)
    x = torch.randn(8, 64, 56, 56, device=DEV)
    g = ops.act_geometry(8, 64, 56, 56, 3, 3, 1, 1)
    _, tab = ops.quantize_act(x, g, tern, 3.0)
    v_old = ops.solve_v1(x.reshape(8, -1), tern, 3, 3.0)
    assert int((tab[0] == v_old).sum()) >= 6, (tab[0], v_old)


def test_fused_activation_quantizer_hands_odd_rows_to_the_generic_kernels():
    """Rows the fused kernel cannot decide -- all elements equal, all zero, everything far below the clamp bound, a few
    distinct values -- are marked (status != 0) and redone by the generic kernels inside the same call: results are
    bit-identical to lsq_solve_v1 + lsq_encode_act; ordinary rows of the same batch keep status 0."""
    from ml_quant_b200 import ops
    torch.manual_seed(32)
    # (the second shape's rows are longer than the generic solver's in-shared-memory layout holds)
    for n, c, h, w in ((8, 64, 28, 28), (8, 64, 56, 56), (8, 96, 14, 14), (8, 32, 10, 10)):
        x = torch.randn(n, c, h, w, device=DEV)
        x[1] = 0.75
        x[2] = 0.0
        x[3] *= 1e-6
        x[4] = torch.randint(0, 3, (c, h, w), device=DEV).float() - 1.0
        x[5] = x[5].abs() + 0.5
        g = ops.act_geometry(n, c, h, w, 3, 3, 1, 1)
        nwords = ops._C.lib().lsq_act_planes_bytes(ops.C.byref(g), 2) // 4
        for tern in (False, True):
            planes, tab, dg = ops.quantize_act(x, g, tern, 2.0, 3, torch.zeros(nwords, dtype=torch.int32, device=DEV), None, diag=True)
            v1 = ops.solve_v1(x.reshape(n, -1), tern, 3, 2.0)
            planes2, v2 = ops.encode_act(x, g, [v1], 2, 2.0, not tern)
            st = dg[:, 0].cpu()
            assert int(st[0]) == 0 and int(st[6]) == 0 and int(st[7]) == 0, st
            assert int((st != 0).sum()) >= 3, st
            odd = st != 0
            assert torch.equal(tab[0].cpu()[odd], v1.cpu()[odd])
            if not tern:
                assert torch.equal(tab[1].cpu()[odd], v2.cpu()[odd])
            # planes of the whole batch equal the generic encoder's for the scales in the table
            planes3, _ = ops.encode_act(x, g, tab[:1].clone(), 2, 2.0, False, torch.zeros(nwords, dtype=torch.int32, device=DEV))
            assert torch.equal(planes[:nwords], planes3[:nwords])
    # shapes outside the fused kernel's domain take the generic kernels for every row: no clamp, short rows
    x = torch.randn(4, 64, 8, 8, device=DEV)
    g = ops.act_geometry(4, 64, 8, 8, 3, 3, 1, 1)
    for alpha in (None, 2.0):
        planes, tab = ops.quantize_act(x[:, :, :4, :4].contiguous() if alpha else x, ops.act_geometry(4, 64, 4 if alpha else 8, 4 if alpha else 8, 3, 3, 1, 1), False, alpha)
        xx = x[:, :, :4, :4].contiguous() if alpha else x
        v1 = ops.solve_v1(xx.reshape(4, -1), False, 3, alpha)
        assert torch.equal(tab[0], v1)


# ---------------------------------------------------------------------------------------------------------------
# multi-plane weight schemes on the packed route (reference weight_quantization.py:37-109)
# ---------------------------------------------------------------------------------------------------------------
def test_multi_plane_weight_schemes_take_the_packed_route(golden_layers):
    """w_quant in {ls-2, ls-T, gf-2, gf-3} with binary activations: one tensor-core binary convolution per weight sign
    plane (no dense fake-quant weights, no F.conv2d), output equal to the oracle's composition -- scales solved by the
    oracle's own train-mode call, so only exact integer accumulators and fp32 epilogues differ: <= 1e-5 of max|y|;
    plain forward and the fused BatchNorm / ReLU / residual forms."""
    runtime_strict()
    from quant.binary.binary_conv import QuantConv2d
    from ml_quant_b200 import ops
    torch.manual_seed(41)
    for xs, ws, cin, cout, st in [('ls-2', 'ls-2', 64, 64, 1), ('ls-1', 'ls-T', 64, 128, 2), ('ls-2', 'gf-2', 128, 128, 1),
                                  ('ls-T', 'gf-3', 64, 64, 1)]:
        m = QuantConv2d(xs, ws, cin, cout, 3, {'kind': 'symmetric', 'alpha': 2.0}, stride=st, padding=1)
        x = torch.randn(3, cin, 14, 14) * 1.2
        w = m.weight.detach()
        wsc, wq = O.quantize_weight(w, ws)                       # the reference's train-mode solve of the weight scales
        names = [f'v{i + 1}' for i in range(len(wsc))]
        with torch.no_grad():
            for nme, v in zip(names, wsc):
                getattr(m.w_approximate, nme).copy_(v)
        m = m.to(DEV).eval()
        xin = x.clamp(-2.0, 2.0)
        ops.reset_counters()
        with torch.no_grad():
            y = m(x.to(DEV))
        nplanes_w = {'ls-2': 2, 'ls-T': 2, 'gf-2': 2, 'gf-3': 3}[ws]
        assert ops.LAUNCHES.get('bconv_tc', 0) == nplanes_w and 'fakequant' not in ops.LAUNCHES, ops.LAUNCHES
        # the oracle's convolution with OUR activation scales (the ls-2 / ls-T pick is under the solver contract)
        g = ops.act_geometry(3, cin, 14, 14, 3, 3, st, 1)
        if xs in ('ls-2', 'ls-T'):
            _, tab = ops.quantize_act(x.to(DEV), g, xs == 'ls-T', 2.0)
            xsc = [tab[0].cpu(), tab[1].cpu()] if xs == 'ls-2' else [tab[0].cpu()]
        else:
            xsc = [ops.row_absmean(xin.reshape(3, -1).to(DEV)).cpu()]
        xq = O.quantize_activation_stored(xin, xs, xsc)
        want = F.conv2d(xq.double(), wq.double(), m.bias.detach().cpu().double(), st, 1).float()
        err = float((y.cpu() - want).abs().max() / want.abs().max())
        assert err < 1e-5, (xs, ws, err)
        # fused forms: BatchNorm prologue, ReLU, residual before / after the activation
        bn = nn.BatchNorm2d(cin).to(DEV).eval()
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.normal_(1, 0.1); bn.bias.normal_(0, 0.1)
        res = torch.randn_like(y)
        with torch.no_grad():
            for after in (True, False):
                got = m.forward_fused(x.to(DEV), bn, nn.ReLU(), res, after)
                conv = m(bn(x.to(DEV)))
                ref = F.relu(conv) + res if after else F.relu(conv + res)
                e2 = float((got - ref).abs().max() / ref.abs().max())
                assert e2 < 1e-5, (xs, ws, after, e2)
    # golden layers with multi-plane weights (reference outputs; 32 input channels: CUDA-core kernel)
    for rec in golden_layers:
        sp = rec['spec']
        if sp['w_quant'] in ('ls-1', 'fp') or sp['x_quant'] == 'fp':
            continue
        clamp = None if sp['alpha'] is None else {'kind': 'symmetric', 'alpha': sp['alpha']}
        m = QuantConv2d(sp['x_quant'], sp['w_quant'], sp['cin'], sp['cout'], sp['k'], clamp, stride=sp['stride'],
                        padding=sp['padding'], bias=sp['bias'])
        m.load_state_dict(rec['state'])
        m = m.to(DEV).eval()
        ops.reset_counters()
        with torch.no_grad():
            y = m(rec['x'].to(DEV)).cpu()
        assert 'fakequant' not in ops.LAUNCHES and ops.LAUNCHES.get('bconv_simple', 0) >= 2, ops.LAUNCHES
        err = float((y - rec['y']).abs().max() / rec['y'].abs().max())
        assert err < (1e-5 if sp['x_quant'] in ('ls-1',) or sp['x_quant'].startswith('gf') else 5e-2), (sp, err)


# ---------------------------------------------------------------------------------------------------------------
# uint8 pixel input (the data format in front of the path; runtime.set_pixel_input)
# ---------------------------------------------------------------------------------------------------------------
def _host_transform(u8, mean, std):
    """torchvision's ToTensor + Normalize on the host (quant/data loaders of the reference): x / 255, then (x - mean) / std."""
    t = u8.to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    return t.sub_(m).div_(s)


def test_uint8_pixels_expand_like_the_host_transform():
    """lsq_u8_expand and lsq_stem_fwd_u8 against ToTensor + Normalize evaluated on the host: bit-identical."""
    from ml_quant_b200 import ops, runtime
    runtime_strict()
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    lut = runtime.pixel_lut(mean, std)
    g = torch.Generator().manual_seed(41)
    for shape in [(3, 3, 224, 224), (2, 3, 61, 75), (1, 3, 7, 9), (5, 3, 32, 32), (2, 3, 33, 223)]:
        u8 = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g)
        want = _host_transform(u8, mean, std).to(DEV)
        got = ops.u8_expand(u8.to(DEV), lut)
        assert torch.equal(got, want), shape
        if shape[2] >= 7 and shape[3] >= 7:
            wt = torch.randn(64, 3, 7, 7, device=DEV) * 0.1
            b = torch.randn(64, device=DEV)
            img = ops.stem_pack(wt)
            assert bool(ops._C.lib().lsq_stem_is_fused(*[shape[0], shape[2], shape[3]]))
            assert torch.equal(ops.stem_fwd_u8(u8.to(DEV), lut, img, b), ops.stem_fwd(want, img, b)), shape
    # one channel, odd sizes (scalar path of the expand kernel)
    u8 = torch.randint(0, 256, (4, 1, 13, 11), dtype=torch.uint8, generator=g)
    lut1 = runtime.pixel_lut((0.1307,), (0.3081,))
    assert torch.equal(ops.u8_expand(u8.to(DEV), lut1), _host_transform(u8, (0.1307,), (0.3081,)).to(DEV))
    with pytest.raises(ValueError):
        ops.u8_expand(torch.zeros(2, 3, 4, 4, device=DEV), lut)               # fp32 input
    with pytest.raises(RuntimeError):
        ops.stem_fwd_u8(torch.zeros(1, 3, 64, 300, dtype=torch.uint8, device=DEV), lut, img, b)   # wider than the one-kernel route


@pytest.mark.parametrize('cfg,hw', [('imagenet_resnet18_ls1w_ls2a', (224, 224)), ('imagenet_resnet18_ls1w_ls2a', (64, 300)),
                                    ('cifar100_resnet18_ls1w_ls2a', (32, 32))])
def test_network_takes_uint8_pixel_batches(cfg, hw):
    """A network prepared with runtime.set_pixel_input classifies uint8 pixel batches exactly as it classifies the fp32
    tensors the host transform makes from them (fused stem, wide-image route, CIFAR stem; eager and as a CUDA graph)."""
    from ml_quant_b200 import runtime
    runtime_strict()
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    model = runtime.build_model(cfg, torch.device(DEV))
    runtime.calibrate(model, (3,) + hw)
    runtime.optimize_for_inference(model)
    runtime.set_pixel_input(model, mean, std)
    g = torch.Generator().manual_seed(43)
    u8 = torch.randint(0, 256, (6, 3) + hw, dtype=torch.uint8, generator=g)
    xf = _host_transform(u8, mean, std).to(DEV)
    with torch.no_grad():
        want = model(xf)
        got = model(u8.to(DEV))
    assert torch.equal(got, want)
    graphed = runtime.GraphedForward(model, u8.to(DEV))
    assert torch.equal(graphed(), want)
    # host batches through the pipeline, 1 byte per pixel
    pipe = runtime.HostPipeline(model, tuple(u8.shape), torch.device(DEV), use_graph=True, dtype=torch.uint8)
    out = pipe.run([u8.pin_memory(), u8.pin_memory()])
    assert torch.equal(out.to(DEV), want)
    # without set_pixel_input a uint8 batch is an error, not a silent cast
    other = runtime.build_model(cfg, torch.device(DEV))
    runtime.optimize_for_inference(other).eval()
    with torch.no_grad(), pytest.raises(RuntimeError):
        other(u8.to(DEV))


def test_reference_on_cuda_against_reference_on_cpu_second_witness():
    """the unmodified reference's opt_v1 on CUDA and on the CPU over the same rows.  Their picks may differ
    (the arg-min is decided among candidates whose fp32 costs tie to ~7 digits), but each pick satisfies the staged contract
    this repository's solver is held to against the other: fp64 cost within (1 + 1e-5).  Runs the staged reference in its
    own process (oracle/ref_witness.py); LSQ_WITNESS_OUT=<file> keeps the JSON line."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'oracle', 'ref_witness.py'), '24', '200704'],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    d = json.loads(line)
    if 'unavailable' in d:
        pytest.skip(d['unavailable'])
    if os.environ.get('LSQ_WITNESS_OUT'):
        open(os.environ['LSQ_WITNESS_OUT'], 'w').write(line + '\n')
    for case in d['cases']:
        assert case['max_fp64_cost_ratio_minus_1'] <= 1e-5, case


def test_plane_mean_matches_torch():
    """lsq_plane_mean (the classifier head's global average pool) against x.mean((2, 3)): <= 1e-6 relative (different,
    fixed summation order), deterministic and batch invariant."""
    from ml_quant_b200 import ops
    torch.manual_seed(51)
    for n, c, h, w in [(512, 512, 7, 7), (3, 64, 4, 4), (2, 5, 1, 1), (2, 3, 33, 17), (256, 512, 4, 4)]:
        x = torch.randn(n, c, h, w, device=DEV) + 0.5
        got = ops.plane_mean(x)
        want = x.double().mean(dim=(2, 3))
        assert got.shape == (n, c)
        assert float((got.double() - want).abs().max()) <= 1e-6 * float(want.abs().max()) + 1e-7
        assert torch.equal(got, ops.plane_mean(x))
        assert torch.equal(got[:1], ops.plane_mean(x[:1].contiguous()))
    with pytest.raises(ValueError):
        ops.plane_mean(torch.zeros(4, 4, device=DEV))
