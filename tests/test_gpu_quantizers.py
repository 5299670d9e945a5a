"""GPU parity of the quantizer kernels against the oracle / golden vectors, through the C ABI.

Parity contract: sign planes and everything downstream of given scales are BIT-EXACT;
scales that are fp32 means agree to 1e-6 relative (different summation order); the ls-2 / ls-T v1
solve is staged: our v1 is one of the reference's own candidates for that row and its
exact (fp64) cost is no worse than the reference pick's within 1e-5.
"""
import pytest
import torch

from oracle import lsq_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
REL = 1e-6   # fp32 mean, different summation order


def _close(a, b, rel=REL):
    return torch.allclose(a.cpu(), b, rtol=rel, atol=0)


def _solver_contract(rows, v_mine, v_ref, tern, skip):
    a = rows[..., ::skip].abs()
    srt, mask = O.candidate_mask(a, tern)
    for r in range(rows.shape[0]):
        cands = torch.masked_select(srt[r, 1:-1], mask[r])
        in_data = bool((cands == v_mine[r]).any())
        edge = tern and bool(a[r].min() > 0.5 * a[r].mean())     # optimal.py:86-118 appended value
        assert in_data or edge or cands.numel() == 0, (r, float(v_mine[r]))
    c_my, c_or = O.exact_cost(rows, v_mine, tern, skip), O.exact_cost(rows, v_ref, tern, skip)
    assert bool((c_my <= c_or * (1 + 1e-5) + 1e-12).all()), float((c_my / c_or).max())


def test_sign_and_ste_kat():
    # reference tests/binary/test_ste.py:13-36
    from quant.binary.ste import binarize
    x = torch.tensor([42, -42, 42, 42, 0, -1, 1, -4.2, 4.2], device=DEV)
    assert binarize(x).tolist() == [1, -1, 1, 1, 1, -1, 1, -1, 1]
    x = torch.tensor([42, -42, 0, -1, 1, -0.2, 0.2], device=DEV, requires_grad=True)
    binarize(x).sum().backward()
    assert x.grad.tolist() == [0, 0, 1, 1, 1, 1, 1]


def test_clamps_and_fp():
    from quant.binary import quantization as Q
    x = torch.tensor([-1.0, 0.0, 1.0, 2.0], device=DEV)
    assert torch.equal(Q.clamp_identity(x), x)
    assert Q.clamp_symmetric(x, 0.5).tolist() == [-0.5, 0, 0.5, 0.5]
    assert torch.equal(Q.QuantizerFP()(x), x)


def test_functions_against_golden(golden_functions):
    from quant.binary import quantization as Q
    from ml_quant_b200 import ops
    for rec in golden_functions:
        x = rec['x']
        xg = x.to(DEV)
        rows = x.reshape(x.shape[0], -1)
        v1, _ = Q.quantizer_ls_1(xg)
        assert _close(v1, rec['ls1']['v1'])
        if 'xq' in rec['ls1']:
            assert torch.equal(Q.quantizer_ls_1(xg, rec['ls1']['v1'].to(DEV))[1].cpu(), rec['ls1']['xq'])
        for skip in (1, 3):
            g2, gt = rec[f'ls2_s{skip}'], rec[f'lsT_s{skip}']
            _solver_contract(rows, ops.solve_v1(rows.to(DEV), False, skip).cpu(), g2['v1'], False, skip)
            _solver_contract(rows, ops.solve_v1(rows.to(DEV), True, skip).cpu(), gt['v1'], True, skip)
            _, v2, _ = Q.quantizer_ls_2(xg, g2['v1'].to(DEV), skip=skip)
            assert _close(v2, g2['v2'])
            if 'xq' in g2:
                assert torch.equal(Q.quantizer_ls_2(xg, g2['v1'].to(DEV), g2['v2'].to(DEV))[2].cpu(), g2['xq'])
                assert torch.equal(Q.quantizer_ls_ternary(xg, gt['v1'].to(DEV))[1].cpu(), gt['xq'])
        for k in (1, 2, 3):
            gg = rec[f'gf{k}']
            # every scale given the reference's previous ones (a flipped residual sign would show here)
            for i in range(k):
                vi = ops.row_absmean(rows.to(DEV), [gg['vs'][j].to(DEV) for j in range(i)])
                assert _close(vi, gg['vs'][i], 2e-6)
            if 'xq' in gg:
                assert torch.equal(Q.quantizer_gf(xg, k, [v.to(DEV) for v in gg['vs']])[1].cpu(), gg['xq'])


def test_ternary_all_equal_kat():
    # reference tests/binary/test_quantization.py:95-111
    from quant.binary import quantization as Q
    x = torch.ones(32, 3, 16, 16, device=DEV) * 2
    assert torch.all(Q.quantizer_ls_ternary(x)[1] == 2.0)
    torch.manual_seed(1234)
    x = torch.rand(32, 3, 16, 16)
    x[1] = 2
    x[9] = -3
    xq = Q.quantizer_ls_ternary(x.to(DEV))[1]
    assert torch.all(xq[1] == 2) and torch.all(xq[9] == -3)


def test_cost_ordering_properties():
    # reference tests/binary/test_quantization.py:36-165 (seed 1234, skip=1), at a GPU-sized batch
    from quant.binary import quantization as Q
    torch.manual_seed(1234)
    x = torch.randn(200, 3, 64, 64, device=DEV)

    def cost(xq):
        return torch.norm((xq - x).view(200, -1), dim=1)
    ls1 = cost(Q.quantizer_ls_1(x)[1])
    ls2 = cost(Q.quantizer_ls_2(x, skip=1)[2])
    lsT = cost(Q.quantizer_ls_ternary(x, skip=1)[1])
    gf = [cost(Q.quantizer_gf(x, k)[1]) for k in (1, 2, 3, 4)]
    assert torch.all(ls2 <= lsT) and torch.all(lsT <= ls1)
    assert torch.all(ls2 <= gf[1]) and torch.all(gf[1] <= ls1)
    assert all(torch.all(gf[i + 1] <= gf[i]) for i in range(3))
    # optimal scales beat random sub-optimal ones, per row
    sub = torch.randn(200, 1, 1, 1, device=DEV).abs() * torch.where(x >= 0, 1.0, -1.0)
    assert torch.all(ls1 <= cost(sub))


@pytest.mark.parametrize('tern', [False, True])
def test_solver_full_size_rows(tern):
    """BASELINE size rows (ImageNet layer-1 activations, 200704 elements, skip 3) against the oracle."""
    from ml_quant_b200 import ops
    torch.manual_seed(5)
    x = torch.randn(6, 64 * 56 * 56).clamp_(-3, 3)
    v, dg = ops.solve_v1(x.to(DEV), tern, 3, diag=True)
    v_ref = O.solve_v1(x, tern, 3, chunk=1).view(-1)
    _solver_contract(x, v.cpu(), v_ref, tern, 3)
    assert int(dg[:, 3].max()) == 0 and int(dg[:, 0].max()) <= 4
    # the clamp fused in the kernel equals clamping first
    y = torch.randn(6, 64 * 28 * 28) * 2
    assert torch.equal(ops.solve_v1(y.to(DEV), tern, 3, alpha=3.0), ops.solve_v1(y.clamp(-3, 3).to(DEV), tern, 3))


def test_solver_is_deterministic_and_batch_invariant():
    from ml_quant_b200 import ops
    torch.manual_seed(6)
    x = torch.randn(32, 64 * 28 * 28, device=DEV)
    a = ops.solve_v1(x, False, 3)
    assert torch.equal(a, ops.solve_v1(x, False, 3))
    assert torch.equal(a[5:9], ops.solve_v1(x[5:9].contiguous(), False, 3))


def test_solver_degenerate_rows():
    from ml_quant_b200 import ops
    x = torch.ones(4, 64, 28, 28) * 1.25
    x[1, :5] = 0.3
    x[2] = torch.randint(0, 4, (64, 28, 28)).float() * 0.5 - 0.75     # 4 discrete levels, heavy ties
    rows = x.reshape(4, -1)
    for tern in (False, True):
        v = ops.solve_v1(rows.to(DEV), tern, 3).cpu()
        v_ref = O.solve_v1(rows[:2], tern, 3, chunk=1).view(-1)
        assert torch.equal(v[:2], v_ref)
        assert torch.isfinite(v).all()
    tiny = torch.randn(3, 2, device=DEV)                                  # n < 3: defined as v1 = 0
    assert ops.solve_v1(tiny, False, 1).tolist() == [0, 0, 0]


def test_activation_quantizer_modules():
    # reference tests/binary/test_activation_quantization.py: constants 2.0 / 4.0 / 2.2 = 0.9*2+0.1*4
    from quant.binary import activation_quantization as A
    two, four = torch.ones(8, 3, 4, 4, device=DEV) * 2, torch.ones(8, 3, 4, 4, device=DEV) * 4
    for cls in (A.ActivationQuantizerLS1, A.ActivationQuantizerLS2, A.ActivationQuantizerLST):
        q = cls('off', 0.9).to(DEV)
        q.train()
        assert torch.all(q(two) == 2.0)
        q.eval()
        assert torch.all(q(four) == 4.0)                  # eval without moving average recomputes
        q = cls('eval_only', 0.9).to(DEV)
        q.train()
        assert torch.all(q(two) == 2.0) and torch.all(q(four) == 4.0)
        q.eval()
        out = q(four)
        if cls is A.ActivationQuantizerLS1:
            assert torch.allclose(out, torch.full_like(out, 2.2))
        q = cls('train_and_eval', 0.9).to(DEV)
        q.train()
        assert torch.all(q(two) == 2.0)
        out = q(four)
        if cls is A.ActivationQuantizerLS1:
            assert torch.allclose(out, torch.full_like(out, 2.2))
    q = A.ActivationQuantizerGF(2, 'eval_only', 0.9).to(DEV)
    q.train()
    q(two)
    assert set(q.state_dict()) == {'moving_avg_module.num_batches_tracked', 'moving_avg_module.momentum',
                                   'moving_avg_module.moving_average'}


def test_weight_quantizer_modules():
    # reference tests/binary/test_weight_quantization.py: train caches, eval re-uses
    from quant.binary import weight_quantization as W
    torch.manual_seed(0)
    w = torch.randn(16, 8, 3, 3, device=DEV)
    for q in (W.WeightQuantizerLS1(16), W.WeightQuantizerLS2(16), W.WeightQuantizerLST(16), W.WeightQuantizerGF(16, 2)):
        q = q.to(DEV)
        q.train()
        a = q(w)
        assert float(q.v1.abs().sum()) > 0
        q.eval()
        assert torch.equal(a, q(w))
        w2 = torch.rand(16, 8, 3, 3, device=DEV) + 0.5          # all positive: same signs, cached scales
        q.train()
        b = q(w2)
        q.eval()
        assert torch.equal(b, q(w2))


# BASELINE.json configs[4]: the 20 distinct conv-weight shapes of torchvision resnet50
RESNET50_WEIGHT_SHAPES = [(64, 3, 7, 7), (64, 64, 1, 1), (64, 64, 3, 3), (64, 256, 1, 1), (128, 128, 3, 3), (128, 256, 1, 1),
                          (128, 512, 1, 1), (256, 64, 1, 1), (256, 256, 3, 3), (256, 512, 1, 1), (256, 1024, 1, 1),
                          (512, 128, 1, 1), (512, 256, 1, 1), (512, 512, 3, 3), (512, 1024, 1, 1), (512, 2048, 1, 1),
                          (1024, 256, 1, 1), (1024, 512, 1, 1), (2048, 512, 1, 1), (2048, 1024, 1, 1)]


@pytest.mark.parametrize('tern', [False, True])
def test_solver_resnet50_weight_sweep(tern):
    """ls-2 / ls-T scale solve on ResNet-50-sized weight tensors (rows = output channels, 64..4608 elements,
    skip 3 and 1, no clamp): every row meets the solver contract against the oracle; ls-1 scales to 1e-6."""
    from ml_quant_b200 import ops
    g = torch.Generator().manual_seed(50)
    for shape in RESNET50_WEIGHT_SHAPES:
        w = torch.randn(*shape, generator=g) * (2.0 / (shape[1] * shape[2] * shape[3])) ** 0.5     # kaiming-like
        rows = w.reshape(shape[0], -1)[:96]                      # the oracle is O(n log n) per row on the CPU
        for skip in (3, 1):
            if (rows.shape[1] + skip - 1) // skip < 3:
                continue
            v = ops.solve_v1(rows.to(DEV), tern, skip).cpu()
            v_ref = O.solve_v1(rows, tern, skip, chunk=32).view(-1)
            _solver_contract(rows, v, v_ref, tern, skip)
        assert _close(ops.row_absmean(rows.to(DEV)), rows.abs().mean(1))


def test_multi_tensor_solver_is_bit_identical_to_single_launches():
    """lsq_solve_v1_multi / lsq_row_absmean_multi over all 53 ResNet-50 weight tensors (one grid per call; tensors
    whose sampled rows exceed the small-row kernel get their own launch inside): same bits as per-tensor calls."""
    from ml_quant_b200 import ops
    g = torch.Generator().manual_seed(51)
    ws = [(torch.randn(s[0], s[1] * s[2] * s[3], generator=g) * 0.05).to(DEV) for s in RESNET50_WEIGHT_SHAPES]
    ws.append(torch.randn(5, 2, generator=g).to(DEV))            # fewer than 3 solve elements: v1 = 0
    ws.append(torch.randn(300, 7000, generator=g).to(DEV))       # long rows: own launch inside the multi call
    for tern in (False, True):
        for skip in (3, 1):
            multi = ops.solve_v1_multi(ws, tern, skip)
            for w, v in zip(ws, multi):
                assert torch.equal(v, ops.solve_v1(w, tern, skip)), (tuple(w.shape), tern, skip)
    for alpha in (None, 0.05):
        multi = ops.row_absmean_multi(ws, alpha=alpha)
        for w, v in zip(ws, multi):
            assert torch.equal(v, ops.row_absmean(w, alpha=alpha)), tuple(w.shape)
    # more tensors than one parameter table holds (112): two grids
    many = [ws[i % 8][: 16 + i] for i in range(130)]
    multi = ops.solve_v1_multi(many, False, 3)
    for w, v in zip(many, multi):
        assert torch.equal(v, ops.solve_v1(w.contiguous(), False, 3))



def test_refresh_weight_scales_matches_train_mode_forward():
    """runtime.refresh_weight_scales (one multi-tensor launch per weight scheme) stores exactly the buffers a
    train-mode pass of each WeightQuantizer* stores (weight_quantization.py:27-34, :51-59, :75-82, :100-109)."""
    import copy
    import torch.nn as nn
    from quant.binary.binary_conv import QuantConv2d
    from ml_quant_b200 import runtime
    torch.manual_seed(5)
    specs = [('ls-1', 16, 32, 3), ('ls-2', 32, 64, 3), ('ls-T', 64, 64, 3), ('ls-2', 64, 128, 1), ('ls-1', 20, 50, 5),
             ('gf-2', 16, 16, 3), ('fp', 8, 8, 3), ('ls-T', 128, 256, 3)]
    net = nn.Sequential(*[QuantConv2d('fp', wq, cin, cout, k) for wq, cin, cout, k in specs]).to(DEV)
    want = copy.deepcopy(net)
    for m in want:
        m.w_approximate.train()
        with torch.no_grad():
            m.w_approximate(m.weight)
    runtime.refresh_weight_scales(net)
    checked = 0
    for a, b in zip(net, want):
        for (name, x), (_, y) in zip(a.w_approximate.named_buffers(), b.w_approximate.named_buffers()):
            assert torch.equal(x, y), (a.w_quant, name)
            assert float(y.abs().sum()) > 0
            checked += 1
    assert checked == 1 + 2 + 1 + 2 + 1 + 2 + 0 + 1
