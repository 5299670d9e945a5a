"""GPU parity of the binary convolution and of whole networks against golden vectors / the oracle."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import lsq_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def runtime_strict():
    from ml_quant_b200.runtime import strict_fp32
    strict_fp32()


def _planes_from_gpu(planes, g, npl):
    """Unpack the GPU plane buffer back to bool [npl, n, c, h, w] for bit-exact comparison."""
    buf = planes.cpu()[: npl * g.nphase * g.vtot * g.cw].view(npl, g.nphase, g.vtot, g.cw)
    out = torch.zeros(npl, g.n, g.c, g.h, g.w, dtype=torch.bool)
    s = g.stride
    yy, xx = torch.meshgrid(torch.arange(g.h), torch.arange(g.w), indexing='ij')
    phase = ((yy % s) * 2 + (xx % s)) if s == 2 else torch.zeros_like(yy)
    for n in range(g.n):
        v = g.lead + (n * g.rows_per_sample + g.ph + yy // s) * g.pitch + xx // s
        words = buf[:, phase, v]                       # [npl, h, w, cw]
        for c in range(g.c):
            out[:, n, c] = ((words[..., c >> 5] >> (c & 31)) & 1).bool()
    return out


def _run_layer(rec, impl):
    from quant.binary.binary_conv import QuantConv2d
    from ml_quant_b200 import ops
    s, st = rec['spec'], rec['state']
    clamp = None if s['alpha'] is None else {'kind': 'symmetric', 'alpha': s['alpha']}
    m = QuantConv2d(s['x_quant'], s['w_quant'], s['cin'], s['cout'], s['k'], clamp, stride=s['stride'],
                    padding=s['padding'], bias=s['bias'])
    m.load_state_dict(st, strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize('impl', [1, 2])
def test_layers_with_injected_scales(golden_layers, impl):
    """Encode + binary conv with the reference's activation scales: planes bit-exact, y within 1e-5 max|y|."""
    from ml_quant_b200 import ops
    ran = 0
    for rec in golden_layers:
        s, st = rec['spec'], rec['state']
        if s['w_quant'] != 'ls-1' or s['x_quant'] == 'fp':
            continue
        m = _run_layer(rec, impl)
        x = rec['x']
        n, c, h, w = x.shape
        g = ops.act_geometry(n, c, h, w, s['k'], s['k'], s['stride'], s['padding'])
        npl = m._num_planes()
        if impl == 2 and not ops.tc_supported(g, npl, s['cout']):
            continue
        sc = rec['x_scales']
        tern = s['x_quant'] == 'ls-T'
        known = sc[:1] if tern else sc
        planes, _ = ops.encode_act(x.to(DEV), g, [v.to(DEV) for v in known[:npl]], npl, s['alpha'], False)
        xin = x if s['alpha'] is None else x.clamp(-s['alpha'], s['alpha'])
        ref_planes = torch.stack(O.bit_planes(xin, s['x_quant'], sc))
        assert torch.equal(_planes_from_gpu(planes, g, npl), ref_planes), s
        table = torch.stack(sc + sc if tern else sc).to(DEV)
        y = ops.bconv2d(planes, g, npl, table, m.packed_weights(), m.w_approximate.v1, m.bias, s['cout'], impl)
        err = float((y.cpu() - rec['y']).abs().max() / rec['y'].abs().max())
        assert err < 1e-5, (s, err)
        ran += 1
    assert ran >= 3


def test_tensor_core_and_cuda_core_kernels_agree():
    """Same integer accumulators -> outputs equal to fp32 rounding of the epilogue (<= 1e-6 max|y|)."""
    from ml_quant_b200 import ops
    torch.manual_seed(11)
    for (n, cin, cout, h, w, st, npl) in [(3, 64, 64, 56, 56, 1, 2), (2, 128, 256, 28, 28, 2, 2), (2, 512, 512, 7, 7, 1, 1),
                                          (5, 256, 128, 15, 13, 1, 2), (2, 64, 128, 17, 19, 2, 1)]:
        x = torch.randn(n, cin, h, w, device=DEV)
        wt = torch.randn(cout, cin, 3, 3, device=DEV)
        g = ops.act_geometry(n, cin, h, w, 3, 3, st, 1)
        assert ops.tc_supported(g, npl, cout)
        v1 = torch.rand(n, device=DEV) + 0.5
        planes, v2 = ops.encode_act(x, g, [v1] if npl == 2 else [], npl, None, True)
        table = torch.stack([v1, v2]) if npl == 2 else v2[None]
        wp = ops.pack_weights(wt)
        ws, b = torch.rand(cout, device=DEV), torch.randn(cout, device=DEV)
        y1 = ops.bconv2d(planes, g, npl, table, wp, ws, b, cout, 1)
        y2 = ops.bconv2d(planes, g, npl, table, wp, ws, b, cout, 2)
        assert float((y1 - y2).abs().max() / y1.abs().max()) < 1e-6


def test_module_end_to_end_layers(golden_layers):
    """QuantConv2d.forward with its own solve.  Scales that are means agree to 1e-6 and the output to
    1e-5 max|y|; for ls-2 / ls-T the output is compared after checking the staged solver contract."""
    runtime_strict()
    from ml_quant_b200 import ops
    for rec in golden_layers:
        s = rec['spec']
        m = _run_layer(rec, 0)
        with torch.no_grad():
            y = m(rec['x'].to(DEV)).cpu()
        scale = rec['y'].abs().max()
        err = float((y - rec['y']).abs().max() / scale)
        if s['x_quant'] in ('ls-2', 'ls-T'):
            xin = rec['x'] if s['alpha'] is None else rec['x'].clamp(-s['alpha'], s['alpha'])
            rows = xin.reshape(xin.shape[0], -1)
            v1 = ops.solve_v1(rows.to(DEV), s['x_quant'] == 'ls-T', 3).cpu()
            if torch.equal(v1, rec['x_scales'][0]):
                assert err < 1e-5, (s, err)
            else:                                   # a different but equally good candidate
                c_my = O.exact_cost(rows, v1, s['x_quant'] == 'ls-T', 3)
                c_or = O.exact_cost(rows, rec['x_scales'][0], s['x_quant'] == 'ls-T', 3)
                assert bool((c_my <= c_or * (1 + 1e-5)).all()) and err < 5e-2, (s, err)
        else:
            assert err < 1e-5, (s, err)


def test_packed_and_generic_routes_agree():
    runtime_strict()
    from quant.binary.binary_conv import QuantConv2d
    torch.manual_seed(2)
    for xs in ('ls-1', 'ls-2', 'ls-T', 'gf-2', 'gf-3'):
        m = QuantConv2d(xs, 'ls-1', 64, 64, 3, {'kind': 'symmetric', 'alpha': 2.0}, padding=1).to(DEV)
        x = torch.randn(3, 64, 20, 20, device=DEV) * 1.5
        with torch.no_grad():
            m.train()
            m(x)
            m.eval()
            a = m(x)
            m.allow_packed = False
            b = m(x)
        assert float((a - b).abs().max() / b.abs().max()) < 1e-5, xs


def test_fp_and_structured_kats():
    # reference tests/binary/test_binary_conv.py:18-67
    from quant.binary.binary_conv import QuantConv2d
    runtime_strict()
    torch.manual_seed(1234)
    x = torch.randn(8, 3, 40, 40, device=DEV, requires_grad=True)
    ref = nn.Conv2d(3, 30, 5).to(DEV)
    mine = QuantConv2d('fp', 'fp', 3, 30, 5).to(DEV)
    mine.weight, mine.bias = nn.Parameter(ref.weight), nn.Parameter(ref.bias)
    assert torch.equal(ref(x), mine(x))
    x = torch.zeros(1, 3, 8, 8)
    x[0, :, :4, 4:] = -1
    x[0, :, 4:, :4] = 2
    x[0, :, 4:, 4:] = -3
    conv = QuantConv2d('fp', 'ls-1', 3, 1, (4, 4), stride=4, bias=False).to(DEV)
    with torch.no_grad():
        conv.weight[0, 0, 0, :3].abs_()          # unbalanced signs so the sums are not ~0
        conv.weight[0, 0, 1, :3].abs_()
    y = conv(x.to(DEV)).squeeze()
    assert y.shape == (2, 2) and y[0, 0] == 0
    assert torch.isclose(y[1, 0], -2 * y[0, 1]) and torch.isclose(y[1, 1], 3 * y[0, 1])
    conv = QuantConv2d('ls-1', 'ls-1', 3, 16, (2, 2)).to(DEV)
    y = conv(torch.randn(4, 3, 8, 8, device=DEV))
    assert bool((y.abs().amax(dim=(2, 3)) <= 12 + conv.bias.abs() + 1e-5).all())


def test_training_step_runs_and_ste_gradients_flow():
    from quant.binary.binary_conv import QuantConv2d
    torch.manual_seed(0)
    m = QuantConv2d('ls-2', 'ls-1', 8, 8, 3, {'kind': 'symmetric', 'alpha': 2.0}, padding=1).to(DEV).train()
    x = torch.randn(4, 8, 10, 10, device=DEV, requires_grad=True)
    m(x).square().mean().backward()
    assert x.grad is not None and float(x.grad.abs().sum()) > 0 and float(m.weight.grad.abs().sum()) > 0
    assert float(m.w_approximate.v1.abs().sum()) > 0


def test_nets_against_golden(golden_nets):
    """Whole networks with the reference's state_dict.  MNIST (fp activations, ls-1 weights) and the
    ls-1 activation ResNet have no ill-posed solve: 1e-4 of max|logit| (sign flips of near-zero BN
    outputs are the only discrete effect).  ls-2 / ls-T: logits within 5e-2 and same top-1."""
    runtime_strict()
    from ml_quant_b200.nets import QLeNet5, QResNet
    for name, rec in golden_nets.items():
        cls, loss = (QLeNet5, F.nll_loss) if name.startswith('mnist') else (QResNet, F.cross_entropy)
        m = cls(loss_fn=loss, **rec['arch'])
        m.load_state_dict(rec['state'])
        m = m.to(DEV).eval()
        with torch.no_grad():
            y = m(rec['x'].to(DEV)).cpu()
        err = float((y - rec['y']).abs().max() / rec['y'].abs().max())
        if name in ('mnist_ls1w_fpa',):
            assert err < 1e-5, (name, err)
        elif name == 'imagenet_ls1':
            assert err < 2e-3, (name, err)
        else:
            assert err < 5e-2, (name, err)
            assert torch.equal(y.argmax(1), rec['y'].argmax(1)), name


def test_full_size_resnet18_against_live_oracle():
    """BASELINE network at full width and resolution (batch 4): GPU forward vs the oracle on the CPU."""
    runtime_strict()
    from ml_quant_b200 import configs, runtime
    model = runtime.build_model('imagenet_resnet18_ls1w_ls1a', torch.device(DEV))
    runtime.calibrate(model, (3, 224, 224), batches=1, batch=8)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 3, 224, 224, generator=g)
    with torch.no_grad():
        y = model(x.to(DEV)).cpu()
    y_ref = O.resnet_forward(sd, configs.arch('imagenet_resnet18_ls1w_ls1a'), x)
    err = float((y - y_ref).abs().max() / y_ref.abs().max())
    # End to end this network is chaotic too, if less than the ls-2 one: every layer takes sign(bn(x)) of ~10^6 values
    # per sample and the few within fp32 rounding of zero flip.  Round 1's epilogue (fmul, fadd) happened to land at
    # 4e-3 for this seed, round 2's (one fma: one rounding less) at 3.7e-2 -- both are "a handful of flipped signs,
    # amplified by 16 random-init layers".  The strict statement is the teacher-forced per-layer test below
    # (test_full_size_network_layer_by_layer_with_own_scales[imagenet_resnet18_ls1w_ls1a]).
    assert err < 0.15, err
    assert int((y.argmax(1) == y_ref.argmax(1)).sum()) >= 3
    # batch sharding is result preserving (per-sample scales, eval BN): bit-identical halves
    with torch.no_grad():
        a = model(x[:2].to(DEV)).cpu()
    assert torch.equal(a, y[:2])


def test_cuda_graph_and_host_pipeline_match_eager():
    from ml_quant_b200 import runtime
    model = runtime.build_model('cifar100_resnet18_ls1w_ls2a', torch.device(DEV))
    runtime.calibrate(model, (3, 32, 32), batches=1, batch=16)
    x = torch.randn(8, 3, 32, 32)
    with torch.no_grad():
        eager = model(x.to(DEV)).cpu()
    fwd = runtime.GraphedForward(model, x.to(DEV))
    assert torch.equal(fwd().cpu(), eager)
    # the pipeline must classify what is uploaded: different batches per step, last one checked, for both
    # buffer parities, with and without graphs
    other = [torch.randn(8, 3, 32, 32).pin_memory() for _ in range(3)]
    hx = x.pin_memory()
    for use_graph in (True, False):
        pipe = runtime.HostPipeline(model, (8, 3, 32, 32), torch.device(DEV), use_graph=use_graph)
        assert torch.equal(pipe.run([other[0], other[1], hx]).clone(), eager)
        assert torch.equal(pipe.run([other[2], hx]).clone(), eager)
        with torch.no_grad():
            want = model(other[1].to(DEV)).cpu()
        assert torch.equal(pipe.run([hx, other[1]]).clone(), want)


def test_fused_prologue_epilogue_matches_composition():
    """forward_fused(x, bn, nonlin, residual) equals nonlin(conv(bn(x))) (+) residual built from the same
    kernels, bit for bit, when bn(x) is formed with one rounding (the kernels use fmaf)."""
    import torch.nn as nn
    from quant.binary.binary_conv import QuantConv2d, bn_affine
    torch.manual_seed(4)
    for xs, stride, nonlin, after in [('ls-2', 1, nn.ReLU(), True), ('ls-1', 2, nn.PReLU(), True),
                                      ('ls-T', 1, nn.ReLU(), False), ('gf-2', 1, nn.PReLU(64), True)]:
        m = QuantConv2d(xs, 'ls-1', 64, 64, 3, {'kind': 'symmetric', 'alpha': 3.0}, stride=stride, padding=1).to(DEV)
        bn = nn.BatchNorm2d(64).to(DEV)
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.5)
            bn.running_var.uniform_(0.5, 2.0)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.3)
            if isinstance(nonlin, nn.PReLU):
                nonlin.weight.uniform_(0.1, 0.4)
        nonlin = nonlin.to(DEV)
        x = torch.randn(5, 64, 18, 18, device=DEV) * 2
        with torch.no_grad():
            m.train()
            m(x)
            m.eval()
            bn.eval()
            a, b = bn_affine(bn)
            xb = (x.double() * a.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1)).float()
            conv = m(xb)
            res = torch.randn_like(conv)
            want = nonlin(conv) + res if after else nonlin(conv + res)
            got = m.forward_fused(x, bn, nonlin, res, after)
            # the plain (non-packed) route composes the same thing from torch ops
            m.allow_packed = False
            slow = m.forward_fused(x, bn, nonlin, res, after)
        assert torch.equal(got, want), (xs, float((got - want).abs().max()))
        assert float((slow - want).abs().max() / want.abs().max()) < 2e-2


def test_optimized_network_matches_plain_network():
    """optimize_for_inference: every fused XnorBasicBlock reproduces the plain block on the same input to
    1e-4 of max|y| (BatchNorm folded with one rounding instead of two); end to end the ls-1 activation
    networks are only required to stay finite: random-init sign networks amplify such perturbations (flipped
    signs of near-zero activations, near-tied v1 picks -- SURVEY.md H1).  state_dict keys are untouched."""
    from ml_quant_b200 import runtime
    from ml_quant_b200.runtime import _xnor_block_fused
    runtime_strict()
    for cfg, shape in [('cifar100_resnet18_ls1w_ls2a', (3, 32, 32)), ('imagenet_resnet18_ls1w_ls1a', (3, 64, 64))]:
        torch.manual_seed(7)
        model = runtime.build_model(cfg, torch.device(DEV))
        runtime.calibrate(model, shape, batches=2, batch=32)
        x = torch.randn(8, *shape, device=DEV)
        with torch.no_grad():
            h = model.blocks[0](x)
            for blk in list(model.blocks)[1:]:
                plain = blk(h)
                fused = _xnor_block_fused(blk, h)
                assert float((plain - fused).abs().max() / plain.abs().max()) < 1e-4, cfg
                h = plain
            plain = model(x)
            fused = runtime.optimize_for_inference(model)(x)
        assert list(model.state_dict()) == list(runtime.build_model(cfg).state_dict())
        assert bool(torch.isfinite(fused).all())
        # no end-to-end numeric bar: a random-init sign network is chaotic (one flipped sign of a near-zero
        # activation changes a sample's logits), the block-by-block check above is the strict one
        assert fused.shape == plain.shape


def test_fused_stem_kernel_matches_torch():
    """lsq_stem_fwd vs conv + bias + max-pool + ReLU in fp32 ATen (cuDNN off): the one-kernel route (tcgen05 kind::f16,
    fp16 hi/lo operand pairs, pooling in the epilogue) for images up to 250 pixels wide, the two-kernel route (kind::tf32
    3xTF32 convolution + pool kernel) beyond; channels with tiny and huge weights exercise the per-row weight scaling."""
    import torch.nn.functional as F
    from ml_quant_b200 import ops
    runtime_strict()
    torch.manual_seed(9)
    L = ops._C.lib()
    for n, h, w in [(3, 224, 224), (2, 64, 64), (2, 61, 75), (1, 32, 40), (2, 250, 250), (1, 260, 300), (5, 7, 7),
                    (300, 96, 96), (2, 33, 223)]:
        x = torch.randn(n, 3, h, w, device=DEV)
        wt = torch.randn(64, 3, 7, 7, device=DEV) * 0.1
        wt[5] *= 1e-3
        wt[6] *= 300.0
        wt[7] = 0.0
        b = torch.randn(64, device=DEV)
        want = F.relu(F.max_pool2d(F.conv2d(x, wt, b, 2, 3), 3, 2, 1))
        assert ops.stem_supported(n, h, w)
        assert bool(L.lsq_stem_is_fused(n, h, w)) == (w <= 250)
        got = ops.stem_fwd(x, ops.stem_pack(wt), b)
        assert got.shape == want.shape
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 5e-6, (n, h, w, err)
        # per output channel as well (a channel with small weights must not inherit the error scale of a large one)
        errc = float(((got - want).abs().amax(dim=(0, 2, 3)) / want.abs().amax(dim=(0, 2, 3)).clamp_min(1e-20)).max())
        assert errc < 1e-5, (n, h, w, errc)
    # an output buffer that is only 4-byte aligned (a C caller's sub-allocation): the pooled rows leave by the scalar walk
    x = torch.randn(2, 3, 64, 64, device=DEV)
    img = ops.stem_pack(wt)
    want = ops.stem_fwd(x, img, b)
    buf = torch.zeros(want.numel() + 1, device=DEV)
    ws = torch.empty(64, device=DEV)
    ops._C.check(L.lsq_stem_fwd(x.data_ptr(), 2, 64, 64, img.data_ptr(), b.data_ptr(), ws.data_ptr(), buf.data_ptr() + 4,
                                torch.cuda.current_stream().cuda_stream), 'lsq_stem_fwd')
    assert torch.equal(buf[1:].view_as(want), want)
    # out-of-range pixels are clamped to the fp16 range on the one-kernel route (documented domain): finite output
    x = torch.randn(1, 3, 32, 32, device=DEV)
    x[0, 0, 3, 3] = 1e30
    assert bool(torch.isfinite(ops.stem_fwd(x, ops.stem_pack(wt), b)).all())


def test_pointwise_strided_conv_matches_torch():
    """lsq_pwconv_fwd (tcgen05 kind::tf32, 3xTF32 split) vs F.conv2d(kernel 1x1, stride) in fp32 ATen."""
    from ml_quant_b200 import ops
    runtime_strict()
    torch.manual_seed(10)
    for n, cin, cout, h, w, st in [(5, 64, 128, 56, 56, 2), (3, 128, 256, 28, 28, 2), (9, 256, 512, 14, 14, 2),
                                   (2, 32, 128, 17, 23, 2), (2, 16, 256, 9, 7, 1), (70, 64, 128, 8, 8, 2)]:
        assert ops.pwconv_supported(cin, cout)
        x = torch.randn(n, cin, h, w, device=DEV)
        wt = torch.randn(cout, cin, device=DEV) * (1.0 / cin) ** 0.5
        b = torch.randn(cout, device=DEV)
        want = F.conv2d(x, wt.view(cout, cin, 1, 1), b, st)
        got = ops.pwconv_fwd(x, ops.pwconv_pack(wt), b, cout, st)
        assert got.shape == want.shape
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 5e-6, (n, cin, cout, h, w, st, err)


def test_quantlinear_matches_oracle():
    """QuantLinear = 1x1 QuantConv2d on [N, in, 1, 1]: tensor-core route (features % 64 == 0) and generic route
    against the oracle's quantizers + F.linear (scales injected through the cached buffers / oracle solve)."""
    from quant.binary.binary_conv import QuantLinear
    torch.manual_seed(12)
    for xs, fin, fout, alpha in [('ls-2', 512, 256, 3.0), ('ls-1', 256, 128, 2.0), ('ls-T', 192, 64, 3.0), ('ls-2', 100, 10, None)]:
        clamp = None if alpha is None else {'kind': 'symmetric', 'alpha': alpha}
        m = QuantLinear(xs, 'ls-1', fin, fout, clamp=clamp).to(DEV)
        x = torch.randn(33, fin) * 1.5
        with torch.no_grad():
            m.train(); m(x.to(DEV)); m.eval()             # caches the weight scales like the reference
            y = m(x.to(DEV)).cpu()
        xin = x if alpha is None else x.clamp(-alpha, alpha)
        x4 = xin.view(33, fin, 1, 1)
        scales, xq = O.quantize_activation(x4, xs, chunk=8)
        w = m.weight.detach().cpu()
        wq = w.abs().mean(1, keepdim=True) * torch.where(w >= 0, 1.0, -1.0)
        want = F.linear(xq.view(33, fin), wq, m.bias.detach().cpu())
        # ls-2 / ls-T: our v1 may be another candidate of the same cost -> compare with OUR scales too
        err = float((y - want).abs().max() / want.abs().max())
        assert err < (1e-5 if xs == 'ls-1' else 5e-2), (xs, fin, fout, err)
        assert y.shape == (33, fout)
    assert m(torch.randn(2, 3, 100, device=DEV)).shape == (2, 3, 10)


def test_tensor_core_kernel_random_shapes():
    """bconv_tc against the CUDA-core kernel (same exact integer accumulators) over random shapes: batch sizes and
    image sizes that straddle tiles, both strides, 1 and 2 planes, every epilogue variant."""
    import random
    from ml_quant_b200 import ops
    rng = random.Random(1234)
    torch.manual_seed(13)
    tried = 0
    while tried < 24:
        cin, cout = rng.choice([64, 128, 192, 256, 512]), rng.choice([64, 128, 256, 384, 512])
        n, h, w = rng.randint(1, 9), rng.randint(5, 40), rng.randint(5, 40)
        st, npl = rng.choice([1, 1, 2]), rng.choice([1, 2])
        g = ops.act_geometry(n, cin, h, w, 3, 3, st, 1)
        if g is None or not ops.tc_supported(g, npl, cout):
            continue
        tried += 1
        x = torch.randn(n, cin, h, w, device=DEV)
        wt = torch.randn(cout, cin, 3, 3, device=DEV)
        v1 = torch.rand(n, device=DEV) + 0.5
        planes, v2 = ops.encode_act(x, g, [v1] if npl == 2 else [], npl, None, True)
        table = torch.stack([v1, v2]) if npl == 2 else v2[None]
        wp = ops.pack_weights(wt)
        ws, b = torch.rand(cout, device=DEV), torch.randn(cout, device=DEV)
        act = rng.choice([0, 1, 2])
        prelu = (torch.rand(rng.choice([1, cout]), device=DEV) * 0.5) if act == 2 else None
        res = torch.randn(n, cout, g.ho, g.wo, device=DEV) if rng.random() < 0.7 else None
        after = rng.random() < 0.5
        y1 = ops.bconv2d(planes, g, npl, table, wp, ws, b, cout, 1, None, res, act, prelu, after)
        y2 = ops.bconv2d(planes, g, npl, table, wp, ws, b, cout, 2, None, res, act, prelu, after)
        err = float((y1 - y2).abs().max() / y1.abs().max())
        assert err < 1e-6, (n, cin, cout, h, w, st, npl, act, res is not None, after, err)


def test_full_size_resnet18_ls2_headline_config():
    """The benchmark's own network (ImageNet ResNet-18, ls-1 weights / ls-2 activations, 224 x 224, random init)
    against the oracle on the CPU, plain and after optimize_for_inference.  This network is chaotic: the oracle's
    OWN logits move by 2-3 % of max|logit| when its input is perturbed by 1e-7 relative
    (scripts/dev/oracle_sensitivity.py), because ls-2 picks v1 among candidates whose costs tie to 7 digits
     and 16 random-init sign layers amplify every flipped pick.  Two correct implementations with
    different fp32 summation orders therefore agree only coarsely end to end (measured 4-8 %); the strict parity
    checks are the per-layer tests.  Here: logits within 15 % of max|logit|, top-1 equal on at least 3 of 4
    samples, and -- exactly -- identical results for a sharded batch and for the CUDA-graph replay."""
    runtime_strict()
    from ml_quant_b200 import configs, runtime
    cfg = 'imagenet_resnet18_ls1w_ls2a'
    model = runtime.build_model(cfg, torch.device(DEV))
    runtime.calibrate(model, (3, 224, 224), batches=1, batch=8)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 3, 224, 224, generator=g)
    y_ref = O.resnet_forward(sd, configs.arch(cfg), x)
    with torch.no_grad():
        y = model(x.to(DEV)).cpu()
        fused_model = runtime.optimize_for_inference(model)
        yf = fused_model(x.to(DEV)).cpu()
        ya = fused_model(x[:2].to(DEV)).cpu()
    for out in (y, yf):
        err = float((out - y_ref).abs().max() / y_ref.abs().max())
        assert err < 0.15, err
        assert int((out.argmax(1) == y_ref.argmax(1)).sum()) >= 3
    assert torch.equal(ya, yf[:2])
    fwd = runtime.GraphedForward(fused_model, x.to(DEV))
    assert torch.equal(fwd().cpu(), yf)


@pytest.mark.parametrize('cfg', ['imagenet_resnet18_ls1w_ls2a', 'imagenet_resnet18_ls1w_ls1a'])
def test_full_size_network_layer_by_layer_with_own_scales(cfg):
    """Teacher-forced full-size check of the headline network and of the XNOR network (BASELINE config 3; ls-1
    activations: v1 = mean|x| to 1e-6, one plane) (batch 2, 224 x 224): every one of the 16 QuantConv2d
    layers, fed with the GPU's own input of that layer, must (a) pick a v1 that meets the solver contract against the
    oracle, (b) reproduce the oracle's v2 for that v1 to 1e-6, and (c) produce the oracle's convolution output for
    those scales to 1e-5 of max|y|.  Together with the exact kernels around them this pins the whole forward,
    independent of the chaotic amplification of v1 tie-picks between layers."""
    runtime_strict()
    from ml_quant_b200 import runtime
    from tests.test_gpu_quantizers import _solver_contract
    model = runtime.build_model(cfg, torch.device(DEV))
    runtime.calibrate(model, (3, 224, 224), batches=1, batch=8)
    layers = runtime.quant_layers(model)
    assert len(layers) == 16
    xnor = cfg.endswith('ls1a')
    rec = {}
    for i, m in enumerate(layers):
        orig = m.quantize_input

        def wrapped(x, g, prologue=None, _i=i, _orig=orig):
            planes, table = _orig(x, g, prologue)
            rec[_i] = [x.detach().cpu(), table.detach().cpu(), None]
            return planes, table
        m.quantize_input = wrapped
        m.register_forward_hook(lambda mod, inp, out, _i=i: rec[_i].__setitem__(2, out.detach().cpu()))
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 3, 224, 224, generator=g)
    with torch.no_grad():
        model(x.to(DEV))
    for i, m in enumerate(layers):
        xi, table, y = rec[i]
        alpha = m.clamp_alpha
        xin = xi.clamp(-alpha, alpha)
        rows = xin.reshape(2, -1)
        if xnor:
            v1 = table[0]
            assert torch.allclose(v1, rows.abs().mean(1), rtol=1e-6, atol=0), (i, v1)
            y_ref, _ = O.plane_conv_identity(xin, m.weight.detach().cpu(), m.bias.detach().cpu(), 'ls-1', [v1],
                                             m.w_approximate.v1.detach().cpu(), m.stride[0], m.padding[0])
        else:
            v1, v2 = table[0], table[1]
            _solver_contract(rows, v1, O.solve_v1(rows, False, 3, chunk=1).view(-1), False, 3)
            b1 = torch.where(xin >= 0, 1.0, -1.0)
            v2_ref = (xin - v1.view(-1, 1, 1, 1) * b1).abs().mean(dim=(1, 2, 3))
            assert torch.allclose(v2, v2_ref, rtol=1e-6, atol=0), (i, v2, v2_ref)
            y_ref, _ = O.plane_conv_identity(xin, m.weight.detach().cpu(), m.bias.detach().cpu(), 'ls-2', [v1, v2],
                                             m.w_approximate.v1.detach().cpu(), m.stride[0], m.padding[0])
        err = float((y - y_ref).abs().max() / y_ref.abs().max())
        assert err < 1e-5, (i, err)


def test_full_size_fused_network_layer_by_layer():
    """The same teacher-forced check for the path the benchmark runs (runtime.optimize_for_inference: BatchNorm
    + clamp folded into the quantizer kernels, ReLU + residual into the convolution epilogue): every fused layer,
    fed with the GPU's own input and residual, matches the oracle's composition for the GPU's scales to 1e-5 of
    max|y|; BatchNorm is formed with one rounding (x * a + b in double, rounded), which is what the kernels' fmaf does."""
    runtime_strict()
    import torch.nn as nn
    from ml_quant_b200 import runtime
    from ml_quant_b200.binary.binary_conv import bn_affine
    from tests.test_gpu_quantizers import _solver_contract
    model = runtime.build_model('imagenet_resnet18_ls1w_ls2a', torch.device(DEV))
    runtime.calibrate(model, (3, 224, 224), batches=1, batch=8)
    runtime.optimize_for_inference(model)
    layers = runtime.quant_layers(model)
    rec = {}
    for i, m in enumerate(layers):
        oq, of = m.quantize_input, m.forward_fused

        def q_wrapped(x, g, prologue=None, _i=i, _oq=oq):
            planes, table = _oq(x, g, prologue)
            rec.setdefault(_i, {})['table'] = table.detach().cpu()
            return planes, table

        def f_wrapped(x, bn=None, nonlin=None, residual=None, residual_after_act=True, _i=i, _of=of):
            out = _of(x, bn, nonlin, residual, residual_after_act)
            rec.setdefault(_i, {}).update(x=x.detach().cpu(), bn=bn, nonlin=nonlin, after=residual_after_act,
                                          res=None if residual is None else residual.detach().cpu(), y=out.detach().cpu())
            return out
        m.quantize_input, m.forward_fused = q_wrapped, f_wrapped
    g = torch.Generator().manual_seed(78)
    x = torch.randn(2, 3, 224, 224, generator=g)
    with torch.no_grad():
        model(x.to(DEV))
    assert len(rec) == 16
    for i, m in enumerate(layers):
        r = rec[i]
        a, b = bn_affine(r['bn'])
        xb = (r['x'].double() * a.cpu().double().view(1, -1, 1, 1) + b.cpu().double().view(1, -1, 1, 1)).float()
        xin = xb.clamp(-m.clamp_alpha, m.clamp_alpha)
        v1, v2 = r['table'][0], r['table'][1]
        rows = xin.reshape(2, -1)
        _solver_contract(rows, v1, O.solve_v1(rows, False, 3, chunk=1).view(-1), False, 3)
        conv, _ = O.plane_conv_identity(xin, m.weight.detach().cpu(), m.bias.detach().cpu(), 'ls-2', [v1, v2],
                                        m.w_approximate.v1.detach().cpu(), m.stride[0], m.padding[0])
        assert isinstance(r['nonlin'], nn.ReLU)
        res = r['res'] if r['res'] is not None else torch.zeros_like(conv)
        want = F.relu(conv) + res if r['after'] else F.relu(conv + res)
        err = float((r['y'] - want).abs().max() / want.abs().max())
        assert err < 1e-5, (i, err)


def test_packed_checkpoint_round_trip():
    """export_packed / load_packed: sign images + scales instead of fp32 weights; a fresh
    model loaded from the packed form gives bit-identical logits, a train-mode weight re-solve returns the same
    v1, and the quantized weights take 1/32 of their fp32 size.  The oracle runs the unpacked state too."""
    from ml_quant_b200 import ops, runtime
    cfg = 'cifar100_resnet18_ls1w_ls2a'
    model = runtime.build_model(cfg, torch.device(DEV))
    runtime.calibrate(model, (3, 32, 32), batches=1, batch=16)
    x = torch.randn(8, 3, 32, 32).to(DEV)
    with torch.no_grad():
        want = model(x)
    packed = runtime.export_packed(model)
    assert all(not v.is_cuda for v in packed.values())
    layers = runtime.quant_layers(model)
    qbytes = sum(m.weight.numel() * 4 for m in layers)
    pbytes = sum(v.numel() * 4 for k, v in packed.items() if k.endswith('weight_bits'))
    assert len([k for k in packed if k.endswith('weight_bits')]) == len(layers) == 16
    assert pbytes * 32 == qbytes
    # the sign image is the documented layout: bit c&31 of word c>>5 of [cout, tap, word] is W >= 0
    name, m0 = next((n, m) for n, m in model.named_modules() if m is layers[0])
    bits = packed[f'{name}.weight_bits']
    w = m0.weight.detach().cpu()
    cout, cin, kh, kw = w.shape
    got = torch.stack([(bits[:, :, c >> 5] >> (c & 31)) & 1 for c in range(cin)], 1).bool()     # [cout, cin, taps]
    assert torch.equal(got, w.reshape(cout, cin, kh * kw) >= 0)
    fresh = runtime.build_model(cfg, torch.device(DEV), seed=7)
    res = runtime.load_packed(fresh, packed)
    fresh.eval()
    assert not res.missing_keys and not res.unexpected_keys
    with torch.no_grad():
        assert torch.equal(fresh(x), want)
        fm = runtime.quant_layers(fresh)[3]
        assert torch.equal(ops.row_absmean(fm.weight.detach().reshape(fm.out_channels, -1)), fm.w_approximate.v1)
    runtime.optimize_for_inference(fresh)
    runtime.optimize_for_inference(model)
    with torch.no_grad():
        assert torch.equal(fresh(x), model(x))
    with pytest.raises(ValueError):
        runtime.load_packed(fresh, {k: v for k, v in packed.items() if k != '_format'})
