"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python oracle/gen_golden.py
The GPU box has no /root/reference; the committed fixtures are what travels.  Every fixture stores
the inputs, the reference outputs and enough intermediate scales for the staged parity contract
(DESIGN.md "Parity contract"): function level, QuantConv2d layer level, whole-net level.
"""
import os
import sys

import torch
import yaml

REF = os.environ.get('ML_QUANT_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _ref():
    # make sure "quant" resolves to the reference, not to this repo's drop-in shim
    for k in [k for k in sys.modules if k == 'quant' or k.startswith('quant.')]:
        del sys.modules[k]
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') != os.path.abspath(os.path.join(OUT, '..', '..'))]
    sys.path.insert(0, REF)
    import quant.binary.quantization as q
    import quant.binary.optimal as o
    import quant.binary.binary_conv as bc
    import quant.models.resnet as rn
    import quant.models.lenet as ln
    assert q.__file__.startswith(REF), q.__file__
    return q, o, bc, rn, ln


def functions(q, o):
    """Quantizer functions on small ragged shapes (row length not a multiple of 3, 4 or 32)."""
    cases = []
    shapes = [(5, 3, 7, 9), (4, 16, 6, 6), (3, 2, 5, 5), (4, 64, 14, 14), (1, 64, 56, 56), (16, 32, 3, 3)]
    for si, shp in enumerate(shapes):
        g = torch.Generator().manual_seed(1234 + si)
        x = torch.randn(*shp, generator=g)
        if si % 2 == 1:
            x = x.clamp(-2, 2)          # exact ties at the clamp value
        if si == 3:
            x[1] = 0.75                  # an all-equal row (ternary edge case, optimal.py:86-118)
            x[2, :, :7] = 0.0            # zeros: sign(0) = +1
        rec = {'x': x}
        v1, xq = q.quantizer_ls_1(x)
        rec['ls1'] = {'v1': v1, 'xq': xq}
        for skip in (1, 3):
            v1, v2, xq = q.quantizer_ls_2(x, skip=skip)
            rec[f'ls2_s{skip}'] = {'v1': v1, 'v2': v2, 'xq': xq}
            v1, xq = q.quantizer_ls_ternary(x, skip=skip)
            rec[f'lsT_s{skip}'] = {'v1': v1, 'xq': xq}
            for tern in (False, True):
                a = x.view(x.shape[0], -1)[..., ::skip].abs()
                if a.shape[1] >= 3:
                    mask, vs = o.compute_mask(a, tern)
                    rec[f'cand_s{skip}_t{int(tern)}'] = {'counts': mask.sum(1), 'values': vs}
        for k in (1, 2, 3):
            vs, xq = q.quantizer_gf(x, k)
            rec[f'gf{k}'] = {'vs': torch.stack(vs), 'xq': xq}
        if x.numel() > 20000:            # keep the fixture small: scales pin the result, xq follows
            for v in rec.values():
                if isinstance(v, dict):
                    v.pop('xq', None)
        cases.append(rec)
    torch.save(cases, os.path.join(OUT, 'functions.pt'))
    print('functions.pt', len(cases))


def _x_scales(scheme, xin):
    import quant.binary.quantization as q
    if scheme == 'fp':
        return []
    if scheme == 'ls-1':
        return [q.quantizer_ls_1(xin)[0]]
    if scheme == 'ls-2':
        return list(q.quantizer_ls_2(xin)[:2])
    if scheme == 'ls-T':
        return [q.quantizer_ls_ternary(xin)[0]]
    return list(q.quantizer_gf(xin, int(scheme.split('-')[1]))[0])


def layers(bc):
    """Eval-mode QuantConv2d forwards (weight scales cached by one train-mode call first)."""
    specs = [
        # x_quant, w_quant, cin, cout, k, stride, pad, bias, alpha, N, H, W
        ('ls-2', 'ls-1', 64, 64, 3, 1, 1, True, 3.0, 3, 14, 14),
        ('ls-2', 'ls-1', 64, 128, 3, 2, 1, True, 3.0, 2, 14, 14),
        ('ls-1', 'ls-1', 64, 64, 3, 1, 1, True, 2.0, 3, 9, 11),
        ('ls-T', 'ls-1', 128, 128, 3, 1, 1, True, 2.0, 2, 7, 7),
        ('gf-2', 'ls-1', 64, 64, 3, 1, 1, False, None, 2, 8, 8),
        ('fp', 'ls-1', 20, 50, 5, 1, 0, True, None, 4, 12, 12),
        ('ls-2', 'ls-1', 20, 50, 5, 1, 0, True, 2.0, 4, 12, 12),
        ('ls-2', 'ls-2', 32, 32, 3, 1, 1, True, None, 2, 6, 6),
        ('ls-1', 'gf-2', 32, 48, 1, 1, 0, False, None, 2, 5, 5),
        ('ls-2', 'ls-1', 128, 256, 3, 2, 1, True, 3.0, 2, 14, 14),
    ]
    out = []
    for i, (xs, ws, cin, cout, k, st, pd, bias, alpha, n, h, w) in enumerate(specs):
        torch.manual_seed(500 + i)
        clamp = None if alpha is None else {'kind': 'symmetric', 'alpha': alpha}
        m = bc.QuantConv2d(xs, ws, cin, cout, k, clamp, stride=st, padding=pd, bias=bias)
        x = torch.randn(n, cin, h, w) * 1.5
        with torch.no_grad():
            m.train()
            m(x)                           # caches w_approximate.v*
            m.eval()
            y = m(x)
            xin = m.clamping_fn(x)
            xsc = _x_scales(xs, xin)
        rec = {'spec': dict(x_quant=xs, w_quant=ws, cin=cin, cout=cout, k=k, stride=st, padding=pd,
                            bias=bias, alpha=alpha),
               'state': {kk: vv.clone() for kk, vv in m.state_dict().items()},
               'x': x, 'x_scales': xsc, 'y': y}
        out.append(rec)
    torch.save(out, os.path.join(OUT, 'layers.pt'))
    print('layers.pt', len(out))


def ma_layers(bc):
    """QuantConv2d with a moving-average activation policy (activation_quantization.py:68-102): two train-mode
    steps track the per-batch mean scales, then the eval forward quantizes with the STORED scales (:90-98)."""
    out = []
    i = 0
    for mode in ('eval_only', 'train_and_eval'):
        for xs in ('ls-1', 'ls-2', 'ls-T', 'gf-2'):
            torch.manual_seed(900 + i)
            i += 1
            m = bc.QuantConv2d(xs, 'ls-1', 64, 64, 3, {'kind': 'symmetric', 'alpha': 2.0}, mode, 0.9, padding=1)
            xt = [torch.randn(4, 64, 9, 9) * 1.3 for _ in range(2)]
            x = torch.randn(3, 64, 9, 9) * 1.3
            with torch.no_grad():
                m.train()
                yt = [m(t) for t in xt]
                m.eval()
                y = m(x)
            out.append({'spec': dict(x_quant=xs, mode=mode, momentum=0.9, alpha=2.0),
                        'state': {kk: vv.clone() for kk, vv in m.state_dict().items()},
                        'x_train': xt, 'y_train_last': yt[-1], 'x': x, 'y': y})
    torch.save(out, os.path.join(OUT, 'ma_layers.pt'))
    print('ma_layers.pt', len(out))


def _calibrate(model, shape, seeds=(100, 101)):
    model.train()
    with torch.no_grad():
        for s in seeds:
            g = torch.Generator().manual_seed(s)
            model(torch.randn(*shape, generator=g))
    model.eval()


def nets(rn, ln):
    import torch.nn.functional as F
    ex = os.path.join(REF, 'examples')
    out = {}
    # CIFAR-100 ResNet ls1-weight/ls2-activation, width reduced 64 -> 16, one block per stage, to keep the fixture small
    for name, path in [('cifar_ls2', 'cifar100/cifar100_ls1_weight_ls2_activation_kd.yaml'),
                       ('cifar_lsT', 'cifar100/cifar100_ls1_weight_lsT_activation_kd.yaml'),
                       ('imagenet_ls1', 'imagenet/imagenet_ls1_kd.yaml')]:
        arch = yaml.safe_load(open(os.path.join(ex, path)))['model']['arch_config']
        arch['layer0']['n_in_channels'] = 16
        arch['num_blocks'] = [1, 1, 1, 1]
        arch['output_classes'] = 10
        torch.manual_seed(0)
        model = rn.QResNet(loss_fn=F.cross_entropy, **arch)
        hw = 64 if name.startswith('imagenet') else 32
        _calibrate(model, (8, 3, hw, hw))
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(4, 3, hw, hw, generator=g)
        with torch.no_grad():
            y = model(x)
        out[name] = {'arch': arch, 'state': model.state_dict(), 'x': x, 'y': y}
    # MNIST LeNet-5, BASELINE.json configs[0]
    arch = yaml.safe_load(open(os.path.join(ex, 'mnist/mnist_ls1_weight_fp_activation.yaml')))['model']['arch_config']
    arch['conv2_filters'] = 16   # fc1 shrinks 800x500 -> 256x160; keeps the fixture small
    torch.manual_seed(0)
    model = ln.QLeNet5(loss_fn=F.nll_loss, **arch)
    _calibrate(model, (16, 1, 28, 28))
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(64, 1, 28, 28, generator=g)
    with torch.no_grad():
        y = model(x)
    out['mnist_ls1w_fpa'] = {'arch': arch, 'state': model.state_dict(), 'x': x, 'y': y}
    torch.save(out, os.path.join(OUT, 'nets.pt'))
    print('nets.pt', {k: tuple(v['y'].shape) for k, v in out.items()})


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    q, o, bc, rn, ln = _ref()
    only = set(sys.argv[1:])         # e.g. `python oracle/gen_golden.py ma_layers` adds one fixture, leaves the rest
    if not only or 'functions' in only:
        functions(q, o)
    if not only or 'layers' in only:
        layers(bc)
    if not only or 'nets' in only:
        nets(rn, ln)
    if not only or 'ma_layers' in only:
        ma_layers(bc)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, 'KiB')
