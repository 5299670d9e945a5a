"""CPU reference arm of bench.py: time the UNMODIFIED reference package (oracle/_ref/reference_full/quant, staged by
oracle/make_ref.py) -- or, when it is absent, the oracle port -- on the host cores.  Test / measurement infrastructure.

    python oracle/ref_arm.py <config> <images_per_step> <steps> <warmup>        -> one JSON line on stdout

Runs in its own process because the reference's package is also called ``quant``: here sys.path holds the staged
reference FIRST, so ``import quant`` is the reference itself (its own QResNet, QuantConv2d, quantizers, F.conv2d).
The network is built from the same arch_config (ml_quant_b200/configs.py = the reference's YAML), seeded the same way
and calibrated like runtime.calibrate (train-mode forwards populate w_approximate.v1 and the BatchNorm statistics).
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, '_ref', 'reference_full')


def main():
    config, images, steps, warmup = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    have_ref = os.path.isdir(os.path.join(REF, 'quant', 'binary'))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') not in (ROOT, HERE)]
    if have_ref:
        sys.path.insert(0, REF)
    sys.path.append(ROOT)
    import torch
    import torch.nn.functional as F
    from ml_quant_b200 import configs          # plain dictionaries; imports nothing of the product path
    torch.set_num_threads(os.cpu_count() or 1)
    arch = configs.arch(config)
    shape = configs.input_shape(config)
    torch.manual_seed(0)
    if have_ref:
        import quant
        assert os.path.abspath(quant.__file__).startswith(REF), quant.__file__
        from quant.models.resnet import QResNet
        model = QResNet(loss_fn=F.cross_entropy, **arch)
        model.train()
        with torch.no_grad():
            for i in range(2):
                g = torch.Generator().manual_seed(100 + i)
                model(torch.randn(8, *shape, generator=g))
        model.eval()
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(images, *shape, generator=g)

        def fwd():
            with torch.no_grad():
                return model(x)
        kind = 'reference'
    else:
        from oracle import lsq_oracle as O
        from ml_quant_b200 import nets
        model = nets.QResNet(loss_fn=F.cross_entropy, **arch)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        for k in list(sd):
            if k.endswith('w_approximate.v1'):
                sd[k] = sd[k[:-len('w_approximate.v1')] + 'weight'].abs().mean(dim=(1, 2, 3))
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(images, *shape, generator=g)

        def fwd():
            with torch.no_grad():
                return O.resnet_forward(sd, arch, x)
        kind = 'port'
    for _ in range(warmup):
        fwd()
    t = time.perf_counter()
    for _ in range(steps):
        fwd()
    dt = time.perf_counter() - t
    print(json.dumps({'images_per_s': images * steps / dt, 's_per_step': dt / steps, 'kind': kind,
                      'cores': os.cpu_count() or 1, 'images_per_step': images, 'steps': steps}))


if __name__ == '__main__':
    main()
