"""Step-by-step GPU bring-up diagnostics (not a pytest file): python tests/gpu_bringup.py [stage ...]

Each stage prints PASS/FAIL lines and never stops at the first failure; the tensor-core stages run
in child processes so a device trap cannot hide the other results.  Output -> gpurun_out/bringup.log.
"""
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import lsq_oracle as O  # noqa: E402
from ml_quant_b200 import ops  # noqa: E402
from ml_quant_b200.binary import quantization as Q  # noqa: E402
from ml_quant_b200.binary.binary_conv import QuantConv2d  # noqa: E402

DEV = 'cuda:0'


def runtime_strict():
    from ml_quant_b200.runtime import strict_fp32
    strict_fp32()


def report(name, ok, extra=''):
    print(f"{'PASS' if ok else 'FAIL'} {name} {extra}", flush=True)


def stage_quant():
    torch.manual_seed(0)
    for shape in [(5, 3, 7, 9), (4, 64, 14, 14), (8, 64, 56, 56), (64, 64, 3, 3)]:
        x = torch.randn(*shape).clamp(-2.5, 2.5)
        xg = x.to(DEV)
        v1o, xqo = O.quant_ls1(x)
        v1, xq = Q.quantizer_ls_1(xg)
        report(f'ls1 v1 {shape}', torch.allclose(v1.cpu(), v1o, rtol=1e-6, atol=0), f'maxrel={((v1.cpu()-v1o).abs()/v1o).max():.2e}')
        _, xq2 = Q.quantizer_ls_1(xg, v1o.to(DEV))
        report(f'ls1 xq(injected v1) bit-exact {shape}', torch.equal(xq2.cpu(), xqo))
        # ls-2 with injected v1: v2 and xq
        v1o, v2o, xqo = O.quant_ls2(x, skip=3, chunk=4)
        _, v2, _ = Q.quantizer_ls_2(xg, v1o.to(DEV))
        report(f'ls2 v2(injected v1) {shape}', torch.allclose(v2.cpu(), v2o, rtol=1e-6, atol=0), f'maxrel={((v2.cpu()-v2o).abs()/v2o).max():.2e}')
        _, _, xq = Q.quantizer_ls_2(xg, v1o.to(DEV), v2o.to(DEV))
        report(f'ls2 xq(injected) bit-exact {shape}', torch.equal(xq.cpu(), xqo))
        v1t, xqt = O.quant_lsT(x, skip=3, chunk=4)
        _, xq = Q.quantizer_ls_ternary(xg, v1t.to(DEV))
        report(f'lsT xq(injected) bit-exact {shape}', torch.equal(xq.cpu(), xqt))
        vso, xqo = O.quant_gf(x, 3)
        vs, xq = Q.quantizer_gf(xg, 3)
        report(f'gf3 scales {shape}', all(torch.allclose(a.cpu(), b, rtol=2e-6, atol=0) for a, b in zip(vs, vso)))
        _, xq = Q.quantizer_gf(xg, 3, [v.to(DEV) for v in vso])
        report(f'gf3 xq(injected) bit-exact {shape}', torch.equal(xq.cpu(), xqo))


def stage_solve():
    torch.manual_seed(1)
    cases = [((6, 3, 16, 16), 1), ((6, 3, 16, 16), 3), ((8, 64, 28, 28), 3), ((4, 64, 56, 56), 3), ((4, 64, 28, 28), 1),
             ((64, 64, 3, 3), 3), ((3, 2, 5, 5), 1)]
    for shape, skip in cases:
        for tern in (False, True):
            x = torch.randn(*shape).clamp(-3, 3)
            rows = x.reshape(shape[0], -1)
            t = time.time()
            v_or = O.solve_v1(rows, tern, skip, chunk=2).view(-1)
            t_or = time.time() - t
            try:
                v, dg = ops.solve_v1(rows.to(DEV), tern, skip, diag=True)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                report(f'solve {shape} skip={skip} tern={tern}', False, repr(e))
                continue
            v = v.cpu()
            same = int((v == v_or).sum())
            c_my = O.exact_cost(rows, v, tern, skip)
            c_or = O.exact_cost(rows, v_or, tern, skip)
            ratio = float((c_my / c_or).max())
            a = rows[..., ::skip].abs()
            srt, mask = O.candidate_mask(a, tern)
            inset = 0
            for r in range(rows.shape[0]):
                cands = torch.masked_select(srt[r, 1:-1], mask[r])
                inset += int(bool((cands == v[r]).any()) or (tern and v[r] not in a[r]))
            ok = ratio <= 1 + 1e-5 and inset == rows.shape[0]
            report(f'solve {shape} skip={skip} tern={int(tern)}', ok,
                   f'identical={same}/{rows.shape[0]} in_candidate_set={inset} cost_ratio_max={ratio:.8f} '
                   f'diag[0]={dg[0].tolist()} oracle_s={t_or:.2f}')
    # degenerate rows
    x = torch.ones(4, 64, 28, 28) * 1.25
    x[1, :5] = 0.3
    for tern in (False, True):
        v = ops.solve_v1(x.reshape(4, -1).to(DEV), tern, 3).cpu()
        v_or = O.solve_v1(x.reshape(4, -1), tern, 3, chunk=1).view(-1)
        report(f'solve constant rows tern={int(tern)}', torch.equal(v, v_or), f'{v.tolist()} vs {v_or.tolist()}')
    x = torch.ones(32, 3, 16, 16) * 2
    report('lsT all-equal KAT', bool(torch.all(Q.quantizer_ls_ternary(x.to(DEV))[1] == 2.0)))
    # timing at the benchmark size
    x = torch.randn(512, 64 * 56 * 56, device=DEV).clamp_(-3, 3)
    for _ in range(2):
        ops.solve_v1(x, False, 3)
    torch.cuda.synchronize()
    t = time.time()
    v, dg = ops.solve_v1(x, False, 3, diag=True)
    torch.cuda.synchronize()
    print(f'INFO solve 512x200704 skip3: {1e3*(time.time()-t):.2f} ms, passes max {int(dg[:,0].max())} collected mean {float(dg[:,1].float().mean()):.0f} flags {int(dg[:,3].max())}')


def _layer_case(xs, cin, cout, k, st, pd, alpha, n, h, w, impl, seed=0):
    torch.manual_seed(seed)
    clamp = None if alpha is None else {'kind': 'symmetric', 'alpha': alpha}
    m = QuantConv2d(xs, 'ls-1', cin, cout, k, clamp, stride=st, padding=pd, bias=True)
    x = torch.randn(n, cin, h, w) * 1.5
    xin = x if alpha is None else x.clamp(-alpha, alpha)
    wv1 = m.weight.detach().abs().mean(dim=(1, 2, 3))
    scales, _ = O.quantize_activation(xin, xs, chunk=2)
    y_ref, ints = O.plane_conv_identity(xin, m.weight.detach(), m.bias.detach(), xs, scales, wv1, st, pd)
    m = m.to(DEV).eval()
    m.w_approximate.v1.copy_(wv1.to(DEV))
    xg = x.to(DEV)
    g = ops.act_geometry(n, cin, h, w, k, k, st, pd)
    npl = m._num_planes()
    known = scales[:1] if xs == 'ls-T' else scales
    planes, _ = ops.encode_act(xg, g, [s.to(DEV) for s in known[:npl]], npl, alpha, False)
    table = torch.stack(scales + scales if xs == 'ls-T' else scales).to(DEV)
    y = ops.bconv2d(planes, g, npl, table, m.packed_weights(), m.w_approximate.v1, m.bias, cout, impl)
    torch.cuda.synchronize()
    err = float((y.cpu() - y_ref).abs().max() / y_ref.abs().max())
    return err


def stage_conv(impl):
    cases = [('ls-2', 64, 64, 3, 1, 1, 3.0, 3, 14, 14), ('ls-1', 64, 64, 3, 1, 1, 2.0, 3, 9, 11),
             ('ls-2', 64, 128, 3, 2, 1, 3.0, 2, 14, 14), ('ls-T', 128, 128, 3, 1, 1, 2.0, 2, 7, 7),
             ('ls-2', 128, 256, 3, 2, 1, 3.0, 2, 14, 14), ('ls-2', 256, 256, 3, 1, 1, 3.0, 2, 14, 14),
             ('ls-2', 64, 64, 3, 1, 1, 3.0, 4, 56, 56), ('ls-2', 512, 512, 3, 1, 1, 3.0, 3, 7, 7),
             ('ls-2', 64, 128, 3, 2, 1, 3.0, 2, 7, 9)]
    if impl == 1:
        cases += [('ls-2', 20, 50, 5, 1, 0, 2.0, 4, 12, 12), ('gf-3', 16, 24, 3, 1, 1, None, 2, 8, 8),
                  ('ls-2', 16, 32, 3, 2, 1, 3.0, 2, 9, 9)]
    for c in cases:
        try:
            err = _layer_case(*c, impl)
            report(f'bconv impl={impl} {c}', err < 1e-5, f'err={err:.2e}')
        except Exception as e:  # noqa: BLE001
            report(f'bconv impl={impl} {c}', False, repr(e)[:300])
            if 'CUDA' in repr(e) or 'cuda' in repr(e):
                traceback.print_exc()
                return


def stage_module():
    """QuantConv2d end to end (own scales) against the oracle forward."""
    runtime_strict()
    for xs, cin, cout, st in [('ls-2', 64, 64, 1), ('ls-1', 64, 128, 2), ('ls-T', 64, 64, 1), ('gf-2', 64, 64, 1)]:
        torch.manual_seed(3)
        m = QuantConv2d(xs, 'ls-1', cin, cout, 3, {'kind': 'symmetric', 'alpha': 3.0}, stride=st, padding=1)
        x = torch.randn(4, cin, 28, 28) * 1.5
        wv1 = m.weight.detach().abs().mean(dim=(1, 2, 3))
        y_ref = O.quant_conv2d(x, m.weight.detach(), m.bias.detach(), xs, 'ls-1', [wv1], 3.0, st, 1, chunk=2)
        m = m.to(DEV).eval()
        m.w_approximate.v1.copy_(wv1.to(DEV))
        with torch.no_grad():
            y = m(x.to(DEV))
            m.allow_packed = False
            y_gen = m(x.to(DEV))
        e1 = float((y.cpu() - y_ref).abs().max() / y_ref.abs().max())
        e2 = float((y_gen.cpu() - y_ref).abs().max() / y_ref.abs().max())
        report(f'module {xs} {cin}->{cout} s{st}', e1 < 1e-3, f'packed err={e1:.2e} generic err={e2:.2e}')


def stage_solve_timing():
    for c, hw in [(64, 56), (128, 28), (256, 14), (512, 7)]:
        x = torch.randn(512, c * hw * hw, device=DEV).clamp_(-3, 3)
        for tern in (False, True):
            for _ in range(2):
                ops.solve_v1(x, tern, 3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                v, dg = ops.solve_v1(x, tern, 3, diag=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print('INFO phase kcycles mean [pop+zero, hist_g, hist_l, reduce, scan+flag, thread0, collect_g, collect_l, creduce, sort+eval, tail]:',
                  [round(float(dg[:, 4 + i].float().mean()) / 1e3, 1) for i in range(11)], flush=True)
            print(f'INFO solve 512x{c*hw*hw} tern={int(tern)}: {ms:.3f} ms  ({x.numel()*4/ms/1e6:.0f} GB/s of fp32 input) passes max {int(dg[:,0].max())} '
                  f'collected mean {float(dg[:,1].float().mean()):.0f} cands mean {float(dg[:,2].float().mean()):.1f} flags {int(dg[:,3].max())}', flush=True)


STAGES = {'quant': stage_quant, 'solve': stage_solve, 'conv1': lambda: stage_conv(1), 'conv2': lambda: stage_conv(2),
          'module': stage_module, 'solve_timing': stage_solve_timing}

if __name__ == '__main__':
    todo = sys.argv[1:]
    if todo:
        for s in todo:
            try:
                STAGES[s]()
            except Exception:  # noqa: BLE001
                print(f'FAIL stage {s} crashed', flush=True)
                traceback.print_exc()
    else:
        print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
        for s in ['quant', 'solve', 'conv1', 'conv2', 'module']:
            print(f'===== stage {s}', flush=True)
            r = subprocess.run(['timeout', '600', sys.executable, os.path.abspath(__file__), s], cwd=ROOT)
            print(f'===== stage {s} exit {r.returncode}', flush=True)
