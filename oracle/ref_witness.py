"""Second witness for the staged v1 contract: the UNMODIFIED reference's opt_v1 (oracle/_ref/reference_full,
staged by oracle/make_ref.py) run on the CPU and on CUDA over the same seeded rows -- how far does the reference disagree with
itself?  Test infrastructure.  Own process because the reference's package is also called ``quant``.

    python oracle/ref_witness.py [rows] [len]      -> one JSON line
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, '_ref', 'reference_full')


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    ln = int(sys.argv[2]) if len(sys.argv) > 2 else 200704
    if not os.path.isdir(os.path.join(REF, 'quant', 'binary')):
        print(json.dumps({'unavailable': 'oracle/_ref/reference_full is not staged (python oracle/make_ref.py)'}))
        return
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') not in (ROOT, HERE)]
    sys.path.insert(0, REF)
    sys.path.append(ROOT)
    import torch
    import quant
    assert os.path.abspath(quant.__file__).startswith(REF), quant.__file__
    from quant.binary.optimal import opt_v1
    from oracle.lsq_oracle import exact_cost
    out = {'rows': rows, 'len': ln, 'skip': 3, 'cases': []}
    for name, alpha, ternary in [('ls-2 clamp 3', 3.0, False), ('ls-T clamp 3', 3.0, True), ('ls-2 no clamp', None, False)]:
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(rows, ln, generator=g)
        if alpha is not None:
            x = x.clamp(-alpha, alpha)
        v_cpu = opt_v1(x, ternary, 3).reshape(-1)
        v_gpu = opt_v1(x.cuda(), ternary, 3).reshape(-1).cpu()
        c_cpu = exact_cost(x, v_cpu, ternary, 3)
        c_gpu = exact_cost(x, v_gpu, ternary, 3)
        differ = v_cpu != v_gpu
        rel = ((v_cpu - v_gpu).abs() / v_cpu.abs().clamp_min(1e-30))
        out['cases'].append({
            'case': name, 'rows_with_different_pick': int(differ.sum()),
            'max_rel_diff_of_v1': float(rel.max()), 'median_rel_diff_where_different': float(rel[differ].median()) if bool(differ.any()) else 0.0,
            'max_fp64_cost_ratio_minus_1': float(((c_gpu / c_cpu) - 1.0).abs().max()),
            'v1_cpu': [float(v) for v in v_cpu[:4]], 'v1_cuda': [float(v) for v in v_gpu[:4]]})
    print(json.dumps(out))


if __name__ == '__main__':
    main()
