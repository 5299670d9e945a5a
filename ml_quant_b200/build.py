"""Build ml_quant_b200/csrc/*.cu into the in-tree C-ABI library liblsq_b200.so (sm_100a only).

    python -m ml_quant_b200.build [--force]

The library links only the CUDA runtime; PyTorch is not involved (ctypes binds it, see _C.py).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liblsq_b200.so')
SOURCES = ['lsq_quant.cu', 'lsq_solve.cu', 'lsq_qact.cu', 'lsq_bconv.cu', 'lsq_bconv_tc.cu', 'lsq_stem.cu', 'lsq_pwconv.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--use_fast_math=false']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'lsq_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    flags += [f for f in os.environ.get('LSQ_NVCC_EXTRA', '').split() if f]      # e.g. -DLSQ_TC_DIAG (development builds)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    headers.append(os.path.join(HERE, '..', 'include', 'lsq_b200.h'))
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(CSRC, src[:-3] + '.o')
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), newest_header):
            return obj
        cmd = [_nvcc()] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
        subprocess.run(cmd, check=True)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([_nvcc(), '-shared', '-o', LIB] + objs + ['-lcudart'], check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
