"""Stage the parts of the reference checkout that must travel to the GPU box into oracle/_ref/ (git-ignored).

    python oracle/make_ref.py          # in the build container, where /root/reference is mounted

What is staged, and why (nothing here is product code; no product module imports from oracle/):
  * oracle/_ref/reference/quant/{common,data,utils}  -- the reference's orchestration layer (SURVEY.md section 2 rows
    9-14, out of scope to rebuild): `quant.common.tasks.classification_task`, checkpoints, metrics ...  They are the
    CALLERS of the hot path; the drop-in `quant` package of this repository (quant/__init__.py) extends its search
    path with this directory when $ML_QUANT_REFERENCE points at it, so the reference's examples and tests run
    unmodified on top of the B200 modules.  quant/binary, quant/models and quant/__init__.py are deliberately NOT
    staged: those names must resolve to this repository.
  * oracle/_ref/reference/{tests,examples}           -- the reference's own pytest suite and example drivers / YAMLs,
    run by scripts/run_reference_tests.py against the shim (VERDICT r1 missing #1).
  * oracle/_ref/reference_full/quant                  -- the whole unmodified reference package, used ONLY by the CPU
    reference arm of bench.py (`--impl reference`, cpu_baseline.kind = "reference").
The GPU box has no /root/reference; gpurun ships oracle/_ref with the snapshot (it is not in .gpurunignore).
"""
import os
import shutil
import sys

REF = os.environ.get('ML_QUANT_REFERENCE_SRC', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')


def main() -> int:
    if not os.path.isdir(os.path.join(REF, 'quant')):
        print(f'{REF}/quant not found: nothing staged (oracle/_ref is only made in the build container)')
        return 0
    ign = shutil.ignore_patterns('__pycache__', '*.pyc')
    shim = os.path.join(OUT, 'reference')
    full = os.path.join(OUT, 'reference_full')
    for d in (shim, full):
        if os.path.isdir(d):
            shutil.rmtree(d)
    os.makedirs(os.path.join(shim, 'quant'))
    for sub in ('common', 'data', 'utils'):
        shutil.copytree(os.path.join(REF, 'quant', sub), os.path.join(shim, 'quant', sub), ignore=ign)
    for sub in ('tests', 'examples'):
        shutil.copytree(os.path.join(REF, sub), os.path.join(shim, sub), ignore=ign)
    shutil.copytree(os.path.join(REF, 'quant'), os.path.join(full, 'quant'), ignore=ign)
    n = sum(len(fs) for _, _, fs in os.walk(OUT))
    print(f'staged {n} files under {OUT}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
