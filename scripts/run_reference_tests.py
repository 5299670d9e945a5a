"""Run the reference's own pytest files (staged by oracle/make_ref.py) against this repository's `quant` shim.

    python scripts/run_reference_tests.py [pytest args ...]      # on a GPU box; log -> gpurun_out/ref_tests.log

tests/binary, tests/models, tests/utils/test_moving_average.py and tests/common/test_tasks.py of the reference
(train -> checkpoint -> restore -> skip-training -> KD student on RandomQuantDataLoader), unmodified.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref', 'reference')


def main() -> int:
    if not os.path.isdir(os.path.join(REF, 'tests')):
        print('oracle/_ref/reference is missing: run `python oracle/make_ref.py` in the build container first')
        return 2
    env = dict(os.environ)
    env['ML_QUANT_REFERENCE'] = REF
    # The staged tree comes first so that `tests` is the reference's test package (this repository has one of the
    # same name); its quant/ has no __init__.py (a namespace portion), so `import quant` still resolves to this
    # repository's regular package, whose __path__ then picks up quant.common / data / utils from the staged tree.
    env['PYTHONPATH'] = os.pathsep.join([REF, os.path.join(ROOT, 'scripts'), ROOT, env.get('PYTHONPATH', '')])
    targets = [os.path.join(REF, 'tests', t) for t in
               ('binary', 'models', os.path.join('utils', 'test_moving_average.py'), os.path.join('common', 'test_tasks.py'))]
    cmd = [sys.executable, '-m', 'pytest', '-c', os.devnull, '--rootdir', REF, '-p', 'ref_tests_plugin', '-p', 'no:cacheprovider',
           '-q', '-rA', '--import-mode=importlib'] + sys.argv[1:] + targets
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    log = os.path.join(ROOT, 'gpurun_out', 'ref_tests.log')
    with open(log, 'w') as f:
        f.write('$ ' + ' '.join(cmd) + '\n')
        f.flush()
        rc = subprocess.run(cmd, env=env, cwd=REF, stdout=f, stderr=subprocess.STDOUT).returncode
        f.write(f'\nexit code {rc}\n')
    print(open(log).read()[-3000:])
    return rc


if __name__ == '__main__':
    sys.exit(main())
