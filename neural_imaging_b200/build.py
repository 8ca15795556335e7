"""Build libni_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
TAG = os.environ.get('NI_BUILD_TAG', '')          # development variants (tools/): e.g. NI_BUILD_TAG=prof -> build_prof/, libni_b200_prof.so
OBJ = os.path.join(HERE, 'build' + ('_' + TAG if TAG else ''))
LIB = os.path.join(HERE, 'libni_b200' + ('_' + TAG if TAG else '') + '.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
# a tagged build is a DEVELOPMENT variant: -DNI_DEV compiles in the environment switches, the generation-2 tcgen05 kernels and the
# hardware probes of csrc/dev/ (declared in include/ni_b200_dev.h). The shipping libni_b200.so has none of them.
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-I', CSRC] + (['-DNI_DEV'] if TAG else []) + os.environ.get('NI_NVCC_EXTRA', '').split()      # e.g. -DNI_TC_PROFILE


def _sources():
    src = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    if TAG:
        src += sorted(os.path.join('dev', f) for f in os.listdir(os.path.join(CSRC, 'dev')) if f.endswith('.cu'))
    return src


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: {}\n{}\n{}'.format(' '.join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    if force or jobs or _stale(LIB, objs):
        run([NVCC, '-shared', '-o', LIB] + objs + ['-lcudart'])   # no -lcuda: driver entry points are resolved at run time
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
