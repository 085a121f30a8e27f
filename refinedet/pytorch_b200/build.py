"""Build librefinedet_b200.so in-tree with nvcc for sm_100a.

Replaces the reference's ``utils/build.py`` + ``make.sh`` (distutils monkey-patch, Cython,
``-arch=sm_52``): one plain nvcc invocation per translation unit, no Cython, no numpy C-API.

    python -m refinedet.pytorch_b200.build          # build if stale
    python -m refinedet.pytorch_b200.build --force

``-fmad=false`` keeps every fp32 multiply/add un-contracted so the operation order matches the
reference's CPU arithmetic; ``-lineinfo`` lets ncu map SASS to source.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'librefinedet_b200.so')
SOURCES = ['rd_detect.cu', 'rd_match.cu', 'rd_loss.cu', 'rd_select.cu']
HEADERS = ['rd_common.cuh', 'rd_nms_core.cuh']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-fmad=false', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '-I', INCLUDE]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def _deps():
    files = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    files.append(os.path.join(INCLUDE, 'refinedet_b200.h'))
    files.append(os.path.abspath(__file__))
    return files


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False):
    """Compile the CUDA sources into ``lib/librefinedet_b200.so``; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(LIB_DIR, exist_ok=True)
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + '.o')
        extra = os.environ.get('RD_NVCC_EXTRA', '').split()          # tuning experiments (-DRD_SMALL_THREADS=160 ...)
        cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + '.tmp.%d' % os.getpid()
    cmd = [nvcc, '-shared', '-Xcompiler', '-fPIC', '-o', tmp] + objs + \
          ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
