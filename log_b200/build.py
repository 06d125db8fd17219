"""Build the sm_100a shared library in-tree with nvcc (no JIT cache, no torch extension machinery).

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, 'csrc')
LIB_DIR = os.path.join(ROOT, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'liblog_b200_raster.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas=-warn-spills']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(ROOT, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile log_b200/csrc/*.cu -> log_b200/_lib/liblog_b200_raster.so for sm_100a."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)')
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + '.tmp'
    extra = os.environ.get('LGR_NVCC_EXTRA', '').split()      # e.g. -DLGR_BWD_MIN_CTAS=5 for tuning experiments
    cmd = [nvcc] + NVCC_FLAGS + extra + ['-shared', '-o', tmp] + sources()
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
        print(' '.join(cmd))
    env = dict(os.environ)
    env.pop('CC', None)
    env.pop('CXX', None)
    subprocess.check_call(cmd, env=env)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
