"""
In-tree build of libcopter_b200.so (the C-ABI CUDA library) for sm_100a.  nvcc cross-compiles
without a GPU, so this runs in the build container; the .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = [os.path.join(PKG, 'csrc', 'copter_kernels.cu')]
HDR = [os.path.join(ROOT, 'include', 'copter_b200.h'), os.path.join(PKG, 'csrc', 'copter_physics.cuh'), os.path.join(PKG, 'csrc', 'copter_core.h'),
       os.path.join(PKG, 'csrc', 'copter_policy.cuh'), os.path.join(PKG, 'csrc', 'copter_policy_tc.cuh')]
LIB = os.path.join(PKG, 'libcopter_b200.so')

NVCC_FLAGS = ["-std=c++17", "-O3", "-fmad=false", "-DCOPTER_NO_CONTRACT=1", '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found: libcopter_b200.so cannot be built')


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in SRC + HDR + [__file__])


def build(force=False, verbose=False):
    """Compiles every CUDA source of the package into libcopter_b200.so. Returns its path."""
    if not force and not is_stale():
        return LIB
    src = [s for s in SRC if os.path.exists(s)]
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB + '.tmp'] + src
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), r.stderr))
    os.replace(LIB + '.tmp', LIB)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
