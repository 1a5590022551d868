"""Build the CUDA extension IN-TREE for sm_100a (``ravnest_b200/ops/rvn_C.so``).

``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` cross-compiles on a CPU-only box; the built
``.so`` is git-ignored but travels to the GPU box with the working tree.  ``python -m ravnest_b200.ops.build``.
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD_DIR = os.path.join(HERE, '_build')
TARGET = os.path.join(HERE, 'rvn_C.so')

CUDA_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '--use_fast_math', '-std=c++17',
              '-Xptxas', '-v', '--expt-relaxed-constexpr']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu'))) + [os.path.join(CSRC, 'binding.cpp')]


def stale():
    if not os.path.exists(TARGET):
        return True
    built = os.path.getmtime(TARGET)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh'))
    return any(os.path.getmtime(s) > built for s in deps)


def build_all(verbose=False, force=False):
    if not force and not stale():
        return TARGET
    os.environ.setdefault('CUDA_HOME', '/usr/local/cuda')
    os.environ.setdefault('MAX_JOBS', str(min(8, os.cpu_count() or 4)))
    from torch.utils import cpp_extension
    os.makedirs(BUILD_DIR, exist_ok=True)
    cpp_extension.load(name='rvn_C', sources=sources(), extra_cuda_cflags=CUDA_FLAGS, extra_cflags=['-O3', '-std=c++17'],
                       extra_ldflags=['-L/usr/lib/x86_64-linux-gnu', '-l:libstdc++.so.6', '-lm'], build_directory=BUILD_DIR, verbose=verbose, is_python_module=False)
    shutil.copyfile(os.path.join(BUILD_DIR, 'rvn_C.so'), TARGET)
    return TARGET


if __name__ == '__main__':
    print(build_all(verbose='-q' not in sys.argv, force='--force' in sys.argv))
