"""In-tree build of libpb200.so (sm_100a only; nvcc cross-compiles without a GPU).

    python -m prospect_public_b200.build [--force]

The shared object is written next to the sources' package (prospect_public_b200/_lib/) so it
travels to the GPU box with the repo snapshot; nothing is JIT-cached outside the tree.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libpb200.so')
SOURCES = [os.path.join(CSRC, 'pb200_kernels.cu')]
HEADERS = [os.path.join(CSRC, h) for h in ('pb200_device.cuh', 'pb200_wide.cuh', 'pb200_umma.cuh', 'pb200_hosted.cuh')] + \
          [os.path.join(os.path.dirname(PKG), 'include', 'pb200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v',
]


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isfile(cand) or cand == 'nvcc'):
            return cand
    raise RuntimeError('nvcc not found')


def is_stale():
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > built for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into _lib/libpb200.so.  Returns the path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + SOURCES + ['-o', LIB_PATH]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
