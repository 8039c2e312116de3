"""Builds libimagine_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libimagine_b200.so')
SOURCES = ['api.cu', 'fields.cu', 'los.cu', 'los_il4.cu', 'los_rl.cu', 'likelihood.cu', 'gram_batch.cu', 'dense.cu']
HEADERS = ['common.cuh', 'linalg.cuh', 'los.cuh', os.path.join('..', '..', 'include', 'imagine_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libimagine_b200.so cannot be built')


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile the CUDA sources into imagine_b200/libimagine_b200.so; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ['-o', LIB_PATH] + SOURCES
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force=True, verbose=True))
