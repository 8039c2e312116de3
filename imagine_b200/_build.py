"""Builds libimagine_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

Each .cu is compiled to an object under csrc/build/ (in parallel, only when stale), then linked.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(CSRC, 'build')
LIB_PATH = os.path.join(HERE, 'libimagine_b200.so')
SOURCES = ['api.cu', 'fields.cu', 'los.cu', 'los_rl.cu', 'likelihood.cu', 'gram_batch.cu', 'dense.cu', 'comm.cu']
HEADERS = ['common.cuh', 'linalg.cuh', 'los.cuh', os.path.join('..', '..', 'include', 'imagine_b200.h')]
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMPILE_FLAGS = ARCH_FLAGS + ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']
LINK_FLAGS = ARCH_FLAGS + ['-shared', '-cudart', 'static', '-ldl']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libimagine_b200.so cannot be built')


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + '.o')


def _headers_mtime():
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def _obj_stale(src, hdr_time):
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return os.path.getmtime(os.path.join(CSRC, src)) > t or hdr_time > t


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def _compile(nvcc, src, verbose):
    cmd = [nvcc] + COMPILE_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', _obj(src)]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed on %s:\n%s%s' % (src, proc.stdout, proc.stderr))
    return proc.stderr


def build(force=False, verbose=False):
    """Compile the CUDA sources into imagine_b200/libimagine_b200.so; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = _headers_mtime()
    todo = [s for s in SOURCES if force or _obj_stale(s, hdr_time)]
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        logs = list(ex.map(lambda s: _compile(nvcc, s, verbose), todo))
    if verbose:
        for s, log in zip(todo, logs):
            print('----', s)
            print(log)
    cmd = [nvcc] + LINK_FLAGS + ['-o', LIB_PATH] + [_obj(s) for s in SOURCES]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc link failed:\n' + proc.stdout + proc.stderr)
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build(force='--incremental' not in sys.argv, verbose=True))
