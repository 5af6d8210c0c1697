"""Builds libvipnerf_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m vipnerf_b200.build [--force]

The shared object is written next to this file so that it travels with the source tree (it is git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(HERE, 'build')
LIB_PATH = os.path.join(HERE, 'libvipnerf_b200.so')
SOURCES = ['api.cu', 'stage_kernels.cu', 'frame_kernels.cu', 'prior_kernels.cu', 'mlp_fp32.cu', 'train_kernels.cu', 'gemm_tc.cu', 'mlp_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr']
# extra -D switches for experiments (e.g. VIPNERF_NVCC_DEFINES="VIPNERF_ONES_4K"), part of the build digest
NVCC_FLAGS += ['-D' + d for d in os.environ.get('VIPNERF_NVCC_DEFINES', '').split() if d]


def find_nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(nvcc):
        raise RuntimeError('nvcc not found: the CUDA extension cannot be built')
    return nvcc


def _digest(paths) -> str:
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _all_inputs():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    files.append(os.path.join(os.path.dirname(HERE), 'include', 'vipnerf.h'))
    return files


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every CUDA source for sm_100a and links the shared library; returns its path."""
    stamp = os.path.join(OBJ_DIR, 'stamp.txt')
    digest = _digest(_all_inputs())
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp) and open(stamp).read() == digest:
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJ_DIR, src + '.log'), 'w') as f:
            f.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, '-shared', '-o', LIB_PATH, *objs, '-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
