"""
In-tree build of libdemux_b200.so (sm_100a only) with nvcc.  No torch / pybind involvement: the library is a
plain C-ABI shared object (include/demux_b200.h) loaded with ctypes by `demuxalot_b200._native`.

    python -m demuxalot_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / 'csrc'
LIB_PATH = CSRC / 'libdemux_b200.so'
SOURCES = ['api.cu', 'builder.cu', 'table.cu', 'estep.cu', 'estep_pairs.cu', 'estep_pairs_warp.cu', 'estep_pairs_strip.cu', 'mstep.cu',
           'snp_aggregate.cu', 'comm.cu']
HEADERS = [CSRC / 'common.cuh', CSRC.parent.parent / 'include' / 'demux_b200.h']

HOST_SRC = CSRC.parent / 'csrc_host' / 'bam_counter.cpp'
HOST_LIB_PATH = CSRC.parent / 'csrc_host' / 'libdemux_io.so'

NVCC_FLAGS = [
    '-std=c++17', '-O3', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '--fmad=true',            # contraction only where the source allows it; parity-critical code uses __f*_rn
    '-Xcompiler', '-fPIC', '-Xcompiler', '-O2',
    '-Xptxas', '-v',
    '-Wno-deprecated-gpu-targets',
]


def find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError('nvcc not found: libdemux_b200.so cannot be built')


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS + [Path(__file__)]
    return any(d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    objects = []
    log_dir = CSRC / 'build'
    log_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = log_dir / (src + '.o')
        cmd = [nvcc, *NVCC_FLAGS, '-c', str(CSRC / src), '-o', str(obj)]
        procs.append((src, obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, cmd, proc in procs:
        out, _ = proc.communicate()
        (log_dir / (src + '.ptxas.log')).write_text(out)
        if verbose:
            print(out)
        if proc.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{" ".join(cmd)}\n{out}')
        objects.append(str(obj))
    link = [nvcc, '-shared', '-o', str(LIB_PATH), *objects, '-gencode', 'arch=compute_100a,code=sm_100a', '-ldl']
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{" ".join(link)}\n{res.stdout}')
    return LIB_PATH


def build_host(force: bool = False) -> Path:
    """libdemux_io.so: the native input stage (g++, zlib; no CUDA)."""
    header = CSRC.parent.parent / 'include' / 'demux_io.h'
    if not force and HOST_LIB_PATH.exists() and \
            HOST_LIB_PATH.stat().st_mtime >= max(HOST_SRC.stat().st_mtime, header.stat().st_mtime):
        return HOST_LIB_PATH
    cxx = os.environ.get('CXX') or shutil.which('g++') or shutil.which('c++')
    if not cxx:
        raise RuntimeError('no C++ compiler found: libdemux_io.so cannot be built')
    cmd = [cxx, '-O3', '-std=c++17', '-shared', '-fPIC', '-o', str(HOST_LIB_PATH), str(HOST_SRC), '-lz']
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'host build failed:\n{" ".join(cmd)}\n{res.stdout}')
    return HOST_LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
    print(build_host(force='--force' in sys.argv))
