"""Builds lib/libb200q.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

Every source is compiled to its own object under lib/obj/ (in parallel, only when it or one of the headers it
includes changed) and the objects are linked into the shared library: a change to the kernel generator or the
planner does not recompile the three-minute tile-kernel translation unit."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(LIBDIR, 'obj')
LIB = os.path.join(LIBDIR, 'libb200q.so')
PUBLIC = os.path.join(os.path.dirname(HERE), 'include', 'b200q.h')
_COMMON = ['b200q_program.h', 'b200q_planner.h']
# source -> headers of csrc/ it includes (besides include/b200q.h)
SOURCES = {
    'b200q_lib.cu': _COMMON + ['b200q_tile_body.h', 'b200q_jit.h', 'b200q_codegen.h'],
    'b200q_qudit.cu': ['b200q_qudit_geom.h'],
    'b200q_qudit_fused.cu': [],
    'b200q_qudit_sector.cu': ['b200q_qudit_geom.h'],
    'b200q_dense_tc.cu': [],
    'b200q_sample.cu': [],
    'b200q_planner.cpp': _COMMON,
    'b200q_codegen.cpp': _COMMON + ['b200q_codegen.h'],
    'b200q_jit.cpp': _COMMON + ['b200q_codegen.h', 'b200q_jit.h'],
}
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _obj(src: str) -> str:
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + '.o')


def _stale(src: str) -> bool:
    obj = _obj(src)
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [os.path.join(CSRC, src), PUBLIC] + [os.path.join(CSRC, h) for h in SOURCES[src]]
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    return any(_stale(s) or os.path.getmtime(_obj(s)) > os.path.getmtime(LIB) for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    todo = [s for s in SOURCES if force or _stale(s)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o',
                                                                              _obj(src)]
        if src == 'b200q_lib.cu':
            cmd += ['--threads', '4']
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, todo))
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC'] + [
        _obj(s) for s in SOURCES] + ['-o', LIB, '-ldl']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
