"""Builds the TEST-ONLY host emulator of the tile kernel body (g++, no CUDA needed)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, '_build')
LIB = os.path.join(OUT, 'libb200q_hostemu.so')


def build(force=False):
    srcs = [os.path.join(HERE, 'hostemu.cpp'), os.path.join(ROOT, 'deepquantum_b200', 'csrc', 'b200q_planner.cpp')]
    deps = srcs + [os.path.join(ROOT, 'deepquantum_b200', 'csrc', f)
                   for f in ('b200q_tile_body.h', 'b200q_program.h', 'b200q_planner.h', 'b200q_qudit_geom.h')] + [
        os.path.join(ROOT, 'include', 'b200q.h')]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-o', LIB] + srcs
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
