#!/usr/bin/env python
"""Static SASS statistics of the specialised pass kernels of a benchmark circuit (no GPU needed): compiles the
kernels into a scratch cache directory and prints the opcode mix summed over all passes -- the generated code is
straight-line inside the tile loop, so static counts track executed instructions per thread and tile.

  python tools/jit_sass_stats.py [--nqubit 28] [--depth 40]
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nqubit', type=int, default=28)
    ap.add_argument('--depth', type=int, default=40)
    ap.add_argument('--chunk-bits', type=int, default=0)
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix='b200q_sass_')
    os.environ['B200Q_JIT_CACHE'] = d
    import torch
    from deepquantum_b200 import circuit as circ
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import jit_precompile
    circ.PLAN_OPTIONS.update(chunk_bits=a.chunk_bits)
    cir = jit_precompile.bench_circuit(a.nqubit, a.depth)
    plan = cir._get_program().plan(torch.complex64)
    ok = plan.compile(0)
    tot = collections.Counter()
    regs = []
    for f in sorted(os.listdir(d)):
        if not f.endswith('.cubin'):
            continue
        path = os.path.join(d, f)
        sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
        for line in sass.splitlines():
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if m:
                tot[m.group(2).split('.')[0]] += 1
        res = subprocess.run(['cuobjdump', '-res-usage', path], capture_output=True, text=True).stdout
        m = re.search(r'REG:(\d+) STACK:(\d+)', res)
        if m:
            regs.append((int(m.group(1)), int(m.group(2))))
    n = sum(tot.values())
    print(f'{plan.n_passes} passes, {ok} kernels, {n} SASS instructions ({n / max(1, ok):.0f} per pass); '
          f'registers max {max(r for r, _ in regs)}, kernels with stack {sum(1 for _, s in regs if s)}')
    for op, c in tot.most_common(18):
        print(f'  {op:10s} {c:7d}  {100.0 * c / n:5.1f} %')


if __name__ == '__main__':
    main()
