#!/usr/bin/env python
"""Generate and compile (NVRTC, sm_100a -- no GPU needed) the specialised pass kernels of the benchmark circuits,
so that the cubin cache next to the library (deepquantum_b200/lib/jit_cache/) is warm when bench.py runs.

  python tools/jit_precompile.py [--nqubit 30] [--depth 40] [--dump-dir DIR]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import deepquantum_b200 as dq  # noqa: E402
from deepquantum_b200 import workloads as wl  # noqa: E402


def bench_circuit(n, depth):
    spec = wl.random_clifford_rx_spec(n, depth)
    cir = dq.QubitCircuit(n)
    for e in spec:
        if e['g'] == 'rx':
            cir.rx(e['w'][0], encode=True)
        else:
            wl.apply_spec(cir, [e])
    cir.observable([0], 'z')
    cir.observable([n // 2, n - 1], 'zz')
    return cir


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nqubit', type=int, default=30)
    ap.add_argument('--depth', type=int, default=40)
    ap.add_argument('--dump-dir', default='')
    ap.add_argument('--threads', type=int, default=0)
    a = ap.parse_args()
    cir = bench_circuit(a.nqubit, a.depth)
    prog = cir._get_program()
    plan = prog.plan(torch.complex64)
    if a.dump_dir:
        os.makedirs(a.dump_dir, exist_ok=True)
        for i in range(plan.n_passes):
            with open(os.path.join(a.dump_dir, f'pass{i:02d}.cu'), 'w') as f:
                f.write(plan.codegen(i))
    t0 = time.time()
    ok = plan.compile(a.threads)
    print(f'{a.nqubit} qubits depth {a.depth}: {plan.n_passes} passes, {ok} specialised kernels compiled in '
          f'{time.time() - t0:.1f} s; status {plan.jit_status()}')


if __name__ == '__main__':
    main()
