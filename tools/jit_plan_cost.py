#!/usr/bin/env python
"""Offline cost model of a plan as run by the specialised kernels (no GPU): passes, rounds and the packed FP32
instructions per thread the generated sources contain, by origin.  Used to tune planner options (environment
variables B200Q_XC1_PENALTY, B200Q_MIN_ROUND_GATES, B200Q_DEFER_DIAG, B200Q_FREE_PHASE).

  python tools/jit_plan_cost.py [--nqubit 28] [--depth 40]
"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nqubit', type=int, default=28)
    ap.add_argument('--depth', type=int, default=40)
    ap.add_argument('--chunk-bits', type=int, default=0)
    a = ap.parse_args()
    import torch
    import jit_precompile
    from deepquantum_b200 import circuit as circ
    circ.PLAN_OPTIONS.update(chunk_bits=a.chunk_bits)
    cir = jit_precompile.bench_circuit(a.nqubit, a.depth)
    plan = cir._get_program().plan(torch.complex64)
    tot = dict(rot=0, had=0, rho=0, matz=0, scale=0, other=0)
    rounds = 0
    for i in range(plan.n_passes):
        s = plan.codegen(i)
        body = s[s.index('// ---- round'):s.index('#if defined(__CUDACC__)\nextern')]
        rounds += len(re.findall(r'^// ---- round', body, re.M))
        nfa, nhad, ncm, nsc = (len(re.findall(k + r'\(', body)) for k in ('vfa', 'vhad', 'vcm', 'vsc'))
        tot['rot'] += nfa
        tot['had'] += 2 * nhad
        tot['rho'] += 4 * ncm
        nmz = len(re.findall(r'const V s_ =', body))
        tot['matz'] += 16 * nmz
        tot['scale'] += nsc - 16 * nmz
        tot['other'] += len(re.findall(r'v(fma|mul|add|sub)\(', body))
    total = sum(tot.values())
    print(f'{a.nqubit}q depth {a.depth}: passes {plan.n_passes} rounds {rounds} packed/thread total {total} '
          f'({total / plan.n_passes:.0f} per pass) {tot}')


if __name__ == '__main__':
    main()
