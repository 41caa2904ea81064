#!/usr/bin/env python
"""Dense 5- and 6-target gates (complex64): the tensor-core contraction (b200q_dense_tc_apply: tcgen05.mma, 3-product
TF32 split) against the CUDA-core dense pass on the same state -- time per gate, achieved HBM fraction, and the
rel-L2 difference of each against a complex128 host contraction at 20 qubits.

  python tools/dense_tc_bench.py [--nqubit 28]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from deepquantum_b200 import _lib as L  # noqa: E402
from deepquantum_b200 import engine  # noqa: E402


def rand_u(k, rng):
    q, _ = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
    return q


def tc_apply(st, n, u, targets, ctrl=0, adjoint=0):
    t = (C.c_int32 * len(targets))(*targets)
    L.check(L.load().b200q_dense_tc_apply(st.data_ptr(), n, u.data_ptr(), t, len(targets), ctrl, adjoint,
                                          torch.cuda.current_stream().cuda_stream))


def core_apply(st, n, u, targets, controls=()):
    os.environ['B200Q_DENSE_TC'] = '0'
    try:
        engine.apply_gate_(st, n, u, targets, controls)
    finally:
        os.environ.pop('B200Q_DENSE_TC', None)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nqubit', type=int, default=28)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    import statevec_oracle as so
    # ---- accuracy at 20 qubits against the complex128 host contraction
    n = 20
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    for k, wires, ctrl_wires in ((6, [3, 19, 0, 7, 11, 15], []), (5, [18, 2, 9, 5, 12], [0]), (4, [1, 10, 19, 6], [])):
        u = rand_u(k, rng)
        if ctrl_wires:
            ref = so.evolve_state_controlled(psi.reshape(1, -1), u, n, wires, ctrl_wires).reshape(-1)
        else:
            ref = so.evolve_state(psi.reshape(1, -1), u, n, wires).reshape(-1)
        targets = [n - 1 - w for w in reversed(wires)]
        ctrl = sum(1 << (n - 1 - c) for c in ctrl_wires)
        ud = torch.tensor(u, dtype=torch.complex64, device='cuda').reshape(-1).contiguous()
        st_tc = torch.tensor(psi, dtype=torch.complex64, device='cuda')
        tc_apply(st_tc, n, ud, targets, ctrl)
        out = st_tc.cpu().numpy().astype(np.complex128)
        row = {'check': f'{k} targets, 20 qubits', 'tc_vs_c128_host': float(np.linalg.norm(out - ref) / np.linalg.norm(ref))}
        if k >= 5:
            st_cc = torch.tensor(psi, dtype=torch.complex64, device='cuda')
            core_apply(st_cc, n, ud, targets, [n - 1 - c for c in ctrl_wires])
            oc = st_cc.cpu().numpy().astype(np.complex128)
            row['cuda_core_vs_c128_host'] = float(np.linalg.norm(oc - ref) / np.linalg.norm(ref))
        print(json.dumps(row), flush=True)
    # ---- speed at --nqubit
    n = a.nqubit
    st = torch.zeros(2**n, dtype=torch.complex64, device='cuda')
    st[0] = 1
    peak = 6551.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    for k in (6, 5):
        u = torch.tensor(rand_u(k, rng), dtype=torch.complex64, device='cuda').reshape(-1).contiguous()
        targets = [n - 1, 3, n // 2, 7, 1, n - 4][:k]
        ms_tc = timed(lambda: tc_apply(st, n, u, targets))
        ms_cc = timed(lambda: core_apply(st, n, u, targets))
        gb = 2 * (2**n) * 8 / 1e9
        print(json.dumps({'bench': f'{k} targets, {n} qubits', 'ms_tensor_core': ms_tc, 'ms_cuda_core': ms_cc,
                          'frac_hbm_tensor_core': gb / (ms_tc * 1e-3) / peak,
                          'frac_hbm_cuda_core': gb / (ms_cc * 1e-3) / peak}), flush=True)


if __name__ == '__main__':
    main()
