"""Density-matrix workload on one GPU: n-qubit noisy Clifford+RX circuit (one channel per qubit per layer) run as
a 2n-qubit amplitude vector.  One JSON line per size: passes, ms, fraction of the HBM roofline per pass."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepquantum_b200 as dq  # noqa: E402
from deepquantum_b200 import workloads as wl  # noqa: E402
from tools.bench_configs import PEAK, timed  # noqa: E402


def run(n, depth, channels=None):
    spec = wl.noisy_circuit_spec(n, depth, channels=channels)
    cir = dq.QubitCircuit(n, den_mat=True)
    wl.apply_spec(cir, spec)
    for q in range(n - 1):
        cir.observable([q, q + 1], 'z')
    cir.to('cuda')
    prog = cir._get_program()
    plan = prog.plan(torch.complex64)
    with torch.no_grad():
        ms = timed(lambda: cir(), reps=3, warm=2)
        ms_e = timed(lambda: cir.expectation(), reps=3, warm=1)
        mats = prog.low.build_matrices(torch.complex64, 'cuda')
        buf = torch.empty(1, 4**n, dtype=torch.complex64, device='cuda')
        dq.engine.init_basis_(buf, 2 * n, 1, 0)
        ms_k = timed(lambda: plan.run(buf, mats, 1, 0), reps=3, warm=1)        # kernels only (state not reset)
        ms_m = timed(lambda: prog.low.build_matrices(torch.complex64, 'cuda'), reps=3, warm=1)
        rho = cir.state
        tr = rho.diagonal().sum().real.item()
    bytes_pass = 2 * (4**n) * 8
    from deepquantum_b200.operation import DenMatLowering
    print(json.dumps({'config': f'noisy Clifford+RX, {n} qubits (rho = {2 * n}-qubit vector), depth {depth}, complex64',
                      'channels': list(channels) if channels else 'all seven', 'pauli_bell': DenMatLowering.PAULI_BELL, 'damping_svd': DenMatLowering.DAMPING_SVD,
                      'source_ops': len(spec), 'kernel_gate_records': len(prog.structs), 'passes': plan.n_passes,
                      'ms_forward': ms, 'ms_kernels': ms_k, 'ms_matrix_build': ms_m, 'ms_per_pass': ms_k / plan.n_passes,
                      'ops_per_s': len(spec) / ms * 1e3,
                      'frac_of_hbm_per_pass': plan.n_passes * bytes_pass / (ms_k * 1e-3) / PEAK,
                      'ms_expectation': ms_e, 'trace': tr}), flush=True)


if __name__ == '__main__':
    sizes = [int(a) for a in sys.argv[1:]] or [12, 14]
    for n in sizes:
        run(n, 10)
        if os.environ.get('DENMAT_ONLY_MIXED') != '1':
            run(n, 10, channels=('depolarizing',))
