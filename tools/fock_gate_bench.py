"""Per-gate timing of the Fock tensor kernels at config-5 size (8 modes, cutoff 10, complex64, 0.8 GB state):
block-structured kernel (b200q_qudit_apply_structured) against the generic ELL kernel (b200q_qudit_apply), by gate
class and mode position.  One JSON line per case: ms, GB/s of the 1.6 GB a pass has to move, fraction of the copy peak.

    python tools/fock_gate_bench.py [--double]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepquantum_b200 import _lib as L  # noqa: E402
from deepquantum_b200 import photonic as ph  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        return 6551.0


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    double = '--double' in sys.argv
    n, d = 8, 10
    cdt = torch.complex128 if double else torch.complex64
    st = torch.randn(1, d**n, dtype=cdt, device='cuda')
    st /= st.norm()
    nbytes = 2 * d**n * (16 if double else 8)
    pk = peak()
    cases = [('squeezer', ph.Squeezing([0.3, 1.0], n, [w], d)) for w in (0, 4, 7)]
    cases += [('phase shifter', ph.PhaseShift(0.7, n, [w], d)) for w in (0, 7)]
    cases += [('beamsplitter', ph.BeamSplitter([0.3, 1.0], n, list(w), d)) for w in ((0, 1), (3, 4), (4, 3), (1, 6), (6, 7), (0, 7))]
    cases += [('two-mode squeezer', ph.Squeezing2([0.2, 1.0], n, list(w), d)) for w in ((3, 4), (6, 7))]
    for name, op in cases:
        m = op.update_matrix_state().reshape(d**len(op.wires), -1).to(cdt).to('cuda')
        row = {'gate': name, 'modes': op.wires, 'dtype': 'c128' if double else 'c64'}
        for label, structure in (('structured', op._structure), ('generic', L.QUDIT_GENERAL)):
            ms = timed(lambda: ph.qudit_apply_(st, n, d, m, op.wires, 1, structure))
            row[f'ms_{label}'] = round(ms, 4)
            row[f'frac_{label}'] = round(nbytes / (ms * 1e-3) / 1e9 / pk, 3)
        print(json.dumps(row), flush=True)
    # config 5 as a whole: matrix assembly (batched torch evaluation of the Fock matrices) against the 36 gate passes
    import deepquantum_b200 as dq
    from deepquantum_b200 import workloads as wl
    cir = dq.QumodeCircuit(n, 'vac', cutoff=d, backend='fock', basis=False)
    for e in wl.fock_interferometer_spec(n):
        if e['g'] == 's':
            cir.s(e['w'][0], e['p'][0], e['p'][1])
        else:
            cir.bs(e['w'], e['p'])
    cir.to('cuda', torch.float64 if double else torch.float32)
    with torch.no_grad():
        ms_build = timed(lambda: cir.build_matrices(cdt, 'cuda'))
        mats = cir.build_matrices(cdt, 'cuda')
        flat = cir.init_state.state.reshape(1, -1).to(cdt).contiguous().clone()

        def gates_only():
            for op, m in zip(cir.operators, mats):
                ph.qudit_apply_(flat, n, d, m, op.wires, 1, op._structure)
        ms_gates = timed(gates_only)
        ms_all = timed(lambda: cir())
    print(json.dumps({'config5': f'{n} modes x cutoff {d}', 'dtype': 'c128' if double else 'c64', 'gates': len(cir.operators),
                      'ms_forward': round(ms_all, 3), 'ms_build_matrices': round(ms_build, 3),
                      'ms_gate_passes': round(ms_gates, 3),
                      'frac_gate_passes': round(len(cir.operators) * nbytes / (ms_gates * 1e-3) / 1e9 / pk, 3)}), flush=True)


def mesh():
    """Clements mesh of (phase shifter, beamsplitter) pairs -- the unit cell of a programmable interferometer -- at
    config-5 size: 28 pairs = 56 gates.  With the diagonal neighbours folded into the beamsplitter blocks: 28 passes."""
    import deepquantum_b200 as dq
    n, d = 8, 10
    cir = dq.QumodeCircuit(n, [1, 0, 1, 0, 1, 0, 1, 0], cutoff=d, backend='fock', basis=False)
    g = torch.Generator().manual_seed(1)
    for layer in range(n):
        for a in range(layer % 2, n - 1, 2):
            cir.ps(a, float(torch.rand(1, generator=g) * 6))
            cir.bs([a, a + 1], [float(torch.rand(1, generator=g) * 6), float(torch.rand(1, generator=g) * 6)])
    cir.to('cuda')
    with torch.no_grad():
        ms = timed(lambda: cir())
    st = cir.fock_plan_stats()
    print(json.dumps({'mesh': 'Clements, (ps, bs) pairs, 8 modes x cutoff 10', 'gates': st['gates'], 'passes': st['passes'],
                      'ms_forward': round(ms, 3), 'group': os.environ.get('B200Q_FOCK_GROUP', '1'),
                      'fold': os.environ.get('B200Q_FOCK_FOLD', '1')}), flush=True)


if __name__ == '__main__':
    if '--mesh' in sys.argv:
        mesh()
    else:
        main()
