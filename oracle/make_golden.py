#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(TuringQ/deepquantum, /root/reference/src) on CPU in the build container.

Usage:  python oracle/make_golden.py            (needs /root/reference; not run on the GPU box)

TEST INFRASTRUCTURE ONLY.  The fixtures are committed; the tests never need the reference again.
Everything is seeded, but the values depend on the torch CPU kernels only to rounding.

Files written
  gate_matrices.npz   the local matrix of every gate family at fixed parameters (complex128)
  circuits.npz        final states of seeded circuits (specs stored as JSON) in c128 and c64
  qaoa.npz            QAOA MaxCut loss + gradient (reference autograd) at n = 6, 8, 10
  fock.npz            Fock tensor backend (squeezer / beamsplitter / phase shifter) final states
  measure.npz         qmath.measure(with_prob=True): states, measured wires and the probabilities the reference
                      attaches to every outcome it drew (all wires, wire subsets, batched states)
  denmat.npz          density-matrix circuits (den_mat=True): every gate family + the seven channels, final rho in
                      c128 and c64, Pauli-string expectations, measure(with_prob=True) probabilities
  hamiltonian.npz     circuits with HamiltonianGate blocks (Pauli-sum and matrix form, controlled), states and rho
  qasm3.npz           OpenQASM 3.0 programs (def / ctrl @ / pow() @ / all stdgates names), the reference's import of
                      them run to a final state, and the reference's export of a seeded circuit
  fock2.npz           Fock tensor path with the beamsplitter family (mzi, bs_theta, bs_phi, bs_rx, bs_ry, bs_h, dc, h),
                      rotations (r, f) and Kerr gates (k, ck): final states and the local Fock matrices
  dist_adjoint.npz    the reference's own test circuit of the differentiable sharded expectation
                      (tests/test_circuit.py:87-139), dense autograd: expectation values and d/d(data), n = 4 and 6
  unitary.npz         QubitCircuit.get_unitary() of the all-gate-families circuit (5 qubits) and a random circuit
  reset.npz           circuits with Reset (postselect 0 / 1, several wires, a wire whose outcome has probability 0, all
                      wires, 2-D data): final states in c64 (the reference cannot run them in c128)
  dense_grad.npz      trainable / data-fed dense blocks on 3 and 4 wires (HamiltonianGate, LatentGate, one controlled):
                      loss and its reference-autograd gradient w.r.t. every parameter and the data, n = 6
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from ref_loader import load_reference  # noqa: E402

from deepquantum_b200 import workloads as wl  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
dq = load_reference()
torch.manual_seed(1234)


def all_gates_spec(n=5, seed=7):
    """Every gate family of circuit.py:899-1537, with controls on both sides of the targets."""
    g = torch.Generator().manual_seed(seed)

    def r(k=1):
        return (torch.rand(k, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()

    def runitary(k):
        a = torch.randn(2**k, 2**k, generator=g, dtype=torch.float64) + 1j * torch.randn(
            2**k, 2**k, generator=g, dtype=torch.float64)
        q, _ = torch.linalg.qr(a)
        return q.real.tolist(), q.imag.tolist()

    spec = [{'g': 'hlayer'}]
    for i, name in enumerate(['x', 'y', 'z', 'h', 's', 'sdg', 't', 'tdg']):
        spec.append({'g': name, 'w': [i % n]})
        spec.append({'g': name, 'w': [(i + 2) % n], 'c': [(i + 4) % n]})
    spec.append({'g': 'x', 'w': [2], 'c': [0, 4]})
    spec.append({'g': 'h', 'w': [0], 'c': [1, 3, 4]})
    for name in ['rx', 'ry', 'rz', 'p']:
        spec.append({'g': name, 'w': [1], 'p': r()})
        spec.append({'g': name, 'w': [4], 'c': [2], 'p': r()})
        spec.append({'g': name, 'w': [0], 'c': [3, 1], 'p': r()})
    spec.append({'g': 'u3', 'w': [3], 'p': r(3)})
    spec.append({'g': 'u3', 'w': [2], 'c': [4], 'p': r(3)})
    for plane in ['xy', 'yz', 'zx']:
        spec.append({'g': 'j', 'w': [1], 'p': r(), 'plane': plane})
    for name in ['cx', 'cy', 'cz', 'ch', 'cs', 'csdg', 'ct', 'ctdg', 'cnot']:
        spec.append({'g': name, 'w': [3, 0]})
        spec.append({'g': name, 'w': [1, 4]})
    for name in ['crx', 'cry', 'crz', 'cp']:
        spec.append({'g': name, 'w': [4, 2], 'p': r()})
    spec.append({'g': 'cu', 'w': [0, 3], 'p': r(3)})
    for name in ['swap', 'iswap']:
        spec.append({'g': name, 'w': [0, 3]})
        spec.append({'g': name, 'w': [4, 1], 'c': [2]})
    for name in ['rxx', 'ryy', 'rzz', 'rxy', 'rbs']:
        spec.append({'g': name, 'w': [1, 3], 'p': r()})
        spec.append({'g': name, 'w': [4, 0], 'p': r()})
        spec.append({'g': name, 'w': [2, 1], 'c': [4], 'p': r()})
    for name in ['crxx', 'cryy', 'crzz', 'crxy']:
        spec.append({'g': name, 'w': [2, 4, 0], 'p': r()})
    spec.append({'g': 'toffoli', 'w': [0, 2, 4]})
    spec.append({'g': 'toffoli', 'w': [4, 1, 0]})
    spec.append({'g': 'ccx', 'w': [3, 1, 2]})
    spec.append({'g': 'fredkin', 'w': [1, 3, 0]})
    spec.append({'g': 'cswap', 'w': [4, 0, 2]})
    for k, w in [(1, [2]), (2, [3, 1]), (3, [4, 0, 2]), (4, [1, 3, 0, 4])]:
        re, im = runitary(k)
        spec.append({'g': 'any', 'w': w, 'u_re': re, 'u_im': im})
    re, im = runitary(2)
    spec.append({'g': 'any', 'w': [0, 2], 'c': [4], 'u_re': re, 'u_im': im})
    spec += [{'g': 'xlayer', 'w': [0, 2]}, {'g': 'ylayer'}, {'g': 'zlayer', 'w': [1]}]
    spec += [{'g': 'rxlayer', 'p': r(n)}, {'g': 'rylayer', 'w': [0, 3], 'p': r(2)}, {'g': 'rzlayer', 'p': r(n)}]
    spec.append({'g': 'u3layer', 'w': [1, 4], 'p': r(6)})
    spec.append({'g': 'cxlayer', 'pairs': [[0, 1], [4, 2]]})
    spec.append({'g': 'cnot_ring'})
    spec.append({'g': 'cnot_ring', 'minmax': [1, 4], 'step': 2, 'reverse': True})
    return spec


def run_ref(spec, n, double, init=None):
    cir = dq.QubitCircuit(n) if init is None else dq.QubitCircuit(n, init_state=init)
    wl.apply_spec(cir, spec, torch.complex128 if double else torch.complex64)
    if double:
        cir.to(torch.double)
    with torch.no_grad():
        out = cir()
    return out.reshape(out.shape[0], -1).numpy() if out.ndim == 3 else out.reshape(-1).numpy()


# ----------------------------------------------------------------------------------------------
def gate_matrices():
    th = [0.37, 1.21, -2.05]
    out = {}
    for name, cls in [('x', dq.PauliX), ('y', dq.PauliY), ('z', dq.PauliZ), ('h', dq.Hadamard), ('s', dq.SGate),
                      ('sdg', dq.SDaggerGate), ('t', dq.TGate), ('tdg', dq.TDaggerGate), ('cnot', dq.CNOT),
                      ('swap', dq.Swap), ('iswap', dq.ImaginarySwap), ('toffoli', dq.Toffoli),
                      ('fredkin', dq.Fredkin)]:
        out[name] = cls().to(torch.double).matrix.numpy()
    for name, cls in [('rx', dq.Rx), ('ry', dq.Ry), ('rz', dq.Rz), ('p', dq.PhaseShift), ('rxx', dq.Rxx),
                      ('ryy', dq.Ryy), ('rzz', dq.Rzz), ('rxy', dq.Rxy), ('rbs', dq.ReconfigurableBeamSplitter)]:
        gate = cls(inputs=th[0]).to(torch.double)
        out[name] = gate.update_matrix().detach().numpy()
        out[name + '_inv'] = gate.inverse().update_matrix().detach().numpy()
    u3 = dq.U3Gate(inputs=th).to(torch.double)
    out['u3'] = u3.update_matrix().detach().numpy()
    out['u3_inv'] = u3.inverse().update_matrix().detach().numpy()
    for plane in ['xy', 'yz', 'zx']:
        out['j_' + plane] = dq.gate.ProjectionJ(inputs=th[0], plane=plane).to(torch.double).update_matrix().numpy()
    out['theta'] = np.array(th)
    np.savez_compressed(os.path.join(OUT, 'gate_matrices.npz'), **out)
    print('gate_matrices', len(out))


def circuits():
    cases = {}
    cases['all_gates_n5'] = (5, all_gates_spec(5, 7))
    cases['all_gates_n5_b'] = (5, all_gates_spec(5, 11)[::-1])
    cases['c1_plumbing_n12'] = (12, wl.c1_plumbing_spec(12))
    for n, depth in [(3, 6), (6, 8), (9, 10), (12, 12), (14, 40)]:
        cases[f'random_n{n}_d{depth}'] = (n, wl.random_clifford_rx_spec(n, depth))
    cases['random_cx_n10_d10'] = (10, wl.random_clifford_rx_spec(10, 10, seed=5, two_qubit='cx'))
    out = {}
    for name, (n, spec) in cases.items():
        out[name + '/spec'] = np.array(json.dumps({'n': n, 'spec': spec}))
        out[name + '/c128'] = run_ref(spec, n, True)
        out[name + '/c64'] = run_ref(spec, n, False)
        print(name, n, len(spec), float(np.linalg.norm(out[name + '/c128'])))
    # batched initial state + non-trivial initial state (circuit.py:199-226)
    g = torch.Generator().manual_seed(99)
    n = 6
    init = torch.randn(3, 2**n, generator=g, dtype=torch.float64) + 1j * torch.randn(3, 2**n, generator=g,
                                                                                      dtype=torch.float64)
    init = init / init.norm(dim=-1, keepdim=True)
    spec = wl.random_clifford_rx_spec(n, 5, seed=3)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    cir.to(torch.double)
    with torch.no_grad():
        res = cir(state=init.unsqueeze(-1))
    out['batched_n6/spec'] = np.array(json.dumps({'n': n, 'spec': spec}))
    out['batched_n6/init'] = init.numpy()
    out['batched_n6/c128'] = res.reshape(3, -1).numpy()
    np.savez_compressed(os.path.join(OUT, 'circuits.npz'), **out)


def qaoa():
    out = {}
    for n, p in [(6, 2), (8, 4), (10, 3)]:
        edges, weights, layout = wl.qaoa_maxcut_structure(n, p, seed=wl.SEED + n)
        cir = dq.QubitCircuit(n)
        wl.build_qaoa(cir, edges, p, barriers=False)  # reference Barrier breaks .to(double)
        cir.to(torch.double)
        params = torch.tensor([0.1 + 0.05 * k for k in range(p)] + [1.0 - 0.1 * k for k in range(p)],
                              dtype=torch.float64, requires_grad=True)
        data = wl.qaoa_data(params, weights, layout)
        state = cir(data)
        exp = cir.expectation()
        w = torch.tensor(weights, dtype=torch.float64)
        loss = 0.5 * (w * (exp.reshape(-1) - 1)).sum()  # examples/qaoa.py:55-57 (negative cut value)
        loss.backward()
        out[f'n{n}_p{p}/meta'] = np.array(json.dumps({'n': n, 'p': p, 'edges': edges, 'weights': weights,
                                                       'seed': wl.SEED + n}))
        out[f'n{n}_p{p}/params'] = params.detach().numpy()
        out[f'n{n}_p{p}/state'] = state.detach().reshape(-1).numpy()
        out[f'n{n}_p{p}/expectation'] = exp.detach().reshape(-1).numpy()
        out[f'n{n}_p{p}/loss'] = loss.detach().numpy()
        out[f'n{n}_p{p}/grad'] = params.grad.numpy()
        print('qaoa', n, p, float(loss), params.grad.tolist())
    np.savez_compressed(os.path.join(OUT, 'qaoa.npz'), **out)


def fock():
    out = {}
    for nmode, cutoff in [(2, 5), (3, 4), (4, 6), (5, 8)]:
        spec = wl.fock_interferometer_spec(nmode, seed=wl.SEED + nmode)
        spec.append({'g': 'ps', 'w': [0], 'p': [0.77]})
        for double in (True, False):
            cir = dq.QumodeCircuit(nmode, 'vac', cutoff=cutoff, backend='fock', basis=False)
            for e in spec:
                if e['g'] == 's':
                    cir.s(e['w'][0], r=e['p'][0], theta=e['p'][1])
                elif e['g'] == 'bs':
                    cir.bs(e['w'], e['p'])
                elif e['g'] == 'ps':
                    cir.ps(e['w'][0], e['p'][0])
            if double:
                cir.to(torch.double)
            with torch.no_grad():
                st = cir()
            out[f'm{nmode}_c{cutoff}/' + ('c128' if double else 'c64')] = st.reshape(-1).numpy()
        out[f'm{nmode}_c{cutoff}/spec'] = np.array(json.dumps({'nmode': nmode, 'cutoff': cutoff, 'spec': spec}))
        # local gate matrices (photonic/gate.py:192-194, 347-374, 1091-1114) for the oracle pin
        s = dq.photonic.Squeezing(inputs=[0.31, 1.3], nmode=1, wires=[0], cutoff=cutoff).to(torch.double)
        out[f'm{nmode}_c{cutoff}/s_matrix'] = s.update_matrix_state().detach().numpy()
        b = dq.photonic.BeamSplitter(inputs=[0.6, 2.2], nmode=2, wires=[0, 1], cutoff=cutoff).to(torch.double)
        out[f'm{nmode}_c{cutoff}/bs_matrix'] = b.update_matrix_state().detach().numpy()
        print('fock', nmode, cutoff, float(np.linalg.norm(out[f'm{nmode}_c{cutoff}/c128'])))
    np.savez_compressed(os.path.join(OUT, 'fock.npz'), **out)


FOCK2_SPEC = [
    {'g': 's', 'w': [0], 'p': [0.31, 0.4]}, {'g': 's', 'w': [1], 'p': [0.22, 1.3]}, {'g': 's', 'w': [2], 'p': [0.4, 2.2]},
    {'g': 'mzi', 'w': [0, 1], 'p': [0.3, 0.4]}, {'g': 'mzi', 'w': [1, 2], 'p': [1.3, 0.7], 'phi_first': False},
    {'g': 'bs_theta', 'w': [0, 1], 'p': [0.5]}, {'g': 'bs_phi', 'w': [1, 2], 'p': [0.6]},
    {'g': 'bs_rx', 'w': [2, 0], 'p': [0.7]}, {'g': 'bs_ry', 'w': [1, 2], 'p': [0.8]}, {'g': 'bs_h', 'w': [0, 1], 'p': [0.9]},
    {'g': 'dc', 'w': [1, 2]}, {'g': 'h', 'w': [0, 2]}, {'g': 'r', 'w': [0], 'p': [0.3]},
    {'g': 'r', 'w': [1], 'p': [0.45], 'inv_mode': True}, {'g': 'f', 'w': [2]}, {'g': 'k', 'w': [0], 'p': [0.2]},
    {'g': 'ck', 'w': [1, 2], 'p': [0.1]}, {'g': 'bs', 'w': [0, 2], 'p': [0.6, 2.2]}, {'g': 'ck', 'w': [2, 0], 'p': [0.35]},
    {'g': 'd', 'w': [1], 'p': [0.25, 0.9]}, {'g': 'd', 'w': [2], 'p': [0.4, -1.7]}, {'g': 'bs_rx', 'w': [1, 2], 'p': [1.1]},
    {'g': 'd', 'w': [0], 'p': [0.15, 0.0]}, {'g': 's2', 'w': [0, 1], 'p': [0.2, 0.8]}, {'g': 's2', 'w': [2, 0], 'p': [0.12, -2.0]},
    {'g': 'bs', 'w': [1, 2], 'p': [1.0, 0.3]}]


def apply_fock_spec(cir, spec):
    for e in spec:
        g, w, prm = e['g'], e['w'], e.get('p', [])
        if g in ('s', 'd'):
            getattr(cir, g)(w[0], r=prm[0], theta=prm[1])
        elif g == 's2':
            cir.s2(w, r=prm[0], theta=prm[1])
        elif g in ('bs', 'mzi'):
            getattr(cir, g)(w, prm, **({'phi_first': e['phi_first']} if 'phi_first' in e else {}))
        elif g in ('bs_theta', 'bs_phi', 'bs_rx', 'bs_ry', 'bs_h', 'ck'):
            getattr(cir, g)(w, prm[0])
        elif g in ('dc', 'h'):
            getattr(cir, g)(w)
        elif g == 'r':
            cir.r(w[0], prm[0], inv_mode=e.get('inv_mode', False))
        elif g == 'f':
            cir.f(w[0])
        elif g in ('ps', 'k'):
            getattr(cir, g)(w[0], prm[0])
        else:
            raise ValueError(g)


def fock2():
    out = {}
    for nmode, cutoff in [(3, 4), (3, 6)]:
        for double in (True, False):
            cir = dq.QumodeCircuit(nmode, 'vac', cutoff=cutoff, backend='fock', basis=False)
            apply_fock_spec(cir, FOCK2_SPEC)
            if double:
                cir.to(torch.double)
            with torch.no_grad():
                st = cir()
            out[f'm{nmode}_c{cutoff}/' + ('c128' if double else 'c64')] = st.reshape(-1).numpy()
            if double:
                for i, op in enumerate(cir.operators):
                    out[f'm{nmode}_c{cutoff}/mat{i}'] = op.update_matrix_state().detach().numpy()
        print('fock2', nmode, cutoff, float(np.linalg.norm(out[f'm{nmode}_c{cutoff}/c128'])))
    out['spec'] = np.array(json.dumps(FOCK2_SPEC))
    np.savez_compressed(os.path.join(OUT, 'fock2.npz'), **out)


def measure():
    """Reference qmath.measure (qmath.py:568-638) with with_prob=True: the counts are random, the attached
    probabilities and the key convention (sorted wires, wire 0 first) are what the oracle is pinned to."""
    out = {}
    g = torch.Generator().manual_seed(99)
    cases = [('all5', 5, None, 1), ('sub6', 6, [4, 1, 3], 1), ('one7', 7, 2, 1), ('batch5', 5, [0, 3], 3)]
    for name, n, wires, batch in cases:
        st = torch.randn(batch, 2**n, generator=g, dtype=torch.float64) + 1j * torch.randn(
            batch, 2**n, generator=g, dtype=torch.float64)
        st = st / st.norm(dim=-1, keepdim=True)
        arg = st[0] if batch == 1 else st
        res = dq.qmath.measure(arg, shots=4096, with_prob=True, wires=wires)
        res = [res] if batch == 1 else res
        out[name + '/state'] = st.numpy()
        out[name + '/meta'] = json.dumps({'n': n, 'wires': wires, 'batch': batch})
        for b, d in enumerate(res):
            keys = sorted(d)
            out[f'{name}/keys{b}'] = np.array(keys)
            out[f'{name}/counts{b}'] = np.array([d[k][0] for k in keys])
            out[f'{name}/probs{b}'] = np.array([float(d[k][1]) for k in keys])
    np.savez_compressed(os.path.join(OUT, 'measure.npz'), **out)
    print('measure.npz:', len(out), 'arrays')


def hamiltonian_spec(n=6, seed=21):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(4, 4, generator=g, dtype=torch.float64) + 1j * torch.randn(4, 4, generator=g, dtype=torch.float64)
    h2 = (a + a.mH) / 2
    a = torch.randn(8, 8, generator=g, dtype=torch.float64) + 1j * torch.randn(8, 8, generator=g, dtype=torch.float64)
    h3 = (a + a.mH) / 2
    return [{'g': 'hlayer'},
            {'g': 'hamiltonian', 'ham': [[0.5, 'x0y1'], [-1, 'z2y1']], 'p': [0.7]},
            {'g': 'rx', 'w': [3], 'p': [1.1]},
            {'g': 'hamiltonian', 'w': [4, 1], 'c': [5], 'p': [0.37], 'h_re': h2.real.tolist(), 'h_im': h2.imag.tolist()},
            {'g': 'cnot', 'w': [5, 0]},
            {'g': 'hamiltonian', 'ham': [[0.3, 'Z2Z5'], [0.8, 'x3'], [-0.25, 'y4z2']], 'p': [1.9]},
            {'g': 'hamiltonian', 'w': [0, 5, 3], 'p': [0.21], 'h_re': h3.real.tolist(), 'h_im': h3.imag.tolist()},
            {'g': 'hamiltonian', 'ham': [1.5, 'y3'], 'c': [0, 1], 'p': [0.4]}]


def hamiltonian():
    out = {}
    n, spec = 6, hamiltonian_spec()
    for double in (True, False):
        tag = 'c128' if double else 'c64'
        out['ham6/' + tag] = run_ref(spec, n, double)
        cir = dq.QubitCircuit(n, den_mat=True)
        wl.apply_spec(cir, spec, torch.complex128 if double else torch.complex64)
        if double:
            cir.to(torch.double)
        with torch.no_grad():
            out['ham6/rho_' + tag] = cir().numpy()
    # the inverse circuit undoes it (t -> -t), reference circuit.py:530-555
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec, torch.complex128)
    cir.to(torch.double)
    with torch.no_grad():
        out['ham6/inv_c128'] = cir.inverse()(state=cir()).reshape(-1).numpy()
    out['ham6/spec'] = np.array(json.dumps({'n': n, 'spec': spec}))
    np.savez_compressed(os.path.join(OUT, 'hamiltonian.npz'), **out)
    print('hamiltonian.npz:', len(out), 'arrays', float(np.linalg.norm(out['ham6/c128'])), abs(out['ham6/inv_c128'][0]))


QASM_PROGRAM = """OPENQASM 3.0;
include "stdgates.inc";
qubit[5] q;
bit[5] c;
def entangle(a, b) x0, x1 {
  rx(a) x0;
  ry(b / 2 + pi / 8) x1;
  cx x0, x1;
  rzz(a * b) x0, x1;
}
def layer(t) a0, a1, a2 {
  entangle(t, 2 * t) a0, a1;
  ctrl @ entangle(-t, 0.5) a2, a1, a0;
  u(t, 0.2, -0.3) a2;
}
h q[0];
h q[1];
h q[2];
h q[3];
h q[4];
x q[1];
y q[2];
z q[3];
s q[0];
sdg q[1];
t q[2];
tdg q[3];
p(0.7) q[4];
rx(pi / 3) q[0];
ry(-0.4) q[1];
rz(1.9) q[2];
swap q[0], q[3];
cx q[4], q[0];
cz q[1], q[2];
ccx q[0], q[1], q[4];
cswap q[2], q[3], q[4];
rxx(0.3) q[0], q[2];
ryy(0.5) q[1], q[3];
layer(0.37) q[3], q[0], q[2];
pow(2) @ entangle(0.2, 0.9) q[4], q[1];
pow(-1) @ t q[0];
pow(-2) @ rz(0.3) q[1];
pow(3) @ s q[2];
ctrl @ ctrl @ ry(0.8) q[0], q[4], q[2];
ctrl @ pow(2) @ rxx(0.15) q[3], q[1], q[2];
pow(0.5) @ x q[3];
pow(-0.25) @ entangle(0.6, 0.1) q[2], q[4];
ctrl @ pow(1.5) @ h q[0], q[1];
c[0] = measure q[0];
c[3] = measure q[3];
"""


def qasm3():
    """Reference qasm3.py: import (:166-472) and export (:117-156).  `inv @` is left out on purpose: the reference
    turns it into a no-op for integer powers (qasm3.py:303-306, 330-331 negate twice) -- see
    deepquantum_b200/qasm3.py.  One statement per line: the reference parses lines, not statements."""
    import importlib
    q3 = importlib.import_module('deepquantum.qasm3')
    out = {'program': np.array(QASM_PROGRAM)}
    cir = q3.qasm3_to_cir(QASM_PROGRAM)
    cir.to(torch.double)      # (a reference circuit holding a Barrier cannot be moved: operation.py:166 / utils.py:47)
    with torch.no_grad():
        out['state_c128'] = cir().reshape(-1).numpy()
    out['n_ops'] = np.array(len(cir.operators))
    out['wires_measure'] = np.array(cir.wires_measure)
    spec = all_gates_spec(5)
    rc = dq.QubitCircuit(5)
    wl.apply_spec(rc, spec)
    rc.measure(wires=[1, 3])
    text = q3.cir_to_qasm3(rc)
    out['export_spec'] = np.array(json.dumps({'n': 5, 'spec': spec}))
    out['export_text'] = np.array(text)
    back = q3.qasm3_to_cir(text)
    back.to(torch.double)
    with torch.no_grad():
        out['export_reimport_state_c128'] = back().reshape(-1).numpy()
    np.savez_compressed(os.path.join(OUT, 'qasm3.npz'), **out)
    print('qasm3.npz: ops', int(out['n_ops']), 'norm', float(np.linalg.norm(out['state_c128'])))


def denmat():
    """Reference density-matrix path (qmath.py:509-540, operation.py:221-262, 594-600, channel.py)."""
    out = {}
    cases = {'noisy3': (3, wl.noisy_circuit_spec(3, 4, seed=wl.SEED + 3)),
             'noisy5': (5, wl.noisy_circuit_spec(5, 6, seed=wl.SEED + 5)),
             'allgates5': (5, all_gates_spec(5) + wl.noisy_circuit_spec(5, 1, seed=wl.SEED + 55)),
             'noisy7': (7, wl.noisy_circuit_spec(7, 5, seed=wl.SEED + 7))}
    obs = [([0], 'z'), ([1], 'x'), ([2], 'y'), ([0, 2], 'zz'), ([0, 1, 2], 'xyz'), ([2, 0], 'yx')]
    for name, (n, spec) in cases.items():
        for double in (True, False):
            cir = dq.QubitCircuit(n, den_mat=True)
            wl.apply_spec(cir, spec, torch.complex128 if double else torch.complex64)
            for w, b in obs:
                cir.observable(w, b)
            if double:
                cir.to(torch.double)
            with torch.no_grad():
                rho = cir()
                exp = cir.expectation()
            tag = 'c128' if double else 'c64'
            out[f'{name}/{tag}'] = rho.numpy()
            out[f'{name}/exp_{tag}'] = exp.numpy()
            if double:
                for key, wires in (('all', None), ('sub', [2, 0])):
                    res = dq.qmath.measure(rho, shots=2048, with_prob=True, wires=wires, den_mat=True)
                    keys = sorted(res)
                    out[f'{name}/meas_{key}_keys'] = np.array(keys)
                    out[f'{name}/meas_{key}_probs'] = np.array([float(res[k][1]) for k in keys])
        out[f'{name}/spec'] = np.array(json.dumps({'n': n, 'spec': spec, 'obs': obs}))
        print('denmat', name, n, len(spec), 'trace', float(np.trace(out[f'{name}/c128']).real),
              'purity', float(np.trace(out[f'{name}/c128'] @ out[f'{name}/c128']).real))
    # a mixed initial state given as a matrix, and a batch of them
    g = torch.Generator().manual_seed(5)
    a = torch.randn(2, 8, 8, generator=g, dtype=torch.float64) + 1j * torch.randn(2, 8, 8, generator=g,
                                                                                    dtype=torch.float64)
    rho0 = a @ a.mH
    rho0 = rho0 / rho0.diagonal(dim1=-2, dim2=-1).sum(-1).reshape(2, 1, 1)
    spec = cases['noisy3'][1]
    cir = dq.QubitCircuit(3, den_mat=True)
    wl.apply_spec(cir, spec, torch.complex128)
    cir.to(torch.double)
    with torch.no_grad():
        out['mixed3/init'] = rho0.numpy()
        out['mixed3/single'] = cir(state=rho0[0]).numpy()
        out['mixed3/batch'] = cir(state=rho0).numpy()
    np.savez_compressed(os.path.join(OUT, 'denmat.npz'), **out)
    print('denmat.npz:', len(out), 'arrays')


def dist_adjoint_build(cir):
    """The circuit of the reference's own test of the differentiable sharded expectation
    (tests/test_circuit.py:87-139): every encoded gate family, controls, Toffoli / Fredkin / swap, X / Y observables."""
    cir.rxlayer(encode=True)
    cir.rylayer(encode=True)
    cir.rzlayer(encode=True)
    cir.u3layer(encode=True)
    cir.hlayer()
    cir.cnot_ring()
    cir.toffoli(0, 1, 2)
    cir.fredkin(2, 1, 0)
    cir.swap([2, 3])
    cir.rx(0, controls=[1, 2, 3], encode=True)
    cir.ry(1, controls=[0, 2, 3], encode=True)
    cir.rz(2, controls=[0, 1, 3], encode=True)
    cir.rxx([0, 1], controls=[2, 3], encode=True)
    cir.ryy([1, 2], controls=[0, 3], encode=True)
    cir.rzz([2, 3], controls=[0, 1], encode=True)
    cir.rxy([3, 0], controls=[1, 2], encode=True)
    cir.observable(0)
    cir.observable(1, 'x')
    cir.observable([2, 3], 'xy')
    return cir


def dist_adjoint():
    """Expectation values and d(sum)/d(data) of the dense reference circuit (plain autograd, the `cir2` half of
    tests/test_circuit.py:87-139) in float64, plus a 6-qubit variant whose gates reach the rank bits of 2 and 4 ranks."""
    out = {}
    for name, n, ndata in (('ref_test_n4', 4, 10), ('ref_test_n6', 6, 10)):
        data = torch.arange(ndata, dtype=torch.float64, requires_grad=True)
        cir = dist_adjoint_build(dq.QubitCircuit(n, reupload=True))
        cir.to(torch.double)
        cir(data=data)
        exp = cir.expectation()
        exp.sum().backward()
        out[f'{name}/data'] = data.detach().numpy()
        out[f'{name}/expectation'] = exp.detach().reshape(-1).numpy()
        out[f'{name}/grad'] = data.grad.numpy()
        print('dist_adjoint', name, exp.detach().reshape(-1).tolist())
    np.savez_compressed(os.path.join(OUT, 'dist_adjoint.npz'), **out)


def unitary():
    """`QubitCircuit.get_unitary()` (reference circuit.py:467) of the all-gate-families circuit at 5 qubits and of a
    random Clifford+RX circuit at 6 qubits."""
    out = {}
    # the reference's get_unitary() of a CONTROLLED UAnyGate disagrees with its own forward pass (probe: forward =
    # oracle to 2e-16, get_unitary() off by 0.5); those entries are left out of this fixture
    all_gates = [e for e in all_gates_spec(5) if not (e['g'] in ('any', 'latent') and e.get('c'))]
    for name, n, spec in (('all_gates_n5', 5, all_gates), ('random_n6_d3', 6, wl.random_clifford_rx_spec(6, 3))):
        cir = dq.QubitCircuit(n)
        wl.apply_spec(cir, spec, torch.complex128)
        cir.to(torch.double)
        out[f'{name}/spec'] = np.array(json.dumps({'n': n, 'spec': spec}))
        out[f'{name}/unitary'] = cir.get_unitary().detach().numpy()
    np.savez_compressed(os.path.join(OUT, 'unitary.npz'), **out)


def reset_build(cir):
    """Shared with tests/test_reset.py."""
    cir.hlayer()
    cir.rxlayer(encode=True)
    cir.cnot_ring()
    cir.reset(1)                          # postselect 0
    cir.rylayer(encode=True)
    cir.cnot(0, 3)
    cir.reset([0, 3], postselect=1)
    cir.x(2)
    cir.reset(2)                          # |1> for sure: probability of the postselected outcome is exactly 0
    cir.u3layer()
    cir.cz(4, 2)
    return cir


def reset():
    """complex64 only: the reference's `.to(torch.double)` fails on a circuit that holds a Reset (a module without
    buffers, operation.py:166 -> utils.py:47, the same defect as for Barrier)."""
    out = {}
    n = 5
    cir = reset_build(dq.QubitCircuit(n))
    gp = torch.Generator().manual_seed(6)
    for op in cir.operators:
        for prm in op.parameters():
            with torch.no_grad():
                prm.copy_(torch.rand(prm.shape, generator=gp) * 6)
    data = torch.rand(2 * n, generator=torch.Generator().manual_seed(7)) * 6
    with torch.no_grad():
        out['state/c64'] = cir(data=data).reshape(-1).numpy()
        out['data'] = data.numpy()
        out['params'] = np.concatenate([prm.detach().reshape(-1).numpy() for op in cir.operators for prm in op.parameters()])
        data2 = torch.stack([data, data.flip(0), data * 0.5])
        out['data2'] = data2.numpy()
        out['state2/c64'] = cir(data=data2).reshape(3, -1).numpy()
        full = dq.QubitCircuit(3)
        full.hlayer()
        full.reset()
        full.rx(1, 0.4)
        out['full/c64'] = full().reshape(-1).numpy()
    np.savez_compressed(os.path.join(OUT, 'reset.npz'), **out)
    print('reset.npz', {k: v.shape for k, v in out.items()}, float(np.linalg.norm(out['state/c64'])))


def dense_grad_build(cir, h3):
    """Trainable dense blocks on 3 and 4 wires between ordinary layers (shared with tests/test_gpu_parity.py)."""
    cir.hlayer()
    cir.hamiltonian(h3, wires=[0, 4, 2])                              # trainable evolution time, 8 x 8 matrix form
    cir.rxlayer()
    cir.latent(wires=[1, 3, 5])                                       # trainable 8 x 8 latent matrix
    cir.cnot_ring()
    cir.hamiltonian([[0.6, 'x1z2y3z4'], [-0.3, 'z1x4']], controls=0)  # 4 wires + a control, trainable time
    cir.rylayer(encode=True)
    cir.observable(0)
    cir.observable(1, 'x')
    cir.observable([2, 5], 'zy')
    return cir


def dense_grad():
    out = {}
    n = 6
    g = torch.Generator().manual_seed(77)
    a = torch.randn(8, 8, generator=g, dtype=torch.float64) + 1j * torch.randn(8, 8, generator=g, dtype=torch.float64)
    h3 = (a + a.mH) / 2
    cir = dense_grad_build(dq.QubitCircuit(n), h3)
    cir.to(torch.double)
    for i, op in enumerate(cir.operators):
        for name, prm in op.named_parameters():
            with torch.no_grad():
                prm.copy_(torch.randn(prm.shape, generator=g, dtype=torch.float64).to(prm.dtype)
                          if not prm.is_complex() else
                          torch.randn(prm.shape, generator=g, dtype=torch.float64)
                          + 1j * torch.randn(prm.shape, generator=g, dtype=torch.float64))
            out[f'param/{i}/{name}'] = prm.detach().numpy().copy()
    data = torch.tensor([0.31, 0.2, -0.4, 0.9, 1.3, -0.7], dtype=torch.float64, requires_grad=True)
    state = cir(data=data)
    exp = cir.expectation()
    w = torch.tensor([1.0, -0.7, 0.45], dtype=torch.float64)
    loss = (w * exp.reshape(-1)).sum()
    loss.backward()
    out['h3'] = h3.numpy()
    out['data'] = data.detach().numpy()
    out['weights'] = w.numpy()
    out['state'] = state.detach().reshape(-1).numpy()
    out['expectation'] = exp.detach().reshape(-1).numpy()
    out['loss'] = loss.detach().numpy()
    out['grad/data'] = data.grad.numpy()
    for i, op in enumerate(cir.operators):
        for name, prm in op.named_parameters():
            out[f'grad/{i}/{name}'] = prm.grad.resolve_conj().numpy().copy()
    np.savez_compressed(os.path.join(OUT, 'dense_grad.npz'), **out)
    print('dense_grad', float(loss), data.grad.tolist(), [k for k in out if k.startswith('grad/')])


if __name__ == '__main__':
    which = sys.argv[1:] or ['gates', 'circuits', 'qaoa', 'fock', 'measure', 'denmat', 'hamiltonian', 'qasm3', 'fock2',
                             'dist_adjoint', 'unitary', 'dense_grad', 'reset']
    if 'dense_grad' in which:
        dense_grad()
    if 'reset' in which:
        reset()
    if 'dist_adjoint' in which:
        dist_adjoint()
    if 'unitary' in which:
        unitary()
    if 'denmat' in which:
        denmat()
    if 'hamiltonian' in which:
        hamiltonian()
    if 'qasm3' in which:
        qasm3()
    if 'fock2' in which:
        fock2()
    if 'measure' in which:
        measure()
    if 'gates' in which:
        gate_matrices()
    if 'circuits' in which:
        circuits()
    if 'qaoa' in which:
        qaoa()
    if 'fock' in which:
        fock()
