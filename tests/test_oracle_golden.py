"""Pin the CPU oracle (oracle/) against fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

import gates_np
import statevec_oracle as so

from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_gate_matrices_match_reference():
    g = _load('gate_matrices.npz')
    th = g['theta']
    f = gates_np.f32
    for name in ('x', 'y', 'z', 'h', 's', 'sdg', 't', 'tdg'):
        assert np.array_equal(gates_np.CONST_1Q[name], g[name]), name
    assert np.array_equal(gates_np.CNOT, g['cnot'])
    assert np.array_equal(gates_np.SWAP, g['swap'])
    assert np.array_equal(gates_np.ISWAP, g['iswap'])
    assert np.array_equal(gates_np.TOFFOLI, g['toffoli'])
    assert np.array_equal(gates_np.FREDKIN, g['fredkin'])
    for name in ('rx', 'ry', 'rz', 'p', 'rxx', 'ryy', 'rzz', 'rxy', 'rbs'):
        fn = getattr(gates_np, name)
        np.testing.assert_allclose(fn(f(th[0])), g[name], rtol=0, atol=1e-15, err_msg=name)
        # inverse() of a parametric gate negates theta (gate.py:395-400, 417-421)
        np.testing.assert_allclose(fn(-f(th[0])), g[name + '_inv'], rtol=0, atol=1e-15, err_msg=name)
    np.testing.assert_allclose(gates_np.u3(f(th[0]), f(th[1]), f(th[2])), g['u3'], atol=1e-15)
    np.testing.assert_allclose(gates_np.u3(-f(th[0]), -f(th[2]), -f(th[1])), g['u3_inv'], atol=1e-15)
    for plane in ('xy', 'yz', 'zx'):
        np.testing.assert_allclose(gates_np.j(f(th[0]), plane), g['j_' + plane], atol=1e-15)


def _cases():
    g = _load('circuits.npz')
    return sorted({k.split('/')[0] for k in g.files if k.endswith('/spec') and not k.startswith('batched')})


@pytest.mark.parametrize('case', _cases())
def test_circuit_states_match_reference(case):
    g = _load('circuits.npz')
    meta = json.loads(str(g[case + '/spec']))
    n, spec = meta['n'], meta['spec']
    ops = gates_np.lower_spec(spec, n)
    out = so.run_circuit(ops, n)
    ref = g[case + '/c128']
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < 1e-12, err
    # the reference's own complex64 path sits at its fp32 rounding floor
    ref32 = g[case + '/c64']
    err32 = np.linalg.norm(out - ref32) / np.linalg.norm(ref)
    assert err32 < 3e-6, err32


def test_batched_initial_state():
    g = _load('circuits.npz')
    meta = json.loads(str(g['batched_n6/spec']))
    ops = gates_np.lower_spec(meta['spec'], meta['n'])
    out = so.run_circuit(ops, meta['n'], state=g['batched_n6/init'])
    np.testing.assert_allclose(out, g['batched_n6/c128'], atol=1e-13)


def test_controlled_equals_block_diag():
    """operation.py:265-272: a controlled gate is block_diag(I, U) on controls+wires."""
    rng = np.random.default_rng(0)
    n = 5
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    u = gates_np.u3(0.3, 1.1, -0.4)
    a = so.evolve_state_controlled(psi, u, n, [1], [3, 0])
    big = np.eye(8, dtype=complex)
    big[6:, 6:] = u
    b = so.evolve_state(psi, big, n, [3, 0, 1])
    np.testing.assert_allclose(a, b, atol=1e-14)


def test_qaoa_expectation_and_adjoint_gradient():
    import torch

    from deepquantum_b200 import workloads as wl

    g = _load('qaoa.npz')
    for key in sorted({k.split('/')[0] for k in g.files}):
        meta = json.loads(str(g[key + '/meta']))
        n, p, edges, weights = meta['n'], meta['p'], meta['edges'], meta['weights']
        _, _, layout = wl.qaoa_maxcut_structure(n, p, seed=meta['seed'])
        params = g[key + '/params']
        data = wl.qaoa_data(torch.tensor(params), weights, layout).numpy()
        ops, dops, which = [], [], []
        it = iter(range(len(data)))
        h = gates_np.H
        for q in range(n):
            ops.append((h, [q], [])); dops.append(None); which.append(None)
        for k in range(p):
            for ei, (a, b) in enumerate(edges):
                ops.append((gates_np.CNOT, [a, b], [])); dops.append(None); which.append(None)
                i = next(it)
                th = data[i]
                ops.append((gates_np.rz(th), [b], []))
                dops.append(np.diag([-0.5j * np.exp(-0.5j * th), 0.5j * np.exp(0.5j * th)])); which.append(i)
                ops.append((gates_np.CNOT, [a, b], [])); dops.append(None); which.append(None)
            for q in range(n):
                i = next(it)
                th = data[i]
                ops.append((gates_np.rx(th), [q], []))
                c, s = np.cos(th / 2), np.sin(th / 2)
                dops.append(0.5 * np.array([[-s, -1j * c], [-1j * c, -s]])); which.append(i)
        psi = so.run_circuit(ops, n)
        np.testing.assert_allclose(psi, g[key + '/state'], atol=1e-12)
        exps = np.array([so.expectation_pauli(psi, n, [a, b], 'z') for a, b in edges])
        np.testing.assert_allclose(exps, g[key + '/expectation'], atol=1e-12)
        w = np.array(weights)
        loss = 0.5 * np.sum(w * (exps - 1))
        np.testing.assert_allclose(loss, g[key + '/loss'], atol=1e-12)
        # adjoint gradient wrt every encoded angle, chained to the 2p parameters (qaoa_data)
        gang = so.adjoint_gradient(ops, dops, n, [([a, b], 'z') for a, b in edges], weights=list(0.5 * w))
        grad = np.zeros(2 * p)
        for gi, i in zip(gang, which):
            if i is None:
                continue
            kind, k, idx = layout[i]
            if kind == 'gamma':
                grad[k] += gi * 2 * weights[idx]
            else:
                grad[p + k] += gi * 2
        np.testing.assert_allclose(grad, g[key + '/grad'], atol=1e-10)


@pytest.mark.parametrize('case', _cases())
def test_torch_port_is_pinned_to_the_reference(case):
    """`oracle/torch_port.run_ops` -- the timed CPU baseline of bench.py and the checker of the large GPU parity
    tests -- against the reference's own states: bit-equal in complex128, at the float32 rounding floor in
    complex64 (ATen blocks the complex64 matmul of the permuted view differently from the reference's call)."""
    import torch
    import torch_port
    g = _load('circuits.npz')
    meta = json.loads(str(g[case + '/spec']))
    n, spec = meta['n'], meta['spec']
    ops = gates_np.lower_spec(spec, n)
    out128, done, _ = torch_port.run_ops(ops, n, dtype=torch.complex128)
    assert done == len(ops)
    ref = g[case + '/c128']
    assert np.linalg.norm(out128.numpy() - ref) / np.linalg.norm(ref) < 1e-13
    out64, _, _ = torch_port.run_ops(ops, n, dtype=torch.complex64)
    assert np.linalg.norm(out64.numpy() - g[case + '/c64']) / np.linalg.norm(ref) < 2e-6


def test_oracle_unitary_matches_reference_get_unitary():
    """The oracle applied to every basis state reproduces `QubitCircuit.get_unitary()` of the reference
    (circuit.py:467-477; fixture tests/golden/unitary.npz)."""
    g = _load('unitary.npz')
    for case in sorted({k.split('/')[0] for k in g.files}):
        meta = json.loads(str(g[case + '/spec']))
        n, spec = meta['n'], meta['spec']
        ops = gates_np.lower_spec(spec, n)
        u = np.stack([so.run_circuit(ops, n, state=np.eye(2**n, dtype=np.complex128)[j]) for j in range(2**n)], axis=1)
        assert np.abs(u - g[case + '/unitary']).max() < 1e-12
