"""CombinedSingleGate (reference gate.py:1790-1903) and shot-based expectation values (reference
circuit.py:400-426, qmath.py:863-871)."""
import numpy as np
import pytest
import torch

import deepquantum_b200 as dq
import gates_np
import statevec_oracle as so
from helpers import emu_run_program


def _combined(n, wire, controls=None):
    gates = [dq.Rx(0.3), dq.Hadamard(), dq.Rz(1.1), dq.U3Gate([0.2, 0.5, 0.9]), dq.TGate()]
    return dq.CombinedSingleGate(gates, nqubit=n, wires=[wire], controls=controls)


def _oracle_ops(wire, controls):
    f = gates_np.f32
    mats = [gates_np.rx(f(0.3)), gates_np.H, gates_np.rz(f(1.1)), gates_np.u3(f(0.2), f(0.5), f(0.9)), gates_np.T]
    return [(m, [wire], list(controls or [])) for m in mats]


def test_combined_single_gate_is_the_product_of_its_members():
    comb = _combined(1, 0).to(torch.double)
    ref = np.eye(2)
    for m, _, _ in _oracle_ops(0, None):
        ref = m @ ref
    np.testing.assert_allclose(comb.update_matrix().numpy(), ref, atol=1e-14)
    assert comb.npara == 1 + 1 + 3
    inv = comb.inverse()
    np.testing.assert_allclose((inv.update_matrix() @ comb.update_matrix()).numpy(), np.eye(2), atol=2e-7)
    comb.add(dq.PauliY().to(torch.double))
    np.testing.assert_allclose(comb.update_matrix().numpy(), gates_np.Y @ ref, atol=1e-14)


@pytest.mark.parametrize('controls', [None, [0, 3]])
def test_combined_single_gate_in_a_circuit(controls):
    n = 5
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.add(_combined(n, 2, controls))
    cir.cnot(2, 4)
    cir.to(torch.double)
    out, stats = emu_run_program(cir._get_program(), n, np.complex128)
    ops = [(gates_np.H, [w], []) for w in range(n)] + _oracle_ops(2, controls) + [(gates_np.CNOT, [2, 4], [])]
    np.testing.assert_allclose(out[0], so.run_circuit(ops, n), atol=1e-13)
    assert len(cir._get_program().structs) == n + 2     # ONE record for the five member gates


@pytest.mark.gpu
def test_gpu_combined_gate_gradient_reaches_member_parameters():
    n = 4
    rx = dq.Rx(0.3, requires_grad=True)
    comb = dq.CombinedSingleGate([dq.Hadamard(), rx, dq.SGate()], nqubit=n, wires=[1])
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.add(comb)
    cir.cnot(1, 2)
    cir.ry(2, 0.7)
    cir.observable([2], 'z')
    cir.to(torch.double).to('cuda')
    cir()
    cir.expectation().sum().backward()
    grad = rx.theta.grad.item()
    eps = 1e-6
    vals = []
    with torch.no_grad():
        for d in (eps, -2 * eps):
            rx.theta += d
            cir()
            vals.append(cir.expectation().sum().item())
    assert abs(grad - (vals[0] - vals[1]) / (2 * eps)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('den_mat', [False, True])
def test_gpu_sampled_expectation_converges_to_the_exact_value(den_mat):
    n = 5
    cir = dq.QubitCircuit(n, den_mat=den_mat)
    cir.hlayer()
    cir.rx(0, 0.4)
    cir.cnot(0, 1)
    cir.ry(2, 1.2)
    cir.u3(3, [0.3, 0.8, 1.4])
    if den_mat:
        cir.depolarizing(1, 0.3)
    cir.observable([0, 1], 'zz')
    cir.observable([2], 'x')
    cir.observable([3, 2], 'yz')
    cir.to('cuda')
    cir()
    exact = cir.expectation().cpu().numpy()
    torch.manual_seed(3)
    shots = 40000
    est = cir.expectation(shots=shots).cpu().numpy()
    assert est.shape == exact.shape
    assert np.abs(est - exact).max() < 5 / np.sqrt(shots)     # 5 sigma of a +-1 variable
