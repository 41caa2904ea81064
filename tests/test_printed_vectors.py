"""The known answers the reference PRINTS for this path: README.md:114-124 and
tutorials/basics.ipynb cells 5-18.  They pin the oracle and the product's host logic (lowering + matrices, run
through the CPU emulator of the kernel body) to values that were not produced in this container."""
import numpy as np
import torch

import deepquantum_b200 as dq
import gates_np
import statevec_oracle as so
from helpers import emu_run_program

S = 0.5 ** 0.5


def _run(cir):
    out, _ = emu_run_program(cir._get_program(), cir.nqubit, np.complex64)
    return out[0]


def test_readme_example_state_and_expectation():
    """README.md:114-124: h(0); cnot(0, 1); rx(1, 0.2) prints [0.7036, -0.0706j, -0.0706j, 0.7036] and <Z0> = 0."""
    printed = np.array([0.7036, -0.0706j, -0.0706j, 0.7036])
    ops = [(gates_np.H, [0], []), (gates_np.CNOT, [0, 1], []), (gates_np.rx(gates_np.f32(0.2)), [1], [])]
    psi = so.run_circuit(ops, 2)
    np.testing.assert_allclose(psi, printed, atol=5e-5)
    assert abs(so.expectation_pauli(psi, 2, [0], 'z')) < 1e-7
    cir = dq.QubitCircuit(2)
    cir.h(0)
    cir.cnot(0, 1)
    cir.rx(1, 0.2)
    cir.observable(0)
    np.testing.assert_allclose(_run(cir), printed, atol=5e-5)


def test_tutorial_single_gate_answers():
    """basics.ipynb cells 5-18: X|1> = |0>; Rx(pi/2)|1> = [-0.7071j, 0.7071]; Rx(pi)|1> = [-1j, -4.3711e-08]
    (the float32 parameter shows); Rx(pi/2) on wire 1 of |11> = [0, 0, -0.7071j, 0.7071]."""
    one = dq.QubitState(nqubit=1, state=[0, 1]).state
    np.testing.assert_array_equal(one.numpy(), [[0], [1]])
    np.testing.assert_array_equal(dq.PauliX().matrix.numpy(), [[0, 1], [1, 0]])
    rx = dq.Rx(torch.pi / 2)
    np.testing.assert_allclose((rx.matrix @ one).numpy().ravel(), [-S * 1j, S], atol=1e-7)
    np.testing.assert_allclose((rx.get_matrix(torch.pi) @ one).numpy().ravel(), [-1j, -4.3711e-08], atol=1e-12)
    np.testing.assert_allclose(gates_np.rx(gates_np.f32(np.pi)) @ [0, 1], [-1j, -4.3711e-08], atol=1e-12)
    for gate, init, printed in ((lambda c: c.x(0), [0, 1], [1, 0]),
                                (lambda c: c.rx(0, torch.pi / 2), [0, 1], [-S * 1j, S])):
        cir = dq.QubitCircuit(1, init_state=init)
        gate(cir)
        out, _ = emu_run_program(cir._get_program(), 1, np.complex64, state=np.array(init, dtype=np.complex64))
        np.testing.assert_allclose(out[0], printed, atol=1e-7)
    cir = dq.QubitCircuit(2, init_state=[0, 0, 0, 1])
    cir.rx(1, torch.pi / 2)
    out, _ = emu_run_program(cir._get_program(), 2, np.complex64, state=np.array([0, 0, 0, 1], dtype=np.complex64))
    np.testing.assert_allclose(out[0], [0, 0, -S * 1j, S], atol=1e-7)
    np.testing.assert_allclose(so.run_circuit([(gates_np.rx(gates_np.f32(np.pi / 2)), [1], [])], 2,
                                              state=np.array([0, 0, 0, 1], dtype=complex)), [0, 0, -S * 1j, S],
                               atol=1e-7)


def test_tutorial_permutation_gate_matrices():
    """basics.ipynb cells 20-22: CNOT / Toffoli / Fredkin print as 0/1 permutation matrices, the controlled forms keep
    the 2x2 (4x4) target matrix."""
    np.testing.assert_array_equal(dq.CNOT(wires=[0, 1]).matrix.numpy(), gates_np.CNOT)
    np.testing.assert_array_equal(dq.PauliX(nqubit=2, wires=[1], controls=[0]).matrix.numpy(), gates_np.X)
    np.testing.assert_array_equal(dq.Toffoli(wires=[0, 1, 2]).matrix.numpy(), gates_np.TOFFOLI)
    np.testing.assert_array_equal(dq.Fredkin(wires=[0, 1, 2]).matrix.numpy(), gates_np.FREDKIN)
    np.testing.assert_array_equal(dq.Swap(nqubit=3, wires=[1, 2], controls=[0]).matrix.numpy(), gates_np.SWAP)
    assert gates_np.CNOT.tolist() == [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]
