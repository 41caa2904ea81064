"""HamiltonianGate blocks `exp(-i H t)` (reference gate.py:2867-3024,
circuit.py:1450-1476): oracle and product lowering against fixtures from the unmodified reference."""
import json
import os

import numpy as np
import pytest
import torch

import deepquantum_b200 as dq
import denmat_oracle as do
import gates_np
import statevec_oracle as so
from conftest import GOLDEN
from deepquantum_b200 import workloads as wl
from helpers import emu_run_program


def _load():
    g = np.load(os.path.join(GOLDEN, 'hamiltonian.npz'))
    m = json.loads(str(g['ham6/spec']))
    return g, m['n'], m['spec']


def _build(n, spec, double, den_mat=False):
    cir = dq.QubitCircuit(n, den_mat=den_mat)
    wl.apply_spec(cir, spec, torch.complex128 if double else torch.complex64)
    if double:
        cir.to(torch.double)
    return cir


def test_oracle_matches_reference():
    g, n, spec = _load()
    out = so.run_circuit(gates_np.lower_spec(spec, n), n)
    assert np.linalg.norm(out - g['ham6/c128']) < 1e-12
    rho = do.run_spec(spec, n)
    assert np.linalg.norm(rho - g['ham6/rho_c128']) < 1e-12


def test_gate_module_matches_reference_conventions():
    gate = dq.HamiltonianGate([[0.5, 'x0y1'], [-1, 'z3y1']], t=0.3, nqubit=4)
    assert gate.wires == [0, 1, 2, 3] and gate.minmax == [0, 3] and gate.npara == 1
    assert gate.ham_tsr.dtype == torch.cfloat and gate.t.dtype == torch.float
    gate.to(torch.double)
    assert gate.ham_tsr.dtype == torch.cdouble and gate.t.dtype == torch.double
    u = gate.update_matrix()
    np.testing.assert_allclose((u @ u.mH).numpy(), np.eye(16), atol=1e-12)
    np.testing.assert_allclose((gate.inverse().update_matrix() @ u).numpy(), np.eye(16), atol=1e-12)
    ref = gates_np.lower_entry({'g': 'hamiltonian', 'ham': [[0.5, 'x0y1'], [-1, 'z3y1']], 'p': [0.3]}, 4)[0][0]
    np.testing.assert_allclose(u.numpy(), ref, atol=1e-13)
    with pytest.raises(AssertionError):
        dq.HamiltonianGate(torch.tensor([[0, 1], [0, 0]], dtype=torch.cfloat), t=0.1, nqubit=1, wires=[0])
    cir = dq.QubitCircuit(3)
    cir.hamiltonian([1.0, 'z0z2'])          # trainable evolution time
    assert cir.npara == 1 and isinstance(cir.operators[0].t, torch.nn.Parameter)


@pytest.mark.parametrize('double', [True, False])
def test_lowering_matches_reference(double):
    g, n, spec = _load()
    cdt, tol = (np.complex128, 1e-12) if double else (np.complex64, 3e-6)
    out, _ = emu_run_program(_build(n, spec, double)._get_program(), n, cdt)
    assert np.linalg.norm(out[0] - g['ham6/c128']) < tol
    rho, _ = emu_run_program(_build(n, spec, double, den_mat=True)._get_program(), 2 * n, cdt)
    ref = g['ham6/rho_c128']
    assert np.linalg.norm(rho[0].reshape(ref.shape) - ref) < tol


def test_inverse_circuit():
    g, n, spec = _load()
    cir = _build(n, spec, True)
    out, _ = emu_run_program((cir + cir.inverse())._get_program(), n, np.complex128)
    np.testing.assert_allclose(out[0], g['ham6/inv_c128'], atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('double', [True, False])
def test_gpu_matches_reference(double):
    g, n, spec = _load()
    tol = 1e-10 if double else 3e-6
    cir = _build(n, spec, double).to('cuda')
    out = cir().reshape(-1).cpu().numpy()
    assert np.linalg.norm(out - g['ham6/c128']) < tol
    rho = _build(n, spec, double, den_mat=True).to('cuda')().cpu().numpy()
    assert np.linalg.norm(rho - g['ham6/rho_c128']) < tol
    if double:
        back = cir.inverse()(state=cir()).reshape(-1).cpu().numpy()
        np.testing.assert_allclose(back, g['ham6/inv_c128'], atol=1e-10)


@pytest.mark.gpu
def test_gpu_gradient_of_evolution_time():
    """d<Z0>/dt through the adjoint sweep against torch autograd on the dense matrices (complex128)."""
    n = 4
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.hamiltonian([[0.7, 'x0z1'], [0.4, 'y1']])    # 2 wires: the reverse sweep differentiates dense blocks up to k = 2
    cir.rx(1, 0.3)
    cir.observable([0], 'z')
    cir.to(torch.double).to('cuda')
    t = cir.operators[n].t
    cir()
    val = cir.expectation().sum()
    val.backward()
    grad = t.grad.item()
    eps = 1e-6
    with torch.no_grad():
        t += eps
        cir()
        up = cir.expectation().sum().item()
        t -= 2 * eps
        cir()
        dn = cir.expectation().sum().item()
    assert abs(grad - (up - dn) / (2 * eps)) < 1e-6
