"""Reset (reference gate.py:3027-3094, circuit.py:1603-1607): a state-dependent projection that
splits the circuit into separately fused programs.  The reference fixture is complex64 only (its `.to(double)` fails on
a circuit holding a Reset), so the numpy restatement `statevec_oracle.reset_wires` is pinned on it and then carries the
complex128 comparison."""
import os

import numpy as np
import pytest
import torch

import deepquantum_b200 as dq
import statevec_oracle as so
from conftest import GOLDEN
from helpers import emu_run_program


def _g():
    return np.load(os.path.join(GOLDEN, 'reset.npz'))


def _build(g, rdtype=torch.float32):
    """The circuit of oracle/make_golden.py `reset_build`, parameters from the fixture."""
    n = 5
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.rxlayer(encode=True)
    cir.cnot_ring()
    cir.reset(1)
    cir.rylayer(encode=True)
    cir.cnot(0, 3)
    cir.reset([0, 3], postselect=1)
    cir.x(2)
    cir.reset(2)
    cir.u3layer()
    cir.cz(4, 2)
    flat, k = torch.tensor(g['params']), 0
    for op in cir.operators:
        for prm in op.parameters():
            with torch.no_grad():
                prm.copy_(flat[k:k + prm.numel()].reshape(prm.shape))
            k += prm.numel()
    assert k == flat.numel()
    cir.to(rdtype)
    return cir


def _host_run(cir, data, cdtype):
    """Segments through the CPU-stepped kernel body, Reset through the oracle."""
    n = cir.nqubit
    data = torch.as_tensor(data)
    batch = data.shape[0] if data.ndim == 2 else 1
    if data.ndim == 2:
        cir._encode_batched(data.to(cir.init_state.state.real.dtype if hasattr(cir.init_state, 'state') else data.dtype))
    else:
        cir.encode(data)
    x = np.zeros((batch, 2**n), dtype=cdtype)
    x[:, 0] = 1
    for seg in cir._get_segments():
        if isinstance(seg, dq.Reset):
            x = so.reset_wires(x, n, seg.wires, seg.postselect).astype(cdtype)
        else:
            x, _ = emu_run_program(seg, n, cdtype, state=x, batch=batch)
    return x


def test_reset_host_logic_matches_the_reference():
    g = _g()
    cir = _build(g)
    out = _host_run(cir, torch.tensor(g['data']), np.complex64)[0]
    assert np.abs(out - g['state/c64']).max() < 3e-6
    out2 = _host_run(cir, torch.tensor(g['data2']), np.complex64)
    assert np.abs(out2 - g['state2/c64']).max() < 3e-6
    # every reset wire is |0> afterwards up to the gates that follow; the state stays normalised
    assert abs(np.linalg.norm(out) - 1) < 1e-5
    with pytest.raises(NotImplementedError):
        cir._get_program()
    with pytest.raises(NotImplementedError):
        cir.inverse()


def test_oracle_reset_branches():
    """probability-0 branch, postselect 1, all wires."""
    psi = np.zeros(8, dtype=np.complex128)
    psi[0b010] = 0.6
    psi[0b011] = 0.8j
    out = so.reset_wires(psi, 3, [1], 0)          # wire 1 is |1> for sure: the other branch, undivided
    assert np.allclose(out[[0b000, 0b001]], [0.6, 0.8j]) and np.count_nonzero(out) == 2
    out = so.reset_wires(psi, 3, [2], 1)          # p(1) = 0.64
    assert np.allclose(out[0b010], 1j) and np.count_nonzero(out) == 1
    assert np.allclose(so.reset_wires(psi, 3, [0, 1, 2]), np.eye(8)[0])


@pytest.mark.gpu
@pytest.mark.parametrize('rdtype', [torch.float32, torch.float64])
def test_gpu_reset_matches_the_reference(rdtype):
    g = _g()
    cdt = np.complex64 if rdtype == torch.float32 else np.complex128
    cir = _build(g, rdtype).to('cuda')
    ref1 = g['state/c64'] if rdtype == torch.float32 else _host_run(_build(g, rdtype), torch.tensor(g['data']).double(), cdt)[0]
    ref2 = g['state2/c64'] if rdtype == torch.float32 else _host_run(_build(g, rdtype), torch.tensor(g['data2']).double(), cdt)
    tol = 3e-6 if rdtype == torch.float32 else 1e-10
    out = cir(torch.tensor(g['data'], device='cuda', dtype=rdtype)).reshape(-1).cpu().numpy()
    assert np.abs(out - ref1).max() < tol
    out2 = cir(torch.tensor(g['data2'], device='cuda', dtype=rdtype)).reshape(3, -1).cpu().numpy()
    assert np.abs(out2 - ref2).max() < tol
    if rdtype == torch.float64:                   # and the c128 result agrees with the reference's c64 one
        assert np.abs(out - g['state/c64']).max() < 3e-6
    full = dq.QubitCircuit(3)
    full.hlayer()
    full.reset()
    full.rx(1, 0.4)
    full.to('cuda', rdtype)
    assert np.abs(full().reshape(-1).cpu().numpy() - g['full/c64']).max() < 3e-6


@pytest.mark.gpu
def test_gpu_sampled_reset_is_a_projective_measurement():
    """postselect=None (not reproducible draw by draw): the wires end in |0>, the state is normalised, and it equals
    the projection on ONE outcome of the measured wires."""
    n = 6
    torch.manual_seed(3)
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.rxlayer()
    cir.cnot_ring()
    pre = dq.QubitCircuit(n)
    pre.operators = torch.nn.Sequential(*list(cir.operators))      # the same gate objects, without the reset
    pre.npara = cir.npara
    cir.reset([1, 4], postselect=None)
    cir.to('cuda', torch.double)
    pre.to('cuda', torch.double)
    before = pre().detach().reshape([2] * n).cpu().numpy()
    after = cir().reshape([2] * n).cpu().numpy()
    assert abs(np.linalg.norm(after) - 1) < 1e-10
    assert np.abs(after[:, 1]).max() == 0 and np.abs(after[:, :, :, :, 1]).max() == 0
    cands = []
    for a in (0, 1):
        for b in (0, 1):
            proj = before[:, a, :, :, b, :]
            cands.append(np.linalg.norm(after[:, 0, :, :, 0, :] - proj / np.linalg.norm(proj)))
    assert min(cands) < 1e-10


def test_move_is_reset_then_swap():
    """`cir.move(a, b)` (reference gate.py:3141-3168: Reset on wires[1], then Swap) through the host logic against the
    oracle: the state of wire a ends up on wire b and wire a is |0>."""
    import gates_np
    n = 4
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.rx(1, 0.7)
    cir.cnot(1, 2)
    cir.move(1, 3)
    cir.ry(0, 0.3)
    cir.to(torch.double)
    assert [type(op).__name__ for op in cir.operators][-3:] == ['Reset', 'Swap', 'Ry']
    out = _host_run(cir, torch.zeros(0), np.complex128)[0]
    h = np.array([[1, 1], [1, -1]]) * np.float32(2 ** -0.5).astype(np.float64)
    ops = [(h, [w], []) for w in range(n)] + [(gates_np.rx(np.float32(0.7)), [1], []), (gates_np.X, [2], [1])]
    psi = so.run_circuit(ops, n)
    psi = so.reset_wires(psi, n, [3], 0)
    swap = np.eye(4)[[0, 2, 1, 3]]
    psi = so.run_circuit([(swap, [1, 3], []), (gates_np.ry(np.float32(0.3)), [0], [])], n, psi)
    assert np.abs(out - psi).max() < 1e-7
    t = out.reshape([2] * n)
    assert np.abs(t[:, 1]).max() < 1e-12          # wire 1 is |0> after the move


@pytest.mark.gpu
def test_gpu_move_matches_the_host_path():
    n = 4

    def build():
        cir = dq.QubitCircuit(n)
        cir.hlayer()
        cir.rx(1, 0.7)
        cir.cnot(1, 2)
        cir.move(1, 3)
        cir.ry(0, 0.3)
        cir.to(torch.double)
        return cir
    ref = _host_run(build(), torch.zeros(0), np.complex128)[0]
    out = build().to('cuda')().reshape(-1).cpu().numpy()
    assert np.abs(out - ref).max() < 1e-12
