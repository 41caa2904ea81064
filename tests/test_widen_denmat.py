"""Density-matrix path (reference qmath.py:509-540, operation.py:221-262, 594-600,
channel.py): oracle pinned by fixtures from the unmodified reference; product lowering (row gate + conjugated
column gate, channels as superoperator gates on a 2n-qubit amplitude vector) checked on the CPU through the
emulator of the kernel body and on the GPU through the C ABI."""
import json
import os

import numpy as np
import pytest
import torch

import deepquantum_b200 as dq
import denmat_oracle as do
from conftest import GOLDEN
from deepquantum_b200 import workloads as wl
from helpers import emu_run_program

CASES = ['noisy3', 'noisy5', 'allgates5', 'noisy7']


def _g():
    return np.load(os.path.join(GOLDEN, 'denmat.npz'))


def _meta(g, case):
    m = json.loads(str(g[case + '/spec']))
    return m['n'], m['spec'], m['obs']


# ---- oracle vs reference fixtures ------------------------------------------------------------------------------
@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference(case):
    g = _g()
    n, spec, obs = _meta(g, case)
    rho = do.run_spec(spec, n)
    ref = g[case + '/c128']
    assert np.linalg.norm(rho - ref) / np.linalg.norm(ref) < 1e-12
    exp = [do.expectation_pauli(rho, n, w, b) for w, b in obs]
    np.testing.assert_allclose(exp, g[case + '/exp_c128'], atol=1e-12)
    for key, wires in (('all', None), ('sub', [2, 0])):
        p = do.measure_probs(rho, n, wires)
        idx = [int(k, 2) for k in g[f'{case}/meas_{key}_keys']]
        np.testing.assert_allclose(p[idx], g[f'{case}/meas_{key}_probs'], atol=1e-12)


def test_oracle_mixed_initial_state():
    g = _g()
    n, spec, _ = _meta(g, 'noisy3')
    np.testing.assert_allclose(do.run_spec(spec, n, g['mixed3/init'][0]), g['mixed3/single'], atol=1e-13)


# ---- product host logic on the CPU (emulated kernel body) -------------------------------------------------------
def _build(n, spec, double, obs=()):
    cir = dq.QubitCircuit(n, den_mat=True)
    wl.apply_spec(cir, spec, torch.complex128 if double else torch.complex64)
    for w, b in obs:
        cir.observable(w, b)
    if double:
        cir.to(torch.double)
    return cir


@pytest.mark.parametrize('svd', [False, True])
def test_kraus_operators_match_oracle(svd, monkeypatch):
    from deepquantum_b200.operation import DenMatLowering
    monkeypatch.setattr(DenMatLowering, 'DAMPING_SVD', svd)
    th = [0.3, 0.9, 0.5, 1.2]
    pairs = [(dq.channel.BitFlip, 'bit_flip', 1), (dq.channel.PhaseFlip, 'phase_flip', 1),
             (dq.channel.Depolarizing, 'depolarizing', 1), (dq.channel.Pauli, 'pauli', 4),
             (dq.channel.AmplitudeDamping, 'amp_damp', 1), (dq.channel.PhaseDamping, 'phase_damp', 1),
             (dq.channel.GeneralizedAmplitudeDamping, 'gen_amp_damp', 2)]
    for cls, name, k in pairs:
        ch = cls(inputs=th[:k] if k > 1 else th[0]).to(torch.double)
        ks = ch.update_matrix().numpy()
        ref = np.stack(do.kraus(name, th[:k]))
        np.testing.assert_allclose(ks, ref, atol=1e-15, err_msg=name)
        # trace preservation and the superoperator the lowering uses
        np.testing.assert_allclose(sum(k_.conj().T @ k_ for k_ in ks), np.eye(2), atol=1e-7)
        full = sum(np.kron(k_, k_.conj()) for k_ in ref).reshape(2, 2, 2, 2)     # [row', col', row, col]
        low = ch._lowered_matrix().numpy()
        if ch._diagonal_kraus:
            low = low.reshape(4, 4)
            np.testing.assert_allclose(low, full.reshape(4, 4), atol=1e-15)
            assert np.count_nonzero(low - np.diag(np.diagonal(low))) == 0
        elif ch._pauli_kraus:   # [diagonal in the Bell basis | exact Hadamard]
            r = 0.5 ** 0.5
            bell = np.kron(np.array([[r, r], [r, -r]]), np.eye(2)) @ np.array(
                [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])       # H(row) . CX(row->col)
            dmat, had = low[:16].reshape(4, 4), low[16:]
            assert np.count_nonzero(dmat - np.diag(np.diagonal(dmat))) == 0
            np.testing.assert_allclose(bell.T @ dmat @ bell, full.reshape(4, 4), atol=1e-15, err_msg=name)
            np.testing.assert_allclose(had, [r, r, r, -r], atol=1e-16)
        elif ch._damping_kraus and svd:   # [Vh | diagonal (flipped parity frame) | U]: parity-0 block U diag Vh, parity-1 scalar
            vh, dmat, u = low[:4].reshape(2, 2), low[4:20].reshape(4, 4), low[20:].reshape(2, 2)
            for rot in (vh, u):
                assert abs(np.linalg.det(rot) - 1) < 1e-12 and abs(rot[0, 0] - rot[1, 1]) < 1e-12
            dd = np.diagonal(dmat)
            rebuilt = np.zeros((2, 2, 2, 2), dtype=complex)
            m0 = u @ np.diag([dd[1], dd[3]]) @ vh
            for a in range(2):
                for b in range(2):
                    rebuilt[a, a, b, b] = m0[a, b]
                rebuilt[a, 1 - a, a, 1 - a] = dd[0]
            assert dd[0] == dd[2]
            np.testing.assert_allclose(rebuilt, full, atol=1e-14, err_msg=name)
        else:       # parity blocks [M1 | M0]; everything outside them is zero
            m1, m0 = low[:4].reshape(2, 2), low[4:].reshape(2, 2)
            rebuilt = np.zeros((2, 2, 2, 2), dtype=complex)
            for a in range(2):
                for b in range(2):
                    rebuilt[a, a, b, b] = m0[a, b]
                    rebuilt[a, 1 - a, b, 1 - b] = m1[a, b]
            np.testing.assert_allclose(rebuilt, full, atol=1e-15, err_msg=name)


def test_custom_channel_uses_dense_superoperator():
    """A channel outside the parity-preserving family (Kraus operators H-rotated) goes through the dense 2-target op."""
    class Rotated(dq.Channel):
        def get_matrix(self, theta):
            p = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
            h = torch.tensor([[1, 1], [1, -1]], dtype=torch.cfloat) / 2**0.5
            z = torch.tensor([[1, 0], [0, -1]], dtype=torch.cfloat)
            return torch.stack([torch.sqrt(1 - p) * torch.eye(2, dtype=torch.cfloat), torch.sqrt(p) * (h @ z)])

    n = 3
    cir = dq.QubitCircuit(n, den_mat=True)
    cir.hlayer()
    cir.rx(1, 0.4)
    cir.add(Rotated(0.6, nqubit=n, wires=[1]))
    cir.cnot(1, 2)
    cir.add(Rotated(0.3, nqubit=n, wires=[2]))
    cir.to(torch.double)
    out, _ = emu_run_program(cir._get_program(), 2 * n, np.complex128)
    rho = np.zeros(4**n, dtype=complex)
    rho[0] = 1
    hm = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    import gates_np
    for w in range(n):
        rho = do.evolve_den_mat(rho, gates_np.H, n, [w])
    rho = do.evolve_den_mat(rho, gates_np.rx(gates_np.f32(0.4)), n, [1])
    for th, w in ((0.6, 1), (0.3, 2)):
        if w == 2:
            rho = do.evolve_den_mat(rho, gates_np.X, n, [2], [1])
        p = np.sin(np.float64(np.float32(th))) ** 2
        hz = (hm.astype(np.complex64) @ np.diag([1, -1]).astype(np.complex64)).astype(complex)
        rho = do.apply_channel(rho, [np.sqrt(1 - p) * np.eye(2), np.sqrt(p) * hz], n, w)
    np.testing.assert_allclose(out[0], rho, atol=1e-7)


@pytest.mark.parametrize('case', CASES)
def test_lowering_matches_reference(case):
    g = _g()
    n, spec, _ = _meta(g, case)
    ref = g[case + '/c128']
    cir = _build(n, spec, True)
    prog = cir._get_program()
    assert prog.low.state_qubits == 2 * n
    out, stats = emu_run_program(prog, 2 * n, np.complex128, chunk_bits=11 if 2 * n > 12 else 0)
    assert np.linalg.norm(out[0].reshape(ref.shape) - ref) / np.linalg.norm(ref) < 1e-12, stats
    cir32 = _build(n, spec, False)
    out32, _ = emu_run_program(cir32._get_program(), 2 * n, np.complex64, chunk_bits=11 if 2 * n > 12 else 0)
    assert np.linalg.norm(out32[0].reshape(ref.shape) - ref) / np.linalg.norm(ref) < 3e-6
    # the c64 reference itself is this far from its c128 run
    ref32 = g[case + '/c64']
    assert np.linalg.norm(out32[0].reshape(ref.shape) - ref32) / np.linalg.norm(ref) < 3e-6


@pytest.mark.parametrize('bell', [True, False])
def test_pauli_channels_bell_basis_and_parity_block_lowerings_agree(bell, monkeypatch):
    from deepquantum_b200.operation import DenMatLowering
    monkeypatch.setattr(DenMatLowering, 'PAULI_BELL', bell)
    monkeypatch.setattr(DenMatLowering, 'DAMPING_SVD', bell)
    g = _g()
    n, spec, _ = _meta(g, 'noisy5')
    ref = g['noisy5/c128']
    prog = _build(n, spec, True)._get_program()
    kinds = {r[0] for r in prog.low.records if isinstance(r[0], str)}
    assert ('super_pauli' in kinds) == bell and ('super_damping' in kinds) == bell
    out, _ = emu_run_program(prog, 2 * n, np.complex128)
    assert np.linalg.norm(out[0].reshape(ref.shape) - ref) / np.linalg.norm(ref) < 1e-12


def test_lowering_fuses_row_and_column_gates():
    """Row and column records act on disjoint bits: a whole noisy layer fits one pass."""
    n = 5
    cir = _build(n, wl.noisy_circuit_spec(n, 1, seed=3), False)
    prog = cir._get_program()
    out, stats = emu_run_program(prog, 2 * n, np.complex64)
    assert stats['passes'] <= 3 and len(prog.structs) > 2 * n


def test_lowering_mixed_and_batched_initial_state():
    g = _g()
    n, spec, _ = _meta(g, 'noisy3')
    cir = _build(n, spec, True)
    init = g['mixed3/init']
    out, _ = emu_run_program(cir._get_program(), 2 * n, np.complex128, state=init.reshape(2, -1), batch=2)
    np.testing.assert_allclose(out.reshape(2, 8, 8), g['mixed3/batch'], atol=1e-13)
    np.testing.assert_allclose(out[0].reshape(8, 8), g['mixed3/single'], atol=1e-13)


def test_host_api_of_density_matrices():
    st = dq.QubitState(2, 'entangle', den_mat=True)
    assert st.state.shape == (4, 4) and abs(st.state[0, 3] - 0.5) < 1e-7
    assert dq.state.is_density_matrix(st.state)
    assert not dq.state.is_density_matrix(torch.eye(4, dtype=torch.cfloat))
    cir = dq.QubitCircuit(2, den_mat=True)
    cir.h(0)
    cir.bit_flip(0, 0.2)
    assert isinstance(cir.operators[1], dq.channel.BitFlip) and cir.operators[1].den_mat
    with pytest.raises(AssertionError):
        dq.QubitCircuit(2).bit_flip(0, 0.2)          # reference circuit.py:1542
    with pytest.raises(AssertionError):
        cir.get_unitary()                            # reference circuit.py:463-465
    with pytest.raises(dq.B200QError):
        cir()                                        # CPU tensors: no fallback
    assert abs(float(cir.operators[1].prob) - np.sin(np.float32(0.2))**2) < 1e-7


# ---- GPU parity ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('double', [True, False])
def test_gpu_density_matrix_matches_reference(case, double):
    g = _g()
    n, spec, obs = _meta(g, case)
    cir = _build(n, spec, double, [tuple(o) for o in obs])
    cir.to('cuda')
    rho = cir()
    assert rho.shape == (2**n, 2**n)
    ref = g[case + '/c128']
    err = np.linalg.norm(rho.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < (1e-10 if double else 3e-6), err
    exp = cir.expectation().cpu().numpy()
    np.testing.assert_allclose(exp, g[case + '/exp_c128'], atol=1e-10 if double else 5e-6)
    if double:
        for key, wires in (('all', None), ('sub', [2, 0])):
            res = cir.measure(shots=4096, with_prob=True, wires=wires)
            pref = dict(zip([str(k) for k in g[f'{case}/meas_{key}_keys']], g[f'{case}/meas_{key}_probs']))
            full = do.measure_probs(ref, n, wires)
            assert sum(v[0] for v in res.values()) == 4096
            for k, (cnt, p) in res.items():
                assert abs(float(p) - full[int(k, 2)]) < 1e-9
                if k in pref:
                    assert abs(float(p) - pref[k]) < 1e-9


@pytest.mark.gpu
def test_gpu_mixed_batched_state_and_standalone_modules():
    g = _g()
    n, spec, _ = _meta(g, 'noisy3')
    cir = _build(n, spec, True).to('cuda')
    init = torch.from_numpy(g['mixed3/init']).cuda()
    np.testing.assert_allclose(cir(state=init).cpu().numpy(), g['mixed3/batch'], atol=1e-12)
    np.testing.assert_allclose(cir(state=init[0]).cpu().numpy(), g['mixed3/single'], atol=1e-12)
    # gate / channel modules called on a density matrix (reference operation.py:221-229, 594-608)
    rho = init[0]
    gate = dq.Rx(0.4, nqubit=3, wires=[1], controls=[2], den_mat=True).to(torch.double).to('cuda')
    chan = dq.channel.AmplitudeDamping(0.7, nqubit=3, wires=[0]).to(torch.double).to('cuda')
    out = chan(gate(rho)).cpu().numpy()
    ref = do.evolve_den_mat(g['mixed3/init'][0].reshape(-1), gate.update_matrix().detach().cpu().numpy(), 3, [1], [2])
    ref = do.apply_channel(ref, do.kraus('amp_damp', [0.7]), 3, 0).reshape(8, 8)
    np.testing.assert_allclose(out, ref, atol=1e-12)


@pytest.mark.gpu
def test_gpu_density_matrix_large_properties():
    """12 qubits (a 24-qubit amplitude vector, 128 MiB): trace 1, Hermitian, purity < 1, and equal to |psi><psi|
    when there is no channel."""
    n = 12
    spec = wl.noisy_circuit_spec(n, 3, seed=9)
    cir = _build(n, spec, False).to('cuda')
    rho = cir()
    tr = rho.diagonal().sum()
    assert abs(tr.real.item() - 1) < 1e-4 and abs(tr.imag.item()) < 1e-5
    assert (rho - rho.mH).abs().max().item() < 1e-6
    purity = (rho.abs() ** 2).sum().item()
    assert 0 < purity < 0.9
    pure = wl.random_clifford_rx_spec(n, 3, seed=9)
    c1 = dq.QubitCircuit(n, den_mat=True)
    wl.apply_spec(c1, pure)
    c2 = dq.QubitCircuit(n)
    wl.apply_spec(c2, pure)
    rho = c1.to('cuda')()
    psi = c2.to('cuda')()
    assert (rho - psi @ psi.mH).abs().max().item() < 2e-6


def test_lowering_batched_data_with_encoded_channels():
    """2-D data: encoder gates AND encoder channels get one column block per sample (explicit batch)."""
    n = 3
    cir = dq.QubitCircuit(n, den_mat=True)
    cir.hlayer()
    cir.rx(0, encode=True)
    cir.bit_flip(0, encode=True)
    cir.cnot(0, 2)
    cir.gen_amp_damp(2, encode=True)
    cir.phase_damp(1, encode=True)
    cir.to(torch.double)
    data = torch.tensor([[0.3, 0.5, 0.2, 0.9, 0.4], [1.1, 0.25, 0.7, 0.35, 0.8]], dtype=torch.float64)
    cir._encode_batched(data)
    out, _ = emu_run_program(cir._get_program(), 2 * n, np.complex128, batch=2)
    for b in range(2):
        d = data[b].tolist()
        spec = [{'g': 'hlayer'}, {'g': 'rx', 'w': [0], 'p': d[0:1], 'exact': True},
                {'g': 'bit_flip', 'w': [0], 'p': d[1:2]}, {'g': 'cnot', 'w': [0, 2]},
                {'g': 'gen_amp_damp', 'w': [2], 'p': d[2:4]}, {'g': 'phase_damp', 'w': [1], 'p': d[4:5]}]
        rho = np.zeros(4**n, dtype=complex)
        rho[0] = 1
        for e in spec:
            if e['g'] in do.CHANNELS:
                rho = do.apply_channel(rho, do.kraus(e['g'], e['p'], exact=True), n, e['w'][0])
            else:
                import gates_np
                for m, w, c in gates_np.lower_entry(e, n):
                    rho = do.evolve_den_mat(rho, m, n, w, c)
        np.testing.assert_allclose(out[b], rho, atol=1e-13)


@pytest.mark.parametrize('init', ['equal', 'entangle'])
def test_named_initial_density_matrices(init):
    cir = dq.QubitCircuit(3, init_state=init, den_mat=True)
    cir.rx(0, 0.3)
    cir.cnot(0, 2)
    cir.amp_damp(2, 0.4)
    cir.to(torch.double)
    st0 = cir.init_state.state.numpy()
    assert st0.shape == (8, 8) and abs(np.trace(st0) - 1) < 1e-7
    out, _ = emu_run_program(cir._get_program(), 6, np.complex128, state=st0.reshape(1, -1))
    spec = [{'g': 'rx', 'w': [0], 'p': [0.3]}, {'g': 'cnot', 'w': [0, 2]}, {'g': 'amp_damp', 'w': [2], 'p': [0.4]}]
    np.testing.assert_allclose(out[0].reshape(8, 8), do.run_spec(spec, 3, st0), atol=1e-14)


def test_composition_and_encoding_of_density_matrix_circuits():
    a = dq.QubitCircuit(2, den_mat=True)
    a.h(0)
    a.bit_flip(0, 0.2)
    b = dq.QubitCircuit(2, den_mat=True)
    b.cnot(0, 1)
    b.phase_damp(1, 0.5)
    c = a + b
    assert c.den_mat and len(c.operators) == 4 and c._get_program().low.state_qubits == 4
    inv = b.inverse()
    assert inv.den_mat and [type(o).__name__ for o in inv.operators] == ['PhaseDamping', 'CNOT']
    a.add(b)
    assert len(a.operators) == 4
    st = dq.QubitState(2, torch.tensor([1, 2, 3, 4], dtype=torch.cfloat), den_mat=True)
    assert st.state.shape == (4, 4) and abs(st.state.diagonal().sum() - 1) < 1e-6
    assert dq.QubitState(2, torch.randn(3, 4, dtype=torch.cfloat), den_mat=True).state.shape == (3, 4, 4)
    e = dq.QubitCircuit(2, den_mat=True)
    e.rx(0, encode=True)
    e.depolarizing(1, encode=True)
    e.pauli(0, encode=True)
    assert e.ndata == 6
    e.encode(torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.6]))
    np.testing.assert_allclose(e.operators[1].theta.numpy(), [0.2], atol=1e-7)
    np.testing.assert_allclose(e.operators[2].theta.numpy(), [0.3, 0.4, 0.5, 0.6], atol=1e-7)
    out, _ = emu_run_program(e._get_program(), 4, np.complex64)
    spec = [{'g': 'rx', 'w': [0], 'p': [0.1]}, {'g': 'depolarizing', 'w': [1], 'p': [0.2]},
            {'g': 'pauli', 'w': [0], 'p': [0.3, 0.4, 0.5, 0.6]}]
    np.testing.assert_allclose(out[0].reshape(4, 4), do.run_spec(spec, 2), atol=2e-7)


def test_two_wire_custom_channel():
    """A correlated two-qubit channel (user subclass of Channel): one dense 4-target superoperator record."""
    class CorrelatedFlip(dq.Channel):
        def get_matrix(self, theta):
            p = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
            x = torch.tensor([[0, 1], [1, 0]], dtype=torch.cfloat)
            y = torch.tensor([[0, -1j], [1j, 0]], dtype=torch.cfloat)
            return torch.stack([torch.sqrt(1 - p) * torch.eye(4, dtype=torch.cfloat), torch.sqrt(p) * torch.kron(x, y)])

    n = 3
    cir = dq.QubitCircuit(n, den_mat=True)
    cir.hlayer()
    cir.rx(0, 0.4)
    cir.cnot(0, 2)
    cir.add(CorrelatedFlip(0.6, nqubit=n, wires=[2, 0]))
    cir.ry(1, 0.9)
    cir.add(CorrelatedFlip(0.3, nqubit=n, wires=[0, 1]))
    cir.to(torch.double)
    prog = cir._get_program()
    assert [r[0] for r in prog.low.records if isinstance(r[0], str)] == ['super', 'super']
    out, _ = emu_run_program(prog, 2 * n, np.complex128)
    import gates_np
    rho = np.zeros(4**n, dtype=complex)
    rho[0] = 1
    for w in range(n):
        rho = do.evolve_den_mat(rho, gates_np.H, n, [w])
    rho = do.evolve_den_mat(rho, gates_np.rx(gates_np.f32(0.4)), n, [0])
    rho = do.evolve_den_mat(rho, gates_np.X, n, [2], [0])
    xy = np.kron(gates_np.X, gates_np.Y)

    def flip(r, th, wires):
        p = np.sin(np.float64(np.float32(th))) ** 2
        return sum(do.evolve_den_mat(r, k, n, wires) for k in (np.sqrt(1 - p) * np.eye(4), np.sqrt(p) * xy))

    rho = flip(rho, 0.6, [2, 0])
    rho = do.evolve_den_mat(rho, gates_np.ry(gates_np.f32(0.9)), n, [1])
    rho = flip(rho, 0.3, [0, 1])
    np.testing.assert_allclose(out[0], rho, atol=1e-7)
    assert abs(out[0].reshape(8, 8).trace() - 1) < 1e-6      # float32-rounded Hadamard constants, like the reference
