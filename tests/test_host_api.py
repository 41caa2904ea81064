"""Host side of the product package on CPU: gate matrices, lowering, matrix assembly, plan caching,
`.to()` semantics, batching -- executed through the TEST-ONLY CPU emulator of the kernel body and
compared with fixtures from the unmodified reference."""
import json
import os

import numpy as np
import pytest
import torch

import deepquantum_b200 as dq
from conftest import GOLDEN
from deepquantum_b200 import workloads as wl
from helpers import emu_run_program


def _g(name):
    return np.load(os.path.join(GOLDEN, name))


def test_gate_matrices_match_reference():
    g = _g('gate_matrices.npz')
    th = g['theta']
    consts = {'x': dq.PauliX, 'y': dq.PauliY, 'z': dq.PauliZ, 'h': dq.Hadamard, 's': dq.SGate, 'sdg': dq.SDaggerGate,
              't': dq.TGate, 'tdg': dq.TDaggerGate, 'cnot': dq.CNOT, 'swap': dq.Swap, 'iswap': dq.ImaginarySwap,
              'toffoli': dq.Toffoli, 'fredkin': dq.Fredkin}
    for name, cls in consts.items():
        m = cls().to(torch.double).matrix.numpy()
        assert np.array_equal(m, g[name]), name
    params = {'rx': dq.Rx, 'ry': dq.Ry, 'rz': dq.Rz, 'p': dq.PhaseShift, 'rxx': dq.Rxx, 'ryy': dq.Ryy,
              'rzz': dq.Rzz, 'rxy': dq.Rxy, 'rbs': dq.ReconfigurableBeamSplitter}
    for name, cls in params.items():
        gate = cls(inputs=float(th[0])).to(torch.double)
        np.testing.assert_allclose(gate.update_matrix().numpy(), g[name], atol=1e-15, err_msg=name)
        np.testing.assert_allclose(gate.inverse().update_matrix().numpy(), g[name + '_inv'], atol=1e-15)
    u3 = dq.U3Gate(inputs=[float(x) for x in th]).to(torch.double)
    np.testing.assert_allclose(u3.update_matrix().numpy(), g['u3'], atol=1e-15)
    np.testing.assert_allclose(u3.inverse().update_matrix().numpy(), g['u3_inv'], atol=1e-15)
    for plane in ('xy', 'yz', 'zx'):
        j = dq.ProjectionJ(inputs=float(th[0]), plane=plane).to(torch.double)
        np.testing.assert_allclose(j.update_matrix().numpy(), g['j_' + plane], atol=1e-15)


def _cases():
    g = _g('circuits.npz')
    return sorted({k.split('/')[0] for k in g.files if k.endswith('/spec') and not k.startswith('batched')})


@pytest.mark.parametrize('case', _cases())
def test_circuit_lowering_matches_reference(case):
    g = _g('circuits.npz')
    meta = json.loads(str(g[case + '/spec']))
    n, spec = meta['n'], meta['spec']
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec, torch.complex128)
    cir.to(torch.double)
    out, stats = emu_run_program(cir._get_program(), n, np.complex128, chunk_bits=11 if n > 12 else 0)
    ref = g[case + '/c128']
    err = np.linalg.norm(out[0] - ref) / np.linalg.norm(ref)
    assert err < 1e-12, (err, stats)
    cir32 = dq.QubitCircuit(n)
    wl.apply_spec(cir32, spec)
    out32, _ = emu_run_program(cir32._get_program(), n, np.complex64, chunk_bits=11 if n > 12 else 0)
    assert np.linalg.norm(out32[0] - ref) / np.linalg.norm(ref) < 3e-6


def test_inverse_circuit_roundtrip():
    n = 6
    spec = wl.random_clifford_rx_spec(n, 6, seed=11)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    cir.u3(2, [0.3, 0.2, 0.1], controls=[4])
    cir.to(torch.double)
    full = cir + cir.inverse()
    out, _ = emu_run_program(full._get_program(), n, np.complex128)
    expect = np.zeros(2**n)
    expect[0] = 1
    # constants are float32-rounded even in the complex128 path (like the reference), so H*H != 1 exactly
    np.testing.assert_allclose(out[0], expect, atol=5e-6)


def test_batched_data_matrices():
    """2-D data: one matrix set per sample (the reference vmaps `_forward_helper`, circuit.py:227-241)."""
    n = 4
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.rxlayer(encode=True)
    cir.cnot_ring()
    cir.rz(1, encode=True)
    cir.to(torch.double)
    data = torch.tensor([[0.1, 0.2, 0.3, 0.4, 0.5], [1.1, 1.2, 1.3, 1.4, 1.5], [2.1, 2.2, 2.3, 2.4, 2.5]],
                        dtype=torch.float64)
    cir._encode_batched(data)
    out, _ = emu_run_program(cir._get_program(), n, np.complex128, batch=3)
    for b in range(3):
        single = dq.QubitCircuit(n)
        single.hlayer()
        single.rxlayer(inputs=data[b, :4].tolist())
        single.cnot_ring()
        single.rz(1, float(data[b, 4]))
        single.to(torch.double)
        for gate, v in zip([op for op in single.operators if op.npara], data[b]):
            gate.init_para(v)   # exact float64 angles
        ref, _ = emu_run_program(single._get_program(), n, np.complex128)
        np.testing.assert_allclose(out[b], ref[0], atol=1e-14)


def test_plan_is_cached_and_invalidated():
    cir = dq.QubitCircuit(5)
    cir.hlayer()
    p1 = cir._get_program()
    assert cir._get_program() is p1
    cir.cnot(0, 1)
    assert cir._get_program() is not p1


def test_to_double_and_state_dict():
    cir = dq.QubitCircuit(3)
    cir.h(0)
    cir.rx(1)
    cir.rz(2, 0.5)
    assert cir.init_state.state.dtype == torch.complex64
    cir.to(torch.double)
    assert cir.init_state.state.dtype == torch.complex128
    assert cir.operators[0].matrix.dtype == torch.complex128
    assert cir.operators[1].theta.dtype == torch.float64
    assert isinstance(cir.operators[1].theta, torch.nn.Parameter)      # trainable (circuit.py:1047-1051)
    assert not isinstance(cir.operators[2].theta, torch.nn.Parameter)  # fixed input -> buffer
    assert cir.npara == 2 and cir.ndata == 0
    sd = cir.state_dict()
    assert 'operators.1.theta' in sd and 'init_state.state' in sd


def test_lazy_zero_state_is_not_materialised():
    st = dq.QubitState(30)
    assert 'state' not in st._buffers
    st.to(torch.double)
    assert st.dtype == torch.complex128


def test_cx_diag_cx_peephole():
    """`cnot; rz; cnot` (examples/qaoa.py:36-40) is lowered to ONE 2-target diagonal whose entries are gathered from
    the Rz matrix: same state as the un-fused lowering and as the oracle, also for the inverse circuit, with the
    gradient flowing through the gather."""
    import gates_np
    import statevec_oracle as so
    from deepquantum_b200.operation import Lowering
    n = 7
    th = [0.3, 1.1, -2.0, 0.7]

    def build():
        cir = dq.QubitCircuit(n)
        cir.hlayer()
        for i, (a, b) in enumerate([(0, 1), (2, 5), (6, 0), (3, 4)]):
            cir.cnot(a, b)
            cir.rz(b, th[i])
            cir.cnot(a, b)
            cir.rx(a, 0.4 + i)
        cir.cnot(1, 2)
        cir.s(2)                 # constant diagonal in the middle: fused too
        cir.cnot(1, 2)
        cir.cnot(1, 3)
        cir.rz(2, 0.9)           # different target: NOT the pattern
        cir.cnot(1, 3)
        return cir

    cir = build()
    cir.to(torch.double)
    prog = cir._get_program()
    assert prog.ngates == len(cir.operators) and len(prog.structs) == prog.ngates - 2 * 5
    out, _ = emu_run_program(prog, n, np.complex128)
    ops = [(op.update_matrix().detach().numpy(), op.wires, op.controls) for op in cir.operators]
    ref = so.run_circuit(ops, n)
    assert np.linalg.norm(out[0] - ref) < 1e-13
    Lowering.FUSE_CX_DIAG_CX = False
    try:
        plain = build()
        plain.to(torch.double)
        p2 = plain._get_program()
        assert len(p2.structs) == p2.ngates
        out2, _ = emu_run_program(p2, n, np.complex128)
    finally:
        Lowering.FUSE_CX_DIAG_CX = True
    assert np.linalg.norm(out[0] - out2[0]) < 1e-13
    inv = cir.inverse()
    inv.to(torch.double)
    back, _ = emu_run_program(inv._get_program(), n, np.complex128, state=out[0])
    e0 = np.zeros(2**n)
    e0[0] = 1
    assert np.linalg.norm(back[0] - e0) < 1e-6   # the float32-rounded Hadamard constant (gate.py:1069) is not unitary
    # the derived block is a differentiable gather of the Rz matrix
    rz = [op for op in cir.operators if isinstance(op, dq.Rz)][0]
    rz.theta.requires_grad_(True)
    m = prog.low.build_matrices(torch.complex128, 'cpu')
    m[prog.low.n_primary:].abs().sum().backward()
    assert rz.theta.grad is not None


def test_encoder_and_plain_gates_of_one_class_with_batched_data():
    """`rylayer(encode=True); rylayer()` with 2-D data stacks [batch] and 0-d parameters of one class."""
    n, nb = 3, 4
    cir = dq.QubitCircuit(n)
    cir.rylayer(encode=True)
    cir.rylayer()
    cir.cnot(0, 1)
    cir.to(torch.double)
    data = torch.rand(nb, n, dtype=torch.double)
    prog = cir._get_program()
    cir._encode_batched(data)
    try:
        mats = prog.low.build_matrices(torch.complex128, 'cpu')
    finally:
        for op in cir.encoders:
            for g in op.gates:
                g._batched = None
    assert mats.ndim == 2 and mats.shape[0] == nb
    for b in range(nb):
        cir.encode(data[b])
        one = prog.low.build_matrices(torch.complex128, 'cpu')
        assert torch.allclose(mats[b], one, atol=1e-14)
    out, _ = emu_run_program_batched(cir, prog, n, data)
    for b in range(nb):
        cir.encode(data[b])
        ref, _ = emu_run_program(prog, n, np.complex128)
        assert np.abs(out[b] - ref[0]).max() < 1e-12


def emu_run_program_batched(cir, prog, n, data):
    cir._encode_batched(data)
    try:
        return emu_run_program(prog, n, np.complex128, batch=data.shape[0])
    finally:
        for op in cir.encoders:
            for g in op.gates:
                g._batched = None


def test_fock_transforms_with_exact_zero_entries():
    """complex 0**0 is NaN in PyTorch: a beamsplitter at theta = 0 (identity) and a two-mode squeezer at
    r = 0 must give finite transforms and gradients."""
    from deepquantum_b200 import photonic as ph
    d = 4
    theta = torch.zeros(1, dtype=torch.double, requires_grad=True)
    phi = torch.full((1,), 0.3, dtype=torch.double)
    t = ph.bs_matrix_state(ph.BeamSplitter.mixing_matrix(theta, phi), d)
    assert torch.isfinite(t.real).all() and torch.isfinite(t.imag).all()
    eye = torch.eye(d * d, dtype=t.dtype).reshape(d, d, d, d)
    assert torch.allclose(t[0], eye, atol=1e-12)
    t.abs().sum().backward()
    assert torch.isfinite(theta.grad).all()
    r = torch.zeros(1, dtype=torch.double, requires_grad=True)
    s2 = ph.squeezing2_matrix_state(r, torch.full((1,), 0.7, dtype=torch.double), d)
    assert torch.isfinite(s2.real).all() and torch.isfinite(s2.imag).all()
    assert torch.allclose(s2[0], eye.to(s2.dtype), atol=1e-12)
    s2.abs().sum().backward()
    assert torch.isfinite(r.grad).all()


def test_inverse_keeps_layer_encoders():
    """`inverse(encode=True)` re-registers the member gates of an encoding LAYER (reference circuit.py:530-555):
    the data of the inverse circuit reaches the same number of gates, in reversed order."""
    cir = dq.QubitCircuit(3)
    cir.rxlayer(encode=True)
    cir.cnot(0, 1)
    cir.ry(2, encode=True)
    inv = cir.inverse(encode=True)
    assert inv.ndata == 4 and len(inv.encoders) == 4
    data = torch.tensor([0.1, 0.2, 0.3, 0.4])
    inv.encode(data)
    # reversed operators: ry(2), cnot, rx(2), rx(1), rx(0); encoders in that order take data[0..3]; inverse angles
    assert type(inv.operators[0]).__name__ == 'Ry' and abs(float(inv.operators[0].theta) - 0.1) < 1e-7
    thetas = [float(op.theta) for op in inv.operators[2:]]
    assert np.allclose(thetas, [0.2, 0.3, 0.4], atol=1e-7)


def _dense_grad_circuit(g):
    """The circuit of tests/golden/dense_grad.npz (oracle/make_golden.py `dense_grad_build`), parameters from the
    fixture."""
    n = 6
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.hamiltonian(torch.tensor(g['h3']), wires=[0, 4, 2])
    cir.rxlayer()
    cir.latent(wires=[1, 3, 5])
    cir.cnot_ring()
    cir.hamiltonian([[0.6, 'x1z2y3z4'], [-0.3, 'z1x4']], controls=0)
    cir.rylayer(encode=True)
    cir.observable(0)
    cir.observable(1, 'x')
    cir.observable([2, 5], 'zy')
    cir.to(torch.double)
    for key in g.files:
        if key.startswith('param/'):
            _, i, name = key.split('/')
            with torch.no_grad():
                getattr(cir.operators[int(i)], name).copy_(torch.tensor(g[key]))
    return cir


def test_trainable_dense_blocks_are_planned_for_their_cotangent():
    """Trainable dense gates on 3-4 wires carry B200Q_GATE_GRAD (a pass of their own whose reverse step accumulates
    the full cotangent); the forward result does not change (CPU-stepped kernel body against the reference state)."""
    from deepquantum_b200 import _lib as L
    g = np.load(os.path.join(GOLDEN, 'dense_grad.npz'))
    cir = _dense_grad_circuit(g)
    cir.encode(torch.tensor(g['data']))
    prog = cir._get_program()
    flagged = [(s.n_targets, bool(s.flags & L.GATE_GRAD)) for s in prog.structs if s.n_targets >= 3]
    assert flagged == [(3, True), (3, True), (4, True)]
    assert not any(s.flags & L.GATE_GRAD for s in prog.structs if s.n_targets < 3)
    out, stats = emu_run_program(prog, 6, np.complex128)
    assert np.linalg.norm(out[0] - g['state']) < 1e-10
    assert stats['direct'] >= 3
