"""Fock tensor path, beamsplitter family / rotations / Kerr / displacement gates (reference photonic/gate.py:414-877, 1336-1489, 2291-2483,
photonic/circuit.py:2026-2245, 2471-2520, 2628-2680): local Fock matrices and final states against fixtures from
the unmodified reference -- on the CPU through the emulator of the qudit kernel's geometry, on the GPU through the
kernel."""
import json
import os

import numpy as np
import pytest
import torch

import deepquantum_b200 as dq
from conftest import GOLDEN
from test_fock import _emu_qudit

KEYS = ['m3_c4', 'm3_c6']


def _g():
    return np.load(os.path.join(GOLDEN, 'fock2.npz'))


def _apply(cir, spec):
    for e in spec:
        g, w, prm = e['g'], e['w'], e.get('p', [])
        if g in ('s', 'd'):
            getattr(cir, g)(w[0], prm[0], prm[1])
        elif g == 's2':
            cir.s2(w, prm[0], prm[1])
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
        else:
            getattr(cir, g)(w[0], prm[0])


def _circuit(key, double):
    g = _g()
    n, d = int(key[1]), int(key.split('_c')[1])
    cir = dq.QumodeCircuit(n, 'vac', cutoff=d, backend='fock', basis=False)
    _apply(cir, json.loads(str(g['spec'])))
    if double:
        cir.to(torch.double)
    return g, n, d, cir


@pytest.mark.parametrize('key', KEYS)
def test_fock_matrices_match_reference(key):
    g, n, d, cir = _circuit(key, True)
    mats = cir.build_matrices(torch.complex128, 'cpu')
    for i, (op, m) in enumerate(zip(cir.operators, mats)):
        ref = g[f'{key}/mat{i}'].reshape(m.shape)
        np.testing.assert_allclose(m.numpy(), ref, atol=1e-13, err_msg=f'{i} {type(op).__name__}')
        np.testing.assert_allclose(op.update_matrix_state().reshape(m.shape).numpy(), ref, atol=1e-13)


@pytest.mark.parametrize('key', KEYS)
def test_fock_states_match_reference(key):
    g, n, d, cir = _circuit(key, True)
    st = np.zeros(d**n, dtype=np.complex128)
    st[0] = 1
    for op, m in zip(cir.operators, cir.build_matrices(torch.complex128, 'cpu')):
        st = _emu_qudit(st, n, d, m.numpy(), op.wires)
    ref = g[key + '/c128']
    assert np.linalg.norm(st - ref) / np.linalg.norm(ref) < 1e-12


def test_builder_conventions():
    cir = dq.QumodeCircuit(2, 'vac', cutoff=3)
    cir.bs_theta([0, 1])
    cir.mzi([0, 1], encode=True)
    cir.dc([0, 1])
    assert isinstance(cir.operators[0].theta, torch.nn.Parameter) and cir.operators[0].npara == 1
    assert not isinstance(cir.operators[1].theta, torch.nn.Parameter)
    assert abs(float(cir.operators[0].phi) - np.float32(np.pi / 2)) == 0          # float32 constant, like the reference
    assert cir.operators[2].convention == 'rx' and abs(float(cir.operators[2].theta) - np.float32(np.pi / 2)) == 0
    with pytest.raises(AssertionError):
        cir.bs_rx([0, 1], 0.3, mu=0.0, sigma=0.1)


def test_encoders_route_data_like_the_reference():
    """`forward(data, state)` of the reference (photonic/circuit.py:405-431): 1-D data feeds the encoder gates in
    the order they were added; trainable gates keep their parameters."""
    cir = dq.QumodeCircuit(3, 'vac', cutoff=3)
    cir.ps(0, encode=True)
    cir.bs([0, 1])
    cir.s(2, encode=True)
    cir.bs_theta([1, 2], encode=True)
    assert cir.ndata == 4 and cir.npara == 2 and len(cir.encoders) == 3
    free = [float(cir.operators[1].theta), float(cir.operators[1].phi)]
    cir.encode(torch.tensor([0.1, 0.2, 0.3, 0.4]))
    assert abs(float(cir.operators[0].theta) - 0.1) < 1e-7
    assert abs(float(cir.operators[2].r) - 0.2) < 1e-7 and abs(float(cir.operators[2].theta) - 0.3) < 1e-7
    assert abs(float(cir.operators[3].theta) - 0.4) < 1e-7
    assert abs(float(cir.operators[3].phi) - np.float32(np.pi / 2)) == 0
    assert [float(cir.operators[1].theta), float(cir.operators[1].phi)] == free
    mats = cir.build_matrices(torch.complex128, 'cpu')
    ref = dq.photonic.PhaseShift(0.1, 3, 0, 3).update_matrix_state()
    assert torch.allclose(mats[0], ref.to(torch.complex128), atol=1e-7)
    with pytest.raises(AssertionError):
        cir.encode(torch.tensor([0.1, 0.2]))


@pytest.mark.gpu
@pytest.mark.parametrize('key', KEYS)
@pytest.mark.parametrize('double', [True, False])
def test_gpu_fock_states_match_reference(key, double):
    g, n, d, cir = _circuit(key, double)
    cir.to('cuda')
    out = cir().reshape(-1).cpu().numpy()
    ref = g[key + '/c128']
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < (1e-10 if double else 2e-6), err


def test_fock_state_constructors():
    """`'vac'`, a basis state, a superposition list (reference photonic/state.py:62-96, not normalised)."""
    st = dq.photonic.FockState([(0.6, [1, 0, 0]), (0.8j, [0, 1, 1])], nmode=3, cutoff=4).state
    assert st.shape == (1, 4, 4, 4) and abs(st[0, 1, 0, 0] - 0.6) < 1e-7 and abs(st[0, 0, 1, 1] - 0.8j) < 1e-7
    assert abs(float(st.abs().sum()) - 1.4) < 1e-6
    assert dq.photonic.FockState([(0.6, [1, 0, 0]), (0.8, [0, 1, 1])]).state.shape == (1, 3, 3, 3)   # cutoff = 2 + 1
    assert dq.photonic.FockState([1, 0, 2]).state[0, 1, 0, 2] == 1
    assert dq.photonic.FockState('vac', nmode=2, cutoff=3).state[0, 0, 0] == 1


def _c5_circuit(nmode, cutoff, rdtype):
    import deepquantum_b200 as dq
    from deepquantum_b200 import workloads as wl
    spec = wl.fock_interferometer_spec(nmode)
    cir = dq.QumodeCircuit(nmode, 'vac', cutoff=cutoff, backend='fock', basis=False)
    for e in spec:
        if e['g'] == 's':
            cir.s(e['w'][0], e['p'][0], e['p'][1])
        else:
            cir.bs(e['w'], e['p'])
    cir.to('cuda', rdtype)
    return cir


@pytest.mark.gpu
@pytest.mark.parametrize('nmode,cutoff,rdtype', [(6, 10, torch.float32), (6, 10, torch.float64), (5, 7, torch.float64),
                                                 (7, 4, torch.float32)])
def test_fused_fock_passes_against_host_contraction(nmode, cutoff, rdtype):
    """Config-5 circuit (squeezers + Clements mesh) through the FUSED Fock passes (b200q_qudit_fused) against the
    oracle's qudit contraction on the host (numpy restatement of evolve_state with qudit = cutoff, qmath.py:485-506)
    with the same matrices, and against the one-gate-per-pass kernel."""
    import statevec_oracle as so
    from deepquantum_b200 import photonic as ph
    cir = _c5_circuit(nmode, cutoff, rdtype)
    old = ph.FUSE_FOCK
    ph.FUSE_FOCK = True
    try:
        out = cir().reshape(-1).cpu().numpy().astype(np.complex128)
    finally:
        ph.FUSE_FOCK = old
    stats = cir.fock_plan_stats()
    assert 0 < stats['passes'] < stats['gates']
    cdt = torch.complex128 if rdtype == torch.float64 else torch.complex64
    mats = cir.build_matrices(cdt, 'cuda')
    psi = np.zeros((1, cutoff**nmode), dtype=np.complex128)
    psi[0, 0] = 1
    for op, m in zip(cir.operators, mats):
        psi = so.evolve_state(psi, m.cpu().numpy().astype(np.complex128), nmode, list(op.wires), cutoff)
    ref = psi.reshape(-1)
    tol = 1e-12 if rdtype == torch.float64 else 2e-6
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < tol
    old = ph.FUSE_FOCK
    ph.FUSE_FOCK = False
    try:
        unfused = cir().reshape(-1).cpu().numpy().astype(np.complex128)
    finally:
        ph.FUSE_FOCK = old
    assert np.linalg.norm(out - unfused) / np.linalg.norm(ref) < tol


@pytest.mark.gpu
def test_fused_fock_full_size_c5_matches_per_gate_kernel():
    """BASELINE config 5 at full size (8 modes, cutoff 10, 10^8 amplitudes): fused passes against the per-gate kernel
    (itself pinned by the reference fixtures at smaller sizes), and the norm."""
    from deepquantum_b200 import photonic as ph
    cir = _c5_circuit(8, 10, torch.float32)
    ph_old = ph.FUSE_FOCK
    ph.FUSE_FOCK = True
    try:
        out = cir().reshape(-1).clone()
    finally:
        ph.FUSE_FOCK = ph_old
    ph.FUSE_FOCK = False
    try:
        ref = cir().reshape(-1)
    finally:
        ph.FUSE_FOCK = ph_old
    err = float((out - ref).norm() / ref.norm())
    assert err < 2e-6, err
    assert abs(float((out.real**2 + out.imag**2).sum()) - 1) < 1e-3     # truncated Fock space: squeezing leaks a little
    # Size-independent physics of the circuit, checked on all 10^8 amplitudes against HOST arithmetic: the squeezers
    # create photons in pairs and every (truncated) beamsplitter matrix is block-diagonal in the photon number of its
    # two modes, complete -- hence unitary -- for sectors of at most cutoff - 1 photons.  So (a) no amplitude with an
    # odd total photon number, and (b) the distribution P(N) for N <= 9 after the whole mesh equals the convolution of
    # the single-mode squeezed-vacuum distributions (from the same 10 x 10 squeezer matrices, on the host).
    d, n = 10, 8
    dig = torch.arange(d, device='cuda')
    total = torch.zeros([d] * n, dtype=torch.int64, device='cuda')
    for m in range(n):
        shape = [1] * n
        shape[m] = d
        total = total + dig.reshape(shape)
    prob = (ref.real.double()**2 + ref.imag.double()**2)
    p_n = torch.zeros(n * (d - 1) + 1, dtype=torch.float64, device='cuda').index_add_(0, total.reshape(-1), prob).cpu().numpy()
    assert p_n[1::2].max() < 1e-12
    mats = cir.build_matrices(torch.complex128, 'cuda')
    conv = np.ones(1)
    for op, mtx in zip(cir.operators, mats):
        if len(op.wires) == 1:
            conv = np.convolve(conv, np.abs(mtx.cpu().numpy()[:, 0])**2)
    assert np.abs(p_n[:d] - conv[:d]).max() < 2e-6, (p_n[:d], conv[:d])


@pytest.mark.gpu
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('nmode,cutoff', [(4, 5), (3, 10), (3, 16), (5, 3)])
def test_structured_fock_kernels_against_host_contraction(nmode, cutoff, rdtype):
    """Every block structure of b200q_qudit_apply_structured (DIAG, DENSE1, NUMBER, DIFFERENCE) on every mode position
    (first / middle / last, both wire orders), batch of 2, against the oracle's qudit contraction on the host and against
    the generic ELL kernel; cutoff 17 exercises the documented fall-back to the generic kernel."""
    import statevec_oracle as so
    from deepquantum_b200 import _lib as L
    from deepquantum_b200 import photonic as ph
    n, d = nmode, cutoff
    cdt = torch.complex128 if rdtype == torch.float64 else torch.complex64
    g = torch.Generator().manual_seed(11)
    rnd = lambda s=1.0: float(torch.rand(1, generator=g) * s)   # noqa: E731
    gates = []
    for w in range(n):
        gates.append(ph.Squeezing([rnd(0.4), rnd(6)], n, [w], d))
        gates.append(ph.PhaseShift(rnd(6), n, [w], d))
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    for a, b in pairs:
        gates.append(ph.BeamSplitter([rnd(6), rnd(6)], n, [a, b], d))
    for a, b in pairs[::3]:
        gates.append(ph.Squeezing2([rnd(0.3), rnd(6)], n, [a, b], d))
        gates.append(ph.CrossKerr(rnd(1), n, [a, b], d))
        gates.append(ph.MZI([rnd(6), rnd(6)], n, [b, a], d))
    gates.append(ph.Displacement([rnd(0.3), rnd(6)], n, [n - 1], d))
    gates.append(ph.Kerr(rnd(1), n, [0], d))
    psi = torch.randn(2, d**n, generator=g, dtype=torch.float64) + 1j * torch.randn(2, d**n, generator=g,
                                                                                     dtype=torch.float64)
    psi = (psi / psi.norm(dim=1, keepdim=True)).to(cdt)
    ref = psi.numpy().astype(np.complex128)
    st = psi.to('cuda').contiguous()
    gen = st.clone()
    kinds = set()
    for op in gates:
        m = op.update_matrix_state().reshape(d**len(op.wires), d**len(op.wires)).to(cdt)
        kinds.add(op._structure)
        ph.qudit_apply_(st, n, d, m.to('cuda'), op.wires, 2, op._structure)
        ph.qudit_apply_(gen, n, d, m.to('cuda'), op.wires, 2, L.QUDIT_GENERAL)
        ref = so.evolve_state(ref, m.numpy().astype(np.complex128), n, list(op.wires), d)
    assert kinds == {L.QUDIT_DIAG, L.QUDIT_DENSE1, L.QUDIT_NUMBER, L.QUDIT_DIFFERENCE}
    scale = np.linalg.norm(ref)
    tol = 1e-11 if rdtype == torch.float64 else 2e-5
    err = np.linalg.norm(st.cpu().numpy() - ref) / scale
    err_gen = np.linalg.norm(gen.cpu().numpy() - ref) / scale
    assert err < tol and err_gen < tol, (err, err_gen)


@pytest.mark.gpu
@pytest.mark.parametrize('cutoff', [2, 3, 10, 16])
def test_native_fock_matrices_match_the_torch_recurrences(cutoff):
    """b200q_fock_bs_matrix / b200q_fock_squeezing_matrix (one launch per gate class) against the batched torch
    restatement of the reference recurrences (photonic/gate.py:347-374, 1091-1114), itself pinned by the reference
    fixtures (fock.npz, fock2.npz), in complex128 on the device -- including theta = 0 / r = 0 (exact zeros)."""
    from deepquantum_b200 import photonic as ph
    g = torch.Generator().manual_seed(2)
    n = 9
    th = torch.rand(n, generator=g, dtype=torch.float64) * 6
    phi = torch.rand(n, generator=g, dtype=torch.float64) * 6
    th[0] = 0.0
    r = torch.rand(n, generator=g, dtype=torch.float64) * 0.8
    r[1] = 0.0
    th, phi, r = th.cuda(), phi.cuda(), r.cuda()
    u = ph.BeamSplitter.mixing_matrix(th, phi)
    old = ph.NATIVE_FOCK_MATRICES
    try:
        ph.NATIVE_FOCK_MATRICES = True
        bs_n, sq_n = ph.bs_matrix_state(u, cutoff), ph.squeezing_matrix_state(r, phi, cutoff)
        mzi_n = ph.bs_matrix_state(ph.MZI.mixing_matrix(th, phi, False), cutoff)
        ph.NATIVE_FOCK_MATRICES = False
        bs_t, sq_t = ph.bs_matrix_state(u, cutoff), ph.squeezing_matrix_state(r, phi, cutoff)
        mzi_t = ph.bs_matrix_state(ph.MZI.mixing_matrix(th, phi, False), cutoff)
    finally:
        ph.NATIVE_FOCK_MATRICES = old
    for a, b in ((bs_n, bs_t), (sq_n, sq_t), (mzi_n, mzi_t)):
        assert a.shape == b.shape and a.dtype == b.dtype
        # both run the same recurrence in float64; its rounding noise grows with the cutoff (6e-13 at 16)
        assert float((a - b).abs().max()) < max(1e-13, 1e-15 * cutoff**4) * max(1.0, float(b.abs().max()))
        assert torch.equal(a == 0, b == 0)          # the structural zeros are exact in both


def test_fock_group_planner_keeps_the_gate_order():
    """`plan_fock_groups`: every gate exactly once, and any two gates that share a mode keep their circuit order
    (random circuits of structured / unstructured one- and two-mode gates)."""
    from deepquantum_b200 import _lib as L
    from deepquantum_b200 import photonic as ph
    rng = np.random.default_rng(4)
    sizes = []
    for trial in range(60):
        n = int(rng.integers(2, 7))
        info = []
        for _ in range(int(rng.integers(1, 40))):
            if rng.integers(3) == 0:
                a, b = rng.choice(n, size=2, replace=False)
                info.append(([int(a), int(b)], int(rng.choice([L.QUDIT_NUMBER, L.QUDIT_DIFFERENCE, L.QUDIT_DIAG, L.QUDIT_GENERAL]))))
            else:
                info.append(([int(rng.integers(n))], int(rng.choice([L.QUDIT_DENSE1, L.QUDIT_DIAG, L.QUDIT_GENERAL]))))
        groups = ph.plan_fock_groups(info, n, 4)
        pos = {i: (gi, k) for gi, grp in enumerate(groups) for k, i in enumerate(grp)}
        assert sorted(pos) == list(range(len(info)))
        for i in range(len(info)):
            for j in range(i + 1, len(info)):
                if set(info[i][0]) & set(info[j][0]):
                    assert pos[i] < pos[j], (trial, i, j)
        for grp in groups:
            assert len(grp) <= ph.GROUP_MAX_OPS
            if len(grp) > 1:
                two = [i for i in grp if len(info[i][0]) == 2]
                assert len(two) == 1 and all(info[i][1] != L.QUDIT_GENERAL for i in grp)
                assert all(set(info[i][0]) <= set(info[two[0]][0]) for i in grp)
            sizes.append(len(grp))
    assert max(sizes) >= 3


@pytest.mark.gpu
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('nmode,cutoff', [(4, 5), (3, 10), (5, 3), (3, 16)])
def test_grouped_fock_gates_against_host_contraction(nmode, cutoff, rdtype):
    """Groups of a two-mode gate with the one-mode gates around it (b200q_qudit_apply_group: squeezers / displacements /
    phase shifters / Kerr before and after beamsplitters, MZIs, two-mode squeezers and cross-Kerr gates, on every mode
    pair incl. the lowest mode and reversed wires) through `QumodeCircuit.forward` against the oracle's contraction."""
    import statevec_oracle as so
    from deepquantum_b200 import photonic as ph
    n, d = nmode, cutoff
    g = torch.Generator().manual_seed(9)
    rnd = lambda s=1.0: float(torch.rand(1, generator=g) * s)   # noqa: E731
    cir = dq.QumodeCircuit(n, [(0.6, [1] + [0] * (n - 1)), (0.8, [0] * (n - 1) + [2 if d > 2 else 1])], cutoff=d)
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    for k, (a, b) in enumerate(pairs):
        cir.s(a, rnd(0.3), rnd(6))
        cir.ps(b, rnd(6))
        if k % 3 == 0:
            cir.d(b, rnd(0.2), rnd(6))
        if k % 4 == 0:
            cir.bs([a, b], [rnd(6), rnd(6)])
        elif k % 4 == 1:
            cir.mzi([a, b], [rnd(6), rnd(6)])
        elif k % 4 == 2:
            cir.s2([a, b], rnd(0.2), rnd(6))
        else:
            cir.ck([a, b], rnd(1))
        cir.k(a, rnd(1))
        cir.ps(b, rnd(6))
    cir.to('cuda', rdtype)
    out = cir().reshape(-1).cpu().numpy().astype(np.complex128)
    stats = cir.fock_plan_stats()
    assert stats['passes'] < stats['gates'] and max(stats['gates_per_pass']) >= 3
    cdt = torch.complex128 if rdtype == torch.float64 else torch.complex64
    mats = cir.build_matrices(cdt, 'cuda')
    psi = cir.init_state.state.reshape(1, -1).cpu().numpy().astype(np.complex128)
    for op, m in zip(cir.operators, mats):
        psi = so.evolve_state(psi, m.cpu().numpy().astype(np.complex128), n, list(op.wires), d)
    ref = psi.reshape(-1)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < (1e-11 if rdtype == torch.float64 else 3e-5), err
    old = ph.GROUP_FOCK
    ph.GROUP_FOCK = False
    try:
        cir.__dict__['_group_key'] = None
        single = cir().reshape(-1).cpu().numpy().astype(np.complex128)
    finally:
        ph.GROUP_FOCK = old
        cir.__dict__['_group_key'] = None
    assert np.linalg.norm(single - ref) / np.linalg.norm(ref) < (1e-11 if rdtype == torch.float64 else 3e-5)
