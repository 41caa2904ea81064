"""Photonic Fock-tensor path on CPU: the vectorised Fock transformation matrices against the reference's
recurrences (golden fock.npz), and the qudit kernel's index geometry (stepped by the TEST-ONLY emulator)
against the oracle's evolve_state with qudit = cutoff."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import statevec_oracle as so
from conftest import GOLDEN
from helpers import hostemu

import deepquantum_b200 as dq
from deepquantum_b200 import photonic as ph


def _g():
    return np.load(os.path.join(GOLDEN, 'fock.npz'))


@pytest.mark.parametrize('key', ['m2_c5', 'm3_c4', 'm4_c6', 'm5_c8'])
def test_fock_matrices_match_reference(key):
    g = _g()
    d = json.loads(str(g[key + '/spec']))['cutoff']
    f32 = lambda v: float(np.float32(v))   # noqa: E731  (the reference stores parameters as float32)
    s = ph.squeezing_matrix_state(torch.tensor([f32(0.31)], dtype=torch.float64),
                                  torch.tensor([f32(1.3)], dtype=torch.float64), d)[0].numpy()
    np.testing.assert_allclose(s, g[key + '/s_matrix'], atol=1e-13)
    u = ph.BeamSplitter.mixing_matrix(torch.tensor([f32(0.6)], dtype=torch.float64),
                                      torch.tensor([f32(2.2)], dtype=torch.float64))
    b = ph.bs_matrix_state(u, d)[0].numpy()
    np.testing.assert_allclose(b, g[key + '/bs_matrix'], atol=1e-13)


def _emu_qudit(state, nmode, d, matrix, wires):
    lib = hostemu()
    lib.hostemu_qudit.restype = C.c_int
    lib.hostemu_qudit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int,
                                  C.c_int64]
    st = np.ascontiguousarray(state).copy()
    m = np.ascontiguousarray(matrix.astype(st.dtype))
    w = (C.c_int32 * len(wires))(*wires)
    batch = st.shape[0] if st.ndim == 2 else 1
    rc = lib.hostemu_qudit(st.ctypes.data, nmode, d, 0 if st.dtype == np.complex64 else 1, m.ctypes.data, w, len(wires),
                           batch)
    assert rc == 0
    return st


@pytest.mark.parametrize('nmode,d', [(1, 4), (2, 3), (4, 5), (5, 3), (3, 10)])
def test_qudit_geometry_every_mode(nmode, d):
    rng = np.random.default_rng(nmode * 10 + d)
    psi = rng.normal(size=(2, d**nmode)) + 1j * rng.normal(size=(2, d**nmode))
    for w in range(nmode):
        m = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        ref = so.evolve_state(psi, m, nmode, [w], d)
        np.testing.assert_allclose(_emu_qudit(psi, nmode, d, m, [w]), ref, atol=1e-12)
    if nmode >= 2:
        for a in range(nmode):
            for b in range(nmode):
                if a == b:
                    continue
                m = rng.normal(size=(d * d, d * d)) + 1j * rng.normal(size=(d * d, d * d))
                ref = so.evolve_state(psi, m, nmode, [a, b], d)
                np.testing.assert_allclose(_emu_qudit(psi, nmode, d, m, [a, b]), ref, atol=1e-11)


@pytest.mark.parametrize('key', ['m2_c5', 'm3_c4', 'm4_c6'])
def test_fock_circuit_matches_reference(key):
    """Whole interferometer (squeezers + beamsplitter mesh + phase shifter): product matrices + emulated kernel
    geometry against the reference's final Fock state."""
    g = _g()
    meta = json.loads(str(g[key + '/spec']))
    n, d = meta['nmode'], meta['cutoff']
    cir = dq.QumodeCircuit(n, 'vac', cutoff=d, backend='fock', basis=False)
    for e in meta['spec']:
        if e['g'] == 's':
            cir.s(e['w'][0], e['p'][0], e['p'][1])
        elif e['g'] == 'bs':
            cir.bs(e['w'], e['p'])
        else:
            cir.ps(e['w'][0], e['p'][0])
    mats = cir.build_matrices(torch.complex128, 'cpu')
    st = np.zeros(d**n, dtype=np.complex128)
    st[0] = 1
    for op, m in zip(cir.operators, mats):
        st = _emu_qudit(st, n, d, m.numpy(), op.wires)
    ref = g[key + '/c128']
    assert np.linalg.norm(st - ref) / np.linalg.norm(ref) < 1e-12
