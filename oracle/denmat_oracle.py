"""CPU restatement (numpy) of the reference's density-matrix path.  TEST INFRASTRUCTURE ONLY: imported by
tests/, never by the product package.

Pinned by tests/golden/denmat.npz (written by oracle/make_golden.py from the UNMODIFIED reference).

  evolve_den_mat            reference src/deepquantum/qmath.py:509-540  (U on the row wires, conj(U) on the
                            column wires of rho reshaped to 2n two-level axes)
  controlled gates          reference src/deepquantum/operation.py:232-262 (the all-ones control slice on each side)
  channels                  reference src/deepquantum/operation.py:594-600 (sum over Kraus operators of
                            evolve_den_mat), Kraus operators src/deepquantum/channel.py:46-55, 88-97, 138-149,
                            200-212, 252-263, 303-314, 367-383
  expectation               reference src/deepquantum/qmath.py:855-856  Tr(O rho)
  measurement probabilities reference src/deepquantum/qmath.py:600-602, 620  |diag(rho)|, marginalised
"""
import numpy as np

import gates_np
import statevec_oracle as so

_I = np.eye(2, dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)

CHANNELS = ('bit_flip', 'phase_flip', 'depolarizing', 'pauli', 'amp_damp', 'phase_damp', 'gen_amp_damp')


def kraus(name, params, exact=False):
    """Kraus operators of a channel builder (`prob = sin(theta)^2`, channel.py).  Parameters given as Python
    floats reach the reference as float32 (`inputs_to_tensor`), hence the rounding unless `exact`."""
    th = np.array([x if exact else gates_np.f32(x) for x in params], dtype=np.float64)
    p = np.sin(th) ** 2
    if name == 'bit_flip':
        return [np.sqrt(1 - p[0]) * _I, np.sqrt(p[0]) * _X]
    if name == 'phase_flip':
        return [np.sqrt(1 - p[0]) * _I, np.sqrt(p[0]) * _Z]
    if name == 'depolarizing':
        s = np.sqrt(p[0] / 3)
        return [np.sqrt(1 - p[0]) * _I, s * _X, s * _Y, s * _Z]
    if name == 'pauli':
        p = p / p.sum()
        return [np.sqrt(p[k]) * m for k, m in enumerate((_I, _X, _Y, _Z))]
    if name == 'amp_damp':
        return [np.array([[1, 0], [0, np.sqrt(1 - p[0])]], dtype=np.complex128),
                np.array([[0, np.sqrt(p[0])], [0, 0]], dtype=np.complex128)]
    if name == 'phase_damp':
        return [np.array([[1, 0], [0, np.sqrt(1 - p[0])]], dtype=np.complex128),
                np.array([[0, 0], [0, np.sqrt(p[0])]], dtype=np.complex128)]
    if name == 'gen_amp_damp':
        q, g = p[0], p[1]
        return [np.sqrt(q) * np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=np.complex128),
                np.sqrt(q) * np.array([[0, np.sqrt(g)], [0, 0]], dtype=np.complex128),
                np.sqrt(1 - q) * np.array([[np.sqrt(1 - g), 0], [0, 1]], dtype=np.complex128),
                np.sqrt(1 - q) * np.array([[0, 0], [np.sqrt(g), 0]], dtype=np.complex128)]
    raise ValueError(name)


def evolve_den_mat(rho, matrix, nqubit, wires, controls=()):
    """rho: flat 4^n vector (row index major).  Left-multiply on the row wires, conj on the column wires."""
    two_n = 2 * nqubit
    op = (np.asarray(matrix), list(wires), list(controls))
    rho = so.apply_op(rho, op, two_n)
    op = (np.asarray(matrix).conj(), [w + nqubit for w in wires], [c + nqubit for c in controls])
    return so.apply_op(rho, op, two_n)


def apply_channel(rho, ks, nqubit, wire):
    return sum(evolve_den_mat(rho, k, nqubit, [wire]) for k in ks)


def run_spec(spec, nqubit, rho=None):
    """Replay a circuit spec (workloads.apply_spec format, plus channel entries) on a density matrix.
    Returns rho as a [2^n, 2^n] complex128 array."""
    dim = 2**nqubit
    if rho is None:
        rho = np.zeros(dim * dim, dtype=np.complex128)
        rho[0] = 1
    else:
        rho = np.asarray(rho, dtype=np.complex128).reshape(-1).copy()
    for e in spec:
        if e['g'] in CHANNELS:
            rho = apply_channel(rho, kraus(e['g'], e['p']), nqubit, e['w'][0])
        else:
            for matrix, wires, controls in gates_np.lower_entry(e, nqubit):
                rho = evolve_den_mat(rho, matrix, nqubit, wires, controls)
    return rho.reshape(dim, dim)


def expectation_pauli(rho, nqubit, wires, basis):
    """Tr(P rho) for the Pauli string with `basis[i]` on `wires[i]`."""
    paulis = {'x': _X, 'y': _Y, 'z': _Z}
    basis = basis if len(basis) == len(wires) else basis * len(wires)
    v = np.asarray(rho, dtype=np.complex128).reshape(-1)
    for w, b in zip(wires, basis):
        v = so.apply_op(v, (paulis[b], [w], []), 2 * nqubit)   # left multiplication only
    dim = 2**nqubit
    return float(np.trace(v.reshape(dim, dim)).real)


def measure_probs(rho, nqubit, wires=None):
    """Probabilities of the outcomes of `wires` (ascending wire order, first wire = most significant key bit)."""
    p = np.abs(np.diagonal(np.asarray(rho)))
    if wires is None:
        return p
    wires = sorted(wires)
    rest = [q for q in range(nqubit) if q not in wires]
    return p.reshape([2] * nqubit).transpose(wires + rest).reshape(2**len(wires), -1).sum(-1)
