"""numpy restatement of the reference gate matrices (gate.py) and of the circuit "spec" lowering.

TEST INFRASTRUCTURE ONLY (see statevec_oracle.py).  Pinned by `tests/golden/gate_matrices.npz`,
which `oracle/make_golden.py` dumped from the unmodified reference.

A circuit *spec* is a JSON-able list of entries ``{"g": name, "w": [wires], "c": [controls],
"p": [params]}`` whose names are the reference's builder methods (circuit.py:899-1537).  The same
spec is replayed on the reference (make_golden.py), on the product package (tests) and lowered here
to the ``(matrix, wires, controls)`` triples that `Gate.op_state` consumes (operation.py:191-197).
"""
from __future__ import annotations

import numpy as np

SQ2 = 2**0.5


def _c(x):
    return np.asarray(x, dtype=np.complex128)


def _c32(x):
    """The reference registers its constant matrices as complex64 buffers (gate.py:841-1367) and
    `cir.to(torch.double)` only widens them (operation.py:156-169): the complex128 path therefore
    carries float32-rounded constants.  Reproduced here."""
    return np.asarray(x, dtype=np.complex64).astype(np.complex128)


def f32(x):
    """Gate parameters are stored as float32 tensors unless a tensor is passed
    (`inputs_to_tensor`, gate.py:384-391)."""
    return float(np.float32(x))


# --- constant gates: gate.py:841, 916, 995, 1069, 1143, 1233, 1303, 1367 -------------------------
X = _c([[0, 1], [1, 0]])
Y = _c([[0, -1j], [1j, 0]])
Z = _c([[1, 0], [0, -1]])
H = _c32(np.asarray([[1, 1], [1, -1]], dtype=np.complex64) / np.float32(SQ2))
S = _c([[1, 0], [0, 1j]])
SDG = _c([[1, 0], [0, -1j]])
T = _c32([[1, 0], [0, (1 + 1j) / SQ2]])
TDG = _c32([[1, 0], [0, (1 - 1j) / SQ2]])
# gate.py:1934, 2006, 2069
CNOT = _c([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
SWAP = _c([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
ISWAP = _c([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])


def _perm_matrix(dim, swaps):
    m = np.eye(dim, dtype=np.complex128)
    for a, b in swaps:
        m[[a, b]] = m[[b, a]]
    return m


TOFFOLI = _perm_matrix(8, [(6, 7)])  # gate.py:2521-2536
FREDKIN = _perm_matrix(8, [(5, 6)])  # gate.py:2691-2706


# --- parametric gates ----------------------------------------------------------------------------
def u3(theta, phi, lambd):  # gate.py:594-602
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return _c([[c, -np.exp(1j * lambd) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lambd)) * c]])


def p(theta):  # PhaseShift, gate.py:737-742
    return _c([[1, 0], [0, np.exp(1j * theta)]])


def rx(theta):  # gate.py:1443-1448
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return _c([[c, -1j * s], [-1j * s, c]])


def ry(theta):  # gate.py:1538-1543
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return _c([[c, -s], [s, c]])


def rz(theta):  # gate.py:1634-1639
    return _c([[np.exp(-1j * theta / 2), 0], [0, np.exp(1j * theta / 2)]])


def j(theta, plane='xy'):  # ProjectionJ, gate.py:1751-1767
    if plane in ('xy', 'yx'):
        e = np.exp(-1j * theta)
        return _c([[1, e], [1, -e]]) / SQ2
    if plane in ('yz', 'zy'):
        cps = np.cos(theta / 2) + np.sin(theta / 2)
        cms = np.cos(theta / 2) - np.sin(theta / 2)
        return _c([[cps, -1j * cms], [cms, 1j * cps]]) / SQ2
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return _c([[c, s], [s, -c]])


def rxx(theta):  # gate.py:2139-2146
    c, s = np.cos(theta / 2), -1j * np.sin(theta / 2)
    return _c([[c, 0, 0, s], [0, c, s, 0], [0, s, c, 0], [s, 0, 0, c]])


def ryy(theta):  # gate.py:2212-2219
    c, s = np.cos(theta / 2), 1j * np.sin(theta / 2)
    return _c([[c, 0, 0, s], [0, c, -s, 0], [0, -s, c, 0], [s, 0, 0, c]])


def rzz(theta):  # gate.py:2295-2300
    em, ep = np.exp(-1j * theta / 2), np.exp(1j * theta / 2)
    return np.diag(_c([em, ep, ep, em]))


def rxy(theta):  # gate.py:2366-2373
    c, s = np.cos(theta / 2), -1j * np.sin(theta / 2)
    return _c([[1, 0, 0, 0], [0, c, s, 0], [0, s, c, 0], [0, 0, 0, 1]])


def rbs(theta):  # ReconfigurableBeamSplitter, gate.py:2455-2462
    c, s = np.cos(theta), np.sin(theta)
    return _c([[1, 0, 0, 0], [0, c, s, 0], [0, -s, c, 0], [0, 0, 0, 1]])


CONST_1Q = {'x': X, 'y': Y, 'z': Z, 'h': H, 's': S, 'sdg': SDG, 't': T, 'tdg': TDG}
PARAM_1Q = {'u3': u3, 'p': p, 'rx': rx, 'ry': ry, 'rz': rz, 'j': j}
CONST_2Q = {'swap': SWAP, 'iswap': ISWAP}
PARAM_2Q = {'rxx': rxx, 'ryy': ryy, 'rzz': rzz, 'rxy': rxy, 'rbs': rbs}
# builders whose first argument(s) are controls of a single-qubit gate (circuit.py:922-1197)
CONTROLLED_ALIAS = {
    'cx': 'x', 'cy': 'y', 'cz': 'z', 'ch': 'h', 'cs': 's', 'csdg': 'sdg', 'ct': 't', 'ctdg': 'tdg',
    'crx': 'rx', 'cry': 'ry', 'crz': 'rz', 'cp': 'p', 'cu': 'u3', 'ccx': 'x',
}
CONTROLLED_ALIAS_2Q = {'crxx': 'rxx', 'cryy': 'ryy', 'crzz': 'rzz', 'crxy': 'rxy', 'cswap': 'swap'}


def lower_entry(entry, nqubit):
    """One spec entry -> list of (matrix, wires, controls) (layers expand to several gates)."""
    g = entry['g']
    w = list(entry.get('w', []))
    c = list(entry.get('c', []))
    prm = entry.get('p', [])
    if not entry.get('exact', False):
        prm = [f32(x) for x in prm]
    if g == 'barrier':  # gate.py:3112-3114, a no-op
        return []
    if g in CONTROLLED_ALIAS:  # w = [controls..., target]
        return lower_entry({'g': CONTROLLED_ALIAS[g], 'w': w[-1:], 'c': w[:-1], 'p': prm, 'exact': True}, nqubit)
    if g in CONTROLLED_ALIAS_2Q:  # w = [control, t1, t2]
        return lower_entry({'g': CONTROLLED_ALIAS_2Q[g], 'w': w[-2:], 'c': w[:-2], 'p': prm, 'exact': True}, nqubit)
    if g in CONST_1Q:
        return [(CONST_1Q[g], w, c)]
    if g in PARAM_1Q:
        if g == 'j':
            return [(j(prm[0], entry.get('plane', 'xy')), w, c)]
        return [(PARAM_1Q[g](*prm), w, c)]
    if g == 'cnot':  # dense 4x4 on [control, target], circuit.py:1179-1182
        return [(CNOT, w, c)]
    if g in CONST_2Q:
        return [(CONST_2Q[g], w, c)]
    if g in PARAM_2Q:
        return [(PARAM_2Q[g](*prm), w, c)]
    if g == 'toffoli':
        return [(TOFFOLI, w, c)]
    if g == 'fredkin':
        return [(FREDKIN, w, c)]
    if g == 'any':
        u = np.asarray(entry['u_re']) + 1j * np.asarray(entry['u_im'])
        return [(u, w, c)]
    if g == 'hamiltonian':  # exp(-i H t), gate.py:2952-2994
        from scipy.linalg import expm
        if 'ham' in entry:
            # the Pauli sum is accumulated in complex64 (gate.py:2962-2978), over the wires min..max it names
            paulis = {'x': X.astype(np.complex64), 'y': Y.astype(np.complex64), 'z': Z.astype(np.complex64)}
            pairs = entry['ham']
            if len(pairs) == 2 and isinstance(pairs[1], str):   # a single [coeff, string] pair, gate.py:2931-2932
                pairs = [pairs]
            named = [int(i) for _, s_ in pairs for i in s_[1::2]]
            lo, hi = min(named), max(named)
            ham = None
            for coeff, string in pairs:
                lst = [np.eye(2, dtype=np.complex64)] * nqubit
                for wire, key in zip(string[1::2], string[::2]):
                    lst[int(wire)] = paulis[key.lower()]
                term = lst[lo]
                for m in lst[lo + 1:hi + 1]:
                    term = np.kron(term, m)
                term = (term * np.float32(coeff)).astype(np.complex64)
                ham = term if ham is None else (ham + term).astype(np.complex64)
            w = list(range(lo, hi + 1))
        else:
            ham = np.asarray(entry['h_re']) + 1j * np.asarray(entry['h_im'])
        return [(expm(-1j * ham.astype(np.complex128) * prm[0]), w, c)]
    # ---- layers (layer.py:204-483): one gate per wire -------------------------------------------
    if g in ('xlayer', 'ylayer', 'zlayer', 'hlayer'):
        ws = w if w else list(range(nqubit))
        return [(CONST_1Q[g[0]], [q], []) for q in ws]
    if g in ('rxlayer', 'rylayer', 'rzlayer'):
        ws = w if w else list(range(nqubit))
        return [(PARAM_1Q[g[:2]](prm[i]), [q], []) for i, q in enumerate(ws)]
    if g == 'u3layer':
        ws = w if w else list(range(nqubit))
        return [(u3(*prm[3 * i:3 * i + 3]), [q], []) for i, q in enumerate(ws)]
    if g == 'cxlayer':  # layer.py:412-443, dense CNOTs on the given pairs
        return [(CNOT, list(pair), []) for pair in entry['pairs']]
    if g == 'cnot_ring':  # layer.py:446-483
        lo, hi = entry.get('minmax') or [0, nqubit - 1]
        step = entry.get('step', 1)
        reverse = entry.get('reverse', False)
        nw = hi - lo + 1
        if reverse:
            pairs = [[lo + i, lo + (i - step) % nw] for i in range(nw - 1, -1, -1)]
        else:
            pairs = [[lo + i, lo + (i + step) % nw] for i in range(nw)]
        return [(CNOT, pr, []) for pr in pairs]
    raise ValueError(f'unknown spec entry {g!r}')


def lower_spec(spec, nqubit):
    ops = []
    for entry in spec:
        ops.extend(lower_entry(entry, nqubit))
    return ops
