"""CPU oracle for the statevector gate-application path (numpy, complex128 by default).

TEST INFRASTRUCTURE ONLY -- this is the *checker*, never the product.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
it.  The product package `deepquantum_b200` never does, and fails loudly without its CUDA library.

Every function restates one function of the reference (TuringQ/deepquantum @ 727c44d, v4.5.0) and
cites it.  The arithmetic the reference delegates to ATen (`permute`/`reshape`/`mm`/`cat`, torch
>= 2.4 per pyproject.toml:10, source not under /root/reference) is a plain complex matrix product,
restated here with `numpy.tensordot`.

Parity pin: `tests/test_oracle_golden.py` checks these functions against fixtures in
`tests/golden/*.npz` that `oracle/make_golden.py` produced by running the unmodified reference in
the build container (every gate family, the five BASELINE configs at reduced size, gradients, the
sharded path on 2/4/8 gloo ranks, the Fock tensor path).

Conventions (reference `operation.py:45-55`, `distributed.py:18`): a state of n qudits of dimension
d is a flat vector of d**n amplitudes; wire 0 is the MOST significant digit.  For qubits, wire w is
bit (n-1-w) of the flat index.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------------------
# qmath.py:485-506  evolve_state
# ----------------------------------------------------------------------------------------------
def evolve_state(state: np.ndarray, matrix: np.ndarray, nqudit: int, wires, qudit: int = 2) -> np.ndarray:
    """psi' = (matrix on `wires`) psi.

    Restates `qmath.evolve_state` (qmath.py:485-506): the wires are moved to the front, the state is
    viewed as a (d^k, d^(n-k)) matrix, left-multiplied, and moved back.  The matrix row/column
    index enumerates `wires` in the given order with wires[0] as the most significant digit.
    `state` is flat (d**n,) or batched (batch, d**n); the result has the same shape.
    """
    wires = list(wires)
    k = len(wires)
    batched = state.ndim == 2
    psi = state.reshape((-1,) + (qudit,) * nqudit)
    axes = [w + 1 for w in wires]
    mat = np.asarray(matrix).reshape((qudit,) * (2 * k))
    # contract the column indices of `mat` with the target axes of psi
    out = np.tensordot(mat, psi, axes=(list(range(k, 2 * k)), axes))
    # tensordot puts the k new axes first, then batch, then the remaining wires in order
    rest = [a for a in range(nqudit + 1) if a not in axes]
    src_order = axes + rest
    out = np.transpose(out, np.argsort(src_order))
    out = np.ascontiguousarray(out).reshape(-1, qudit**nqudit)
    return out if batched else out[0]


# ----------------------------------------------------------------------------------------------
# operation.py:203-219  Gate.op_state_control
# ----------------------------------------------------------------------------------------------
def evolve_state_controlled(state: np.ndarray, matrix: np.ndarray, nqubit: int, wires, controls) -> np.ndarray:
    """Controlled gate: `matrix` acts on `wires` only where every control wire is |1>.

    Restates `Gate.op_state_control` (operation.py:203-219), which permutes to
    [targets, rest, controls], multiplies only the last (all-ones) control slice and concatenates.
    """
    controls = list(controls)
    if not controls:
        return evolve_state(state, matrix, nqubit, wires)
    batched = state.ndim == 2
    psi = state.reshape((-1,) + (2,) * nqubit).copy()
    sel = [slice(None)] * (nqubit + 1)
    for c in controls:
        sel[c + 1] = 1
    sub = psi[tuple(sel)]
    # wires of the sub-tensor: the control axes were dropped, renumber the target wires
    remaining = [w for w in range(nqubit) if w not in controls]
    sub_wires = [remaining.index(w) for w in wires]
    nsub = nqubit - len(controls)
    bsz = sub.shape[0]
    new = evolve_state(np.ascontiguousarray(sub).reshape(bsz, -1), matrix, nsub, sub_wires)
    psi[tuple(sel)] = new.reshape(sub.shape)
    out = psi.reshape(-1, 2**nqubit)
    return out if batched else out[0]


def apply_op(state, op, nqubit):
    """Apply one lowered op `(matrix, wires, controls)` (the triple `Gate.op_state` consumes,
    operation.py:191-197)."""
    matrix, wires, controls = op
    return evolve_state_controlled(state, matrix, nqubit, wires, controls)


def run_circuit(ops, nqubit: int, state: np.ndarray | None = None, dtype=np.complex128) -> np.ndarray:
    """The `nn.Sequential` loop of `QubitCircuit._forward_helper` (circuit.py:244-263) over lowered
    ops.  Default initial state |0...0> (`QubitState`, state.py:31-44)."""
    if state is None:
        state = np.zeros(2**nqubit, dtype=dtype)
        state[0] = 1.0
    else:
        state = np.asarray(state, dtype=dtype)
        state = state.reshape(-1) if state.size == 2**nqubit else state.reshape(-1, 2**nqubit)
    for op in ops:
        state = apply_op(state, (np.asarray(op[0], dtype=dtype), op[1], op[2]), nqubit)
    return state


# ----------------------------------------------------------------------------------------------
# qmath.py:830-860 expectation / layer.py:127-165 Observable
# ----------------------------------------------------------------------------------------------
_PAULI = {
    'x': np.array([[0, 1], [1, 0]], dtype=np.complex128),
    'y': np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    'z': np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def expectation_pauli(state: np.ndarray, nqubit: int, wires, basis: str) -> float:
    """<psi| P_wires |psi> for a Pauli string: apply the Pauli gates wire by wire
    (layer.py:156-166) and take `state.mH @ (O state)` (qmath.py:858)."""
    wires = list(wires)
    if len(basis) == 1:
        basis = basis * len(wires)
    phi = state
    for w, b in zip(wires, basis):
        phi = evolve_state(phi, _PAULI[b.lower()], nqubit, [w])
    return float(np.real(np.vdot(state, phi)))


def inner_product(bra: np.ndarray, ket: np.ndarray) -> complex:
    """<bra|ket> (`inner_product_dist`, distributed.py:288-294, on one shard)."""
    return complex(np.vdot(bra, ket))


# ----------------------------------------------------------------------------------------------
# adjoint.py:19-83  AdjointExpectation (dense restatement for a single observable list)
# ----------------------------------------------------------------------------------------------
def adjoint_gradient(ops, dops, nqubit, observables, weights=None, dtype=np.complex128):
    """Gradient of sum_k w_k <psi|O_k|psi> with respect to the parameter of every op that has a
    derivative matrix, by the adjoint method (adjoint.py:47-83).

    ops:  list of (matrix, wires, controls); dops: list of dU/dtheta (same shape as matrix) or None.
    observables: list of (wires, basis).  Returns a list of floats (None where dops[i] is None).
    """
    psi = run_circuit(ops, nqubit, dtype=dtype)
    if weights is None:
        weights = [1.0] * len(observables)
    lam = np.zeros_like(psi)
    for (wires, basis), wgt in zip(observables, weights):
        phi = psi
        if len(basis) == 1:
            basis = basis * len(wires)
        for w, b in zip(wires, basis):
            phi = evolve_state(phi, _PAULI[b.lower()], nqubit, [w])
        lam = lam + wgt * phi
    grads = [None] * len(ops)
    for i in range(len(ops) - 1, -1, -1):
        matrix, wires, controls = ops[i]
        udag = np.asarray(matrix, dtype=dtype).conj().T
        psi = evolve_state_controlled(psi, udag, nqubit, wires, controls)  # un-apply (adjoint.py:60)
        if dops[i] is not None:
            # a controlled gate's derivative acts only inside the all-ones control block
            mu = _apply_derivative(psi, np.asarray(dops[i], dtype=dtype), nqubit, wires, controls)
            grads[i] = 2.0 * float(np.real(np.vdot(lam, mu)))  # adjoint.py:66-73
        lam = evolve_state_controlled(lam, udag, nqubit, wires, controls)
    return grads


def _apply_derivative(state, dmat, nqubit, wires, controls):
    if not controls:
        return evolve_state(state, dmat, nqubit, wires)
    psi = state.reshape((2,) * nqubit)
    out = np.zeros_like(psi)
    sel = [slice(None)] * nqubit
    for c in controls:
        sel[c] = 1
    sub = np.ascontiguousarray(psi[tuple(sel)]).reshape(-1)
    remaining = [w for w in range(nqubit) if w not in controls]
    sub_wires = [remaining.index(w) for w in wires]
    new = evolve_state(sub, dmat, nqubit - len(controls), sub_wires)
    out[tuple(sel)] = new.reshape(psi[tuple(sel)].shape)
    return out.reshape(-1)


# ----------------------------------------------------------------------------------------------
# distributed.py  -- sharded layout helpers (rank = high-order bits, state.py:358-360)
# ----------------------------------------------------------------------------------------------
def shard(state: np.ndarray, world_size: int, rank: int) -> np.ndarray:
    """Rank r holds flat indices [r*2^(n-g), (r+1)*2^(n-g)) (state.py:358-360)."""
    per = state.shape[-1] // world_size
    return state[..., rank * per:(rank + 1) * per]


def unshard(shards) -> np.ndarray:
    return np.concatenate(list(shards), axis=-1)


# ----------------------------------------------------------------------------------------------
# gate.py:3027-3094  Reset.op_state (postselect 0 / 1)
# ----------------------------------------------------------------------------------------------
def reset_wires(state: np.ndarray, nqubit: int, wires, postselect: int = 0) -> np.ndarray:
    """Reset `wires` to |0>: per wire, keep the `postselect` branch divided by the square root of its probability (the
    OTHER branch, undivided, when that probability is exactly 0) and relabel it |0>.  All wires: |0...0>.
    `state` is flat (2**n,) or batched (batch, 2**n)."""
    batched = state.ndim == 2
    psi = state.reshape((-1,) + (2,) * nqubit).copy()
    if len(wires) == nqubit:
        out = np.zeros_like(psi.reshape(psi.shape[0], -1))
        out[:, 0] = 1
        return out if batched else out[0]
    for w in wires:
        x = np.moveaxis(psi, w + 1, 0)                                # (2, batch, ...)
        probs = (np.abs(x) ** 2).reshape(2, x.shape[1], -1).sum(-1)   # (2, batch)
        mask = 1 - np.sign(probs[postselect])
        norm = np.sqrt(probs[postselect] + mask)
        shape = (-1,) + (1,) * (nqubit - 1)
        s0 = ((1 - mask).reshape(shape) * x[postselect] + mask.reshape(shape) * x[1 - postselect]) / norm.reshape(shape)
        psi = np.moveaxis(np.stack([s0, np.zeros_like(s0)]), 0, w + 1)
    out = np.ascontiguousarray(psi).reshape(psi.shape[0], -1)
    return out if batched else out[0]
