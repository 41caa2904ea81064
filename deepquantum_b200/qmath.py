"""Boundary functions with the reference's names (qmath.py): `evolve_state`, `expectation`,
`inverse_permutation`, `multi_kron`.  The contraction itself runs in libb200q.so."""
from __future__ import annotations

import torch

from . import _lib as L
from . import engine
from .state import amplitude_encoding  # noqa: F401  (re-exported like the reference)


def inverse_permutation(permute_shape: list[int]) -> list[int]:
    inv = [0] * len(permute_shape)
    for pos, axis in enumerate(permute_shape):
        inv[axis] = pos
    return inv


def multi_kron(lst: list[torch.Tensor]) -> torch.Tensor:
    out = lst[0]
    for m in lst[1:]:
        out = torch.kron(out, m)
    return out


def evolve_state(state: torch.Tensor, matrix: torch.Tensor, nqudit: int, wires: list[int],
                 qudit: int = 2) -> torch.Tensor:
    """Drop-in for `qmath.evolve_state` (reference qmath.py:485-506).

    `state` is `[batch, d, ..., d]` (any strides), `matrix` is `d^k x d^k` with `wires[0]` the most
    significant matrix digit.  Returns a new tensor of the same shape; the input is not modified."""
    shape = state.shape
    flat = state.reshape(-1, qudit**nqudit).contiguous().clone()
    batch = flat.shape[0]
    if qudit == 2:
        engine.apply_gate_(flat, nqudit, matrix, engine.wires_to_targets(nqudit, wires), (), L.GATE_MAT, False, batch)
    else:
        from .photonic import qudit_apply_
        qudit_apply_(flat, nqudit, qudit, matrix, wires, batch)
    return flat.reshape(shape)


def evolve_state_controlled(state: torch.Tensor, matrix: torch.Tensor, nqubit: int, wires: list[int],
                            controls: list[int]) -> torch.Tensor:
    """Drop-in for `Gate.op_state_control` (reference operation.py:203-219) as a free function."""
    shape = state.shape
    flat = state.reshape(-1, 2**nqubit).contiguous().clone()
    engine.apply_gate_(flat, nqubit, matrix, engine.wires_to_targets(nqubit, wires),
                       [nqubit - 1 - c for c in controls], L.GATE_MAT, False, flat.shape[0])
    return flat.reshape(shape)
