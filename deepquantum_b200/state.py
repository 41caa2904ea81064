"""QubitState with the reference's interface (state.py:14-78), state vectors and density matrices."""
from __future__ import annotations

from typing import Any

import torch
from torch import nn

from .operation import apply_complex_fix


def amplitude_encoding(data: Any, nqubit: int) -> torch.Tensor:
    """Normalised amplitude vector(s) `[batch, 2^n, 1]` from data (reference qmath.py:455-482)."""
    if not isinstance(data, (torch.Tensor, nn.Parameter)):
        data = torch.tensor(data)
    single = data.ndim == 1 or (data.ndim == 2 and data.shape[-1] == 1)
    batch = 1 if single else data.shape[0]
    data = data.reshape(batch, -1)
    size = data.shape[1]
    n = 2**nqubit
    state = torch.zeros(batch, n, dtype=data.dtype, device=data.device) + 0j
    data = nn.functional.normalize(data[:, :n], p=2, dim=-1)
    if n > size:
        state[:, :size] = data[:, :]
    else:
        state[:, :] = data[:, :]
    return state.unsqueeze(-1)


def is_power_of_two(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


def is_density_matrix(rho: torch.Tensor) -> bool:
    """Hermitian, trace one, positive semi-definite; 2-D or batched 3-D (reference qmath.py:117-153)."""
    if not isinstance(rho, torch.Tensor) or rho.ndim not in (2, 3):
        return False
    if not (is_power_of_two(rho.shape[-2]) and is_power_of_two(rho.shape[-1])) or rho.shape[-1] != rho.shape[-2]:
        return False
    if rho.ndim == 2:
        rho = rho.unsqueeze(0)
    if not torch.allclose(rho, rho.mH):
        return False
    trace = rho.diagonal(dim1=-2, dim2=-1).sum(-1)
    if not torch.allclose(trace, torch.ones_like(trace)):
        return False
    tol = 1e-6 if rho.dtype == torch.complex64 else 1e-10
    return bool(torch.all(torch.linalg.eigvalsh(rho) > -tol))


LAZY_NQUBIT = 24  # named initial states above this size are materialised on demand only


class QubitState(nn.Module):
    """A pure state of n qubits: `'zeros'`, `'equal'`, `'entangle'/'GHZ'/'ghz'`, or amplitude data
    (reference state.py:14-78).  Named states of more than LAZY_NQUBIT qubits are not stored: `state`
    is built on access, on the dtype/device the module was moved to (a 30-qubit complex128 `'zeros'`
    buffer would pin 16 GiB that the engine never reads -- it fills |0...0> with a kernel)."""

    def __init__(self, nqubit: int = 1, state: Any = 'zeros', den_mat: bool = False) -> None:
        super().__init__()
        self.nqubit = nqubit
        self.den_mat = den_mat
        self.kind = state if isinstance(state, str) else 'data'
        if isinstance(state, str) and state not in ('zeros', 'equal', 'entangle', 'GHZ', 'ghz'):
            raise ValueError(f'unknown initial state {state!r}')
        self._lazy = isinstance(state, str) and nqubit > (LAZY_NQUBIT // 2 if den_mat else LAZY_NQUBIT)
        self.register_buffer('_probe', torch.zeros(1, dtype=torch.cfloat), persistent=False)
        if self._lazy:
            return
        if isinstance(state, str):
            vec = self._named(state, nqubit, torch.cfloat, 'cpu')
        else:
            if not isinstance(state, torch.Tensor):
                state = torch.tensor(state, dtype=torch.cfloat)
            ndim = state.ndim
            if den_mat and state.shape[-1] == 2**nqubit and is_density_matrix(state):
                self.register_buffer('state', state)
                return
            vec = amplitude_encoding(data=state, nqubit=nqubit)
            if vec.ndim > ndim:
                vec = vec.squeeze(0)
        if den_mat:
            vec = vec @ vec.mH
        self.register_buffer('state', vec)

    @staticmethod
    def _named(kind, nqubit, dtype, device):
        if kind == 'zeros':
            vec = torch.zeros((2**nqubit, 1), dtype=dtype, device=device)
            vec[0] = 1
        elif kind == 'equal':
            vec = torch.full((2**nqubit, 1), 1.0, dtype=dtype, device=device) / (2**nqubit) ** 0.5
        else:
            vec = torch.zeros((2**nqubit, 1), dtype=dtype, device=device)
            vec[0] = 1 / 2**0.5
            vec[-1] = 1 / 2**0.5
        return vec

    def __getattr__(self, name):
        if name == 'state' and self.__dict__.get('_lazy', False):
            probe = self._buffers['_probe']
            vec = self._named(self.kind, self.nqubit, probe.dtype, probe.device)
            return vec @ vec.mH if self.den_mat else vec
        return super().__getattr__(name)

    @property
    def dtype(self) -> torch.dtype:
        return self._buffers['_probe'].dtype

    @property
    def device(self) -> torch.device:
        return self._buffers['_probe'].device

    def _apply(self, fn: Any) -> 'QubitState':
        tensors = {'_probe': self._buffers.pop('_probe')}
        if 'state' in self._buffers:
            tensors['state'] = self._buffers.pop('state')
        super()._apply(fn)
        for key, value in apply_complex_fix(fn, tensors).items():
            self.register_buffer(key, value, persistent=(key == 'state'))
        return self

    def forward(self) -> None:
        pass
