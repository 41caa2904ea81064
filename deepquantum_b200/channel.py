"""Quantum channels with the reference's interface (channel.py:16-383): same class names, parameters
(`prob = sin(theta)^2`) and Kraus operators.  On the device a channel is ONE dense gate on the (row, column) wire
pair of the density matrix -- the superoperator `sum_i K_i (x) conj(K_i)` (`operation.Channel._lowered_matrix`) --
inside the same fused passes as the unitary gates; the reference evolves one copy of the state per Kraus
operator and sums them (operation.py:594-600)."""
from __future__ import annotations

from typing import Any

import torch

from .operation import Channel


def _kraus(entries) -> torch.Tensor:
    """[[k00, k01, k10, k11], ...] of broadcastable real/complex tensors `[...]` (or the constants 0 / 1) ->
    `[..., n_kraus, 2, 2]` complex, with ONE stack over all entries (constants share one tensor: the assembly of a
    noisy circuit is launch-bound, every avoided small kernel counts)."""
    tensors = [e for k in entries for e in k if isinstance(e, torch.Tensor)]
    shape = torch.broadcast_shapes(*[e.shape for e in tensors])
    ref = tensors[0]
    cdtype = torch.complex64 if ref.dtype in (torch.float32, torch.complex64) else torch.complex128
    consts = {}

    def full(e):
        if isinstance(e, torch.Tensor):
            return e.to(cdtype).expand(shape)
        if e not in consts:
            consts[e] = torch.full(shape, e, dtype=cdtype, device=ref.device)
        return consts[e]
    flat = torch.stack([full(e) for k in entries for e in k], dim=-1)
    return flat.reshape(*shape, len(entries), 2, 2)


class _OneParam(Channel):
    """Built-in channels: `_kraus_of(prob)` gives the Kraus operators for `prob = sin(theta)^2` of shape
    `[..., npara]`, vectorised over leading dimensions, so that all channels of one class in a circuit are evaluated
    by ONE batched call per forward (`_batched_matrix`, consumed by `Lowering.build_matrices`)."""
    _name = None
    _parity_kraus = True   # every Kraus operator below is diagonal or anti-diagonal
    _batched = None

    def __init__(self, inputs: Any = None, nqubit: int = 1, wires=None, tsr_mode: bool = False,
                 requires_grad: bool = False) -> None:
        super().__init__(inputs=inputs, name=self._name, nqubit=nqubit, wires=wires, tsr_mode=tsr_mode,
                         requires_grad=requires_grad)

    @staticmethod
    def _kraus_of(prob: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def get_matrix(self, theta: Any) -> torch.Tensor:
        theta = self.inputs_to_tensor(theta).reshape(-1)
        return self._kraus_of(torch.sin(theta) ** 2)

    def _param_list(self):
        if self._batched is not None:      # [batch, npara] set by QubitCircuit for 2-D data
            return [self._batched[:, i] for i in range(self._batched.shape[1])]
        theta = self.theta.reshape(-1)
        return [theta[i] for i in range(self.npara)]

    @classmethod
    def _batched_matrix(cls, p: torch.Tensor) -> torch.Tensor:
        """`[..., N, npara]` angles -> `[..., N, 1, size]` lowered blocks (`Channel._lower_kraus`)."""
        low = cls._lower_kraus(cls._kraus_of(torch.sin(p) ** 2))
        return low.reshape(*low.shape[:-1], 1, low.shape[-1]) if low.ndim == p.ndim else \
            low.reshape(*p.shape[:-1], 1, -1)


class BitFlip(_OneParam):
    """rho -> (1-p) rho + p X rho X (reference channel.py:16-55)."""
    _name = 'BitFlip'
    _pauli_kraus = True

    @staticmethod
    def _kraus_of(prob):
        p = prob[..., 0]
        a, b = torch.sqrt(1 - p), torch.sqrt(p)
        return _kraus([[a, 0, 0, a], [0, b, b, 0]])


class PhaseFlip(_OneParam):
    """rho -> (1-p) rho + p Z rho Z (reference channel.py:58-97)."""
    _name = 'PhaseFlip'
    _diagonal_kraus = True

    @staticmethod
    def _kraus_of(prob):
        p = prob[..., 0]
        a, b = torch.sqrt(1 - p), torch.sqrt(p)
        return _kraus([[a, 0, 0, a], [b, 0, 0, -b]])


class Depolarizing(_OneParam):
    """rho -> (1-p) rho + p/3 (X rho X + Y rho Y + Z rho Z) (reference channel.py:100-149)."""
    _name = 'Depolarizing'
    _pauli_kraus = True

    @staticmethod
    def _kraus_of(prob):
        p = prob[..., 0]
        a, s = torch.sqrt(1 - p), torch.sqrt(p / 3)
        return _kraus([[a, 0, 0, a], [0, s, s, 0], [0, -1j * s, 1j * s, 0], [s, 0, 0, -s]])


class Pauli(_OneParam):
    """rho -> sum_k p_k P_k rho P_k with normalised `p = sin(theta)^2` (reference channel.py:152-212)."""
    _name = 'Pauli'
    _pauli_kraus = True

    def __init__(self, inputs: Any = None, nqubit: int = 1, wires=None, tsr_mode: bool = False,
                 requires_grad: bool = False) -> None:
        super().__init__(inputs=inputs, nqubit=nqubit, wires=wires, tsr_mode=tsr_mode, requires_grad=requires_grad)
        self.npara = 4

    @property
    def prob(self):
        prob = torch.sin(self.theta) ** 2
        return prob / prob.sum()

    def inputs_to_tensor(self, inputs: Any = None) -> torch.Tensor:
        if inputs is None:
            inputs = torch.rand(4) * torch.pi
        elif not isinstance(inputs, torch.Tensor):
            inputs = torch.tensor(inputs, dtype=torch.float).reshape(-1)[:4]
        return inputs

    @staticmethod
    def _kraus_of(prob):
        prob = prob / prob.sum(-1, keepdim=True)
        si, sx, sy, sz = (torch.sqrt(prob[..., k]) for k in range(4))
        return _kraus([[si, 0, 0, si], [0, sx, sx, 0], [0, -1j * sy, 1j * sy, 0], [sz, 0, 0, -sz]])

    def extra_repr(self) -> str:
        p = self.prob
        return f'wires={self.wires}, px={p[1].item()}, py={p[2].item()}, pz={p[3].item()}'


class AmplitudeDamping(_OneParam):
    """K0 = diag(1, sqrt(1-p)), K1 = sqrt(p) |0><1| (reference channel.py:215-263)."""
    _name = 'AmplitudeDamping'
    _damping_kraus = True

    @staticmethod
    def _kraus_of(prob):
        p = prob[..., 0]
        return _kraus([[torch.ones_like(p), 0, 0, torch.sqrt(1 - p)], [0, torch.sqrt(p), 0, 0]])


class PhaseDamping(_OneParam):
    """K0 = diag(1, sqrt(1-p)), K1 = sqrt(p) |1><1| (reference channel.py:266-314)."""
    _name = 'PhaseDamping'
    _diagonal_kraus = True

    @staticmethod
    def _kraus_of(prob):
        p = prob[..., 0]
        return _kraus([[torch.ones_like(p), 0, 0, torch.sqrt(1 - p)], [0, 0, 0, torch.sqrt(p)]])


class GeneralizedAmplitudeDamping(_OneParam):
    """Four Kraus operators with probability `p = sin(theta_0)^2` and rate `gamma = sin(theta_1)^2`
    (reference channel.py:317-383)."""
    _name = 'GeneralizedAmplitudeDamping'
    _damping_kraus = True

    def __init__(self, inputs: Any = None, nqubit: int = 1, wires=None, tsr_mode: bool = False,
                 requires_grad: bool = False) -> None:
        super().__init__(inputs=inputs, nqubit=nqubit, wires=wires, tsr_mode=tsr_mode, requires_grad=requires_grad)
        self.npara = 2

    def inputs_to_tensor(self, inputs: Any = None) -> torch.Tensor:
        if inputs is None:
            inputs = torch.rand(2) * torch.pi
        elif not isinstance(inputs, torch.Tensor):
            inputs = torch.tensor(inputs, dtype=torch.float).reshape(-1)[:2]
        return inputs

    @staticmethod
    def _kraus_of(prob):
        p, g = prob[..., 0], prob[..., 1]
        sp, sq, sg, s1g = torch.sqrt(p), torch.sqrt(1 - p), torch.sqrt(g), torch.sqrt(1 - g)
        return _kraus([[sp, 0, 0, sp * s1g], [0, sp * sg, 0, 0], [sq * s1g, 0, 0, sq], [0, 0, sq * sg, 0]])

    def extra_repr(self) -> str:
        return f'wires={self.wires}, probability={self.prob[0].item()}, rate={self.prob[1].item()}'


__all__ = ['BitFlip', 'PhaseFlip', 'Depolarizing', 'Pauli', 'AmplitudeDamping', 'PhaseDamping',
           'GeneralizedAmplitudeDamping']
