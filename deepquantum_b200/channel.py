"""Quantum channels with the reference's interface (channel.py:16-383): same class names, parameters
(`prob = sin(theta)^2`) and Kraus operators.  On the device a channel is ONE dense gate on the (row, column) wire
pair of the density matrix -- the superoperator `sum_i K_i (x) conj(K_i)` (`operation.Channel._lowered_matrix`) --
inside the same fused passes as the unitary gates; the reference evolves one copy of the state per Kraus
operator and sums them (operation.py:594-600)."""
from __future__ import annotations

from typing import Any

import torch

from .operation import Channel

_I = torch.tensor([[1, 0], [0, 1]], dtype=torch.cfloat)
_X = torch.tensor([[0, 1], [1, 0]], dtype=torch.cfloat)
_Y = torch.tensor([[0, -1j], [1j, 0]], dtype=torch.cfloat)
_Z = torch.tensor([[1, 0], [0, -1]], dtype=torch.cfloat)


def _paulis(device):
    return _I.to(device), _X.to(device), _Y.to(device), _Z.to(device)


def _mat(entries) -> torch.Tensor:
    return torch.stack(entries).reshape(2, 2)


class _OneParam(Channel):
    _name = None
    _parity_kraus = True   # every Kraus operator below is diagonal or anti-diagonal

    def __init__(self, inputs: Any = None, nqubit: int = 1, wires=None, tsr_mode: bool = False,
                 requires_grad: bool = False) -> None:
        super().__init__(inputs=inputs, name=self._name, nqubit=nqubit, wires=wires, tsr_mode=tsr_mode,
                         requires_grad=requires_grad)


class BitFlip(_OneParam):
    """rho -> (1-p) rho + p X rho X (reference channel.py:16-55)."""
    _name = 'BitFlip'

    def get_matrix(self, theta: Any) -> torch.Tensor:
        prob = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
        i, x, _, _ = _paulis(prob.device)
        return torch.stack([torch.sqrt(1 - prob) * i, torch.sqrt(prob) * x])


class PhaseFlip(_OneParam):
    """rho -> (1-p) rho + p Z rho Z (reference channel.py:58-97)."""
    _name = 'PhaseFlip'
    _diagonal_kraus = True

    def get_matrix(self, theta: Any) -> torch.Tensor:
        prob = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
        i, _, _, z = _paulis(prob.device)
        return torch.stack([torch.sqrt(1 - prob) * i, torch.sqrt(prob) * z])


class Depolarizing(_OneParam):
    """rho -> (1-p) rho + p/3 (X rho X + Y rho Y + Z rho Z) (reference channel.py:100-149)."""
    _name = 'Depolarizing'

    def get_matrix(self, theta: Any) -> torch.Tensor:
        prob = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
        i, x, y, z = _paulis(prob.device)
        s = torch.sqrt(prob / 3)
        return torch.stack([torch.sqrt(1 - prob) * i, s * x, s * y, s * z])


class Pauli(_OneParam):
    """rho -> sum_k p_k P_k rho P_k with normalised `p = sin(theta)^2` (reference channel.py:152-212)."""
    _name = 'Pauli'

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

    def get_matrix(self, theta: Any) -> torch.Tensor:
        prob = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
        prob = prob / prob.sum()
        return torch.stack([torch.sqrt(prob[k:k + 1]) * m for k, m in enumerate(_paulis(prob.device))])

    def extra_repr(self) -> str:
        p = self.prob
        return f'wires={self.wires}, px={p[1].item()}, py={p[2].item()}, pz={p[3].item()}'


class AmplitudeDamping(_OneParam):
    """K0 = diag(1, sqrt(1-p)), K1 = sqrt(p) |0><1| (reference channel.py:215-263)."""
    _name = 'AmplitudeDamping'

    def get_matrix(self, theta: Any) -> torch.Tensor:
        prob = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
        m0, m1 = torch.zeros_like(prob), torch.ones_like(prob)
        return torch.stack([_mat([m1, m0, m0, torch.sqrt(1 - prob)]), _mat([m0, torch.sqrt(prob), m0, m0])]) + 0j


class PhaseDamping(_OneParam):
    """K0 = diag(1, sqrt(1-p)), K1 = sqrt(p) |1><1| (reference channel.py:266-314)."""
    _name = 'PhaseDamping'
    _diagonal_kraus = True

    def get_matrix(self, theta: Any) -> torch.Tensor:
        prob = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
        m0, m1 = torch.zeros_like(prob), torch.ones_like(prob)
        return torch.stack([_mat([m1, m0, m0, torch.sqrt(1 - prob)]), _mat([m0, m0, m0, torch.sqrt(prob)])]) + 0j


class GeneralizedAmplitudeDamping(_OneParam):
    """Four Kraus operators with probability `p = sin(theta_0)^2` and rate `gamma = sin(theta_1)^2`
    (reference channel.py:317-383)."""
    _name = 'GeneralizedAmplitudeDamping'

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

    def get_matrix(self, theta: Any) -> torch.Tensor:
        prob = torch.sin(self.inputs_to_tensor(theta).reshape(-1)) ** 2
        p, g = prob[0], prob[1]
        m0, m1 = torch.zeros_like(p), torch.ones_like(p)
        k0 = torch.sqrt(p) * _mat([m1, m0, m0, torch.sqrt(1 - g)])
        k1 = torch.sqrt(p) * _mat([m0, torch.sqrt(g), m0, m0])
        k2 = torch.sqrt(1 - p) * _mat([torch.sqrt(1 - g), m0, m0, m1])
        k3 = torch.sqrt(1 - p) * _mat([m0, m0, torch.sqrt(g), m0])
        return torch.stack([k0, k1, k2, k3]) + 0j

    def extra_repr(self) -> str:
        return f'wires={self.wires}, probability={self.prob[0].item()}, rate={self.prob[1].item()}'


__all__ = ['BitFlip', 'PhaseFlip', 'Depolarizing', 'Pauli', 'AmplitudeDamping', 'PhaseDamping',
           'GeneralizedAmplitudeDamping']
