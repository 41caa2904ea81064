"""The gate zoo with the reference's class names, constructor signatures and matrices (gate.py).

Every class keeps the reference conventions -- constant matrices are complex64 buffers that `.to()`
widens (so the complex128 path carries float32-rounded constants exactly like the reference,
gate.py:841-1367), parameters are float32 unless a tensor is passed (gate.py:384-391) -- and adds
the *structure class* the fusion planner may rely on (dense / diagonal / Pauli-X permutation).
Permutation gates (CNOT, Toffoli, Swap, Fredkin) are lowered to controlled-X records: the result is
bit-identical to multiplying by the reference's dense 0/1 matrix.
"""
from __future__ import annotations

from copy import copy
from typing import Any

import torch
from torch import nn

from . import _lib as L
from .operation import Gate, Lowering


# -------------------------------------------------------------------------------------------------
# bases
# -------------------------------------------------------------------------------------------------
class SingleGate(Gate):
    def __init__(self, name=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                         den_mat=den_mat, tsr_mode=tsr_mode)
        assert len(self.wires) == 1


class DoubleGate(Gate):
    def __init__(self, name=None, nqubit=2, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False) -> None:
        if wires is None:
            wires = [0, 1]
        assert len(wires) == 2
        super().__init__(name=name, nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                         den_mat=den_mat, tsr_mode=tsr_mode)


class DoubleControlGate(DoubleGate):
    def __init__(self, name=None, nqubit=2, wires=None, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=wires, controls=None, condition=False, den_mat=den_mat,
                         tsr_mode=tsr_mode)


class TripleGate(Gate):
    def __init__(self, name=None, nqubit=3, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False) -> None:
        if wires is None:
            wires = [0, 1, 2]
        assert len(wires) == 3
        super().__init__(name=name, nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                         den_mat=den_mat, tsr_mode=tsr_mode)


class ArbitraryGate(Gate):
    def __init__(self, name=None, nqubit=1, wires=None, minmax=None, controls=None, den_mat=False,
                 tsr_mode=False) -> None:
        self.nqubit = nqubit
        if wires is None:
            if minmax is None:
                minmax = [0, nqubit - 1]
            self._check_minmax(minmax)
            wires = list(range(minmax[0], minmax[1] + 1))
        super().__init__(name=name, nqubit=nqubit, wires=wires, controls=controls, condition=False, den_mat=den_mat,
                         tsr_mode=tsr_mode)
        self.minmax = [min(self.wires), max(self.wires)]
        self.inv_mode = False

    def inverse(self) -> 'ArbitraryGate':
        gate = copy(self)
        gate.inv_mode = not self.inv_mode
        gate.name = self.name + '_dagger' if isinstance(self.name, str) else self.name
        return gate


class _Parametric:
    """Shared behaviour of parametric gates (reference gate.py:341-520): parameters are scalar tensors
    (`nn.Parameter` or buffer), `inverse()` is a shallow copy with `inv_mode` flipped."""

    _matrix_source = 'group'
    _pnames = ('theta',)
    _fast_encode = True      # QubitCircuit.encode may set the buffer directly (classes overriding init_para opt out)

    def _setup_parametric(self, inputs, requires_grad):
        self.npara = len(self._pnames)
        self.requires_grad = requires_grad
        self.inv_mode = False
        self._batched = None
        self.init_para(inputs)

    def inputs_to_tensor(self, inputs: Any = None) -> torch.Tensor:
        while isinstance(inputs, list):
            inputs = inputs[0]
        if inputs is None:
            inputs = torch.rand(1)[0] * 4 * torch.pi
        elif not isinstance(inputs, (torch.Tensor, nn.Parameter)):
            inputs = torch.tensor(inputs, dtype=torch.float)
        return inputs

    def _set(self, name, value):
        if self.requires_grad:
            setattr(self, name, nn.Parameter(value))
        else:
            if name in self._parameters:
                del self._parameters[name]
            self.register_buffer(name, value)

    def init_para(self, inputs: Any = None) -> None:
        self._batched = None
        self._set('theta', self.inputs_to_tensor(inputs))
        self._matrix_cache = None     # `matrix` is refreshed lazily: encode() re-parametrises hundreds of gates

    @property
    def matrix(self) -> torch.Tensor:
        """Detached local matrix of the current parameters (the reference refreshes it eagerly in init_para,
        gate.py:411-415; here it is computed on first access so that `encode` costs no kernel launches)."""
        if self.__dict__.get('_matrix_cache') is None:
            self.update_matrix()
        return self.__dict__['_matrix_cache']

    @matrix.setter
    def matrix(self, value) -> None:
        self.__dict__['_matrix_cache'] = value

    def _signed(self):
        """Parameter tensors with `inv_mode` applied (gate.py:395-400)."""
        return [-self.theta if self.inv_mode else self.theta]

    def _param_list(self):
        if self._batched is not None:  # [batch, npara] set by QubitCircuit for 2-D data
            cols = [self._batched[:, i] for i in range(self._batched.shape[1])]
            return self._signed_cols(cols)
        return [t.reshape(()) for t in self._signed()]

    def _signed_cols(self, cols):
        return [-c for c in cols] if self.inv_mode else cols

    def update_matrix(self) -> torch.Tensor:
        p = torch.stack([t.reshape(()) for t in self._signed()])
        matrix = self._batched_matrix(p.unsqueeze(0))[0]
        self.matrix = matrix.detach()
        return matrix

    def get_matrix(self, *inputs) -> torch.Tensor:
        p = torch.stack([self.inputs_to_tensor(x).reshape(()) for x in inputs])
        return self._batched_matrix(p.unsqueeze(0))[0]

    def get_derivative(self, *inputs) -> torch.Tensor:
        """d(matrix)/d(parameters) via the jacobian of the small matrix (reference gate.py:402-406)."""
        from torch.autograd.functional import jacobian
        p = torch.stack([self.inputs_to_tensor(x).reshape(()) for x in inputs]).detach()
        jac = jacobian(lambda q: torch.view_as_real(self._batched_matrix(q.unsqueeze(0))[0]), p)
        out = jac[..., 0, :] + 1j * jac[..., 1, :]   # [d, d, npara]
        return out.squeeze(-1) if out.shape[-1] == 1 else out.permute(2, 0, 1)

    def inverse(self):
        gate = copy(self)
        gate.inv_mode = not self.inv_mode
        return gate

    def extra_repr(self) -> str:
        vals = ', '.join(f'{n}={t.item():.6g}' for n, t in zip(self._pnames, self._signed()))
        s = f'wires={self.wires}, {vals}'
        return s if self.controls == [] else s + f', controls={self.controls}'


class ParametricSingleGate(_Parametric, SingleGate):
    def __init__(self, name=None, inputs=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        SingleGate.__init__(self, name=name, nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                            den_mat=den_mat, tsr_mode=tsr_mode)
        self._setup_parametric(inputs, requires_grad)


class ParametricDoubleGate(_Parametric, DoubleGate):
    def __init__(self, name=None, inputs=None, nqubit=2, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        DoubleGate.__init__(self, name=name, nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                            den_mat=den_mat, tsr_mode=tsr_mode)
        self._setup_parametric(inputs, requires_grad)


def _cplx(re, im=None):
    return torch.complex(re, torch.zeros_like(re) if im is None else im)


def _mat(entries, d):
    """entries: list of d*d tensors [..., N] -> [..., N, d, d]."""
    return torch.stack(entries, dim=-1).reshape(*entries[0].shape, d, d)


# -------------------------------------------------------------------------------------------------
# single-qubit gates
# -------------------------------------------------------------------------------------------------
class U3Gate(ParametricSingleGate):
    """U3(theta, phi, lambda) (reference gate.py:523-674)."""

    _pnames = ('theta', 'phi', 'lambd')

    def __init__(self, inputs=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='U3Gate', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    def inputs_to_tensor(self, inputs: Any = None):
        if inputs is None:
            theta = torch.rand(1)[0] * torch.pi
            phi = torch.rand(1)[0] * 2 * torch.pi
            lambd = torch.rand(1)[0] * 2 * torch.pi
            return theta, phi, lambd
        if isinstance(inputs, (torch.Tensor, nn.Parameter)) and inputs.ndim == 0:
            return inputs
        if not isinstance(inputs, (torch.Tensor, nn.Parameter)) and not isinstance(inputs, (list, tuple)):
            return torch.tensor(inputs, dtype=torch.float)
        out = []
        for x in inputs:
            out.append(x if isinstance(x, (torch.Tensor, nn.Parameter)) else torch.tensor(x, dtype=torch.float))
        return tuple(out)

    _fast_encode = False

    def init_para(self, inputs: Any = None) -> None:
        self._batched = None
        theta, phi, lambd = self.inputs_to_tensor(inputs)
        self._set('theta', theta)
        self._set('phi', phi)
        self._set('lambd', lambd)
        self._matrix_cache = None

    def _signed(self):
        if self.inv_mode:  # gate.py:606-609
            return [-self.theta, -self.lambd, -self.phi]
        return [self.theta, self.phi, self.lambd]

    def _signed_cols(self, cols):
        return [-cols[0], -cols[2], -cols[1]] if self.inv_mode else cols

    def get_matrix(self, theta, phi, lambd) -> torch.Tensor:
        p = torch.stack([torch.as_tensor(x, dtype=torch.float).reshape(()) if not isinstance(x, torch.Tensor)
                         else x.reshape(()) for x in (theta, phi, lambd)])
        return self._batched_matrix(p.unsqueeze(0))[0]

    @staticmethod
    def _batched_matrix(p):
        theta, phi, lambd = p[..., 0], p[..., 1], p[..., 2]
        c, s = _cplx(torch.cos(theta / 2)), _cplx(torch.sin(theta / 2))
        e_il = torch.exp(1j * lambd)
        e_ip = torch.exp(1j * phi)
        e_ipl = torch.exp(1j * (phi + lambd))
        return _mat([c, -e_il * s, e_ip * s, e_ipl * c], 2)


class PhaseShift(ParametricSingleGate):
    """diag(1, e^{i theta}) (reference gate.py:677-753)."""

    _kind = L.GATE_DIAG

    def __init__(self, inputs=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='PhaseShift', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        one = _cplx(torch.ones_like(theta))
        zero = torch.zeros_like(one)
        return _mat([one, zero, zero, torch.exp(1j * theta)], 2)


class _ConstSingle(SingleGate):
    _shared_const = True
    _matrix_entries = None
    _default_name = None

    def __init__(self, nqubit=1, wires=None, controls=None, condition=False, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name=self._default_name, nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                         den_mat=den_mat, tsr_mode=tsr_mode)
        self.register_buffer('matrix', self._make_matrix())

    @classmethod
    def _make_matrix(cls):
        return torch.tensor(cls._matrix_entries, dtype=torch.cfloat)


class Identity(Gate):
    def __init__(self, nqubit=1, wires=None, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name='Identity', nqubit=nqubit, wires=wires, controls=None, den_mat=den_mat,
                         tsr_mode=tsr_mode)
        self.register_buffer('matrix', torch.eye(2**self.nqubit, dtype=torch.cfloat))

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        return

    def get_unitary(self) -> torch.Tensor:
        return self.matrix

    def forward(self, x: Any) -> Any:
        return x


class PauliX(_ConstSingle):
    _kind = L.GATE_X
    _default_name = 'PauliX'
    _matrix_entries = [[0, 1], [1, 0]]


class PauliY(_ConstSingle):
    _hint = L.GATE_RXLIKE
    _default_name = 'PauliY'
    _matrix_entries = [[0, -1j], [1j, 0]]


class PauliZ(_ConstSingle):
    _kind = L.GATE_DIAG
    _hint = L.GATE_PHASE_Z
    _default_name = 'PauliZ'
    _matrix_entries = [[1, 0], [0, -1]]


class Hadamard(_ConstSingle):
    _hint = L.GATE_REAL | L.GATE_HADAMARD
    _default_name = 'Hadamard'

    @classmethod
    def _make_matrix(cls):  # gate.py:1069: complex64 tensor divided by a Python float
        return torch.tensor([[1, 1], [1, -1]], dtype=torch.cfloat) / 2**0.5


class SGate(_ConstSingle):
    _kind = L.GATE_DIAG
    _hint = L.GATE_PHASE_S
    _default_name = 'SGate'
    _matrix_entries = [[1, 0], [0, 1j]]

    def inverse(self):
        return SDaggerGate(nqubit=self.nqubit, wires=self.wires, controls=self.controls, tsr_mode=self.tsr_mode).to(
            self.matrix.device, self.matrix.real.dtype)


class SDaggerGate(_ConstSingle):
    _kind = L.GATE_DIAG
    _hint = L.GATE_PHASE_SDG
    _default_name = 'SDaggerGate'
    _matrix_entries = [[1, 0], [0, -1j]]

    def inverse(self):
        return SGate(nqubit=self.nqubit, wires=self.wires, controls=self.controls, tsr_mode=self.tsr_mode).to(
            self.matrix.device, self.matrix.real.dtype)


class TGate(_ConstSingle):
    _kind = L.GATE_DIAG
    _default_name = 'TGate'
    _matrix_entries = [[1, 0], [0, (1 + 1j) / 2**0.5]]

    def inverse(self):
        return TDaggerGate(nqubit=self.nqubit, wires=self.wires, controls=self.controls, tsr_mode=self.tsr_mode).to(
            self.matrix.device, self.matrix.real.dtype)


class TDaggerGate(_ConstSingle):
    _kind = L.GATE_DIAG
    _default_name = 'TDaggerGate'
    _matrix_entries = [[1, 0], [0, (1 - 1j) / 2**0.5]]

    def inverse(self):
        return TGate(nqubit=self.nqubit, wires=self.wires, controls=self.controls, tsr_mode=self.tsr_mode).to(
            self.matrix.device, self.matrix.real.dtype)


class Rx(ParametricSingleGate):
    """exp(-i theta X / 2) (reference gate.py:1389-1480)."""

    _hint = L.GATE_RXLIKE | L.GATE_ROTATION

    def __init__(self, inputs=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='Rx', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        c = _cplx(torch.cos(theta / 2))
        misin = _cplx(torch.zeros_like(theta), -torch.sin(theta / 2))
        return _mat([c, misin, misin, c], 2)


class Ry(ParametricSingleGate):
    _hint = L.GATE_REAL | L.GATE_ROTATION

    def __init__(self, inputs=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='Ry', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        c, s = torch.cos(theta / 2), torch.sin(theta / 2)
        return _mat([_cplx(c), _cplx(-s), _cplx(s), _cplx(c)], 2)


class Rz(ParametricSingleGate):
    _kind = L.GATE_DIAG

    def __init__(self, inputs=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='Rz', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        em, ep = torch.exp(-1j * theta / 2), torch.exp(1j * theta / 2)
        zero = torch.zeros_like(em)
        return _mat([em, zero, zero, ep], 2)


class ProjectionJ(ParametricSingleGate):
    """Measurement-plane rotation J (reference gate.py:1674-1787)."""

    def __init__(self, inputs=None, nqubit=1, wires=None, plane='xy', controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        self.plane = plane.lower()
        assert self.plane in ('xy', 'yx', 'yz', 'zy', 'zx', 'xz'), f'Unsupported measurement plane: {plane}'
        super().__init__(name='ProjectionJ', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    _matrix_source = 'dyn'   # the plane is per instance, so no per-class batching

    def update_matrix(self) -> torch.Tensor:
        theta = -self.theta if self.inv_mode else self.theta  # the reference's convention (gate.py:395-400)
        matrix = self._plane_matrix(theta.reshape(1))[0]
        self.matrix = matrix.detach()
        return matrix

    def get_matrix(self, theta) -> torch.Tensor:
        return self._plane_matrix(self.inputs_to_tensor(theta).reshape(1))[0]

    def _plane_matrix(self, theta):
        if self.plane in ('xy', 'yx'):
            one = _cplx(torch.ones_like(theta))
            e = torch.exp(-1j * theta)
            return _mat([one, e, one, -e], 2) / 2**0.5
        if self.plane in ('yz', 'zy'):
            cps = _cplx(torch.cos(theta / 2) + torch.sin(theta / 2))
            cms = _cplx(torch.cos(theta / 2) - torch.sin(theta / 2))
            return _mat([cps, -1j * cms, cms, 1j * cps], 2) / 2**0.5
        c, s = _cplx(torch.cos(theta / 2)), _cplx(torch.sin(theta / 2))
        return _mat([c, s, s, -c], 2)


# -------------------------------------------------------------------------------------------------
# two- and three-qubit gates
# -------------------------------------------------------------------------------------------------
class CNOT(DoubleControlGate):
    """Dense 4x4 CNOT on [control, target] (reference gate.py:1906-1960); a permutation, lowered to a
    controlled amplitude swap."""

    def __init__(self, nqubit=2, wires=None, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name='CNOT', nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
        self.register_buffer('matrix', torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]) + 0j)

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        low.add(self, L.GATE_X, [self.wires[1]], [self.wires[0]])


class Swap(DoubleGate):
    def __init__(self, nqubit=2, wires=None, controls=None, condition=False, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name='Swap', nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                         den_mat=den_mat, tsr_mode=tsr_mode)
        self.register_buffer('matrix', torch.tensor([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) + 0j)

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        a, b = self.wires
        for c, t in ((a, b), (b, a), (a, b)):
            low.add(self, L.GATE_X, [t], [c] + self.controls)


class ImaginarySwap(DoubleGate):
    def __init__(self, nqubit=2, wires=None, controls=None, condition=False, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name='ImaginarySwap', nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                         den_mat=den_mat, tsr_mode=tsr_mode)
        self.register_buffer('matrix', torch.tensor([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]))


def _diag4(a, b, c, d):
    z = torch.zeros_like(a)
    return _mat([a, z, z, z, z, b, z, z, z, z, c, z, z, z, z, d], 4)


class Rxx(ParametricDoubleGate):
    def __init__(self, inputs=None, nqubit=2, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='Rxx', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        c = _cplx(torch.cos(theta / 2))
        s = _cplx(torch.zeros_like(theta), -torch.sin(theta / 2))
        z = torch.zeros_like(c)
        return _mat([c, z, z, s, z, c, s, z, z, s, c, z, s, z, z, c], 4)


class Ryy(ParametricDoubleGate):
    def __init__(self, inputs=None, nqubit=2, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='Ryy', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        c = _cplx(torch.cos(theta / 2))
        s = _cplx(torch.zeros_like(theta), torch.sin(theta / 2))
        z = torch.zeros_like(c)
        return _mat([c, z, z, s, z, c, -s, z, z, -s, c, z, s, z, z, c], 4)


class Rzz(ParametricDoubleGate):
    _kind = L.GATE_DIAG

    def __init__(self, inputs=None, nqubit=2, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='Rzz', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        em, ep = torch.exp(-1j * theta / 2), torch.exp(1j * theta / 2)
        return _diag4(em, ep, ep, em)


class Rxy(ParametricDoubleGate):
    def __init__(self, inputs=None, nqubit=2, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='Rxy', inputs=inputs, nqubit=nqubit, wires=wires, controls=controls,
                         condition=condition, den_mat=den_mat, tsr_mode=tsr_mode, requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        c = _cplx(torch.cos(theta / 2))
        s = _cplx(torch.zeros_like(theta), -torch.sin(theta / 2))
        z, o = torch.zeros_like(c), torch.ones_like(c)
        return _mat([o, z, z, z, z, c, s, z, z, s, c, z, z, z, z, o], 4)


class ReconfigurableBeamSplitter(ParametricDoubleGate):
    def __init__(self, inputs=None, nqubit=2, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name='ReconfigurableBeamSplitter', inputs=inputs, nqubit=nqubit, wires=wires,
                         controls=controls, condition=condition, den_mat=den_mat, tsr_mode=tsr_mode,
                         requires_grad=requires_grad)

    @staticmethod
    def _batched_matrix(p):
        theta = p[..., 0]
        c, s = _cplx(torch.cos(theta)), _cplx(torch.sin(theta))
        z, o = torch.zeros_like(c), torch.ones_like(c)
        return _mat([o, z, z, z, z, c, s, z, z, -s, c, z, z, z, z, o], 4)


def _perm8(swap):
    m = torch.eye(8)
    m[[swap[0], swap[1]]] = m[[swap[1], swap[0]]]
    return m + 0j


class Toffoli(TripleGate):
    """Dense 8x8 CCX on [control1, control2, target] (reference gate.py:2482-2649)."""

    def __init__(self, nqubit=3, wires=None, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name='Toffoli', nqubit=nqubit, wires=wires, controls=None, den_mat=den_mat,
                         tsr_mode=tsr_mode)
        self.register_buffer('matrix', _perm8((6, 7)))

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        low.add(self, L.GATE_X, [self.wires[2]], [self.wires[0], self.wires[1]])


class Fredkin(TripleGate):
    """Dense 8x8 controlled-SWAP on [control, target1, target2] (reference gate.py:2652-2742)."""

    def __init__(self, nqubit=3, wires=None, den_mat=False, tsr_mode=False) -> None:
        super().__init__(name='Fredkin', nqubit=nqubit, wires=wires, controls=None, den_mat=den_mat,
                         tsr_mode=tsr_mode)
        self.register_buffer('matrix', _perm8((5, 6)))

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        c0, a, b = self.wires
        for c, t in ((a, b), (b, a), (a, b)):
            low.add(self, L.GATE_X, [t], [c, c0])


class UAnyGate(ArbitraryGate):
    """Arbitrary unitary on `wires` (reference gate.py:2745-2788)."""

    def __init__(self, unitary, nqubit=1, wires=None, minmax=None, controls=None, name='UAnyGate', den_mat=False,
                 tsr_mode=False) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=wires, minmax=minmax, controls=controls, den_mat=den_mat,
                         tsr_mode=tsr_mode)
        if not isinstance(unitary, torch.Tensor):
            unitary = torch.tensor(unitary, dtype=torch.cfloat).reshape(-1, 2 ** len(self.wires))
        assert unitary.dtype in (torch.cfloat, torch.cdouble)
        assert unitary.shape[-1] == unitary.shape[-2] == 2 ** len(self.wires)
        err = (unitary @ unitary.mH - torch.eye(unitary.shape[-1], dtype=unitary.dtype, device=unitary.device)).abs()
        assert float(err.max()) < 1e-4, 'Please check the unitary matrix'
        self.register_buffer('matrix', unitary)

    def update_matrix(self) -> torch.Tensor:
        return self.matrix.mH if self.inv_mode else self.matrix

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        low.add(self, self._kind, self.wires, self.controls, adjoint=inverse != self.inv_mode)


class LatentGate(ArbitraryGate):
    """Unitary obtained from the SVD of a trainable latent matrix (reference gate.py:2791-2864)."""

    _matrix_source = 'dyn'

    def __init__(self, inputs=None, nqubit=1, wires=None, minmax=None, controls=None, name='LatentGate',
                 den_mat=False, tsr_mode=False, requires_grad=False) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=wires, minmax=minmax, controls=controls, den_mat=den_mat,
                         tsr_mode=tsr_mode)
        self.requires_grad = requires_grad
        self.init_para(inputs)

    def inputs_to_tensor(self, inputs=None) -> torch.Tensor:
        dim = 2 ** len(self.wires)
        if inputs is None:
            inputs = torch.randn(dim, dim)
        elif not isinstance(inputs, (torch.Tensor, nn.Parameter)):
            inputs = torch.tensor(inputs, dtype=torch.float)
        assert inputs.shape[-1] == inputs.shape[-2] == dim
        return inputs

    def get_matrix(self, inputs) -> torch.Tensor:
        latent = self.inputs_to_tensor(inputs) + 0j
        u, _, vh = torch.linalg.svd(latent)
        return u @ vh

    def update_matrix(self) -> torch.Tensor:
        latent = self.latent.mH if self.inv_mode else self.latent
        matrix = self.get_matrix(latent)
        self.matrix = matrix.detach()
        return matrix

    def init_para(self, inputs=None) -> None:
        latent = self.inputs_to_tensor(inputs)
        if self.requires_grad:
            self.latent = nn.Parameter(latent)
        else:
            self.register_buffer('latent', latent)
        self.update_matrix()
        self.npara = self.latent.numel()


class CombinedSingleGate(SingleGate):
    """Product of single-qubit gates on the same wire, applied as ONE 2x2 (reference gate.py:1790-1903).  The
    member gates keep their parameters (autograd chains through the matrix product)."""

    _matrix_source = 'dyn'

    def __init__(self, gates, name=None, nqubit=1, wires=None, controls=None, condition=False, den_mat=False,
                 tsr_mode=False) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=wires, controls=controls, condition=condition,
                         den_mat=den_mat, tsr_mode=tsr_mode)
        self.gates = nn.ModuleList()
        for gate in gates:
            self._adopt(gate)
            self.gates.append(gate)
        self.update_npara()
        self.update_matrix()

    def _adopt(self, gate) -> None:
        gate.nqubit, gate.wires, gate.controls = self.nqubit, self.wires, self.controls
        gate.condition, gate.den_mat, gate.tsr_mode = self.condition, self.den_mat, self.tsr_mode

    def get_matrix(self) -> torch.Tensor:
        matrix = None
        for gate in self.gates:
            matrix = gate.update_matrix() if matrix is None else gate.update_matrix() @ matrix
        return matrix

    def update_matrix(self) -> torch.Tensor:
        matrix = self.get_matrix()
        self.matrix = matrix.detach()
        return matrix

    def update_npara(self) -> None:
        self.npara = sum(g.npara for g in self.gates)

    def init_para(self, inputs=None) -> None:
        count = 0
        for gate in self.gates:
            if gate.npara:
                gate.init_para(None if inputs is None else inputs[count:count + gate.npara])
            count += gate.npara
        self.update_matrix()

    def add(self, gate) -> None:
        self._adopt(gate)
        self.gates.append(gate)
        self.update_npara()
        self.update_matrix()

    def inverse(self) -> 'CombinedSingleGate':
        return CombinedSingleGate(gates=[g.inverse() for g in reversed(self.gates)], name=self.name,
                                  nqubit=self.nqubit, wires=self.wires, controls=self.controls,
                                  condition=self.condition, den_mat=self.den_mat, tsr_mode=self.tsr_mode)

    def _apply(self, fn: Any) -> 'CombinedSingleGate':
        nn.Module._apply(self, fn)
        return self


class HamiltonianGate(ArbitraryGate):
    """`exp(-i H t)` on `wires` (reference gate.py:2867-3024): `hamiltonian` is a Pauli-sum list such as
    `[[0.5, 'x0y1'], [-1, 'z3y1']]` (then the gate spans the min..max wire it names) or a Hermitian matrix on
    `wires` / `minmax`.  The matrix exponential is a torch call on the 2^k x 2^k block, differentiable in `t`;
    the block is applied by the dense k-target kernel path (k <= 4)."""

    _matrix_source = 'dyn'

    def __init__(self, hamiltonian, t=None, nqubit=1, wires=None, minmax=None, controls=None,
                 name='HamiltonianGate', den_mat=False, tsr_mode=False, requires_grad=False) -> None:
        self.nqubit = nqubit
        self.ham_lst = None
        if isinstance(hamiltonian, list):
            self.ham_lst = hamiltonian
            wires = None
            minmax = self.get_minmax(hamiltonian)
        super().__init__(name=name, nqubit=nqubit, wires=wires, minmax=minmax, controls=controls, den_mat=den_mat,
                         tsr_mode=tsr_mode)
        self.npara = 1
        self.requires_grad = requires_grad
        self.register_buffer('x', torch.tensor([[0, 1], [1, 0]], dtype=torch.cfloat))
        self.register_buffer('y', torch.tensor([[0, -1j], [1j, 0]], dtype=torch.cfloat))
        self.register_buffer('z', torch.tensor([[1, 0], [0, -1]], dtype=torch.cfloat))
        self.init_para([hamiltonian, t])

    def _apply(self, fn: Any) -> 'HamiltonianGate':
        from .operation import apply_complex_fix
        names = [k for k in ('x', 'y', 'z', 'ham_tsr') if k in self._buffers]
        tensors = {k: self._buffers.pop(k) for k in names}
        nn.Module._apply(self, fn)
        for key, value in apply_complex_fix(fn, tensors).items():
            self.register_buffer(key, value)
        return self

    @staticmethod
    def _convert_hamiltonian(hamiltonian: list) -> list:
        if len(hamiltonian) == 2 and isinstance(hamiltonian[1], str):
            hamiltonian = [hamiltonian]
        assert all(isinstance(i, list) for i in hamiltonian), 'Invalid input type'
        for pair in hamiltonian:
            assert isinstance(pair[1], str), 'Invalid input type'
        return hamiltonian

    def get_minmax(self, hamiltonian: list) -> list[int]:
        lo, hi = self.nqubit - 1, 0
        for pair in self._convert_hamiltonian(hamiltonian):
            for i in pair[1][1::2]:
                lo, hi = min(lo, int(i)), max(hi, int(i))
        return [lo, hi]

    def inputs_to_tensor(self, inputs=None):
        if inputs is None:
            return self.ham_tsr, torch.rand(1)[0]
        ham, t = inputs
        if ham is None:
            ham_tsr = self.ham_tsr
        elif isinstance(ham, list):
            ham = self._convert_hamiltonian(ham)
            paulis = {'x': self.x, 'y': self.y, 'z': self.z}
            identity = torch.eye(2, dtype=self.x.dtype, device=self.x.device)
            lo, hi = self.get_minmax(ham)
            ham_tsr = None
            for coeff, string in ham:
                lst = [identity] * self.nqubit
                for wire, key in zip(string[1::2], string[::2]):
                    lst[int(wire)] = paulis[key.lower()]
                term = lst[lo]
                for m in lst[lo + 1:hi + 1]:
                    term = torch.kron(term, m)
                ham_tsr = term * coeff if ham_tsr is None else ham_tsr + term * coeff
        elif not isinstance(ham, torch.Tensor):
            ham_tsr = torch.tensor(ham, dtype=self.x.dtype, device=self.x.device)
        else:
            ham_tsr = ham
        assert torch.allclose(ham_tsr, ham_tsr.mH)
        if t is None:
            t = torch.rand(1)[0]
        elif not isinstance(t, (torch.Tensor, nn.Parameter)):
            t = torch.tensor(t, dtype=torch.float)
        return ham_tsr, t

    def get_matrix(self, hamiltonian, t) -> torch.Tensor:
        ham, t = self.inputs_to_tensor([hamiltonian, t])
        return torch.linalg.matrix_exp(-1j * ham * t)

    def update_matrix(self) -> torch.Tensor:
        t = -self.t if self.inv_mode else self.t
        matrix = self.get_matrix(self.ham_tsr, t)
        assert matrix.shape[-1] == matrix.shape[-2] == 2 ** len(self.wires)
        self.matrix = matrix.detach()
        return matrix

    def get_derivative(self, t) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            t = torch.tensor(t, dtype=torch.float)
        du = torch.autograd.functional.jacobian(lambda v: torch.view_as_real(self.get_matrix(self.ham_tsr, v)),
                                                t.squeeze())
        return du[..., 0] + du[..., 1] * 1j

    def init_para(self, inputs=None) -> None:
        ham, t = self.inputs_to_tensor(inputs)
        self.register_buffer('ham_tsr', ham)
        if self.requires_grad:
            self.t = nn.Parameter(t)
        else:
            if 't' in self._parameters:
                del self._parameters['t']
            self.register_buffer('t', t)
        self.update_matrix()

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        # `inv_mode` is folded into the matrix (t -> -t); a further inversion by the circuit is an adjoint record
        low.add(self, self._kind, self.wires, self.controls, adjoint=inverse)


class Barrier(Gate):
    """No-op (reference gate.py:3097-3126); never a fusion barrier."""

    def __init__(self, nqubit=1, wires=None, name='Barrier') -> None:
        if wires is None:
            wires = list(range(nqubit))
        super().__init__(name=name, nqubit=nqubit, wires=wires)

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        return

    def forward(self, x: Any) -> Any:
        return x


class Reset(Gate):
    """Reset of `wires` to |0> (reference gate.py:3027-3094), a NON-unitary, state-dependent operation.

    `postselect` 0 / 1: every wire in turn is projected on that outcome, renormalised by the outcome's probability and
    relabelled |0> (a wire whose outcome has probability exactly 0 keeps the other branch, as in the reference);
    `postselect=None`: the outcome of all wires is sampled per state.  All wires of the circuit: the state becomes
    |0...0>.  A circuit runs the gates before and after a Reset as separate fused programs (circuit.py); the reduction
    and the 2 x 2 (2^k x 2^k) projection matrix are evaluated on the device, the projection itself is one more pass of
    the same kernels.  Forward only."""

    def __init__(self, nqubit: int = 1, wires=None, postselect: int | None = 0, tsr_mode: bool = False) -> None:
        if wires is None:
            wires = list(range(nqubit))
        super().__init__(name='Reset', nqubit=nqubit, wires=wires, tsr_mode=tsr_mode)
        assert postselect in (0, 1, None)
        self.postselect = postselect

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        raise NotImplementedError('Reset splits the circuit into separate programs (QubitCircuit handles it); it cannot '
                                  'be part of a fused program, an inverse circuit or a unitary')

    def inverse(self):
        raise NotImplementedError('Reset has no inverse')

    def apply_(self, x: torch.Tensor, batch: int) -> None:
        """In place on a contiguous device tensor `[batch, 2^nqubit]`."""
        from . import _lib as L
        from . import engine
        n = self.nqubit
        if len(self.wires) == n:
            engine.init_basis_(x, n, batch, 0)
            return
        rdt = x.real.dtype
        if self.postselect is None:
            wires = sorted(self.wires)
            k = len(wires)
            assert k <= L.MAX_TARGETS, 'sampled reset of more than 6 wires at once'
            prob = (x.real**2 + x.imag**2).reshape([batch] + [2] * n)
            other = [1 + w for w in range(n) if w not in wires]
            prob = (prob.sum(other) if other else prob).reshape(batch, 2**k)
            sample = torch.multinomial(prob.double(), 1)                                    # [batch, 1]
            mats = torch.zeros(batch, 2**k, 2**k, dtype=rdt, device=x.device)
            mats[:, 0, :].scatter_(1, sample, prob.gather(1, sample).to(rdt).rsqrt())
            self._apply_matrix(x, batch, wires, mats)
            return
        ps = self.postselect
        for w in self.wires:
            bit = n - 1 - w
            masks = torch.tensor([1 << bit, 0], dtype=torch.int64, device=x.device)
            red = engine.expectation_z(x, n, masks, batch)                                  # [batch, 2]: p0 - p1, p0 + p1
            p = (red[:, 1] + (1 - 2 * ps) * red[:, 0]) / 2                                  # probability of `postselect`
            empty = (p <= 0).to(p.dtype)                                                    # 1 - sign(p) of the reference
            norm = torch.sqrt(p.clamp_min(0) + empty)
            keep, other = ((1 - empty) / norm).to(rdt), (empty / norm).to(rdt)
            mats = torch.zeros(batch, 2, 2, dtype=rdt, device=x.device)
            mats[:, 0, ps] = keep
            mats[:, 0, 1 - ps] = other
            self._apply_matrix(x, batch, [w], mats)

    def _apply_matrix(self, x, batch, wires, mats) -> None:
        from . import _lib as L
        from . import engine
        n = self.nqubit
        key = (tuple(wires), x.dtype)
        plans = self.__dict__.setdefault('_plans', {})
        if key not in plans:
            targets = engine.wires_to_targets(n, wires)
            plans[key] = engine.FusedPlan(n, x.dtype, [L.make_gate(L.GATE_MAT, targets, (), 0)])
        m = mats.to(x.dtype).reshape(batch, -1).contiguous()
        plans[key].run(x, m, batch, m.shape[1])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import engine
        engine.require_cuda(x, 'the state')
        n = self.nqubit
        flat = x.reshape(-1, 2**n).contiguous().clone()
        with torch.no_grad():
            self.apply_(flat, flat.shape[0])
        return flat.reshape(x.shape)
