"""Layers and Observable with the reference's interface (layer.py).  A layer is a group of gates on
disjoint wires -- a natural fusion unit: the planner puts a whole layer into one pass."""
from __future__ import annotations

from copy import deepcopy
from typing import Any

import torch
from torch import nn

from .gate import CNOT, Hadamard, PauliX, PauliY, PauliZ, Rx, Ry, Rz, U3Gate
from .operation import Layer


class SingleLayer(Layer):
    def __init__(self, name=None, nqubit=1, wires=None, den_mat=False, tsr_mode=False) -> None:
        if wires is None:
            wires = [[i] for i in range(nqubit)]
        super().__init__(name=name, nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
        for wire in self.wires:
            assert len(wire) == 1


class ParametricSingleLayer(SingleLayer):
    def __init__(self, name=None, nqubit=1, wires=None, den_mat=False, tsr_mode=False, requires_grad=True) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
        self.requires_grad = requires_grad

    def inverse(self):
        layer = deepcopy(self)
        gates = nn.Sequential()
        for gate in self.gates[::-1]:
            gates.append(gate.inverse())
        layer.gates = gates
        layer.wires = self.wires[::-1]
        return layer


class DoubleLayer(Layer):
    def __init__(self, name=None, nqubit=2, wires=None, den_mat=False, tsr_mode=False) -> None:
        if wires is None:
            wires = [[i, i + 1] for i in range(0, nqubit - 1, 2)]
        super().__init__(name=name, nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
        for wire in self.wires:
            assert len(wire) == 2


class Observable(SingleLayer):
    """Pauli-string observable (reference layer.py:127-165)."""

    def __init__(self, nqubit=1, wires=None, basis='z', den_mat=False, tsr_mode=False) -> None:
        super().__init__(name='Observable', nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
        basis = basis.lower()
        self.basis = basis * len(self.wires) if len(basis) == 1 else basis
        assert len(self.wires) == len(self.basis), 'The number of wires is not equal to the number of bases'
        for i, wire in enumerate(self.wires):
            cls = {'x': PauliX, 'y': PauliY, 'z': PauliZ}.get(self.basis[i])
            if cls is None:
                raise ValueError('Use illegal measurement basis')
            self.gates.append(cls(nqubit=nqubit, wires=wire, den_mat=den_mat, tsr_mode=True))


def _const_layer(gate_cls, layer_name):
    class _L(SingleLayer):
        def __init__(self, nqubit=1, wires=None, den_mat=False, tsr_mode=False) -> None:
            super().__init__(name=layer_name, nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
            for wire in self.wires:
                self.gates.append(gate_cls(nqubit=nqubit, wires=wire, den_mat=den_mat, tsr_mode=True))

    _L.__name__ = _L.__qualname__ = layer_name
    return _L


XLayer = _const_layer(PauliX, 'XLayer')
YLayer = _const_layer(PauliY, 'YLayer')
ZLayer = _const_layer(PauliZ, 'ZLayer')
HLayer = _const_layer(Hadamard, 'HLayer')


def _param_layer(gate_cls, layer_name, per_gate):
    class _L(ParametricSingleLayer):
        def __init__(self, nqubit=1, wires=None, inputs: Any = None, den_mat=False, tsr_mode=False,
                     requires_grad=True) -> None:
            super().__init__(name=layer_name, nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode,
                             requires_grad=requires_grad)
            for i, wire in enumerate(self.wires):
                if inputs is None:
                    theta = None
                elif per_gate == 1:
                    theta = inputs[i]
                else:
                    theta = inputs[per_gate * i:per_gate * (i + 1)]
                    if isinstance(theta, torch.Tensor):
                        theta = list(theta)
                gate = gate_cls(inputs=theta, nqubit=nqubit, wires=wire, den_mat=den_mat, tsr_mode=True,
                                requires_grad=requires_grad)
                self.gates.append(gate)
                self.npara += gate.npara

    _L.__name__ = _L.__qualname__ = layer_name
    return _L


RxLayer = _param_layer(Rx, 'RxLayer', 1)
RyLayer = _param_layer(Ry, 'RyLayer', 1)
RzLayer = _param_layer(Rz, 'RzLayer', 1)
U3Layer = _param_layer(U3Gate, 'U3Layer', 3)


class CnotLayer(DoubleLayer):
    def __init__(self, nqubit=2, wires=None, name='CnotLayer', den_mat=False, tsr_mode=False) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
        for wire in self.wires:
            self.gates.append(CNOT(nqubit=nqubit, wires=wire, den_mat=den_mat, tsr_mode=True))

    def inverse(self):
        return CnotLayer(nqubit=self.nqubit, wires=list(reversed(self.wires)), name=self.name,
                         tsr_mode=self.tsr_mode)


class CnotRing(CnotLayer):
    def __init__(self, nqubit=2, minmax=None, step=1, reverse=False, den_mat=False, tsr_mode=False) -> None:
        if minmax is None:
            minmax = [0, nqubit - 1]
        self.nqubit = nqubit
        self._check_minmax(minmax)
        assert minmax[0] < minmax[1]
        self.minmax, self.step, self.reverse = minmax, step, reverse
        nw = minmax[1] - minmax[0] + 1
        if reverse:
            wires = [[minmax[0] + i, minmax[0] + (i - step) % nw] for i in range(nw - 1, -1, -1)]
        else:
            wires = [[minmax[0] + i, minmax[0] + (i + step) % nw] for i in range(nw)]
        super().__init__(nqubit=nqubit, wires=wires, name='CnotRing', den_mat=den_mat, tsr_mode=tsr_mode)
