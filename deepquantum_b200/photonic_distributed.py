"""Fock state tensor sharded over the ranks of the default process group (reference
photonic/state.py:623-685 `DistributedFockState`, photonic/distributed.py:31-103 `local_gate` / `dist_swap_gate` /
`dist_gate`, photonic/circuit.py:2849-2925 `DistributedQumodeCircuit`).

Layout as in the reference: `world_size = cutoff^g`; the first `g` modes are *global* -- their photon numbers are
the base-`cutoff` digits of the rank -- and every rank holds a `[cutoff] * (nmode - g)` tensor of the other modes.
A gate on local modes is one launch of the qudit kernel on the shard.  What differs from the reference is the
schedule: it swaps every global target mode in before the gate and back out after it (two all-to-alls among
`cutoff` ranks per global target per gate, distributed.py:92-101); here the circuit keeps a logical -> physical mode
map, a global mode is swapped in ONCE when a gate needs it, the local mode it evicts is the one whose next use is
farthest away (look-ahead over the gate list), and the reference layout is restored by swaps at the end of the
circuit.  Every rank enters every collective.
"""
from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist
from torch import nn

from .communication import comm_get_rank, comm_get_world_size
from .operation import apply_complex_fix
from .photonic import QumodeCircuit, qudit_apply_


def _digits(value: int, base: int, ndigit: int) -> list[int]:
    out = []
    for _ in range(ndigit):
        out.append(value % base)
        value //= base
    return out[::-1]              # most significant first


class CudaQuditExecutor:
    """Applies a local gate with the qudit kernel (csrc/b200q_qudit.cu)."""

    takes_structure = True    # `apply` accepts the block structure of the gate class (photonic._FockGate._structure)

    def apply(self, amps: torch.Tensor, nmode_local: int, cutoff: int, matrix: torch.Tensor, wires,
              structure: int = 0) -> None:
        qudit_apply_(amps.reshape(1, -1), nmode_local, cutoff, matrix, wires, 1, structure)


class DistributedFockState(nn.Module):
    """Fock state of `nmode` modes over `world_size = cutoff^g` ranks (reference photonic/state.py:623-685):
    `'vac'` / `'zeros'`, a Fock basis state `[1, 0, 0]`, or a superposition `[(amp, [1, 0]), ...]`."""

    def __init__(self, state: Any, nmode: int | None = None, cutoff: int | None = None) -> None:
        super().__init__()
        self.world_size = comm_get_world_size()
        self.rank = comm_get_rank()
        if state in ('vac', 'zeros'):
            state = [(1, [0] * nmode)]
        assert isinstance(state, list)
        if all(isinstance(i, int) for i in state):
            state = [(1.0, state)]
        assert all(isinstance(i, tuple) for i in state)
        nphoton = 0
        for _, occ in state:
            nphoton = max(nphoton, sum(occ))
            if nmode is None:
                nmode = len(occ)
        if cutoff is None:
            cutoff = nphoton + 1
        self.state, self.nmode, self.cutoff = state, nmode, cutoff
        g, w = 0, 1
        while w < self.world_size:
            w *= cutoff
            g += 1
        assert w == self.world_size, 'world_size must be a power of cutoff'
        assert cutoff**nmode >= self.world_size
        self.nmode_global, self.nmode_local = g, nmode - g
        self.num_amps_per_node = cutoff**self.nmode_local
        amps = torch.zeros([cutoff] * self.nmode_local, dtype=torch.cfloat)
        self.register_buffer('amps', amps)
        self.register_buffer('buffer', torch.zeros_like(amps))
        self.reset()

    def _apply(self, fn: Any) -> 'DistributedFockState':
        tensors = {k: self._buffers.pop(k) for k in ('amps', 'buffer')}
        super()._apply(fn)
        for key, value in apply_complex_fix(fn, tensors).items():
            self.register_buffer(key, value)
        return self

    def reset(self) -> None:
        self.amps.zero_()
        self.buffer.zero_()
        for amp, occ in self.state:
            owner = 0
            for digit in occ[:self.nmode_global]:
                owner = owner * self.cutoff + digit
            if owner == self.rank:
                self.amps[tuple(occ[self.nmode_global:])] = amp


def swap_global_local(state: DistributedFockState, global_pos: int, local_axis: int) -> None:
    """Exchange global digit `global_pos` (0 = most significant rank digit) with local tensor axis `local_axis`:
    one all-to-all among the `cutoff` ranks that differ in that digit (the mixed branch of the reference's
    `dist_swap_gate`, photonic/distributed.py:68-80).  Ranks outside the group exchange nothing but still enter."""
    d, g = state.cutoff, state.nmode_global
    x = state.amps.movedim(local_axis, 0).contiguous()              # [d, rest]: slice j goes to the rank with digit j
    shape = x.shape
    if state.world_size > 1:
        digits = _digits(state.rank, d, g)
        weight = d ** (g - 1 - global_pos)
        base_rank = state.rank - digits[global_pos] * weight
        sizes = [0] * state.world_size
        for j in range(d):
            sizes[base_rank + j * weight] = x[0].numel()
        out = state.buffer.reshape(-1)
        dist.all_to_all_single(out, x.reshape(-1), sizes, sizes)   # slices arrive in increasing source-digit order
        y = out.reshape(shape)
    else:
        y = x
    new = y.movedim(0, local_axis).contiguous()
    if state.world_size > 1:
        state.buffer = state.amps
    state.amps = new


class DistributedQumodeCircuit(QumodeCircuit):
    """`QumodeCircuit` on a sharded Fock tensor (reference photonic/circuit.py:2849-2925): `forward` is in place and
    `no_grad`, and returns the `DistributedFockState` whose `.amps` is this rank's shard in the reference layout."""

    def __init__(self, nmode: int, init_state: Any, cutoff: int | None = None, name: str | None = None) -> None:
        nn.Module.__init__(self)
        self.nmode, self.name, self.backend, self.basis = nmode, name, 'fock', False
        self.operators = nn.Sequential()
        self.state = None
        self.npara = 0
        self.cutoff = cutoff
        self._executor = None
        self.set_init_state(init_state)

    def set_init_state(self, init_state: Any = None) -> None:
        if isinstance(init_state, DistributedFockState):
            self.init_state = init_state
        else:
            self.init_state = DistributedFockState(init_state, self.nmode, self.cutoff)
        self.cutoff = self.init_state.cutoff

    def _next_use(self, start: int, logical_mode: int) -> int:
        for i in range(start, len(self.operators)):
            if logical_mode in self.operators[i].wires:
                return i
        return len(self.operators) + logical_mode          # never used again: evict first (ties by mode index)

    @torch.no_grad()
    def forward(self, state: DistributedFockState | None = None) -> DistributedFockState:
        if state is None:
            self.init_state.reset()
        else:
            self.init_state = state
        st = self.init_state
        ex = self._executor or CudaQuditExecutor()
        n, g, nl, d = self.nmode, st.nmode_global, st.nmode_local, self.cutoff
        mats = self.build_matrices(st.amps.dtype, st.amps.device)
        phys = list(range(n))            # phys[logical mode] = physical slot; slots < g are rank digits
        for i, (op, m) in enumerate(zip(self.operators, mats)):
            assert len(op.wires) <= nl, 'a gate needs all of its modes local at once'
            for w in op.wires:
                if phys[w] < g:          # swap the global mode in; evict the local mode used farthest in the future
                    candidates = [q for q in range(n) if phys[q] >= g and q not in op.wires]
                    victim = max(candidates, key=lambda q: self._next_use(i + 1, q))
                    swap_global_local(st, phys[w], phys[victim] - g)
                    phys[w], phys[victim] = phys[victim], phys[w]
            if getattr(ex, 'takes_structure', False):
                ex.apply(st.amps, nl, d, m, [phys[w] - g for w in op.wires], op._structure)
            else:
                ex.apply(st.amps, nl, d, m, [phys[w] - g for w in op.wires])
        # restore the reference layout: logical mode q at slot q
        def swap_modes(a, b):            # logical modes: a on a rank digit, b on a local axis
            swap_global_local(st, phys[a], phys[b] - g)
            phys[a], phys[b] = phys[b], phys[a]

        for slot in range(g):            # 1. the right logical mode into every global slot
            if phys[slot] == slot:
                continue
            current = phys.index(slot)   # the logical mode sitting on this rank digit
            if phys[slot] >= g:
                swap_modes(current, slot)
            else:                        # wanted mode on another rank digit: route through a local axis
                helper = next(q for q in range(n) if phys[q] >= g)
                swap_modes(current, helper)
                swap_modes(slot, current)
                swap_modes(helper, slot)
        order = [phys[q] - g for q in range(g, n)]           # 2. local axes back in order (a local permute)
        if order != list(range(nl)):
            st.amps = st.amps.permute(order).contiguous()
        self.state = st
        return st


__all__ = ['DistributedFockState', 'DistributedQumodeCircuit', 'swap_global_local', 'CudaQuditExecutor']
