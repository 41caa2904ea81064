"""Operation / Gate base classes with the reference's interface (operation.py:16-409), lowered to
the b200q C ABI instead of permute/reshape/matmul.

Statevector path plus the density-matrix path (section 8f rank 2): a density matrix of
n qubits is run through the SAME kernels as a 2n-qubit amplitude vector (`DenMatLowering`): `U rho U^dagger` is
`U` on the row wire and `conj(U)` on the column wire (reference qmath.py:509-540), a Kraus channel is the
superoperator `sum_i K_i (x) conj(K_i)` on the (row, column) wire pair (reference operation.py:594-600), lowered by
its structure to a diagonal record, to Bell-basis / parity-block records, or to one dense 2-target record.
MPS arguments are accepted for signature compatibility and rejected with NotImplementedError.
"""
from __future__ import annotations

import os
from copy import copy
from typing import Any

import torch
from torch import nn

from . import _lib as L
from . import engine

dtype_map = {torch.float: torch.cfloat, torch.double: torch.cdouble}  # reference __init__.py:115-118


def apply_complex_fix(fn: Any, tensors: dict) -> dict:
    """`.to()` / `.double()` semantics of the reference (utils.py:45-50): probe which real dtype and
    device `fn` maps to and move complex buffers to the matching complex dtype."""
    first = next(iter(tensors.values()))
    probe = fn(torch.empty(0, dtype=first.real.dtype, device=first.device))
    target = dtype_map.get(probe.dtype, probe.dtype)
    return {k: v.to(probe.device, target) for k, v in tensors.items()}


class Operation(nn.Module):
    """Base class of quantum operations (reference operation.py:16-113)."""

    def __init__(self, name=None, nqubit: int = 1, wires=None, den_mat: bool = False, tsr_mode: bool = False) -> None:
        super().__init__()
        self.name = name
        self.nqubit = nqubit
        self.wires = wires
        self.den_mat = den_mat
        self.tsr_mode = tsr_mode
        self.npara = 0

    def tensor_rep(self, x: torch.Tensor) -> torch.Tensor:
        """[..., 2^n(,1)] -> [batch, 2, ..., 2] (reference operation.py:45-55)."""
        if self.den_mat:
            assert x.shape[-1] == 2**self.nqubit and x.shape[-2] == 2**self.nqubit
            return x.reshape([-1] + [2] * (2 * self.nqubit))
        if x.ndim == 1:
            assert x.shape[-1] == 2**self.nqubit
        else:
            assert x.shape[-1] == 2**self.nqubit or x.shape[-2] == 2**self.nqubit
        return x.reshape([-1] + [2] * self.nqubit)

    def vector_rep(self, x: torch.Tensor) -> torch.Tensor:
        return x.reshape(-1, 2**self.nqubit, 1)

    def matrix_rep(self, x: torch.Tensor) -> torch.Tensor:
        return x.reshape(-1, 2**self.nqubit, 2**self.nqubit)

    def get_unitary(self) -> torch.Tensor:
        raise NotImplementedError

    def init_para(self) -> None:
        pass

    def set_nqubit(self, nqubit: int) -> None:
        self.nqubit = nqubit

    def set_wires(self, wires) -> None:
        self.wires = self._convert_indices(wires)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.tensor_rep(x) if self.tsr_mode else self.vector_rep(x)

    def _convert_indices(self, indices) -> list[int]:
        if isinstance(indices, int):
            indices = [indices]
        assert isinstance(indices, list), 'Invalid input type'
        assert all(isinstance(i, int) for i in indices), 'Invalid input type'
        if len(indices) > 0:
            assert min(indices) > -1 and max(indices) < self.nqubit, 'Invalid input'
        assert len(set(indices)) == len(indices), 'Invalid input'
        return indices

    def _check_minmax(self, minmax: list[int]) -> None:
        assert isinstance(minmax, list)
        assert len(minmax) == 2
        assert all(isinstance(i, int) for i in minmax)
        assert -1 < minmax[0] <= minmax[1] < self.nqubit


class Lowering:
    """Collects the C-ABI gate records of a sequence of gates plus where each record's matrix comes
    from: a constant buffer, a per-forward `update_matrix()` call, or a batched per-class evaluation
    of all gates of one parametric class (one vectorised torch call instead of one per gate)."""

    def __init__(self, nqubit: int):
        self.nqubit = nqubit
        self.records = []      # (kind, targets, controls, adjoint, block, index, size)
        self.const = []        # tensors
        self.dynamic = []      # gates whose update_matrix() is evaluated every forward
        self.groups = {}       # class -> list of gates
        self.sources = []      # gate object per record (None for X records)
        self._const_index = {}
        self._const_cache = {}
        self._idx_cache = {}
        self.derived = []      # (source record index): 4x4 diagonals built from a 2x2 diagonal, see _peephole
        self.n_source_gates = 0

    def add(self, gate: 'Gate', kind: int, wires, controls, adjoint: bool = False, matrix_of: 'Gate | None' = None):
        n = self.nqubit
        targets = engine.wires_to_targets(n, wires)
        ctrl = [n - 1 - c for c in controls]
        size = 4 ** len(wires)
        src = matrix_of if matrix_of is not None else gate
        if kind == L.GATE_X:
            block, idx = 'none', 0
        elif src._matrix_source == 'const':
            # gates of a class with a class-wide constant matrix (H, S, T, ...) share one buffer entry
            key = type(src) if getattr(src, '_shared_const', False) else id(src)
            if key not in self._const_index:
                self._const_index[key] = len(self.const)
                self.const.append(src)
            block, idx = 'const', self._const_index[key]
        elif src._matrix_source == 'group':
            lst = self.groups.setdefault(type(src), [])
            block, idx = type(src), len(lst)
            lst.append(src)
        else:
            block, idx = 'dyn', len(self.dynamic)
            self.dynamic.append(src)
        hint = getattr(src, '_hint', 0) if (kind in (L.GATE_MAT, L.GATE_DIAG) and len(wires) == 1) else 0
        if (kind in (L.GATE_MAT, L.GATE_DIAG) and len(targets) >= 3 and block not in ('none', 'const')
                and (getattr(src, 'requires_grad', False) or getattr(src, '_data_ref', None) is not None
                     or getattr(src, '_batched', None) is not None)):
            hint |= L.GATE_GRAD   # trainable / data-fed dense gate on >= 3 wires: own pass, full cotangent in the reverse sweep
        self.records.append((kind, tuple(targets), tuple(ctrl), bool(adjoint), block, idx, size, hint))
        self.sources.append(src if kind != L.GATE_X else None)

    FUSE_CX_DIAG_CX = True

    def _peephole(self):
        """CX(a->b) . D(b) . CX(a->b)  ==  diag(d0, d1, d1, d0) on (b, a) for a 1-target diagonal D = diag(d0, d1):
        the `cnot; rz; cnot` idiom of the reference's QAOA / ZZ-feature-map circuits (examples/qaoa.py:36-40)
        becomes ONE diagonal op -- no amplitude moves, no tile-bit requirement -- whose matrix entries are
        gathered from D's matrix in `build_matrices` (so autograd, forward and adjoint, chains through D)."""
        recs, srcs, out_r, out_s = self.records, self.sources, [], []
        i = 0
        while i < len(recs):
            a = recs[i]
            if (self.FUSE_CX_DIAG_CX and i + 2 < len(recs) and a[0] == L.GATE_X and len(a[2]) == 1
                    and recs[i + 2][:3] == a[:3] and recs[i + 1][0] == L.GATE_DIAG and recs[i + 1][1] == a[1]
                    and recs[i + 1][2] == () and recs[i + 1][4] != 'none'):
                d = recs[i + 1]
                # matrix-index bit 0 acts on the target b, bit 1 on the control a
                out_r.append((L.GATE_DIAG, (a[1][0], a[2][0]), (), d[3], 'derived', len(self.derived), 16, 0))
                out_s.append(srcs[i + 1])
                self.derived.append(d)
                i += 3
            else:
                out_r.append(a)
                out_s.append(srcs[i])
                i += 1
        self.records, self.sources = out_r, out_s

    def finalize(self):
        """Assign matrix-buffer offsets; returns (GateStruct list, layout description)."""
        self.n_source_gates = len(self.records)
        self._peephole()
        sizes = {'const': [0] * len(self.const), 'dyn': [0] * len(self.dynamic)}
        for cls, lst in self.groups.items():
            sizes[cls] = [0] * len(lst)
        sizes['derived'] = [16] * len(self.derived)
        for kind, _t, _c, _a, block, idx, size, _h in self.records:
            if block != 'none':
                sizes[block][idx] = size
        for _k, _t, _c, _a, block, idx, size, _h in self.derived:   # the source matrices must be laid out too
            sizes[block][idx] = size
        bases, offs, total = {}, {}, 0
        for block in ['const', 'dyn'] + list(self.groups) + ['derived']:
            bases[block] = total
            acc, lst = 0, []
            for s in sizes[block]:
                lst.append(acc)
                acc += s
            offs[block] = lst
            total += acc
        self.offsets = [0 if r[4] == 'none' else bases[r[4]] + offs[r[4]][r[5]] for r in self.records]
        self.total = max(total, 1)
        # gather indices of the derived diagonals: entry (j, j) of diag(d0, d1, d1, d0) <- source (0,0) / (1,1);
        # the off-diagonal entries (never read by the kernel) <- source (0,1), a structural zero
        self._derived_src = None
        if self.derived:
            idx = []
            for d in self.derived:
                so = bases[d[4]] + offs[d[4]][d[5]]
                blockidx = [so + 1] * 16
                for j, e in ((0, 0), (1, 3), (2, 3), (3, 0)):
                    blockidx[5 * j] = so + e
                idx.extend(blockidx)
            self._derived_src = idx
            self.n_primary = bases['derived']
        return self._make_structs()

    def _make_structs(self):
        return [L.make_gate(kind, targets, ctrl, off, adj, hint)
                for (kind, targets, ctrl, adj, _b, _i, _s, hint), off in zip(self.records, self.offsets)]

    @property
    def state_qubits(self) -> int:
        """Qubits of the amplitude vector the lowered program runs on."""
        return self.nqubit

    def _gather_from_data(self, cls, lst):
        """All parameters of a gate class as ONE gather from the encoded data vector, when every gate of the class
        took its parameters from the same `data` tensor in the last `encode` (and is not inverted / batched)."""
        ref = getattr(lst[0], '_data_ref', None)
        if ref is None or not hasattr(lst[0], '_pnames'):    # channels: parameters are stacked one by one
            return None
        data = ref[0]
        idx = []
        for g in lst:
            r = getattr(g, '_data_ref', None)
            if r is None or r[0] is not data or g._batched is not None or getattr(g, 'inv_mode', False):
                return None
            if len(g._pnames) != len(r[1]) or any(getattr(g, nm).data_ptr() != data[i].data_ptr()
                                                  for nm, i in zip(g._pnames[:1], r[1][:1])):
                return None
            idx.extend(r[1])
        key = (cls, tuple(idx), str(data.device))
        cached = self._idx_cache.get(cls)
        if cached is None or cached[0] != key[1] or cached[1].device != data.device:
            cached = (key[1], torch.tensor(idx, dtype=torch.int64, device=data.device))
            self._idx_cache[cls] = cached
        return data.index_select(0, cached[1])

    def structure_key(self):
        return tuple(r[:4] + (r[6], r[7]) for r in self.records)

    def build_matrices(self, cdtype: torch.dtype, device, batch: int | None = None) -> torch.Tensor:
        """Flat device buffer of every matrix ([total] or [batch, total]); differentiable w.r.t. the gate
        parameters."""
        parts = []
        if self.const:
            key = (cdtype, str(device))
            if key not in self._const_cache:   # invalidated by QubitCircuit._apply (dtype / device moves)
                self._const_cache[key] = torch.cat([g.matrix.reshape(-1).to(device=device, dtype=cdtype)
                                                    for g in self.const])
            parts.append(self._const_cache[key])
        if self.dynamic:
            parts.append(torch.cat([getattr(g, '_lowered_matrix', g.update_matrix)().reshape(-1)
                                    for g in self.dynamic]).to(device=device,
                                                                                             dtype=cdtype))
        batched = False
        for cls, lst in self.groups.items():
            p = self._gather_from_data(cls, lst)
            if p is None:
                plist = [t for g in lst for t in g._param_list()]
                nbs = {t.shape[0] for t in plist if t.ndim == 1}
                if nbs:   # encoder gates of a batched forward next to plain gates of the same class: broadcast
                    assert len(nbs) == 1, 'gates of one class were encoded with different batch sizes'
                    nb = nbs.pop()
                    plist = [t if t.ndim == 1 else t.reshape(1).expand(nb) for t in plist]
                p = torch.stack(plist)                   # [N*npara] or [N*npara, batch]
            if p.ndim == 2:
                batched = True
                p = p.transpose(0, 1).reshape(p.shape[1], len(lst), -1)
            else:
                p = p.reshape(len(lst), -1)
            # evaluate in float64 and round ONCE to the state dtype: CUDA's float32 sin/cos are 1-2 ulp, which
            # at depth 40 costs more accuracy than the whole complex64 simulation (measured: 2.0e-6 vs 7e-7)
            m = cls._batched_matrix(p.to(device=device, dtype=torch.float64))        # [..., N, d, d]
            parts.append(m.reshape(*m.shape[:-3], -1).to(cdtype))
        if not parts:
            return torch.zeros(1, dtype=cdtype, device=device)
        if batched:
            nb = next(x.shape[0] for x in parts if x.ndim == 2)
            parts = [x if x.ndim == 2 else x.unsqueeze(0).expand(nb, -1) for x in parts]
            flat = torch.cat(parts, dim=-1).contiguous()
        else:
            flat = torch.cat(parts)
        if self._derived_src:
            key = ('derived', str(device))
            if key not in self._idx_cache:
                self._idx_cache[key] = torch.tensor(self._derived_src, dtype=torch.int64, device=device)
            flat = torch.cat([flat, flat.index_select(-1, self._idx_cache[key])], dim=-1)
        return flat


class DenMatLowering(Lowering):
    """Lowering of a density-matrix circuit onto the 2n-qubit amplitude vector `rho[i, j] -> index i * 2^n + j`
    (row wire w = bit 2n-1-w, column wire w = bit n-1-w): every gate record becomes the pair `U` on the row wires,
    `conj(U)` on the column wires (reference qmath.py:509-540 does the two matrix products one after the other, each
    with a permute + reshape copy of the 4^n-element state); a channel becomes the superoperator
    `sum_i K_i (x) conj(K_i)` on (row wires, column wires) (reference operation.py:594-600 evolves one copy of the
    state per Kraus operator and sums them), emitted by `add_super` as the cheapest record sequence its structure
    allows.  The matrix buffer is `[flat | conj(flat)]`.  Row and column records act on
    disjoint bits, so the planner is free to fuse them into the same pass."""

    # A/B switch: False lowers Pauli channels through the generic parity blocks
    PAULI_BELL = os.environ.get('B200Q_DENMAT_PAULI_BELL', '1') != '0'
    # measured slower (14-qubit mixed workload: 34 passes / 94.1 ms against 28 passes / 80.6 ms): off by default
    DAMPING_SVD = os.environ.get('B200Q_DENMAT_DAMPING_SVD', '0') != '0'

    def add_super(self, chan: 'Channel', wires) -> None:
        n = self.nqubit
        w2 = list(wires) + [w + n for w in wires]
        targets = engine.wires_to_targets(2 * n, w2)
        # channels whose Kraus operators are all diagonal (phase flip, phase damping) have a DIAGONAL superoperator:
        # no tile bits, no amplitude moves, fusable anywhere
        kind, size = 'super', 4 ** len(w2)
        if getattr(chan, '_diagonal_kraus', False):
            kind = 'super_diag'
        elif getattr(chan, '_pauli_kraus', False) and len(wires) == 1 and self.PAULI_BELL:
            # Pauli channels sum_k p_k P_k rho P_k: the four superoperators P_k (x) conj(P_k) commute, their common
            # eigenbasis is the Bell basis, so the channel is CX(row->col) . H(row) . diag . H(row) . CX(row->col):
            # five records, all of them in-place ("lean") kernel ops
            kind, size = 'super_pauli', 20
        elif getattr(chan, '_damping_kraus', False) and len(wires) == 1 and self.DAMPING_SVD:
            # (generalised) amplitude damping: the parity-1 block is a scalar, the parity-0 block a real 2x2 =
            # rotation . diagonal . rotation (SVD): CX, X(col), c-Ry, diag, c-Ry, X(col), CX -- lean ops only
            kind, size = 'super_damping', 24
        elif getattr(chan, '_parity_kraus', False) and len(wires) == 1:
            # every Kraus operator diagonal or anti-diagonal: the superoperator keeps the parity row ^ column, so it is
            # CX(row->col) . [M1 on row if parity 1, M0 on row if parity 0] . CX(row->col) -- register-kind ops only
            # (a dense 2-target op needs its own kind of round inside a pass and blocks the fusion of its
            # neighbours: 12-qubit noisy workload of tools/bench_denmat.py, 47 passes / 232 rounds with dense
            # superoperators, 21 passes / 69 rounds with the parity blocks)
            kind, size = 'super_parity', 8
        if hasattr(chan, '_batched_matrix'):    # built-in channels: one batched evaluation per class and forward
            lst = self.groups.setdefault(type(chan), [])
            block, idx = type(chan), len(lst)
            lst.append(chan)
        else:
            block, idx = 'dyn', len(self.dynamic)
            self.dynamic.append(chan)
        self.records.append((kind, tuple(targets), (), False, block, idx, size, 0))
        self.sources.append(chan)

    @property
    def state_qubits(self) -> int:
        return 2 * self.nqubit

    def _make_structs(self):
        n, out = self.nqubit, []
        for (kind, targets, ctrl, adj, _b, _i, _s, hint), off in zip(self.records, self.offsets):
            if kind in ('super', 'super_diag'):
                out.append(L.make_gate(L.GATE_MAT if kind == 'super' else L.GATE_DIAG, targets, (), off, False, 0))
                continue
            if kind == 'super_pauli':
                col, row = targets
                cx = L.make_gate(L.GATE_X, [col], [row], 0, False, 0)
                had = L.make_gate(L.GATE_MAT, [row], [], off + 16, False, L.GATE_REAL | L.GATE_HADAMARD)
                out += [cx, had, L.make_gate(L.GATE_DIAG, targets, (), off, False, 0), had, cx]
                continue
            if kind == 'super_damping':
                col, row = targets
                cx = L.make_gate(L.GATE_X, [col], [row], 0, False, 0)
                flip = L.make_gate(L.GATE_X, [col], [], 0, False, 0)
                rot = L.GATE_REAL | L.GATE_ROTATION
                out += [cx, flip, L.make_gate(L.GATE_MAT, [row], [col], off, False, rot),
                        L.make_gate(L.GATE_DIAG, targets, (), off + 4, False, 0),
                        L.make_gate(L.GATE_MAT, [row], [col], off + 20, False, rot), flip, cx]
                continue
            if kind == 'super_parity':
                col, row = targets
                cx = L.make_gate(L.GATE_X, [col], [row], 0, False, 0)
                flip = L.make_gate(L.GATE_X, [col], [], 0, False, 0)
                out += [cx, L.make_gate(L.GATE_MAT, [row], [col], off, False, 0), flip,
                        L.make_gate(L.GATE_MAT, [row], [col], off + 4, False, 0), flip, cx]
                continue
            out.append(L.make_gate(kind, [t + n for t in targets], [c + n for c in ctrl], off, adj, hint))
            out.append(L.make_gate(kind, targets, ctrl, 0 if kind == L.GATE_X else off + self.total, adj,
                                   L.conj_hint(hint)))
        return out

    def build_matrices(self, cdtype: torch.dtype, device, batch: int | None = None) -> torch.Tensor:
        flat = super().build_matrices(cdtype, device, batch)
        return torch.cat([flat, flat.conj().resolve_conj()], dim=-1).contiguous()


class Gate(Operation):
    """Base class of gates (reference operation.py:116-409).

    Class attributes consumed by the lowering:
      _kind           B200Q gate kind the planner may assume (structure only, never values)
      _matrix_source  'const' (registered `matrix` buffer), 'group' (`_batched_matrix` over all gates of
                      the class) or 'dyn' (`update_matrix()` every forward)
    """

    _kind = L.GATE_MAT
    _matrix_source = 'const'
    _hint = 0   # B200Q_GATE_REAL / B200Q_GATE_RXLIKE structure hint of the class's 2x2 matrix

    def __init__(self, name=None, nqubit: int = 1, wires=None, controls=None, condition: bool = False,
                 den_mat: bool = False, tsr_mode: bool = False) -> None:
        self.nqubit = nqubit
        if wires is None:
            wires = [0]
        if controls is None:
            controls = []
        wires = self._convert_indices(wires)
        controls = self._convert_indices(controls)
        for wire in wires:
            assert wire not in controls, 'Use repeated wires'
        if condition:
            raise NotImplementedError('conditional measurement (condition=True) is outside the accelerated path')
        super().__init__(name=name, nqubit=nqubit, wires=wires, den_mat=den_mat, tsr_mode=tsr_mode)
        self.controls = controls
        self.condition = condition

    # ---- dtype / device moves: complex buffers follow the real dtype (operation.py:156-169) ---------
    def _apply(self, fn: Any) -> 'Gate':
        if self.npara > 0:
            super()._apply(fn)
        else:
            tensor = self._buffers.pop('matrix') if 'matrix' in self._buffers else None
            super()._apply(fn)
            if tensor is not None:
                self.register_buffer('matrix', apply_complex_fix(fn, {'matrix': tensor})['matrix'])
        return self

    def set_controls(self, controls) -> None:
        self.controls = self._convert_indices(controls)

    def get_matrix(self, inputs: Any) -> torch.Tensor:
        return self.matrix

    def update_matrix(self) -> torch.Tensor:
        return self.matrix

    def get_derivative(self, inputs: Any) -> torch.Tensor:
        return torch.zeros_like(self.matrix)

    def get_unitary(self) -> torch.Tensor:
        """Global 2^n x 2^n unitary (small n only), built column by column like the non-local branch of
        the reference (gate.py:326-331): apply the gate to the identity."""
        dim = 2**self.nqubit
        matrix = self.update_matrix()
        eye = torch.eye(dim, dtype=matrix.dtype, device=matrix.device)
        out = self._apply_to_batch(eye.contiguous().clone(), dim)
        return out.T

    # ---- lowering -----------------------------------------------------------------------------------
    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        low.add(self, self._kind, self.wires, self.controls, adjoint=inverse)

    def _apply_to_batch(self, flat: torch.Tensor, batch: int) -> torch.Tensor:
        """Apply this gate in place to a contiguous [batch, 2^n] tensor."""
        low = Lowering(self.nqubit)
        self._lower(low)
        low.finalize()
        mats = low.build_matrices(flat.dtype, flat.device)
        for (kind, targets, ctrl, adj, block, _i, size, _h), off in zip(low.records, low.offsets):
            m = None if block == 'none' else mats[off:off + size]
            engine.apply_gate_(flat, self.nqubit, m, targets, ctrl, kind, adj, batch)
        return flat

    def op_state(self, x: torch.Tensor) -> torch.Tensor:
        """Out-of-place application to a `[batch, 2, ..., 2]` tensor (reference operation.py:191-197)."""
        shape = x.shape
        flat = x.reshape(-1, 2**self.nqubit).contiguous().clone()
        flat = self._apply_to_batch(flat, flat.shape[0])
        x = flat.reshape(shape)
        if not self.tsr_mode:
            x = self.vector_rep(x).squeeze(0)
        return x

    def op_den_mat(self, x: torch.Tensor) -> torch.Tensor:
        """Out-of-place `U rho U^dagger` on a `[batch, 2, ..., 2]` (2n axes) tensor (reference operation.py:221-229)."""
        shape = x.shape
        flat = x.reshape(-1, 4**self.nqubit).contiguous().clone()
        _run_den_mat(self, flat)
        x = flat.reshape(shape)
        if not self.tsr_mode:
            x = self.matrix_rep(x).squeeze(0)
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not isinstance(x, torch.Tensor):
            return self.op_dist_state(x)
        if not self.tsr_mode:
            x = self.tensor_rep(x)
        if self.den_mat:
            assert x.ndim == 2 * self.nqubit + 1
            return self.op_den_mat(x)
        assert x.ndim == self.nqubit + 1
        return self.op_state(x)

    def op_dist_state(self, x):
        """Sharded state (reference operation.py:265-272): delegated to the distributed state object."""
        return x.apply_gate(self)

    def inverse(self) -> 'Gate':
        return self

    def extra_repr(self) -> str:
        s = f'wires={self.wires}'
        return s if self.controls == [] else s + f', controls={self.controls}'


class Layer(Operation):
    """A set of gates on disjoint wires (reference operation.py:412-522)."""

    def __init__(self, name=None, nqubit: int = 1, wires=None, den_mat: bool = False, tsr_mode: bool = False) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=None, den_mat=den_mat, tsr_mode=tsr_mode)
        self.nqubit = nqubit
        if wires is None:
            wires = [[0]]
        self.wires = self._convert_indices(wires)
        self.gates = nn.Sequential()

    def _convert_indices(self, indices) -> list[list[int]]:
        if isinstance(indices, int):
            indices = [[indices]]
        assert isinstance(indices, list), 'Invalid input type'
        if all(isinstance(i, int) for i in indices):
            indices = [[i] for i in indices]
        assert all(isinstance(i, list) for i in indices), 'Invalid input type'
        for idx in indices:  # per-gate checks only: rings may reuse a wire across gates (operation.py:496-507)
            assert all(isinstance(i, int) for i in idx), 'Invalid input type'
            assert min(idx) > -1 and max(idx) < self.nqubit, 'Invalid input'
            assert len(set(idx)) == len(idx), 'Invalid input'
        return indices

    def set_nqubit(self, nqubit: int) -> None:
        self.nqubit = nqubit
        for g in self.gates:
            g.nqubit = nqubit

    def init_para(self, inputs: Any = None) -> None:
        count = 0
        for g in self.gates:
            if inputs is None:
                g.init_para()
            else:
                g.init_para(inputs[count:count + g.npara])
            count += g.npara

    def update_npara(self) -> None:
        self.npara = sum(g.npara for g in self.gates)

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        for g in (reversed(self.gates) if inverse else self.gates):
            g._lower(low, inverse)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.tsr_mode:
            x = self.tensor_rep(x)
        shape = x.shape
        if self.den_mat:
            flat = x.reshape(-1, 4**self.nqubit).contiguous().clone()
            _run_den_mat(self, flat)
            x = flat.reshape(shape)
            return x if self.tsr_mode else self.matrix_rep(x).squeeze(0)
        flat = x.reshape(-1, 2**self.nqubit).contiguous().clone()
        for g in self.gates:
            g._apply_to_batch(flat, flat.shape[0])
        x = flat.reshape(shape)
        if not self.tsr_mode:
            return self.vector_rep(x).squeeze(0)
        return x

    def inverse(self) -> 'Layer':
        return self


def _run_den_mat(op: Operation, flat: torch.Tensor) -> None:
    """Apply one operation (gate, layer or channel) in place to a contiguous `[batch, 4^n]` density matrix."""
    engine.require_cuda(flat, 'the density matrix')
    low = DenMatLowering(op.nqubit)
    op._lower(low)
    structs = low.finalize()
    with torch.no_grad():
        mats = low.build_matrices(flat.dtype, flat.device)
        engine.FusedPlan(2 * op.nqubit, flat.dtype, structs).run(flat, mats, flat.shape[0], 0)


_BELL_CACHE = {}


class Channel(Operation):
    """Base class of quantum channels (reference operation.py:525-625).  `get_matrix` / `update_matrix` return the
    stacked Kraus operators like the reference; the lowering uses the superoperator built from them."""

    def __init__(self, inputs: Any = None, name=None, nqubit: int = 1, wires=None, tsr_mode: bool = False,
                 requires_grad: bool = False) -> None:
        self.nqubit = nqubit
        if wires is None:
            wires = [0]
        wires = self._convert_indices(wires)
        super().__init__(name=name, nqubit=nqubit, wires=wires, den_mat=True, tsr_mode=tsr_mode)
        self.npara = 1
        self.requires_grad = requires_grad
        self.init_para(inputs)

    @property
    def prob(self):
        return torch.sin(self.theta) ** 2

    def inputs_to_tensor(self, inputs: Any = None) -> torch.Tensor:
        while isinstance(inputs, list):
            inputs = inputs[0]
        if inputs is None:
            inputs = torch.rand(1)[0] * torch.pi
        elif not isinstance(inputs, torch.Tensor):
            inputs = torch.tensor(inputs, dtype=torch.float)
        return inputs

    def get_matrix(self, theta: Any) -> torch.Tensor:
        raise NotImplementedError

    def update_matrix(self) -> torch.Tensor:
        matrix = self.get_matrix(self.theta)
        self.matrix = matrix.detach()
        return matrix

    _diagonal_kraus = False   # all Kraus operators diagonal: the superoperator is a diagonal gate
    _parity_kraus = False     # all Kraus operators diagonal or anti-diagonal (one wire): two 2x2 parity blocks
    _pauli_kraus = False      # Kraus operators proportional to Pauli matrices: diagonal in the Bell basis
    _damping_kraus = False    # parity-1 block proportional to the identity, parity-0 block real (amplitude damping)

    @classmethod
    def _lower_kraus(cls, k: torch.Tensor) -> torch.Tensor:
        """Kraus operators `[..., n_kraus, d, d]` -> flat lowered blocks `[..., size]`: the superoperator
        `sum_i K_i (x) conj(K_i)` as `[d^2, d^2]` (row wires are the high matrix-index bits), or, for
        parity-preserving one-wire channels, its two 2x2 blocks `[M1 | M0]` acting on the row bit when
        row ^ column = 1 / 0 (see `DenMatLowering.add_super`)."""
        d = k.shape[-1]
        sup = torch.einsum('...iab,...icd->...acbd', k, k.conj())          # [..., row', col', row, col]
        if cls._pauli_kraus and d == 2 and DenMatLowering.PAULI_BELL:
            # [4x4 diagonal in the Bell basis | exact Hadamard]: B = (H on row) . CX(row->col), diag = B S B^T
            key = (sup.dtype, str(k.device))
            if key not in _BELL_CACHE:     # constants: built (and copied to the device) once, not every forward
                r = 0.5 ** 0.5
                _BELL_CACHE[key] = (
                    torch.tensor([[r, 0, 0, r], [0, r, r, 0], [r, 0, 0, -r], [0, r, -r, 0]], dtype=sup.real.dtype,
                                 device=k.device).to(sup.dtype),
                    torch.tensor([r, r, r, -r], dtype=sup.real.dtype, device=k.device).to(sup.dtype))
            bell, had = _BELL_CACHE[key]
            s4 = sup.reshape(*sup.shape[:-4], 4, 4)
            dd = torch.einsum('ij,...jk,ik->...i', bell, s4, bell)
            return torch.cat([torch.diag_embed(dd).reshape(*dd.shape[:-1], 16), had.expand(*dd.shape[:-1], 4)], dim=-1)
        if cls._parity_kraus and not cls._diagonal_kraus and d == 2:
            i = torch.arange(2, device=k.device)
            m0 = sup[..., i[:, None], i[:, None], i[None, :], i[None, :]]
            m1 = sup[..., i[:, None], 1 - i[:, None], i[None, :], 1 - i[None, :]]
            if cls._damping_kraus and DenMatLowering.DAMPING_SVD:
                # [Vh (rotation) | 4x4 diagonal in the flipped parity frame | U (rotation)], M0 = U diag(sig) Vh
                u, sig, vh = torch.linalg.svd(m0.real)
                du, dv = torch.linalg.det(u).sign(), torch.linalg.det(vh).sign()      # reflections -> rotations
                u = torch.cat([u[..., :, :1], u[..., :, 1:] * du[..., None, None]], dim=-1)
                vh = torch.cat([vh[..., :1, :], vh[..., 1:, :] * dv[..., None, None]], dim=-2)
                sig = torch.stack([sig[..., 0], sig[..., 1] * du * dv], dim=-1)
                scal = m1[..., 0, 0].real                                              # parity-1 block = scal * 1
                dd = torch.stack([scal, sig[..., 0], scal, sig[..., 1]], dim=-1)       # index row * 2 + flipped parity
                parts = [vh.reshape(*vh.shape[:-2], 4), torch.diag_embed(dd).reshape(*dd.shape[:-1], 16),
                         u.reshape(*u.shape[:-2], 4)]
                return torch.cat(parts, dim=-1).to(sup.dtype)
            return torch.cat([m1.reshape(*m1.shape[:-2], 4), m0.reshape(*m0.shape[:-2], 4)], dim=-1)
        return sup.reshape(*sup.shape[:-4], d**4)

    def _lowered_matrix(self) -> torch.Tensor:
        return self._lower_kraus(self.update_matrix())

    def init_para(self, inputs: Any = None) -> None:
        theta = self.inputs_to_tensor(inputs)
        if self.requires_grad:
            self.theta = nn.Parameter(theta)
        else:
            self.register_buffer('theta', theta)
        self.update_matrix()

    def _lower(self, low: Lowering, inverse: bool = False) -> None:
        assert isinstance(low, DenMatLowering), 'Channels act on density matrices (den_mat=True)'
        low.add_super(self, self.wires)

    def op_den_mat(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        flat = x.reshape(-1, 4**self.nqubit).contiguous().clone()
        _run_den_mat(self, flat)
        x = flat.reshape(shape)
        if not self.tsr_mode:
            x = self.matrix_rep(x).squeeze(0)
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.tsr_mode:
            x = self.tensor_rep(x)
        assert x.ndim == 2 * self.nqubit + 1
        return self.op_den_mat(x)

    def inverse(self) -> 'Channel':
        return self

    def extra_repr(self) -> str:
        return f'wires={self.wires}, probability={self.prob.item()}'


__all__ = ['Operation', 'Gate', 'Layer', 'Channel', 'Lowering', 'DenMatLowering', 'apply_complex_fix', 'dtype_map', 'copy']
