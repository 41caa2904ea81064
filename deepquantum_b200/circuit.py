"""QubitCircuit with the reference's builder / forward / expectation interface (circuit.py:81-1622),
executed as fused passes of hand-written sm_100a kernels.

`cir.h(0); cir.cnot(0, 1); cir.rx(1, 0.2); cir()` is a drop-in: the builder methods create the same
gate modules (same names, arguments and matrices), `forward` lowers `self.operators` to a gate
program once (cached), computes every gate matrix with a handful of vectorised torch calls (still
differentiable), and runs the program through `libb200q.so` -- the state is read and written once
per fused group instead of at least twice per gate (reference qmath.py:503-504).
"""
from __future__ import annotations

from copy import copy
from typing import Any

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import engine
from .gate import (Reset, Barrier, CNOT, Fredkin, Hadamard, HamiltonianGate, ImaginarySwap, LatentGate, PauliX, PauliY, PauliZ,
                   PhaseShift, ProjectionJ, ReconfigurableBeamSplitter, Rx, Rxx, Rxy, Ry, Ryy, Rz, Rzz, SDaggerGate,
                   SGate, Swap, TDaggerGate, TGate, Toffoli, U3Gate, UAnyGate)
from .layer import (CnotLayer, CnotRing, HLayer, Observable, RxLayer, RyLayer, RzLayer, U3Layer, XLayer, YLayer,
                    ZLayer)
from .channel import (AmplitudeDamping, BitFlip, Depolarizing, GeneralizedAmplitudeDamping, Pauli, PhaseDamping,
                      PhaseFlip)
from .operation import Channel, DenMatLowering, Gate, Layer, Lowering, Operation
from .state import QubitState, amplitude_encoding

# planner options used by every circuit (overridable for A/B measurements in bench.py)
PLAN_OPTIONS = {'chunk_bits': 0, 'low_bits': 0, 'max_rounds': 0, 'fuse': True}


class _Program:
    """Lowered gate program of a circuit + its fused plans per dtype."""

    def __init__(self, nqubit: int, ops, inverse: bool = False, den_mat: bool = False):
        self.low = DenMatLowering(nqubit) if den_mat else Lowering(nqubit)
        seq = list(ops)
        for op in (reversed(seq) if inverse else seq):
            op._lower(self.low, inverse)
        self.structs = self.low.finalize()
        self.plans = {}

    def plan(self, dtype: torch.dtype, **override) -> engine.FusedPlan:
        """Fused plan for `dtype`; `override` (e.g. chunk_bits=11 for the reverse sweep, which holds two tiles per CTA)
        replaces entries of PLAN_OPTIONS."""
        opts = {**PLAN_OPTIONS, **override}
        key = (dtype, tuple(sorted(opts.items())))
        if key not in self.plans:
            self.plans[key] = engine.FusedPlan(self.low.state_qubits, dtype, self.structs, **opts)
        return self.plans[key]

    @property
    def ngates(self) -> int:
        """Gate applications of the SOURCE circuit (the lowering may fuse `cnot; rz; cnot` into one diagonal)."""
        return self.low.n_source_gates


class QubitCircuit(Operation):
    """Quantum circuit on n qubits: state vectors, or density matrices with `den_mat=True` (forward only; run as a
    2n-qubit amplitude vector through the same fused kernels, see `operation.DenMatLowering`)."""

    def __init__(self, nqubit: int, init_state: Any = 'zeros', name: str | None = None, den_mat: bool = False,
                 reupload: bool = False, mps: bool = False, chi: int | None = None, shots: int = 1024) -> None:
        super().__init__(name=name, nqubit=nqubit, wires=None, den_mat=den_mat)
        if mps:
            raise NotImplementedError('the MPS back-end is outside the accelerated path (SURVEY.md section 2, #12)')
        self.reupload = reupload
        self.mps = mps
        self.chi = chi
        self.shots = shots
        self.set_init_state(init_state)
        self.operators = nn.Sequential()
        self.encoders = []
        self.observables = nn.ModuleList()
        self.state = None
        self.ndata = 0
        self.depth = np.array([0] * nqubit)
        self.wires_measure = []
        self.wires_condition = []
        self._program = None
        self._program_len = -1

    # ---------------------------------------------------------------------------------------------
    def set_init_state(self, init_state: Any) -> None:
        if isinstance(init_state, QubitState):
            assert self.nqubit == init_state.nqubit
            self.init_state = init_state
            self.den_mat = init_state.den_mat
        else:
            self.init_state = QubitState(nqubit=self.nqubit, state=init_state, den_mat=self.den_mat)

    def __add__(self, rhs: 'QubitCircuit') -> 'QubitCircuit':
        assert self.nqubit == rhs.nqubit
        cir = QubitCircuit(nqubit=self.nqubit, init_state=self.init_state, name=self.name, den_mat=self.den_mat,
                           reupload=self.reupload)
        cir.operators = self.operators + rhs.operators
        cir.encoders = self.encoders + rhs.encoders
        cir.observables = rhs.observables
        cir.npara = self.npara + rhs.npara
        cir.ndata = self.ndata + rhs.ndata
        cir.depth = self.depth + rhs.depth
        cir.wires_measure = rhs.wires_measure
        return cir

    def _apply(self, fn):
        super()._apply(fn)
        if self._program is not None:
            self._program.low._const_cache.clear()
        return self

    # ---------------------------------------------------------------------------------------------
    def _get_program(self) -> _Program:
        if self._program is None or self._program_len != len(self.operators):
            if any(isinstance(op, Reset) for op in self.operators):
                raise NotImplementedError('a circuit with Reset is not one linear program (no unitary / inverse / '
                                          'sharded form); forward() runs it as separate programs')
            self._program = _Program(self.nqubit, self.operators, den_mat=self.den_mat)
            self._program_len = len(self.operators)
        return self._program

    def _get_segments(self):
        """Gates between Reset operations as separate fused programs: [_Program | Reset, ...] (cached like _program)."""
        if self.__dict__.get('_segments') is None or self.__dict__.get('_segments_key') != len(self.operators):
            segs, cur = [], []
            for op in self.operators:
                if isinstance(op, Reset):
                    if cur:
                        segs.append(_Program(self.nqubit, cur, den_mat=self.den_mat))
                        cur = []
                    segs.append(op)
                else:
                    cur.append(op)
            if cur:
                segs.append(_Program(self.nqubit, cur, den_mat=self.den_mat))
            self.__dict__['_segments'] = segs
            self.__dict__['_segments_key'] = len(self.operators)
        return self.__dict__['_segments']

    def forward(self, data: torch.Tensor | None = None, state: Any = None) -> torch.Tensor:
        """Run the circuit; returns the final state `[2^n, 1]` (or `[batch, 2^n, 1]`), like the reference
        (circuit.py:180-242).  2-D `data` is an explicit batch (the reference uses `vmap` here)."""
        if state is None:
            state = self.init_state
        lazy_zero = isinstance(state, QubitState) and state.kind == 'zeros'
        if lazy_zero:   # never materialise |0...0>: the engine fills it with a kernel
            state_t = torch.empty(0, 1, dtype=state.dtype, device=state.device)
        else:
            state_t = state.state if isinstance(state, QubitState) else state
        if self.ndata == 0:
            data = None
        if data is None or data.ndim == 1:
            self.encode(data)
            out = self._run(state_t, None, lazy_zero)
        else:
            assert data.ndim == 2
            assert lazy_zero or state_t.ndim in (2, 3)
            self._encode_batched(data)
            try:
                out = self._run(state_t, data.shape[0], lazy_zero)
            finally:
                for op in self.encoders:
                    for g in (op.gates if isinstance(op, Layer) else [op]):
                        g._batched = None
            self.encode(data[-1])
        self.state = out
        return out

    def _run(self, state_t: torch.Tensor, data_batch: int | None, lazy_zero: bool) -> torch.Tensor:
        n = 2 * self.nqubit if self.den_mat else self.nqubit
        if self.den_mat and not lazy_zero:
            assert state_t.shape[-1] == 2**self.nqubit and state_t.shape[-2] == 2**self.nqubit
        engine.require_cuda(state_t, 'the circuit state (move the circuit with cir.to("cuda"))')
        cdtype = state_t.dtype
        if any(isinstance(op, Reset) for op in self.operators):
            return self._run_with_resets(state_t, data_batch, lazy_zero)
        prog = self._get_program()
        mats = prog.low.build_matrices(cdtype, state_t.device)
        batched_state = state_t.ndim == 3 and not lazy_zero
        nb_state = state_t.shape[0] if batched_state else 1
        batch = data_batch if data_batch is not None else nb_state
        if data_batch is not None and batched_state:
            assert nb_state == data_batch
        if lazy_zero:
            x = torch.empty(batch, 2**n, dtype=cdtype, device=state_t.device)
            engine.init_basis_(x, n, batch, 0)
        else:
            x = state_t.reshape(nb_state, 2**n)
            if nb_state != batch:
                x = x.expand(batch, -1)
            x = x.contiguous()
        mbs = mats.shape[-1] if mats.ndim == 2 else 0
        if mats.ndim == 2 and mats.shape[0] != batch:
            raise ValueError('batch of data and batch of states differ')
        if not self.den_mat and torch.is_grad_enabled() and (mats.requires_grad or x.requires_grad):
            from .adjoint import CircuitFunction
            y = CircuitFunction.apply(x, mats, prog, batch, mbs)
        else:
            # density matrices are forward only: channels are not invertible, so the reverse sweep of adjoint.py
            # (which un-computes the state with U^dagger) does not apply; the result carries no grad_fn
            y = x.detach() if lazy_zero else x.detach().clone()
            prog.plan(cdtype).run(y, mats.detach(), batch, mbs)
        y = y.reshape(batch, 2**self.nqubit, -1)
        if data_batch is None and not batched_state:
            y = y.squeeze(0)
        return y

    def _run_with_resets(self, state_t: torch.Tensor, data_batch: int | None, lazy_zero: bool) -> torch.Tensor:
        """Circuits with Reset (reference gate.py:3027-3094): program, projection, program, ... on one state buffer.
        Forward only: the projection depends on the state, and the reverse sweep needs invertible steps."""
        assert not self.den_mat
        n, cdtype = self.nqubit, state_t.dtype
        batched_state = state_t.ndim == 3 and not lazy_zero
        nb_state = state_t.shape[0] if batched_state else 1
        batch = data_batch if data_batch is not None else nb_state
        if lazy_zero:
            x = torch.empty(batch, 2**n, dtype=cdtype, device=state_t.device)
            engine.init_basis_(x, n, batch, 0)
        else:
            x = state_t.detach().reshape(nb_state, 2**n)
            x = (x.expand(batch, -1) if nb_state != batch else x).contiguous().clone()
        with torch.no_grad():
            for seg in self._get_segments():
                if isinstance(seg, Reset):
                    seg.apply_(x, batch)
                    continue
                mats = seg.low.build_matrices(cdtype, state_t.device)
                mbs = mats.shape[-1] if mats.ndim == 2 else 0
                if mats.ndim == 2 and mats.shape[0] != batch:
                    raise ValueError('batch of data and batch of states differ')
                seg.plan(cdtype).run(x, mats.detach(), batch, mbs)
        y = x.reshape(batch, 2**n, -1)
        if data_batch is None and not batched_state:
            y = y.squeeze(0)
        return y

    # ---------------------------------------------------------------------------------------------
    def encode(self, data: torch.Tensor | None) -> None:
        """Route `data` slices into the encoder gates (reference circuit.py:265-293).  Each gate also remembers
        which elements of `data` it took, so the matrix assembly can gather all angles of a gate class with one
        index_select instead of stacking hundreds of 0-d views."""
        if data is None:
            return
        if not self.reupload:
            assert len(data) >= self.ndata, 'The circuit needs more data, or consider data re-uploading'
        count = 0
        ndat = len(data)
        singles = data.split(1) if isinstance(data, torch.Tensor) and data.ndim == 1 else None   # all [1] views at once
        for op in self.encoders:
            # fast path (hundreds of one-parameter encoder gates per forward): same effect as op.init_para(data[i:i+1])
            # for a buffer-backed parameter, without nn.Module.__setattr__ / register_buffer per gate
            if (singles is not None and op.npara == 1 and count < ndat and getattr(op, '_fast_encode', False)
                    and not op.requires_grad and 'theta' in op._buffers):
                op._batched = None
                op._buffers['theta'] = singles[count]
                op.__dict__['_matrix_cache'] = None
                op._data_ref = (data, (count,))
                count = (count + 1) % ndat
                continue
            count_up = count + op.npara
            if self.reupload and count_up > ndat:
                n = int(np.ceil(count_up / ndat))
                op.init_para(torch.cat([data] * n)[count:count_up])
                idx = [i % ndat for i in range(count, count_up)]
            else:
                op.init_para(data[count:count_up])
                idx = list(range(count, count_up))
            k = 0
            for g in (op.gates if isinstance(op, Layer) else [op]):
                g._data_ref = (data, tuple(idx[k:k + g.npara]))
                k += g.npara
            count = count_up % ndat

    def _encode_batched(self, data: torch.Tensor) -> None:
        """2-D data: every encoder gate gets a `[batch, npara]` column block (explicit batch instead of the
        reference's vmap, circuit.py:227-241)."""
        width = data.shape[1]
        if not self.reupload:
            assert width >= self.ndata, 'The circuit needs more data, or consider data re-uploading'
        count = 0
        for op in self.encoders:
            gates = op.gates if isinstance(op, Layer) else [op]
            for g in gates:
                count_up = count + g.npara
                if count_up > width:
                    idx = [i % width for i in range(count, count_up)]
                    g._batched = data[:, idx]
                else:
                    g._batched = data[:, count:count_up]
                count = count_up % width

    def init_para(self) -> None:
        for op in self.operators:
            op.init_para()

    def init_encoder(self) -> None:
        for op in self.encoders:
            op.init_para()

    def reset_circuit(self, init_state: Any = 'zeros') -> None:
        self.set_init_state(init_state)
        self.operators = nn.Sequential()
        self.encoders = []
        self.observables = nn.ModuleList()
        self.state = None
        self.ndata = 0
        self.npara = 0
        self.depth = np.array([0] * self.nqubit)
        self.wires_measure = []
        self._program = None

    def amplitude_encoding(self, data: Any) -> torch.Tensor:
        return amplitude_encoding(data, self.nqubit)

    # ---------------------------------------------------------------------------------------------
    def observable(self, wires=None, basis: str = 'z') -> None:
        self.observables.append(Observable(nqubit=self.nqubit, wires=wires, basis=basis, den_mat=self.den_mat,
                                           tsr_mode=False))

    def reset_observable(self) -> None:
        self.observables = nn.ModuleList()

    def expectation(self, shots: int | None = None) -> torch.Tensor:
        """Exact expectation values of the Pauli-string observables on the final state: `[n_obs]` or
        `[batch, n_obs]` (reference circuit.py:381-428, qmath.py:830-860).  Z-strings are reduced by ONE fused
        kernel pass over the state for all observables; X/Y factors are rotated to Z on a copy first."""
        assert len(self.observables) > 0, 'There is no observable'
        assert isinstance(self.state, torch.Tensor), 'There is no final state'
        if shots is not None:
            return self._sampled_expectation(shots)
        from .adjoint import expectation_z
        n = self.nqubit
        st = self.state
        batched = st.ndim == 3
        dm = self.den_mat
        flat = st.reshape(-1, 4**n if dm else 2**n)
        batch = flat.shape[0]
        groups = {}
        for k, ob in enumerate(self.observables):
            rot = tuple(sorted((w[0], b) for w, b in zip(ob.wires, ob.basis) if b != 'z'))
            mask = 0
            for w in ob.wires:
                mask |= 1 << (n - 1 - w[0])
            groups.setdefault(rot, []).append((k, mask))
        out = [None] * len(self.observables)
        for rot, items in groups.items():
            phi = flat
            if rot:
                # R^dagger Z R = X for R = Ry(-pi/2), = Y for R = Rx(pi/2), evaluated in the state's precision: the
                # reference applies the Pauli matrices themselves (layer.py:156-166), which is exact -- a rotation by
                # its float32-rounded Hadamard constant (gate.py:1069) would cost 3e-8 in complex128
                basis_cir = QubitCircuit(n, den_mat=dm)
                half = torch.tensor(torch.pi / 2, dtype=flat.real.dtype)
                for w, b in rot:
                    if b == 'y':
                        basis_cir.rx(w, half)
                    else:
                        basis_cir.ry(w, -half)
                basis_cir.to(flat.device, flat.real.dtype)
                phi = basis_cir(state=flat.reshape(batch, 2**n, -1)).reshape(batch, -1)
            if dm:
                # Tr(Z-string rho) = sum_i (+-) Re rho_ii (reference qmath.py:856): the diagonal (2^n of the 4^n
                # entries, a strided view) goes through the same reduction kernel as amplitudes sqrt(rho_ii)
                diag = phi.reshape(batch, 2**n, 2**n).diagonal(dim1=-2, dim2=-1).real
                phi = torch.sqrt(diag.clamp_min(0)).to(flat.dtype).contiguous()
            masks = torch.tensor([m for _, m in items], dtype=torch.int64, device=flat.device)
            vals = expectation_z(phi, n, masks, batch)           # [batch, n_items] float64
            for j, (k, _) in enumerate(items):
                out[k] = vals[:, j]
        res = torch.stack(out, dim=-1).to(flat.real.dtype)
        return res if batched else res.squeeze(0)

    def _sampled_expectation(self, shots: int) -> torch.Tensor:
        """Shot-based estimate (reference circuit.py:400-426): rotate the observable's X / Y factors to Z, sample
        its wires on the device, average the parities of the outcomes."""
        from .qmath import measure as _measure, sample2expval
        self.shots = shots
        st = self.state
        rdtype, device = st.real.dtype, st.device
        out = []
        for ob in self.observables:
            basis_cir = QubitCircuit(self.nqubit, den_mat=self.den_mat)
            for w, b in zip(ob.wires, ob.basis):
                if b == 'y':
                    basis_cir.sdg(w[0])
                if b in ('x', 'y'):
                    basis_cir.h(w[0])
            basis_cir.to(device, rdtype)
            with torch.no_grad():
                rotated = basis_cir(state=st)
            samples = _measure(rotated, shots=shots, wires=sum(ob.wires, []), den_mat=self.den_mat)
            if isinstance(samples, list):
                expval = torch.cat([sample2expval(s).to(device, rdtype) for s in samples])
            else:
                expval = sample2expval(samples).to(device, rdtype)
                if st.ndim == 2:
                    expval = expval.squeeze(0)
            out.append(expval)
        return torch.stack(out, dim=-1)

    def measure(self, shots: int | None = None, with_prob: bool = False, wires=None, block_size: int = 2**24):
        """Measure the final state (reference circuit.py:338-379 -> qmath.measure, qmath.py:568-638): sampled on
        the device by the two-level inverse-CDF kernels (csrc/b200q_sample.cu)."""
        from .qmath import measure as _measure
        shots = self.shots if shots is None else shots
        self.shots = shots
        wires = list(range(self.nqubit)) if wires is None else self._convert_indices(wires)
        self.wires_measure = wires
        if self.state is None:
            return None
        assert isinstance(self.state, torch.Tensor), 'There is no final state'
        return _measure(self.state, shots=shots, with_prob=with_prob, wires=wires, den_mat=self.den_mat,
                        block_size=block_size)

    def get_unitary(self) -> torch.Tensor:
        """Global unitary (small n): the circuit applied to the identity (reference circuit.py:467-477)."""
        assert not self.den_mat
        dim = 2**self.nqubit
        eye = torch.eye(dim, dtype=self.init_state.dtype, device=self.init_state.device)
        prog = self._get_program()
        mats = prog.low.build_matrices(eye.dtype, eye.device)
        y = eye.contiguous().clone()
        prog.plan(eye.dtype).run(y, mats, dim, 0)
        return y.T

    def get_amplitude(self, bits: str) -> torch.Tensor:
        assert not self.den_mat
        assert isinstance(self.state, torch.Tensor), 'There is no final state'
        assert len(bits) == self.nqubit
        idx = int(bits, 2)
        return self.state.reshape(-1, 2**self.nqubit)[:, idx].squeeze(0) if self.state.ndim == 3 else \
            self.state.reshape(-1)[idx]

    def get_prob(self, bits: str, wires=None) -> torch.Tensor:
        assert isinstance(self.state, torch.Tensor), 'There is no final state'
        n = self.nqubit
        wires = list(range(n)) if wires is None else self._convert_indices(wires)
        assert len(bits) == len(wires)
        if self.den_mat:   # probabilities are the diagonal of rho
            diag = self.state.reshape(-1, 2**n, 2**n).diagonal(dim1=-2, dim2=-1).real.reshape(-1, *([2] * n))
            sel = [slice(None)] * (n + 1)
            for w, b in zip(wires, bits):
                sel[w + 1] = int(b)
            p = diag[tuple(sel)].reshape(diag.shape[0], -1).sum(-1)
            return p if self.state.ndim == 3 else p.squeeze(0)
        flat = self.state.reshape(-1, *([2] * n))
        sel = [slice(None)] * (n + 1)
        for w, b in zip(wires, bits):
            sel[w + 1] = int(b)
        sub = flat[tuple(sel)].reshape(flat.shape[0], -1)
        p = (sub.real**2 + sub.imag**2).sum(-1)
        return p if self.state.ndim == 3 else p.squeeze(0)

    def inverse(self, encode: bool = False) -> 'QubitCircuit':
        """Inverse circuit (reference circuit.py:530-555): reversed operators, each `op.inverse()`."""
        cir = QubitCircuit(nqubit=self.nqubit, name=self.name, den_mat=self.den_mat, reupload=self.reupload)
        # a layer is stored flattened in `operators` and as ONE entry of `encoders`: membership is by member gate
        enc_ids = {id(g) for e in self.encoders for g in (e.gates if isinstance(e, Layer) else [e])}
        for op in reversed(self.operators):
            op_inv = op.inverse()
            cir.operators.append(op_inv)
            if encode and id(op) in enc_ids:
                cir.encoders.append(op_inv)
        cir.npara, cir.ndata = (self.npara, self.ndata) if encode else (self.npara + self.ndata, 0)
        cir.depth = self.depth.copy()
        cir.to(self.init_state.device, torch.empty(0, dtype=self.init_state.dtype).real.dtype)
        return cir

    def max_depth(self) -> int:
        return int(max(self.depth))

    # ---------------------------------------------------------------------------------------------
    def add(self, op: Operation, encode: bool = False, wires=None, controls=None) -> None:
        """Append a gate, a layer or another circuit (reference circuit.py:820-897)."""
        assert isinstance(op, Operation)
        self._program = None
        if wires is not None:
            assert isinstance(op, Gate)
            controls = [] if controls is None else controls
            wires = self._convert_indices(wires)
            controls = self._convert_indices(controls)
            for wire in wires:
                assert wire not in controls, 'Use repeated wires'
            assert len(wires) == len(op.wires), 'Invalid input'
            op = copy(op)
            op.wires = wires
            op.controls = controls
        if isinstance(op, QubitCircuit):
            assert self.nqubit == op.nqubit
            self.operators += op.operators
            self.encoders += op.encoders
            self.observables = op.observables
            self.npara += op.npara
            self.ndata += op.ndata
            self.depth += op.depth
            self.wires_measure = op.wires_measure
            return
        op.tsr_mode = True
        if isinstance(op, Gate):
            self.operators.append(op)
            for i in op.wires + op.controls:
                self.depth[i] += 1
        elif isinstance(op, Layer):
            self.operators.extend(op.gates)
            for wire in op.wires:
                for i in wire:
                    self.depth[i] += 1
        elif isinstance(op, Channel):
            assert self.den_mat, 'Channels need the density-matrix representation (den_mat=True)'
            self.operators.append(op)
        else:
            raise NotImplementedError(f'{type(op).__name__} is outside the accelerated statevector path')
        if encode:
            assert not op.requires_grad, 'Please set requires_grad of the operation to be False'
            self.encoders.append(op)
            self.ndata += op.npara
        else:
            self.npara += op.npara

    # ---- builders (reference circuit.py:899-1622) ------------------------------------------------
    def _fixed(self, cls, wires, controls=None, condition=False):
        self.add(cls(nqubit=self.nqubit, wires=wires, controls=controls, condition=condition))

    def _param(self, cls, wires, inputs, controls=None, condition=False, encode=False, **kw):
        requires_grad = (not encode) and inputs is None
        self.add(cls(inputs=inputs, nqubit=self.nqubit, wires=wires, controls=controls, condition=condition,
                     requires_grad=requires_grad, **kw), encode=encode)

    def u3(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(U3Gate, wires, inputs, controls, condition, encode)

    def cu(self, control, target, inputs=None, encode=False):
        self._param(U3Gate, [target], inputs, [control], False, encode)

    def p(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(PhaseShift, wires, inputs, controls, condition, encode)

    def cp(self, control, target, inputs=None, encode=False):
        self._param(PhaseShift, [target], inputs, [control], False, encode)

    def x(self, wires, controls=None, condition=False):
        self._fixed(PauliX, wires, controls, condition)

    def y(self, wires, controls=None, condition=False):
        self._fixed(PauliY, wires, controls, condition)

    def z(self, wires, controls=None, condition=False):
        self._fixed(PauliZ, wires, controls, condition)

    def h(self, wires, controls=None, condition=False):
        self._fixed(Hadamard, wires, controls, condition)

    def s(self, wires, controls=None, condition=False):
        self._fixed(SGate, wires, controls, condition)

    def sdg(self, wires, controls=None, condition=False):
        self._fixed(SDaggerGate, wires, controls, condition)

    def t(self, wires, controls=None, condition=False):
        self._fixed(TGate, wires, controls, condition)

    def tdg(self, wires, controls=None, condition=False):
        self._fixed(TDaggerGate, wires, controls, condition)

    def ch(self, control, target):
        self._fixed(Hadamard, [target], [control])

    def cs(self, control, target):
        self._fixed(SGate, [target], [control])

    def csdg(self, control, target):
        self._fixed(SDaggerGate, [target], [control])

    def ct(self, control, target):
        self._fixed(TGate, [target], [control])

    def ctdg(self, control, target):
        self._fixed(TDaggerGate, [target], [control])

    def rx(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(Rx, wires, inputs, controls, condition, encode)

    def ry(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(Ry, wires, inputs, controls, condition, encode)

    def rz(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(Rz, wires, inputs, controls, condition, encode)

    def crx(self, control, target, inputs=None, encode=False):
        self._param(Rx, [target], inputs, [control], False, encode)

    def cry(self, control, target, inputs=None, encode=False):
        self._param(Ry, [target], inputs, [control], False, encode)

    def crz(self, control, target, inputs=None, encode=False):
        self._param(Rz, [target], inputs, [control], False, encode)

    def j(self, wires, inputs=None, plane='xy', controls=None, condition=False, encode=False):
        self._param(ProjectionJ, wires, inputs, controls, condition, encode, plane=plane)

    def cnot(self, control, target):
        self.add(CNOT(nqubit=self.nqubit, wires=[control, target]))

    def cx(self, control, target):
        self._fixed(PauliX, [target], [control])

    def cy(self, control, target):
        self._fixed(PauliY, [target], [control])

    def cz(self, control, target):
        self._fixed(PauliZ, [target], [control])

    def swap(self, wires, controls=None, condition=False):
        self._fixed(Swap, wires, controls, condition)

    def iswap(self, wires, controls=None, condition=False):
        self._fixed(ImaginarySwap, wires, controls, condition)

    def rxx(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(Rxx, wires, inputs, controls, condition, encode)

    def ryy(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(Ryy, wires, inputs, controls, condition, encode)

    def rzz(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(Rzz, wires, inputs, controls, condition, encode)

    def rxy(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(Rxy, wires, inputs, controls, condition, encode)

    def rbs(self, wires, inputs=None, controls=None, condition=False, encode=False):
        self._param(ReconfigurableBeamSplitter, wires, inputs, controls, condition, encode)

    def crxx(self, control, target1, target2, inputs=None, encode=False):
        self._param(Rxx, [target1, target2], inputs, [control], False, encode)

    def cryy(self, control, target1, target2, inputs=None, encode=False):
        self._param(Ryy, [target1, target2], inputs, [control], False, encode)

    def crzz(self, control, target1, target2, inputs=None, encode=False):
        self._param(Rzz, [target1, target2], inputs, [control], False, encode)

    def crxy(self, control, target1, target2, inputs=None, encode=False):
        self._param(Rxy, [target1, target2], inputs, [control], False, encode)

    def toffoli(self, control1, control2, target):
        self.add(Toffoli(nqubit=self.nqubit, wires=[control1, control2, target]))

    def ccx(self, control1, control2, target):
        self._fixed(PauliX, [target], [control1, control2])

    def fredkin(self, control, target1, target2):
        self.add(Fredkin(nqubit=self.nqubit, wires=[control, target1, target2]))

    def cswap(self, control, target1, target2):
        self._fixed(Swap, [target1, target2], [control])

    def any(self, unitary, wires=None, minmax=None, controls=None, name='uany'):
        self.add(UAnyGate(unitary=unitary, nqubit=self.nqubit, wires=wires, minmax=minmax, controls=controls,
                          name=name))

    def latent(self, wires=None, minmax=None, inputs=None, controls=None, encode=False, name='latent'):
        requires_grad = (not encode) and inputs is None
        self.add(LatentGate(inputs=inputs, nqubit=self.nqubit, wires=wires, minmax=minmax, controls=controls,
                            name=name, requires_grad=requires_grad), encode=encode)

    def hamiltonian(self, hamiltonian, t=None, wires=None, minmax=None, controls=None, encode=False,
                    name='hamiltonian'):
        """`exp(-i H t)` (reference circuit.py:1450-1476)."""
        requires_grad = (not encode) and t is None
        self.add(HamiltonianGate(hamiltonian=hamiltonian, t=t, nqubit=self.nqubit, wires=wires, minmax=minmax,
                                 controls=controls, name=name, den_mat=self.den_mat, requires_grad=requires_grad),
                 encode=encode)

    def _const_layer(self, cls, wires):
        self.add(cls(nqubit=self.nqubit, wires=wires))

    def _param_layer(self, cls, wires, inputs, encode):
        requires_grad = (not encode) and inputs is None
        self.add(cls(nqubit=self.nqubit, wires=wires, inputs=inputs, requires_grad=requires_grad), encode=encode)

    def xlayer(self, wires=None):
        self._const_layer(XLayer, wires)

    def ylayer(self, wires=None):
        self._const_layer(YLayer, wires)

    def zlayer(self, wires=None):
        self._const_layer(ZLayer, wires)

    def hlayer(self, wires=None):
        self._const_layer(HLayer, wires)

    def rxlayer(self, wires=None, inputs=None, encode=False):
        self._param_layer(RxLayer, wires, inputs, encode)

    def rylayer(self, wires=None, inputs=None, encode=False):
        self._param_layer(RyLayer, wires, inputs, encode)

    def rzlayer(self, wires=None, inputs=None, encode=False):
        self._param_layer(RzLayer, wires, inputs, encode)

    def u3layer(self, wires=None, inputs=None, encode=False):
        self._param_layer(U3Layer, wires, inputs, encode)

    def cxlayer(self, wires=None):
        self.add(CnotLayer(nqubit=self.nqubit, wires=wires))

    def cnot_ring(self, minmax=None, step=1, reverse=False):
        self.add(CnotRing(nqubit=self.nqubit, minmax=minmax, step=step, reverse=reverse))

    def barrier(self, wires=None):
        self.add(Barrier(nqubit=self.nqubit, wires=wires))

    # ---- channels (reference circuit.py:1540-1601) ------------------------------------------------
    def _channel(self, cls, wires, inputs, encode):
        assert self.den_mat
        requires_grad = (not encode) and inputs is None
        self.add(cls(inputs=inputs, nqubit=self.nqubit, wires=wires, requires_grad=requires_grad), encode=encode)

    def bit_flip(self, wires, inputs=None, encode=False):
        self._channel(BitFlip, wires, inputs, encode)

    def phase_flip(self, wires, inputs=None, encode=False):
        self._channel(PhaseFlip, wires, inputs, encode)

    def depolarizing(self, wires, inputs=None, encode=False):
        self._channel(Depolarizing, wires, inputs, encode)

    def pauli(self, wires, inputs=None, encode=False):
        self._channel(Pauli, wires, inputs, encode)

    def amp_damp(self, wires, inputs=None, encode=False):
        self._channel(AmplitudeDamping, wires, inputs, encode)

    def phase_damp(self, wires, inputs=None, encode=False):
        self._channel(PhaseDamping, wires, inputs, encode)

    def gen_amp_damp(self, wires, inputs=None, encode=False):
        self._channel(GeneralizedAmplitudeDamping, wires, inputs, encode)

    def reset(self, wires=None, postselect: int | None = 0) -> None:
        """Reset `wires` to |0> (reference circuit.py:1603-1607, gate.py:3027-3094); forward only."""
        assert not self.den_mat, 'Currently NOT supported'
        self.add(Reset(nqubit=self.nqubit, wires=wires, postselect=postselect))

    def move(self, wire1: int, wire2: int, postselect: int | None = 0) -> None:
        """Move (reference circuit.py:1619-1622, gate.py:3141-3168): a reset of `wire2` followed by a swap of the two
        wires.  Added as its two constituent operations (the reference wraps the same pair in one `Move` module; its
        quasi-probability decomposition for circuit cutting is outside the accelerated path)."""
        self.reset(wire2, postselect=postselect)
        self.swap([wire1, wire2])


class _ShardedExpectation(torch.autograd.Function):
    """Exact expectation values of a sharded final state, differentiable w.r.t. the matrix buffer by the adjoint
    (un-computing) method: the counterpart of the reference's `AdjointExpectation` (adjoint.py:19-83), which is
    wired to `DistributedQubitCircuit.expectation` only (circuit.py:1734-1738)."""

    @staticmethod
    def forward(ctx, mats, cir):
        ctx.cir = cir
        ctx.save_for_backward(mats)
        with torch.no_grad():
            return cir._expectation_values()

    @staticmethod
    def backward(ctx, grad):
        (mats,) = ctx.saved_tensors
        with torch.no_grad():
            gm = ctx.cir._expectation_backward(mats, grad)
        return gm.to(mats.dtype), None


class DistributedQubitCircuit(QubitCircuit):
    """Circuit on a statevector sharded over the ranks of the default process group (reference
    circuit.py:1625-1770).  `forward` is in place and `no_grad`, like the reference; it returns the
    `DistributedQubitState` whose `.amps` is this rank's slice of the final state in the reference layout."""

    def __init__(self, nqubit: int, name: str | None = None, reupload: bool = False, shots: int = 1024) -> None:
        super().__init__(nqubit=nqubit, init_state='zeros', name=name, reupload=reupload, shots=shots)
        self._sharded = None
        self._executor = None

    def set_init_state(self, init_state='zeros') -> None:
        from .distributed import DistributedQubitState
        if isinstance(init_state, DistributedQubitState):
            self.init_state = init_state
        else:
            self.init_state = DistributedQubitState(self.nqubit)

    def _apply(self, fn):
        nn.Module._apply(self, fn)
        if self._program is not None:
            self._program.low._const_cache.clear()
        return self

    @torch.no_grad()
    def forward(self, data: torch.Tensor | None = None, state=None):
        from .distributed import CudaExecutor, ShardedProgram
        if state is None:
            self.init_state.reset()
        else:
            self.init_state = state
        st = self.init_state
        with torch.enable_grad():
            self.encode(data)
        prog = self._get_program()
        if self._executor is None:
            engine.require_cuda(st.amps, 'the distributed state (move the circuit with cir.to(f"cuda:{local_rank}"))')
            self._executor = CudaExecutor()
        # with peer-mapped shards (NVLink) the exchanges are bit permutations done by the fused last pass of a segment
        # (shards below 2^6 amplitudes are padded to the register bits of a thread: those keep the NCCL transposes)
        mode = 'perm' if (hasattr(self._executor, 'run_plan_exchange') and st.world_size > 1
                          and st.log_num_amps_per_node - st.log_num_nodes >= 1 and st.log_num_amps_per_node >= 6
                          and getattr(st, 'enable_peer_exchange', lambda: False)()) else 'pswap'
        if self._sharded is None or self._sharded.low is not prog.low or self._sharded.mode != mode:
            self._sharded = ShardedProgram(prog.low, self.nqubit, st.world_size, st.rank, mode)
        mats = prog.low.build_matrices(st.amps.dtype, st.amps.device).detach()
        self._sharded.run(st, mats, self._executor, getattr(self, '_marks', None))
        self.state = st
        return st

    # ---- expectation values of the sharded state (reference circuit.py:1706-1758) -----------------------------
    def _clone_state(self, st):
        """A sharded state with its own shard + receive buffer holding a copy of `st.amps`."""
        from .distributed import DistributedQubitState
        new = DistributedQubitState(self.nqubit)
        new.register_buffer('amps', st.amps.detach().clone())
        new.register_buffer('buffer', torch.zeros_like(st.amps))
        return new

    def _basis_circuits(self, ob):
        """(rotation to the Z basis, its inverse) of an observable with X / Y factors, None for a Z string
        (reference circuit.py:1741-1749 builds the same circuit for the shot-based estimate)."""
        if set(ob.basis) == {'z'}:
            return None
        cache = self.__dict__.setdefault('_basis_cache', {})
        key = (tuple(w[0] for w in ob.wires), ob.basis)
        st = self.state
        if key not in cache or cache[key][2] != (st.amps.dtype, str(st.amps.device)):
            # R^dagger Z R = X for R = Ry(-pi/2), = Y for R = Rx(pi/2): rotations evaluated in the state's precision (the
            # reference's H is a float32-rounded constant, unitary to 6e-8 only, gate.py:1069)
            fwd, inv = DistributedQubitCircuit(self.nqubit), DistributedQubitCircuit(self.nqubit)
            rdtype = torch.empty(0, dtype=st.amps.dtype).real.dtype
            half = torch.tensor(torch.pi / 2, dtype=rdtype)
            for w, b in zip(ob.wires, ob.basis):
                if b == 'x':
                    fwd.ry(w[0], -half)
                    inv.ry(w[0], half)
                elif b == 'y':
                    fwd.rx(w[0], half)
                    inv.rx(w[0], -half)
            for c in (fwd, inv):
                c.to(st.amps.device, rdtype)
                c._executor = self._executor
            cache[key] = (fwd, inv, (st.amps.dtype, str(st.amps.device)))
        return cache[key][:2]

    def _ob_mask(self, ob) -> int:
        m = 0
        for w in ob.wires:
            m |= 1 << (self.nqubit - 1 - w[0])
        return m

    def _expectation_values(self) -> torch.Tensor:
        """<psi| O_k |psi> for every observable: Z strings in ONE fused reduction over the shard, observables with
        X / Y factors on a rotated copy; one all-reduce (reference distributed.py:288-294)."""
        import torch.distributed as dist
        st, ex = self.state, self._executor
        nl = st.log_num_amps_per_node
        dev = st.amps.device
        vals = torch.zeros(len(self.observables), dtype=torch.float64, device=dev)
        zidx = [k for k, ob in enumerate(self.observables) if set(ob.basis) == {'z'}]
        if zidx:
            mt = torch.tensor([self._ob_mask(self.observables[k]) for k in zidx], dtype=torch.int64, device=dev)
            vals[zidx] = ex.expectation_z(st.amps, nl, mt, st.rank << nl)
        for k, ob in enumerate(self.observables):
            if k in zidx:
                continue
            fwd, _ = self._basis_circuits(ob)
            tmp = fwd(state=self._clone_state(st))
            mt = torch.tensor([self._ob_mask(ob)], dtype=torch.int64, device=dev)
            vals[k] = ex.expectation_z(tmp.amps, nl, mt, st.rank << nl)[0]
        if dist.is_initialized() and st.world_size > 1:
            dist.all_reduce(vals)
        return vals.to(st.amps.real.dtype)

    def _expectation_backward(self, mats: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
        """Cotangent of the matrix buffer for L = sum_k grad_k <O_k>: seed lambda = 2 sum_k grad_k O_k psi, then the
        reverse sweep over the sharded schedule (reference adjoint.py:47-83; exchanges replayed backwards) and one
        all-reduce of the per-rank shares."""
        import torch.distributed as dist
        st, ex = self.state, self._executor
        nl = st.log_num_amps_per_node
        dev = st.amps.device
        g = grad.reshape(-1).to(torch.float64)
        lam = self._clone_state(st)
        lam.amps.zero_()
        zidx = [k for k, ob in enumerate(self.observables) if set(ob.basis) == {'z'}]
        if zidx:
            mt = torch.tensor([self._ob_mask(self.observables[k]) for k in zidx], dtype=torch.int64, device=dev)
            lam.amps += ex.apply_z_weights(st.amps, nl, mt, 2.0 * g[zidx], st.rank << nl).reshape(-1)
        for k, ob in enumerate(self.observables):
            if k in zidx:
                continue
            fwd, inv = self._basis_circuits(ob)
            tmp = fwd(state=self._clone_state(st))
            mt = torch.tensor([self._ob_mask(ob)], dtype=torch.int64, device=dev)
            tmp.amps.copy_(ex.apply_z_weights(tmp.amps, nl, mt, 2.0 * g[k:k + 1], st.rank << nl).reshape(-1))
            tmp = inv(state=tmp)
            lam.amps += tmp.amps
        psi = self._clone_state(st)
        if self._sharded.mode == 'perm':
            assert psi.enable_peer_exchange() and lam.enable_peer_exchange()
        gm = self._sharded.run_adjoint(psi, lam, mats.detach(), ex)
        if dist.is_initialized() and st.world_size > 1:
            flat = torch.view_as_real(gm.contiguous())
            dist.all_reduce(flat)
            gm = torch.view_as_complex(flat)
        return gm

    def expectation(self, shots: int | None = None) -> torch.Tensor:
        """Expectation values of the observables on the sharded final state (reference circuit.py:1706-1758).
        `shots=None`: exact and differentiable w.r.t. the gate parameters / encoded data by the adjoint method;
        otherwise a shot-based estimate (rank 0 returns the values, the other ranks empty tensors, like the
        reference)."""
        assert len(self.observables) > 0, 'There is no observable'
        assert self.state is not None, 'There is no final state'
        if shots is not None:
            return self._sampled_expectation(shots)
        prog = self._get_program()
        st = self.state
        with torch.enable_grad():
            mats = prog.low.build_matrices(st.amps.dtype, st.amps.device)
        if torch.is_grad_enabled() and mats.requires_grad:
            return _ShardedExpectation.apply(mats, self)
        with torch.no_grad():
            return self._expectation_values()

    def _sampled_expectation(self, shots: int) -> torch.Tensor:
        """Shot-based estimate (reference circuit.py:1737-1756): rotate a copy to the observable's eigenbasis, sample
        its wires with `measure_dist`, average the parities on rank 0."""
        from .distributed import measure_dist
        from .qmath import sample2expval
        self.shots = shots
        st = self.state
        rdtype, dev = st.amps.real.dtype, st.amps.device
        out = []
        for ob in self.observables:
            circs = self._basis_circuits(ob)
            tmp = st if circs is None else circs[0](state=self._clone_state(st))
            samples = measure_dist(tmp, shots=shots, wires=[w[0] for w in ob.wires], executor=self._executor)
            if st.rank == 0:
                out.append(sample2expval(samples).to(dev, rdtype).squeeze(0))
            else:
                out.append(torch.tensor([], dtype=rdtype, device=dev))
        return torch.stack(out, dim=-1)

    def measure(self, shots: int | None = None, with_prob: bool = False, wires=None, block_size: int = 2**24):
        """Measure the sharded final state (reference circuit.py:1677-1704 -> measure_dist)."""
        from .distributed import measure_dist
        shots = self.shots if shots is None else shots
        self.shots = shots
        wires = list(range(self.nqubit)) if wires is None else self._convert_indices(wires)
        self.wires_measure = wires
        if self.state is None:
            return None
        return measure_dist(self.state, shots=shots, with_prob=with_prob, wires=wires, block_size=block_size,
                            executor=self._executor)

    def cnot(self, control: int, target: int) -> None:
        self.cx(control, target)      # reference circuit.py:1764-1766: global controls then need no exchange

    def toffoli(self, control1: int, control2: int, target: int) -> None:
        self.ccx(control1, control2, target)
