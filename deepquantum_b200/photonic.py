"""Photonic Fock-tensor back-end (the `basis=False`, pure-state path of the reference's QumodeCircuit):
per-mode gate contraction on a `[batch, cutoff, ..., cutoff]` state tensor
(photonic/circuit.py:405-431 -> photonic/operation.py:142-146 -> qmath.evolve_state with qudit=cutoff).

Only what that hot path needs is mirrored: `QumodeCircuit(nmode, 'vac', cutoff, backend='fock', basis=False)`
with the `ps` / `bs` / `s` builders, the beamsplitter family (`mzi`, `bs_theta`, `bs_phi`, `bs_rx`, `bs_ry`, `bs_h`,
`dc`, `h`), the rotations `r` / `f`, the Kerr gates `k` / `ck`, the displacement `d` and the two-mode squeezer `s2`, and the gate classes behind them, whose
Fock-space transformation matrices follow the same recurrences (arXiv:2004.11002 Eq. 51-52, 74-75) but are
evaluated with a handful of vectorised torch calls for ALL gates of a class at once, on the device -- the
reference's per-element Python loops (photonic/gate.py:356-373, 1098-1114) cost 33 ms per beamsplitter,
100x the kernel time on a B200.  Gaussian / Bosonic / permanent back-ends are out of scope.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any

import torch
from torch import nn

from . import _lib as L
from . import engine
from .operation import apply_complex_fix

# A/B switch: 1 = fused Fock passes (b200q_qudit_fused: several gates per read + write of the state).  Measured on C5
# (8 modes, cutoff 10): 13 passes instead of 36 launches but 40.8 ms against 39.7 ms -- the fused kernel executes
# 330 thread instructions per amplitude and pass (ELL gather from the shared tile), so it only matches the per-gate
# kernel; off by default until its contraction is register-blocked
FUSE_FOCK = os.environ.get('B200Q_FOCK_FUSE', '0') != '0'


STRUCTURED_FOCK = os.environ.get('B200Q_FOCK_STRUCTURED', '1') != '0'   # block-structured kernels by gate class


def qudit_apply_(flat: torch.Tensor, nmode: int, d: int, matrix: torch.Tensor, wires, batch: int = 1,
                 structure: int = L.QUDIT_GENERAL) -> None:
    """In-place `evolve_state(state, matrix, nmode, wires, qudit=d)` on a contiguous [batch, d^nmode] tensor.
    `structure` (L.QUDIT_*): block structure of the matrix known from the gate CLASS (b200q_qudit_apply_structured)."""
    engine.require_cuda(flat, 'the Fock state tensor')
    if not flat.is_contiguous() or flat.numel() != batch * d**nmode:
        raise L.B200QError('state must be contiguous with batch * cutoff^nmode elements')
    m = matrix.to(flat.dtype).contiguous()
    w = (C.c_int32 * len(wires))(*[int(x) for x in wires])
    L.check(L.load().b200q_qudit_apply_structured(flat.data_ptr(), nmode, d, engine.dtype_code(flat.dtype), m.data_ptr(),
                                                  w, len(wires), structure if STRUCTURED_FOCK else L.QUDIT_GENERAL,
                                                  batch, engine._stream(flat)))


GROUP_FOCK = os.environ.get('B200Q_FOCK_GROUP', '1') != '0'   # one pass for a two-mode gate + its one-mode neighbours
GROUP_MAX_OPS = 4


def qudit_apply_group_(flat: torch.Tensor, nmode: int, d: int, tile_modes, ops, batch: int = 1) -> None:
    """`ops` = [(matrix, wires, structure), ...] applied in order, all inside the two `tile_modes`, with one read and one
    write of the state (b200q_qudit_apply_group)."""
    engine.require_cuda(flat, 'the Fock state tensor')
    if not flat.is_contiguous() or flat.numel() != batch * d**nmode:
        raise L.B200QError('state must be contiguous with batch * cutoff^nmode elements')
    keep = [m.to(flat.dtype).contiguous() for m, _, _ in ops]          # alive until the launch is enqueued
    arr = (L.QuditOpStruct * len(ops))()
    for a, m, (_, wires, structure) in zip(arr, keep, ops):
        a.n_targets = len(wires)
        for j, w in enumerate(wires):
            a.modes[j] = int(w)
        a.structure = int(structure)
        a.matrix = m.data_ptr()
    tm = (C.c_int32 * 2)(int(tile_modes[0]), int(tile_modes[1]))
    L.check(L.load().b200q_qudit_apply_group(flat.data_ptr(), nmode, d, engine.dtype_code(flat.dtype), tm, arr, len(ops),
                                             batch, engine._stream(flat)))


def plan_fock_groups(ops_info, nmode: int, d: int):
    """Group the gates of a Fock circuit for `b200q_qudit_apply_group`.  `ops_info[i] = (wires, structure)`.
    A structured two-mode gate absorbs the structured one-mode gates waiting on its two modes (they commute with
    everything in between, which does not touch those modes) and, afterwards, the one-mode gates that follow it directly
    on its modes.  A group is only formed where it saves a pass over the state: three or more gates, two when the pair
    contains the lowest mode (staged in shared memory anyway), or any number of DIAGONAL neighbours (phase shifter /
    Kerr next to a beamsplitter: folded into the gate's packed blocks, the gate then costs what it costs alone).  Returns [[gate indices in execution order], ...]."""
    out, absorbing = [], []          # absorbing[g]: group g is a two-mode group that may still take one-mode gates
    pend = {}                        # mode -> one-mode structured gates waiting for a two-mode gate on that mode
    last = {}                        # mode -> index in `out` of the last group that touched the mode

    def emit(group, modes, can_absorb):
        out.append(group)
        absorbing.append(can_absorb)
        for m in modes:
            last[m] = len(out) - 1

    def flush(m):
        for j in pend.pop(m, []):
            emit([j], [m], False)

    for i, (wires, structure) in enumerate(ops_info):
        ok = GROUP_FOCK and structure != L.QUDIT_GENERAL and d <= 16 and nmode >= 2
        if ok and len(wires) == 1:
            m = wires[0]
            g = last.get(m)
            if (g is not None and absorbing[g] and len(out[g]) < GROUP_MAX_OPS and not pend.get(m)
                    and (len(out[g]) >= 2 or (nmode - 1) in ops_info[out[g][-1]][0] or structure == L.QUDIT_DIAG)):
                out[g].append(i)     # follows its two-mode gate directly on this mode
            else:
                pend.setdefault(m, []).append(i)
        elif ok and len(wires) == 2:
            a, b = wires
            pre = pend.get(a, []) + pend.get(b, [])
            lowest = (nmode - 1) in (a, b)
            all_diag = all(ops_info[j][1] == L.QUDIT_DIAG for j in pre)     # folded into the gate's blocks: free
            if pre and len(pre) + 1 <= GROUP_MAX_OPS and (len(pre) >= 2 or lowest or all_diag):
                pend.pop(a, None)
                pend.pop(b, None)
                emit(sorted(pre) + [i], [a, b], True)
            else:
                flush(a)
                flush(b)
                emit([i], [a, b], True)
        else:
            for m in wires:
                flush(m)
            emit([i], list(wires), False)
    for m in sorted(pend):
        flush(m)
    return out


FOCK_TILE_MAX = 12288     # amplitudes of a fused pass's shared-memory tile (b200q_qudit_fused)


def plan_fock_passes(gate_modes, nmode: int, d: int, max_gates: int = L.QUDIT_FUSED_MAX_GATES):
    """Group the gates of a Fock circuit (`gate_modes[i]` = modes of gate i, circuit order) into fused passes: each pass
    owns T tile modes (cutoff^T amplitudes staged in shared memory) and runs, in order, every pending gate whose modes
    lie in the tile and that does not have to wait for an earlier gate outside it (gates on disjoint modes commute).
    The last mode is kept in every tile whenever there is room: global accesses are then runs of `cutoff` amplitudes.
    Returns [(sorted tile modes, [gate indices])]; the reference applies the same gates one by one
    (photonic/circuit.py:405-431), each with its own pass over the state."""
    T = 1
    while T < nmode and d ** (T + 1) <= FOCK_TILE_MAX:
        T += 1
    n = len(gate_modes)
    done = [False] * n
    passes = []

    def scan(tile):
        blocked, out = set(), []
        for i in range(n):
            if done[i]:
                continue
            ms = set(gate_modes[i])
            if ms <= tile and not (ms & blocked):
                out.append(i)
                if len(out) == max_gates:
                    break
            else:
                blocked |= ms
        return out

    while not all(done):
        first = done.index(False)
        tile = set(gate_modes[first])
        assert len(tile) <= T, 'gate does not fit a fused tile'
        if len(tile) < T:
            tile.add(nmode - 1)
        while len(tile) < T:
            # the earliest pending gate that is not executable yet and whose missing modes still fit
            blocked, add = set(), None
            for i in range(n):
                if done[i]:
                    continue
                ms = set(gate_modes[i])
                if ms <= tile and not (ms & blocked):
                    continue
                if not (ms & blocked) and len(tile | ms) <= T:
                    add = ms - tile
                    break
                blocked |= ms
            if add is None:
                add = {next(m for m in range(nmode - 1, -1, -1) if m not in tile)}
            tile |= add
        ids = scan(tile)
        for i in ids:
            done[i] = True
        passes.append((sorted(tile), ids))
    return passes


def qudit_fused_(flat: torch.Tensor, nmode: int, d: int, tile_modes, gates, matrices: torch.Tensor, batch: int = 1) -> None:
    """One fused pass (`b200q_qudit_fused`): `gates` = [(modes, element offset of the matrix in `matrices`)]."""
    engine.require_cuda(flat, 'the Fock state tensor')
    arr = (L.QuditGateStruct * len(gates))()
    for g, (modes, off) in zip(arr, gates):
        g.n_targets = len(modes)
        for j, m in enumerate(modes):
            g.modes[j] = int(m)
        g.mat_offset = int(off)
    tm = (C.c_int32 * len(tile_modes))(*[int(x) for x in tile_modes])
    L.check(L.load().b200q_qudit_fused(flat.data_ptr(), nmode, d, engine.dtype_code(flat.dtype), tm, len(tile_modes), arr,
                                       len(gates), matrices.data_ptr(), batch, engine._stream(flat)))


# ---- Fock-space transformation matrices, batched over gates ------------------------------------------------
def ps_matrix_state(theta: torch.Tensor, d: int) -> torch.Tensor:
    """[N] -> [N, d, d]: diag(exp(i theta n))   (photonic/gate.py:192-194)."""
    n = torch.arange(d, dtype=theta.dtype, device=theta.device)
    return torch.diag_embed(torch.exp(1j * theta[:, None] * n[None, :]))


NATIVE_FOCK_MATRICES = os.environ.get('B200Q_FOCK_NATIVE_MATRICES', '1') != '0'


def _native_matrices_ok(t: torch.Tensor, d: int, dmax: int) -> bool:
    """One launch per gate class instead of the torch recurrences: forward-only (no autograd graph), float64 parameters
    on the device."""
    return (NATIVE_FOCK_MATRICES and t.is_cuda and d <= dmax and t.shape[0] >= 1
            and not (torch.is_grad_enabled() and t.requires_grad)
            and t.dtype in (torch.float64, torch.complex128))


def squeezing_matrix_state(r: torch.Tensor, theta: torch.Tensor, d: int) -> torch.Tensor:
    """[N], [N] -> [N, d, d]   (photonic/gate.py:1091-1114, arXiv:2004.11002 Eq. 51-52): column n+1 from
    columns n and n-1, vectorised over rows and gates."""
    if _native_matrices_ok(r, d, 64) and theta.dtype == torch.float64:
        prm = torch.stack([r, theta], dim=1).contiguous()
        out = torch.empty(r.shape[0], d, d, dtype=torch.complex128, device=r.device)
        L.check(L.load().b200q_fock_squeezing_matrix(prm.data_ptr(), r.shape[0], d, L.C128, out.data_ptr(),
                                                     engine._stream(out)))
        return out
    rt = r.dtype
    sq = torch.sqrt(torch.arange(d, dtype=rt, device=r.device))
    sech = 1 / torch.cosh(r)
    e_it_tanh = torch.exp(1j * theta) * torch.tanh(r)
    e_m_it_tanh = torch.exp(-1j * theta) * torch.tanh(r)
    cols = []
    col0 = [torch.sqrt(sech) + 0j]
    for m in range(1, d):      # rank 1: only even rows are non-zero
        if m % 2 == 0:
            col0.append(-sq[m - 1] / sq[m] * e_it_tanh * col0[m - 2])
        else:
            col0.append(torch.zeros_like(col0[0]))
    cols.append(torch.stack(col0, dim=-1))                     # [N, d]
    mm = torch.arange(d, device=r.device)
    for n in range(d - 1):     # rank 2: T[m, n+1] for (m + n) odd
        prev = cols[n]
        shifted = torch.cat([torch.zeros_like(prev[:, :1]), prev[:, :-1]], dim=-1)      # T[m-1, n]
        term = sq[None, :] / sq[n + 1] * sech[:, None] * shifted
        if n >= 1:
            term = term + sq[n] / sq[n + 1] * e_m_it_tanh[:, None] * cols[n - 1]
        mask = ((mm + n) % 2 == 1).to(rt)
        cols.append(term * mask[None, :])
    return torch.stack(cols, dim=-1)                           # [N, d(m), d(n)]


def displacement_matrix_state(r: torch.Tensor, theta: torch.Tensor, d: int) -> torch.Tensor:
    """[N], [N] -> [N, d, d]   (photonic/gate.py:1431-1451, arXiv:2004.11002 Eq. 57-58): column 0 is the coherent
    state, column n+1 follows from column n and its shift; vectorised over rows and gates."""
    sq = torch.sqrt(torch.arange(d, dtype=r.dtype, device=r.device))
    alpha = r * torch.exp(1j * theta)
    alpha_c = r * torch.exp(-1j * theta)
    col = [torch.exp(-(r**2) / 2) + 0j]
    for m in range(d - 1):
        col.append(alpha / sq[m + 1] * col[m])
    cur = torch.stack(col, dim=-1)                       # [N, d]
    cols = [cur]
    for n in range(d - 1):
        shifted = torch.cat([torch.zeros_like(cur[:, :1]), cur[:, :-1]], dim=-1)        # T[m - 1, n], zero for m = 0
        cur = (-alpha_c[:, None] * cur + sq[None, :] * shifted) / sq[n + 1]
        cols.append(cur)
    return torch.stack(cols, dim=-1)                     # [N, m, n]


def _int_powers(base: torch.Tensor, d: int) -> torch.Tensor:
    """[N] -> [N, d]: base^0 .. base^(d-1) by a cumulative product.  `base ** arange(d)` evaluates complex 0**0 as
    NaN in PyTorch, which poisons the whole transform (and its gradient) for an exact zero entry: a beamsplitter at
    theta = 0, a squeezer at r = 0."""
    ones = torch.ones_like(base)[:, None]
    if d == 1:
        return ones
    return torch.cumprod(torch.cat([ones, base[:, None].expand(-1, d - 1)], dim=1), dim=1)


def squeezing2_matrix_state(r: torch.Tensor, theta: torch.Tensor, d: int) -> torch.Tensor:
    """[N], [N] -> [N, d, d, d, d] (index m, n, p, q; photonic/gate.py:1258-1290, arXiv:2004.11002 Eq. 64-67).
    Only entries with m - n = p - q are non-zero.  With A_q[m, n] = T[m, n, q + m - n, q] the rank-4 recurrence reads
    A_q[m, n] = sech sqrt(n / q) A_{q-1}[m, n-1] - e^{-i theta} tanh sqrt((q + m - n) / q) A_{q-1}[m, n]: a sweep over
    q, vectorised over (m, n) and the gates; A_0 (ranks 2 and 3) is a sweep over m."""
    nb = r.shape[0]
    sq = torch.sqrt(torch.arange(d, dtype=r.dtype, device=r.device))
    sech = (1 / torch.cosh(r))[:, None]
    t_p = (torch.exp(1j * theta) * torch.tanh(r))[:, None]
    t_m = (torch.exp(-1j * theta) * torch.tanh(r))[:, None]
    idx = torch.arange(d, device=r.device)
    # A_0[m, n] = T[m, n, m - n, 0], m >= n
    diag = sech * _int_powers(t_p[:, 0], d)                              # [N, n]: T[n, n, 0, 0]
    rows = []
    for m in range(d):
        if m == 0:
            row = torch.zeros(nb, d, dtype=diag.dtype, device=r.device)
        else:                                                            # sech sqrt(m / (m - n)) A_0[m - 1, n], n < m
            fac = torch.where(idx < m, sq[m] / sq[(m - idx).clamp(min=1)], torch.zeros_like(sq))
            row = sech * fac[None, :] * rows[m - 1]
        row = torch.where(idx[None, :] == m, diag, row)
        rows.append(row)
    a = torch.stack(rows, dim=1)                                         # [N, m, n]
    out = torch.zeros(nb, d, d, d, d, dtype=a.dtype, device=r.device)
    mm, nn_ = idx[:, None].expand(d, d), idx[None, :].expand(d, d)
    for q in range(d):
        pp = q + mm - nn_
        valid = (pp >= 0) & (pp < d)
        if q > 0:
            shifted = torch.cat([torch.zeros_like(a[:, :, :1]), a[:, :, :-1]], dim=2)    # A_{q-1}[m, n - 1]
            a = (sech[:, :, None] * (sq[nn_] / sq[q])[None] * shifted
                 - t_m[:, :, None] * (sq[pp.clamp(0, d - 1)] / sq[q])[None] * a)
            a = a * valid[None]
        sel = valid.nonzero(as_tuple=True)
        out[:, sel[0], sel[1], pp[sel], q] = a[:, sel[0], sel[1]]
    return out


def bs_matrix_state(u: torch.Tensor, d: int) -> torch.Tensor:
    """[N, 2, 2] mode-mixing unitaries -> [N, d, d, d, d] Fock transformation tensors T[m, n, p, q]
    (photonic/gate.py:347-374, arXiv:2004.11002 Eq. 74-75): the q-recurrence vectorised over (m, n, p, gate)."""
    if _native_matrices_ok(u, d, 16):
        uc = u.contiguous()
        out = torch.empty(u.shape[0], d, d, d, d, dtype=torch.complex128, device=u.device)
        L.check(L.load().b200q_fock_bs_matrix(uc.data_ptr(), u.shape[0], d, L.C128, out.data_ptr(), engine._stream(out)))
        return out
    rt = u.real.dtype
    dev = u.device
    N = u.shape[0]
    idx = torch.arange(d, device=dev)
    sq = torch.sqrt(idx.to(rt))
    lf = torch.lgamma(idx.to(rt) + 1)
    m, n, p = torch.meshgrid(idx, idx, idx, indexing='ij')
    # rank 3 (q = 0, p = m + n): sqrt(p! / (m! n!)) u00^m u10^n  -- the closed form of the reference's recurrence
    coef = torch.exp(0.5 * (lf[p] - lf[m] - lf[n])) * (p == m + n).to(rt)                 # [d, d, d]
    pw0 = _int_powers(u[:, 0, 0], d)                                                      # [N, d]
    pw1 = _int_powers(u[:, 1, 0], d)
    t0 = coef[None] * pw0[:, :, None, None] * pw1[:, None, :, None]                       # [N, m, n, p]
    slices = [t0]
    sm = (sq[:, None, None]).expand(d, d, d)
    sn = (sq[None, :, None]).expand(d, d, d)
    for q in range(1, d):
        prev = slices[-1]
        sh_m = torch.cat([torch.zeros_like(prev[:, :1]), prev[:, :-1]], dim=1)            # T[m-1, n, p, q-1]
        sh_n = torch.cat([torch.zeros_like(prev[:, :, :1]), prev[:, :, :-1]], dim=2)      # T[m, n-1, p, q-1]
        cur = (sm[None] / sq[q]) * u[:, 0, 1][:, None, None, None] * sh_m + \
              (sn[None] / sq[q]) * u[:, 1, 1][:, None, None, None] * sh_n
        mask = (m + n - p == q).to(rt)
        slices.append(cur * mask[None])
    return torch.stack(slices, dim=-1)                                                    # [N, m, n, p, q]


class _FockGate(nn.Module):
    """Base of the photonic gates on the Fock tensor path (reference photonic/operation.py:60-271).

    `_structure`: block structure of the class's Fock matrix (include/b200q.h B200Q_QUDIT_*), a property of the gate
    CLASS (photon-number conservation of the beamsplitter family, ...), never read off the values."""
    _structure = L.QUDIT_GENERAL

    def __init__(self, name, nmode, wires, cutoff):
        super().__init__()
        self.name, self.nmode, self.cutoff = name, nmode, cutoff
        self.wires = [wires] if isinstance(wires, int) else list(wires)
        assert all(0 <= w < nmode for w in self.wires) and len(set(self.wires)) == len(self.wires)

    def _to_tensor(self, x):
        return x if isinstance(x, (torch.Tensor, nn.Parameter)) else torch.tensor(x, dtype=torch.float)

    def op_state_tensor(self, x: torch.Tensor) -> torch.Tensor:
        """Out-of-place application to `[batch, cutoff, ..., cutoff]` (photonic/operation.py:142-146)."""
        nt = len(self.wires)
        matrix = self.update_matrix_state().reshape(self.cutoff**nt, self.cutoff**nt)
        flat = x.reshape(-1, self.cutoff**self.nmode).contiguous().clone()
        qudit_apply_(flat, self.nmode, self.cutoff, matrix, self.wires, flat.shape[0], self._structure)
        return flat.reshape(x.shape)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.op_state_tensor(x)

    def _free_names(self) -> list:
        """Names of the free parameters, in the order `inputs` lists them."""
        if hasattr(self, '_free'):
            return [self._free]
        order = [nm for nm in ('r', 'kappa', 'theta', 'phi') if nm in self._buffers or nm in self._parameters]
        return order[:self.npara]

    def init_para(self, inputs: Any = None) -> None:
        """Sets the free parameters from `inputs` (data encoding, reference photonic/gate.py `init_para`)."""
        names = self._free_names()
        if inputs is None:
            return
        vals = inputs if isinstance(inputs, torch.Tensor) else torch.as_tensor(inputs, dtype=torch.float)
        vals = vals.reshape(-1)
        assert vals.numel() == len(names), f'{self.name} takes {len(names)} parameter(s)'
        for nm, v in zip(names, vals):
            if nm in self._parameters:
                with torch.no_grad():
                    self._parameters[nm].copy_(v)
            else:
                self._buffers[nm] = v.to(self._buffers[nm].device)

    def extra_repr(self) -> str:
        return f'wires={self.wires}'


class PhaseShift(_FockGate):
    """diag(exp(i theta n)); `inv_mode` rotates clockwise (theta -> -theta, reference photonic/gate.py:135-199)."""
    _structure = L.QUDIT_DIAG

    def __init__(self, inputs: Any = None, nmode: int = 1, wires=None, cutoff: int = 2, requires_grad: bool = False,
                 inv_mode: bool = False):
        super().__init__('PhaseShift', nmode, [0] if wires is None else wires, cutoff)
        self.inv_mode = inv_mode
        theta = torch.rand(1)[0] * 2 * torch.pi if inputs is None else self._to_tensor(inputs)
        if requires_grad:
            self.theta = nn.Parameter(theta)
        else:
            self.register_buffer('theta', theta)
        self.npara = 1

    def _params(self):
        return [-self.theta if self.inv_mode else self.theta]

    @staticmethod
    def _batched_matrix_state(p, d):
        return ps_matrix_state(p[:, 0], d)

    def update_matrix_state(self) -> torch.Tensor:
        return ps_matrix_state(self._params()[0].reshape(1).double(), self.cutoff)[0]


class BeamSplitter(_FockGate):
    """BS(theta, phi): mode-mixing matrix [[cos, -e^{-i phi} sin], [e^{i phi} sin, cos]] (photonic/gate.py:331-339)."""
    _structure = L.QUDIT_NUMBER

    def __init__(self, inputs: Any = None, nmode: int = 2, wires=None, cutoff: int = 2, requires_grad: bool = False):
        super().__init__('BeamSplitter', nmode, [0, 1] if wires is None else wires, cutoff)
        assert len(self.wires) == 2
        if inputs is None:
            theta, phi = torch.rand(1)[0] * 2 * torch.pi, torch.rand(1)[0] * 2 * torch.pi
        else:
            theta, phi = self._to_tensor(inputs[0]), self._to_tensor(inputs[1])
        for nm, v in (('theta', theta), ('phi', phi)):
            if requires_grad:
                setattr(self, nm, nn.Parameter(v))
            else:
                self.register_buffer(nm, v)
        self.npara = 2

    def _params(self):
        return [self.theta, self.phi]

    @staticmethod
    def mixing_matrix(theta, phi):
        cos, sin = torch.cos(theta) + 0j, torch.sin(theta) + 0j
        return torch.stack([cos, -torch.exp(-1j * phi) * sin, torch.exp(1j * phi) * sin, cos], dim=-1).reshape(
            *theta.shape, 2, 2)

    @staticmethod
    def _batched_matrix_state(p, d):
        return bs_matrix_state(BeamSplitter.mixing_matrix(p[:, 0], p[:, 1]), d)

    def update_matrix_state(self) -> torch.Tensor:
        p = torch.stack([self.theta.reshape(()), self.phi.reshape(())]).double().unsqueeze(0)
        return self._batched_matrix_state(p, self.cutoff)[0]


class Squeezing(_FockGate):
    _structure = L.QUDIT_DENSE1
    def __init__(self, inputs: Any = None, nmode: int = 1, wires=None, cutoff: int = 2, requires_grad: bool = False):
        super().__init__('Squeezing', nmode, [0] if wires is None else wires, cutoff)
        if inputs is None:
            r, theta = torch.rand(1)[0], torch.rand(1)[0] * 2 * torch.pi
        else:
            r, theta = self._to_tensor(inputs[0]), self._to_tensor(inputs[1])
        for nm, v in (('r', r), ('theta', theta)):
            if requires_grad:
                setattr(self, nm, nn.Parameter(v))
            else:
                self.register_buffer(nm, v)
        self.npara = 2

    def _params(self):
        return [self.r, self.theta]

    @staticmethod
    def _batched_matrix_state(p, d):
        return squeezing_matrix_state(p[:, 0], p[:, 1], d)

    def update_matrix_state(self) -> torch.Tensor:
        p = torch.stack([self.r.reshape(()), self.theta.reshape(())]).double().unsqueeze(0)
        return self._batched_matrix_state(p, self.cutoff)[0]


class Displacement(Squeezing):
    """D(r, theta) (reference photonic/gate.py:1336-1489); same parameter handling as the squeezer."""

    def __init__(self, inputs: Any = None, nmode: int = 1, wires=None, cutoff: int = 2, requires_grad: bool = False):
        super().__init__(inputs, nmode, wires, cutoff, requires_grad)
        self.name = 'Displacement'

    @staticmethod
    def _batched_matrix_state(p, d):
        return displacement_matrix_state(p[:, 0], p[:, 1], d)


class Squeezing2(BeamSplitter):
    """Two-mode squeezing S2(r, theta) (reference photonic/gate.py:1157-1333); parameters handled like BS(theta, phi)."""
    _structure = L.QUDIT_DIFFERENCE

    def __init__(self, inputs: Any = None, nmode: int = 2, wires=None, cutoff: int = 2, requires_grad: bool = False):
        if inputs is None:
            inputs = [torch.rand(1)[0], torch.rand(1)[0] * 2 * torch.pi]
        super().__init__(inputs, nmode, wires, cutoff, requires_grad)
        self.name = 'Squeezing2'

    @staticmethod
    def _batched_matrix_state(p, d):
        return squeezing2_matrix_state(p[:, 0], p[:, 1], d)


class MZI(BeamSplitter):
    """Mach-Zehnder interferometer, phase shifter before (`phi_first`) or after the theta stage
    (reference photonic/gate.py:414-516)."""

    def __init__(self, inputs: Any = None, nmode: int = 2, wires=None, cutoff: int = 2, phi_first: bool = True,
                 requires_grad: bool = False):
        super().__init__(inputs, nmode, wires, cutoff, requires_grad)
        self.name = 'MZI'
        self.phi_first = phi_first
        self._variant = bool(phi_first)

    @staticmethod
    def mixing_matrix(theta, phi, phi_first=True):
        cos, sin = torch.cos(theta / 2) + 0j, torch.sin(theta / 2) + 0j
        pre = 1j * torch.exp(1j * theta / 2)
        e_ip = torch.exp(1j * phi)
        mat = (pre[..., None] * torch.stack([e_ip * sin, cos, e_ip * cos, -sin], dim=-1)).reshape(*theta.shape, 2, 2)
        return mat if phi_first else mat.transpose(-1, -2)

    @staticmethod
    def _batched_matrix_state(p, d, phi_first=True):
        return bs_matrix_state(MZI.mixing_matrix(p[:, 0], p[:, 1], phi_first), d)

    def update_matrix_state(self) -> torch.Tensor:
        p = torch.stack([self.theta.reshape(()), self.phi.reshape(())]).double().unsqueeze(0)
        return self._batched_matrix_state(p, self.cutoff, self.phi_first)[0]


class _BeamSplitterOneParam(BeamSplitter):
    """Beamsplitter with one free angle; the other is a float32 constant buffer like in the reference
    (`torch.pi / 2` / `torch.pi / 4` pass through `torch.tensor(..., dtype=torch.float)`)."""
    _free, _fixed_value = 'theta', 0.0

    def __init__(self, inputs: Any = None, nmode: int = 2, wires=None, cutoff: int = 2, requires_grad: bool = False):
        _FockGate.__init__(self, type(self).__name__, nmode, [0, 1] if wires is None else wires, cutoff)
        assert len(self.wires) == 2
        free = torch.rand(1)[0] * 2 * torch.pi if inputs is None else self._to_tensor(inputs)
        while isinstance(free, (list, tuple)):
            free = self._to_tensor(free[0])
        fixed = torch.tensor(self._fixed_value, dtype=torch.float).to(free.device, free.dtype)
        if requires_grad:
            setattr(self, self._free, nn.Parameter(free))
        else:
            self.register_buffer(self._free, free)
        self.register_buffer('phi' if self._free == 'theta' else 'theta', fixed)
        self.npara = 1


class BeamSplitterTheta(_BeamSplitterOneParam):
    """BS(theta, phi = pi/2) (reference photonic/gate.py:519-613)."""
    _free, _fixed_value = 'theta', torch.pi / 2


class BeamSplitterPhi(_BeamSplitterOneParam):
    """BS(theta = pi/4, phi) (reference photonic/gate.py:616-710)."""
    _free, _fixed_value = 'phi', torch.pi / 4


class BeamSplitterSingle(_FockGate):
    """One-angle beamsplitters in the `'rx'`, `'ry'` or `'h'` convention (reference photonic/gate.py:713-877)."""
    _structure = L.QUDIT_NUMBER

    def __init__(self, inputs: Any = None, nmode: int = 2, wires=None, cutoff: int = 2, convention: str = 'rx',
                 requires_grad: bool = False):
        super().__init__('BeamSplitterSingle', nmode, [0, 1] if wires is None else wires, cutoff)
        assert len(self.wires) == 2 and convention in ('rx', 'ry', 'h')
        self.convention = convention
        self._variant = convention
        theta = torch.rand(1)[0] * 2 * torch.pi if inputs is None else self._to_tensor(inputs)
        if requires_grad:
            self.theta = nn.Parameter(theta)
        else:
            self.register_buffer('theta', theta)
        self.npara = 1

    def _params(self):
        return [self.theta]

    @staticmethod
    def mixing_matrix(theta, convention):
        cos, sin = torch.cos(theta / 2) + 0j, torch.sin(theta / 2) + 0j
        entries = {'rx': [cos, 1j * sin, 1j * sin, cos], 'ry': [cos, -sin, sin, cos], 'h': [cos, sin, sin, -cos]}
        return torch.stack(entries[convention], dim=-1).reshape(*theta.shape, 2, 2)

    @staticmethod
    def _batched_matrix_state(p, d, convention='rx'):
        return bs_matrix_state(BeamSplitterSingle.mixing_matrix(p[:, 0], convention), d)

    def update_matrix_state(self) -> torch.Tensor:
        return self._batched_matrix_state(self.theta.reshape(1, 1).double(), self.cutoff, self.convention)[0]

    def extra_repr(self) -> str:
        return f'wires={self.wires}, theta={self.theta.item()}, convention={self.convention}'


class Kerr(_FockGate):
    """diag(exp(i kappa n^2)) (reference photonic/gate.py:2291-2383)."""
    _structure = L.QUDIT_DIAG

    def __init__(self, inputs: Any = None, nmode: int = 1, wires=None, cutoff: int = 2, requires_grad: bool = False):
        super().__init__(type(self).__name__, nmode, self._default_wires() if wires is None else wires, cutoff)
        kappa = torch.rand(1)[0] * 2 * torch.pi if inputs is None else self._to_tensor(inputs)
        if requires_grad:
            self.kappa = nn.Parameter(kappa)
        else:
            self.register_buffer('kappa', kappa)
        self.npara = 1

    @staticmethod
    def _default_wires():
        return [0]

    def _params(self):
        return [self.kappa]

    @staticmethod
    def _batched_matrix_state(p, d):
        n = torch.arange(d, dtype=p.dtype, device=p.device)
        return torch.diag_embed(torch.exp(1j * p[:, :1] * (n * n)[None, :]))

    def update_matrix_state(self) -> torch.Tensor:
        return self._batched_matrix_state(self.kappa.reshape(1, 1).double(), self.cutoff)[0]


class CrossKerr(Kerr):
    """diag(exp(i kappa n1 n2)) on two modes (reference photonic/gate.py:2386-2483)."""

    @staticmethod
    def _default_wires():
        return [0, 1]

    @staticmethod
    def _batched_matrix_state(p, d):
        n = torch.arange(d, dtype=p.dtype, device=p.device)
        return torch.diag_embed(torch.exp(1j * p[:, :1] * torch.kron(n, n)[None, :]))


class FockState(nn.Module):
    """Fock state tensor `[1, cutoff, ..., cutoff]`: `'vac'`, a list of photon numbers, a superposition
    `[(amplitude, [photon numbers]), ...]` (not normalised, like the reference) or a complex tensor
    (photonic/state.py:20-119, `basis=False`)."""

    def __init__(self, state: Any = 'vac', nmode: int | None = None, cutoff: int | None = None):
        super().__init__()
        if isinstance(state, str):
            assert state in ('vac', 'zeros') and nmode is not None and cutoff is not None
            occ = [0] * nmode
        elif isinstance(state, torch.Tensor) and state.is_complex():
            occ = None
            nmode = state.ndim - 1 if nmode is None else nmode
            cutoff = state.shape[-1] if cutoff is None else cutoff
        else:
            assert isinstance(state, (list, tuple))
            terms = [(1.0, list(state))] if all(isinstance(i, int) for i in state) else list(state)
            assert all(isinstance(i, tuple) for i in terms)
            occ = terms
            nmode = len(terms[0][1]) if nmode is None else nmode
            cutoff = max(sum(o) for _, o in terms) + 1 if cutoff is None else cutoff
        self.nmode, self.cutoff = nmode, cutoff
        if occ is None:
            t = state.reshape([-1] + [cutoff] * nmode)
        else:
            t = torch.zeros([1] + [cutoff] * nmode, dtype=torch.cfloat)
            for amp, o in ([(1.0, occ)] if isinstance(occ[0], int) else occ):
                t[(0,) + tuple(int(x) for x in o)] = amp
        self.register_buffer('state', t)

    def _apply(self, fn):
        tensor = self._buffers.pop('state')
        super()._apply(fn)
        self.register_buffer('state', apply_complex_fix(fn, {'state': tensor})['state'])
        return self


class QumodeCircuit(nn.Module):
    """Photonic circuit on the Fock TENSOR path only (`backend='fock'`, `basis=False`, pure state)."""

    def __init__(self, nmode: int, init_state: Any = 'vac', cutoff: int | None = None, backend: str = 'fock',
                 basis: bool = False, den_mat: bool = False, name: str | None = None):
        super().__init__()
        if backend != 'fock' or basis or den_mat:
            raise NotImplementedError('deepquantum_b200 accelerates the Fock tensor path only '
                                      "(backend='fock', basis=False, den_mat=False)")
        self.nmode, self.name, self.backend, self.basis = nmode, name, backend, basis
        if isinstance(init_state, FockState):
            self.init_state = init_state
            cutoff = init_state.cutoff if cutoff is None else cutoff
        else:
            assert cutoff is not None, 'cutoff is required'
            self.init_state = FockState(init_state, nmode=nmode, cutoff=cutoff)
        self.cutoff = cutoff
        self.operators = nn.Sequential()
        self.state = None
        self.npara = 0
        self.ndata = 0
        self.encoders = []

    def add(self, op: _FockGate, encode: bool = False) -> None:
        self.operators.append(op)
        if encode:
            assert not any(p.requires_grad for p in op.parameters()), 'an encoder gate cannot be trainable'
            self.encoders.append(op)
            self.ndata += op.npara
        else:
            self.npara += op.npara

    def encode(self, data: Any) -> None:
        """Routes `data` to the encoder gates in the order they were added (reference photonic/circuit.py:`encode`)."""
        if data is None:
            return
        data = data if isinstance(data, torch.Tensor) else torch.as_tensor(data, dtype=torch.float)
        assert data.ndim == 1, 'batched data is outside the accelerated Fock tensor path'
        assert data.numel() >= self.ndata, 'The circuit needs more data'
        count = 0
        for op in self.encoders:
            op.init_para(data[count:count + op.npara])
            count += op.npara

    def ps(self, wires: int, inputs: Any = None, encode: bool = False) -> None:
        self.add(PhaseShift(inputs, self.nmode, wires, self.cutoff, requires_grad=inputs is None and not encode), encode=encode)

    def bs(self, wires: list[int], inputs: Any = None, encode: bool = False) -> None:
        self.add(BeamSplitter(inputs, self.nmode, wires, self.cutoff, requires_grad=inputs is None and not encode), encode=encode)

    def s(self, wires: int, r: Any = None, theta: Any = None, encode: bool = False) -> None:
        inputs = None if r is None else [r, 0.0 if theta is None else theta]
        self.add(Squeezing(inputs, self.nmode, wires, self.cutoff, requires_grad=inputs is None and not encode), encode=encode)

    # ---- the beamsplitter family, rotations and Kerr gates (reference photonic/circuit.py:2026-2245, 2471-2520,
    # 2628-2680); `mu` / `sigma` (the reference's gate-noise model) are accepted and must stay unset -------------
    def _one(self, cls, wires, inputs, encode, mu, sigma, **kw):
        assert mu is None and sigma is None, 'gate noise is outside the accelerated Fock tensor path'
        self.add(cls(inputs, self.nmode, wires, self.cutoff, requires_grad=inputs is None and not encode, **kw), encode=encode)

    def mzi(self, wires, inputs=None, phi_first=True, encode=False, mu=None, sigma=None):
        self._one(MZI, wires, inputs, encode, mu, sigma, phi_first=phi_first)

    def bs_theta(self, wires, inputs=None, encode=False, mu=None, sigma=None):
        self._one(BeamSplitterTheta, wires, inputs, encode, mu, sigma)

    def bs_phi(self, wires, inputs=None, encode=False, mu=None, sigma=None):
        self._one(BeamSplitterPhi, wires, inputs, encode, mu, sigma)

    def bs_rx(self, wires, inputs=None, encode=False, mu=None, sigma=None):
        self._one(BeamSplitterSingle, wires, inputs, encode, mu, sigma, convention='rx')

    def bs_ry(self, wires, inputs=None, encode=False, mu=None, sigma=None):
        self._one(BeamSplitterSingle, wires, inputs, encode, mu, sigma, convention='ry')

    def bs_h(self, wires, inputs=None, encode=False, mu=None, sigma=None):
        self._one(BeamSplitterSingle, wires, inputs, encode, mu, sigma, convention='h')

    def dc(self, wires, mu=None, sigma=None):
        """Directional coupler: `bs_rx(pi / 2)`."""
        self._one(BeamSplitterSingle, wires, torch.pi / 2, False, mu, sigma, convention='rx')

    def h(self, wires, mu=None, sigma=None):
        """Photonic Hadamard: `bs_h(pi / 2)`."""
        self._one(BeamSplitterSingle, wires, torch.pi / 2, False, mu, sigma, convention='h')

    def r(self, wires, inputs=None, encode=False, inv_mode=False, mu=None, sigma=None):
        self._one(PhaseShift, wires, inputs, encode, mu, sigma, inv_mode=inv_mode)

    def f(self, wires, mu=None, sigma=None):
        """Fourier gate: rotation by pi / 2."""
        self._one(PhaseShift, wires, torch.pi / 2, False, mu, sigma)

    def k(self, wires, inputs=None, encode=False, mu=None, sigma=None):
        self._one(Kerr, wires, inputs, encode, mu, sigma)

    def ck(self, wires, inputs=None, encode=False, mu=None, sigma=None):
        self._one(CrossKerr, wires, inputs, encode, mu, sigma)

    def d(self, wires, r=None, theta=None, encode=False, mu=None, sigma=None):
        assert mu is None and sigma is None, 'gate noise is outside the accelerated Fock tensor path'
        if r is None and theta is None:
            inputs = None
        else:
            inputs = [torch.rand(1)[0] if r is None else r, 0.0 if theta is None else theta]
        self.add(Displacement(inputs, self.nmode, wires, self.cutoff, requires_grad=inputs is None and not encode), encode=encode)

    def s2(self, wires, r=None, theta=None, encode=False, mu=None, sigma=None):
        assert mu is None and sigma is None, 'gate noise is outside the accelerated Fock tensor path'
        if r is None and theta is None:
            inputs = None
        else:
            inputs = [torch.rand(1)[0] if r is None else r, 0.0 if theta is None else theta]
        self.add(Squeezing2(inputs, self.nmode, wires, self.cutoff, requires_grad=inputs is None and not encode), encode=encode)

    def build_matrices(self, cdtype, device):
        """All Fock transformation matrices of the circuit: one batched evaluation per gate class (and variant)."""
        groups = {}
        for i, op in enumerate(self.operators):
            groups.setdefault((type(op), getattr(op, '_variant', None)), []).append(i)
        mats = [None] * len(self.operators)
        for (cls, variant), ids in groups.items():
            p = torch.stack([torch.stack([t.reshape(()) for t in self.operators[i]._params()]) for i in ids])
            p = p.to(device=device, dtype=torch.float64)
            m = cls._batched_matrix_state(p, self.cutoff) if variant is None else \
                cls._batched_matrix_state(p, self.cutoff, variant)
            nt = len(self.operators[ids[0]].wires)
            m = m.reshape(len(ids), self.cutoff**nt, self.cutoff**nt).to(cdtype)
            for j, i in enumerate(ids):
                mats[i] = m[j]
        return mats

    def forward(self, data: Any = None, state: Any = None) -> torch.Tensor:
        """Final Fock state tensor `[batch, cutoff, ..., cutoff]` (photonic/circuit.py:405-431: `forward(data,
        state)`; 1-D `data` feeds the gates added with `encode=True`)."""
        if isinstance(data, FockState) and state is None:   # forward(state) of the earlier signature
            data, state = None, data
        self.encode(data)
        x = self.init_state.state if state is None else (state.state if isinstance(state, FockState) else state)
        engine.require_cuda(x, 'the Fock state (move the circuit with cir.to("cuda"))')
        d, n = self.cutoff, self.nmode
        flat = x.reshape(-1, d**n).contiguous().clone()
        with torch.no_grad():
            mats = self.build_matrices(flat.dtype, flat.device)
            fuse = (FUSE_FOCK and len(self.operators) > 0 and d * d <= 256 and d * d <= FOCK_TILE_MAX and n >= 2
                    and all(1 <= len(op.wires) <= 2 for op in self.operators))
            if fuse:
                key = tuple(tuple(op.wires) for op in self.operators)
                if self.__dict__.get('_fock_plan_key') != key:
                    self.__dict__['_fock_plan'] = plan_fock_passes([tuple(op.wires) for op in self.operators], n, d)
                    self.__dict__['_fock_plan_key'] = key
                buf = torch.cat([m.reshape(-1) for m in mats]).contiguous()
                offs, acc = [], 0
                for m in mats:
                    offs.append(acc)
                    acc += m.numel()
                for tile, ids in self.__dict__['_fock_plan']:
                    qudit_fused_(flat, n, d, tile, [(self.operators[i].wires, offs[i]) for i in ids], buf, flat.shape[0])
            else:
                ops = list(self.operators)
                key = tuple((tuple(op.wires), op._structure) for op in ops)
                if self.__dict__.get('_group_key') != key:
                    self.__dict__['_groups'] = plan_fock_groups([(list(w), st) for w, st in key], n, d)
                    self.__dict__['_group_key'] = key
                for grp in self.__dict__['_groups']:
                    if len(grp) == 1:
                        op = ops[grp[0]]
                        qudit_apply_(flat, n, d, mats[grp[0]], op.wires, flat.shape[0], op._structure)
                    else:
                        pair = next(ops[i].wires for i in grp if len(ops[i].wires) == 2)
                        qudit_apply_group_(flat, n, d, pair, [(mats[i], ops[i].wires, ops[i]._structure) for i in grp],
                                           flat.shape[0])
        self.state = flat.reshape([-1] + [d] * n)
        return self.state

    def fock_plan_stats(self) -> dict:
        """Passes and gates per pass of the fused plan of the last forward (diagnostics / bench)."""
        plan = self.__dict__.get('_fock_plan') or []
        if not plan and self.__dict__.get('_groups'):
            grp = self.__dict__['_groups']
            return {'gates': len(self.operators), 'passes': len(grp), 'gates_per_pass': [len(x) for x in grp]}
        return {'gates': len(self.operators), 'passes': len(plan), 'gates_per_pass': [len(ids) for _, ids in plan],
                'tiles': [t for t, _ in plan]}
