"""Thin Python layer over the C ABI (include/b200q.h): plans, single-gate application, reductions.

PyTorch is used for device memory and streams only; every compute call goes to libb200q.so on the
current CUDA stream.  There is no CPU fallback: CPU tensors are rejected with B200QError.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import B200QError

_checked_devices = set()


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.complex64:
        return L.C64
    if dtype == torch.complex128:
        return L.C128
    raise B200QError(f'unsupported state dtype {dtype}; expected complex64 or complex128')


def require_cuda(t: torch.Tensor, what: str = 'state') -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise B200QError(f'{what} must be a CUDA tensor: deepquantum_b200 runs its gate kernels on sm_100a only and '
                         'has no CPU fallback (use the reference package for CPU simulation)')
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _checked_devices:
        L.check(L.load().b200q_device_check(dev))
        _checked_devices.add(dev)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def wires_to_targets(nqubit: int, wires) -> list[int]:
    """Reference wires (wires[0] = most significant matrix bit, qmath.py:497-504) -> bit position of each
    matrix-index bit, least significant first."""
    return [nqubit - 1 - w for w in reversed(list(wires))]


class FusedPlan:
    """A fused pass schedule for a fixed gate *structure* (kinds, targets, controls); the matrix values are
    supplied at run time from a device buffer."""

    def __init__(self, nqubit: int, dtype: torch.dtype, gates: list, chunk_bits: int = 0, low_bits: int = 0,
                 max_rounds: int = 0, fuse: bool = True, structured: bool = True, coalesce_bits: int | None = None):
        lib = L.load()
        self.nqubit = nqubit
        self.dtype = dtype
        self.ngates = len(gates)
        arr = (L.GateStruct * max(1, len(gates)))(*gates)
        opt = L.PlanOptions()
        opt.chunk_bits, opt.low_bits, opt.max_rounds, opt.fuse = chunk_bits, low_bits, max_rounds, int(fuse)
        opt.reserved[0] = 0 if structured else 1      # A/B switch: general op codes only (non-lean kernel)
        if coalesce_bits is not None:
            opt.reserved[1] = 1 + coalesce_bits       # lane-owned chunk bits of the rounds that touch global memory
        handle = C.c_void_p()
        L.check(lib.b200q_plan_create(nqubit, dtype_code(dtype), arr, len(gates), C.byref(opt), C.byref(handle)))
        self._h = handle
        st = L.PlanStats()
        L.check(lib.b200q_plan_get_stats(self._h, C.byref(st)))
        self.stats = {n: getattr(st, n) for n, _ in L.PlanStats._fields_}

    @property
    def n_passes(self) -> int:
        return self.stats['n_passes']

    def pass_gates(self, i: int) -> int:
        return L.load().b200q_plan_pass_gates(self._h, i)

    def pass_gate_ids(self, i: int) -> list:
        """Indices of the gates (of the list the plan was created from) that pass `i` applies."""
        buf = (C.c_int32 * 64)()
        n = L.load().b200q_plan_pass_gate_ids(self._h, i, buf, 64)
        if n < 0:
            L.check(n)
        return list(buf[:min(n, 64)])

    def export(self) -> bytes:
        need = C.c_size_t()
        lib = L.load()
        L.check(lib.b200q_plan_export(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        L.check(lib.b200q_plan_export(self._h, buf, need.value, C.byref(need)))
        return buf.raw

    def compile(self, threads: int = 0, exchange_variant: bool = False) -> int:
        """Compile the specialised kernel of every pass now (NVRTC, sm_100a; needs no GPU).  Returns the number
        of passes that have one.  `run` does this lazily for states of >= 20 qubits."""
        rc = L.load().b200q_plan_compile(self._h, threads, int(exchange_variant))
        if rc < 0:
            L.check(rc)
        return rc

    def jit_status(self) -> dict:
        """{'specialised': passes that run a run-time compiled kernel, 'failed': passes whose compilation failed}."""
        ok, bad = C.c_int32(), C.c_int32()
        L.check(L.load().b200q_plan_jit_status(self._h, C.byref(ok), C.byref(bad)))
        return {'specialised': ok.value, 'failed': bad.value}

    def codegen(self, i: int, remote: bool = False) -> str:
        """Generated CUDA source of pass i (diagnostics)."""
        lib = L.load()
        n, smem = C.c_size_t(), C.c_size_t()
        L.check(lib.b200q_plan_codegen(self._h, i, int(remote), None, 0, C.byref(n), C.byref(smem)))
        buf = C.create_string_buffer(n.value)
        L.check(lib.b200q_plan_codegen(self._h, i, int(remote), buf, n.value, C.byref(n), C.byref(smem)))
        return buf.value.decode()

    def run(self, state: torch.Tensor, mats: torch.Tensor | None, batch: int = 1, mat_batch_stride: int = 0,
            first: int | None = None, last: int | None = None) -> None:
        """Apply the plan IN PLACE to `state` (contiguous, batch * 2^n complex elements)."""
        require_cuda(state)
        if state.dtype != self.dtype or not state.is_contiguous() or state.numel() != batch << self.nqubit:
            raise B200QError('state must be a contiguous tensor of batch * 2^n elements of the plan dtype')
        mptr = None
        if mats is not None:
            if mats.dtype != self.dtype or not mats.is_contiguous() or mats.device != state.device:
                raise B200QError('matrix buffer must be contiguous, of the plan dtype, on the state device')
            mptr = mats.data_ptr()
        lib = L.load()
        if first is None and last is None:
            rc = lib.b200q_plan_run(self._h, state.data_ptr(), mptr, batch, mat_batch_stride, _stream(state))
        else:
            rc = lib.b200q_plan_run_range(self._h, first or 0, self.n_passes if last is None else last,
                                          state.data_ptr(), mptr, batch, mat_batch_stride, _stream(state))
        L.check(rc)

    def run_exchange(self, state: torch.Tensor, mats: torch.Tensor | None, peer_ptrs, rank: int, perm=None) -> None:
        """Run the plan on `state` with the last pass storing every chunk into the receive buffer of the rank that
        owns it after the block transpose (`peer_ptrs[r]`: device pointer of rank r's buffer, valid on this
        device).  `state` is left in an intermediate layout; the caller swaps shard and buffer afterwards."""
        require_cuda(state)
        if state.dtype != self.dtype or not state.is_contiguous() or state.numel() != 1 << self.nqubit:
            raise B200QError('state must be a contiguous tensor of 2^n elements of the plan dtype')
        arr = (C.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
        mptr = mats.data_ptr() if mats is not None else None
        pm = None if perm is None else (C.c_uint8 * len(perm))(*[int(x) for x in perm])
        L.check(L.load().b200q_plan_run_exchange(self._h, state.data_ptr(), mptr, arr, len(peer_ptrs), rank, pm,
                                                 _stream(state)))

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                L.load().b200q_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def apply_gate_(state: torch.Tensor, nqubit: int, matrix: torch.Tensor | None, targets, controls=(), kind=L.GATE_MAT,
                adjoint: bool = False, batch: int = 1, mat_batch_stride: int = 0) -> None:
    """In-place single gate: the kernel behind `evolve_state` / `Gate.op_state_control`."""
    require_cuda(state)
    if not state.is_contiguous() or state.numel() != batch << nqubit:
        raise B200QError('state must be contiguous with batch * 2^n elements')
    t = (C.c_int32 * len(targets))(*[int(x) for x in targets])
    c = (C.c_int32 * max(1, len(controls)))(*[int(x) for x in controls]) if len(controls) else None
    mptr = None
    if matrix is not None:
        if matrix.dtype != state.dtype:
            matrix = matrix.to(state.dtype)
        matrix = matrix.contiguous()
        mptr = matrix.data_ptr()
    L.check(L.load().b200q_apply_gate(state.data_ptr(), nqubit, dtype_code(state.dtype), kind, mptr, t, len(targets), c,
                                      len(controls), int(adjoint), batch, mat_batch_stride, _stream(state)))


def norm2(state: torch.Tensor, nqubit: int, batch: int = 1) -> torch.Tensor:
    require_cuda(state)
    out = torch.empty(batch, dtype=torch.float64, device=state.device)
    L.check(L.load().b200q_norm2(state.data_ptr(), nqubit, dtype_code(state.dtype), batch, out.data_ptr(),
                                 _stream(state)))
    return out


SAMPLE_BLOCK_BITS = 12   # 4 096 amplitudes (32 KiB of complex64) per block of the two-level inverse CDF


def block_mass(state: torch.Tensor, nqubit: int, batch: int = 1, block_bits: int | None = None) -> torch.Tensor:
    """[batch, 2^(n - block_bits)] float64: sum of |a|^2 over consecutive blocks (one read of the state)."""
    require_cuda(state)
    bb = min(SAMPLE_BLOCK_BITS if block_bits is None else block_bits, nqubit)
    out = torch.empty(batch, 1 << (nqubit - bb), dtype=torch.float64, device=state.device)
    L.check(L.load().b200q_block_mass(state.data_ptr(), nqubit, dtype_code(state.dtype), batch, bb, out.data_ptr(),
                                      _stream(state)))
    return out


def sample_indices(state: torch.Tensor, nqubit: int, uniforms: torch.Tensor, block_bits: int | None = None,
                   mass: torch.Tensor | None = None) -> torch.Tensor:
    """Inverse-CDF sampling of basis-state indices of ONE state (2^n amplitudes) from float64 uniforms in
    [0, 1): index = first i with sum_{j<=i} |a_j|^2 > u * sum_j |a_j|^2 (reference: torch.multinomial in
    qmath.block_sample, qmath.py:543-565; same distribution, explicit randomness)."""
    require_cuda(state)
    bb = min(SAMPLE_BLOCK_BITS if block_bits is None else block_bits, nqubit)
    if mass is None:
        mass = block_mass(state, nqubit, 1, bb)[0]
    cdf = torch.cumsum(mass, 0)
    u = uniforms.to(device=state.device, dtype=torch.float64).contiguous() * cdf[-1]
    blk = torch.searchsorted(cdf, u, right=True).clamp_(max=cdf.numel() - 1)
    residual = (u - (cdf[blk] - mass[blk])).contiguous()
    out = torch.empty(u.numel(), dtype=torch.int64, device=state.device)
    L.check(L.load().b200q_sample_blocks(state.data_ptr(), nqubit, dtype_code(state.dtype), bb, blk.data_ptr(),
                                         residual.data_ptr(), u.numel(), out.data_ptr(), _stream(state)))
    return out


def marginal_probs(state: torch.Tensor, nqubit: int, mask: int, keys_sorted: torch.Tensor) -> torch.Tensor:
    """float64 [n_keys]: sum of |a_i|^2 over the indices with (i & mask) == key, keys sorted ascending."""
    require_cuda(state)
    out = torch.empty(keys_sorted.numel(), dtype=torch.float64, device=state.device)
    for k0 in range(0, keys_sorted.numel(), 2048):
        part = keys_sorted[k0:k0 + 2048].contiguous()
        L.check(L.load().b200q_marginal_probs(state.data_ptr(), nqubit, dtype_code(state.dtype), mask, part.data_ptr(),
                                              part.numel(), out[k0:].data_ptr(), _stream(state)))
    return out


def inner_product(bra: torch.Tensor, ket: torch.Tensor, nqubit: int, batch: int = 1) -> torch.Tensor:
    require_cuda(bra)
    require_cuda(ket)
    out = torch.empty(batch, 2, dtype=torch.float64, device=bra.device)
    L.check(L.load().b200q_inner_product(bra.data_ptr(), ket.data_ptr(), nqubit, dtype_code(bra.dtype), batch,
                                         out.data_ptr(), _stream(bra)))
    return torch.view_as_complex(out)


def expectation_z(state: torch.Tensor, nqubit: int, masks: torch.Tensor, batch: int = 1,
                  index_offset: int = 0) -> torch.Tensor:
    """[batch, n_masks] float64: sum_i |psi_i|^2 (-1)^popcount(i & mask)."""
    require_cuda(state)
    out = torch.empty(batch, masks.numel(), dtype=torch.float64, device=state.device)
    L.check(L.load().b200q_expectation_z(state.data_ptr(), nqubit, dtype_code(state.dtype), batch, masks.data_ptr(),
                                         masks.numel(), index_offset, out.data_ptr(), _stream(state)))
    return out


def apply_z_weights(state: torch.Tensor, nqubit: int, masks: torch.Tensor, weights: torch.Tensor, batch: int = 1,
                    index_offset: int = 0) -> torch.Tensor:
    require_cuda(state)
    out = torch.empty_like(state)
    weights = weights.to(torch.float64).contiguous()
    L.check(L.load().b200q_apply_z_weights(state.data_ptr(), out.data_ptr(), nqubit, dtype_code(state.dtype), batch,
                                           masks.data_ptr(), weights.data_ptr(), masks.numel(), index_offset,
                                           _stream(state)))
    return out


def init_basis_(state: torch.Tensor, nqubit: int, batch: int = 1, index: int = 0) -> None:
    require_cuda(state)
    L.check(L.load().b200q_init_basis(state.data_ptr(), nqubit, dtype_code(state.dtype), batch, index, _stream(state)))
