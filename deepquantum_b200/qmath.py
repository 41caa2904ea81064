"""Boundary functions with the reference's names (qmath.py): `evolve_state`, `expectation`,
`inverse_permutation`, `multi_kron`.  The contraction itself runs in libb200q.so."""
from __future__ import annotations

import torch

from . import _lib as L
from . import engine
from .state import amplitude_encoding  # noqa: F401  (re-exported like the reference)


def inverse_permutation(permute_shape: list[int]) -> list[int]:
    inv = [0] * len(permute_shape)
    for pos, axis in enumerate(permute_shape):
        inv[axis] = pos
    return inv


def multi_kron(lst: list[torch.Tensor]) -> torch.Tensor:
    out = lst[0]
    for m in lst[1:]:
        out = torch.kron(out, m)
    return out


def evolve_state(state: torch.Tensor, matrix: torch.Tensor, nqudit: int, wires: list[int],
                 qudit: int = 2) -> torch.Tensor:
    """Drop-in for `qmath.evolve_state` (reference qmath.py:485-506).

    `state` is `[batch, d, ..., d]` (any strides), `matrix` is `d^k x d^k` with `wires[0]` the most
    significant matrix digit.  Returns a new tensor of the same shape; the input is not modified."""
    shape = state.shape
    flat = state.reshape(-1, qudit**nqudit).contiguous().clone()
    batch = flat.shape[0]
    if qudit == 2:
        engine.apply_gate_(flat, nqudit, matrix, engine.wires_to_targets(nqudit, wires), (), L.GATE_MAT, False, batch)
    else:
        from .photonic import qudit_apply_
        qudit_apply_(flat, nqudit, qudit, matrix, wires, batch)
    return flat.reshape(shape)


def evolve_state_controlled(state: torch.Tensor, matrix: torch.Tensor, nqubit: int, wires: list[int],
                            controls: list[int]) -> torch.Tensor:
    """Drop-in for `Gate.op_state_control` (reference operation.py:203-219) as a free function."""
    shape = state.shape
    flat = state.reshape(-1, 2**nqubit).contiguous().clone()
    engine.apply_gate_(flat, nqubit, matrix, engine.wires_to_targets(nqubit, wires),
                       [nqubit - 1 - c for c in controls], L.GATE_MAT, False, flat.shape[0])
    return flat.reshape(shape)


def sample2expval(sample: dict) -> torch.Tensor:
    """Average parity of the sampled bit strings (reference qmath.py:863-871)."""
    total = exp = 0
    for bitstring, ncount in sample.items():
        exp += ncount * (-1) ** (bitstring.count('1') % 2)
        total += ncount
    return torch.tensor([exp / total])


def measure(state: torch.Tensor, shots: int = 1024, with_prob: bool = False, wires=None, den_mat: bool = False,
            block_size: int = 2**24, generator: torch.Generator | None = None):
    """Drop-in for `qmath.measure` (reference qmath.py:568-638) on device statevectors.

    Same return format (bit string of the measured wires in ascending wire order -> count, or (count, prob)
    with `with_prob`).  The state is read once for the block masses and every shot then scans one 32 KiB
    block (`b200q_block_mass`, `b200q_sample_blocks`); marginals on a wire subset are obtained by sampling the
    full index and keeping the measured bits, their exact probabilities by `b200q_marginal_probs`.
    `block_size` is accepted for signature compatibility (the block size here is fixed by the kernel);
    `generator` optionally seeds the uniforms (CPU generator)."""
    if den_mat:
        # probabilities are |diag(rho)| (reference qmath.py:600-602, 620): sample the amplitude vector sqrt(|rho_ii|)
        # with the same kernels -- 2^n of the 4^n entries, a strided view
        assert state.ndim in (2, 3) and state.shape[-1] == state.shape[-2], 'Please input density matrices'
        state = torch.sqrt(state.diagonal(dim1=-2, dim2=-1).abs()).to(state.dtype)
        if state.ndim == 1:
            state = state.unsqueeze(-1)
    is_single = state.ndim == 1 or (state.ndim == 2 and state.shape[-1] == 1)
    batch = 1 if is_single else state.shape[0]
    flat = state.reshape(batch, -1)
    dim = flat.shape[-1]
    assert dim & (dim - 1) == 0, 'The length of the quantum state is not in the form of 2^n'
    n = dim.bit_length() - 1
    if wires is not None:
        if isinstance(wires, int):
            wires = [wires]
        assert isinstance(wires, list)
        wires = sorted(wires)
    meas = list(range(n)) if wires is None else wires
    nbits = len(meas)
    flat = flat.contiguous()
    results = []
    for b in range(batch):
        st = flat[b]
        u = torch.rand(shots, dtype=torch.float64, generator=generator)
        idx = engine.sample_indices(st, n, u)
        if nbits == n:
            keys = idx
        else:
            keys = torch.zeros_like(idx)
            for j, w in enumerate(meas):
                keys |= ((idx >> (n - 1 - w)) & 1) << (nbits - 1 - j)
        vals, counts = torch.unique(keys, return_counts=True)
        probs = None
        if with_prob:
            if nbits == n:
                a = st[vals]
                probs = (a.real.double()**2 + a.imag.double()**2)
            else:
                mask = 0
                for w in meas:
                    mask |= 1 << (n - 1 - w)
                dep = torch.zeros_like(vals)   # key bits deposited at the measured index bits
                for j, w in enumerate(meas):
                    dep |= ((vals >> (nbits - 1 - j)) & 1) << (n - 1 - w)
                order = torch.argsort(dep)
                p_sorted = engine.marginal_probs(st, n, mask, dep[order].contiguous())
                probs = torch.empty_like(p_sorted)
                probs[order] = p_sorted
            probs = probs.to(st.real.dtype)
        d = {}
        vl, cl = vals.tolist(), counts.tolist()
        for i, (v, c) in enumerate(zip(vl, cl)):
            key = format(v, f'0{nbits}b')
            d[key] = (c, probs[i]) if with_prob else c
        results.append(d)
    return results[0] if batch == 1 else results


def expectation(state: torch.Tensor, observable, den_mat: bool = False, chi: int | None = None) -> torch.Tensor:
    """Drop-in for `qmath.expectation` (reference qmath.py:830-860): <psi| O |psi> (or Tr(O rho)) of one `Observable`
    on a device state `[2^n, 1]` / `[batch, 2^n, 1]` (density matrices `[.., 2^n, 2^n]`); a real scalar, or `[batch]`.
    Z strings are one fused reduction pass over the state, X / Y factors are rotated to Z on a copy.  Differentiable
    w.r.t. the state like the reference's.  Matrix-product states are out of scope (statevector path only)."""
    if isinstance(state, list):
        raise NotImplementedError('matrix product states are outside the statevector path of deepquantum_b200')
    from .circuit import QubitCircuit
    n = observable.nqubit
    cir = QubitCircuit(n, den_mat=den_mat)
    cir.observables.append(observable)
    cir.state = state
    return cir.expectation()[..., 0]
