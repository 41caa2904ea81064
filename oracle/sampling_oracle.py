"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the sampling step that follows the gate-application
path: `qmath.measure` / `block_sample` (reference qmath.py:543-638).

The reference draws with torch.multinomial (its random stream cannot be reproduced by another
implementation), so the oracle states the DISTRIBUTION the reference samples from -- |a|^2 on the full index,
marginalised over the un-measured wires with the measured wires in ascending order as the key bits
(qmath.py:609-624) -- and an inverse-CDF sampler driven by explicit uniforms, which is what the CUDA kernels
implement (csrc/b200q_sample.cu).  Pinned against reference outputs in tests/golden/measure.npz
(`with_prob=True` probabilities returned by the unmodified reference).
"""
from __future__ import annotations

import numpy as np


def probabilities(state: np.ndarray, nqubit: int, wires=None) -> np.ndarray:
    """|a|^2, marginalised onto `wires` (sorted ascending like qmath.py:611; wire 0 = most significant bit)."""
    p = np.abs(np.asarray(state, dtype=np.complex128).reshape(-1))**2
    if wires is None:
        return p
    wires = sorted([wires] if isinstance(wires, int) else list(wires))
    rest = [i for i in range(nqubit) if i not in wires]
    p = p.reshape([2] * nqubit).transpose(wires + rest).reshape(2**len(wires), -1).sum(-1)   # qmath.py:623-624
    return p


def sample_indices(state: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """Inverse CDF on the full index: first i with sum_{j<=i} |a_j|^2 > u * total."""
    p = np.abs(np.asarray(state, dtype=np.complex128).reshape(-1))**2
    cdf = np.cumsum(p)
    idx = np.searchsorted(cdf, np.asarray(uniforms, dtype=np.float64) * cdf[-1], side='right')
    return np.minimum(idx, len(p) - 1)


def keys_of(indices: np.ndarray, nqubit: int, wires=None) -> np.ndarray:
    """Measured bits of a full index as the integer the reference formats with bin() (qmath.py:628)."""
    if wires is None:
        return np.asarray(indices)
    wires = sorted([wires] if isinstance(wires, int) else list(wires))
    out = np.zeros_like(np.asarray(indices))
    for j, w in enumerate(wires):
        out |= ((np.asarray(indices) >> (nqubit - 1 - w)) & 1) << (len(wires) - 1 - j)
    return out


def measure(state: np.ndarray, nqubit: int, uniforms: np.ndarray, wires=None, with_prob: bool = False) -> dict:
    """Counter of bit strings (qmath.py:626-632) for the given uniforms."""
    keys = keys_of(sample_indices(state, uniforms), nqubit, wires)
    nbits = nqubit if wires is None else len([wires] if isinstance(wires, int) else wires)
    probs = probabilities(state, nqubit, wires)
    out = {}
    for k in keys.tolist():
        s = format(k, f'0{nbits}b')
        out[s] = out.get(s, 0) + 1
    if with_prob:
        out = {s: (c, float(probs[int(s, 2)])) for s, c in out.items()}
    return out
