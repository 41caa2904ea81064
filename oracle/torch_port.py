"""CPU restatement of the reference's gate-application loop on torch CPU tensors, used ONLY as the
timed CPU baseline (`bench.py` cpu_baseline / --impl reference) and as a second checker.

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.

It follows the reference's data movement step by step, because that is what its cost is made of:
`evolve_state` (qmath.py:485-506) moves the target wires to the front, materialises the permuted
state as a (2^k, 2^(n-k)) matrix (the `aten::copy_` half of the profile), multiplies by the gate
matrix with ATen's complex `mm` (the other half), and returns a permuted VIEW, so the next gate pays
the copy again; `op_state_control` (operation.py:203-219) multiplies only the all-ones control slice
and re-assembles the state with `torch.cat`.  The matmul runs on every host thread torch is given.
"""
from __future__ import annotations

import torch


def evolve_state(state: torch.Tensor, matrix: torch.Tensor, nqubit: int, wires) -> torch.Tensor:
    """state: [batch, 2, ..., 2] (any strides) -> same shape, as a permuted view (qmath.py:497-506)."""
    k = len(wires)
    front = [w + 1 for w in wires]
    order = front + [a for a in range(nqubit + 1) if a not in front]
    mat2d = state.permute(order).reshape(2**k, -1)          # contiguous copy
    out = (matrix @ mat2d).reshape([2] * k + [-1] + [2] * (nqubit - k))
    back = [0] * (nqubit + 1)
    for pos, axis in enumerate(order):
        back[axis] = pos
    return out.permute(back)


def evolve_state_controlled(state, matrix, nqubit, wires, controls):
    """operation.py:203-219: permute to [targets, rest, controls], act on the last control slice, cat."""
    k, c = len(wires), len(controls)
    front = [w + 1 for w in wires]
    tail = [w + 1 for w in controls]
    order = front + [a for a in range(nqubit + 1) if a not in front and a not in tail] + tail
    x = state.permute(order).reshape(2**k, -1, 2**c)
    x = torch.cat([x[:, :, :-1], (matrix @ x[:, :, -1]).unsqueeze(-1)], dim=-1)
    x = x.reshape([2] * k + [-1] + [2] * (nqubit - k - c) + [2] * c)
    back = [0] * (nqubit + 1)
    for pos, axis in enumerate(order):
        back[axis] = pos
    return x.permute(back)


def run_ops(ops, nqubit: int, state: torch.Tensor | None = None, dtype=torch.complex64, max_gates=None,
            time_budget_s: float | None = None):
    """Apply lowered ops `(matrix, wires, controls)` like the nn.Sequential loop (circuit.py:261).
    Returns (state [2^n], gates_applied, seconds)."""
    import time
    if state is None:
        state = torch.zeros(2**nqubit, dtype=dtype)
        state[0] = 1
    x = state.reshape([1] + [2] * nqubit)
    mats = [torch.as_tensor(m).to(dtype) for m, _, _ in ops]
    done = 0
    t0 = time.perf_counter()
    with torch.no_grad():
        for (m, wires, controls), mt in zip(ops, mats):
            if max_gates is not None and done >= max_gates:
                break
            if time_budget_s is not None and done >= 2 and time.perf_counter() - t0 > time_budget_s:
                break
            x = evolve_state_controlled(x, mt, nqubit, wires, controls) if controls else evolve_state(x, mt, nqubit,
                                                                                                     wires)
            done += 1
        out = x.reshape(-1)   # vector_rep: the final contiguous copy (operation.py:57-59)
    return out, done, time.perf_counter() - t0
