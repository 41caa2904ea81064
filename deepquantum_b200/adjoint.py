"""Autograd through the fused engine by the adjoint (reversible) method (reference adjoint.py:19-83).

The reference's single-device circuits rely on PyTorch autograd through permute/mm/cat, which saves
one full state per gate -- impossible at 30 qubits.  Here the backward
pass un-applies the gates from the final state while propagating the cotangent, so only three
states are alive (psi, lambda, and the saved output).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib as L
from . import engine


ADJOINT_CHUNK_BITS = int(os.environ.get('B200Q_ADJOINT_CHUNK_BITS', '0'))   # 0: the forward plan's tiles


class ExpectationZFunction(torch.autograd.Function):
    """[batch, n_masks] = sum_i |psi_i|^2 (-1)^popcount(i & mask)  with its exact cotangent."""

    @staticmethod
    def forward(ctx, state, nqubit, masks, batch):
        ctx.save_for_backward(state, masks)
        ctx.nqubit, ctx.batch = nqubit, batch
        return engine.expectation_z(state, nqubit, masks, batch)

    @staticmethod
    def backward(ctx, grad):
        state, masks = ctx.saved_tensors
        # d/d(conj psi_i) of sum_k g_k sum_i |psi_i|^2 s_k(i), in PyTorch's convention (dL/dRe + i dL/dIm)
        lam = engine.apply_z_weights(state.detach().contiguous(), ctx.nqubit, masks, 2.0 * grad.contiguous(),
                                     ctx.batch)
        return lam.reshape(state.shape), None, None, None


def expectation_z(state: torch.Tensor, nqubit: int, masks: torch.Tensor, batch: int) -> torch.Tensor:
    state = state.contiguous()
    if torch.is_grad_enabled() and state.requires_grad:
        return ExpectationZFunction.apply(state, nqubit, masks, batch)
    return engine.expectation_z(state, nqubit, masks, batch)


class CircuitFunction(torch.autograd.Function):
    """y = U(mats) x for a whole fused program; backward by un-computation."""

    @staticmethod
    def forward(ctx, x, mats, prog, batch, mbs):
        y = x.detach().clone()
        md = mats.detach()
        prog.plan(y.dtype).run(y, md, batch, mbs)
        ctx.save_for_backward(y, md)
        ctx.prog, ctx.batch, ctx.mbs = prog, batch, mbs
        ctx.x_needs_grad = x.requires_grad
        return y

    @staticmethod
    def backward(ctx, grad_y):
        y, mats = ctx.saved_tensors
        prog = ctx.prog
        # the reverse sweep stages TWO tiles (psi, lambda) per CTA: with 32 KiB tiles (chunk_bits 11) two CTAs share an SM
        # and overlap their load / compute / store phases; the forward plan keeps its 64 KiB tiles.  Any partition of
        # the same gate list into passes un-computes the same unitary.
        cb = ADJOINT_CHUNK_BITS
        plan = prog.plan(y.dtype, chunk_bits=cb) if cb else prog.plan(y.dtype)
        batch, mbs = ctx.batch, ctx.mbs
        psi = y.clone().reshape(batch, -1)
        lam = grad_y.contiguous().clone().reshape(batch, -1)
        # cotangent of the matrix buffer, always accumulated in double precision
        grad_m = torch.zeros(mats.shape, dtype=torch.complex128, device=mats.device)
        # only gates whose matrix is computed from parameters / data need a gradient
        def _needs(r):
            block = prog.low.derived[r[5]][4] if r[4] == 'derived' else r[4]
            return 0 if block in ('none', 'const') else 1
        need = (C.c_uint8 * max(1, len(prog.low.records)))(*[_needs(r) for r in prog.low.records])
        if not ctx.needs_input_grad[1]:
            need = (C.c_uint8 * max(1, len(prog.low.records)))()
        lib = L.load()
        # the reverse sweep handles one state per call: a batch (2-D data, the reference's vmap, circuit.py:227-241,
        # or a batch of initial states) is swept sample by sample; with per-sample matrices (mbs != 0) every sample
        # has its own cotangent row, shared matrices accumulate into the same buffer
        for b in range(batch):
            m_b = mats[b] if mbs else mats
            g_b = grad_m[b] if mbs else grad_m
            L.check(lib.b200q_adjoint_run(plan._h, psi[b].data_ptr(), lam[b].data_ptr(), m_b.data_ptr(), g_b.data_ptr(),
                                          need, engine._stream(psi)))
        gm = grad_m.to(mats.dtype) if ctx.needs_input_grad[1] else None
        return (lam.reshape(grad_y.shape) if ctx.x_needs_grad else None), gm, None, None, None
