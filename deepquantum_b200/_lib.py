"""ctypes binding of libb200q.so (the C ABI in include/b200q.h).

The library is built in-tree by `__graft_entry__.build()` / `deepquantum_b200/build.py` into
`deepquantum_b200/lib/`.  There is NO CPU fallback: every compute entry point raises if the library
or an sm_100 device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

C64, C128 = 0, 1
GATE_MAT, GATE_DIAG, GATE_X = 0, 1, 2
GATE_ADJOINT = 1
GATE_REAL = 2
GATE_RXLIKE = 4
GATE_HADAMARD = 8
GATE_ROTATION = 16
GATE_PHASE_SHIFT = 5          # 1-target diagonal gates: exactly diag(1, i^q), q = 1 (S), 2 (Z), 3 (S^dagger)
GATE_PHASE_MASK = 3 << GATE_PHASE_SHIFT
GATE_PHASE_S, GATE_PHASE_Z, GATE_PHASE_SDG = 1 << 5, 2 << 5, 3 << 5
QUDIT_GENERAL, QUDIT_DIAG, QUDIT_DENSE1, QUDIT_NUMBER, QUDIT_DIFFERENCE = 0, 1, 2, 3, 4
GATE_GRAD = 128   # the cotangent of this gate will be asked for (dense gates on >= 3 targets get their own pass)


def conj_hint(hint: int) -> int:
    """Structure hint of the complex-conjugated matrix (column records of the density-matrix lowering)."""
    q = (hint & GATE_PHASE_MASK) >> GATE_PHASE_SHIFT
    return (hint & ~GATE_PHASE_MASK) | (((4 - q) & 3) << GATE_PHASE_SHIFT)
MAX_TARGETS = 6

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libb200q.so')


class QuditOpStruct(C.Structure):
    """b200q_qudit_op_t (include/b200q.h)."""
    _fields_ = [('n_targets', C.c_int32), ('modes', C.c_int32 * 2), ('structure', C.c_int32), ('matrix', C.c_uint64)]


class GateStruct(C.Structure):
    _fields_ = [
        ('kind', C.c_int32),
        ('n_targets', C.c_int32),
        ('targets', C.c_int32 * MAX_TARGETS),
        ('controls', C.c_uint64),
        ('mat_offset', C.c_int64),
        ('flags', C.c_int32),
        ('reserved', C.c_int32),
    ]


class PlanOptions(C.Structure):
    _fields_ = [
        ('chunk_bits', C.c_int32),
        ('low_bits', C.c_int32),
        ('max_rounds', C.c_int32),
        ('fuse', C.c_int32),
        ('reserved', C.c_int32 * 4),
    ]


class PlanStats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('n_gates', 'n_passes', 'n_rounds', 'n_ops', 'n_direct_ops', 'tile_bits',
                                         'threads_per_cta', 'smem_bytes')]


class QuditGateStruct(C.Structure):
    _fields_ = [('n_targets', C.c_int32), ('modes', C.c_int32 * 2), ('reserved', C.c_int32), ('mat_offset', C.c_int64)]


QUDIT_FUSED_MAX_GATES = 16


class B200QError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    'b200q_version': (C.c_char_p, []),
    'b200q_last_error': (C.c_char_p, []),
    'b200q_device_check': (C.c_int, [C.c_int]),
    'b200q_plan_create': (C.c_int, [C.c_int, C.c_int, C.POINTER(GateStruct), C.c_int, C.POINTER(PlanOptions),
                                    C.POINTER(C.c_void_p)]),
    'b200q_plan_destroy': (None, [C.c_void_p]),
    'b200q_plan_get_stats': (C.c_int, [C.c_void_p, C.POINTER(PlanStats)]),
    'b200q_plan_pass_gates': (C.c_int, [C.c_void_p, C.c_int]),
    'b200q_plan_export': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    'b200q_plan_run': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    'b200q_plan_run_range': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_void_p]),
    'b200q_plan_run_exchange': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                          C.POINTER(C.c_uint8), C.c_void_p]),
    'b200q_plan_codegen': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t)]),
    'b200q_plan_compile': (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    'b200q_plan_jit_status': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'b200q_apply_gate': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int,
                                   C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p]),
    'b200q_dense_tc_apply': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_uint64, C.c_int,
                                       C.c_void_p]),
    'b200q_norm2': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    'b200q_inner_product': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    'b200q_expectation_z': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_uint64,
                                      C.c_void_p, C.c_void_p]),
    'b200q_apply_z_weights': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_uint64, C.c_void_p]),
    'b200q_init_basis': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_void_p]),
    'b200q_adjoint_run': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_uint8), C.c_void_p]),
    'b200q_qudit_apply_group': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                          C.POINTER(QuditOpStruct), C.c_int, C.c_int64, C.c_void_p]),
    'b200q_plan_pass_gate_ids': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int]),
    'b200q_fock_bs_matrix': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'b200q_fock_squeezing_matrix': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'b200q_qudit_apply_structured': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32),
                                               C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    'b200q_qudit_apply': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int,
                                    C.c_int64, C.c_void_p]),
    'b200q_qudit_fused': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int,
                                    C.POINTER(QuditGateStruct), C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    'b200q_block_mass': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    'b200q_sample_blocks': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p]),
    'b200q_marginal_probs': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load libb200q.so (raises B200QError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200QError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                         '(nvcc, sm_100a).  deepquantum_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().b200q_last_error().decode()
        raise B200QError(f'b200q error {rc}: {msg}')


def make_gate(kind, targets, controls=(), mat_offset=0, adjoint=False, hint=0) -> GateStruct:
    g = GateStruct()
    g.kind = kind
    g.n_targets = len(targets)
    for j, t in enumerate(targets):
        g.targets[j] = int(t)
    mask = 0
    for c in controls:
        mask |= 1 << int(c)
    g.controls = mask
    g.mat_offset = int(mat_offset)
    g.flags = (GATE_ADJOINT if adjoint else 0) | hint
    return g
