"""The C-ABI library loads and exports every symbol include/b200q.h declares; the host-only entry
points (planner, argument validation) work without a GPU.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gates_np
from helpers import lower_ops

from deepquantum_b200 import _lib as L
from deepquantum_b200 import build as libbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    libbuild.build()
    return L.load()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, 'include', 'b200q.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(b200q_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/b200q.h but not exported'
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    assert b'sm_100a' in lib.b200q_version()


def test_struct_layout_matches_header(lib):
    assert C.sizeof(L.GateStruct) == 4 + 4 + 4 * 6 + 8 + 8 + 4 + 4
    assert C.sizeof(L.PlanOptions) == 32
    assert C.sizeof(L.PlanStats) == 32


def test_planner_host_only(lib):
    n = 20
    ops = []
    for w in range(n):
        ops.append((gates_np.H, [w], []))
        ops.append((gates_np.X, [(w + 1) % n], [w]))
        ops.append((gates_np.rz(0.1 * w), [w], []))
    arr, ng, _ = lower_ops(ops, n, np.complex64)
    h = C.c_void_p()
    L.check(lib.b200q_plan_create(n, L.C64, arr, ng, None, C.byref(h)))
    st = L.PlanStats()
    L.check(lib.b200q_plan_get_stats(h, C.byref(st)))
    assert st.n_gates == ng and 1 <= st.n_passes < ng / 4
    assert sum(lib.b200q_plan_pass_gates(h, i) for i in range(st.n_passes)) == ng
    need = C.c_size_t()
    L.check(lib.b200q_plan_export(h, None, 0, C.byref(need)))
    assert need.value % st.n_passes == 0 and need.value // st.n_passes < 4096   # fits kernel parameter space
    lib.b200q_plan_destroy(h)


def test_invalid_arguments_are_rejected(lib):
    h = C.c_void_p()
    g = L.make_gate(L.GATE_MAT, [3], [3])          # control == target (operation.py:98: 'Use repeated wires')
    assert lib.b200q_plan_create(5, L.C64, C.byref(g), 1, None, C.byref(h)) < 0
    assert b'control' in lib.b200q_last_error()
    g = L.make_gate(L.GATE_MAT, [7])               # target out of range (operation.py:97)
    assert lib.b200q_plan_create(5, L.C64, C.byref(g), 1, None, C.byref(h)) < 0
    g = L.make_gate(L.GATE_X, [1, 2])
    assert lib.b200q_plan_create(5, L.C64, C.byref(g), 1, None, C.byref(h)) < 0
    assert lib.b200q_plan_create(5, 7, C.byref(g), 1, None, C.byref(h)) < 0
    assert lib.b200q_norm2(None, 5, L.C64, 1, None, None) < 0


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    import torch

    import deepquantum_b200 as dq
    cir = dq.QubitCircuit(3)
    cir.h(0)
    cir.cnot(0, 1)
    with pytest.raises(dq.B200QError):
        cir()
    with pytest.raises(dq.B200QError):
        dq.Hadamard(nqubit=2, wires=[0])(torch.tensor([1, 0, 0, 0], dtype=torch.cfloat))
    with pytest.raises(dq.B200QError):
        dq.evolve_state(torch.zeros(1, 2, 2, dtype=torch.cfloat), torch.eye(2, dtype=torch.cfloat), 2, [0])


def test_reference_side_stub_imports_and_installs():
    """INTEGRATION.md section 2 as an importable file (tools/patch_reference.py): binds the C-ABI symbols and, when
    the reference tree is mounted (build container only), installs itself into the reference and leaves a CPU
    circuit unchanged."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('patch_reference', os.path.join(root, 'tools', 'patch_reference.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ('evolve_state', 'op_state_control', 'evolve_den_mat', 'install'):
        assert callable(getattr(mod, name))
    mod._self_check()
