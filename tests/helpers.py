"""Shared test helpers: lowering oracle ops to C-ABI gate records (test-side, classification by
matrix structure) and loading the host emulator."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deepquantum_b200 import _lib as L  # noqa: E402

_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)


def classify(matrix):
    m = np.asarray(matrix)
    if m.shape == (2, 2) and np.array_equal(m, _X):
        return L.GATE_X
    if np.count_nonzero(m - np.diag(np.diagonal(m))) == 0:
        return L.GATE_DIAG
    return L.GATE_MAT


def lower_ops(ops, nqubit, dtype=np.complex128, hints=True):
    """(matrix, wires, controls[, {'grad': bool, 'adjoint': bool}]) tuples -> (GateStruct array, flat matrix buffer).
    'adjoint': the STORED matrix is `matrix`, the gate applied is its conjugate transpose; 'grad': B200Q_GATE_GRAD."""
    gates, mats, off = [], [], 0
    for entry in ops:
        matrix, wires, controls = entry[:3]
        extra = entry[3] if len(entry) > 3 else {}
        m = np.asarray(matrix, dtype=dtype)
        kind = classify(m)
        k = len(wires)
        targets = [nqubit - 1 - w for w in reversed(wires)]   # matrix LSB first
        ctr = [nqubit - 1 - c for c in controls]
        hint = 0
        if kind == L.GATE_MAT and k == 1 and hints:
            if np.all(m.imag == 0):
                hint = L.GATE_REAL
                if m[0, 0] != 0 and m[0, 0] == m[0, 1] == m[1, 0] == -m[1, 1]:
                    hint |= L.GATE_HADAMARD
                elif m[0, 0] == m[1, 1] and m[0, 1] == -m[1, 0] and abs(np.linalg.det(m) - 1) < 1e-6:
                    hint |= L.GATE_ROTATION
            elif m[0, 0].imag == 0 and m[1, 1].imag == 0 and m[0, 1].real == 0 and m[1, 0].real == 0:
                hint = L.GATE_RXLIKE
                if m[0, 0] == m[1, 1] and m[0, 1] == m[1, 0] and abs(np.linalg.det(m) - 1) < 1e-6:
                    hint |= L.GATE_ROTATION
        if kind == L.GATE_DIAG and k == 1 and hints and m[0, 0] == 1 and m[0, 1] == 0 and m[1, 0] == 0:
            hint = {1j: L.GATE_PHASE_S, -1: L.GATE_PHASE_Z, -1j: L.GATE_PHASE_SDG}.get(complex(m[1, 1]), 0)
        if extra.get('grad'):
            hint |= L.GATE_GRAD
        gates.append(L.make_gate(kind, targets, ctr, off, bool(extra.get('adjoint')), hint))
        mats.append(m.reshape(-1))
        off += m.size
    arr = (L.GateStruct * max(1, len(gates)))(*gates)
    buf = np.concatenate(mats) if mats else np.zeros(1, dtype=dtype)
    return arr, len(gates), np.ascontiguousarray(buf.astype(dtype))


_emu = None


def hostemu():
    global _emu
    if _emu is None:
        sys.path.insert(0, os.path.join(ROOT, 'tests', 'native'))
        import build as emu_build
        lib = C.CDLL(emu_build.build())
        lib.hostemu_run.restype = C.c_int
        lib.hostemu_run.argtypes = [C.c_int, C.c_int, C.POINTER(L.GateStruct), C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int),
                                    C.c_char_p, C.c_int]
        _emu = lib
    return _emu


def emu_run(ops, nqubit, cdtype, state=None, chunk_bits=0, low_bits=0, max_rounds=0, fuse=1, batch=1):
    """Run lowered ops through planner + CPU-stepped kernel body; returns (state, stats)."""
    arr, ng, mats = lower_ops(ops, nqubit, cdtype)
    if state is None:
        state = np.zeros((batch, 2**nqubit), dtype=cdtype)
        state[:, 0] = 1
    else:
        state = np.ascontiguousarray(np.asarray(state, dtype=cdtype).reshape(batch, 2**nqubit)).copy()
    stats = (C.c_int * 4)()
    err = C.create_string_buffer(256)
    rc = hostemu().hostemu_run(nqubit, L.C64 if cdtype == np.complex64 else L.C128, arr, ng, chunk_bits, low_bits,
                               max_rounds, fuse, state.ctypes.data, mats.ctypes.data, batch, 0, stats, err, 256)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return state, {'passes': stats[0], 'rounds': stats[1], 'ops': stats[2], 'direct': stats[3]}


def emu_run_program(prog, nqubit, cdtype, state=None, batch=1, **opts):
    """Run a product-side lowered program (`deepquantum_b200.circuit._Program`) through the CPU-stepped
    kernel body.  TEST-ONLY executor for the host logic (lowering, matrix assembly, plan)."""
    import torch
    tdt = torch.complex64 if cdtype == np.complex64 else torch.complex128
    mats_t = prog.low.build_matrices(tdt, 'cpu').detach()
    mbs = mats_t.shape[-1] if mats_t.ndim == 2 else 0
    mats = np.ascontiguousarray(mats_t.numpy())
    arr = (L.GateStruct * max(1, len(prog.structs)))(*prog.structs)
    if state is None:
        state = np.zeros((batch, 2**nqubit), dtype=cdtype)
        state[:, 0] = 1
    else:
        state = np.ascontiguousarray(np.asarray(state, dtype=cdtype).reshape(batch, 2**nqubit)).copy()
    stats = (C.c_int * 4)()
    err = C.create_string_buffer(256)
    rc = hostemu().hostemu_run(nqubit, L.C64 if cdtype == np.complex64 else L.C128, arr, len(prog.structs),
                               opts.get('chunk_bits', 0), opts.get('low_bits', 0), opts.get('max_rounds', 0),
                               opts.get('fuse', 1), state.ctypes.data, mats.ctypes.data, batch, mbs, stats, err, 256)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return state, {'passes': stats[0], 'rounds': stats[1], 'ops': stats[2], 'direct': stats[3]}


def emu_adjoint(ops, nqubit, cdtype, psi_final, lam_final, chunk_bits=0, need=None):
    """Reverse sweep through the CPU-stepped kernel body.  Returns (psi_initial, lam_initial, grad) with grad a
    complex128 array laid out like the matrix buffer."""
    arr, ng, mats = lower_ops(ops, nqubit, cdtype)
    psi = np.ascontiguousarray(np.asarray(psi_final, dtype=cdtype)).copy()
    lam = np.ascontiguousarray(np.asarray(lam_final, dtype=cdtype)).copy()
    grad = np.zeros(mats.size, dtype=np.complex128)
    lib = hostemu()
    lib.hostemu_adjoint.restype = C.c_int
    lib.hostemu_adjoint.argtypes = [C.c_int, C.c_int, C.POINTER(L.GateStruct), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
    err = C.create_string_buffer(256)
    needp = None
    if need is not None:
        need = np.ascontiguousarray(np.asarray(need, dtype=np.uint8))
        needp = need.ctypes.data
    rc = lib.hostemu_adjoint(nqubit, L.C64 if cdtype == np.complex64 else L.C128, arr, ng, chunk_bits,
                             psi.ctypes.data, lam.ctypes.data, mats.ctypes.data, grad.ctypes.data, needp, err, 256)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return psi, lam, grad


# ---- generated (specialised) pass kernels, stepped on the CPU ---------------------------------------------------
def codegen_sources(nqubit, cdtype, arr, ngates, chunk_bits=0, remote_last=False, **planopts):
    """Plan the gates with the product library and return, per pass, (source, pass header dict).  Host only."""
    lib = L.load()
    opt = L.PlanOptions()
    opt.chunk_bits = chunk_bits
    opt.fuse = planopts.get('fuse', 1)
    opt.low_bits = planopts.get('low_bits', 0)
    opt.max_rounds = planopts.get('max_rounds', 0)
    plan = C.c_void_p()
    L.check(lib.b200q_plan_create(nqubit, L.C64 if cdtype == np.complex64 else L.C128, arr, ngates, C.byref(opt),
                                  C.byref(plan)))
    try:
        st = L.PlanStats()
        L.check(lib.b200q_plan_get_stats(plan, C.byref(st)))
        need = C.c_size_t()
        L.check(lib.b200q_plan_export(plan, None, 0, C.byref(need)))
        raw = (C.c_uint8 * need.value)()
        L.check(lib.b200q_plan_export(plan, raw, need.value, C.byref(need)))
        psize = need.value // max(1, st.n_passes)
        out = []
        for i in range(st.n_passes):
            hdr = bytes(raw[i * psize:i * psize + 5])
            info = {'n_bits': hdr[0], 'n_qubits': hdr[1], 'tile_bits': hdr[2], 'n_rounds': hdr[3], 'n_ops': hdr[4]}
            n = C.c_size_t()
            smem = C.c_size_t()
            remote = int(remote_last and i == st.n_passes - 1)
            rc = lib.b200q_plan_codegen(plan, i, remote, None, 0, C.byref(n), C.byref(smem))
            if rc != 0:
                out.append((None, info))
                continue
            buf = C.create_string_buffer(n.value)
            L.check(lib.b200q_plan_codegen(plan, i, remote, buf, n.value, C.byref(n), C.byref(smem)))
            info['stats'] = lib.b200q_last_error().decode()
            info['smem'] = smem.value
            out.append((buf.value.decode(), info))
        return out
    finally:
        lib.b200q_plan_destroy(plan)


def compile_generated_host(source):
    """g++ build of one generated pass source (its host half): returns the loaded library."""
    import hashlib
    import subprocess
    d = os.path.join(ROOT, 'tests', 'native', '_build', 'gen')
    os.makedirs(d, exist_ok=True)
    h = hashlib.sha1(source.encode()).hexdigest()[:20]
    so = os.path.join(d, h + '.so')
    if not os.path.exists(so):
        src = os.path.join(d, h + '.cpp')
        with open(src, 'w') as f:
            f.write(source)
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-x', 'c++', src, '-o', so + '.tmp'])
        os.replace(so + '.tmp', so)
    lib = C.CDLL(so)
    lib.b200qj_emulate.restype = None
    lib.b200qj_emulate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_uint64, C.c_int,
                                   C.c_void_p]
    return lib


def gen_run(ops, nqubit, cdtype, state=None, chunk_bits=0, batch=1, **planopts):
    """Run lowered ops through planner + GENERATED pass kernels stepped on the CPU (passes the generator does not
    cover go through the emulator of the generic kernel body).  Returns (state, list of pass infos)."""
    arr, ng, mats = lower_ops(ops, nqubit, cdtype)
    if state is None:
        state = np.zeros((batch, 2**nqubit), dtype=cdtype)
        state[:, 0] = 1
    else:
        state = np.ascontiguousarray(np.asarray(state, dtype=cdtype).reshape(batch, 2**nqubit)).copy()
    vs = 1 if cdtype == np.complex64 else 0
    infos = []
    for src, info in codegen_sources(nqubit, cdtype, arr, ng, chunk_bits=chunk_bits, **planopts):
        assert src is not None, 'pass not covered by the generator'
        lib = compile_generated_host(src)
        tile_shift = info['n_bits'] - info['tile_bits']
        lib.b200qj_emulate(state.ctypes.data, mats.ctypes.data, (2**nqubit) >> vs, 0, tile_shift,
                           (1 << tile_shift) * batch, 1, None)
        infos.append(info)
    return state, infos
