"""The per-pass kernel generator (csrc/b200q_codegen.cpp) checked on the CPU: the generated source of every pass
is compiled with g++ (its host half defines the packed-FP32 primitives as plain float pairs) and stepped thread by
thread on a numpy state, then compared with the oracle.  Covers the Pauli-frame bookkeeping (thread-level CNOTs,
Hadamard / rotation propagation, CNOT renaming between register slots), the lane slot of complex64, diagonal ops
in every selector configuration, register / thread / tile-level controls and the in-tile dense contraction."""
import numpy as np
import pytest

import gates_np
import statevec_oracle as so
from helpers import gen_run

TOL = {np.complex128: 1e-12, np.complex64: 3e-6}


def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return psi / np.linalg.norm(psi)


@pytest.mark.parametrize('fuse', [0, 1])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_generated_every_bit_position(fuse, cdtype):
    n = 14 if cdtype == np.complex128 else 15
    rng = np.random.default_rng(5)
    psi = _rand_state(n, 5)
    ops = []
    for w in range(n):
        ops.append((gates_np.u3(*rng.uniform(0, 6, 3)), [w], []))
        ops.append((gates_np.X, [w], [(w + 3) % n, (w + 7) % n]))
        ops.append((gates_np.rz(0.7 + w), [w], []))
        ops.append((gates_np.rzz(0.2 + w), [w, (w + 5) % n], [(w + 1) % n]))
        ops.append((gates_np.ry(0.4 + w), [(w + 2) % n], [w]))
    if fuse:   # one-gate passes are slow to compile one by one: keep the dense blocks for the fused run
        for k, wires, ctr in [(2, [n - 1, 2], []), (2, [0, n - 2], [5]), (3, [1, n - 1, 6], []), (4, [3, 0, n - 3, 8], [1])]:
            q, _ = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
            ops.append((q, wires, ctr))
    else:
        ops = ops[:25]
    ref = so.run_circuit(ops, n, state=psi)
    out, infos = gen_run(ops, n, cdtype, state=psi, chunk_bits=11, fuse=fuse)
    err = np.linalg.norm(out[0] - ref)
    assert err < TOL[cdtype], (err, infos)


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_generated_structured_ops_and_frame(seed, cdtype):
    """H / Rx / Ry / S / CNOT mixes over the full 4 pi period with thread-level, register-slot and lane controls:
    every Pauli-frame rule of the generator is exercised against the oracle."""
    n = 15 if cdtype == np.complex64 else 14
    rng = np.random.default_rng(100 + seed)
    psi = _rand_state(n, seed)
    ops = []
    for _ in range(160):
        w = int(rng.integers(n))
        c = int((w + 1 + rng.integers(n - 1)) % n)
        th = float(rng.uniform(0, 4 * np.pi))
        kind = int(rng.integers(11))
        if rng.integers(5) == 0:
            w = n - 1   # index bit 0: the complex64 lane slot
            c = int(rng.integers(n - 1))
        elif rng.integers(5) == 0:
            c = n - 1
            w = int(rng.integers(n - 1))
        if kind == 0:
            ops.append((gates_np.H, [w], []))
        elif kind == 1:
            ops.append((gates_np.rx(th), [w], []))
        elif kind == 2:
            ops.append((gates_np.ry(th), [w], []))
        elif kind in (3, 9, 10):
            ops.append((gates_np.X, [w], [c]))
        elif kind == 4:
            ops.append((gates_np.rx(th), [w], [c]))
        elif kind == 5:
            ops.append((gates_np.ry(th), [w], [c]))
        elif kind == 6:
            ops.append((gates_np.H, [w], [c]))
        elif kind == 7:
            ops.append((gates_np.rx(th).conj().T, [w], []))
        elif kind == 8:
            ops.append((gates_np.S, [w], []))
        if rng.integers(4) == 0:
            ops.append((gates_np.rz(th), [w], [c] if rng.integers(2) else []))
        if rng.integers(6) == 0:
            ops.append((gates_np.p(th), [c], [w]))
        if rng.integers(8) == 0:
            ops.append((np.diag([1, 1, 1, -1]).astype(complex), [w, c], []))
        if rng.integers(8) == 0:
            ops.append((gates_np.rzz(th), [c, w], []))
        if rng.integers(10) == 0:
            ops.append((gates_np.X, [w], []))
        if rng.integers(12) == 0:
            ops.append((gates_np.Z, [w], []))
    ref = so.run_circuit(ops, n, psi)
    out, infos = gen_run(ops, n, cdtype, state=psi, chunk_bits=11)
    err = np.linalg.norm(out[0] - ref) / np.linalg.norm(ref)
    assert err < TOL[cdtype], (err, infos)
    total = ' '.join(i['stats'] for i in infos)
    assert 'frame_x' in total


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_generated_c2_like_circuit(cdtype):
    """The headline generator's gate mix (H / S / RX layers + CNOTs on random matchings) at 16 qubits."""
    n, depth = 16, 8
    rng = np.random.default_rng(11)
    ops = []
    for _ in range(depth):
        for q in range(n):
            k = int(rng.integers(3))
            ops.append(((gates_np.H, gates_np.S, gates_np.rx(float(rng.uniform(0, 2 * np.pi))))[k], [q], []))
        perm = rng.permutation(n)
        for i in range(0, n - 1, 2):
            ops.append((gates_np.X, [int(perm[i + 1])], [int(perm[i])]))
    ref = so.run_circuit(ops, n)
    out, infos = gen_run(ops, n, cdtype, chunk_bits=11)
    err = np.linalg.norm(out[0] - ref) / np.linalg.norm(ref)
    assert err < TOL[cdtype], (err, infos)


def test_generated_batch_of_states():
    n, batch = 13, 3
    rng = np.random.default_rng(3)
    st = np.stack([_rand_state(n, 40 + b) for b in range(batch)])
    ops = [(gates_np.H, [2], []), (gates_np.X, [5], [2]), (gates_np.rx(0.3), [12], []), (gates_np.rz(1.1), [0], [12]),
           (gates_np.ry(2.2), [7], [0])]
    ref = np.stack([so.run_circuit(ops, n, state=st[b]) for b in range(batch)])
    out, _ = gen_run(ops, n, np.complex128, state=st, chunk_bits=11, batch=batch)
    assert np.abs(out - ref).max() < 1e-12


def _remote_struct(nl, cdtype, bufs, world, rank, perm):
    """b200q_remote_t for one rank (same construction as b200q_plan_run_exchange, csrc/b200q_lib.cu)."""
    import ctypes as C

    class Remote(C.Structure):
        _fields_ = [('peer', C.c_void_p * 8), ('base', C.c_uint64), ('perm', C.c_uint8 * 40), ('n_chunk_bits', C.c_int32),
                    ('enabled', C.c_int32)]
    vs = 1 if cdtype == np.complex64 else 0
    g = world.bit_length() - 1
    R = Remote()
    for q in range(world):
        R.peer[q] = bufs[q].ctypes.data
    R.n_chunk_bits = nl - vs
    for j in range(vs, nl):
        R.perm[j - vs] = perm[j] - vs
    base = 0
    for k in range(g):
        if (rank >> k) & 1:
            pos = perm[nl + k] - vs
            base |= (1 << pos) if pos < R.n_chunk_bits else (1 << (40 + pos - R.n_chunk_bits))
    R.base = base
    R.enabled = 1
    return R


@pytest.mark.parametrize('world', [2, 8])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_generated_fused_exchange_bit_permutation(cdtype, world):
    """The fused-exchange variant of the generated write-back round: every chunk goes to (rank, position) =
    permutation of the distributed index bits; expected = segment on every shard, then a numpy bit permutation."""
    import ctypes as C
    from helpers import codegen_sources, compile_generated_host, lower_ops
    nl = 13
    g = world.bit_length() - 1
    nt = nl + g
    rng = np.random.default_rng(10 + world)
    ops = []
    for _ in range(40):
        w = int(rng.integers(nl))
        c = int((w + 1 + rng.integers(nl - 1)) % nl)
        k = int(rng.integers(5))
        ops.append([(gates_np.H, [w], []), (gates_np.rx(float(rng.uniform(0, 12))), [w], []), (gates_np.X, [w], [c]),
                    (gates_np.S, [w], []), (gates_np.u3(*rng.uniform(0, 6, 3)), [w], [c])][k])
    shards = [(rng.normal(size=2**nl) + 1j * rng.normal(size=2**nl)).astype(cdtype) for _ in range(world)]
    vs = 1 if cdtype == np.complex64 else 0
    perm = list(range(vs)) + [int(x) + vs for x in rng.permutation(nt - vs)]
    local = np.concatenate([so.run_circuit(ops, nl, state=s.astype(np.complex128)) for s in shards])
    old = np.arange(2**nt, dtype=np.int64)
    new = np.zeros_like(old)
    for j in range(nt):
        new |= ((old >> j) & 1) << perm[j]
    full = np.empty_like(local)
    full[new] = local
    arr, ng, mats = lower_ops(ops, nl, cdtype)
    srcs = codegen_sources(nl, cdtype, arr, ng, chunk_bits=11, remote_last=True)
    bufs = [np.full(2**nl, np.nan + 0j, dtype=cdtype) for _ in range(world)]
    for r in range(world):
        st = np.ascontiguousarray(shards[r].copy())
        R = _remote_struct(nl, cdtype, bufs, world, r, perm)
        for i, (src, info) in enumerate(srcs):
            lib = compile_generated_host(src)
            ts = info['n_bits'] - info['tile_bits']
            lib.b200qj_emulate(st.ctypes.data, mats.ctypes.data, (2**nl) >> vs, 0, ts, 1 << ts, 1,
                               C.addressof(R) if i == len(srcs) - 1 else None)
    got = np.concatenate(bufs)
    assert not np.isnan(got).any()
    assert np.linalg.norm(got - full) / np.linalg.norm(full) < TOL[cdtype]
