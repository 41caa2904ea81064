"""Planner + tile-kernel body, stepped on the CPU (tests/native/hostemu.cpp), against the oracle."""
import json
import os

import numpy as np
import pytest

import gates_np
import statevec_oracle as so
from conftest import GOLDEN
from helpers import emu_run


def _golden_cases():
    g = np.load(os.path.join(GOLDEN, 'circuits.npz'))
    return sorted({k.split('/')[0] for k in g.files if k.endswith('/spec') and not k.startswith('batched')})


@pytest.mark.parametrize('case', _golden_cases())
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_golden_circuits(case, cdtype):
    g = np.load(os.path.join(GOLDEN, 'circuits.npz'))
    meta = json.loads(str(g[case + '/spec']))
    n, spec = meta['n'], meta['spec']
    ops = gates_np.lower_spec(spec, n)
    out, stats = emu_run(ops, n, cdtype, chunk_bits=11 if n > 12 else 0)
    ref = g[case + '/c128']
    err = np.linalg.norm(out[0] - ref) / np.linalg.norm(ref)
    assert err < (1e-12 if cdtype == np.complex128 else 3e-6), (err, stats)
    assert stats['passes'] >= 1


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
@pytest.mark.parametrize('n', [1, 2, 3, 4, 5, 6, 7])
def test_tiny_states(n, cdtype):
    rng = np.random.default_rng(n)
    ops = []
    for _ in range(12):
        w = int(rng.integers(n))
        ops.append((gates_np.u3(*rng.uniform(0, 6, 3)), [w], []))
        if n > 1:
            c = int((w + 1 + rng.integers(n - 1)) % n)
            ops.append((gates_np.X, [w], [c]))
            ops.append((gates_np.rz(0.3), [c], [w]))
    ref = so.run_circuit(ops, n)
    out, _ = emu_run(ops, n, cdtype)
    assert np.linalg.norm(out[0] - ref) < (1e-12 if cdtype == np.complex128 else 2e-6)


@pytest.mark.parametrize('fuse', [0, 1])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_every_bit_position_and_unfused(fuse, cdtype):
    """1-target, controlled, diagonal and dense k-target gates on every wire of a 14/15-qubit state
    (more than one tile with the smallest tile size), fused and one-gate-per-pass."""
    n = 14 if cdtype == np.complex128 else 15
    rng = np.random.default_rng(5)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    ops = []
    for w in range(n):
        ops.append((gates_np.u3(*rng.uniform(0, 6, 3)), [w], []))
        ops.append((gates_np.X, [w], [(w + 3) % n, (w + 7) % n]))
        ops.append((gates_np.rz(0.7 + w), [w], []))
        ops.append((gates_np.rzz(0.2 + w), [w, (w + 5) % n], [(w + 1) % n]))
        ops.append((gates_np.ry(0.4 + w), [(w + 2) % n], [w]))
    q, _ = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    ops.append((q, [n - 1, 2], []))
    ops.append((q, [0, n - 2], [5]))
    q3, _ = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))
    ops.append((q3, [1, n - 1, 6], []))
    q4, _ = np.linalg.qr(rng.normal(size=(16, 16)) + 1j * rng.normal(size=(16, 16)))
    ops.append((q4, [3, 0, n - 3, 8], [1]))
    ref = so.run_circuit(ops, n, state=psi)
    out, stats = emu_run(ops, n, cdtype, state=psi, chunk_bits=11, fuse=fuse)
    err = np.linalg.norm(out[0] - ref)
    assert err < (1e-12 if cdtype == np.complex128 else 3e-6), (err, stats)
    if fuse:
        assert stats['passes'] < len(ops) / 3
    else:
        assert stats['passes'] == len(ops)


def test_batched_states():
    g = np.load(os.path.join(GOLDEN, 'circuits.npz'))
    meta = json.loads(str(g['batched_n6/spec']))
    ops = gates_np.lower_spec(meta['spec'], meta['n'])
    out, _ = emu_run(ops, meta['n'], np.complex128, state=g['batched_n6/init'], batch=3)
    np.testing.assert_allclose(out, g['batched_n6/c128'], atol=1e-13)


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_adjoint_sweep_matches_torch_autograd(cdtype):
    """Reverse sweep (un-compute + cotangent of every gate matrix + cotangent of the input state) against
    PyTorch autograd through the CPU port of the reference's own contraction (oracle/torch_port.py)."""
    import torch

    import torch_port
    from helpers import emu_adjoint, lower_ops

    n = 13 if cdtype == np.complex128 else 14      # two tiles at chunk_bits = 11... and multi-pass
    rng = np.random.default_rng(7)
    ops = []
    # exactly unitary H: the un-computation applies U^dagger, which inverts U only if U is unitary
    # (the reference's float32-rounded constant H is unitary to 6e-8 only)
    hmat = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2.0)
    for w in range(n):
        ops.append((gates_np.u3(*rng.uniform(0, 6, 3)), [w], []))
        ops.append((gates_np.rx(0.3 + w), [(w + 4) % n], []))
        ops.append((hmat, [(w + 2) % n], [(w + 5) % n]))
        ops.append((gates_np.X, [w], [(w + 3) % n]))
        ops.append((gates_np.rz(0.7 + w), [w], []))
        ops.append((gates_np.rzz(0.2 + w), [w, (w + 5) % n], []))
        ops.append((gates_np.p(0.4 + w), [(w + 1) % n], [w]))
        ops.append((gates_np.ry(0.4 + w), [(w + 2) % n], [w, (w + 6) % n]))
    ops.append((gates_np.rxx(0.8), [n - 1, 2], []))
    ops.append((gates_np.rxy(0.5), [0, n - 2], [5]))
    psi0 = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi0 /= np.linalg.norm(psi0)
    wvec = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)   # L = Re <w|psi> + <psi|D|psi>
    diag = rng.normal(size=2**n)

    tm = [torch.tensor(np.asarray(m, dtype=np.complex128), requires_grad=True) for m, _, _ in ops]
    x0 = torch.tensor(psi0, requires_grad=True)
    x = x0.reshape([1] + [2] * n)
    for (m, wires, ctr), t in zip(ops, tm):
        x = torch_port.evolve_state_controlled(x, t, n, wires, ctr) if ctr else torch_port.evolve_state(x, t, n, wires)
    psi = x.reshape(-1)
    loss = (torch.tensor(wvec).conj() * psi).sum().real + (torch.tensor(diag) * (psi.real**2 + psi.imag**2)).sum()
    loss.backward()
    lam_final = wvec + 2 * diag * psi.detach().numpy()           # PyTorch cotangent dL/dRe + i dL/dIm
    tol = 1e-10 if cdtype == np.complex128 else 3e-4

    psi_in, lam_in, grad = emu_adjoint(ops, n, cdtype, psi.detach().numpy(), lam_final, chunk_bits=11)
    assert np.linalg.norm(psi_in - psi0) < (1e-10 if cdtype == np.complex128 else 1e-4)
    assert np.linalg.norm(lam_in - x0.grad.numpy()) / np.linalg.norm(x0.grad.numpy()) < tol
    off = 0
    for (m, wires, ctr), t in zip(ops, tm):
        m = np.asarray(m)
        g = grad[off:off + m.size].reshape(m.shape)
        ref = t.grad.numpy()
        if np.array_equal(m, gates_np.X):
            off += m.size
            continue
        if np.count_nonzero(m - np.diag(np.diagonal(m))) == 0:      # diagonal gates: only the diagonal is read
            g, ref = np.diagonal(g), np.diagonal(ref)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(g - ref).max() / scale < tol, (wires, ctr, g, ref)
        off += m.size


@pytest.mark.parametrize('fuse', [1, 3])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
@pytest.mark.parametrize('seed', [0, 1, 2])
def test_structured_butterflies(seed, cdtype, fuse):
    """Hadamard (add/sub + deferred scalar) and rotation (three in-place shears, matrix negated when
    cos < 0) fast paths: angles over the full 4*pi period, mixed with CNOT relabelling (flip states),
    thread-level / global controls (sign goes to the thread phase) and inverses."""
    n = 15 if cdtype == np.complex64 else 14
    rng = np.random.default_rng(100 + seed)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    ops = []
    for _ in range(160):
        w = int(rng.integers(n))
        c = int((w + 1 + rng.integers(n - 1)) % n)
        th = float(rng.uniform(0, 4 * np.pi))
        kind = int(rng.integers(9))
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
        elif kind == 3:
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
    ref = so.run_circuit(ops, n, psi)
    out, stats = emu_run(ops, n, cdtype, state=psi, chunk_bits=11, fuse=fuse)
    err = np.linalg.norm(out[0] - ref) / np.linalg.norm(ref)
    assert err < (1e-12 if cdtype == np.complex128 else 3e-6), (err, stats)


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_adjoint_lean_ops_with_pending_scalars(cdtype):
    """Reverse sweep on the lean op set: un-controlled Hadamards whose gradient is NOT requested run as bare
    add/sub butterflies (their scalar stays pending on psi and lambda and is folded into the other ops' cotangents
    at flush time), thread-level diagonals only touch the pending phase, rotations over the full 4*pi period."""
    torch = pytest.importorskip('torch')
    import torch_port
    from helpers import emu_adjoint

    n = 13 if cdtype == np.complex128 else 14
    rng = np.random.default_rng(11)
    hmat = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2.0)
    ops, need = [], []
    for w in range(n):
        ops.append((hmat, [w], []));                                   need.append(0)
    for layer in range(3):
        for w in range(n):
            c = (w + 3 + layer) % n
            ops.append((gates_np.X, [w], [c]));                         need.append(0)
            ops.append((gates_np.rz(float(rng.uniform(0, 12))), [w], [])); need.append(1)
            ops.append((gates_np.X, [w], [c]));                         need.append(0)
            ops.append((gates_np.rx(float(rng.uniform(0, 12))), [c], [])); need.append(1)
            if w % 3 == 0:
                ops.append((hmat, [(w + 1) % n], []));                  need.append(0)
                ops.append((gates_np.ry(float(rng.uniform(0, 12))), [w], [c])); need.append(1)
                ops.append((gates_np.S, [c], []));                      need.append(0)
                ops.append((gates_np.rzz(0.3 + w), [w, c], []));        need.append(1)
    psi0 = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi0 /= np.linalg.norm(psi0)
    wvec = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    tm = [torch.tensor(np.asarray(m, dtype=np.complex128), requires_grad=True) for m, _, _ in ops]
    x0 = torch.tensor(psi0, requires_grad=True)
    x = x0.reshape([1] + [2] * n)
    for (m, wires, ctr), t in zip(ops, tm):
        x = torch_port.evolve_state_controlled(x, t, n, wires, ctr) if ctr else torch_port.evolve_state(x, t, n, wires)
    psi = x.reshape(-1)
    ((torch.tensor(wvec).conj() * psi).sum().real).backward()
    tol = 1e-10 if cdtype == np.complex128 else 3e-4
    psi_in, lam_in, grad = emu_adjoint(ops, n, cdtype, psi.detach().numpy(), wvec, chunk_bits=11, need=need)
    assert np.linalg.norm(psi_in - psi0) < (1e-10 if cdtype == np.complex128 else 1e-4)
    assert np.linalg.norm(lam_in - x0.grad.numpy()) / np.linalg.norm(x0.grad.numpy()) < tol
    off = 0
    for (m, wires, ctr), t, nd in zip(ops, tm, need):
        m = np.asarray(m)
        g = grad[off:off + m.size].reshape(m.shape)
        ref = t.grad.numpy()
        off += m.size
        if not nd:
            continue
        if np.count_nonzero(m - np.diag(np.diagonal(m))) == 0:
            g, ref = np.diagonal(g), np.diagonal(ref)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(g - ref).max() / scale < tol, (wires, ctr, g, ref)


def _exchange_all_ranks(ops, nl, cdtype, shards, world, perm=None, coalesce_bits=-1):
    """Every rank's plan with the fused exchange, all ranks in one address space (test-only emulator)."""
    import ctypes as C
    from deepquantum_b200 import _lib as L
    from helpers import hostemu, lower_ops
    arr, ng, mats = lower_ops(ops, nl, cdtype)
    states = [np.ascontiguousarray(s.copy()) for s in shards]
    bufs = [np.full(2**nl, np.nan + 0j, dtype=cdtype) for _ in range(world)]
    lib = hostemu()
    lib.hostemu_run_exchange_rank.restype = C.c_int
    lib.hostemu_run_exchange_rank.argtypes = [C.c_int, C.c_int, C.POINTER(L.GateStruct), C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int,
                                              C.POINTER(C.c_uint8), C.c_char_p, C.c_int]
    bp = (C.c_void_p * world)(*[b.ctypes.data for b in bufs])
    pm = None if perm is None else (C.c_uint8 * len(perm))(*perm)
    err = C.create_string_buffer(256)
    for r in range(world):
        rc = lib.hostemu_run_exchange_rank(nl, L.C64 if cdtype == np.complex64 else L.C128, arr, ng, 11,
                                           coalesce_bits, states[r].ctypes.data, bp, mats.ctypes.data, world, r, pm,
                                           err, 256)
        assert rc == 0, err.value.decode()
    return bufs


def _segment_ops(nl, rng, count=60):
    ops = []
    for _ in range(count):
        w = int(rng.integers(nl))
        c = int((w + 1 + rng.integers(nl - 1)) % nl)
        k = int(rng.integers(5))
        ops.append([(gates_np.H, [w], []), (gates_np.rx(float(rng.uniform(0, 12))), [w], []), (gates_np.X, [w], [c]),
                    (gates_np.S, [w], []), (gates_np.u3(*rng.uniform(0, 6, 3)), [w], [c])][k])
    return ops


@pytest.mark.parametrize('world', [2, 4, 8])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_fused_pass_and_block_transpose(cdtype, world):
    """`b200q_plan_run_exchange` with perm = NULL: the last pass of a local segment scatters every chunk straight
    into the receive buffer of the rank that owns it after the block transpose.  Expected = ordinary run of the
    segment on every shard followed by the all-to-all transpose in numpy."""
    nl = 14
    rng = np.random.default_rng(world)
    ops = _segment_ops(nl, rng)
    shards = [(rng.normal(size=2**nl) + 1j * rng.normal(size=2**nl)).astype(cdtype) for _ in range(world)]
    expect_local = [emu_run(ops, nl, cdtype, state=s, chunk_bits=11)[0][0] for s in shards]
    blk = 2**nl // world
    expect = [np.concatenate([expect_local[src][dst * blk:(dst + 1) * blk] for src in range(world)])
              for dst in range(world)]
    bufs = _exchange_all_ranks(ops, nl, cdtype, shards, world, coalesce_bits=3)
    tol = 1e-12 if cdtype == np.complex128 else 2e-6
    for dst in range(world):
        assert not np.isnan(bufs[dst]).any()
        assert np.linalg.norm(bufs[dst] - expect[dst]) / np.linalg.norm(expect[dst]) < tol


@pytest.mark.parametrize('world', [1, 2, 8])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_fused_pass_and_bit_permutation(cdtype, world):
    """The exchange as an arbitrary permutation of the bits of the distributed index (rank bits included):
    expected = segment on every shard, then numpy transposition of the full state seen as a [2]*n tensor."""
    nl = 13
    g = world.bit_length() - 1
    nt = nl + g
    rng = np.random.default_rng(10 + world)
    ops = _segment_ops(nl, rng, 40)
    shards = [(rng.normal(size=2**nl) + 1j * rng.normal(size=2**nl)).astype(cdtype) for _ in range(world)]
    vs = 1 if cdtype == np.complex64 else 0
    perm = list(range(vs)) + [int(x) + vs for x in rng.permutation(nt - vs)]      # bit j -> perm[j]
    local = np.concatenate([emu_run(ops, nl, cdtype, state=s, chunk_bits=11)[0][0] for s in shards])
    # full[new_index] = local[old_index], new_index = sum bit_j(old) << perm[j]
    old = np.arange(2**nt, dtype=np.int64)
    new = np.zeros_like(old)
    for j in range(nt):
        new |= ((old >> j) & 1) << perm[j]
    full = np.empty_like(local)
    full[new] = local
    bufs = _exchange_all_ranks(ops, nl, cdtype, shards, world, perm=perm)
    got = np.concatenate(bufs)
    assert not np.isnan(got).any()
    tol = 1e-12 if cdtype == np.complex128 else 2e-6
    assert np.linalg.norm(got - full) / np.linalg.norm(full) < tol


@pytest.mark.parametrize('world,circuit', [(2, 'fixture'), (4, 'fixture'), (8, 'fixture'), (4, 'c2'), (8, 'c2')])
def test_sharded_perm_schedule_all_ranks_in_one_process(world, circuit):
    """The 'perm' schedule of the sharded path (exchanges = bit permutations done by the fused last pass of a
    segment, layout restored by one permuting exchange): every rank's `ShardedProgram` stepped in lockstep in
    ONE process, local segments and fused exchanges executed by the CPU emulator of the kernel body, against the
    dense oracle.  (The NCCL / 'pswap' schedule is covered by tests/test_distributed_gloo.py.)"""
    import ctypes as C
    torch = pytest.importorskip('torch')
    import deepquantum_b200 as dq
    from deepquantum_b200 import _lib as L
    from deepquantum_b200.distributed import ShardedProgram
    from helpers import hostemu
    from test_distributed_gloo import _build

    n = 10 if circuit == 'fixture' else 13
    g = world.bit_length() - 1
    nl = n - g
    if circuit == 'fixture':
        dense = _build(dq.QubitCircuit(n), n)
    else:   # the bench generator, deep enough that segments span several passes: the scheduler cuts their sparse tails
        from deepquantum_b200 import workloads as wl
        dense = dq.QubitCircuit(n)
        wl.apply_spec(dense, wl.random_clifford_rx_spec(n, 14), torch.complex128)
    dense.to(torch.double)
    low = dense._get_program().low
    mats = low.build_matrices(torch.complex128, 'cpu').detach()
    ops = [(op.update_matrix().detach().numpy(), op.wires, op.controls) for op in dense.operators]
    ref = so.run_circuit(ops, n)
    lib = hostemu()
    lib.hostemu_run_exchange_rank.restype = C.c_int
    lib.hostemu_run_exchange_rank.argtypes = [C.c_int, C.c_int, C.POINTER(L.GateStruct), C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int,
                                              C.POINTER(C.c_uint8), C.c_char_p, C.c_int]

    class State:
        def __init__(self, r):
            self.amps = torch.zeros(2**nl, dtype=torch.complex128)
            self.buffer = torch.zeros(2**nl, dtype=torch.complex128)
            if r == 0:
                self.amps[0] = 1.0

        def enable_peer_exchange(self):
            return True

        def peer_buffer_ptrs(self):
            return [s.buffer.data_ptr() for s in states]

    class Exec:
        def make_plan(self, nlocal, dtype, structs, exchange=False):
            return list(structs)

        def _arr(self, structs):
            return (L.GateStruct * max(1, len(structs)))(*structs)

        def run_plan(self, plan, amps, m):
            err = C.create_string_buffer(256)
            mm = np.ascontiguousarray(m.numpy())
            rc = lib.hostemu_run(nl, L.C128, self._arr(plan), len(plan), 11, 0, 0, 1, amps.data_ptr(), mm.ctypes.data,
                                 1, 0, None, err, 256)
            assert rc == 0, err.value.decode()

        def run_plan_exchange(self, plan, amps, m, peers, rank, perm=None):
            err = C.create_string_buffer(256)
            mm = np.ascontiguousarray(m.numpy())
            bp = (C.c_void_p * world)(*peers)
            pm = None if perm is None else (C.c_uint8 * len(perm))(*perm)
            rc = lib.hostemu_run_exchange_rank(nl, L.C128, self._arr(plan), len(plan), 11, 3, amps.data_ptr(), bp,
                                               mm.ctypes.data, world, rank, pm, err, 256)
            assert rc == 0, err.value.decode()

    states = [State(r) for r in range(world)]
    os.environ['B200Q_SHARD_TRIM'] = '20' if circuit == 'c2' else '0'      # the segment cut is an option (off by default)
    try:
        progs = [ShardedProgram(low, n, world, r, 'perm') for r in range(world)]
    finally:
        os.environ.pop('B200Q_SHARD_TRIM', None)
    ex = Exec()
    for p in progs:
        p.fused_exchanges, p._skip_next = 0, False
    assert all(len(p.steps) == len(progs[0].steps) for p in progs)
    assert any(s[0] == 'xperm' for s in progs[0].steps)
    assert len({p.n_deferred for p in progs}) == 1          # every rank cuts its segments at the same gates
    if circuit == 'c2':
        assert progs[0].n_deferred > 0
    for si in range(len(progs[0].steps)):
        what = [progs[r].run_step(si, states[r], mats, ex, True) for r in range(world)]
        assert len(set(what)) == 1, what
        if what[0] == 'exchange':
            for r in range(world):
                ShardedProgram.commit_exchange(states[r])
    got = torch.cat([s.amps for s in states]).numpy()
    assert np.linalg.norm(got - ref) < 1e-12, np.linalg.norm(got - ref)


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_dense_gates_on_five_and_six_targets(cdtype):
    """UAnyGate-style dense blocks on 5 and 6 wires (reference gate.py:2745-2790 accepts any k): a pass of their own
    between fused tile passes, plain and adjoint, with a control, mixed with ordinary gates."""
    from helpers import lower_ops, hostemu
    n = 13
    rng = np.random.default_rng(8)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)

    def rand_u(k):
        q, _ = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
        return q

    ops = [(gates_np.H, [w], []) for w in range(n)]
    ops.append((rand_u(5), [0, 12, 3, 7, 5], []))
    ops += [(gates_np.rx(0.3 + w), [w], []) for w in range(n)]
    ops.append((gates_np.X, [2], [9]))
    ops.append((rand_u(6), [11, 1, 4, 8, 2, 6], [10]))
    ops.append((rand_u(5).conj().T, [4, 3, 2, 1, 0], []))
    ops += [(gates_np.ry(0.1 * w), [w], [(w + 1) % n]) for w in range(0, n, 3)]
    ref = so.run_circuit(ops, n, state=psi)
    out, stats = emu_run(ops, n, cdtype, state=psi, chunk_bits=11)
    err = np.linalg.norm(out[0] - ref) / np.linalg.norm(ref)
    assert err < (1e-12 if cdtype == np.complex128 else 5e-6), (err, stats)
    assert stats['direct'] >= 3


def _dense_grad_case(n, rng):
    """Ops (with trainable dense gates on 3, 4 and 5 wires, one controlled, one stored as its adjoint), a random
    input state and the autograd reference: (ops, psi0, psi_final, lam_final, x0.grad, [matrix grads])."""
    import torch

    import torch_port

    def rand_u(k):
        q, _ = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
        return q

    hmat = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2.0)
    ops = [(hmat, [w], []) for w in range(n)]
    ops.append((rand_u(3), [1, n - 1, 4], [], {'grad': True}))
    ops += [(gates_np.rx(0.3 + w), [w], []) for w in range(n)]
    ops.append((rand_u(4), [n - 2, 0, 5, 2], [7], {'grad': True}))
    ops.append((gates_np.X, [2], [9]))
    ops.append((rand_u(3), [3, 6, 8], [], {'grad': True, 'adjoint': True}))
    ops.append((rand_u(5), [0, n - 1, 3, 7, 5], [], {'grad': True}))
    ops.append((rand_u(3), [2, 3, 4], []))          # no cotangent wanted: stays inside a fused pass
    ops += [(gates_np.ry(0.1 * w + 0.2), [w], [(w + 1) % n]) for w in range(0, n, 3)]
    psi0 = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi0 /= np.linalg.norm(psi0)
    wvec = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    diag = rng.normal(size=2**n)
    tm = [torch.tensor(np.asarray(e[0], dtype=np.complex128), requires_grad=True) for e in ops]
    x0 = torch.tensor(psi0, requires_grad=True)
    x = x0.reshape([1] + [2] * n)
    for e, t in zip(ops, tm):
        u = t.conj().transpose(0, 1) if (len(e) > 3 and e[3].get('adjoint')) else t
        x = torch_port.evolve_state_controlled(x, u, n, e[1], e[2]) if e[2] else torch_port.evolve_state(x, u, n, e[1])
    psi = x.reshape(-1)
    loss = (torch.tensor(wvec).conj() * psi).sum().real + (torch.tensor(diag) * (psi.real**2 + psi.imag**2)).sum()
    loss.backward()
    lam_final = wvec + 2 * diag * psi.detach().numpy()
    return ops, psi0, psi.detach().numpy(), lam_final, x0.grad.numpy(), [t.grad.resolve_conj().numpy() for t in tm]


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_adjoint_sweep_dense_gates_on_three_to_five_targets(cdtype):
    """Cotangents of trainable dense gates on 3-5 wires (UAnyGate / LatentGate / HamiltonianGate with requires_grad,
    reference gate.py:2745-2931; the reference differentiates them by autograd): planned with B200Q_GATE_GRAD they
    get a pass of their own whose reverse step accumulates the full 2^k x 2^k cotangent."""
    from helpers import emu_adjoint
    n = 13
    ops, psi0, psi_f, lam_f, x0_grad, mgrads = _dense_grad_case(n, np.random.default_rng(12))
    need = [1 if (len(e) > 3 and e[3].get('grad')) or len(e[1]) == 1 else 0 for e in ops]
    need = [0 if np.array_equal(np.asarray(e[0]), gates_np.X) else v for e, v in zip(ops, need)]
    psi_in, lam_in, grad = emu_adjoint(ops, n, cdtype, psi_f, lam_f, chunk_bits=11, need=need)
    tol = 1e-10 if cdtype == np.complex128 else 3e-4
    assert np.linalg.norm(psi_in - psi0) < (1e-10 if cdtype == np.complex128 else 1e-4)
    assert np.linalg.norm(lam_in - x0_grad) / np.linalg.norm(x0_grad) < tol
    off, checked = 0, 0
    for e, ref, nd in zip(ops, mgrads, need):
        m = np.asarray(e[0])
        if nd and len(e[1]) >= 3:
            g = grad[off:off + m.size].reshape(m.shape)
            assert np.abs(g - ref).max() / max(1.0, np.abs(ref).max()) < tol, (e[1], e[2])
            checked += 1
        off += m.size
    assert checked == 4
