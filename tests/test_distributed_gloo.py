"""The sharded path on CPU: world_size 2 and 4 gloo ranks.  The host logic under test is the product's own
(`deepquantum_b200.distributed`: qubit map, per-rank gate rewriting, block-transpose schedule,
`torch.distributed` all-to-all); only the LOCAL fused plan is executed by the test-only CPU emulator of the
kernel body instead of the GPU.  Compared against the dense oracle."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


class EmuExecutor:
    """TEST-ONLY executor: plans with the product planner, steps the kernel body on the CPU."""

    def make_plan(self, nlocal, dtype, structs):
        return (nlocal, dtype, list(structs))

    def run_plan(self, plan, amps, mats):
        from helpers import hostemu
        from deepquantum_b200 import _lib as L
        nlocal, dtype, structs = plan
        arr = (L.GateStruct * max(1, len(structs)))(*structs)
        a = amps.numpy()
        m = np.ascontiguousarray(mats.numpy())
        err = C.create_string_buffer(256)
        rc = hostemu().hostemu_run(nlocal, L.C64 if dtype == torch.complex64 else L.C128, arr, len(structs), 11, 0, 0,
                                   1, a.ctypes.data, m.ctypes.data, 1, 0, None, err, 256)
        assert rc == 0, err.value.decode()

    # local pieces of the differentiable expectation: numpy stands in for the reductions, the CPU-stepped reverse
    # sweep of the kernel body for b200q_adjoint_run
    def expectation_z(self, amps, nlocal, masks, index_offset):
        p = np.abs(amps.numpy())**2
        idx = np.arange(p.size, dtype=np.int64) | index_offset
        out = [float((p * (1 - 2 * (np.array([bin(int(i) & int(m)).count('1') & 1 for i in idx])))).sum())
               for m in masks.tolist()]
        return torch.tensor(out, dtype=torch.float64)

    def apply_z_weights(self, amps, nlocal, masks, weights, index_offset):
        a = amps.numpy()
        idx = np.arange(a.size, dtype=np.int64) | index_offset
        f = np.zeros(a.size)
        for m, w in zip(masks.tolist(), weights.reshape(-1).tolist()):
            f += w * (1 - 2 * np.array([bin(int(i) & int(m)).count('1') & 1 for i in idx]))
        return torch.from_numpy(a * f)

    def run_plan_adjoint(self, plan, psi, lam, mats, grad, need):
        from helpers import hostemu
        from deepquantum_b200 import _lib as L
        nlocal, dtype, structs = plan
        arr = (L.GateStruct * max(1, len(structs)))(*structs)
        lib = hostemu()
        lib.hostemu_adjoint.restype = C.c_int
        lib.hostemu_adjoint.argtypes = [C.c_int, C.c_int, C.POINTER(L.GateStruct), C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        nd = np.ascontiguousarray(np.asarray(need, dtype=np.uint8))
        m = np.ascontiguousarray(mats.numpy())
        err = C.create_string_buffer(256)
        rc = lib.hostemu_adjoint(nlocal, L.C64 if dtype == torch.complex64 else L.C128, arr, len(structs), 11,
                                 psi.numpy().ctypes.data, lam.numpy().ctypes.data, m.ctypes.data,
                                 grad.numpy().ctypes.data, nd.ctypes.data if len(need) else None, err, 256)
        assert rc == 0, err.value.decode()

    # local pieces of measure_dist: the numpy oracle stands in for csrc/b200q_sample.cu
    def block_mass(self, amps, nlocal):
        p = np.abs(amps.numpy())**2
        bb = min(12, nlocal)
        return torch.from_numpy(p.reshape(-1, 2**bb).sum(1))

    def sample_indices(self, amps, nlocal, uniforms, mass):
        import sampling_oracle as smp
        return torch.from_numpy(smp.sample_indices(amps.numpy(), uniforms.numpy()).astype(np.int64))

    def marginal_probs(self, amps, nlocal, mask, keys_sorted):
        p = np.abs(amps.numpy())**2
        idx = np.arange(p.size) & mask
        return torch.tensor([p[idx == int(k)].sum() for k in keys_sorted.tolist()], dtype=torch.float64)


def dq_mod():
    import deepquantum_b200
    return deepquantum_b200


def _build(cir, n, extra=False):
    """A circuit that exercises every sharded case: global 1-target gates, global / local controls, global
    diagonal gates (1 and 2 targets, controlled), dense 2- and 3-target gates touching global wires, swaps."""
    g = torch.Generator().manual_seed(11)
    r = lambda: float(torch.rand(1, generator=g) * 6)   # noqa: E731
    cir.hlayer()
    for w in range(n):
        cir.rx(w, r())
    cir.cnot(0, n - 1)
    cir.cnot(n - 1, 0)
    cir.cx(1, 2)
    cir.rz(0, r())
    cir.rz(1, r(), controls=[n - 2])
    cir.rzz([0, n - 1], r())
    cir.rzz([1, 0], r())
    cir.cz(0, 1)
    cir.toffoli(0, 1, n - 1)
    cir.toffoli(n - 1, 2, 0)
    cir.rxx([0, n - 1], r())
    cir.ryy([1, 2], r(), controls=[0])
    cir.swap([0, n - 2])
    cir.swap([0, 1])
    cir.u3(1, [r(), r(), r()], controls=[0, n - 1])
    cir.p(0, r(), controls=[1])
    q, _ = torch.linalg.qr(torch.randn(8, 8, generator=g, dtype=torch.float64) + 1j * torch.randn(8, 8, generator=g,
                                                                                                 dtype=torch.float64))
    cir.any(q, wires=[n - 1, 0, 3])
    if extra:   # gloo runs only (tests/dist_gpu_worker.py keeps the circuit it was verified with on the GPUs)
        cir.hamiltonian([[0.4, 'x0z3'], [-0.7, 'y1']], t=r())                  # dense block spanning wires 0..3
        cir.hamiltonian([1.1, 'z0'], t=r(), controls=[n - 1])
        cir.add(dq_mod().CombinedSingleGate([dq_mod().Rx(0.3), dq_mod().Hadamard(), dq_mod().Rz(1.1)], nqubit=n,
                                            wires=[0], controls=[2]))
    cir.ylayer()
    for w in range(n):
        cir.ry(w, r())
    cir.cnot_ring()
    cir.t(0)
    cir.s(1)
    cir.h(0)
    return cir


def _worker(rank, world, port, n, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import deepquantum_b200 as dq
    r, w, _ = dq.setup_distributed('gloo')
    assert (r, w) == (rank, world)
    try:
        cir = _build(dq.DistributedQubitCircuit(n), n, extra=True)
        cir.to(torch.double)
        cir._executor = EmuExecutor()
        st = cir()
        assert st.amps.shape == (2**n // world,)
        shards = [torch.empty_like(st.amps) for _ in range(world)]
        if world > 1:
            dist.all_gather(shards, st.amps.contiguous())
        else:
            shards = [st.amps]
        import json
        meas = {}
        for name, wires in (('all', None), ('sub', [n - 1, 0, 2]), ('glob', [0])):
            torch.manual_seed(5)     # rank 0 draws the uniforms
            meas[name] = cir.measure(shots=300, with_prob=True, wires=wires)
            assert rank == 0 or meas[name] == {}
        if rank == 0:
            np.savez(os.path.join(outdir, 'out.npz'), state=torch.cat(shards).numpy(),
                     steps=np.array([s[0] for s in cir._sharded.steps]), measure=json.dumps(meas))
    finally:
        dq.cleanup_distributed()


def _worker_adjoint(rank, world, port, outdir, n=4):
    """The reference's own test of the differentiable sharded expectation (tests/test_circuit.py:87-139)."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import deepquantum_b200 as dq
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    dq.setup_distributed('gloo')
    try:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'dist_adjoint.npz'))
        data = torch.tensor(g[f'ref_test_n{n}/data'], dtype=torch.float64, requires_grad=True)
        cir = dq.DistributedQubitCircuit(n, reupload=True)
        cir.rxlayer(encode=True); cir.rylayer(encode=True); cir.rzlayer(encode=True); cir.u3layer(encode=True)
        cir.hlayer(); cir.cnot_ring(); cir.toffoli(0, 1, 2); cir.fredkin(2, 1, 0); cir.swap([2, 3])
        cir.rx(0, controls=[1, 2, 3], encode=True); cir.ry(1, controls=[0, 2, 3], encode=True)
        cir.rz(2, controls=[0, 1, 3], encode=True); cir.rxx([0, 1], controls=[2, 3], encode=True)
        cir.ryy([1, 2], controls=[0, 3], encode=True); cir.rzz([2, 3], controls=[0, 1], encode=True)
        cir.rxy([3, 0], controls=[1, 2], encode=True)
        cir.observable(0); cir.observable(1, 'x'); cir.observable([2, 3], 'xy')
        cir.to(torch.double)
        cir._executor = EmuExecutor()
        cir(data=data)
        exp = cir.expectation()
        exp.sum().backward()
        if rank == 0:
            np.savez(os.path.join(outdir, 'adj.npz'), expectation=exp.detach().numpy(), grad=data.grad.numpy())
    finally:
        dq.cleanup_distributed()


@pytest.mark.parametrize('world,n', [(1, 4), (2, 4), (2, 6), (4, 6)])
def test_sharded_expectation_is_differentiable(world, n, tmp_path):
    """Expectation values and the gradient w.r.t. the encoded data of the sharded circuit = the reference's dense
    autograd result (fixture written by oracle/make_golden.py from the unmodified reference; the reference's own test
    compares in float32 with `torch.allclose`); X / Y observables, controlled and multi-target parametric gates, gates on
    the rank bits."""
    import subprocess
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), 'adjoint', str(r), str(world), str(port),
                               str(tmp_path), str(n)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = []
    for pr in procs:
        try:
            o, _ = pr.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    assert all(pr.returncode == 0 for pr in procs), '\n'.join(logs)
    res = np.load(os.path.join(tmp_path, 'adj.npz'))
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'dist_adjoint.npz'))
    assert np.abs(res['expectation'] - g[f'ref_test_n{n}/expectation']).max() < 1e-10
    # the reverse sweep un-computes with U^dagger; the reference's Hadamard is a float32-rounded constant (unitary to
    # 6e-8, gate.py:1069), which bounds the gradient at ~1e-7 here -- the floor of the reference's own adjoint too
    assert np.abs(res['grad'] - g[f'ref_test_n{n}/grad']).max() < 5e-7, (res['grad'], g[f'ref_test_n{n}/grad'])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world', [1, 2, 4])
def test_sharded_circuit_matches_dense_oracle(world, tmp_path):
    import subprocess

    import statevec_oracle as so

    import deepquantum_b200 as dq
    n = 7
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), str(r), str(world), str(port), str(n),
                               str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = []
    for pr in procs:
        try:
            o, _ = pr.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    assert all(pr.returncode == 0 for pr in procs), '\n'.join(logs)
    res = np.load(os.path.join(tmp_path, 'out.npz'))
    # dense reference: the same builder calls on the single-device circuit, lowered to oracle ops
    dense = _build(dq.QubitCircuit(n), n, extra=True)
    dense.to(torch.double)
    ops = [(op.update_matrix().detach().numpy(), op.wires, op.controls) for op in dense.operators]
    ref = so.run_circuit(ops, n)
    got = res['state']
    assert np.linalg.norm(got - ref) < 1e-12, np.linalg.norm(got - ref)
    if world > 1:
        assert 'swap' in list(res['steps'])      # the circuit does target global qubits
    # measure_dist: same uniforms (seeded on rank 0) -> the oracle's inverse-CDF counts and marginal probabilities
    import json

    import sampling_oracle as smp
    meas = json.loads(str(res['measure']))
    torch.manual_seed(5)
    u = torch.rand(300, dtype=torch.float64).numpy()
    for name, wires in (('all', None), ('sub', [n - 1, 0, 2]), ('glob', [0])):
        exp = smp.measure(ref, n, u, wires=wires, with_prob=True)
        got_m = meas[name]
        assert set(got_m) == set(exp), name
        for k, (c, p) in exp.items():
            assert got_m[k][0] == c and abs(got_m[k][1] - p) < 1e-12, (name, k, got_m[k], (c, p))


if __name__ == '__main__':
    if sys.argv[1] == 'adjoint':
        _worker_adjoint(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6]))
    else:
        _worker(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
