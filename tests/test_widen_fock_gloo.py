"""Sharded Fock tensor path on CPU: `world_size = cutoff^g` gloo ranks run the product's
own host logic (`deepquantum_b200.photonic_distributed`: mode map, look-ahead eviction, all-to-all swaps of a rank
digit with a local axis, layout restoration); only the LOCAL gate application is executed by the test-only CPU
emulator of the qudit kernel's geometry instead of the GPU.  Compared against the dense oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


class EmuQuditExecutor:
    """TEST-ONLY executor: steps the qudit kernel's index geometry on the CPU."""

    def apply(self, amps, nmode_local, cutoff, matrix, wires):
        from test_fock import _emu_qudit
        out = _emu_qudit(amps.numpy().reshape(-1), nmode_local, cutoff, matrix.numpy(), list(wires))
        amps.reshape(-1).copy_(torch.from_numpy(out))


def _build(cir, n):
    """Gates on every mode, two-mode gates across the global / local boundary and between two global modes."""
    g = torch.Generator().manual_seed(3)
    r = lambda s=1.0: float(torch.rand(1, generator=g) * s)   # noqa: E731
    for w in range(n):
        cir.s(w, r(0.4), r(6))
    cir.d(0, r(0.3), r(6))
    cir.bs([0, 1], [r(6), r(6)])
    cir.bs([n - 1, 0], [r(6), r(6)])
    cir.mzi([1, 2], [r(6), r(6)])
    cir.ps(0, r(6))
    cir.bs_rx([2, n - 1], r(6))
    cir.ck([0, n - 2], r(1))
    cir.bs([1, 0], [r(6), r(6)])
    cir.k(1, r(1))
    cir.d(n - 1, r(0.3), r(6))
    cir.bs_h([0, 2], r(6))
    return cir


def _worker(rank, world, port, n, cutoff, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import deepquantum_b200 as dq
    from deepquantum_b200.photonic_distributed import DistributedQumodeCircuit
    dq.setup_distributed('gloo')
    try:
        init = [(0.6, [1] + [0] * (n - 1)), (0.8, [0] * (n - 1) + [1])]
        cir = _build(DistributedQumodeCircuit(n, init, cutoff=cutoff), n)
        cir.to(torch.double)
        cir._executor = EmuQuditExecutor()
        st = cir()
        assert tuple(st.amps.shape) == (cutoff,) * st.nmode_local
        shards = [torch.empty_like(st.amps) for _ in range(world)]
        if world > 1:
            dist.all_gather(shards, st.amps.contiguous())
        else:
            shards = [st.amps]
        if rank == 0:
            np.save(os.path.join(outdir, 'state.npy'), torch.stack(shards).reshape(-1).numpy())
    finally:
        dq.cleanup_distributed()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world,cutoff,n', [(1, 3, 4), (2, 2, 5), (4, 2, 5), (3, 3, 4), (9, 3, 4)])
def test_sharded_fock_circuit_matches_dense_oracle(world, cutoff, n, tmp_path):
    import statevec_oracle as so

    import deepquantum_b200 as dq
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), str(r), str(world), str(port), str(n),
                               str(cutoff), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
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
    got = np.load(os.path.join(tmp_path, 'state.npy'))
    dense = _build(dq.QumodeCircuit(n, 'vac', cutoff=cutoff, backend='fock', basis=False), n)
    dense.to(torch.double)
    psi = np.zeros([cutoff] * n, dtype=np.complex128)
    psi[(1,) + (0,) * (n - 1)] = 0.6
    psi[(0,) * (n - 1) + (1,)] = 0.8
    psi = psi.reshape(-1)
    for op, m in zip(dense.operators, dense.build_matrices(torch.complex128, 'cpu')):
        psi = so.evolve_state(psi, m.numpy(), n, op.wires, cutoff)
    assert np.linalg.norm(got - psi.reshape(-1)) < 1e-12, np.linalg.norm(got - psi.reshape(-1))


if __name__ == '__main__':
    _worker(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
