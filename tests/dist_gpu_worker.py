"""torchrun worker: DistributedQubitCircuit over NCCL.  The gathered state is compared with the CPU ORACLE
(oracle/torch_port.run_ops, complex128, rank 0 host cores) on the C2 generator at 22 qubits depth 20 -- shards of
2^21 .. 2^19 amplitudes, every local pass with several groups of non-tile bits and run by the specialised
kernels -- in complex128 and complex64, and with the single-GPU engine on the n = 12 all-gate-families circuit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)

import deepquantum_b200 as dq  # noqa: E402
from deepquantum_b200 import workloads as wl  # noqa: E402
from test_distributed_gloo import _build  # noqa: E402


def main():
    rank, world, local_rank = dq.setup_distributed('nccl')
    dev = torch.device('cuda', local_rank)
    ok = True
    # (1) every sharded case at n = 12 against the single-GPU engine
    for n, builder in ((12, lambda c, n=12: _build(c, n)),
                       (22, lambda c: wl.apply_spec(c, wl.random_clifford_rx_spec(22, 12)))):
        cir = dq.DistributedQubitCircuit(n)
        builder(cir)
        cir.observable([0], 'z')
        cir.observable([1, n - 1], 'zz')
        cir.to(dev, torch.double)
        st = cir()
        exp = cir.expectation()
        torch.manual_seed(3)
        meas = cir.measure(shots=400, with_prob=True, wires=[0, 1, n - 1])   # wires 0.. are global (rank bits)
        shards = [torch.empty_like(st.amps) for _ in range(world)]
        dist.all_gather(shards, st.amps.contiguous())
        full = torch.cat(shards)
        if rank == 0:
            dense = dq.QubitCircuit(n)
            builder(dense)
            dense.observable([0], 'z')
            dense.observable([1, n - 1], 'zz')
            dense.to(dev, torch.double)
            with torch.no_grad():
                for a, b in zip(dense.parameters(), cir.parameters()):
                    a.copy_(b)
                for a, b in zip(dense.buffers(), cir.buffers()):
                    if a.shape == b.shape and a.dtype == b.dtype:
                        a.copy_(b)
            ref = dense().reshape(-1)
            err = float((full - ref).norm())
            e_err = float((dense.expectation().reshape(-1) - exp.reshape(-1)).abs().max())
            print(f'n={n} world={world} |sharded - dense| = {err:.3e}  expectation diff {e_err:.3e} '
                  f'schedule {cir._sharded.stats()}')
            ok = ok and err < 1e-10 and e_err < 1e-10
            # measure_dist against the oracle's inverse CDF of the gathered state with the same uniforms
            import sampling_oracle as smp
            torch.manual_seed(3)
            u = torch.rand(400, dtype=torch.float64).numpy()
            want = smp.measure(full.cpu().numpy(), n, u, wires=[0, 1, n - 1], with_prob=True)
            m_ok = set(want) == set(meas) and all(
                abs(meas[k][0] - c) <= 1 and abs(meas[k][1] - p) < 1e-10 for k, (c, p) in want.items())
            print(f'n={n} measure_dist keys {len(meas)} ok={m_ok}')
            ok = ok and m_ok
    # (2) the C2 generator sharded over the ranks against the oracle (parity at n = 20-24 with W ranks)
    import gates_np
    import torch_port
    n, depth = 22, 20
    spec = wl.random_clifford_rx_spec(n, depth)
    ref = None
    if rank == 0:
        out, done, _ = torch_port.run_ops(gates_np.lower_spec(spec, n), n, dtype=torch.complex128)
        ref = out.numpy()
    for rdtype, tol in ((torch.double, 1e-10), (torch.float, 2e-6)):
        cir = dq.DistributedQubitCircuit(n)
        wl.apply_spec(cir, spec)
        cir.to(dev, rdtype)
        st = cir()
        shards = [torch.empty_like(st.amps) for _ in range(world)]
        dist.all_gather(shards, st.amps.contiguous())
        if rank == 0:
            full = torch.cat(shards).cpu().numpy().astype(np.complex128)
            err = float(np.linalg.norm(full - ref) / np.linalg.norm(ref))
            print(f'n={n} depth={depth} world={world} {rdtype}: rel-L2 vs oracle {err:.3e} schedule {cir._sharded.stats()}')
            ok = ok and err < tol
    # (3) differentiable expectation of the sharded circuit (reference circuit.py:1706-1738, adjoint.py:19-83):
    #     the reference's own test circuit against its dense-autograd fixture, and a 21-qubit circuit whose
    #     exchanges run over the NVLink peer mappings against the single-GPU reverse sweep
    from test_distributed_gloo import _free_port  # noqa: F401  (module import keeps sys.path consistent)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'dist_adjoint.npz'))
    for n in (4, 6):
        if 2**n < 4 * world * world:
            continue
        data = torch.tensor(g[f'ref_test_n{n}/data'], dtype=torch.float64, device=dev, requires_grad=True)
        cir = dq.DistributedQubitCircuit(n, reupload=True)
        cir.rxlayer(encode=True); cir.rylayer(encode=True); cir.rzlayer(encode=True); cir.u3layer(encode=True)
        cir.hlayer(); cir.cnot_ring(); cir.toffoli(0, 1, 2); cir.fredkin(2, 1, 0); cir.swap([2, 3])
        cir.rx(0, controls=[1, 2, 3], encode=True); cir.ry(1, controls=[0, 2, 3], encode=True)
        cir.rz(2, controls=[0, 1, 3], encode=True); cir.rxx([0, 1], controls=[2, 3], encode=True)
        cir.ryy([1, 2], controls=[0, 3], encode=True); cir.rzz([2, 3], controls=[0, 1], encode=True)
        cir.rxy([3, 0], controls=[1, 2], encode=True)
        cir.observable(0); cir.observable(1, 'x'); cir.observable([2, 3], 'xy')
        cir.to(dev, torch.double)
        cir(data=data)
        exp = cir.expectation()
        exp.sum().backward()
        if rank == 0:
            e_err = float(np.abs(exp.detach().cpu().numpy() - g[f'ref_test_n{n}/expectation']).max())
            g_err = float(np.abs(data.grad.cpu().numpy() - g[f'ref_test_n{n}/grad']).max())
            print(f'n={n} world={world} sharded expectation vs reference {e_err:.2e}, gradient {g_err:.2e}')
            ok = ok and e_err < 1e-10 and g_err < 5e-7
    n = 21
    gen = torch.Generator().manual_seed(4)
    angles = (torch.rand(3 * n, generator=gen, dtype=torch.float64) * 6).to(dev)

    def build(c):
        c.rxlayer(encode=True)
        c.cnot_ring()
        c.rylayer(encode=True)
        for q in range(0, n - 1, 2):
            c.cnot(q + 1, q)
        c.rzlayer(encode=True)
        c.rxlayer()
        c.cnot_ring(step=3)
        c.observable([0, n - 1], 'zz')
        c.observable(1, 'x')
        c.observable([2, 10], 'yz')
        return c

    grads = []
    for cls in (dq.DistributedQubitCircuit, dq.QubitCircuit):
        torch.manual_seed(1)     # same initial parameters of the trainable rxlayer
        cir = build(cls(n))
        cir.to(dev, torch.double)
        d = angles.clone().requires_grad_(True)
        cir(data=d)
        e = cir.expectation().reshape(-1)
        (e * torch.tensor([1.0, -0.5, 2.0], dtype=torch.float64, device=dev)).sum().backward()
        grads.append((e.detach(), d.grad.clone(), torch.stack([p.grad.reshape(()) for p in cir.parameters()])))
    if rank == 0:
        e_err = float((grads[0][0] - grads[1][0]).abs().max())
        d_err = float((grads[0][1] - grads[1][1]).abs().max())
        p_err = float((grads[0][2] - grads[1][2]).abs().max())
        print(f'n={n} world={world} sharded vs single-GPU: expectation {e_err:.2e}, d/d data {d_err:.2e}, '
              f'd/d parameters {p_err:.2e}')
        ok = ok and e_err < 1e-10 and d_err < 1e-8 and p_err < 1e-8
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dq.cleanup_distributed()
    if rank == 0 and ok:
        print('SHARDED_OK')
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()
