"""torchrun worker of tests/test_gpu_parity.py::test_sharded_fock_two_gpus: the sharded Fock tensor path
(reference photonic/distributed.py:65-78, photonic/state.py:623-685) on one GPU per rank over NCCL, compared on
rank 0 with the HOST oracle (numpy restatement of evolve_state with qudit = cutoff, oracle/statevec_oracle.py) run
on the same Fock matrices, and with the single-GPU circuit.  Prints FOCK_SHARDED_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import deepquantum_b200 as dq  # noqa: E402
import statevec_oracle as so  # noqa: E402
from test_widen_fock_gloo import _build  # noqa: E402


def main():
    rank, world, local = dq.setup_distributed('nccl')
    torch.cuda.set_device(local)
    ok = True
    for cutoff, nl in ((2, 12), (4, 6)):
        n = nl + {2: world.bit_length() - 1, 4: (world.bit_length() - 1 + 1) // 2}[cutoff]
        if cutoff**(n - nl) != world:
            continue                                   # the rank count must be a power of the cutoff
        init = [(0.6, [1] + [0] * (n - 1)), (0.8, [0] * (n - 1) + [1])]
        cir = _build(dq.DistributedQumodeCircuit(n, init, cutoff=cutoff), n).to(f'cuda:{local}')
        st = cir()
        shards = [torch.empty_like(st.amps) for _ in range(world)]
        dist.all_gather(shards, st.amps.contiguous())
        if rank == 0:
            dense = _build(dq.QumodeCircuit(n, init, cutoff=cutoff, backend='fock', basis=False), n).to(f'cuda:{local}')
            one_gpu = dense().reshape(-1)
            got = torch.stack(shards).reshape(-1)
            mats = dense.build_matrices(torch.complex128, f'cuda:{local}')
            psi = dense.init_state.state.reshape(1, -1).cpu().numpy().astype(np.complex128)
            for op, m in zip(dense.operators, mats):
                psi = so.evolve_state(psi, m.cpu().numpy(), n, list(op.wires), cutoff)
            ref = psi.reshape(-1)
            e_or = np.linalg.norm(got.cpu().numpy() - ref) / np.linalg.norm(ref)
            e_1g = float((got - one_gpu).norm())
            print(f'fock cutoff={cutoff} n={n} world={world} vs_oracle={e_or:.2e} vs_one_gpu={e_1g:.2e}')
            ok = ok and e_or < 2e-6 and e_1g < 1e-5
    if rank == 0:
        print('FOCK_SHARDED_OK' if ok else 'FOCK_SHARDED_MISMATCH')
    dq.cleanup_distributed()


if __name__ == '__main__':
    main()
