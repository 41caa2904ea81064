"""Multi-GPU check of the sharded Fock tensor path (NOT part of pytest: the path has only run on gloo ranks so far).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_fock_gpu_check.py      # cutoff 2
    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/dist_fock_gpu_check.py
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/dist_fock_gpu_check.py

Every rank runs the sharded circuit on its GPU (NCCL all-to-all swaps + the qudit kernel); rank 0 also runs the same
circuit on one GPU and compares the gathered shards.  Prints FOCK_SHARDED_OK or the error norm."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import deepquantum_b200 as dq  # noqa: E402
from test_widen_fock_gloo import _build  # noqa: E402


def main():
    rank, world, local = dq.setup_distributed('nccl')
    torch.cuda.set_device(local)
    cutoff = 2
    n = 12 + (world.bit_length() - 1)                 # 2^12 amplitudes per rank
    init = [(0.6, [1] + [0] * (n - 1)), (0.8, [0] * (n - 1) + [1])]
    cir = _build(dq.DistributedQumodeCircuit(n, init, cutoff=cutoff), n).to(f'cuda:{local}')
    st = cir()
    shards = [torch.empty_like(st.amps) for _ in range(world)]
    dist.all_gather(shards, st.amps.contiguous())
    if rank == 0:
        dense = _build(dq.QumodeCircuit(n, init, cutoff=cutoff, backend='fock', basis=False), n).to(f'cuda:{local}')
        ref = dense().reshape(-1)
        got = torch.stack(shards).reshape(-1)
        err = (got - ref).norm().item()
        print('FOCK_SHARDED_OK' if err < 1e-5 else f'FOCK_SHARDED_MISMATCH {err:.3e}', f'world={world} n={n} err={err:.2e}')
    dq.cleanup_distributed()


if __name__ == '__main__':
    main()
