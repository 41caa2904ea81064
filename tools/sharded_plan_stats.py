"""Host-only: local passes / segments / transposes of the sharded schedule of rank 0 for a world size (no GPU, no
process group): python tools/sharded_plan_stats.py <nqubit> <depth> <world>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepquantum_b200 as dq  # noqa: E402
from deepquantum_b200 import engine, workloads as wl  # noqa: E402
from deepquantum_b200.distributed import ShardedProgram  # noqa: E402


class PlanOnly:
    def __init__(self, exchange_bits=3):
        self.exchange_bits = exchange_bits

    def make_plan(self, nlocal, dtype, structs, exchange=False):
        return engine.FusedPlan(nlocal, dtype, structs, coalesce_bits=self.exchange_bits if exchange else None)

    def run_plan(self, plan, amps, mats):
        pass

    def run_plan_exchange(self, plan, amps, mats, peers, rank, perm=None):
        pass


class FakeState:
    def __init__(self, peer):
        self.amps = torch.zeros(1, dtype=torch.complex64)
        self.buffer = torch.zeros(1, dtype=torch.complex64)
        self.peer = peer

    def enable_peer_exchange(self):
        return self.peer

    def peer_buffer_ptrs(self):
        return [0]


def main():
    n, depth, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    spec = wl.random_clifford_rx_spec(n, depth)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    low = cir._get_program().low
    mats = low.build_matrices(torch.complex64, 'cpu').detach()
    import torch.distributed as dist
    dist.all_reduce = lambda *a, **k: None        # plan-only: no process group
    for mode in ('pswap', 'perm'):
        sp = ShardedProgram(low, n, world, 0, mode)
        sp.run(FakeState(mode == 'perm'), mats, PlanOnly())
        passes = sum(p.n_passes for p in sp.plans.values())
        per_seg = [p.n_passes for p in sp.plans.values()]
        print(f'n={n} depth={depth} world={world} mode={mode}: passes {passes} segments {sp.n_segments} '
              f'exchanges {sp.n_swaps} per segment {per_seg}')
    one = cir._get_program().plan(torch.complex64)
    print(f'  single device: {one.n_passes} passes')


if __name__ == '__main__':
    main()
