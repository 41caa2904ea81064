"""torchrun micro-benchmark of the block-transpose exchange variants (NCCL over NVLink)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepquantum_b200 as dq  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, lr = dq.setup_distributed('nccl')
    for nl in (27, 30):
        amps = torch.randn(2**nl, dtype=torch.complex64, device='cuda')
        buf = torch.empty_like(amps)
        chunk = amps.numel() // world

        def a2a():
            dist.all_to_all_single(buf, amps)

        def p2p():
            ops = []
            for p in range(world):
                if p == rank:
                    continue
                ops.append(dist.P2POp(dist.isend, amps[p * chunk:(p + 1) * chunk], p))
                ops.append(dist.P2POp(dist.irecv, buf[p * chunk:(p + 1) * chunk], p))
            reqs = dist.batch_isend_irecv(ops)
            buf[rank * chunk:(rank + 1) * chunk].copy_(amps[rank * chunk:(rank + 1) * chunk])
            for r in reqs:
                r.wait()

        sent = (world - 1) * chunk * 8
        for name, fn in (('all_to_all_single', a2a), ('batch_isend_irecv', p2p)):
            ms = timed(fn)
            if rank == 0:
                print(json.dumps({'variant': name, 'world': world, 'n_local': nl, 'ms': ms,
                                  'GBps_sent_per_rank': sent / ms / 1e6}))
        del amps, buf
        torch.cuda.empty_cache()
    dq.cleanup_distributed()


if __name__ == '__main__':
    main()
