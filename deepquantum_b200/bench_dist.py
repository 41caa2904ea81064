"""bench.py for N > 1 ranks (launched by torchrun, one rank per GPU, NCCL): the SAME circuit as the
single-GPU headline (30 qubits, depth 40 unless --nqubit/--depth say otherwise) with the high-order qubit
index sharded over the ranks -> strong scaling."""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist


def run(args):
    import deepquantum_b200 as dq
    from deepquantum_b200 import circuit as circ
    from deepquantum_b200 import workloads as wl
    import bench as B

    rank, world, local_rank = dq.setup_distributed('nccl')
    dev = torch.device('cuda', local_rank)
    n = args.nqubit or 30
    depth = args.depth or 40
    g = world.bit_length() - 1
    circ.PLAN_OPTIONS.update(chunk_bits=args.chunk_bits, fuse=not args.no_fuse)
    spec = wl.random_clifford_rx_spec(n, depth)
    cir = dq.DistributedQubitCircuit(n)
    angles = []
    for e in spec:
        if e['g'] == 'rx':
            cir.rx(e['w'][0], encode=True)
            angles.append(e['p'][0])
        else:
            wl.apply_spec(cir, [e])
    cir.observable([0], 'z')
    cir.observable([n // 2, n - 1], 'zz')
    cir.to(dev)
    data_host = torch.tensor(angles, dtype=torch.float32).pin_memory()
    data_dev = data_host.to(dev)
    ngates = cir._get_program().ngates

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        cir(data_dev)
    sync()
    cir._marks = []
    clocks = B.ClockSampler(local_rank)
    clocks.__enter__()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cir(data_dev)
    e1.record()
    sync()
    clocks.__exit__()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    seg_ms = sum(a.elapsed_time(b) for k, a, b in cir._marks if k == 'seg')
    swap_ms = sum(a.elapsed_time(b) for k, a, b in cir._marks if k == 'swap')
    parts = torch.tensor([seg_ms, swap_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(parts, op=dist.ReduceOp.MAX)
    cir._marks = None
    st = cir._sharded.stats()

    # end to end: pinned host angles -> cir(data) -> expectation (all-reduce) -> host
    def e2e_step():
        d = data_host.to(dev, non_blocking=True)
        cir(d)
        return cir.expectation().cpu()

    e2e_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    sync()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)

    if rank == 0:
        total_ms = float(ms[0])
        nl = n - g
        bytes_pass = 2 * (2**nl) * 8
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(B.ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        seg_s = float(parts[0]) * 1e-3
        achieved = st['passes'] * args.steps * bytes_pass / seg_s / 1e9 if seg_s > 0 else 0.0
        shard_bytes = (2**nl) * 8
        line = {
            'metric': B.METRIC, 'value': ngates * args.steps / (total_ms * 1e-3), 'unit': B.UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': f'{n}-qubit random Clifford+RX, depth {depth}, complex64, high-order index sharded '
                                   f'over {world} ranks ({nl} local qubits); same circuit as the 1-GPU headline',
                       'gates': ngates, 'local_passes': st['passes'], 'segments': st['segments'],
                       'block_transposes': st['swaps'], 'shard_bytes': shard_bytes,
                       'fused_pass_exchanges': getattr(cir._sharded, 'fused_exchanges', 0),
                       'exchange_path': ('last pass of each segment stores into the peers\' receive buffers over NVLink '
                                         '(symmetric memory), one 4-byte all-reduce per transpose'
                                         if getattr(cir._sharded, 'fused_exchanges', 0) else
                                         'NCCL grouped send/recv: ' + str(cir.state.__dict__.get('_peer_error', ''))),
                       'l2': 'every pass streams the whole shard; shard > L2 for n_local >= 25 complex64',
                       'ms_local_kernels': float(parts[0]) / args.steps, 'ms_exchange': float(parts[1]) / args.steps,
                       'exchange_bytes_per_rank_per_transpose': shard_bytes * (world - 1) // world},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': None, 'kernel': 'b200q_tile_kernel<float,12> (rank 0, local segments only)',
                         'peak_source': 'MEASURED_PEAKS.json (measured copy)' if peaks else 'fallback 6650'},
            'e2e': {'value': ngates * args.steps / float(e2e_s[0]), 'unit': B.UNIT,
                    'h2d_bytes_per_step': data_host.numel() * 4, 'd2h_bytes_per_step': int(res.numel() * res.element_size()),
                    'note': 'per rank: pinned host angles -> cir(data) -> expectation() (all-reduce) -> host'},
            'gpu_launches': args.steps * (st['passes'] + 1),
            'clocks': clocks.summary(),
        }
        print(json.dumps(line))
    dq.cleanup_distributed()
