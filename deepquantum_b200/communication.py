"""torch.distributed plumbing with the reference's function names (communication.py:9-91).

One process per GPU, launched by torchrun; NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def setup_distributed(backend: str = 'nccl', port: str = '29500') -> tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment; returns (rank, world_size, local_rank)."""
    try:
        rank = int(os.environ['RANK'])
        world_size = int(os.environ['WORLD_SIZE'])
        local_rank = int(os.environ['LOCAL_RANK'])
    except KeyError:
        rank, world_size, local_rank = 0, 1, 0
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', port)
    if backend == 'nccl':
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, world_size=world_size, rank=rank,
                                device_id=torch.device('cuda', local_rank))
    else:
        dist.init_process_group(backend, world_size=world_size, rank=rank)
    return rank, world_size, local_rank


def cleanup_distributed() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()


def comm_get_rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def comm_get_world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def comm_exchange_arrays(send_data: torch.Tensor, recv_data: torch.Tensor, pair_rank: int | None) -> None:
    """Pairwise exchange with `pair_rank` (reference communication.py:58-91).  EVERY rank of the group must
    call this (idle ranks with `pair_rank=None`): unlike the reference's both-global swap branch
    (distributed.py:144-147) no rank ever skips a collective."""
    world = comm_get_world_size()
    rank = comm_get_rank()
    in_splits = [0] * world
    out_splits = [0] * world
    if pair_rank is not None:
        in_splits[pair_rank] = send_data.numel()
        out_splits[pair_rank] = recv_data.numel()
    dist.all_to_all_single(recv_data if pair_rank is not None else recv_data[:0],
                           send_data if pair_rank is not None else send_data[:0], out_splits, in_splits)


def block_transpose(amps: torch.Tensor, buffer: torch.Tensor) -> None:
    """buffer <- all-to-all of the W equal chunks of `amps`: chunk c of rank r lands as chunk r of rank c.
    This swaps the log2(W) rank bits with the top log2(W) local index bits in ONE collective (each rank sends
    (W-1)/W of its shard once), instead of one half-shard exchange per global qubit."""
    world = comm_get_world_size()
    if world == 1:
        buffer.copy_(amps)
        return
    if dist.get_backend() != 'nccl':
        dist.all_to_all_single(buffer, amps)
        return
    # NCCL: grouped point-to-point for the W-1 peers, a plain device copy for the rank's own chunk (measured on 2
    # B200s over NVLink: 506-548 GB/s sent per rank, against 412-428 GB/s for all_to_all_single, which also moves
    # the local chunk through the collective; tools/exchange_bench.py)
    rank = comm_get_rank()
    chunk = amps.numel() // world
    ops = []
    for p in range(world):
        if p == rank:
            continue
        ops.append(dist.P2POp(dist.isend, amps[p * chunk:(p + 1) * chunk], p))
        ops.append(dist.P2POp(dist.irecv, buffer[p * chunk:(p + 1) * chunk], p))
    reqs = dist.batch_isend_irecv(ops)
    buffer[rank * chunk:(rank + 1) * chunk].copy_(amps[rank * chunk:(rank + 1) * chunk])
    for r in reqs:
        r.wait()
