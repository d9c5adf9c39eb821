"""Chain sharding across the GPUs of one box: one process per GPU, no data-path collective.

Chains are independent, so ``pm.sample`` scales by giving every rank a contiguous block of chains
(``chain_offset`` keys the Philox streams by GLOBAL chain index, so a sharded run draws exactly the
numbers a single-GPU run of all chains would).  The only collectives are on the diagnostics /
trace-assembly side (``diagnostics.gather_chains``: all-gather over NCCL on GPUs, gloo in the CPU
tests).  The reference has no multi-device code; this is new.
"""
import os
from typing import Tuple


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_chains(total_chains: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of chains of ``rank``: (local count, global index of its first chain)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total_chains), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return count, offset
