"""Multi-GPU plumbing: one process per GPU, the batch sharded by instance range.

Instances are independent and the circuit is shared and read-only, so the only
collective on this path is ONE broadcast of the compiled plan (opcode/coefficient stream) at
circuit-load time -- NCCL over NVLink when the process group is NCCL.  There is no traffic during
the solve; statuses are gathered by the host only if the caller asks for them.
"""
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous instance range [lo, hi) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_bytes(blob: Optional[bytes], src: int = 0, device: Optional[torch.device] = None) -> bytes:
    """Broadcast a byte string from `src` to every rank of the default process group."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    rank = dist.get_rank()
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src)
    if rank == src:
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def broadcast_circuit(ctx, circuit, input_witnesses: Sequence[int], src: int = 0):
    """Rank `src` holds a CompiledCircuit; every other rank receives the plan blob and loads it."""
    from .solver import CompiledCircuit
    rank = dist.get_rank()
    blob = broadcast_bytes(circuit.serialize() if rank == src else None, src=src)
    if rank == src:
        return circuit
    return CompiledCircuit.from_blob(ctx, blob, input_witnesses)


def gather_status_counts(n_solved: int, n_failed: int, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """Optional epilogue: sum of per-rank (solved, failed) counts."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([n_solved, n_failed], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return int(t[0].item()), int(t[1].item())
