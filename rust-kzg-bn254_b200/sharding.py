"""Multi-GPU sharding of the path (one process per GPU, torch.distributed for the plumbing).

Two natural shardings, neither needs a collective on the data path:
  * by blob        -- ranks take contiguous ranges of a blob batch; outputs are gathered on the host;
  * by point range -- one large MSM is cut into contiguous SRS ranges; every rank produces one affine
                      partial sum (64 B) and the partials are exchanged with ONE all_gather and added
                      locally (G1 addition is not an NCCL reduction operator).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous (first, count) of `total` items for `rank`; the remainder goes to the low ranks."""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_gather_bytes(payload: bytes, device: Optional[str] = None) -> List[bytes]:
    """all_gather of equal-length byte strings (NCCL needs CUDA tensors, gloo CPU tensors)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [payload]
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def reduce_g1_partials(pkg, partials: Sequence[bytes]):
    """Sum of affine partial results, each 65 bytes: x||y Montgomery words + identity flag."""
    acc = None
    for p in partials:
        pt = pkg.g1_from_abi(p[:64], p[64:65])[0]
        acc = pkg.g1_add(acc, pt)
    return acc


def msm_srs_point_range_sharded(pkg, engine, scalars_dev_ptr: int, first: int, count: int):
    """This rank's share of a point-range-sharded MSM followed by the gather-and-add exchange.
    `engine` holds SRS[first, first+count) as its local points 0..count; scalars are device resident."""
    import ctypes as C

    out = C.create_string_buffer(64)
    inf = C.c_uint8(0)
    engine.check(pkg.lib.kzgb_msm_srs_range_dev(engine.h, scalars_dev_ptr, 0, count, out, C.byref(inf)))
    partial = out.raw + bytes([inf.value])
    return reduce_g1_partials(pkg, all_gather_bytes(partial))
