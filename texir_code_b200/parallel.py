"""Multi-GPU plumbing of the hot path: one process per GPU, the scene replicated on
every rank, the points of a call split into contiguous slices, no collective on the data path.

  * IrT generation (forward only): every rank computes its slice; one all_gather of [P/G,3] at the end.
  * Material stage (forward + backward): every rank scatters its pixels' gradients into its own copy of the
    texture-gradient buffers; ONE all-reduce(sum) over a flat view of all of them afterwards.

Works with any torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_points: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """contiguous slice [begin, end) of rank `rank`; slice sizes differ by at most `align`, every begin is
    a multiple of `align` (512 keeps the reference's batch granularity, tracer_o3d_irt.py:156)"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = (num_points + align - 1) // align
    base, extra = divmod(units, world)
    b = rank * base + min(rank, extra)
    e = b + base + (1 if rank < extra else 0)
    return min(b * align, num_points), min(e * align, num_points)


def shard_sizes(num_points: int, world: int, align: int = 1) -> List[int]:
    return [e - b for b, e in (shard_bounds(num_points, r, world, align) for r in range(world))]


def gather_points(local: torch.Tensor, num_points: int, align: int = 1, group=None) -> torch.Tensor:
    """all_gather of per-rank slices [n_r, C] (ragged allowed) into the full [num_points, C] tensor"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(num_points, world, align)
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False) -> int:
    """ONE all-reduce over the gradients of `params` (e.g. materials_a 2048^2x3 + materials_r 4096^2x1 =
    112 MiB fp32), through a single flat buffer.  Returns the number of bytes reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()


def sharded_irradiance(tracer, position: torch.Tensor, normal: torch.Tensor, shift: Optional[torch.Tensor] = None,
                       group=None, align: int = 512) -> torch.Tensor:
    """IrT over all ranks: `position`, `normal`, `shift` are the FULL [P,.] tensors (identical on every rank,
    so that the shift stream is the single-GPU one); returns the full [P,3] result on every rank."""
    P = position.shape[0]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    b, e = shard_bounds(P, rank, world, align)
    if shift is None:
        from .sample_util import draw_shift
        shift = draw_shift(P, position.device)
    local = tracer.irradiance(position[b:e], normal[b:e], shift[b:e])
    return gather_points(local, P, align, group)
